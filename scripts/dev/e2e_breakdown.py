"""Developer probe: where the end-to-end B=1 frame time goes on the host side (perf_counter with a sync after every
segment, so the segments do not overlap -- the sum is an upper bound of the pipelined step)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))
import torch
import bench
from tha4_b200 import synthetic
from tha4_b200.poser.modes import mode_07

dev = torch.device('cuda:0')
tsds, _ = bench.load_state_dicts('mode_07')
poser = mode_07.create_poser(dev, state_dicts=tsds)
poser.get_modules()
ctx = poser.get_context()
image = bench.load_image()
img_host = image.unsqueeze(0).contiguous().pin_memory()
poses_host = synthetic.random_poses(64, seed=5).pin_memory()
out_host = torch.empty((1, 512, 512, 4), dtype=torch.uint8).pin_memory()
img_in = torch.empty((1, 4, 512, 512), device=dev)
pose_in = torch.empty((1, 45), device=dev)
sync = torch.cuda.synchronize
acc = {}


def seg(name, t0):
    sync()
    t1 = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t1 - t0)
    return t1


with torch.no_grad():
    for it in range(40):
        if it == 10:
            acc.clear()
        t = time.perf_counter()
        img_in.copy_(img_host, non_blocking=True)
        pose_in.copy_(poses_host[it:it + 1], non_blocking=True)
        t = seg('h2d', t)
        outs = poser.get_posing_outputs(img_in, pose_in)
        t = seg('get_posing_outputs (incl. cache compare)', t)
        frame = ctx.frame_to_srgb8(outs[0], None, False)
        t = seg('frame_to_srgb8', t)
        out_host.copy_(frame, non_blocking=True)
        t = seg('d2h', t)
        del outs, frame
    n = 30
    for k, v in acc.items():
        print('%-45s %7.3f ms' % (k, v / n * 1e3))
    print('sum %.3f ms' % (sum(acc.values()) / n * 1e3))
    # the same with only the host part of get_posing_outputs timed (no sync inside): how long the call takes to RETURN
    t_ret = 0.0
    for it in range(30):
        sync()
        t0 = time.perf_counter()
        outs = poser.get_posing_outputs(img_in, pose_in)
        t_ret += time.perf_counter() - t0
        del outs
    print('get_posing_outputs returns after %.3f ms (host side incl. the cache-compare sync)' % (t_ret / 30 * 1e3))
    poser.protocol.trust_image_identity = True
    t_ret = 0.0
    for it in range(30):
        sync()
        t0 = time.perf_counter()
        outs = poser.get_posing_outputs(img_in, pose_in)
        t_ret += time.perf_counter() - t0
        del outs
    print('   with trust_image_identity: %.3f ms' % (t_ret / 30 * 1e3))
    print('graphs', ctx.counter('graph_replays'), ctx.counter('graph_captures'))
