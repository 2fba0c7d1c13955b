"""Developer script: where the host-side time of one e2e frame goes (B=1 teacher)."""
import os, sys, time, torch
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT)
sys.path.insert(0, os.path.join(_ROOT, 'tests'))
from tha4_b200 import synthetic
from tha4_b200.poser.modes import mode_07
dev = torch.device('cuda:0')
sds = synthetic.teacher_state_dicts(0)
poser = mode_07.create_poser(dev, state_dicts=sds)
img = synthetic.synthetic_image(0, 1)
poses = synthetic.random_poses(64, seed=3)
img_host = img.pin_memory(); poses_host = poses.pin_memory()
out_host = torch.empty(1, 4, 512, 512).pin_memory()
img_in = torch.empty(1, 4, 512, 512, device=dev); pose_in = torch.empty(1, 45, device=dev)
def sync(): torch.cuda.current_stream().synchronize()
acc = {k: 0.0 for k in ('h2d', 'call', 'gpu_wait', 'd2h')}
N = 40
for i in range(N + 10):
    t0 = time.perf_counter()
    img_in.copy_(img_host, non_blocking=True); pose_in.copy_(poses_host[i % 64:i % 64 + 1], non_blocking=True)
    sync(); t1 = time.perf_counter()
    out = poser.pose(img_in, pose_in)
    t2 = time.perf_counter()
    sync(); t3 = time.perf_counter()
    out_host.copy_(out, non_blocking=True); sync()
    t4 = time.perf_counter()
    if i >= 10:
        acc['h2d'] += t1 - t0; acc['call'] += t2 - t1; acc['gpu_wait'] += t3 - t2; acc['d2h'] += t4 - t3
print({k: round(v / N * 1e3, 3) for k, v in acc.items()}, 'ms per frame; total', round(sum(acc.values()) / N * 1e3, 3))
