"""Developer probe: phase stamps of every halo conv launch of one real B=1 teacher frame (THA4_HALO_DEBUG=2)."""
import os, sys
os.environ['THA4_HALO_DEBUG'] = '2'
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))
import torch
import bench
from tha4_b200 import synthetic
from tha4_b200.poser.modes import mode_07

dev = torch.device('cuda:0')
tsds, _ = bench.load_state_dicts('mode_07')
poser = mode_07.create_poser(dev, state_dicts=tsds)
ctx = poser.get_context()
ctx.set_option('cuda_graphs', 0)
image = bench.load_image().unsqueeze(0).to(dev)
poses = synthetic.random_poses(4, seed=5).to(dev)
with torch.no_grad():
    for i in range(3):
        print('=== frame', i, file=sys.stderr, flush=True)
        poser.get_posing_outputs(image, poses[i:i + 1])
        torch.cuda.synchronize()
