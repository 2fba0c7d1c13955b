"""N>1 host logic on CPU: world_size-2 gloo process group (no GPU): shard arithmetic and the gather of sharded frames.
The poser is a deterministic CPU stub implementing the Poser protocol -- this tests the sharding plumbing, not kernels."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tha4_b200.parallel import ShardedPoseSweep, shard_range
from tha4_b200.poser.poser import Poser


def test_shard_range_covers_everything_once():
    for total in (0, 1, 7, 8, 511, 512):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


class _StubPoser(Poser):
    def get_image_size(self): return 8
    def get_output_length(self): return 1
    def get_pose_parameter_groups(self): return []
    def get_num_parameters(self): return 45
    def to(self, device): return self
    def get_posing_outputs(self, image, pose): return [self.pose(image, pose)]
    def pose(self, image, pose, output_index=0):
        return image * pose.sum(dim=1).view(-1, 1, 1, 1)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, num_poses, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        image = torch.rand(4, 8, 8, generator=g)
        poses = torch.rand(num_poses, 45, generator=g)
        sweep = ShardedPoseSweep(_StubPoser(), chunk=3)
        local = sweep.pose_local(image, poses)
        b, e = sweep.local_range(num_poses)
        assert local.shape[0] == e - b
        full = sweep.pose_all(image, poses)
        ref = _StubPoser().pose(image.unsqueeze(0).expand(num_poses, -1, -1, -1), poses)
        out[rank] = bool(torch.equal(full, ref)) and bool(torch.equal(local, ref[b:e]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('num_poses', [7, 8])
def test_sharded_sweep_two_ranks_gloo(num_poses):
    world = 2
    port = _free_port()
    ctx = mp.get_context('spawn')
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_poses, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(out.get(r) for r in range(world))
