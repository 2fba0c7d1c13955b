"""Host logic of the distillation step for N>1 on CPU (gloo, world size 2): the flat-gradient all-reduce + 1/world scaling
reproduces DDP's gradient averaging, i.e. every rank ends up with the weights a single process would get from the mean
gradient.  The CUDA context is replaced by a CPU stub (this tests the plumbing in tha4_b200/distill.py, not kernels)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import distill_oracle


class _StubCtx:
    device = torch.device('cpu')

    def siren_morpher_train_step(self, image, pose, t0, t2, t3, weights, params, grads, want_losses=True):
        grads.copy_(torch.sin(params * 3.0) * (1.0 + dist.get_rank()))       # rank-dependent "gradient"
        return [0.0, 0.0, 0.0, 0.0]

    def siren_face_morpher_train_step(self, pose, target, mask, weights, params, grads, want_losses=True):
        assert target.shape[1:] == (4, 128, 128) and mask.shape == target.shape
        grads.copy_(torch.cos(params * 2.0) * (1.0 + dist.get_rank()))
        return [0.0, 0.0]

    def adam_step(self, params, grads, m, v, lr, step, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
        g = grads * grad_scale
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        denom = v.sqrt() / (1 - betas[1] ** step) ** 0.5 + eps
        params.addcdiv_(m, denom, value=-lr / (1 - betas[0] ** step))


class _StubTeacher:
    def __init__(self): self.ctx = _StubCtx()
    def get_context(self): return self.ctx
    def get_modules(self): return {}
    def get_posing_outputs(self, image, pose): return [image] * 33


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from tha4_b200 import distill
        from tha4_b200.poser.modes import mode_14
        student = mode_14.load_body_morpher(None)
        torch.manual_seed(0)
        for p in student.parameters():
            p.data.normal_(0, 0.1)
        d = distill.BodyMorpherDistiller.__new__(distill.BodyMorpherDistiller)
        d.teacher, d.student, d.ctx = _StubTeacher(), student, _StubCtx()
        d.flat = distill.flatten_parameters(student)
        d.grad, d.exp_avg, d.exp_avg_sq = (torch.zeros_like(d.flat) for _ in range(3))
        d.betas, d.eps, d.step_count, d.group, d.world = (0.9, 0.999), 1e-8, 0, None, world
        p0 = d.flat.clone()
        d.train_step(torch.zeros(1, 4, 8, 8), torch.zeros(1, 45), [1, 1, 1, 1], lr=1e-3, want_losses=False)
        mean_grad = torch.sin(p0 * 3.0) * (sum(1.0 + r for r in range(world)) / world)
        ref = distill_oracle.adam_reference(p0, [mean_grad], 1e-3)
        out[rank] = float((d.flat - ref).abs().max())
        # parameters stayed views of the flat buffer
        assert torch.equal(torch.cat([p.data.reshape(-1) for p in student.parameters()]), d.flat)
        # face student: same plumbing (teacher crop -> step -> all-reduce -> Adam)
        face = mode_14.load_face_morpher(None)
        for p in face.parameters():
            p.data.normal_(0, 0.1)
        f = distill.FaceMorpherDistiller.__new__(distill.FaceMorpherDistiller)
        f.teacher, f.student, f.ctx = _StubTeacher(), face, _StubCtx()
        f.flat = distill.flatten_parameters(face)
        f.grad, f.exp_avg, f.exp_avg_sq = (torch.zeros_like(f.flat) for _ in range(3))
        f.betas, f.eps, f.step_count, f.group, f.world = (0.9, 0.999), 1e-8, 0, None, world
        q0 = f.flat.clone()
        f.train_step(torch.zeros(1, 4, 192, 192), torch.zeros(1, 45), torch.ones(1, 4, 128, 128), lr=1e-3, want_losses=False)
        mean_grad = torch.cos(q0 * 2.0) * (sum(1.0 + r for r in range(world)) / world)
        ref = distill_oracle.adam_reference(q0, [mean_grad], 1e-3)
        out[rank] = max(out[rank], float((f.flat - ref).abs().max()))
    finally:
        dist.destroy_process_group()


def test_distill_allreduce_matches_mean_gradient_adam():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert len(out) == world and max(out.values()) < 1e-6, dict(out)
