"""CPU oracle of the body-distillation step (autograd on the functional restatement)  --  TEST INFRASTRUCTURE.
Follows siren_morpher_protocols_03.py:102-157,178-214 and siren_morpher_03_trainer.py:32-50 of the reference (body) and
siren_face_morpher_protocols_00.py:48-105, siren_face_morpher_00_trainer.py:112-186 (face).  Pinned against the reference's
own run_training_iteration: oracle/make_golden_distill.py -> tests/golden/distill_lambda00.npz, checked by
tests/test_oracle_pinned.py (fixture everywhere, live reference in the build container)."""
from typing import Dict, List, Sequence, Tuple

import torch
from torch import Tensor

from oracle import tha4_oracle as O


def body_losses_and_grads(student_sd: Dict[str, Tensor], image: Tensor, pose: Tensor, t_posed: Tensor, t_warped: Tensor,
                          t_grid: Tensor, weights: Sequence[float]) -> Tuple[List[float], Tensor]:
    """Returns ([mean |.| of the four terms], flat gradient in state_dict order)."""
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in student_sd.items()}
    outs = O.siren_morpher_03(sd, image, pose)       # blended, alpha, colour, warped, grid_change
    terms = [(outs[0] - t_posed).abs().mean(), (outs[3] - t_warped).abs().mean(), (outs[4] - t_grid).abs().mean(),
             (outs[2] - t_posed).abs().mean()]
    loss = sum(w * t for w, t in zip(weights, terms))
    loss.backward()
    flat = torch.cat([sd[k].grad.reshape(-1) for k in student_sd])
    return [float(t.detach()) for t in terms], flat


def face_losses_and_grads(student_sd: Dict[str, Tensor], pose: Tensor, target: Tensor, mask: Tensor,
                          weights: Sequence[float] = (1.0, 20.0)) -> Tuple[List[float], Tensor]:
    """Face student (siren_face_morpher_protocols_00.py:72-105; siren_face_morpher_00_trainer.py:112-114,168-186;
    shion/base/loss/l1_loss.py:9-24,40-58): returns ([mean |t-o|, mean |(t-o) m|], flat gradient in state_dict order)."""
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in student_sd.items()}
    out = O.siren_face_morpher(sd, pose[:, 0:39])
    terms = [(target - out).abs().mean(), ((target - out) * mask).abs().mean()]
    loss = sum(w * t for w, t in zip(weights, terms))
    loss.backward()
    flat = torch.cat([sd[k].grad.reshape(-1) for k in student_sd])
    return [float(t.detach()) for t in terms], flat


def adam_reference(params: Tensor, grads: Sequence[Tensor], lr: float, betas=(0.9, 0.999), eps=1e-8) -> Tensor:
    p = params.detach().clone().requires_grad_(True)
    opt = torch.optim.Adam([p], lr=lr, betas=betas, eps=eps)
    for g in grads:
        opt.zero_grad()
        p.grad = g.clone()
        opt.step()
    return p.detach()
