"""Distillation inner loop of the body student on the CUDA path -- replaces, as a unit, what the reference does in
`SirenMorpherTrainingProtocol03.run_training_iteration` (src/tha4/nn/siren/morpher/siren_morpher_protocols_03.py:178-214):

    teacher forward under no_grad (mode_07, :102-108)  ->  student forward (:125-135)  ->  SumLoss of four
    time-weighted L1 terms (siren_morpher_03_trainer.py:32-50,237-247)  ->  backward  ->  DDP gradient averaging
    (shion/core/training/distrib/distributed_training_states.py:184-187)  ->  Adam step (optimizer_factories.py:9-17).

One process per GPU; the only collective is ONE all-reduce per step on the flat fp32 gradient buffer (331 567
elements = 1.33 MB), issued through torch.distributed (NCCL over NVLink on B200, gloo in CPU tests of the host logic)."""
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor

from tha4_b200._lib import Context
from tha4_b200.nn.siren.morpher.siren_morpher_03 import SirenMorpher03
from tha4_b200.poser.general_poser_02 import GeneralPoser02

LOSS_TERMS = ('full_blended', 'full_warped', 'full_grid_change', 'full_color_change')   # siren_morpher_03_trainer.py:26-30


def flatten_parameters(module: torch.nn.Module) -> Tensor:
    """Moves the module's parameters into ONE contiguous fp32 buffer (state_dict order) and rebinds every parameter as
    a view of it, so the CUDA step, the gradient all-reduce and Adam all work on a single flat tensor."""
    params = list(module.parameters())
    flat = torch.cat([p.detach().reshape(-1).float() for p in params]).contiguous()
    off = 0
    for p in params:
        n = p.numel()
        p.data = flat[off:off + n].view_as(p)
        off += n
    return flat


class BodyMorpherDistiller:
    def __init__(self, teacher: GeneralPoser02, student: SirenMorpher03, betas=(0.9, 0.999), eps: float = 1e-8,
                 process_group=None):
        self.teacher = teacher
        self.student = student
        self.ctx: Context = teacher.get_context()
        teacher.get_modules()
        student.to(self.ctx.device)
        self.flat = flatten_parameters(student)
        assert self.flat.numel() == 331567
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.betas, self.eps = betas, eps
        self.step_count = 0
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1

    def train_step(self, image: Tensor, pose: Tensor, loss_weights: Sequence[float], lr: float,
                   want_losses: bool = True) -> Optional[Dict[str, float]]:
        """One iteration on this rank's batch (image [b,4,512,512], pose [b,45]; b <= 8 in total across ranks in the
        reference, distiller_config.py:100-104)."""
        with torch.no_grad():
            t = self.teacher.get_posing_outputs(image, pose)            # 33 tensors; 0 posed, 2 warped, 3 grid_change, 5 input
            losses = self.ctx.siren_morpher_train_step(t[5], pose if pose.dim() == 2 else pose.unsqueeze(0), t[0], t[2], t[3],
                                                       loss_weights, self.flat, self.grad, want_losses)
            if self.world > 1:
                dist.all_reduce(self.grad, group=self.group)            # the path's single collective
            self.step_count += 1
            self.ctx.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, lr, self.step_count, self.betas, self.eps,
                               grad_scale=1.0 / self.world)
            self.student._uploaded_key = None                           # inference path must re-pack the new weights
        if not want_losses:
            return None
        out = dict(zip(LOSS_TERMS, losses))
        out['loss'] = sum(w * l for w, l in zip(loss_weights, losses))
        return out

    def state_dict(self) -> Dict[str, Tensor]:
        return self.student.state_dict()
