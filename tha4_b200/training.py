"""Distillation trainer around the CUDA inner loops of tha4_b200/distill.py: phase schedules, the loop gates and the
on-disk checkpoint / snapshot format of the reference's distributed trainer, so that a full distillation run
(the reference's one published number is the 30 h full distillation) can be started, stopped and
resumed on the new path -- and so that training states written by either implementation load in the other.

Reference:
  * state directory layout and contents -- src/tha4/shion/core/training/distrib/distributed_training_states.py:29-88
    (`examples_seen_so_far.txt`, `module_<name>.pt`, `optimizer_<name>.pt`, `rng_state_<rank:08d>.pt`; rank 0 writes the
    shared files, every rank its RNG state, barriers around the save) and :96-152 (load), :216-252 (can_load);
  * loop gates -- src/tha4/shion/core/training/distrib/distributed_trainer.py:145-167 (which state to resume from),
    :310-389 (learning rate per iteration, one training iteration, examples_seen += batch * world, checkpoint and
    snapshot saves at their example counts);
  * body-student phases -- src/tha4/distiller/distiller_config.py:177-232 with the lookup rule of
    src/tha4/nn/siren/morpher/siren_morpher_03_trainer.py:76-103 (a phase applies while examples_seen < its upper
    bound; the last phase applies forever);
  * face-student learning-rate steps -- src/tha4/nn/siren/face_morpher/siren_face_morpher_00_trainer.py:28-52,134-151.

What differs, deliberately: the optimiser state lives in two flat fp32 buffers (the fused Adam kernel's moments); it is
converted to / from `torch.optim.Adam.state_dict()` at the file boundary.  The training batch is a pure function of
`examples_seen_so_far` (the reference restarts its DataLoader iterator at the beginning of the epoch after a resume,
distributed_trainer.py:206-231), which is what makes "stop + resume" reproduce the uninterrupted run exactly.
"""
import logging
import os
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from tha4_b200.distill import FACE_LOSS_WEIGHTS, LOSS_TERMS
from tha4_b200.shion.core.load_save import torch_load, torch_save

KEY_MODULE = 'module'          # siren_morpher_protocols_03.py:21 -- the file names below are derived from it


# ------------------------------------------------------------------------------------------------ schedules
class TrainingPhase:
    """siren_morpher_03_trainer.py:66-74."""

    def __init__(self, num_examples_upper_bound: int, learning_rate: float, loss_weights: Dict[str, float]):
        self.num_examples_upper_bound = num_examples_upper_bound
        self.learning_rate = learning_rate
        self.loss_weights = {term: float(loss_weights.get(term, 0.0)) for term in LOSS_TERMS}     # LossWeights: missing terms weigh 0


class TrainingPhases:
    """siren_morpher_03_trainer.py:76-124: strictly increasing upper bounds; phase i applies while
    examples_seen_so_far < bound_i, the last phase applies from then on."""

    def __init__(self, phases: Sequence[TrainingPhase]):
        assert len(phases) > 0
        for a, b in zip(phases[:-1], phases[1:]):
            assert a.num_examples_upper_bound < b.num_examples_upper_bound
        self.phases = list(phases)

    def phase_at(self, examples_seen_so_far: int) -> TrainingPhase:
        for phase in self.phases[:-1]:
            if examples_seen_so_far < phase.num_examples_upper_bound:
                return phase
        return self.phases[-1]

    def learning_rate(self, examples_seen_so_far: int) -> float:
        return self.phase_at(examples_seen_so_far).learning_rate

    def loss_weights(self, examples_seen_so_far: int) -> List[float]:
        w = self.phase_at(examples_seen_so_far).loss_weights
        return [w[term] for term in LOSS_TERMS]

    def total_examples(self) -> int:
        return self.phases[-1].num_examples_upper_bound


def body_morpher_training_phases() -> TrainingPhases:
    """The six phases of the shipped distiller (distiller_config.py:177-232)."""
    def w(blended, warped, grid, color):
        return {'full_blended': blended, 'full_warped': warped, 'full_grid_change': grid, 'full_color_change': color}
    return TrainingPhases([
        TrainingPhase(200_000, 1e-4, w(0.25, 0.25, 0.5, 2.0)),
        TrainingPhase(400_000, 3e-5, w(0.25, 0.25, 0.5, 2.0)),
        TrainingPhase(600_000, 3e-5, w(1.0, 2.5, 5.0, 1.0)),
        TrainingPhase(800_000, 1e-5, w(1.0, 2.5, 5.0, 1.0)),
        TrainingPhase(1_300_000, 1e-5, w(10.0, 1.0, 1.0, 1.0)),
        TrainingPhase(1_500_000, 3e-6, w(10.0, 1.0, 1.0, 1.0)),
    ])


class FaceMorpherSchedule:
    """siren_face_morpher_00_trainer.py:28-52,134-151: base_lr, /3, /10, /30 at 2x / 5x / 8x the checkpoint interval;
    constant loss weights 1.0 / 20.0 (:168-186)."""

    def __init__(self, num_training_total_examples: int = 1_000_000, num_training_examples_per_checkpoint: int = 100_000,
                 boundaries: Optional[Sequence[int]] = None, base_learning_rate: float = 1e-4):
        assert num_training_total_examples % num_training_examples_per_checkpoint == 0
        self.total = num_training_total_examples
        self.per_checkpoint = num_training_examples_per_checkpoint
        self.boundaries = list(boundaries) if boundaries is not None else [2 * self.per_checkpoint, 5 * self.per_checkpoint, 8 * self.per_checkpoint]
        self.base = base_learning_rate

    def learning_rate(self, examples_seen_so_far: int) -> float:
        for bound, div in zip(self.boundaries, (1.0, 3.0, 10.0)):
            if examples_seen_so_far < bound:
                return self.base / div
        return self.base / 30.0

    def loss_weights(self, examples_seen_so_far: int) -> List[float]:
        return list(FACE_LOSS_WEIGHTS)

    def total_examples(self) -> int:
        return self.total


def get_least_greater_multiple(x: int, m: int) -> int:
    """shion/core/training/util.py:21-29."""
    assert x >= 0 and m > 0
    return (x // m + 1) * m


# ------------------------------------------------------------------------------------------------ state on disk
def examples_seen_so_far_file_name(prefix: str) -> str:
    return prefix + '/examples_seen_so_far.txt'


def module_file_name(prefix: str, module_name: str = KEY_MODULE) -> str:
    return '%s/module_%s.pt' % (prefix, module_name)


def optimizer_file_name(prefix: str, module_name: str = KEY_MODULE) -> str:
    return '%s/optimizer_%s.pt' % (prefix, module_name)


def rng_state_file_name(prefix: str, rank: int) -> str:
    return '%s/rng_state_%08d.pt' % (prefix, rank)


def can_load(prefix: str, world_size: int) -> bool:
    """distributed_training_states.py:216-252 (no accumulated modules on this path)."""
    if not os.path.isdir(prefix):
        return False
    names = [examples_seen_so_far_file_name(prefix), module_file_name(prefix), optimizer_file_name(prefix)]
    names += [rng_state_file_name(prefix, r) for r in range(world_size)]
    return all(os.path.isfile(n) for n in names)


def read_examples_seen_so_far(prefix: str) -> int:
    with open(examples_seen_so_far_file_name(prefix)) as fin:
        return int(fin.readlines()[0])


def adam_state_dict_from_flat(distiller, lr: float) -> dict:
    """The flat moment buffers in `torch.optim.Adam.state_dict()` form (one entry per parameter in state_dict order), built by
    a real torch optimiser so that the file has exactly the layout the installed torch -- and the reference's
    `optimizer.load_state_dict` (distributed_training_states.py:141-146) -- expects."""
    params = list(distiller.student.parameters())
    opt = torch.optim.Adam(params, lr=lr, betas=distiller.betas, eps=distiller.eps)
    if distiller.step_count > 0:
        off = 0
        for p in params:
            n = p.numel()
            opt.state[p] = {'step': torch.tensor(float(distiller.step_count)),
                            'exp_avg': distiller.exp_avg[off:off + n].view_as(p).clone(),
                            'exp_avg_sq': distiller.exp_avg_sq[off:off + n].view_as(p).clone()}
            off += n
    sd = opt.state_dict()
    for st in sd['state'].values():                      # files are device-agnostic: the reference moves them with optimizer_to_device
        for k, v in st.items():
            if torch.is_tensor(v):
                st[k] = v.cpu()
    return sd


def load_adam_state_into_flat(distiller, state_dict: dict) -> float:
    """Inverse of adam_state_dict_from_flat; returns the learning rate stored in the file."""
    params = list(distiller.student.parameters())
    state = state_dict['state']
    distiller.exp_avg.zero_()
    distiller.exp_avg_sq.zero_()
    step = 0
    off = 0
    for i, p in enumerate(params):
        n = p.numel()
        st = state.get(i)
        if st is not None:
            distiller.exp_avg[off:off + n].copy_(st['exp_avg'].reshape(-1))
            distiller.exp_avg_sq[off:off + n].copy_(st['exp_avg_sq'].reshape(-1))
            step = int(float(st['step']))
        off += n
    distiller.step_count = step
    return float(state_dict['param_groups'][0]['lr'])


class DistillTrainingState:
    """`examples_seen_so_far` + the student's weights + the Adam moments of one distiller (the equivalent of
    DistributedTrainingState for KEY_MODULE = the student; the frozen teacher is not part of the state)."""

    def __init__(self, distiller, examples_seen_so_far: int = 0):
        self.distiller = distiller
        self.examples_seen_so_far = examples_seen_so_far

    def save(self, prefix: str, rank: int, barrier: Callable[[], None], lr: float = 0.0):
        if rank == 0:
            os.makedirs(prefix, exist_ok=True)
        barrier()
        torch_save(torch.get_rng_state(), rng_state_file_name(prefix, rank))
        if rank == 0:
            with open(examples_seen_so_far_file_name(prefix), 'wt') as fout:
                fout.write('%d\n' % self.examples_seen_so_far)
            torch_save({k: v.detach().cpu().clone() for k, v in self.distiller.student.state_dict().items()}, module_file_name(prefix))
            torch_save(adam_state_dict_from_flat(self.distiller, lr), optimizer_file_name(prefix))
            logging.info('Saved training state to %s (%d examples)', prefix, self.examples_seen_so_far)
        barrier()

    def load(self, prefix: str, rank: int):
        self.examples_seen_so_far = read_examples_seen_so_far(prefix)
        sd = torch_load(module_file_name(prefix))
        d = self.distiller
        keys = list(d.student.state_dict().keys())
        assert set(keys) == set(sd.keys()), 'module file does not hold the student\'s state_dict'
        d.flat.copy_(torch.cat([sd[k].reshape(-1).float() for k in keys]).to(d.flat.device))
        d.student._uploaded_key = None
        load_adam_state_into_flat(d, torch_load(optimizer_file_name(prefix)))
        d.grad.zero_()
        torch.set_rng_state(torch_load(rng_state_file_name(prefix, rank)))
        logging.info('Loaded training state from %s (%d examples)', prefix, self.examples_seen_so_far)


# ------------------------------------------------------------------------------------------------ batches
class PoseBatches:
    """The training pose stream: a [N,45] pose table, shuffled per epoch the way DistributedSampler(shuffle=True,
    drop_last=True) does (permutation seeded with seed + epoch, rank r takes indices r, r + world, ...), cut into
    per-rank batches -- addressed by `examples_seen_so_far`, so a resumed run continues where the stopped one was."""

    def __init__(self, poses: Tensor, batch_size: int, rank: int, world: int, seed: int = 0):
        assert poses.dim() == 2 and poses.shape[1] == 45
        self.poses, self.batch, self.rank, self.world, self.seed = poses, batch_size, rank, world, seed
        n = (poses.shape[0] // world) * world
        self.epoch_size = (n // (batch_size * world)) * (batch_size * world)        # effective epoch (distributed_trainer.py:193-198)
        assert self.epoch_size > 0, 'pose table smaller than one global batch'
        self._epoch, self._perm = -1, None

    def get(self, examples_seen_so_far: int) -> Tensor:
        epoch, within = divmod(examples_seen_so_far, self.epoch_size)
        if epoch != self._epoch:
            g = torch.Generator().manual_seed(self.seed + epoch)
            self._perm = torch.randperm(self.poses.shape[0], generator=g)
            self._epoch = epoch
        mine = self._perm[self.rank:(self.poses.shape[0] // self.world) * self.world:self.world]
        first = (within // (self.batch * self.world)) * self.batch
        return self.poses[mine[first:first + self.batch]]


# ------------------------------------------------------------------------------------------------ trainer
class DistillTrainer:
    """Loop gates of DistributedTrainer.train (distributed_trainer.py:310-389) around one distiller.

    step_fn(batch_poses, loss_weights, lr) runs ONE training iteration on this rank's batch (BodyMorpherDistiller /
    FaceMorpherDistiller.train_step bound to the character image [and face mask]); everything else -- schedule lookup,
    example counting, checkpoints at `per_checkpoint` multiples under <prefix>/checkpoint/%04d, snapshots every
    `per_snapshot` examples under <prefix>/snapshot, resume from the newest usable state -- happens here."""

    def __init__(self, prefix: str, distiller, schedule, batches: PoseBatches,
                 step_fn: Callable[[Tensor, List[float], float], Optional[dict]],
                 per_checkpoint: int = 100_000, per_snapshot: int = 10_000, rank: int = 0, world: int = 1,
                 log_fn: Optional[Callable[[int, dict], None]] = None):
        total = schedule.total_examples()
        assert total % per_checkpoint == 0
        self.prefix, self.distiller, self.schedule, self.batches, self.step_fn = prefix, distiller, schedule, batches, step_fn
        self.checkpoint_examples = [0] + [per_checkpoint * (i + 1) for i in range(total // per_checkpoint)]
        self.per_snapshot, self.rank, self.world, self.log_fn = per_snapshot, rank, world, log_fn
        self.state = DistillTrainingState(distiller)

    # -- paths (distributed_trainer.py:104-130)
    def snapshot_prefix(self) -> str:
        return self.prefix + '/snapshot'

    def checkpoint_prefix(self, index: int) -> str:
        return '%s/checkpoint/%04d' % (self.prefix, index)

    def barrier(self):
        if self.world > 1 and dist.is_available() and dist.is_initialized():
            dist.barrier()

    # -- resume (distributed_trainer.py:145-167)
    def load_previous_training_state(self, target_checkpoint_examples: int):
        batch_total = self.batches.batch        # the reference compares with the per-rank batch size (get_batch_size())
        candidates = [self.snapshot_prefix()] + [self.checkpoint_prefix(i) for i in range(len(self.checkpoint_examples) - 1, -1, -1)]
        for prefix in candidates:
            if can_load(prefix, self.world) and read_examples_seen_so_far(prefix) - target_checkpoint_examples < batch_total:
                self.state.load(prefix, self.rank)
                return prefix
        self.state.examples_seen_so_far = 0
        self.state.save(self.checkpoint_prefix(0), self.rank, self.barrier, self.schedule.learning_rate(0))
        self.state.load(self.checkpoint_prefix(0), self.rank)
        return None

    def checkpoint_index_to_save(self, examples_seen_so_far: int) -> int:
        index = 0
        for i, n in enumerate(self.checkpoint_examples):
            if n <= examples_seen_so_far:
                index = i
        return index

    def train(self, target_checkpoint_examples: Optional[int] = None, max_iterations: Optional[int] = None) -> int:
        """Runs until `examples_seen_so_far >= target` (default: the last checkpoint), or for `max_iterations` iterations
        (a stop in the middle of a run: no state is written beyond the regular gates).  Returns examples_seen_so_far."""
        if target_checkpoint_examples is None:
            target_checkpoint_examples = self.checkpoint_examples[-1]
        self.load_previous_training_state(target_checkpoint_examples)
        st = self.state
        iterations = 0
        while st.examples_seen_so_far < target_checkpoint_examples:
            if max_iterations is not None and iterations >= max_iterations:
                break
            lr = self.schedule.learning_rate(st.examples_seen_so_far)
            weights = self.schedule.loss_weights(st.examples_seen_so_far)
            out = self.step_fn(self.batches.get(st.examples_seen_so_far), weights, lr)
            if self.log_fn is not None and out is not None:
                self.log_fn(st.examples_seen_so_far, out)
            next_checkpoint = next((n for n in self.checkpoint_examples if n > st.examples_seen_so_far), self.checkpoint_examples[-1])
            next_snapshot = get_least_greater_multiple(st.examples_seen_so_far, self.per_snapshot)
            st.examples_seen_so_far += self.batches.batch * self.world
            iterations += 1
            if st.examples_seen_so_far >= next_checkpoint:
                st.save(self.checkpoint_prefix(self.checkpoint_index_to_save(st.examples_seen_so_far)), self.rank, self.barrier, lr)
                if next_checkpoint != next_snapshot:
                    st.save(self.snapshot_prefix(), self.rank, self.barrier, lr)
            if st.examples_seen_so_far >= next_snapshot:
                st.save(self.snapshot_prefix(), self.rank, self.barrier, lr)
        return st.examples_seen_so_far
