"""Base class of the seven network modules: an `nn.Module` that owns parameters under the reference's state_dict
keys and whose forward is one C-ABI call into libtha4_b200.so (no PyTorch-op fallback)."""
import math
from typing import Dict, Optional

import torch
from torch import Tensor
from torch.nn import Module, Parameter

from tha4_b200._lib import Context, Tha4Error


class _Node(Module):
    """Anonymous container so that dotted reference keys ('body.down_blocks.0.conv0.weight') map onto a module tree."""

    def forward(self, *args, **kwargs):
        raise Tha4Error('container node: call the owning network module instead')


def _init_tensor(shape, role: str) -> Tensor:
    """Reference initialisers by role (see state_dict_spec.py)."""
    t = torch.empty(shape, dtype=torch.float32)
    if role in ('conv', 'conv1', 'convT', 'student_last'):
        fan_in = shape[1] * (shape[2] * shape[3] if len(shape) == 4 else 1)
        return t.normal_(0.0, math.sqrt(2.0 / fan_in))                      # kaiming_normal_ (init_function.py:14-16)
    if role in ('zconv', 'grid_head', 'last'):
        return t.zero_()                                                    # unet.py:26-30; poser_args.py:62-68
    if role in ('linear', 'film'):
        bound = 1.0 / math.sqrt(shape[1])
        return t.uniform_(-bound, bound)                                    # torch.nn.Linear default
    if role == 'norm_w':
        return t.fill_(1.0)
    if role in ('norm_b',):
        return t.zero_()
    if role == 'bias':
        return t.zero_()
    if role == 'siren_first':
        return t.uniform_(-1.0 / shape[1], 1.0 / shape[1])                  # siren.py:32
    if role == 'siren':
        b = math.sqrt(6.0 / shape[1]) / 30.0
        return t.uniform_(-b, b)                                            # siren.py:34-36
    if role == 'siren_bias':
        b = 1.0 / math.sqrt(shape[0]) if len(shape) == 1 else 0.0
        return t.uniform_(-b, b)
    raise ValueError(role)


class NativeModule(Module):
    NET_NAME: str = ''

    def __init__(self, spec):
        super().__init__()
        self._spec = spec
        for key, shape, role in spec:
            parts = key.split('.')
            node = self
            for p in parts[:-1]:
                if p not in node._modules:
                    node.add_module(p, _Node())
                node = node._modules[p]
            node.register_parameter(parts[-1], Parameter(_init_tensor(shape, role)))
        self._ctx: Optional[Context] = None
        self._uploaded_key = None

    # ------------------------------------------------------------------ context / weights
    def attach_context(self, ctx: Context):
        self._ctx = ctx
        self._uploaded_key = None

    def _device(self) -> torch.device:
        return next(self.parameters()).device

    def context(self) -> Context:
        dev = self._device()
        if self._ctx is None or self._ctx.device != torch.device('cuda', dev.index if dev.index is not None else 0):
            if dev.type != 'cuda':
                raise Tha4Error('%s lives on %s: tha4_b200 modules run on CUDA only (no CPU fallback); call .to("cuda")'
                                % (type(self).__name__, dev))
            self._ctx = Context(dev)
            self._uploaded_key = None
        return self._ctx

    def sync_weights(self) -> Context:
        """(Re)packs the parameters into the library when they changed since the last upload."""
        ctx = self.context()
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if key != self._uploaded_key:
            ctx.load_net(self.NET_NAME, self.state_dict())
            ctx.modules.add(self)
            self._uploaded_key = key
        return ctx

    def load_state_dict(self, state_dict: Dict[str, Tensor], strict: bool = True):
        self._uploaded_key = None
        return super().load_state_dict(state_dict, strict=strict)
