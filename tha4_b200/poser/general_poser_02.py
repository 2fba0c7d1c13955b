"""`GeneralPoser02`: the poser object every mode's `create_poser` returns.

Interface of src/tha4/poser/general_poser_02.py:10-98 (constructor keywords, lazy module construction, promotion of a
rank-3 image / rank-1 pose to a batch, optional `subrect` crop, `free()`, `to()`); the implementation differs in that
all modules of a poser are attached to ONE library `Context` (device workspace + packed weights), which is what lets a
mode run its whole pipeline as a single C call, and that moving the poser re-creates that context on the new device."""
from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch import Tensor
from torch.nn import Module

from tha4_b200._lib import Context
from tha4_b200.poser.poser import PoseParameterGroup, Poser
from tha4_b200.shion.core.cached_computation import ComputationState

Rect = Tuple[Tuple[int, int], Tuple[int, int]]      # ((row0, row1), (col0, col1))


class GeneralPoser02(Poser):
    def __init__(self,
                 module_loaders: Dict[str, Callable[[], Module]],
                 device: torch.device,
                 output_length: int,
                 pose_parameters: List[PoseParameterGroup],
                 output_list_func: Callable[[ComputationState], List[Tensor]],
                 subrect: Optional[Rect] = None,
                 default_output_index: int = 0,
                 image_size: int = 256,
                 dtype: torch.dtype = torch.float):
        self._loaders = dict(module_loaders)
        self._pipeline = output_list_func
        self._groups = list(pose_parameters)
        self._declared_outputs = output_length
        self._default_output = default_output_index
        self._crop = subrect
        self._size = image_size
        self._dtype = dtype
        self.device = torch.device(device)
        self.modules: Optional[Dict[str, Module]] = None      # built on first use
        self.context: Optional[Context] = None

    # ------------------------------------------------------------------ lazily built state
    def get_context(self) -> Context:
        if self.context is None:
            self.context = Context(self.device)
        return self.context

    def get_modules(self) -> Dict[str, Module]:
        if self.modules is None:
            ctx = self.get_context()
            built = {}
            for key, make in self._loaders.items():
                module = make()
                module.attach_context(ctx)
                module.to(self.device)
                module.train(False)
                built[key] = module
            self.modules = built
        return self.modules

    def free(self):
        """Drops modules and device workspace; they come back on the next call."""
        self.modules, self.context = None, None

    def to(self, device: torch.device) -> 'GeneralPoser02':
        target = torch.device(device)
        if target != self.device:
            self.free()                      # weights are re-uploaded lazily on the new device
            self.device = target
        return self

    # ------------------------------------------------------------------ schema getters
    def get_image_size(self) -> int:
        return self._size

    def get_output_length(self) -> int:
        return self._declared_outputs

    def get_pose_parameter_groups(self) -> List[PoseParameterGroup]:
        return self._groups

    def get_num_parameters(self) -> int:
        return sum(g.get_arity() for g in self._groups)

    def get_dtype(self) -> torch.dtype:
        return self._dtype

    # ------------------------------------------------------------------ posing
    def get_posing_outputs(self, image: Tensor, pose: Tensor) -> List[Tensor]:
        image = image[None] if image.dim() == 3 else image
        pose = pose[None] if pose.dim() == 1 else pose
        if self._crop is not None:
            (r0, r1), (c0, c1) = self._crop
            image = image[:, :, r0:r1, c0:c1]
        state = ComputationState(modules=self.get_modules(), accumulated_modules={}, batch=[image, pose], outputs={})
        state.context = self.get_context()
        return self._pipeline(state)

    def pose_to_srgb8(self, image: Tensor, pose: Tensor, background=None, rint: bool = False,
                      output_index: Optional[int] = None) -> Tensor:
        """Not in the reference: `pose()` followed, on the GPU, by the display conversion every app applies to the frame
        (character_model_ifacialmocap_puppeteer.py:325-349) -> [B,H,W,4] uint8 sRGB on the device."""
        return self.get_context().frame_to_srgb8(self.pose(image, pose, output_index), background, rint)

    def pose(self, image: Tensor, pose: Tensor, output_index: Optional[int] = None) -> Tensor:
        outputs = self.get_posing_outputs(image, pose)
        return outputs[self._default_output if output_index is None else output_index]
