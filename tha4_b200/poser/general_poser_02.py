"""`GeneralPoser02` -- mirror of src/tha4/poser/general_poser_02.py:10-98 (constructor arguments, lazy module
loading, rank-3/rank-1 -> batched promotion, optional subrect, `free()`, `to()`), with one addition: the modules
of a poser share one library Context so that the poser-level pipelines can run as a single C call."""
from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch import Tensor
from torch.nn import Module

from tha4_b200._lib import Context
from tha4_b200.poser.poser import PoseParameterGroup, Poser
from tha4_b200.shion.core.cached_computation import ComputationState


class GeneralPoser02(Poser):
    def __init__(self,
                 module_loaders: Dict[str, Callable[[], Module]],
                 device: torch.device,
                 output_length: int,
                 pose_parameters: List[PoseParameterGroup],
                 output_list_func: Callable[[ComputationState], List[Tensor]],
                 subrect: Optional[Tuple[Tuple[int, int], Tuple[int, int]]] = None,
                 default_output_index: int = 0,
                 image_size: int = 256,
                 dtype: torch.dtype = torch.float):
        self.dtype = dtype
        self.image_size = image_size
        self.default_output_index = default_output_index
        self.output_list_func = output_list_func
        self.subrect = subrect
        self.pose_parameters = pose_parameters
        self.device = torch.device(device)
        self.module_loaders = module_loaders
        self.modules = None
        self.context: Optional[Context] = None
        self.num_parameters = sum(p.get_arity() for p in self.pose_parameters)
        self.output_length = output_length

    def get_image_size(self) -> int:
        return self.image_size

    def get_context(self) -> Context:
        if self.context is None:
            self.context = Context(self.device)
        return self.context

    def get_modules(self):   # general_poser_02.py:41-49
        if self.modules is None:
            ctx = self.get_context()
            self.modules = {}
            for key in self.module_loaders:
                module = self.module_loaders[key]()
                self.modules[key] = module
                module.attach_context(ctx)
                module.to(self.device)
                module.train(False)
        return self.modules

    def get_pose_parameter_groups(self) -> List[PoseParameterGroup]:
        return self.pose_parameters

    def get_num_parameters(self) -> int:
        return self.num_parameters

    def pose(self, image: Tensor, pose: Tensor, output_index: Optional[int] = None) -> Tensor:
        if output_index is None:
            output_index = self.default_output_index
        return self.get_posing_outputs(image, pose)[output_index]

    def get_posing_outputs(self, image: Tensor, pose: Tensor) -> List[Tensor]:   # general_poser_02.py:63-79
        modules = self.get_modules()
        if len(image.shape) == 3:
            image = image.unsqueeze(0)
        if len(pose.shape) == 1:
            pose = pose.unsqueeze(0)
        if self.subrect is not None:
            image = image[:, :, self.subrect[0][0]:self.subrect[0][1], self.subrect[1][0]:self.subrect[1][1]]
        state = ComputationState(modules=modules, accumulated_modules={}, batch=[image, pose], outputs={})
        state.context = self.get_context()
        return self.output_list_func(state)

    def get_output_length(self) -> int:
        return self.output_length

    def free(self):
        self.modules = None
        self.context = None

    def get_dtype(self) -> torch.dtype:
        return self.dtype

    def to(self, device: torch.device) -> 'GeneralPoser02':
        device = torch.device(device)
        if device == self.device:
            return self
        self.device = device
        self.modules = None      # rebuilt (and re-uploaded) lazily on the new device
        self.context = None
        return self
