"""The `Poser` protocol and the pose-parameter schema, mirroring src/tha4/poser/poser.py:9-161 of the reference
(same class, method and argument names, so GUI code written against the reference runs unchanged)."""
from abc import ABC, abstractmethod
from enum import Enum
from typing import List, Optional, Tuple

import torch
from torch import Tensor


class PoseParameterCategory(Enum):   # poser.py:9-17
    EYEBROW = 1
    EYE = 2
    IRIS_MORPH = 3
    IRIS_ROTATION = 4
    MOUTH = 5
    FACE_ROTATION = 6
    BODY_ROTATION = 7
    BREATHING = 8


class PoseParameterGroup:   # poser.py:20-68
    def __init__(self, group_name: str, parameter_index: int, category: PoseParameterCategory, arity: int = 1,
                 discrete: bool = False, default_value: float = 0.0, range: Optional[Tuple[float, float]] = None):
        assert arity == 1 or arity == 2
        self.parameter_names = [group_name] if arity == 1 else [group_name + '_left', group_name + '_right']
        self.range = (0.0, 1.0) if range is None else range
        self.default_value = default_value
        self.discrete = discrete
        self.arity = arity
        self.category = category
        self.parameter_index = parameter_index
        self.group_name = group_name

    def get_arity(self) -> int:
        return self.arity

    def get_group_name(self) -> str:
        return self.group_name

    def get_parameter_names(self) -> List[str]:
        return self.parameter_names

    def is_discrete(self) -> bool:
        return self.discrete

    def get_range(self) -> Tuple[float, float]:
        return self.range

    def get_default_value(self):
        return self.default_value

    def get_parameter_index(self):
        return self.parameter_index

    def get_category(self) -> PoseParameterCategory:
        return self.category


class PoseParameters:   # poser.py:71-129
    def __init__(self, pose_parameter_groups: List[PoseParameterGroup]):
        self.pose_parameter_groups = pose_parameter_groups

    def get_parameter_index(self, name: str) -> int:
        index = 0
        for group in self.pose_parameter_groups:
            for param_name in group.parameter_names:
                if name == param_name:
                    return index
                index += 1
        raise RuntimeError('Cannot find parameter with name %s' % name)

    def get_parameter_name(self, index: int) -> str:
        assert 0 <= index < self.get_parameter_count()
        for group in self.pose_parameter_groups:
            if index < group.get_arity():
                return group.get_parameter_names()[index]
            index -= group.arity
        raise RuntimeError('Something is wrong here!!!')

    def get_pose_parameter_groups(self):
        return self.pose_parameter_groups

    def get_parameter_count(self):
        return sum(group.arity for group in self.pose_parameter_groups)

    class Builder:
        def __init__(self):
            self.index = 0
            self.pose_parameter_groups = []

        def add_parameter_group(self, group_name: str, category: PoseParameterCategory, arity: int = 1,
                                discrete: bool = False, default_value: float = 0.0,
                                range: Optional[Tuple[float, float]] = None):
            self.pose_parameter_groups.append(
                PoseParameterGroup(group_name, self.index, category, arity, discrete, default_value, range))
            self.index += arity
            return self

        def build(self) -> 'PoseParameters':
            return PoseParameters(self.pose_parameter_groups)


class Poser(ABC):   # poser.py:132-161
    @abstractmethod
    def get_image_size(self) -> int:
        pass

    @abstractmethod
    def get_output_length(self) -> int:
        pass

    @abstractmethod
    def get_pose_parameter_groups(self) -> List[PoseParameterGroup]:
        pass

    @abstractmethod
    def get_num_parameters(self) -> int:
        pass

    @abstractmethod
    def pose(self, image: Tensor, pose: Tensor, output_index: int = 0) -> Tensor:
        pass

    @abstractmethod
    def get_posing_outputs(self, image: Tensor, pose: Tensor) -> List[Tensor]:
        pass

    def get_dtype(self) -> torch.dtype:
        return torch.float

    @abstractmethod
    def to(self, device: torch.device):
        pass
