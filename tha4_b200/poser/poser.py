"""The `Poser` interface and the pose-parameter schema objects.

Interface-compatible with src/tha4/poser/poser.py:9-161 of the reference (class names, method names, argument names and
enum members are what the GUIs and puppeteers call), implemented here as an immutable record type with a table of
accessors and a schema object that indexes its parameters once at construction."""
from abc import ABC, abstractmethod
from enum import Enum
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

PoseParameterCategory = Enum('PoseParameterCategory', [                       # values as in poser.py:9-17
    ('EYEBROW', 1), ('EYE', 2), ('IRIS_MORPH', 3), ('IRIS_ROTATION', 4), ('MOUTH', 5), ('FACE_ROTATION', 6),
    ('BODY_ROTATION', 7), ('BREATHING', 8)])

_SIDES = ('_left', '_right')


class PoseParameterGroup:
    """One slider (arity 1) or a left/right pair of sliders (arity 2) of the pose vector."""
    __slots__ = ('group_name', 'parameter_index', 'category', 'arity', 'discrete', 'default_value', 'range', 'parameter_names')

    def __init__(self, group_name: str, parameter_index: int, category: PoseParameterCategory, arity: int = 1,
                 discrete: bool = False, default_value: float = 0.0, range: Optional[Tuple[float, float]] = None):
        if arity not in (1, 2):
            raise AssertionError('a pose parameter group has one parameter or a left/right pair')
        object.__setattr__(self, 'group_name', group_name)
        object.__setattr__(self, 'parameter_index', parameter_index)
        object.__setattr__(self, 'category', category)
        object.__setattr__(self, 'arity', arity)
        object.__setattr__(self, 'discrete', discrete)
        object.__setattr__(self, 'default_value', default_value)
        object.__setattr__(self, 'range', (0.0, 1.0) if range is None else tuple(range))
        object.__setattr__(self, 'parameter_names', [group_name] if arity == 1 else [group_name + s for s in _SIDES])

    def __setattr__(self, key, value):
        raise AttributeError('PoseParameterGroup is immutable')

    def __repr__(self):
        return 'PoseParameterGroup(%r @%d x%d %s)' % (self.group_name, self.parameter_index, self.arity, self.category.name)


def _accessor(field):
    return lambda self: getattr(self, field)


# accessor methods of the reference API -> record field
for _method, _field in (('get_arity', 'arity'), ('get_group_name', 'group_name'), ('get_parameter_names', 'parameter_names'),
                        ('is_discrete', 'discrete'), ('get_range', 'range'), ('get_default_value', 'default_value'),
                        ('get_parameter_index', 'parameter_index'), ('get_category', 'category')):
    setattr(PoseParameterGroup, _method, _accessor(_field))


class PoseParameters:
    """Ordered list of groups = layout of the pose vector (45 entries for THA4)."""

    def __init__(self, pose_parameter_groups: List[PoseParameterGroup]):
        self.pose_parameter_groups = list(pose_parameter_groups)
        self._names: List[str] = [n for g in self.pose_parameter_groups for n in g.parameter_names]
        self._index: Dict[str, int] = {}
        for i, n in enumerate(self._names):
            self._index.setdefault(n, i)              # first occurrence wins, like a linear search would

    def get_parameter_index(self, name: str) -> int:
        try:
            return self._index[name]
        except KeyError:
            raise RuntimeError('Cannot find parameter with name %s' % name) from None

    def get_parameter_name(self, index: int) -> str:
        assert 0 <= index < len(self._names)
        return self._names[index]

    def get_pose_parameter_groups(self) -> List[PoseParameterGroup]:
        return self.pose_parameter_groups

    def get_parameter_count(self) -> int:
        return len(self._names)

    class Builder:
        """Appends groups and assigns each the running offset into the pose vector."""

        def __init__(self):
            self._groups: List[PoseParameterGroup] = []

        def add_parameter_group(self, group_name: str, category: PoseParameterCategory, arity: int = 1, discrete: bool = False,
                                default_value: float = 0.0, range: Optional[Tuple[float, float]] = None):
            offset = sum(g.arity for g in self._groups)
            self._groups.append(PoseParameterGroup(group_name, offset, category, arity, discrete, default_value, range))
            return self

        def build(self) -> 'PoseParameters':
            return PoseParameters(self._groups)


class Poser(ABC):
    """What a GUI / puppeteer / distillation protocol needs from a poser (poser.py:132-161)."""

    @abstractmethod
    def pose(self, image: Tensor, pose: Tensor, output_index: int = 0) -> Tensor:
        """One output tensor ([B,C,H,W]) for image [B,4,S,S] | [4,S,S] and pose [B,P] | [P]."""

    @abstractmethod
    def get_posing_outputs(self, image: Tensor, pose: Tensor) -> List[Tensor]:
        """Every tensor the pipeline produces, in the mode's fixed order."""

    @abstractmethod
    def get_pose_parameter_groups(self) -> List[PoseParameterGroup]:
        """Layout of the pose vector."""

    @abstractmethod
    def get_num_parameters(self) -> int:
        """Length of the pose vector."""

    @abstractmethod
    def get_image_size(self) -> int:
        """Side of the square input image."""

    @abstractmethod
    def get_output_length(self) -> int:
        """Number of outputs the mode declares."""

    @abstractmethod
    def to(self, device: torch.device):
        """Moves the poser; returns it."""

    def get_dtype(self) -> torch.dtype:
        return torch.float
