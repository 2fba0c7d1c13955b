"""The distilled student poser (two SIRENs) -- mirror of src/tha4/poser/modes/mode_14.py:40-162.
The whole DAG (face SIREN from pose[:, :39], paste at rows 80:208 / cols 192:320, body SIREN, mode_14.py:52-90) is one
C call, `tha4_student_forward`, returning body(5) + face(1)."""
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import torch
from torch import Tensor

from tha4_b200.nn.siren.face_morpher.siren_face_morpher_00 import SirenFaceMorpher00
from tha4_b200.nn.siren.morpher.siren_morpher_03 import SirenMorpher03
from tha4_b200.poser.general_poser_02 import GeneralPoser02
from tha4_b200.poser.modes.pose_parameters import get_pose_parameters
from tha4_b200.shion.core.cached_computation import CachedComputationProtocol, ComputationState
from tha4_b200.shion.core.load_save import torch_load

KEY_FACE_MORPHER = "face_morpher"
KEY_BODY_MORPHER = "body_morpher"


@dataclass
class Keys:   # mode_14.py:21-35
    face_morpher: str = KEY_FACE_MORPHER
    face_morpher_output: str = "face_morpher_output"
    face_morpher_input_image: str = "face_morpher_input_image"
    face_morpher_input_pose: str = "face_morpher_input_pose"
    body_morpher_input_image: str = "body_morpher_input_image"
    body_morpher: str = KEY_BODY_MORPHER
    body_morpher_output: str = "body_morpher_output"
    all_outputs: str = "all_outputs"


@dataclass
class Indices:
    original_image: int = 0
    original_pose: int = 1


class TwoStepPoserComputationProtocol(CachedComputationProtocol):
    def __init__(self, keys: Optional[Keys] = None, indices: Optional[Indices] = None):
        super().__init__()
        self.keys = Keys() if keys is None else keys
        self.indices = Indices() if indices is None else indices

    def compute_func(self):
        def func(state: ComputationState) -> List[Tensor]:
            return self.get_output(self.keys.all_outputs, state)

        return func

    def compute_output(self, key: str, state: ComputationState) -> Any:
        if key == self.keys.all_outputs:
            for name in (self.keys.face_morpher, self.keys.body_morpher):
                state.modules[name].sync_weights()
            image = state.batch[self.indices.original_image]
            # a float16 image selects the fp16 I/O entry point (tha4_student_forward_io, io_dtype = 1): half outputs
            forward = state.context.student_forward_half if image.dtype == torch.float16 else state.context.student_forward
            outputs = forward(image, state.batch[self.indices.original_pose].float())
            state.outputs[self.keys.body_morpher_output] = outputs[0:5]
            state.outputs[self.keys.face_morpher_output] = outputs[5]
            return outputs
        elif key in (self.keys.body_morpher_output, self.keys.face_morpher_output):
            self.get_output(self.keys.all_outputs, state)
            return state.outputs[key]
        elif key == self.keys.face_morpher_input_pose:
            return state.batch[self.indices.original_pose][:, 0:39]
        else:
            raise RuntimeError("Unsupported key: " + key)


def load_face_morpher(file_name: Optional[str] = None, state_dict=None):
    module = SirenFaceMorpher00()
    if state_dict is not None:
        module.load_state_dict(state_dict)
    elif file_name is not None:
        module.load_state_dict(torch_load(file_name))
    return module


def load_body_morpher(file_name: Optional[str] = None, state_dict=None):
    module = SirenMorpher03()
    if state_dict is not None:
        module.load_state_dict(state_dict)
    elif file_name is not None:
        module.load_state_dict(torch_load(file_name))
    return module


def create_poser(
        device: torch.device,
        module_file_names: Optional[Dict[str, str]] = None,
        default_output_index: int = 0,
        state_dicts: Optional[Dict[str, Dict[str, Tensor]]] = None) -> GeneralPoser02:
    """Same signature and defaults as the reference (mode_14.py:134-162) plus in-memory `state_dicts`."""
    if module_file_names is None:
        module_file_names = {}
    if KEY_FACE_MORPHER not in module_file_names:
        module_file_names[KEY_FACE_MORPHER] = "data/character_models/lambda_00/face_morpher.pt"
    if KEY_BODY_MORPHER not in module_file_names:
        module_file_names[KEY_BODY_MORPHER] = "data/character_models/lambda_00/body_morpher.pt"
    sd = state_dicts or {}
    loaders = {
        KEY_FACE_MORPHER: lambda: load_face_morpher(module_file_names[KEY_FACE_MORPHER], sd.get(KEY_FACE_MORPHER)),
        KEY_BODY_MORPHER: lambda: load_body_morpher(module_file_names[KEY_BODY_MORPHER], sd.get(KEY_BODY_MORPHER)),
    }
    return GeneralPoser02(
        image_size=512,
        module_loaders=loaders,
        pose_parameters=get_pose_parameters().get_pose_parameter_groups(),
        output_list_func=TwoStepPoserComputationProtocol().compute_func(),
        subrect=None,
        device=device,
        output_length=5 + 1,
        default_output_index=default_output_index)
