"""Import the *unmodified* reference from /root/reference (build container only; it does not exist on the GPU box).

TEST INFRASTRUCTURE.  Used by oracle/make_golden.py and tests/test_oracle_pinned.py to pin the restatement in
oracle/tha4_oracle.py against the reference itself.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('THA4_REFERENCE_ROOT', '/root/reference')
REFERENCE_SRC = os.path.join(REFERENCE_ROOT, 'src')


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, 'tha4'))


def _stub_matplotlib():
    # tha4.shion.base.image_util imports matplotlib.pyplot at import time (absent here); only needed by trainers.
    if 'matplotlib' in sys.modules:
        return
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        m = types.ModuleType('matplotlib')
        m.pyplot = types.ModuleType('matplotlib.pyplot')
        m.cm = types.ModuleType('matplotlib.cm')
        sys.modules.update({'matplotlib': m, 'matplotlib.pyplot': m.pyplot, 'matplotlib.cm': m.cm})


def load():
    """Puts the reference on sys.path and returns its `tha4` package."""
    if not available():
        raise RuntimeError('reference not present at ' + REFERENCE_SRC)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    _stub_matplotlib()
    import tha4
    return tha4


def build_reference_modules(teacher_sds=None, student_sds=None):
    """Instantiate the reference nn.Modules with the hyper-parameters of mode_07.py:137-269 / mode_14.py:93-131
    and load the given state_dicts (strict)."""
    load()
    import contextlib
    import io
    from tha4.poser.modes import mode_07, mode_14
    from tha4.nn.eyebrow_decomposer.eyebrow_decomposer_00 import EyebrowDecomposer00, EyebrowDecomposer00Args
    from tha4.nn.eyebrow_morphing_combiner.eyebrow_morphing_combiner_00 import EyebrowMorphingCombiner00, \
        EyebrowMorphingCombiner00Args
    from tha4.nn.face_morpher.face_morpher_08 import FaceMorpher08, FaceMorpher08Args
    from tha4.nn.nonlinearity_factory import ReLUFactory
    from tha4.nn.normalization import InstanceNorm2dFactory
    from tha4.nn.util import BlockArgs
    from tha4.nn.common.unet import UnetArgs, AttentionBlockArgs
    from tha4.nn.morpher.morpher_00 import Morpher00Args, Morpher00
    from tha4.nn.upscaler.upscaler_02 import Upscaler02Args, Upscaler02

    def ba(inplace):
        return BlockArgs(initialization_method='he', use_spectral_norm=False,
                         normalization_layer_factory=InstanceNorm2dFactory(),
                         nonlinearity_factory=ReLUFactory(inplace=inplace))

    def ua(mc, mults):
        return UnetArgs(in_channels=4, out_channels=7, model_channels=mc, level_channel_multipliers=mults,
                        level_use_attention=[False] * (len(mults) - 1) + [True], num_res_blocks_per_level=1,
                        num_middle_res_blocks=4, time_embedding_channels=None, cond_input_channels=6,
                        cond_internal_channels=256,
                        attention_block_args=AttentionBlockArgs(num_heads=8, use_new_attention_order=True),
                        dropout_prob=0.0)

    out = {}
    if teacher_sds is not None:
        mods = {
            'eyebrow_decomposer': EyebrowDecomposer00(EyebrowDecomposer00Args(128, 4, 64, 16, 6, 512, ba(True))),
            'eyebrow_morphing_combiner': EyebrowMorphingCombiner00(
                EyebrowMorphingCombiner00Args(128, 4, 12, 64, 16, 6, 512, ba(True))),
            'face_morpher': FaceMorpher08(FaceMorpher08Args(192, 4, 27, 64, 24, 6, 512, ba(False), True)),
            'body_morpher': Morpher00(Morpher00Args(256, 4, 6, ua(64, [1, 2, 4, 4, 4]))),
            'upscaler': Upscaler02(Upscaler02Args(512, 4, 6, ua(32, [1, 2, 4, 8, 8, 8]))),
        }
        for k, m in mods.items():
            m.load_state_dict(teacher_sds[k], strict=True)
            m.train(False)
        out['teacher'] = mods
    if student_sds is not None:
        with contextlib.redirect_stdout(io.StringIO()):
            mods = {'face_morpher': mode_14.load_face_morpher(None), 'body_morpher': mode_14.load_body_morpher(None)}
        for k, m in mods.items():
            m.load_state_dict(student_sds[k], strict=True)
            m.train(False)
        out['student'] = mods
    return out


def reference_poser(mode: str, modules: dict):
    """A reference GeneralPoser02 for 'mode_07' | 'mode_12' | 'mode_14' whose loaders return the given modules."""
    load()
    import importlib
    import torch
    from tha4.poser.general_poser_02 import GeneralPoser02
    from tha4.poser.modes.pose_parameters import get_pose_parameters
    m = importlib.import_module('tha4.poser.modes.' + mode)
    if mode == 'mode_07':
        proto, n = m.FiveStepPoserComputationProtocol(2), 33
        names = ['eyebrow_decomposer', 'eyebrow_morphing_combiner', 'face_morpher', 'body_morpher', 'upscaler']
    elif mode == 'mode_12':
        proto, n = m.FiveStepPoserComputationProtocol(2), 18   # (sic) mode_12 reuses the class name, mode_12.py:41
        names = ['eyebrow_decomposer', 'eyebrow_morphing_combiner', 'face_morpher']
    else:
        proto, n = m.TwoStepPoserComputationProtocol(), 6
        names = ['face_morpher', 'body_morpher']
    loaders = {k: (lambda kk=k: modules[kk]) for k in names}
    return GeneralPoser02(image_size=512, module_loaders=loaders,
                          pose_parameters=get_pose_parameters().get_pose_parameter_groups(),
                          output_list_func=proto.compute_func(), subrect=None, device=torch.device('cpu'),
                          output_length=n, default_output_index=0)
