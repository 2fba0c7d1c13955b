"""Host-side logic that needs no GPU: pose schema, module state_dict layout, poser API surface, no-fallback rule."""
import os

import pytest
import torch

from tha4_b200._lib import Tha4Error
from tha4_b200.nn import state_dict_spec as spec
from tha4_b200.poser.modes import mode_07, mode_12, mode_14
from tha4_b200.poser.modes.pose_parameters import get_pose_parameters
from tha4_b200.poser.poser import PoseParameterCategory, Poser


def test_pose_parameter_schema():
    pp = get_pose_parameters()
    assert pp.get_parameter_count() == 45
    groups = pp.get_pose_parameter_groups()
    assert len(groups) == 30
    assert sum(g.get_arity() for g in groups if g.get_category() == PoseParameterCategory.EYEBROW) == 12
    assert pp.get_parameter_index('eye_wink_left') == 12
    assert pp.get_parameter_index('iris_rotation_x') == 37
    assert pp.get_parameter_index('breathing') == 44
    assert pp.get_parameter_name(26) == 'mouth_aaa'
    assert groups[13].get_default_value() == 1.0                       # mouth_aaa (pose_parameters.py:19)
    assert groups[22].get_range() == (-1.0, 1.0) and groups[0].get_range() == (0.0, 1.0)


@pytest.mark.parametrize('mode,n_out', [(mode_07, 33), (mode_12, 18), (mode_14, 6)])
def test_create_poser_surface(mode, n_out):
    poser = mode.create_poser(torch.device('cuda:0'))
    assert isinstance(poser, Poser)
    assert poser.get_image_size() == 512 and poser.get_num_parameters() == 45
    assert poser.get_output_length() == n_out and poser.get_dtype() == torch.float
    assert len(poser.get_pose_parameter_groups()) == 30


def test_module_state_dict_layout_and_param_counts(teacher_sds, student_sds):
    expected = {'eyebrow_decomposer': 31479434, 'eyebrow_morphing_combiner': 31535878, 'face_morpher': 31605002,
                'body_morpher': 34682119, 'upscaler': 35015655}
    for name, cls in mode_07._CLASSES.items():
        m = cls()
        sd = m.state_dict()
        assert list(sd.keys()) == list(teacher_sds[name].keys())
        assert sum(v.numel() for v in sd.values()) == expected[name]
        m.load_state_dict(teacher_sds[name], strict=True)
    face = mode_14.load_face_morpher(None, student_sds['face_morpher'])
    body = mode_14.load_body_morpher(None, student_sds['body_morpher'])
    assert sum(p.numel() for p in face.parameters()) == 121476
    assert sum(p.numel() for p in body.parameters()) == 331567


def test_shipped_student_checkpoints_load(lambda00_sds):
    mode_14.load_face_morpher(None, lambda00_sds['face_morpher'])
    mode_14.load_body_morpher(None, lambda00_sds['body_morpher'])


def test_reference_zero_init_roles():
    m = mode_07._CLASSES['body_morpher']()
    sd = m.state_dict()
    assert float(sd['body.last.2.weight'].abs().max()) == 0.0                            # unet.py:529
    assert float(sd['body.down_blocks.0.res_blocks.0.conv1.weight'].abs().max()) == 0.0  # unet.py:142
    assert float(sd['body.first_conv.weight'].abs().max()) > 0.0


def test_no_cpu_fallback():
    m = mode_14.load_face_morpher(None)
    with pytest.raises(Tha4Error):
        m(torch.zeros(1, 39))
    if not torch.cuda.is_available():
        poser = mode_14.create_poser(torch.device('cuda:0'))
        with pytest.raises(Exception):
            poser.pose(torch.zeros(4, 512, 512), torch.zeros(45))


def test_product_does_not_import_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tha4_b200')
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith('.py'):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, os.path.join(dirpath, f)


def test_general_poser_02_host_logic(monkeypatch):
    """GeneralPoser02 semantics that need no GPU (general_poser_02.py:41-98): lazy module construction, rank-3 / rank-1
    promotion, subrect crop, default output index, free(), to().  The library context is replaced by a stub."""
    from tha4_b200.poser import general_poser_02 as gp

    class FakeCtx:
        def __init__(self, device): self.device = device

    class FakeModule(torch.nn.Module):
        built = 0
        def __init__(self): super().__init__(); FakeModule.built += 1; self.ctx = None
        def attach_context(self, ctx): self.ctx = ctx

    monkeypatch.setattr(gp, 'Context', FakeCtx)
    seen = {}

    def pipeline(state):
        image, pose = state.batch
        seen.update(image=image, pose=pose, modules=state.modules, ctx=state.context)
        return [image.mean(dim=(1, 2, 3)), pose.sum(dim=1), image]

    poser = gp.GeneralPoser02(module_loaders={'a': FakeModule, 'b': FakeModule}, device=torch.device('cpu'), output_length=3,
                              pose_parameters=get_pose_parameters().get_pose_parameter_groups(), output_list_func=pipeline,
                              subrect=((2, 6), (1, 5)), default_output_index=1, image_size=8)
    assert FakeModule.built == 0 and poser.get_num_parameters() == 45 and poser.get_image_size() == 8 and poser.get_output_length() == 3
    out = poser.pose(torch.arange(4 * 8 * 8, dtype=torch.float).view(4, 8, 8), torch.ones(45))
    assert FakeModule.built == 2 and out.shape == (1,) and out.item() == 45.0              # default_output_index = 1
    assert seen['image'].shape == (1, 4, 4, 4) and seen['pose'].shape == (1, 45)             # promoted + cropped
    assert torch.equal(seen['image'][0, 0], torch.arange(64, dtype=torch.float).view(8, 8)[2:6, 1:5])
    assert all(m.ctx is seen['ctx'] and not m.training for m in seen['modules'].values())
    assert poser.pose(torch.zeros(2, 4, 8, 8), torch.zeros(2, 45), 2).shape == (2, 4, 4, 4) and FakeModule.built == 2
    poser.free()
    poser.get_modules()
    assert FakeModule.built == 4                                                              # rebuilt after free()
    assert poser.to(torch.device('cpu')) is poser and FakeModule.built == 4                   # same device: nothing dropped
