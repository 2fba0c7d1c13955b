"""The 45-parameter pose schema (12 eyebrow + 27 face + 6 rotation/breathing), as in
src/tha4/poser/modes/pose_parameters.py:4-35: same group names, order, arities, ranges and defaults."""
from tha4_b200.poser.poser import PoseParameterCategory as C
from tha4_b200.poser.poser import PoseParameters

_GROUPS = [
    # (name, category, arity, default, range)
    ('eyebrow_troubled', C.EYEBROW, 2), ('eyebrow_angry', C.EYEBROW, 2), ('eyebrow_lowered', C.EYEBROW, 2),
    ('eyebrow_raised', C.EYEBROW, 2), ('eyebrow_happy', C.EYEBROW, 2), ('eyebrow_serious', C.EYEBROW, 2),
    ('eye_wink', C.EYE, 2), ('eye_happy_wink', C.EYE, 2), ('eye_surprised', C.EYE, 2), ('eye_relaxed', C.EYE, 2),
    ('eye_unimpressed', C.EYE, 2), ('eye_raised_lower_eyelid', C.EYE, 2),
    ('iris_small', C.IRIS_MORPH, 2),
    ('mouth_aaa', C.MOUTH, 1, 1.0), ('mouth_iii', C.MOUTH, 1), ('mouth_uuu', C.MOUTH, 1), ('mouth_eee', C.MOUTH, 1),
    ('mouth_ooo', C.MOUTH, 1), ('mouth_delta', C.MOUTH, 1), ('mouth_lowered_corner', C.MOUTH, 2),
    ('mouth_raised_corner', C.MOUTH, 2), ('mouth_smirk', C.MOUTH, 1),
    ('iris_rotation_x', C.IRIS_ROTATION, 1, 0.0, (-1.0, 1.0)), ('iris_rotation_y', C.IRIS_ROTATION, 1, 0.0, (-1.0, 1.0)),
    ('head_x', C.FACE_ROTATION, 1, 0.0, (-1.0, 1.0)), ('head_y', C.FACE_ROTATION, 1, 0.0, (-1.0, 1.0)),
    ('neck_z', C.FACE_ROTATION, 1, 0.0, (-1.0, 1.0)),
    ('body_y', C.BODY_ROTATION, 1, 0.0, (-1.0, 1.0)), ('body_z', C.BODY_ROTATION, 1, 0.0, (-1.0, 1.0)),
    ('breathing', C.BREATHING, 1, 0.0, (0.0, 1.0)),
]


def get_pose_parameters() -> PoseParameters:
    b = PoseParameters.Builder()
    for g in _GROUPS:
        name, cat, arity = g[0], g[1], g[2]
        default = g[3] if len(g) > 3 else 0.0
        rng = g[4] if len(g) > 4 else None
        b.add_parameter_group(name, cat, arity=arity, default_value=default, range=rng)
    return b.build()
