"""Tree descriptor, bounds and masks (reference tests/unit/test_controller.py, tests/integration/test_model.py)."""
import numpy as np
import pytest

from stac_mjx_b200 import mjcf, model, tree
from stac_mjx_b200.mjcf import JNT_BALL, JNT_FREE, JNT_HINGE, JNT_SLIDE

from conftest import get_case


def test_align_joint_dims_kat():
    # reference tests/unit/test_controller.py:23-89 (same inputs, same expected bounds / names)
    types = [JNT_FREE, JNT_HINGE, JNT_BALL, JNT_SLIDE]
    ranges = [[0.0, 0.0], [-0.1, 0.1], [0.0, 1.0], [-0.5, 0.5]]
    names = ["root", "hingejoint", "balljoint", "slidejoint"]
    lb, ub, part_names = tree.align_joint_dims(types, ranges, names)
    inf = np.inf
    np.testing.assert_array_equal(lb, np.array([-inf, -inf, -inf, -1, -1, -1, -1, -0.1, 0, 0, 0, 0, -0.5], np.float32))
    np.testing.assert_array_equal(ub, np.array([inf, inf, inf, 1, 1, 1, 1, 0.1, 1, 1, 1, 1, 0.5], np.float32))
    assert part_names == ["root"] * 7 + ["hingejoint"] + ["balljoint"] * 4 + ["slidejoint"]


def test_unconstrained_defaults_and_min_zero():
    lb, ub, _ = tree.align_joint_dims([JNT_HINGE, JNT_SLIDE, JNT_BALL, JNT_HINGE], [[0, 0], [0, 0], [0, 0], [0.2, 0.9]], list("abcd"))
    np.testing.assert_allclose(lb, [-2 * np.pi, -np.inf, -1, -1, -1, -1, 0.0], rtol=1e-7)  # lb = min(lb, 0), stac.py:88
    np.testing.assert_allclose(ub, [2 * np.pi, np.inf, 1, 1, 1, 1, 0.9], rtol=1e-7)


@pytest.mark.parametrize(
    "name,nbody,njnt,nq,K,nact,depth_act,P",
    [
        ("rodent", 67, 68, 74, 23, 31, 12, 5),
        ("mouse", 225, 224, 230, 34, 181, 85, 0),
        ("celegans", 26, 25, 31, 25, 25, 25, 5),
        ("fly_treadmill", 68, 37, 43, 9, 57, 9, 6),
        ("fly_tethered", 68, 37, 43, 30, 49, 9, 6),
        ("synth_data", 2, 1, 7, 1, 1, 1, 1),
    ],
)
def test_compiled_fixture_dimensions(name, nbody, njnt, nq, K, nact, depth_act, P):
    # SURVEY.md section 8 table (computed there from the XML + YAML independently of this reader)
    c = get_case(name)
    t = c.tree
    assert (t.nbody, t.njnt, t.nq, c.K) == (nbody, njnt, nq, K)
    act = t.active_bodies(c.setup.site_bodies)
    assert len(act) == nact and t.body_depth()[act].max() == depth_act
    assert c.setup.indiv_parts.shape == (P, nq)
    assert (t.body_parent[1:] < np.arange(1, nbody)).all()  # DFS pre-order
    np.testing.assert_allclose(np.linalg.norm(t.body_quat, axis=1), 1.0, atol=1e-12)
    hinge = t.jnt_type == JNT_HINGE
    np.testing.assert_allclose(np.linalg.norm(t.jnt_axis[hinge], axis=1), 1.0, atol=1e-12)


def test_rodent_setup_matches_reference_semantics(rodent):
    s, t = rodent.setup, rodent.tree
    assert list(s.indiv_parts.sum(1)) == [11, 11, 6, 6, 7]  # r_leg, l_leg, r_arm, l_arm, head DOFs
    assert s.root_kp_idx == rodent.kp_names.index("SpineL") and s.trunk_kps.sum() == 8
    assert s.is_regularized.sum() == 15 and s.is_regularized[rodent.kp_names.index("HandL")].all()
    assert np.isneginf(s.lb[:3]).all() and (s.lb[3:7] == -1).all() and (s.ub[3:7] == 1).all()
    assert (s.lb <= 0).all()
    np.testing.assert_array_equal(t.qpos0[:7], [0, 0, 0, 1, 0, 0, 0])
    # SCALE_FACTOR scales body offsets below the first top-level body but neither joint anchors nor sites
    j = t.jnt_names.index("atlas")
    np.testing.assert_allclose(t.jnt_pos[j], [-0.02843583549004804, 0, 0])
    b = t.body_names.index("torso")
    np.testing.assert_allclose(t.body_pos[b], 0.9 * np.array([0.03099526054578288, 2.058524937651458e-07, 0.06508957263119967]))
    k = rodent.kp_names.index("ShoulderL")
    np.testing.assert_allclose(t.site_pos[s.site_idxs[k]], [0.0287, 0.00984, -0.02542])
    # nested default classes: vertebra_1_extend inherits pos from class lumbar and range from lumbar_extend
    j = t.jnt_names.index("vertebra_1_extend")
    np.testing.assert_allclose(t.jnt_pos[j], [0.003, 0, -0.003])
    np.testing.assert_allclose(t.jnt_range[j], [-0.5235987755982988, 0.7853981633974483])
    np.testing.assert_allclose(t.jnt_axis[t.jnt_names.index("atlas")], [0, -1, 0])  # axis from class "atlas"


MJCF_FEATURES = """
<mujoco>
  <compiler angle="degree"/>
  <default>
    <joint axis="0 1 0" range="-10 20"/>
    <default class="arm"><joint pos="0.1 0 0"/>
      <default class="arm_twist"><joint axis="2 0 0" range="0 90"/></default>
    </default>
  </default>
  <worldbody>
    <body name="a" pos="1 2 3" euler="0 0 90">
      <freejoint name="root"/>
      <body name="b" pos="0 0 1" childclass="arm">
        <joint name="j1"/>
        <joint name="j2" class="arm_twist"/>
        <site name="s" pos="0 0 0.5"/>
        <body name="c" quat="2 0 0 0"><joint name="j3" type="slide" range="-1 1"/><joint name="j4" type="ball"/></body>
      </body>
    </body>
  </worldbody>
</mujoco>"""


def test_mjcf_defaults_childclass_degrees():
    t = tree.compile_spec(mjcf.parse_mjcf(MJCF_FEATURES, from_string=True))
    assert t.body_names == ["world", "a", "b", "c"] and list(t.body_parent) == [0, 0, 1, 2]
    assert list(t.jnt_type) == [JNT_FREE, JNT_HINGE, JNT_HINGE, JNT_SLIDE, JNT_BALL]
    assert list(t.jnt_qposadr) == [0, 7, 8, 9, 10] and t.nq == 14
    np.testing.assert_allclose(t.body_quat[1], [np.cos(np.pi / 4), 0, 0, np.sin(np.pi / 4)])
    np.testing.assert_allclose(t.qpos0[:7], [1, 2, 3, np.cos(np.pi / 4), 0, 0, np.sin(np.pi / 4)])
    np.testing.assert_allclose(t.qpos0[10:], [1, 0, 0, 0])
    np.testing.assert_allclose(t.jnt_pos[1], [0.1, 0, 0])  # childclass arm
    np.testing.assert_allclose(t.jnt_axis[1], [0, 1, 0])  # inherited from main
    np.testing.assert_allclose(t.jnt_range[1], np.deg2rad([-10, 20]))
    np.testing.assert_allclose(t.jnt_axis[2], [1, 0, 0])  # normalised
    np.testing.assert_allclose(t.jnt_range[2], np.deg2rad([0, 90]))
    np.testing.assert_allclose(t.jnt_range[3], [-1, 1])  # slide ranges are lengths, not angles
    np.testing.assert_allclose(t.body_quat[3], [1, 0, 0, 0])


def test_mjcf_unsupported_elements_fail_loudly():
    bad = "<mujoco><worldbody><body><frame/></body></worldbody></mujoco>"
    with pytest.raises(NotImplementedError):
        mjcf.parse_mjcf(bad, from_string=True)


def test_fixture_roundtrip(rodent):
    t2 = tree.TreeModel.from_dict(rodent.tree.to_dict())
    for k in ("body_pos", "jnt_axis", "qpos0", "site_pos", "body_parent"):
        np.testing.assert_array_equal(getattr(t2, k), getattr(rodent.tree, k))
