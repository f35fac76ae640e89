"""A small synthetic model exercising every joint type (free, hinge, slide, ball) and 2-3 joints per body."""
import numpy as np

from stac_mjx_b200 import mjcf, tree

MIXED_XML = """
<mujoco>
  <compiler angle="radian"/>
  <worldbody>
    <body name="base" pos="0.1 0.2 0.3" quat="0.9 0.1 -0.2 0.3">
      <freejoint name="root"/>
      <site name="m_base" pos="0.02 0.01 0.03"/>
      <body name="arm" pos="0.1 0 0.02" quat="0.8 0.2 0.1 -0.1">
        <joint name="a_slide" type="slide" axis="0.2 1 0.1" pos="0.01 0 0" range="-0.05 0.08"/>
        <joint name="a_hinge" type="hinge" axis="0 0.6 0.8" pos="0 0.01 0.02" range="-1 1.2"/>
        <site name="m_arm" pos="0.05 0.01 0"/>
        <body name="wrist" pos="0.08 0.01 0">
          <joint name="w_h1" type="hinge" axis="1 0 0" range="-0.8 0.8"/>
          <joint name="w_h2" type="hinge" axis="0 1 0" pos="0.01 0 0" range="-0.9 0.7"/>
          <joint name="w_ball" type="ball" pos="0 0.005 0.01"/>
          <site name="m_wrist" pos="0.03 0 0.01"/>
          <body name="tip" pos="0.04 0 0.01">
            <joint name="t_h" type="hinge" axis="0 0 1" ref="0.2" range="-1.5 1.5"/>
            <site name="m_tip" pos="0.02 0.02 0"/>
          </body>
        </body>
      </body>
      <body name="leg" pos="-0.05 0.03 -0.02">
        <joint name="l_ball" type="ball" pos="0.005 0 0"/>
        <site name="m_leg" pos="0 0.04 -0.03"/>
        <body name="foot" pos="0 0.05 -0.06">
          <joint name="f_slide" type="slide" axis="0 0 1" range="-0.02 0.02"/>
          <site name="m_foot" pos="0.01 0 -0.01"/>
        </body>
      </body>
    </body>
    <body name="loose" pos="0.5 0.5 0.1">
      <joint name="loose_free" type="free"/>
      <site name="m_loose" pos="0.01 0.02 0.03"/>
    </body>
  </worldbody>
</mujoco>"""

KP = ["m_base", "m_arm", "m_wrist", "m_tip", "m_leg", "m_foot", "m_loose"]


def mixed_tree():
    t = tree.compile_spec(mjcf.parse_mjcf(MIXED_XML, from_string=True))
    site_idxs = np.array([t.site_id(n) for n in KP], dtype=np.int32)
    lb, ub, _ = tree.align_joint_dims(t.jnt_type, t.jnt_range, t.jnt_names)
    return t, site_idxs, lb, ub


def random_qpos(t, rng, n):
    q = np.tile(t.qpos0, (n, 1)) + rng.normal(scale=0.15, size=(n, t.nq))
    return q
