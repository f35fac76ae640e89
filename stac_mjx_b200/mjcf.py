"""Minimal MJCF reader: kinematic tree only.

The reference obtains its kinematic tree from MuJoCo (``mujoco.MjSpec.from_file``
-> ``add_site`` -> ``rescale.dm_scale_spec`` -> ``spec.compile()``;
reference ``stac_mjx/stac.py:185-235`` and ``stac_mjx/rescale.py:6-46``).  The
STAC hot path only ever reads the *kinematic* part of the compiled model
(body tree, joints, site offsets), so this module parses exactly that subset
of MJCF and lays it out the way MuJoCo's compiler does:

* body ids in depth-first pre-order, world = 0;
* joints / sites numbered body by body, in document order inside a body;
* ``qpos`` addresses in joint order (free 7, ball 4, slide/hinge 1);
* nested ``<default class>`` inheritance and ``childclass`` resolution;
* ``compiler angle`` (degree is MuJoCo's default), ``eulerseq``;
* axes and quaternions normalised, everything kept in float64 (MuJoCo's
  ``mjtNum``) until the descriptor is cast to float32 for the device.

When the real ``mujoco`` package is importable, `tree.TreeModel.from_mjmodel`
fills the same structure from a compiled ``MjModel`` instead.
"""

from __future__ import annotations

import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

# mujoco.mjtJoint values (the reference indexes dicts with these enums,
# stac_mjx/stac.py:27-52).
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
JNT_QPOS_DIMS = {JNT_FREE: 7, JNT_BALL: 4, JNT_SLIDE: 1, JNT_HINGE: 1}
_JNT_TYPE_BY_NAME = {"free": JNT_FREE, "ball": JNT_BALL, "slide": JNT_SLIDE, "hinge": JNT_HINGE}

_JOINT_DEFAULT_KEYS = ("type", "pos", "axis", "range", "ref", "limited")


def _floats(text: str) -> np.ndarray:
    return np.array([float(t) for t in text.split()], dtype=np.float64)


def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Hamilton product, [w, x, y, z]."""
    return np.array(
        [
            a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
            a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
            a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
            a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
        ]
    )


def _axisangle_quat(axis: np.ndarray, angle: float) -> np.ndarray:
    n = np.linalg.norm(axis)
    if n < 1e-14:
        return np.array([1.0, 0.0, 0.0, 0.0])
    axis = axis / n
    return np.concatenate([[math.cos(angle / 2)], axis * math.sin(angle / 2)])


def _mat_quat(m: np.ndarray) -> np.ndarray:
    """Rotation matrix (columns = frame axes) to unit quaternion."""
    t = np.trace(m)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = [0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s]
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = math.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = [(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s]
    elif m[1, 1] > m[2, 2]:
        s = math.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = [(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s]
    else:
        s = math.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = [(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s]
    q = np.array(q)
    return q / np.linalg.norm(q)


@dataclass
class JointSpec:
    name: str
    type: int
    pos: np.ndarray
    axis: np.ndarray
    range: np.ndarray  # (2,) in radians for hinge
    ref: float


@dataclass
class SiteSpec:
    name: str
    pos: np.ndarray


@dataclass
class BodySpec:
    name: str
    pos: np.ndarray
    quat: np.ndarray
    joints: list[JointSpec] = field(default_factory=list)
    sites: list[SiteSpec] = field(default_factory=list)
    children: list["BodySpec"] = field(default_factory=list)

    def add_site(self, name: str, pos) -> None:
        """Append a site after the body's existing ones (mujoco ``body.add_site``)."""
        self.sites.append(SiteSpec(name, np.asarray(pos, dtype=np.float64).copy()))


@dataclass
class ModelSpec:
    """Editable kinematic spec (the subset of ``mujoco.MjSpec`` the path uses)."""

    worldbody: BodySpec
    timestep: float = 0.002

    def body(self, name: str) -> BodySpec:
        stack = [self.worldbody]
        while stack:
            b = stack.pop()
            if b.name == name:
                return b
            stack.extend(b.children)
        raise KeyError(f"no body named {name!r} in model")

    def scale(self, s: float) -> "ModelSpec":
        """Restatement of ``rescale.dm_scale_spec`` for the kinematic fields.

        Reference ``stac_mjx/rescale.py:21-46``: the recursion starts at the
        *first top-level body* and scales the ``pos`` of every body below it;
        that body itself, its siblings, joint anchors and site offsets are left
        untouched.  Geoms, meshes, actuators and keyframes do not enter the
        kinematics.  Mutates and returns self (the reference copies the spec
        first; callers here build a fresh spec per compile).
        """

        def rec(parent: BodySpec) -> None:
            for child in parent.children:
                child.pos = child.pos * s
                rec(child)

        if self.worldbody.children:
            rec(self.worldbody.children[0])
        return self


class _Defaults:
    """Default-class table for <joint> attributes with nested inheritance."""

    def __init__(self) -> None:
        self.joint: dict[str, dict[str, str]] = {"main": {}}

    def load(self, root: ET.Element) -> None:
        for top in root.findall("default"):
            self._walk(top, self.joint["main"], top.get("class") or "main")

    def _walk(self, el: ET.Element, parent_attrs: dict[str, str], cls: str) -> None:
        attrs = dict(parent_attrs)
        for j in el.findall("joint"):
            for k in _JOINT_DEFAULT_KEYS:
                if j.get(k) is not None:
                    attrs[k] = j.get(k)
        self.joint[cls] = attrs
        for sub in el.findall("default"):
            if sub.get("class") is None:
                raise ValueError("nested <default> needs a class name")
            self._walk(sub, attrs, sub.get("class"))


def _orientation(el: ET.Element, degrees: bool, eulerseq: str) -> np.ndarray:
    if el.get("quat") is not None:
        q = _floats(el.get("quat"))
        return q / np.linalg.norm(q)
    if el.get("euler") is not None:
        e = _floats(el.get("euler"))
        if degrees:
            e = np.deg2rad(e)
        q = np.array([1.0, 0.0, 0.0, 0.0])
        for ang, ax in zip(e, eulerseq):
            axis = np.zeros(3)
            axis["xyz".index(ax.lower())] = 1.0
            r = _axisangle_quat(axis, ang)
            # lower case: intrinsic (rotating frame, post-multiply); upper: extrinsic
            q = quat_mul(q, r) if ax.islower() else quat_mul(r, q)
        return q / np.linalg.norm(q)
    if el.get("axisangle") is not None:
        a = _floats(el.get("axisangle"))
        ang = math.radians(a[3]) if degrees else a[3]
        return _axisangle_quat(a[:3], ang)
    if el.get("xyaxes") is not None:
        a = _floats(el.get("xyaxes"))
        x = a[:3] / np.linalg.norm(a[:3])
        y = a[3:] - x * np.dot(x, a[3:])
        y = y / np.linalg.norm(y)
        z = np.cross(x, y)
        return _mat_quat(np.stack([x, y, z], axis=1))
    if el.get("zaxis") is not None:
        z = _floats(el.get("zaxis"))
        z = z / np.linalg.norm(z)
        z0 = np.array([0.0, 0.0, 1.0])
        axis = np.cross(z0, z)
        s = np.linalg.norm(axis)
        ang = math.atan2(s, float(np.dot(z0, z)))
        if s < 1e-10:
            axis = np.array([1.0, 0.0, 0.0])
        return _axisangle_quat(axis, ang)
    return np.array([1.0, 0.0, 0.0, 0.0])


def parse_mjcf(source: str | Path, *, from_string: bool = False) -> ModelSpec:
    """Parse the kinematic subset of an MJCF file (or XML string)."""
    root = ET.fromstring(source) if from_string else ET.parse(str(source)).getroot()
    if root.tag != "mujoco":
        raise ValueError("not an MJCF document (root element must be <mujoco>)")

    degrees, eulerseq = True, "xyz"  # MuJoCo defaults
    for comp in root.findall("compiler"):
        if comp.get("angle") is not None:
            degrees = comp.get("angle") == "degree"
        if comp.get("eulerseq") is not None:
            eulerseq = comp.get("eulerseq")
        if comp.get("coordinate", "local") != "local":
            raise NotImplementedError("compiler coordinate='global' is not supported")

    timestep = 0.002
    for opt in root.findall("option"):
        if opt.get("timestep") is not None:
            timestep = float(opt.get("timestep"))

    defaults = _Defaults()
    defaults.load(root)

    anon = {"body": 0, "joint": 0, "site": 0}

    def make_joint(el: ET.Element, childclass: str | None) -> JointSpec:
        cls = el.get("class") or childclass or "main"
        if cls not in defaults.joint:
            raise ValueError(f"unknown default class {cls!r}")
        attr = dict(defaults.joint[cls])
        for k in _JOINT_DEFAULT_KEYS:
            if el.get(k) is not None:
                attr[k] = el.get(k)
        if el.tag == "freejoint":
            jtype = JNT_FREE
        else:
            jtype = _JNT_TYPE_BY_NAME[attr.get("type", "hinge")]
        pos = _floats(attr["pos"]) if "pos" in attr else np.zeros(3)
        axis = _floats(attr["axis"]) if "axis" in attr else np.array([0.0, 0.0, 1.0])
        rng = _floats(attr["range"]) if "range" in attr else np.zeros(2)
        ref = float(attr.get("ref", 0.0))
        if jtype == JNT_FREE:
            pos, axis, rng, ref = np.zeros(3), np.array([0.0, 0.0, 1.0]), np.zeros(2), 0.0
        if degrees and jtype in (JNT_HINGE, JNT_BALL):
            rng = np.deg2rad(rng)
            if jtype == JNT_HINGE:
                ref = math.radians(ref)
        n = np.linalg.norm(axis)
        if n > 0:
            axis = axis / n
        name = el.get("name")
        if name is None:
            name = ""
            anon["joint"] += 1
        return JointSpec(name, jtype, pos, axis, rng, ref)

    def make_body(el: ET.Element, childclass: str | None) -> BodySpec:
        childclass = el.get("childclass") or childclass
        name = el.get("name")
        if name is None:
            name = ""
            anon["body"] += 1
        pos = _floats(el.get("pos")) if el.get("pos") is not None else np.zeros(3)
        body = BodySpec(name, pos, _orientation(el, degrees, eulerseq))
        for ch in el:
            if ch.tag in ("joint", "freejoint"):
                body.joints.append(make_joint(ch, childclass))
            elif ch.tag == "site":
                spos = _floats(ch.get("pos")) if ch.get("pos") is not None else np.zeros(3)
                body.sites.append(SiteSpec(ch.get("name") or "", spos))
            elif ch.tag == "body":
                body.children.append(make_body(ch, childclass))
            elif ch.tag in ("frame", "replicate", "include", "composite", "flexcomp", "attach"):
                raise NotImplementedError(f"<{ch.tag}> inside <body> is not supported by this reader")
        return body

    wb = root.find("worldbody")
    if wb is None:
        raise ValueError("MJCF has no <worldbody>")
    world = BodySpec("world", np.zeros(3), np.array([1.0, 0.0, 0.0, 0.0]))
    for ch in wb:
        if ch.tag == "body":
            world.children.append(make_body(ch, None))
        elif ch.tag == "site":
            spos = _floats(ch.get("pos")) if ch.get("pos") is not None else np.zeros(3)
            world.sites.append(SiteSpec(ch.get("name") or "", spos))
        elif ch.tag in ("frame", "replicate", "include", "composite", "flexcomp", "attach"):
            raise NotImplementedError(f"<{ch.tag}> inside <worldbody> is not supported by this reader")
    return ModelSpec(world, timestep)
