"""Compiled kinematic tree ("tree descriptor") for the STAC hot path.

`TreeModel` holds exactly the ``MjModel`` fields the reference reads on the
path (reference ``stac_mjx/stac.py:113-140,219-235``, ``stac_core.py:146``,
MJX ``smooth.kinematics``): body tree, joints, ``qpos0`` and the keypoint
sites.  It is produced either by `compile_spec` from a parsed MJCF
(`mjcf.parse_mjcf`; this image has no ``mujoco``) or by
`TreeModel.from_mjmodel` from a real compiled ``mujoco.MjModel``.

`align_joint_dims` and `part_masks` restate the bound / part-mask logic of
``Stac._align_joint_dims`` (``stac.py:54-88``) and ``Stac.part_opt_setup``
(``stac.py:161-183``) on numpy arrays.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np

from .mjcf import (
    JNT_BALL,
    JNT_FREE,
    JNT_HINGE,
    JNT_QPOS_DIMS,
    JNT_SLIDE,
    BodySpec,
    ModelSpec,
)


@dataclass
class TreeModel:
    """Flat kinematic model. Index conventions follow ``mujoco.MjModel``."""

    body_names: list[str]
    body_parent: np.ndarray  # int32 [nbody]; world's parent is 0
    body_pos: np.ndarray  # float64 [nbody, 3]
    body_quat: np.ndarray  # float64 [nbody, 4] (w, x, y, z)
    body_jntadr: np.ndarray  # int32 [nbody]; -1 when the body has no joint
    body_jntnum: np.ndarray  # int32 [nbody]
    jnt_names: list[str]
    jnt_type: np.ndarray  # int32 [njnt]
    jnt_bodyid: np.ndarray  # int32 [njnt]
    jnt_qposadr: np.ndarray  # int32 [njnt]
    jnt_pos: np.ndarray  # float64 [njnt, 3]
    jnt_axis: np.ndarray  # float64 [njnt, 3]
    jnt_range: np.ndarray  # float64 [njnt, 2]
    qpos0: np.ndarray  # float64 [nq]
    site_names: list[str]
    site_bodyid: np.ndarray  # int32 [nsite]
    site_pos: np.ndarray  # float64 [nsite, 3]
    timestep: float = 0.002

    @property
    def nbody(self) -> int:
        return len(self.body_names)

    @property
    def njnt(self) -> int:
        return len(self.jnt_names)

    @property
    def nq(self) -> int:
        return int(self.qpos0.shape[0])

    @property
    def nsite(self) -> int:
        return len(self.site_names)

    def site_id(self, name: str) -> int:
        """``mujoco.mj_name2id(model, mjOBJ_SITE, name)``."""
        try:
            return self.site_names.index(name)
        except ValueError:
            return -1

    def body_depth(self) -> np.ndarray:
        depth = np.zeros(self.nbody, dtype=np.int32)
        for b in range(1, self.nbody):
            depth[b] = depth[self.body_parent[b]] + 1
        return depth

    def active_bodies(self, site_bodies: Sequence[int]) -> np.ndarray:
        """Bodies (excluding world) whose subtree carries one of ``site_bodies``."""
        act = np.zeros(self.nbody, dtype=bool)
        for b in site_bodies:
            b = int(b)
            while b != 0 and not act[b]:
                act[b] = True
                b = int(self.body_parent[b])
        return np.nonzero(act)[0].astype(np.int32)

    @classmethod
    def from_mjmodel(cls, m) -> "TreeModel":
        """Fill the descriptor from a compiled ``mujoco.MjModel`` (production path)."""
        import mujoco  # noqa: F401  (only available outside this image)

        nb, nj, ns = int(m.nbody), int(m.njnt), int(m.nsite)
        jntadr = np.array(m.body_jntadr, dtype=np.int32)
        return cls(
            body_names=[m.body(i).name for i in range(nb)],
            body_parent=np.array(m.body_parentid, dtype=np.int32),
            body_pos=np.array(m.body_pos, dtype=np.float64),
            body_quat=np.array(m.body_quat, dtype=np.float64),
            body_jntadr=jntadr,
            body_jntnum=np.array(m.body_jntnum, dtype=np.int32),
            jnt_names=[m.joint(i).name for i in range(nj)],
            jnt_type=np.array(m.jnt_type, dtype=np.int32),
            jnt_bodyid=np.array(m.jnt_bodyid, dtype=np.int32),
            jnt_qposadr=np.array(m.jnt_qposadr, dtype=np.int32),
            jnt_pos=np.array(m.jnt_pos, dtype=np.float64),
            jnt_axis=np.array(m.jnt_axis, dtype=np.float64),
            jnt_range=np.array(m.jnt_range, dtype=np.float64),
            qpos0=np.array(m.qpos0, dtype=np.float64),
            site_names=[m.site(i).name for i in range(ns)],
            site_bodyid=np.array(m.site_bodyid, dtype=np.int32),
            site_pos=np.array(m.site_pos, dtype=np.float64),
            timestep=float(m.opt.timestep),
        )

    # -- (de)serialisation used for the committed model fixtures -------------
    def to_dict(self) -> dict:
        out = {}
        for k, v in self.__dict__.items():
            out[k] = v.tolist() if isinstance(v, np.ndarray) else v
        return out

    @classmethod
    def from_dict(cls, d: dict) -> "TreeModel":
        ints = {"body_parent", "body_jntadr", "body_jntnum", "jnt_type", "jnt_bodyid", "jnt_qposadr", "site_bodyid"}
        flt_shapes = {
            "body_pos": (-1, 3),
            "body_quat": (-1, 4),
            "jnt_pos": (-1, 3),
            "jnt_axis": (-1, 3),
            "jnt_range": (-1, 2),
            "qpos0": (-1,),
            "site_pos": (-1, 3),
        }
        kw = {}
        for k, v in d.items():
            if k in ints:
                kw[k] = np.array(v, dtype=np.int32)
            elif k in flt_shapes:
                kw[k] = np.array(v, dtype=np.float64).reshape(flt_shapes[k])
            else:
                kw[k] = v
        return cls(**kw)


def compile_spec(spec: ModelSpec) -> TreeModel:
    """Lay a `ModelSpec` out the way ``MjSpec.compile()`` numbers a model."""
    body_names, parent, bpos, bquat, jntadr, jntnum = [], [], [], [], [], []
    jn, jt, jb, jadr, jpos, jaxis, jrange, qpos0 = [], [], [], [], [], [], [], []
    sn, sb, sp = [], [], []

    def visit(b: BodySpec, pid: int) -> None:
        bid = len(body_names)
        body_names.append(b.name)
        parent.append(pid)
        bpos.append(b.pos)
        bquat.append(b.quat / np.linalg.norm(b.quat))
        jntadr.append(len(jn) if b.joints else -1)
        jntnum.append(len(b.joints))
        for j in b.joints:
            if j.type == JNT_FREE and (pid != 0 or len(b.joints) != 1):
                raise ValueError("free joint must be the only joint of a top-level body")
            jn.append(j.name)
            jt.append(j.type)
            jb.append(bid)
            jadr.append(len(qpos0))
            jpos.append(j.pos)
            jaxis.append(j.axis)
            jrange.append(j.range)
            if j.type == JNT_FREE:
                qpos0.extend(list(b.pos) + list(bquat[bid]))
            elif j.type == JNT_BALL:
                qpos0.extend([1.0, 0.0, 0.0, 0.0])
            else:
                qpos0.append(j.ref)
        for s in b.sites:
            sn.append(s.name)
            sb.append(bid)
            sp.append(s.pos)
        for c in b.children:
            visit(c, bid)

    visit(spec.worldbody, 0)

    def arr(x, shape, dt=np.float64):
        return np.array(x, dtype=dt).reshape(shape)

    return TreeModel(
        body_names=body_names,
        body_parent=arr(parent, (-1,), np.int32),
        body_pos=arr(bpos, (-1, 3)),
        body_quat=arr(bquat, (-1, 4)),
        body_jntadr=arr(jntadr, (-1,), np.int32),
        body_jntnum=arr(jntnum, (-1,), np.int32),
        jnt_names=jn,
        jnt_type=arr(jt, (-1,), np.int32),
        jnt_bodyid=arr(jb, (-1,), np.int32),
        jnt_qposadr=arr(jadr, (-1,), np.int32),
        jnt_pos=arr(jpos, (-1, 3)),
        jnt_axis=arr(jaxis, (-1, 3)),
        jnt_range=arr(jrange, (-1, 2)),
        qpos0=arr(qpos0, (-1,)),
        site_names=sn,
        site_bodyid=arr(sb, (-1,), np.int32),
        site_pos=arr(sp, (-1, 3)),
        timestep=spec.timestep,
    )


# --- bounds and part masks -----------------------------------------------------

_UNCONSTRAINED = {
    JNT_FREE: (np.array([-np.inf] * 3 + [-1.0] * 4), np.array([np.inf] * 3 + [1.0] * 4)),
    JNT_BALL: (-np.ones(4), np.ones(4)),
    JNT_SLIDE: (np.array([-np.inf]), np.array([np.inf])),
    JNT_HINGE: (np.array([-2 * np.pi]), np.array([2 * np.pi])),
}


def align_joint_dims(types, ranges, names) -> tuple[np.ndarray, np.ndarray, list[str]]:
    """Per-qpos bounds and joint names (reference ``stac.py:54-88``).

    Free joints are always unconstrained in translation and [-1, 1] in the
    quaternion; a (0, 0) range means "unconstrained" for the joint type; a
    ranged ball joint broadcasts its range over all four components; and the
    lower bound is finally clamped with ``min(lb, 0)`` (``stac.py:88``).
    Returned as float32, the dtype the reference's jax arrays carry.
    """
    lb, ub, part_names = [], [], []
    for t, rng, name in zip(types, ranges, names):
        t = int(t)
        dims = JNT_QPOS_DIMS[t]
        if t == JNT_FREE:
            lo, hi = _UNCONSTRAINED[t]
        else:
            lo, hi = rng
            if lo == 0 and hi == 0:
                lo, hi = _UNCONSTRAINED[t]
            lo, hi = lo * np.ones(dims), hi * np.ones(dims)
        lb.append(np.asarray(lo, dtype=np.float64))
        ub.append(np.asarray(hi, dtype=np.float64))
        part_names += [name] * dims
    lb = np.minimum(np.concatenate(lb), 0.0).astype(np.float32)
    ub = np.concatenate(ub).astype(np.float32)
    return lb, ub, part_names


def part_masks(part_names: Sequence[str], parts: dict | None) -> np.ndarray:
    """Boolean [P, nq] masks by substring match (reference ``stac.py:161-183``)."""
    if not parts:
        return np.zeros((0, len(part_names)), dtype=bool)
    rows = [[any(p in name for p in plist) for name in part_names] for plist in parts.values()]
    return np.array(rows, dtype=bool).reshape(len(rows), len(part_names))


__all__ = [
    "TreeModel",
    "compile_spec",
    "align_joint_dims",
    "part_masks",
    "JNT_FREE",
    "JNT_BALL",
    "JNT_SLIDE",
    "JNT_HINGE",
]
