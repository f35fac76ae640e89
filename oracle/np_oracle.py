"""Independent (torch, reverse-mode autodiff) restatement used to validate stac_oracle.c.

TEST INFRASTRUCTURE ONLY.  The reference differentiates ``q_loss`` with JAX
reverse-mode AD (``jaxopt`` calls ``jax.value_and_grad``); the C oracle and the
CUDA kernels use an analytic Jacobian-transpose gradient.  This module rebuilds
the forward pass op-by-op from the MJX formulas with torch tensors and lets
``torch.autograd`` produce the gradient, so the analytic forms are checked
against the same kind of machinery the reference uses.  It also restates the
jaxopt 0.8.5 ProjectedGradient loop in plain Python for small cases.

Pure-Python loops: small cases only.
"""

from __future__ import annotations

import math

import numpy as np
import torch

JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3


def rotate(v, q):
    s, u = q[0], q[1:]
    r = 2 * (torch.dot(u, v) * u) + (s * s - torch.dot(u, u)) * v
    return r + 2 * s * torch.linalg.cross(u, v)


def quat_mul(u, v):
    return torch.stack(
        [
            u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3],
            u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2],
            u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1],
            u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0],
        ]
    )


def normalize(x):
    n = torch.linalg.norm(x)
    return x / (n + 1e-6 * (n == 0.0))


def axis_angle_to_quat(axis, angle):
    s, c = torch.sin(angle * 0.5), torch.cos(angle * 0.5)
    return torch.cat([c.reshape(1), axis * s])


class TorchModel:
    def __init__(self, tree, site_bodies, dtype=torch.float64):
        self.t = tree
        self.dtype = dtype
        f = lambda a: torch.tensor(np.asarray(a, dtype=np.float32).astype(np.float64), dtype=dtype)
        self.body_pos, self.body_quat = f(tree.body_pos), f(tree.body_quat)
        self.jnt_pos, self.jnt_axis, self.qpos0 = f(tree.jnt_pos), f(tree.jnt_axis), f(tree.qpos0)
        self.site_bodies = [int(b) for b in site_bodies]

    def kinematics(self, qpos):
        """MJX smooth.kinematics: returns (qpos_normalised, xpos list, xquat list)."""
        t = self.t
        xpos = [torch.zeros(3, dtype=self.dtype)]
        xquat = [torch.tensor([1.0, 0, 0, 0], dtype=self.dtype)]
        qout = qpos.clone()
        for b in range(1, t.nbody):
            p = int(t.body_parent[b])
            pos, quat = self.body_pos[b], self.body_quat[b]
            pos = xpos[p] + rotate(pos, xquat[p])
            quat = quat_mul(xquat[p], quat)
            for jj in range(int(t.body_jntnum[b])):
                j = int(t.body_jntadr[b]) + jj
                adr, typ = int(t.jnt_qposadr[j]), int(t.jnt_type[j])
                if typ == JNT_FREE:
                    pos = qpos[adr : adr + 3]
                    quat = normalize(qpos[adr + 3 : adr + 7])
                    qout = torch.cat([qout[: adr + 3], quat, qout[adr + 7 :]])
                    continue
                anchor = rotate(self.jnt_pos[j], quat) + pos
                if typ == JNT_BALL:
                    qloc = normalize(qpos[adr : adr + 4])
                    qout = torch.cat([qout[:adr], qloc, qout[adr + 4 :]])
                    quat = quat_mul(quat, qloc)
                    pos = anchor - rotate(self.jnt_pos[j], quat)
                elif typ == JNT_HINGE:
                    qloc = axis_angle_to_quat(self.jnt_axis[j], qpos[adr] - self.qpos0[adr])
                    quat = quat_mul(quat, qloc)
                    pos = anchor - rotate(self.jnt_pos[j], quat)
                else:
                    pos = pos + rotate(self.jnt_axis[j], quat) * (qpos[adr] - self.qpos0[adr])
            xpos.append(pos)
            xquat.append(quat)
        return qout, xpos, xquat

    def site_xpos(self, xpos, xquat, site_pos):
        return torch.stack([xpos[b] + rotate(site_pos[k], xquat[b]) for k, b in enumerate(self.site_bodies)])

    def q_loss(self, q, kp, qs_to_opt, kps_to_opt, initial_q, site_pos):
        """reference stac_core.py:27-63"""
        qfull = (1 - qs_to_opt) * initial_q + qs_to_opt * q
        _, xpos, xquat = self.kinematics(qfull)
        markers = self.site_xpos(xpos, xquat, site_pos).flatten()
        residual = (kp - markers) * kps_to_opt
        return torch.sum(torch.square(residual))

    def loss_grad(self, q, q0, qmask, kp, kpmask, site_pos):
        tt = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), dtype=self.dtype)
        qv = tt(q).requires_grad_(True)
        loss = self.q_loss(qv, tt(kp), tt(qmask), tt(kpmask), tt(q0), tt(site_pos))
        (g,) = torch.autograd.grad(loss, qv)
        return float(loss.detach()), g.numpy()

    def projected_gradient(self, q0, lb, ub, qmask, kp, kpmask, site_pos, tol, maxiter=400, maxls=15):
        """jaxopt 0.8.5 ProximalGradient._update_accel / _ls / _error with prox = box clip."""
        lb, ub = np.asarray(lb, np.float64), np.asarray(ub, np.float64)
        eps = float(np.finfo(np.float64 if self.dtype == torch.float64 else np.float32).eps)
        fun = lambda v: self.loss_grad(v, q0, qmask, kp, kpmask, site_pos)
        x = np.asarray(q0, np.float64).copy()
        y, t, step, err, it, nls = x.copy(), 1.0, 1.0, math.inf, 0, 0
        while True:
            fy, gy = fun(y)
            st, halv = step, 0
            while True:
                xn = np.clip(y - st * gy, lb, ub)
                fn, _ = fun(xn)
                nls += 1
                d = xn - y
                if not (st * (fn - fy) > st * np.dot(d, gy) + 0.5 * np.dot(d, d) + eps) or halv >= maxls:
                    break
                st *= 0.5
                halv += 1
            step = 1.0 if st <= 1e-6 else st / 0.5
            tn = 0.5 * (1 + math.sqrt(1 + 4 * t * t))
            y = xn + ((t - 1) / tn) * (xn - x)
            _, gn = fun(xn)
            err = float(np.linalg.norm(np.clip(xn - gn, lb, ub) - xn))
            x, t, it = xn, tn, it + 1
            if not (err > tol and it < maxiter):
                break
        return x, err, it, nls
