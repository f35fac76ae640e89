"""Random hinge trees with welded (jointless) bodies, for the element-folding logic of the register-resident path."""
import numpy as np

from stac_mjx_b200 import mjcf, tree


def random_tree(seed, n_bodies=45, p_welded=0.45, max_hinges=3, n_sites=14, free_root=True):
    """MJCF with a free (or hinged) root, `n_bodies` bodies of which a fraction carries no joint, 1..max_hinges hinges on
    the others, and keypoint sites scattered over jointed AND welded bodies.  Returns (TreeModel, site_idxs, lb, ub)."""
    rng = np.random.default_rng(seed)
    parent = [-1] + [int(rng.integers(max(0, i - 6), i)) for i in range(1, n_bodies)]
    children = {i: [] for i in range(n_bodies)}
    for i in range(1, n_bodies):
        children[parent[i]].append(i)
    site_bodies = set(int(b) for b in rng.choice(n_bodies, size=min(n_sites, n_bodies), replace=False))
    names = []

    def fmt(v):
        return " ".join(f"{x:.6g}" for x in v)

    def emit(i, depth):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        pos = rng.normal(scale=0.05, size=3)
        pad = "  " * (depth + 2)
        s = f'{pad}<body name="b{i}" pos="{fmt(pos)}" quat="{fmt(q)}">\n'
        if i == 0 and free_root:
            s += f'{pad}  <freejoint name="root"/>\n'
        elif i == 0 or rng.random() > p_welded:
            for j in range(int(rng.integers(1, max_hinges + 1))):
                ax = rng.normal(size=3)
                ax /= np.linalg.norm(ax)
                jp = rng.normal(scale=0.01, size=3) if rng.random() < 0.5 else np.zeros(3)
                s += (f'{pad}  <joint name="j{i}_{j}" type="hinge" axis="{fmt(ax)}" pos="{fmt(jp)}" '
                      f'range="{-rng.uniform(0.3, 1.5):.4f} {rng.uniform(0.3, 1.5):.4f}"/>\n')
        if i in site_bodies:
            names.append(f"m{i}")
            s += f'{pad}  <site name="m{i}" pos="{fmt(rng.normal(scale=0.02, size=3))}"/>\n'
        for c in children[i]:
            s += emit(c, depth + 1)
        return s + f"{pad}</body>\n"

    xml = '<mujoco>\n  <compiler angle="radian"/>\n  <worldbody>\n' + emit(0, 0) + "  </worldbody>\n</mujoco>"
    t = tree.compile_spec(mjcf.parse_mjcf(xml, from_string=True))
    site_idxs = np.array([t.site_id(n) for n in names], dtype=np.int32)
    lb, ub, _ = tree.align_joint_dims(t.jnt_type, t.jnt_range, t.jnt_names)
    return t, site_idxs, lb, ub


def n_active(t, site_bodies):
    act = np.zeros(t.nbody, bool)
    for b in site_bodies:
        while b != 0 and not act[b]:
            act[b] = True
            b = t.body_parent[b]
    return int(act.sum()), int((act & (np.asarray(t.body_jntnum) > 0)).sum())
