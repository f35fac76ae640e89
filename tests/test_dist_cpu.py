"""world_size-2 gloo tests of the multi-GPU host logic (clip partition, m-phase all-reduce, all-gather)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, tmp):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from conftest import get_case
    from stac_mjx_b200 import parallel

    c = get_case("rodent")
    g = np.load(ROOT / "tests" / "golden" / "rodent.npz")
    o = c.oracle(np.float32, 1)  # stands in for the GPU statistics kernels on this CPU-only box
    kp = g["kp"][:8]
    q = g["c32_clip_qpos"].reshape(-1, c.tree.nq)[:8]

    class OracleEngine:
        """CPU stand-in with the Engine methods `stac_core._m_opt` uses."""

        K = c.K

        def f32(self, a, shape=None):
            return torch.as_tensor(np.asarray(a, dtype=np.float32))

        def m_stats_buffer(self, kp_, q_):
            s_, z2_ = o.m_stats(kp_.numpy(), q_.numpy())
            return torch.tensor(np.concatenate([s_.reshape(-1), [z2_], [float(len(kp_))]]).astype(np.float32))

        def m_residual(self, kp_, q_, m_):
            return torch.tensor([o.m_residual(kp_.numpy(), q_.numpy(), m_.numpy())], dtype=torch.float32)

    from stac_mjx_b200 import stac_core

    mdl = stac_core.StacModel(engine=OracleEngine(), site_pos=torch.zeros(c.K, 3))
    reg = c.setup.is_regularized
    # m-phase: frames sharded, the 3K+2 statistics all-reduced in place, closed form applied redundantly on every rank,
    # then the 1-float residual all-reduced
    lo, hi = parallel.shard_range(len(kp), rank, ws)
    res = stac_core._m_opt(mdl, None, kp[lo:hi], q[lo:hi], c.setup.initial_offsets, reg, 1.0, reduce_fn=parallel.allreduce_m_stats)
    ref = stac_core._m_opt(mdl, None, kp, q, c.setup.initial_offsets, reg, 1.0)
    np.testing.assert_allclose(res.params.numpy(), ref.params.numpy(), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(float(res.error), float(ref.error), rtol=1e-5)
    mo, eo = o.m_opt(kp, q, c.setup.initial_offsets, reg, 1.0)  # the reference's expanded closed form
    np.testing.assert_allclose(ref.params.numpy(), mo, rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(float(ref.error), float(eo), rtol=2e-3)
    s_all, z2_all = res.params, res.error.reshape(1)
    # q-phase: clips block-partitioned, no collective on the data path; results gathered clip-major
    C = 5
    clips = torch.arange(C * 3, dtype=torch.float32).reshape(C, 3)
    lo, hi = parallel.shard_range(C, rank, ws)
    full = parallel.allgather_blocks(clips[lo:hi].clone(), C)
    assert torch.equal(full, clips)
    # fewer units than ranks: a rank with an empty block takes part in the collectives all the same
    one = torch.arange(4, dtype=torch.float32).reshape(1, 4)
    lo, hi = parallel.shard_range(1, rank, ws)
    assert (hi - lo) == (1 if rank == 0 else 0)
    assert torch.equal(parallel.allgather_blocks(one[lo:hi].clone(), 1), one)
    torch.save((s_all, z2_all), Path(tmp) / f"r{rank}.pt")
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])  # every rank holds identical reduced statistics
