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
    o = c.oracle(np.float32, 1)  # stands in for the GPU statistics kernel on this CPU-only box
    kp = g["kp"][:8]
    q = g["c32_clip_qpos"].reshape(-1, c.tree.nq)[:8]
    # m-phase: frames sharded, 3K+2 numbers all-reduced, closed form applied redundantly on every rank
    lo, hi = parallel.shard_range(len(kp), rank, ws)
    s, z2 = o.m_stats(kp[lo:hi], q[lo:hi])
    s_all, z2_all, T = parallel.allreduce_m_stats(torch.tensor(s), torch.tensor([z2]), hi - lo)
    s_ref, z2_ref = o.m_stats(kp, q)
    assert T == len(kp)
    np.testing.assert_allclose(s_all.numpy(), s_ref, rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(z2_all.numpy()[0], z2_ref, rtol=2e-6)
    # q-phase: clips block-partitioned, no collective on the data path; results gathered clip-major
    C = 5
    clips = torch.arange(C * 3, dtype=torch.float32).reshape(C, 3)
    lo, hi = parallel.shard_range(C, rank, ws)
    full = parallel.allgather_blocks(clips[lo:hi].clone(), C)
    assert torch.equal(full, clips)
    torch.save((s_all, z2_all), Path(tmp) / f"r{rank}.pt")
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])  # every rank holds identical reduced statistics
