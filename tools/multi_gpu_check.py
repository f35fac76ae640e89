"""Multi-GPU check (run under torchrun on >= 2 GPUs):
  * Stac.ik_only with clips block-partitioned over ranks + all-gather == the single-GPU result, bit for bit;
  * Stac.fit_offsets with the m-phase statistics all-reduced over ranks == single-GPU offsets to float32 rounding.
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py
"""
import os, sys, io, contextlib
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import torch.distributed as dist
from stac_mjx_b200 import model, synth, parallel
from stac_mjx_b200.config import Cfg
from stac_mjx_b200.stac import Stac

rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
tree, cfg = model.load_fixture("rodent")
cfg = Cfg(cfg.to_dict())
F, C = 8, 5
cfg.stac.n_frames_per_clip, cfg.stac.continuous, cfg.model.N_ITERS = F, False, 2
kpn = list(cfg.model.KEYPOINT_MODEL_PAIRS.keys())
setup = model.make_setup(tree, cfg.model, kpn)
kp, _, _ = synth.synth_session(tree, setup, C * F, F, seed=3)

def quiet(fn, *a):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a)

stac = Stac(None, cfg, kpn, tree=tree, device=local)
single_ik = quiet(stac.ik_only, kp, setup.initial_offsets)          # no process group yet: every rank runs all clips
single_fit = quiet(stac.fit_offsets, kp[:F * 2])
ccfg = Cfg(cfg.to_dict()); ccfg.stac.continuous = True; ccfg.stac.n_frames_per_clip = 10   # overlapping clips (4 x (10 + 10 look-ahead)) + device epilogues
cstac = Stac(None, ccfg, kpn, tree=tree, device=local)
single_c = quiet(lambda: cstac.ik_only(kp, setup.initial_offsets, edge_effects=True, infer_qvels=True))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
multi_ik = quiet(stac.ik_only, kp, setup.initial_offsets)           # clips sharded over ranks, results all-gathered
multi_fit = quiet(stac.fit_offsets, kp[:F * 2])                     # q-phase replicated, m-phase frames sharded + all-reduce
multi_c = quiet(lambda: cstac.ik_only(kp, setup.initial_offsets, edge_effects=True, infer_qvels=True))
ok = True
for k in ("qpos", "xpos", "marker_sites", "qvel", "kp_data"):
    same = np.array_equal(getattr(single_c, k), getattr(multi_c, k))
    ok &= same
    if rank == 0:
        print(f"continuous ik_only + epilogues {k}: sharded == single-GPU bitwise: {same}")
for k in ("qpos", "xpos", "xquat", "marker_sites"):
    same = np.array_equal(getattr(single_ik, k), getattr(multi_ik, k))
    ok &= same
    if rank == 0:
        print(f"ik_only {k}: sharded == single-GPU bitwise: {same}")
d_off = float(np.abs(single_fit.offsets - multi_fit.offsets).max())
d_site = float(np.abs(single_fit.marker_sites - multi_fit.marker_sites).max())
ok &= d_off < 1e-5
t = torch.tensor([multi_fit.offsets.astype(np.float64).sum()], device="cuda")
lst = [torch.zeros_like(t) for _ in range(ws)]
dist.all_gather(lst, t)
consistent = all(float(x) == float(lst[0]) for x in lst)
ok &= consistent
if rank == 0:
    print(f"fit_offsets: max |offsets(all-reduced) - offsets(single)| = {d_off:.3e} m, marker sites differ by {d_site:.3e} m; "
          f"offsets identical on every rank: {consistent}")
    print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", f"world_size={ws}")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
