"""BASELINE config 3: full STAC fit (N_ITERS x [q-phase pass, closed-form m-phase] + final q-phase pass) on the rodent.
    python tools/fit_bench.py [n_fit_frames] [clip]       (optionally under torchrun for the clip-split schedule)"""
import contextlib, io, json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import torch.distributed as dist
from stac_mjx_b200 import model, synth
from stac_mjx_b200.config import Cfg
from stac_mjx_b200.stac import Stac

n_fit = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
clip = int(sys.argv[2]) if len(sys.argv) > 2 else 250
rank, ws, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if ws > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tree, cfg = model.load_fixture("rodent")
cfg = Cfg(cfg.to_dict())
cfg.stac.n_frames_per_clip = clip
kpn = list(cfg.model.KEYPOINT_MODEL_PAIRS.keys())
setup = model.make_setup(tree, cfg.model, kpn)
kp, _, off_true = synth.synth_session(tree, setup, n_fit, clip, seed=20260101)
stac = Stac(None, cfg, kpn, tree=tree, device=local)

def run(fn, *a):
    with contextlib.redirect_stdout(io.StringIO()):
        fn(*a)  # warm-up (allocations, first launch)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d = fn(*a)
        torch.cuda.synchronize()
    return d, time.perf_counter() - t0

res = {"n_fit_frames": n_fit, "N_ITERS": int(cfg.model.N_ITERS), "world_size": ws}
free = setup.is_regularized[:, 0] == 0
err0 = float(np.linalg.norm((setup.initial_offsets - off_true)[free], axis=1).mean())
if ws == 1:
    d, t = run(stac.fit_offsets, kp)
    res["reference_schedule"] = {"seconds": t, "frames_per_s": (cfg.model.N_ITERS + 1) * n_fit / t,
                                 "offset_err_mm": 1e3 * float(np.linalg.norm((d.offsets - off_true)[free], axis=1).mean())}
d, t = run(stac.fit_offsets_clip_split, kp, clip)
res["clip_split_schedule"] = {"clip": clip, "seconds": t, "frames_per_s": (cfg.model.N_ITERS + 1) * n_fit / t,
                              "offset_err_mm": 1e3 * float(np.linalg.norm((d.offsets - off_true)[free], axis=1).mean())}
res["initial_offset_err_mm"] = 1e3 * err0
if rank == 0:
    print(json.dumps(res))
if ws > 1:
    dist.barrier(); dist.destroy_process_group()
