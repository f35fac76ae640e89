"""Benchmark of the STAC IK hot path: frames/s on the rodent rat23 session (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the fused q-phase (root optimisation on the first frame of every clip, then
1 + P FISTA solves per frame) over one synthetic session of 18 000 frames cut into 72 clips of 250 frames.
With N > 1 (launched by torchrun) every rank runs its own session of the same size: clips are independent,
there is no data-path collective, `scaling` is "weak" and `value` is the total frames of all ranks divided by
the slowest rank's time.  Prints ONE JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "IK frames/sec (rodent rat23, whole box) at 1/2/4/8 B200; marker RMSE parity"  # BASELINE.json:metric
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="rodent")
    ap.add_argument("--frames", type=int, default=18000, help="frames per session (per rank)")
    ap.add_argument("--clip", type=int, default=250)
    ap.add_argument("--seed", type=int, default=20260101)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU work of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def load_case(name):
    from stac_mjx_b200 import model

    tree, cfg = model.load_fixture(name)
    kp_names = list(cfg.model.KEYPOINT_MODEL_PAIRS.keys())
    setup = model.make_setup(tree, cfg.model, kp_names)
    return tree, cfg, setup


def root_kw(tree, cfg, setup):
    has_root = setup.root_kp_idx >= 0 and int(tree.jnt_type[0]) in (0, 2)
    return dict(do_root=1 if has_root else 0, root_kp_idx=setup.root_kp_idx, trunk_kps=setup.trunk_kps,
                root_dims=4 if int(tree.jnt_type[0]) == 2 else 7, tol=float(cfg.model.FTOL), maxiter=int(cfg.model.N_ITER_Q))  # fmt: skip


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (a port of the reference algorithm; the JAX reference cannot be installed offline)
# ------------------------------------------------------------------------------------------------------
def cpu_arm(args, tree, cfg, setup, kp, seconds):
    """Time the CPU restatement (oracle, MJX operation order, float32) with all host threads on a bounded sample."""
    from oracle.oracle import Oracle

    cores = os.cpu_count() or 1
    o = Oracle(tree, setup.site_bodies, np.float32, 0)
    F = args.clip
    kw = root_kw(tree, cfg, setup)
    # calibrate on a few frames, then size the sample: one clip chain per thread, equal length
    probe = kp[: min(8, F)][None]
    t0 = time.perf_counter()
    o.pose_clips(probe, tree.qpos0, setup.initial_offsets, setup.lb, setup.ub, setup.indiv_parts, nthreads=1, **kw)
    per_frame = (time.perf_counter() - t0) / probe.shape[1]
    n_clips = min(cores, kp.shape[0] // F)
    frames = int(max(4, min(F, seconds / max(per_frame, 1e-6))))
    sample = np.ascontiguousarray(kp[: n_clips * F].reshape(n_clips, F, -1)[:, :frames])
    t0 = time.perf_counter()
    ref = o.pose_clips(sample, tree.qpos0, setup.initial_offsets, setup.lb, setup.ub, setup.indiv_parts, nthreads=n_clips, **kw)
    dt = time.perf_counter() - t0
    return {
        "_sites": ref["sites"], "_frames": frames, "_clips": n_clips,
        "value": n_clips * frames / dt,
        "unit": UNIT,
        "cores": n_clips,
        "kind": "port",
        "sample": f"{n_clips} clips x first {frames} frames of the same session, one OpenMP thread per clip, "
        f"oracle/stac_oracle.c float32 MJX-order (host has {cores} logical cores); the JAX reference is not installable offline",
        "seconds": dt,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from stac_mjx_b200 import synth

    tree, cfg, setup = load_case(args.model)
    cores = os.cpu_count() or 1
    n_frames = min(args.frames, cores * args.clip)
    kp, _, _ = synth.synth_session(tree, setup, n_frames, args.clip, seed=args.seed)
    vals = []
    per_step = max(2.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        r = cpu_arm(args, tree, cfg, setup, kp, per_step)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    frames_per_step = vals[-1]["value"] * vals[-1]["seconds"]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean([r["seconds"] for r in vals])), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model}.xml rat23 synthetic {args.frames}-frame session per GPU in {args.clip}-frame clips "
                               f"({args.frames // args.clip} independent chains), root optimisation + {1 + setup.indiv_parts.shape[0]} FISTA solves per frame, "
                               f"FTOL {float(cfg.model.FTOL):g}, N_ITER_Q {int(cfg.model.N_ITER_Q)}",
                   "sample": f"each step times a bounded sample of {int(frames_per_step)} frames of that workload on the host CPU",
                   "clip_frames": args.clip},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": vals[-1]["cores"], "kind": "port", "sample": vals[-1]["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()  # fmt: skip
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows)}  # fmt: skip


def run_ours(args):
    import torch
    import torch.distributed as dist

    from stac_mjx_b200 import flops, synth
    from stac_mjx_b200.engine import Engine

    rank, ws, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if ws > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    tree, cfg, setup = load_case(args.model)
    C, F = args.frames // args.clip, args.clip
    kp, _, _ = synth.synth_session(tree, setup, C * F, F, seed=args.seed + rank)  # every rank: its own session (weak scaling)
    eng = Engine(tree, setup.site_bodies, local)
    kw = root_kw(tree, cfg, setup)
    P = setup.indiv_parts.shape[0]
    K = len(setup.site_idxs)

    kp_host = torch.from_numpy(kp.reshape(C, F, -1)).pin_memory()
    kp_dev = kp_host.to(dev)
    q0 = torch.tensor(np.tile(tree.qpos0.astype(np.float32), (C, 1)), device=dev)
    out = {}
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def step_device():
        qio = q0.clone()
        return eng.pose_clips(kp_dev, qio, setup.initial_offsets, setup.lb, setup.ub, setup.indiv_parts, out=out, **kw)

    # e2e: the call a user makes -- Stac.ik_only(kp_data, offsets) on HOST arrays: pinned staging + H2D of the step's
    # keypoints, the fused kernel, D2H of every result into host memory and the reference's StacData packing
    import contextlib
    import io as _io

    from stac_mjx_b200.config import Cfg
    from stac_mjx_b200.stac import Stac

    ecfg = Cfg(cfg.to_dict())
    ecfg.stac.n_frames_per_clip, ecfg.stac.continuous = F, False
    kp_names = list(cfg.model.KEYPOINT_MODEL_PAIRS.keys())
    with contextlib.redirect_stdout(_io.StringIO()):
        stac = Stac(None, ecfg, kp_names, tree=tree, device=local)
    e2e_bytes = {}

    def step_e2e():
        with contextlib.redirect_stdout(_io.StringIO()):
            d = stac.ik_only(kp, setup.initial_offsets)
        e2e_bytes["d2h"] = d.qpos.nbytes + d.xpos.nbytes + d.xquat.nbytes + d.marker_sites.nbytes + C * F * 4
        return d

    def sync_all():
        torch.cuda.synchronize(dev)
        if ws > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step_device()
    sync_all()

    fp32_peak = eng.fma_peak_tflops()
    sampler = ClockSampler(local)
    if rank == 0:  # one sampler per job is enough for the clocks line (rank 0's GPU)
        sampler.start()

    def timed(fn, steps):
        evs = []
        sync_all()
        for _ in range(steps):
            flush.fill_(1.0)  # evict L2 between timed iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        sync_all()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if ws > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_dev = timed(step_device, args.steps)
    iters, ls = int(out["iters"].sum().item()) + int(out["root_stats"][:, [0, 2]].sum().item()), int(out["ls_evals"].sum().item()) + int(out["root_stats"][:, [1, 3]].sum().item())
    status_bad = int((out["status"] != 0).sum().item())
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    sampler.stop_flag.set()
    if rank == 0:
        sampler.join(timeout=2)

    frames_all = ws * C * F
    value = frames_all * args.steps / (ms_dev * 1e-3)
    e2e_value = frames_all * args.steps / (ms_e2e * 1e-3)
    pc = flops.path_cost(tree, setup.site_bodies)
    flop_step = pc.total(iters, ls, C * F, 1 + P)  # this rank's step
    kernel_s = ms_dev * 1e-3 / args.steps
    achieved = flop_step / kernel_s / 1e12
    hbm_step = flops.hbm_bytes_per_frame(tree, K, 1 + P) * C * F
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this workload, from the committed ncu capture
    tfile = ROOT / "profiles" / "traffic_r1f_bench_launch.csv"
    if tfile.exists() and args.model == "rodent" and C * F == 18000:
        try:
            import csv

            vals = {r[12]: float(r[14]) for r in csv.reader(tfile.open()) if len(r) > 14 and r[12].startswith("dram__bytes")}
            traffic = vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]
        except Exception:
            traffic = None
    h2d = kp_host.numel() * 4
    d2h = int(e2e_bytes.get("d2h", 0))

    cpu = None
    if rank == 0 and ws == 1 and not args.no_cpu_baseline:
        cpu = cpu_arm(args, tree, cfg, setup, kp, args.cpu_seconds)
        cpu.pop("seconds", None)
        # marker RMSE parity (the checker role of the oracle): GPU marker sites of the sampled clips vs (a) the CPU port in
        # MJX operation order just timed -- two float32 orders, so this is the algorithm's rounding noise floor -- and
        # (b) the canonical-order oracle on a small subset, which the kernels reproduce bit for bit
        from oracle.oracle import Oracle

        nc, nf = cpu.pop("_clips"), cpu.pop("_frames")
        gpu_sites = out["sites"][:nc, :nf].cpu().numpy()
        rmse_mjx = float(np.sqrt(np.mean(np.sum((gpu_sites - cpu.pop("_sites")) ** 2, axis=-1))))
        sub = np.ascontiguousarray(kp.reshape(C, F, -1)[:2, :20])
        can = Oracle(tree, setup.site_bodies, np.float32, 1).pose_clips(sub, tree.qpos0, setup.initial_offsets, setup.lb, setup.ub,
                                                                        setup.indiv_parts, nthreads=2, **kw)
        d_can = out["sites"][:2, :20].cpu().numpy() - can["sites"]
        parity = {"marker_rmse_m_vs_cpu_port_mjx_order": rmse_mjx, "sample": f"{nc} clips x {nf} frames",
                  "marker_rmse_m_vs_canonical_oracle": float(np.sqrt(np.mean(np.sum(d_can**2, axis=-1)))),
                  "max_abs_qpos_diff_vs_canonical_oracle": float(np.abs(out["qpos"][:2, :20].cpu().numpy() - can["qpos"]).max()),
                  "canonical_sample": "2 clips x 20 frames", "tolerance_m": 1e-4}
    else:
        parity = None

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": f"{args.model}.xml rat23 synthetic {C * F}-frame session per GPU in {F}-frame clips ({C} independent chains), "
                            f"root optimisation + {1 + P} FISTA solves per frame, FTOL {kw['tol']:g}, N_ITER_Q {kw['maxiter']}",
                "frames_per_gpu": C * F, "clips_per_gpu": C, "clip_frames": F, "parallelism": f"clips sharded over {ws} GPU(s), no data-path collective",
                "l2": "256 MB buffer written between timed iterations (L2 flushed)",
                "iters_per_frame": iters / (C * F), "ls_evals_per_iter": ls / max(iters, 1), "nonfinite_clips": status_bad,
            },
            "roofline": {
                "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak if fp32_peak else None,
                "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write, profiles/traffic_r1f_bench_launch.csv; outputs mostly still in L2)",
                "note": "FP32 CUDA-core bound path (no tensor-core or HBM-bound kernel exists on it): algorithmic flops (stac_mjx_b200/flops.py, "
                        "SURVEY 8(d)) of one launch / CUDA-event duration; peak = FFMA throughput measured in this run by stacb_fma_peak "
                        "(not in MEASURED_PEAKS.json, which holds HBM and bf16 tensor peaks only)",
                "flop_per_launch": flop_step,
                "hbm": {"achieved_GBs": hbm_step / kernel_s / 1e9, "peak_GBs": peaks.get("hbm_gbs"), "algorithmic_bytes_per_launch": hbm_step,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "absent"},
            },
            "cpu_baseline": cpu, "parity": parity,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": args.steps,  # one fused stacb::pose_clips_kernel launch per step
            "clocks": sampler.summary(),
        }  # fmt: skip
        print(json.dumps(line), flush=True)
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
