#!/usr/bin/env python
"""Benchmark of the apply_rectwv_coeff hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle)

A "step" is one pass of the hot path over one batch of synthetic frames per GPU
(default: BASELINE config 2 — 64 float32 2048x2048 frames, grism H + filter H,
resampling 2, one shared RectWaveCoeff).  Ranks own independent batches (weak
scaling, no collective on the data path).  Rank 0 prints ONE JSON line.
"""

import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "rectified+wavecal 2048x2048 frames/sec"
UNIT = "frames/s"
FALLBACK_HBM_GBS = 6650.0        # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="2", help="BASELINE config number (1,2,3,4,4b,5)")
    ap.add_argument("--frames", type=int, default=64, help="frames per GPU per step")
    ap.add_argument("--resampling", type=int, default=2)
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"])
    ap.add_argument("--g", type=int, default=0, help="frames per CTA (tuning)")
    ap.add_argument("--tile-k", type=int, default=0)
    ap.add_argument("--stages", type=int, default=0, help="rows in flight per CTA (tuning)")
    ap.add_argument("--combine", default="none", choices=["none", "mean", "gather", "abba", "abba_median"],
                    help="combined product per step: mean image (torch sum + reduce), gather,\n"
                         "the ABBA A-B mean combination (own kernels + reduce) or the ABBA\n"
                         "median combination (row-slab exchange + stack-median kernel; at most\n"
                         "64 frames per nodding position over all ranks)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-frames", type=int, default=3,
                    help="frames of the CPU-baseline sample (0 disables it)")
    ap.add_argument("--ref-budget-s", type=float, default=200.0)
    ap.add_argument("--no-collective", action="store_true",
                    help="N > 1: skip the timed combined products (collective block)")
    return ap.parse_args()


def config_key(text):
    return int(text) if text.isdigit() else text


# --------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [l for (t, l) in self.lines if t0 <= t <= t1 + 0.15] or [l for _, l in self.lines]
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for row in rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax),
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------- CPU arms
def _oracle_one(args):
    """Worker: the oracle on one frame restricted to some slitlets (faithful rows)."""
    seed, config, resampling, keep = args
    import copy

    from oracle.pipeline import apply_rectwv_coeff as oracle_apply
    from pyemir_b200 import synthetic
    coeff = _cached_coeff(config)
    frame = _cached_frames(config, seed)
    c = copy.copy(coeff)
    c.missing_slitlets = sorted(set(coeff.missing_slitlets)
                                | {i for i in range(1, 56) if i not in keep})
    hdul = synthetic.make_hdulist(frame, c)
    t0 = time.perf_counter()
    oracle_apply(hdul, c, args_resampling=resampling,
                 nmax_order=5 if config == 3 else 4)
    return time.perf_counter() - t0


_COEFF = {}
_FRAMES = {}


def _cached_coeff(config):
    from pyemir_b200 import synthetic
    if config not in _COEFF:
        _COEFF[config] = synthetic.make_config(config)
    return _COEFF[config]


def _cached_frames(config, seed):
    from pyemir_b200 import synthetic
    key = (config, seed)
    if key not in _FRAMES:
        _FRAMES[key] = synthetic.make_frames(_cached_coeff(config), 1, seed=seed)[0]
    return _FRAMES[key]


def cpu_baseline_single(config, resampling, nframes):
    """Oracle, one process, one thread, whole frames (mirrors how the reference runs)."""
    times = [_oracle_one((1000 + i, config, resampling, tuple(range(1, 56))))
             for i in range(nframes)]
    total = sum(times)
    return {"value": nframes / total, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{nframes} whole frames of this workload through the float64 CPU "
                      f"oracle (restatement of pyemir+numina, all bounding-box rows), "
                      f"1 process / 1 thread, {total:.1f} s"}


def run_reference_arm(args):
    """--impl reference: the CPU oracle on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    config = config_key(args.config)
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    per_slitlet_s = 0.11
    nsl = int(args.ref_budget_s / ((args.steps + args.warmup) * per_slitlet_s))
    nsl = max(5, min(55, nsl))
    _cached_coeff(config)
    for w in range(workers):
        _cached_frames(config, 2000 + w)        # generated before the fork, shared
    ctx = mp.get_context("fork")
    step_times = []
    with ctx.Pool(workers) as pool:
        for step in range(args.warmup + args.steps):
            first = (step * nsl) % 55
            keep = tuple(((first + k) % 55) + 1 for k in range(nsl))
            jobs = [(2000 + w, config, args.resampling, keep) for w in range(workers)]
            t0 = time.perf_counter()
            pool.map(_oracle_one, jobs, chunksize=1)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                step_times.append(dt)
    total = sum(step_times)
    frames_equiv = workers * (nsl / 55.0) * len(step_times)
    value = frames_equiv / total
    sample = (f"each step: {workers} worker processes x 1 frame x {nsl} of 55 slitlets "
              f"through the float64 CPU oracle (restatement of pyemir+numina; the real "
              f"reference needs numina+astropy, absent here)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(step_times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, config),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, config):
    from pyemir_b200 import synthetic
    cfg = synthetic.BASELINE_CONFIGS[config]
    out = {
        "workload": (f"BASELINE config {config}: {args.frames} synthetic 2048x2048 float32 "
                     f"frames per GPU per step, grism {cfg['grism']} + filter "
                     f"{cfg['filter_name']}, 55 slitlets, ttd order {cfg['order']}, "
                     f"resampling {args.resampling} (flux-preserving area overlap), one "
                     f"shared RectWaveCoeff"),
        "frames_per_gpu_per_step": args.frames,
        "resampling": args.resampling,
        "l2_policy": "inputs larger than L2 (batch in+out = %.2f GB per step vs 126 MB L2)"
                     % (args.frames * (16.78 + 28.42) / 1e3),
        "combine": args.combine,
    }
    return out


def tuning_of(info):
    """Launch parameters of the hot kernel (not part of the workload: kept out of `config`)."""
    return {"frames_per_cta": info["frames_per_cta"], "threads": info["threads"],
            "tile_k": info["tile_k"], "stages": info["stages"], "footprint": info["footprint"],
            "smem_bytes": info["smem_bytes"], "ctas_per_sm": info["ctas_per_sm"],
            "window_rows": info["window_rows"], "window_cols": info["window_cols"],
            "table_MB": info["table_bytes"] / 1e6,
            "library": os.environ.get("EMIRWV_LIB", "libemirwv.so"),
            "label": os.environ.get("AB_LABEL")}


def e2e_recipe_entries(args, plan, h_in, a_out, a_jmm, F, world, dev, barrier, d2h_sparse):
    """End to end through the entries that follow the recipe further than the path itself:
    RAW uint16 frames in (half the bytes over PCIe; bad-pixel mask, bias, dark and flat applied
    on the device in front of the path), and the whole ABBA product (only the combined A - B
    image comes back).  Same frames/s metric; both results are checked against the
    device-resident kernels before they are timed."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from pyemir_b200 import abba
    from pyemir_b200.prereduce import prereduce_torch

    rng = np.random.default_rng(77)
    raw = torch.empty(h_in.shape, dtype=torch.uint16, pin_memory=True)
    raw_np = raw.numpy()
    raw_np[...] = np.clip(h_in.numpy() * 8.0 + 1000.0, 0, 65535).astype(np.uint16)
    bias = rng.normal(1000.0, 2.0, size=raw_np.shape[1:]).astype(np.float32)
    dark = rng.normal(2.0, 0.5, size=raw_np.shape[1:]).astype(np.float32)
    flat = rng.normal(8.0, 0.2, size=raw_np.shape[1:]).astype(np.float32)
    bpm = (rng.random(raw_np.shape[1:]) < 0.005).astype(np.uint8)
    plan.set_calibration(bpm=bpm, bias=bias, dark=dark, flat=flat)
    labels = abba.labels_of(abba.build_full_set("ABBA", 1, (F + 3) // 4)[:F])
    offsets = [0.0 if i % 4 in (0, 3) else 0.3127 for i in range(F)]
    img = torch.empty(plan.out_shape, dtype=torch.float32, pin_memory=True)

    # -- checks against the device-resident kernels (untimed)
    d_red = prereduce_torch(raw.to(dev), bias, dark, flat, bpm=bpm)
    d_prod, d_jm = plan.run_torch(d_red)
    plan.run_host_raw(raw_np, a_out, a_jmm)
    ok_raw = bool(torch.equal(torch.from_numpy(a_out[F - 1]), d_prod[F - 1].cpu())
                  and torch.equal(torch.from_numpy(a_jmm), d_jm.cpu()))
    want = abba.combine_abba(d_prod, abba.build_full_set("ABBA", 1, (F + 3) // 4)[:F], offsets,
                             "mean").cpu()
    plan.abba_host(raw_np, labels, offsets, "mean", raw=True, out=img.numpy())
    ok_abba = bool(torch.equal(img, want))
    del d_red, d_prod, want
    if not (ok_raw and ok_abba):
        raise SystemExit(f"bench.py: host entries differ from the device path (raw {ok_raw}, abba {ok_abba})")

    def timed(call):
        call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            call()
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return world * F * args.e2e_steps / float(te.item())

    v_raw = timed(lambda: plan.run_host_raw(raw_np, a_out, a_jmm))
    v_abba = timed(lambda: plan.abba_host(raw_np, labels, offsets, "mean", raw=True, out=img.numpy()))
    v_abba_f32 = timed(lambda: plan.abba_host(h_in.numpy(), labels, offsets, "mean", out=img.numpy()))
    plan.set_calibration()
    return {
        "raw_u16": {"value": v_raw, "h2d_bytes_per_step": int(raw_np.nbytes),
                    "d2h_bytes_per_step": d2h_sparse,
                    "call": "emirwv_run_host_raw: RAW uint16 frames in pinned memory -> bad-pixel "
                            "mask, bias, dark, flat (k_prereduce, k_bpm_fix) -> k_rectwv -> the whole "
                            "products delivered like e2e.value (zeros written by host threads); "
                            "identical to the device-resident kernels: %s" % ok_raw},
        "abba_product": {"value": v_abba, "h2d_bytes_per_step": int(raw_np.nbytes),
                         "d2h_bytes_per_step": int(img.numpy().nbytes),
                         "float32_reduced_frames_in": v_abba_f32,
                         "call": "emirwv_abba_host(mean, EMIRWV_HOST_RAW): RAW uint16 frames in, flow, "
                                 "rectification, per-frame vertical offsets (half of the frames "
                                 "shifted), float64 sums in frame order, float32(A) - float32(B): "
                                 "only the combined image comes back; identical to the "
                                 "device-resident kernels: %s" % ok_abba},
    }


def collective_block(args, plan, d_in, d_out, d_jmm, F, world, rank, dev):
    """N > 1: what a combined product costs (BASELINE config 5: an ABBA sequence sharded over the
    ranks with an exchange over NVLink at the end), timed per sequence with CUDA events, max over
    ranks; the headline `value` stays the collective-free rectification."""
    import statistics as st

    import torch
    import torch.distributed as dist

    from pyemir_b200 import abba
    from pyemir_b200.distributed import gather_frames

    npix = plan.out_shape[0] * plan.out_shape[1]
    esz = d_out.element_size()

    def run(fn, reps=3):
        fn()
        times = []
        for _ in range(reps):
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t.item()))
        return st.median(times)

    out = {}
    # (i) mean: every rank rectifies its share of a >= 1024-frame sequence in rounds of F frames,
    # adds them to its float64 A and B sums (k_abba_partial), ONE reduce of the two sums, finish
    rounds = max(1, -(-1024 // (F * world)))
    labels = abba.labels_of(abba.build_full_set("ABBA", 1, (F + 3) // 4)[:F])
    na = world * rounds * labels.count(1)
    nb = world * rounds * labels.count(-1)
    state = {}

    def local_sums():
        sums = None
        for _ in range(rounds):
            plan.run_torch(d_in, d_out, d_jmm)
            sums = abba.partial_sums(d_out, labels, sums=sums)
        state["sums"] = sums

    def mean_all():
        local_sums()
        red = abba.reduce_sums(state["sums"], 0)
        if red is not None:
            state["ref"] = abba.finish(red, na, nb)

    t_local = run(local_sums)
    t_all = run(mean_all)
    red_bytes = 2 * npix * 8
    out["abba_mean"] = {
        "frames_per_sequence": F * world * rounds, "ms_per_sequence": t_all,
        "ms_rectify_and_partial_sums": t_local, "ms_reduce_and_finish": max(t_all - t_local, 0.0),
        "frames_per_s": F * world * rounds / (t_all / 1e3),
        "bytes_reduced_per_rank": red_bytes,
        "nvlink_GBs_per_rank": red_bytes / max(t_all - t_local, 1e-6) / 1e6,
        "nvlink_peak_GBs_per_direction": 900.0,
        "how": "k_rectwv + k_abba_partial per round of %d frames per rank, one NCCL reduce(sum) of "
               "the float64 A and B sums (2 x %.1f MB per rank) to rank 0, k_abba_finish" % (F, npix * 8 / 1e6)}

    # (i') the same product with the exchange fused into the last summing pass over NVLink peer
    # memory (pyemir_b200/peer.py): no collective call on the data path
    try:
        from pyemir_b200.peer import PeerWorkspace, combine_abba_mean_fused
        ws = PeerWorkspace(npix)

        def mean_fused():
            sums = None
            for _ in range(rounds - 1):
                plan.run_torch(d_in, d_out, d_jmm)
                sums = abba.partial_sums(d_out, labels, sums=sums)
            plan.run_torch(d_in, d_out, d_jmm)
            state["fused"] = combine_abba_mean_fused(d_out, labels, na, nb, ws, 0, sums=sums)

        t_fused = run(mean_fused)
        torch.cuda.synchronize()                 # (every rank made the same calls up to here)
        same = bool(torch.equal(state["fused"], state["ref"])) if rank == 0 else None
        timed_out = ws.timed_out()
        ws.close()
        out["abba_mean_fused"] = {
            "frames_per_sequence": F * world * rounds, "ms_per_sequence": t_fused,
            "ms_exchange_and_finish": max(t_fused - t_local, 0.0),
            "frames_per_s": F * world * rounds / (t_fused / 1e3),
            "bytes_pushed_per_rank": int(red_bytes * (world - 1) / world),
            "identical_to_the_nccl_path": same, "timed_out": timed_out,
            "how": "the last k_abba_partial pass of every rank pushes its float64 sums straight into "
                   "the receive buffer of the rank that owns the pixel (stores over NVLink, CUDA IPC "
                   "peer memory), system-scope flags, every owner finishes its slab and stores it "
                   "into the image on rank 0: no collective call"}
    except Exception as exc:                                      # noqa: BLE001
        out["abba_mean_fused"] = {"error": repr(exc)[:300]}

    # (ii) median: 64 images per nodding position over all ranks; row-slab exchange (grouped
    # send/recv), per-slab k_abba_median, gather of the slabs
    per_rank = max(4, (128 // world) // 4 * 4)
    if per_rank <= F:
        lab_all = abba.labels_of(abba.build_full_set("ABBA", 1, per_rank * world // 4))
        counts = [per_rank] * world
        sub = d_out[:per_rank]

        def median_all():
            plan.run_torch(d_in[:per_rank], sub, d_jmm[:per_rank])
            abba.combine_abba_median(sub, lab_all, counts, 0)

        def rect_only():
            plan.run_torch(d_in[:per_rank], sub, d_jmm[:per_rank])

        t_rect = run(rect_only)
        t_med = run(median_all)
        sent = per_rank * npix * esz * (world - 1) / world
        out["abba_median"] = {
            "frames_per_sequence": per_rank * world, "ms_per_sequence": t_med,
            "ms_rectify": t_rect, "ms_exchange_median_gather": max(t_med - t_rect, 0.0),
            "frames_per_s": per_rank * world / (t_med / 1e3),
            "bytes_sent_per_rank": int(sent + npix * 4 / world),
            "nvlink_GBs_per_rank": sent / max(t_med - t_rect, 1e-6) / 1e6,
            "nvlink_peak_GBs_per_direction": 900.0,
            "how": "row-slab exchange (every rank sends 1/world of each of its frames to every "
                   "peer), k_abba_median on the slab, gather of the float32 slabs on rank 0; the "
                   "stack-median kernels take at most 64 images per nodding position"}

    # (iii) plain gather of the rectified products on rank 0
    counts = [F] * world

    def gather_all():
        gather_frames(d_out, counts, 0)

    try:
        t_g = run(gather_all, reps=2)
        into_root = (world - 1) * F * npix * esz
        out["gather"] = {"frames": F * world, "ms": t_g, "bytes_into_root": into_root,
                         "nvlink_GBs_into_root": into_root / t_g / 1e6,
                         "nvlink_peak_GBs_per_direction": 900.0,
                         "how": "grouped NCCL send/recv of every rank's %d products to rank 0" % F}
    except RuntimeError as exc:           # root out of memory for the gathered stack
        out["gather"] = {"error": repr(exc)[:200]}
    limiter = ("mean: the per-rank k_abba_partial pass over the products (HBM) and the float64 "
               "reduce -- with the exchange fused into the last pass over peer memory only the "
               "HBM pass is left; median: the exchange moves (world-1)/world of every product once; "
               "gather: the root's NVLink ingest (900 GB/s per direction)")
    out["limiter"] = limiter
    return out


# --------------------------------------------------------------------- GPU arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from pyemir_b200 import abba, synthetic
    from pyemir_b200.distributed import (bind_to_gpu_numa, gather_frames, init_from_env,
                                         reduce_mean)
    from pyemir_b200.plan import RectWavePlan

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pyemir_b200 has no CPU fallback")
    # before anything is allocated: this rank's CPUs and page-locked buffers next to its GPU
    numa = bind_to_gpu_numa(int(os.environ.get("LOCAL_RANK", "0")))
    rank, world, local_rank = init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    config = config_key(args.config)
    cfg = synthetic.BASELINE_CONFIGS[config]

    coeff = synthetic.make_config(config)
    torch.cuda.synchronize()
    t_plan = time.perf_counter()
    plan = RectWavePlan(coeff, args.resampling, args.dtype, max_order=5,
                        tile_k=args.tile_k, stages=args.stages, device=local_rank)
    plan_first_ms = 1e3 * (time.perf_counter() - t_plan)      # includes CUDA module loading
    plan.set_tuning(args.g)
    info = plan.info()
    F = args.frames
    npdt = np.float32 if args.dtype == "float32" else np.float64
    frames_np = synthetic.make_frames(coeff, F, seed=cfg["seed"] + 31 * rank, dtype=npdt)

    h_in = torch.empty(frames_np.shape, dtype=torch.from_numpy(frames_np[:1]).dtype,
                       pin_memory=True)
    h_in.copy_(torch.from_numpy(frames_np))
    del frames_np
    d_in = h_in.to(dev, non_blocking=True)
    d_out = torch.empty((F,) + plan.out_shape, dtype=d_in.dtype, device=dev)
    d_jmm = torch.empty((F, 55, 2), dtype=torch.int32, device=dev)
    counts = [F] * world

    abba_labels = abba.labels_of(abba.build_full_set("ABBA", 1, (F + 3) // 4)[:F])

    def step():
        plan.run_torch(d_in, d_out, d_jmm)
        if args.combine == "mean":
            reduce_mean(d_out, F * world, 0)
        elif args.combine == "abba":
            abba.combine_abba_mean(d_out, abba_labels, world * abba_labels.count(1),
                                   world * abba_labels.count(-1), 0)
        elif args.combine == "abba_median":
            abba.combine_abba_median(d_out, abba_labels * world, counts, 0)
        elif args.combine == "gather":
            gather_frames(d_out, counts, 0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    events = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    t0 = time.perf_counter()
    events[0].record()
    for i in range(args.steps):
        step()
        events[i + 1].record()
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1)
    elapsed_ms = events[0].elapsed_time(events[-1])
    launch_ms = [events[i].elapsed_time(events[i + 1]) for i in range(args.steps)]
    el = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    elapsed_ms = float(el.item())
    value = world * F * args.steps / (elapsed_ms / 1e3)

    # ---- k_rectwv alone: the same launches told that d_out already holds the zeros outside
    # the wavelength coverage (EMIRWV_RUN_OUT_ZEROED: no k_zero_fill beside it)
    own_events = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    for _ in range(3):
        plan.run_torch(d_in, d_out, d_jmm, out_zeroed=True)
    barrier()
    own_events[0].record()
    for i in range(args.steps):
        plan.run_torch(d_in, d_out, d_jmm, out_zeroed=True)
        own_events[i + 1].record()
    barrier()
    own_ms = own_events[0].elapsed_time(own_events[-1]) / args.steps

    # ---- cost of a plan: first build of the process (above, with CUDA module loading), a
    # second build of the same product (what a new product costs), a cache hit
    from pyemir_b200.plan import clear_plan_cache, get_plan
    t0p = time.perf_counter()
    plan2 = RectWavePlan(coeff, args.resampling, args.dtype, max_order=5, device=local_rank)
    torch.cuda.synchronize()
    plan_cold_ms = 1e3 * (time.perf_counter() - t0p)
    plan2.close()
    clear_plan_cache()
    get_plan(coeff, args.resampling, args.dtype, max_order=5, device=local_rank)
    t0p = time.perf_counter()
    get_plan(coeff, args.resampling, args.dtype, max_order=5, device=local_rank)
    plan_cached_ms = 1e3 * (time.perf_counter() - t0p)
    clear_plan_cache()
    plan_cost = {"first_in_process_ms": plan_first_ms, "new_product_ms": plan_cold_ms,
                 "cached_ms": plan_cached_ms, "table_MB": info["table_bytes"] / 1e6}

    # ---- end to end through the host-buffer C-ABI entry (pinned host in/out)
    e2e = None
    if args.e2e_steps > 0:
        h_out = torch.empty((F,) + plan.out_shape, dtype=h_in.dtype, pin_memory=True)
        h_jmm = torch.empty((F, 55, 2), dtype=torch.int32, pin_memory=True)
        a_in, a_out, a_jmm = h_in.numpy(), h_out.numpy(), h_jmm.numpy()
        plan.run_host(a_in, a_out, a_jmm)                    # warm-up (allocates staging)

        def timed_host(out_zeroed):
            plan.run_host(a_in, a_out, a_jmm, out_zeroed=out_zeroed)      # untimed warm-up
            barrier()
            t0e = time.perf_counter()
            for _ in range(args.e2e_steps):
                plan.run_host(a_in, a_out, a_jmm, out_zeroed=out_zeroed)
            barrier()
            te = torch.tensor([time.perf_counter() - t0e], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return world * F * args.e2e_steps / float(te.item())

        # (1) the whole product delivered into the caller's buffer, no promise asked: the
        # covered columns travel over PCIe, the zeros outside the wavelength coverage are
        # written by host threads of the library while the copies run; before it, the same
        # call with every byte shipped over PCIe (EMIRWV_HOST_ZERO_THREADS=0)
        keep_env = os.environ.get("EMIRWV_HOST_ZERO_THREADS")
        os.environ["EMIRWV_HOST_ZERO_THREADS"] = "0"
        all_dma = timed_host(False)
        if keep_env is None:
            del os.environ["EMIRWV_HOST_ZERO_THREADS"]
        else:
            os.environ["EMIRWV_HOST_ZERO_THREADS"] = keep_env
        a_out[...] = np.float32(7.0) if args.dtype == "float32" else 7.0     # not zeros
        dense = timed_host(False)
        last_dense = a_out[F - 1].copy()
        jmm_dense = a_jmm.copy()
        d_chk, _ = plan.run_torch(d_in[F - 1:F])
        if not np.array_equal(d_chk[0].cpu().numpy(), last_dense):
            raise SystemExit("bench.py: host entry differs from the device-resident kernel")
        # (2) the same call told that the buffer already holds the zeros outside the
        # wavelength coverage (EMIRWV_HOST_OUT_ZEROED: the loop over exposures reuses its
        # pinned buffer); the buffer is cleared first so that the check below sees only
        # what this variant copied
        a_out[...] = 0
        a_jmm[...] = 0
        sparse = timed_host(True)
        same = bool(np.array_equal(a_out[F - 1], last_dense) and np.array_equal(a_jmm, jmm_dense))
        cover = plan.coverage().astype(np.int64)
        d2h_sparse = int((cover[:, 1] - cover[:, 0]).sum() * 38 * a_out.itemsize * F + a_jmm.nbytes)
        e2e = {"value": dense, "unit": UNIT,
               "h2d_bytes_per_step": int(a_in.nbytes),
               "d2h_bytes_per_step": d2h_sparse,
               "call": "emirwv_run_host: pinned host frames in, the whole product delivered "
                       "into the caller's pinned buffer (no promise about its contents): "
                       "covered columns and the JMN/JMX scan over PCIe, the zeros outside the "
                       "wavelength coverage written by 3 host threads of the library",
               "all_bytes_over_pcie": {"value": all_dma,
                                       "d2h_bytes_per_step": int(a_out.nbytes + a_jmm.nbytes),
                                       "call": "the same with EMIRWV_HOST_ZERO_THREADS=0: every "
                                               "byte of the products copied back"},
               "zeroed_buffer": {"value": sparse, "d2h_bytes_per_step": d2h_sparse,
                                 "call": "emirwv_run_host_ex(EMIRWV_HOST_OUT_ZEROED): the caller's "
                                         "buffer already holds the zeros outside the wavelength "
                                         "coverage, only covered columns travel; result identical "
                                         "to the full copy: %s" % same},
               "steps": args.e2e_steps,
               "checksum": float(a_out[0, 1000:1038].sum())}
        if not same:
            raise SystemExit("bench.py: sparse device->host copy differs from the full copy")
        if args.dtype == "float32":
            e2e.update(e2e_recipe_entries(args, plan, h_in, a_out, a_jmm, F, world, dev, barrier,
                                          d2h_sparse))

    collective = None
    if world > 1 and args.dtype == "float32" and not args.no_collective:
        collective = collective_block(args, plan, d_in, d_out, d_jmm, F, world, rank, dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the reference's own use: one frame per call through the drop-in (HDUList in, HDUList
    # out, pageable numpy buffers, header rewritten), plan cached after the first call.
    # Reported beside the batch numbers; a failure here never loses the bench line.
    dropin = None
    if args.e2e_steps > 0 and args.dtype == "float32":
        try:
            from pyemir_b200.apply_rectwv_coeff import apply_rectwv_coeff
            hdul = synthetic.make_hdulist(h_in[0].numpy(), coeff)
            kw = dict(args_resampling=args.resampling, max_order=5, device=local_rank)
            apply_rectwv_coeff(hdul, coeff, **kw)
            times = []
            for _ in range(7):
                t0d = time.perf_counter()
                res = apply_rectwv_coeff(hdul, coeff, **kw)
                times.append(time.perf_counter() - t0d)
            med = statistics.median(times)
            dropin = {"call": "apply_rectwv_coeff(HDUList, RectWaveCoeff) -> HDUList, one frame per "
                              "call, pageable numpy buffers, header rewritten, plan cached",
                      "ms_per_call": 1e3 * med, "frames_per_s": 1.0 / med,
                      "out_shape": list(res[0].data.shape)}
        except Exception as exc:          # noqa: BLE001
            dropin = {"error": repr(exc)}

    # ---- roofline of the dominant kernel (k_rectwv): algorithmic bytes / launch time
    bytes_per_launch = F * (info["in_frame_bytes"] + info["out_frame_bytes"])
    avg_launch_s = statistics.mean(launch_ms) / 1e3
    achieved = bytes_per_launch / avg_launch_s / 1e9
    peak, peak_src = FALLBACK_HBM_GBS, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fd:
            peak = float(json.load(fd)["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        pass
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fd:
            rec = json.load(fd).get(f"config{config}_{args.dtype}_r{args.resampling}")
            if rec and rec.get("frames"):
                traffic = rec["dram_bytes_per_launch"] * F / rec["frames"]
    except (OSError, ValueError):
        pass
    cover = plan.coverage().astype(np.int64)
    esize = 4 if args.dtype == "float32" else 8
    own_bytes = F * (info["in_frame_bytes"] + int((cover[:, 1] - cover[:, 0]).sum()) * 38 * esize)
    own_gbs = own_bytes / (own_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel": "one emirwv_run = k_rectwv (its copy warps also write the zero runs "
                          "outside the wavelength coverage) + k_init_jmm: whole frames in, whole "
                          "products out; the STEP, identical to k_rectwv's own launch",
                "bytes_per_launch": bytes_per_launch,
                "launch_ms": statistics.mean(launch_ms),
                "k_rectwv_alone": {"launch_ms": own_ms, "bytes_per_launch": own_bytes,
                                   "achieved": own_gbs, "frac": own_gbs / peak,
                                   "what": "k_rectwv told that the output already holds the zeros "
                                           "(EMIRWV_RUN_OUT_ZEROED): frames in + only the covered "
                                           "columns of the products out, fewer bytes in less time"}}

    cpu = None
    if world == 1 and args.cpu_frames > 0:
        cpu = cpu_baseline_single(config, args.resampling, args.cpu_frames)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.dtype == "float32" else "f64", "data": "synthetic",
        "config": workload_config(args, config), "tuning": tuning_of(info), "plan": plan_cost,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "dropin_single_frame": dropin,
        # per step: k_init_jmm + k_rectwv (persistent; + k_zero_fill when the output rows are
        # not 16-byte aligned) (+ the combination kernels: abba 2, abba_median 3 per step)
        "gpu_launches": (2 + (0 if (info["out_frame_bytes"] // (55 * 38)) % 16 == 0 else 1)
                         + {"abba": 2, "abba_median": 3}.get(args.combine, 0)) * args.steps,
        "clocks": clocks, "host_placement": numa, "collective": collective,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
