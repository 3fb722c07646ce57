#!/usr/bin/env python
"""bench.py — items/sec reranked on B200 (BASELINE.json metric), one JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE configs[1] — synthetic 100-item /rank requests,
30 scalar features, 500-tree LightGBM-shaped LambdaMART.  One *step* = one pass of the
hot path over a batch of REQUESTS_PER_STEP such requests (the batch's f64 feature
matrix, 393 MB, is larger than the 126 MB L2, so no L2 flush is needed between steps).

  value      whole-job items/s with inputs resident in HBM (device-timed, max over ranks)
  e2e        the same metric through the C ABI with HOST buffers (H2D/D2H inside the timer)
  roofline   algorithmic bytes (SURVEY.md §8d B_item with the measured mean path) per launch
             / kernel duration, against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline / --impl reference: the CPU oracle port of the booster arithmetic (the
             reference's own scorer is a JNI jar that is not in this image) on host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ITEMS = 100
FEATURES = 30
TREES = 500
REQUESTS_PER_STEP = 16384
MODEL_SEED = 1234 + 2
DATA_SEED = 42 + 2
METRIC = "items/sec reranked (100-item req, 500-tree LambdaMART)"
UNIT = "items/s"


def _model_blob():
    from metarank_b200 import synth
    return synth.lightgbm_model_text(TREES, FEATURES, 16, 8, seed=MODEL_SEED)


def _matrix(rows, seed):
    from metarank_b200 import synth
    return synth.feature_matrix(rows, FEATURES, seed=seed)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args, rank, world):
    """CPU arm: the oracle port on all host cores, bounded sample per step."""
    if rank != 0:
        return
    from oracle import oracle
    blob = _model_blob()
    ob = oracle.OracleBooster(0, blob)
    cores = os.cpu_count() or 1
    sample_requests = 1000
    rows = sample_requests * ITEMS
    X = _matrix(rows, DATA_SEED)
    for _ in range(args.warmup):
        ob.predictMat(X, rows, FEATURES, threads=0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ob.predictMat(X, rows, FEATURES, threads=0)
    dt = time.perf_counter() - t0
    v = rows * args.steps / dt
    sample = f"{sample_requests} requests x {ITEMS} items per step (same generator/seed as the GPU arm)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: 100-item requests x 30 scalar features x 500-tree LightGBM LambdaMART "
                               "(predictMat on the assembled matrix)", "items_per_request": ITEMS,
                   "features": FEATURES, "trees": TREES, "requests_per_step": sample_requests},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference scorer (ltrlib -> LightGBM JNI) is not installable here (no JVM, no jars); "
                "this is the C oracle port of its arithmetic, OpenMP over rows like LightGBM's predictor",
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--requests-per-step", type=int, default=REQUESTS_PER_STEP)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import metarank_b200 as mb

    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = mb.Context(local)
    blob = _model_blob()
    booster = mb.LightGBMBooster(ctx, blob, n_features=FEATURES)

    R = args.requests_per_step
    rows = R * ITEMS
    X_host = _matrix(rows, DATA_SEED + rank)  # every rank scores its own requests (weak scaling)
    d_X = torch.from_numpy(X_host).cuda()
    d_out = torch.empty(rows, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    def step():
        booster.predict_device(d_X.data_ptr(), rows, FEATURES, d_out.data_ptr(), sptr)

    launches0 = mb._capi.lib().mr_kernel_launches()
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches1 = mb._capi.lib().mr_kernel_launches()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for a, b in evs:
        a.record(stream)
        step()
        b.record(stream)
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    gpu_launches = mb._capi.lib().mr_kernel_launches() - launches1
    total_ms = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = rows * world * args.steps / (total_ms / 1e3)

    # parity spot check against the oracle on the first requests of this rank (outside the timer)
    got = d_out[: 20 * ITEMS].cpu().numpy()

    # ---- e2e: same metric through the C ABI with host buffers (pinned), H2D + D2H inside the timer
    Xp = torch.from_numpy(X_host).pin_memory()
    out_p = torch.empty(rows, dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    lib = mb._capi.lib()
    import ctypes as C

    def e2e_step():
        mb._capi.check(lib.mr_model_predict_mat(booster._h, C.c_void_p(Xp.data_ptr()), C.c_int32(rows),
                                                C.c_int32(FEATURES), C.c_void_p(out_p.data_ptr())))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = rows * world * e2e_steps / float(t.item())
    e2e_ok = bool(np.array_equal(out_p[: 20 * ITEMS].numpy(), got))

    # ---- single-request latency through the C ABI (p50 of 300 calls, 100 items, host buffers)
    lat = None
    if rank == 0:
        x1 = np.ascontiguousarray(X_host[:ITEMS])
        for _ in range(20):
            booster.predictMat(x1, ITEMS, FEATURES)
        ts = []
        for _ in range(300):
            a = time.perf_counter()
            booster.predictMat(x1, ITEMS, FEATURES)
            ts.append(time.perf_counter() - a)
        lat = {"p50_ms": float(np.percentile(ts, 50) * 1e3), "p99_ms": float(np.percentile(ts, 99) * 1e3),
               "what": "mr_model_predict_mat, 1 request x 100 items, host buffers, incl. H2D/D2H + ctypes"}

    if rank == 0:
        from oracle import oracle
        ob = oracle.OracleBooster(0, blob)
        want = ob.predictMat(X_host[: 20 * ITEMS], 20 * ITEMS, FEATURES)
        parity = bool(np.array_equal(got, want))
        order_ok = all(np.array_equal(ctx.rank_order(got[i * ITEMS:(i + 1) * ITEMS]),
                                      oracle.rank_order(want[i * ITEMS:(i + 1) * ITEMS])) for i in range(20))
        # roofline: algorithmic bytes per item, SURVEY.md §8(d): 8F + T*(dbar*16 + 8) + 8
        dbar = booster.mean_path(X_host[:8192], 8192, FEATURES)
        b_item = 8 * FEATURES + TREES * (dbar * 16 + 8) + 8
        peak, peak_src = _peaks()
        achieved = b_item * rows / (kernel_ms / 1e3) / 1e9
        # CPU baseline on a bounded sample, all host cores (~10-30 s of CPU work)
        cores = os.cpu_count() or 1
        cpu_rows = 2000 * ITEMS
        Xc = X_host[:cpu_rows]
        ob.predictMat(Xc[:10000], 10000, FEATURES, threads=0)
        c0 = time.perf_counter()
        ob.predictMat(Xc, cpu_rows, FEATURES, threads=0)
        cpu_dt = time.perf_counter() - c0
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: 100-item requests x 30 scalar features x 500-tree LightGBM LambdaMART "
                                   "(predictMat on the assembled matrix)",
                       "items_per_request": ITEMS, "features": FEATURES, "trees": TREES,
                       "requests_per_step": R, "parallelism": f"requests sharded over {world} GPU(s), no collective",
                       "l2": "inputs (393 MB/step/GPU) larger than the 126 MB L2; no flush needed"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": rows * FEATURES * 8,
                    "d2h_bytes_per_step": rows * 8, "steps": e2e_steps, "parity_ok": e2e_ok,
                    "what": "mr_model_predict_mat on pinned host buffers, copies inside the timed region"},
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "bytes_per_item": b_item, "mean_path": dbar,
                         "kernel_ms": kernel_ms, "kernel": "gbdt_score_kernel"},
            "cpu_baseline": {"value": cpu_rows / cpu_dt, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{cpu_rows // ITEMS} requests x {ITEMS} items, OpenMP over rows"},
            "clocks": clocks, "latency": lat,
            "parity": {"scores_bit_identical": parity, "ordering_identical": bool(order_ok), "checked_items": 20 * ITEMS},
        }
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file):
            try:
                out["roofline"]["traffic"] = json.load(open(traffic_file)).get("gbdt_score_c2_bytes_per_launch")
            except Exception:
                pass
        print(json.dumps(out), flush=True)

    booster.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
