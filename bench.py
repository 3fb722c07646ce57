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
CATALOGUE = 2_000_000  # item table 496 MB, per-model code rows 168 MB: both larger than the 126 MB L2
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


def _splitmix(x):
    """item-id hashes for the synthetic catalogue (any non-zero u64 is a valid mr_hash64 value)"""
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    x = x ^ (x >> np.uint64(31))
    return np.where(x == 0, np.uint64(1), x).astype(np.uint64)


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
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
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


def _usable_cores():
    """Threads the CPU arm can really use: os.cpu_count() capped by the affinity mask and the cgroup CPU quota
    (a 128-CPU host behind a 16-CPU quota runs 128 OpenMP threads slower than 16)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]  # cgroup v2
        if quota != "max":
            n = min(n, max(1, -(-int(quota) // int(period))))
    except (OSError, ValueError):
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())  # cgroup v1
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0 and per > 0:
                n = min(n, max(1, -(-q // per)))
        except (OSError, ValueError):
            pass
    return max(1, n)


def _best_threads(ob, X, rows):
    """The thread count (usable cores vs. every logical CPU) that scores a sample fastest: the CPU arm gets the
    better of the two."""
    cands = sorted({_usable_cores(), os.cpu_count() or 1})
    best, best_dt = cands[0], None
    n = min(rows, 20000)
    for c in cands:
        ob.predictMat(X[:n], n, FEATURES, threads=c)
        t0 = time.perf_counter()
        ob.predictMat(X[:n], n, FEATURES, threads=c)
        dt = time.perf_counter() - t0
        if best_dt is None or dt < best_dt:
            best, best_dt = c, dt
    return best


def run_reference(args, rank, world):
    """CPU arm: the oracle port on all host cores, bounded sample per step."""
    if rank != 0:
        return
    from oracle import oracle
    blob = _model_blob()
    ob = oracle.OracleBooster(0, blob)
    sample_requests = max(1000, (os.cpu_count() or 1) * 40)
    rows = sample_requests * ITEMS
    cat = _matrix(CATALOGUE, DATA_SEED)
    pick = np.random.Generator(np.random.PCG64(DATA_SEED + 1000)).integers(0, CATALOGUE, rows)
    cores = _best_threads(ob, cat[pick[:20000]], 20000)
    for _ in range(args.warmup):
        ob.predictMat(cat[pick], rows, FEATURES, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        X = cat[pick]  # feature assembly on the CPU = row gather of the stored scalars
        ob.predictMat(X, rows, FEATURES, threads=cores)
    dt = time.perf_counter() - t0
    v = rows * args.steps / dt
    sample = f"{sample_requests} requests x {ITEMS} items per step (same generator/seed as the GPU arm)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: 100-item /rank requests, 30 scalar (number) features gathered from a "
                               "2M-item state table, 500-tree LightGBM LambdaMART", "items_per_request": ITEMS,
                   "features": FEATURES, "trees": TREES, "catalogue_items": CATALOGUE,
                   "requests_per_step": sample_requests},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference scorer (ltrlib -> LightGBM JNI) is not installable here (no JVM, no jars); "
                "this is the C oracle port of its arithmetic, OpenMP over rows like LightGBM's predictor",
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
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
        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from metarank_b200 import features as F

    ctx = mb.Context(local)
    blob = _model_blob()
    booster = mb.LightGBMBooster(ctx, blob, n_features=FEATURES)
    for kv in filter(None, os.environ.get("MR_BENCH_OPTS", "").split(",")):  # tuning aid: "chunk_kb=24,threads=448"
        key, val = kv.split("=")
        booster.set_option(key, int(val))

    # ---- device-resident state: CATALOGUE items x 30 `number` features (Persistence.values on HBM)
    names = [f"f{j}" for j in range(FEATURES)]
    feats = [dict(name=n, type="number", scope="item", source=f"metadata.{n}") for n in names]
    mapping = F.FeatureMapping(ctx, feats, names)
    state = F.DeviceState(ctx, mapping)
    cat = _matrix(CATALOGUE, DATA_SEED)  # same column recipe as SURVEY.md 8d (5 % NaN = missing state)
    item_ids = _splitmix(np.arange(1, CATALOGUE + 1, dtype=np.uint64))
    t_up = time.perf_counter()
    for j0 in range(0, FEATURES, 5):
        state.put_packed(F.pack_number_columns(names[j0:j0 + 5], item_ids, cat[:, j0:j0 + 5]))
    state.flush()
    t_up = time.perf_counter() - t_up

    R = args.requests_per_step
    rows = R * ITEMS
    rng = np.random.Generator(np.random.PCG64(DATA_SEED + 1000 + rank))  # every rank ranks its own requests
    pick = rng.integers(0, CATALOGUE, rows)
    ids_host = item_ids[pick]
    offs_host = (np.arange(R + 1, dtype=np.int32) * ITEMS).astype(np.int32)
    d_ids = torch.from_numpy(ids_host.view(np.int64)).cuda()
    d_offs = torch.from_numpy(offs_host).cuda()
    d_out = torch.empty(rows, dtype=torch.float64, device="cuda")
    d_order = torch.empty(rows, dtype=torch.int32, device="cuda")
    d_feat = torch.empty(rows * FEATURES, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    def step(explain=False):
        # explain=false is the /rank default: the f64 matrix is not materialised, the assemble kernel
        # emits the binned scorer's u16 codes directly
        F.rank_device(state, booster, R, rows, d_offs.data_ptr(), d_ids.data_ptr(), d_out.data_ptr(),
                      d_order.data_ptr(), d_feat.data_ptr() if explain else 0, sptr)

    for _ in range(args.warmup):
        step()
    F.rank_device_status(state, sptr)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches1 = mb._capi.lib().mr_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    gpu_launches = mb._capi.lib().mr_kernel_launches() - launches1
    total_ms = e0.elapsed_time(e1)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    rank_ms = [total_ms / args.steps]
    if world > 1:
        allt = torch.zeros(world, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allt, t)
        rank_ms = [float(x) / args.steps for x in allt.cpu()]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = rows * world * args.steps / (total_ms / 1e3)

    # ---- dominant kernel alone (gbdt_score on the assembled matrix), CUDA events on the launching stream
    step(explain=True)  # materialise the f64 matrix once for the stand-alone kernel timings
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    n_codes = booster.codes_bytes(rows)
    if n_codes:  # binned scorer: time the traversal kernel alone on precomputed codes
        d_codes = torch.empty(n_codes, dtype=torch.uint8, device="cuda")
        booster.bin_device(d_feat.data_ptr(), rows, FEATURES, d_codes.data_ptr(), sptr)
        kernel_name = "gbdt_score_compact_kernel (u16 rank codes, 8-byte nodes)"
    else:
        kernel_name = "gbdt_score_kernel"
    for a, b in evs:
        a.record(stream)
        if n_codes:
            booster.score_codes_device(d_codes.data_ptr(), rows, d_out.data_ptr(), sptr)
        else:
            booster.predict_device(d_feat.data_ptr(), rows, FEATURES, d_out.data_ptr(), sptr)
        b.record(stream)
    # ... and assembly alone (model = None)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(stream)
    for _ in range(args.steps):
        F.rank_device(state, None, R, rows, d_offs.data_ptr(), d_ids.data_ptr(), 0, 0, d_feat.data_ptr(), sptr)
    a1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    assemble_ms = a0.elapsed_time(a1) / args.steps
    step(explain=True)
    F.rank_device_status(state, sptr)
    n_chk = 20
    got = d_out[: n_chk * ITEMS].cpu().numpy()
    got_order = d_order[: n_chk * ITEMS].cpu().numpy()
    got_feat = d_feat[: n_chk * ITEMS * FEATURES].cpu().numpy().reshape(-1, FEATURES)

    # ---- e2e: the same metric through mr_rank with HOST buffers (item-id hashes in, scores + order out)
    rk = F.Ranker(mapping, state)
    # page-locked request/response buffers (what a JVM gets from a registered direct ByteBuffer):
    # the library DMAs them in place instead of staging
    ids_pin = torch.from_numpy(ids_host.view(np.int64)).pin_memory()
    sc_pin = torch.empty(rows, dtype=torch.float64).pin_memory()
    ord_pin = torch.empty(rows, dtype=torch.int32).pin_memory()
    ids_pinned_np = ids_pin.numpy().view(np.uint64)
    arrays = dict(offsets=offs_host, ids=ids_pinned_np, users=np.zeros(R, dtype=np.uint64),
                  sessions=np.zeros(R, dtype=np.uint64), req_f64=np.zeros((R, 1)), req_u64=np.zeros((R, 1), dtype=np.uint64),
                  req_vec=np.zeros((R, 1), dtype=np.float32), req_vp=np.zeros((R, 1), dtype=np.uint8), item_f64=None,
                  n_requests=R, total_items=rows)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        sc_h, ord_h, _ = rk.rank_arrays(arrays, booster, want_order=True, out_scores=sc_pin.numpy(), out_order=ord_pin.numpy())
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sc_h, ord_h, _ = rk.rank_arrays(arrays, booster, want_order=True, out_scores=sc_pin.numpy(), out_order=ord_pin.numpy())
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = rows * world * e2e_steps / float(t.item())
    e2e_ok = bool(np.array_equal(sc_h[: n_chk * ITEMS], got) and np.array_equal(ord_h[: n_chk * ITEMS], got_order))

    # ---- single-request latency through the C ABI (p50 of 300 calls, 100 items, host buffers)
    lat = None
    if rank == 0:
        one = dict(arrays, offsets=offs_host[:2].copy(), ids=ids_host[:ITEMS].copy(), n_requests=1, total_items=ITEMS)
        for _ in range(20):
            rk.rank_arrays(one, booster, want_order=True)
        ts = []
        for _ in range(300):
            a = time.perf_counter()
            rk.rank_arrays(one, booster, want_order=True)
            ts.append(time.perf_counter() - a)
        lat = {"p50_ms": float(np.percentile(ts, 50) * 1e3), "p99_ms": float(np.percentile(ts, 99) * 1e3),
               "what": "mr_rank, 1 request x 100 items: id hashes in -> scores + order out, host buffers, via ctypes"}

    if rank == 0:
        from oracle import oracle
        ob = oracle.OracleBooster(0, blob)
        want_feat = cat[pick[: n_chk * ITEMS]]  # the oracle's assembly of `number` features is the stored scalar
        want = ob.predictMat(want_feat, n_chk * ITEMS, FEATURES)
        feat_ok = bool(np.array_equal(got_feat, want_feat, equal_nan=True))
        parity = bool(np.array_equal(got, want))
        order_ok = all(np.array_equal(got_order[i * ITEMS:(i + 1) * ITEMS],
                                      oracle.rank_order(want[i * ITEMS:(i + 1) * ITEMS])) for i in range(n_chk))
        # roofline of the dominant kernel: algorithmic bytes per item, SURVEY.md 8(d): 8F + T*(dbar*16 + 8) + 8
        dbar = booster.mean_path(want_feat[:2000], 2000, FEATURES)
        b_item = 8 * FEATURES + TREES * (dbar * 16 + 8) + 8
        peak, peak_src = _peaks()
        achieved = b_item * rows / (kernel_ms / 1e3) / 1e9
        # CPU baseline on a bounded sample, all host cores: hash lookup + row gather + tree walk
        cpu_rows = min(rows, max(2000 * ITEMS, (os.cpu_count() or 1) * 40 * ITEMS))
        Xc = cat[pick[:cpu_rows]]
        cores = _best_threads(ob, Xc, cpu_rows)
        c0 = time.perf_counter()
        Xc = cat[pick[:cpu_rows]]  # the gather is part of the CPU path too
        ob.predictMat(Xc, cpu_rows, FEATURES, threads=cores)
        cpu_dt = time.perf_counter() - c0
        info = state.info()
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: 100-item /rank requests, 30 scalar (number) features gathered from a "
                                   "2M-item device-resident state table, 500-tree LightGBM LambdaMART, "
                                   "scores + per-request ordering",
                       "items_per_request": ITEMS, "features": FEATURES, "trees": TREES, "catalogue_items": CATALOGUE,
                       "requests_per_step": R, "parallelism": f"requests sharded over {world} GPU(s), state replicated, no collective",
                       "l2": "every table a step gathers from at random is larger than the 126 MB L2 (2 M items: item rows 496 MB, per-model code rows 168 MB), and a step writes and re-reads 138 MB of codes plus 33 MB of ids / scores / order: no flush needed"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(ids_host.nbytes + offs_host.nbytes + 2 * R * 8),
                    "d2h_bytes_per_step": rows * 12, "steps": e2e_steps, "parity_ok": e2e_ok,
                    "what": "mr_rank: item-id hashes in page-locked host memory -> scores + order in page-locked host memory, "
                            "H2D and D2H copies inside the timer"},
            "gpu_launches": int(gpu_launches), "rank_ms_per_step": rank_ms,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "bytes_per_item": b_item, "mean_path": dbar,
                         "kernel_ms": kernel_ms, "kernel": kernel_name,
                         "step_share": kernel_ms / (total_ms / args.steps), "assemble_ms": assemble_ms},
            "cpu_baseline": {"value": cpu_rows / cpu_dt, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{cpu_rows // ITEMS} requests x {ITEMS} items: numpy row gather + C oracle, OpenMP over rows"},
            "clocks": clocks, "latency": lat,
            "parity": {"features_bit_identical": feat_ok, "scores_bit_identical": parity,
                       "ordering_identical": bool(order_ok), "checked_items": n_chk * ITEMS},
            "state": {"items": int(info.rows[1]), "device_bytes": int(info.device_bytes),
                      "item_row_bytes": int(info.item_row_bytes), "upload_s": t_up},
        }
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file):
            try:
                out["roofline"]["traffic"] = json.load(open(traffic_file)).get("dominant_kernel_c2_bytes_per_launch")
            except Exception:
                pass
        print(json.dumps(out), flush=True)

    state.free()
    mapping.free()
    booster.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
