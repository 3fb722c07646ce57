#!/usr/bin/env python
"""bench.py — items/sec reranked on B200 (BASELINE.json metric), one JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C2|C3|C4|C5] [--no-extras]

Configs are BASELINE.json's (SURVEY.md §8 shorthand); the default and headline is C2:
  C2  100-item /rank requests, 30 scalar features, 500-tree LightGBM LambdaMART (16 384 requests per step)
  C3  ranklens feature set (rate / window_count / interacted_with / diversity ...), 1000-item requests
  C4  bi-encoder cosine (384-d) + 15 numbers, 256-item requests, 200-tree XGBoost
  C5  ONE 10 000-item request, 64 features, 2000 trees, item-sharded over the N GPUs (mr_group_rank: the
      scorer stores into every GPU's exchange buffer over NVLink, no collective call on the data path)
One *step* = one pass of the hot path (lookup -> assemble -> score -> order) over one batch.

  value        whole-job items/s with inputs resident in HBM (device-timed with CUDA events, max over ranks)
  e2e          the same metric through the C ABI with HOST buffers (H2D/D2H inside the timer), calls issued from two
               worker threads the way a server issues them (`one_call_at_a_time`: a single thread)
  roofline     dominant kernel against the pipe that bounds it.  The GBDT scorer serves the ensemble from
               shared memory (one TMA stage per chunk), so its bound is not HBM but the busier of two on-chip pipes:
               the SHARED-MEMORY crossbar — wavefronts the walk needs (counted in-run by mr_model_walk_stats) x 128 B /
               kernel time against SMs x 128 B/clk x SM clock — or the ISSUE slots — warp instructions the walk needs
               (SASS counts x the same statistics) against SMs x 4 schedulers x SM clock.  `frac` is the larger one, both
               are in the block, next to the ncu values of one capture.  SURVEY.md 8(d)'s algorithmic-bytes figure is
               kept under `algorithmic_gbs` (it exceeds the HBM peak by construction and is not a fraction of anything).
  kernels      every kernel of the step, device time from the library's own per-launch events
               (mr_profile_begin/end), with algorithmic HBM bytes and the HBM fraction for the HBM-bound ones
  cpu_baseline / --impl reference: the CPU oracle port (the reference's own scorer is a JNI jar that is not in
               this image) on the host cores, on a bounded sample of the same workload.
  other_configs  (default C2 run only) short measurements of C3 / C4 / C5 with their own parity checks.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "items/sec reranked (100-item req, 500-tree LambdaMART)"
UNIT = "items/s"
SM_COUNT = 148

CONFIG = {
    "C2": dict(items=100, features=30, trees=500, requests_per_step=16384, catalogue=2_000_000, model_seed=1236, data_seed=44,
               workload="C2: 100-item /rank requests, 30 scalar (number) features gathered from a 2M-item device-resident "
                        "state table, 500-tree LightGBM LambdaMART, scores + per-request ordering"),
    "C3": dict(items=1000, features=24, trees=500, requests_per_step=256, catalogue=20_000, model_seed=1237, data_seed=45,
               workload="C3: ranklens feature set (24 columns: numbers, string index, normalized + field-scoped rate, "
                        "interacted_with over 4 fields, position, 5 diversity features, counts/windows), 1000-item requests, "
                        "500-tree LightGBM with a categorical column"),
    "C4": dict(items=256, features=16, trees=200, requests_per_step=512, catalogue=20_000, model_seed=1238, data_seed=77,
               workload="C4: bi-encoder cosine (384-d item embeddings vs the request's query embedding) + 15 numbers, "
                        "256-item requests, 200-tree depth-6 XGBoost (f32)"),
    "C5": dict(items=10_000, features=64, trees=2000, requests_per_step=1, catalogue=50_000, model_seed=1239, data_seed=47,
               workload="C5: ONE 10 000-item mega-request, 64 scalar features, 2000-tree LightGBM, items split over the "
                        "GPUs (mr_group_rank: peer stores from the scoring kernel, device-side wait, full ordering on every GPU)"),
}


def config_block(name, world):
    c = CONFIG[name]
    d = {"workload": c["workload"], "items_per_request": c["items"], "features": c["features"], "trees": c["trees"],
         "catalogue_items": c["catalogue"], "requests_per_step": c["requests_per_step"]}
    if name == "C5":
        d["parallelism"] = "one request item-sharded over the GPUs (strong scaling), state + model replicated"
        d["l2"] = "a single request is latency-bound; the 50 000-row table (26 MB) is L2-resident by nature of the workload"
    else:
        d["parallelism"] = "requests sharded over the GPUs, state + model replicated, no collective"
        d["l2"] = ("every table a step gathers from at random is larger than the 126 MB L2 (2 M items: item rows 496 MB, per-model "
                   "code rows 168 MB), and a step writes and re-reads 138 MB of codes plus 33 MB of ids / scores / order: no flush needed"
                   if name == "C2" else
                   "the step's working set (ids, codes, scores, order, per-request scratch) is rewritten every step; the item "
                   "table is small by the nature of the config and stays L2-resident, as it would in production")
    return d


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
            j = json.load(open(p))
            return float(j["hbm_gbs"]), float(j.get("sm_max_mhz", 1965.0)), "measured"
        except Exception:
            pass
    return 6650.0, 1965.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (and the repeat of it that follows)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            time.sleep(0.3)  # nvidia-smi needs a moment before its first sample
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t_begin, t_end):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, ln in self.lines:
            if t < t_begin or t > t_end + 0.05:
                continue
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
                "samples": len(sm), "window_s": t_end - t_begin, "reasons": sorted(reasons)}


def _usable_cores():
    """Threads the CPU arm can really use: os.cpu_count() capped by the affinity mask and the cgroup CPU quota."""
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


# ======================================================================================= workloads
class Workload:
    """One BASELINE config: synthetic state + model + one batch of requests, on this rank's GPU."""

    name = ""
    kind = 0  # booster kind

    def __init__(self, ctx, rank, world):
        self.ctx, self.rank, self.world = ctx, rank, world
        self.c = CONFIG[self.name]
        self.R = self.c["requests_per_step"]
        self.items = self.c["items"]
        self.rows = self.R * self.items

    # --- to be provided: build(), step(), e2e_step(), parity(), cpu_arm()
    def total_rows(self):
        return self.rows * self.world


def _dev(arrays):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v).view(np.int64) if v.dtype == np.uint64 else np.ascontiguousarray(v)).cuda()
            for k, v in arrays.items() if isinstance(v, np.ndarray)}


class C2(Workload):
    name = "C2"

    def model_blob(self):
        from metarank_b200 import synth
        return synth.lightgbm_model_text(self.c["trees"], self.c["features"], 16, 8, seed=self.c["model_seed"])

    def catalogue(self):
        from metarank_b200 import synth
        return synth.feature_matrix(self.c["catalogue"], self.c["features"], seed=self.c["data_seed"])

    def picks(self, rank):
        return np.random.Generator(np.random.PCG64(self.c["data_seed"] + 1000 + rank)).integers(0, self.c["catalogue"], self.rows)

    def build(self):
        import torch
        import metarank_b200 as mb
        from metarank_b200 import features as F
        c = self.c
        self.blob = self.model_blob()
        self.booster = mb.LightGBMBooster(self.ctx, self.blob, n_features=c["features"])
        for kv in filter(None, os.environ.get("MR_BENCH_OPTS", "").split(",")):  # tuning aid: "chunk_kb=24,threads=448"
            key, val = kv.split("=")
            self.booster.set_option(key, int(val))
        names = [f"f{j}" for j in range(c["features"])]
        feats = [dict(name=n, type="number", scope="item", source=f"metadata.{n}") for n in names]
        self.mapping = F.FeatureMapping(self.ctx, feats, names)
        self.state = F.DeviceState(self.ctx, self.mapping)
        self.cat = self.catalogue()  # same column recipe as SURVEY.md 8d (5 % NaN = missing state)
        self.item_ids = _splitmix(np.arange(1, c["catalogue"] + 1, dtype=np.uint64))
        t0 = time.perf_counter()
        for j0 in range(0, c["features"], 5):
            self.state.put_packed(F.pack_number_columns(names[j0:j0 + 5], self.item_ids, self.cat[:, j0:j0 + 5]))
        self.state.flush()
        self.upload_s = time.perf_counter() - t0
        self.pick = self.picks(self.rank)  # every rank ranks its own requests
        self.ids_host = self.item_ids[self.pick]
        self.offs_host = (np.arange(self.R + 1, dtype=np.int32) * self.items).astype(np.int32)
        self.d_ids = torch.from_numpy(self.ids_host.view(np.int64)).cuda()
        self.d_offs = torch.from_numpy(self.offs_host).cuda()
        self.d_out = torch.empty(self.rows, dtype=torch.float64, device="cuda")
        self.d_order = torch.empty(self.rows, dtype=torch.int32, device="cuda")
        self.d_feat = None
        self.stream = torch.cuda.current_stream().cuda_stream
        self.rk = F.Ranker(self.mapping, self.state)
        self.h2d = int(self.ids_host.nbytes + self.offs_host.nbytes + 2 * self.R * 8)
        self.d2h = self.rows * 12

    def step(self, explain=False):
        import torch
        from metarank_b200 import features as F
        if explain and self.d_feat is None:
            self.d_feat = torch.empty(self.rows * self.c["features"], dtype=torch.float64, device="cuda")
        F.rank_device(self.state, self.booster, self.R, self.rows, self.d_offs.data_ptr(), self.d_ids.data_ptr(),
                      self.d_out.data_ptr(), self.d_order.data_ptr(), self.d_feat.data_ptr() if explain else 0, self.stream,
                      max_items=self.items)

    def status(self):
        from metarank_b200 import features as F
        F.rank_device_status(self.state, self.stream)

    def e2e_prepare(self, n_slots=1):
        import torch
        R, rows = self.R, self.rows
        # page-locked request/response buffers (what a JVM gets from a registered direct ByteBuffer): DMA'd in place;
        # one response buffer pair per in-flight call
        self.ids_pin = torch.from_numpy(self.ids_host.view(np.int64)).pin_memory()
        self.e2e_out = [(torch.empty(rows, dtype=torch.float64).pin_memory(), torch.empty(rows, dtype=torch.int32).pin_memory())
                        for _ in range(n_slots)]
        self.arrays = dict(offsets=self.offs_host, ids=self.ids_pin.numpy().view(np.uint64), users=np.zeros(R, dtype=np.uint64),
                           sessions=np.zeros(R, dtype=np.uint64), req_f64=np.zeros((R, 1)), req_u64=np.zeros((R, 1), dtype=np.uint64),
                           req_vec=np.zeros((R, 1), dtype=np.float32), req_vp=np.zeros((R, 1), dtype=np.uint8), item_f64=None,
                           n_requests=R, total_items=rows)

    def e2e_step(self, slot=0):
        sc, od = self.e2e_out[slot]
        return self.rk.rank_arrays(self.arrays, self.booster, want_order=True, out_scores=sc.numpy(), out_order=od.numpy())[:2]

    def parity(self, n_chk=20):
        """oracle on the first n_chk requests: assembled features, scores, order — all bit for bit."""
        from oracle import oracle
        self.step(explain=True)
        self.status()
        n = n_chk * self.items
        got = self.d_out[:n].cpu().numpy()
        got_order = self.d_order[:n].cpu().numpy()
        got_feat = self.d_feat[: n * self.c["features"]].cpu().numpy().reshape(-1, self.c["features"])
        ob = oracle.OracleBooster(0, self.blob)
        want_feat = self.cat[self.pick[:n]]  # the oracle's assembly of `number` features is the stored scalar
        want = ob.predictMat(want_feat, n, self.c["features"])
        order_ok = all(np.array_equal(got_order[i * self.items:(i + 1) * self.items],
                                      oracle.rank_order(want[i * self.items:(i + 1) * self.items])) for i in range(n_chk))
        self._parity_ref = (got, got_order)
        return {"features_bit_identical": bool(np.array_equal(got_feat, want_feat, equal_nan=True)),
                "scores_bit_identical": bool(np.array_equal(got, want)), "ordering_identical": bool(order_ok),
                "checked_items": n}

    def latency(self):
        one = dict(self.arrays, offsets=self.offs_host[:2].copy(), ids=self.ids_host[:self.items].copy(), n_requests=1,
                   total_items=self.items)
        for _ in range(20):
            self.rk.rank_arrays(one, self.booster, want_order=True)
        ts = []
        for _ in range(300):
            a = time.perf_counter()
            self.rk.rank_arrays(one, self.booster, want_order=True)
            ts.append(time.perf_counter() - a)
        return {"p50_ms": float(np.percentile(ts, 50) * 1e3), "p99_ms": float(np.percentile(ts, 99) * 1e3),
                "what": "mr_rank, 1 request x 100 items: id hashes in -> scores + order out, host buffers, via ctypes"}

    def churn(self, steps=8, frac=0.01):
        """State churn inside the timer: before every step `frac` of the catalogue's rows get new values
        (mr_state_upsert of all 30 columns + mr_state_flush: host table update, scatter upload, and the per-model code rows
        of the touched items re-derived on the next rank call), then the step runs.  Wall clock around upsert + flush + step
        with a device sync per step; parity of the last step against the oracle on the updated values."""
        import torch
        from metarank_b200 import features as F
        from oracle import oracle
        c = self.c
        names = [f"f{j}" for j in range(c["features"])]
        n_upd = max(1, int(c["catalogue"] * frac))
        rng = np.random.Generator(np.random.PCG64(c["data_seed"] + 77))
        packs, rows_upd, vals_upd = [], [], []
        for _ in range(steps + 1):
            rows = rng.choice(c["catalogue"], n_upd, replace=False)
            vals = rng.standard_normal((n_upd, c["features"]))
            packs.append(F.pack_number_columns(names, self.item_ids[rows], vals))
            rows_upd.append(rows); vals_upd.append(vals)
        t_up = t_step = 0.0
        for k in range(steps + 1):  # the first pass warms the path and is not counted
            a = time.perf_counter()
            self.state.put_packed(packs[k]); self.state.flush()
            b = time.perf_counter()
            self.step(); self.status(); torch.cuda.synchronize()
            e = time.perf_counter()
            self.cat[rows_upd[k]] = vals_upd[k]
            if k:
                t_up += b - a; t_step += e - b
        n = min(2000, self.rows)
        got = self.d_out[:n].cpu().numpy()
        want = oracle.OracleBooster(0, self.blob).predictMat(np.ascontiguousarray(self.cat[self.pick[:n]]), n, c["features"])
        return {"what": f"{frac:.0%} of the {c['catalogue']} item rows rewritten (all {c['features']} columns) + mr_state_flush before every step, "
                        "inside the timer; wall clock, device sync per step", "steps": steps, "rows_updated_per_step": int(n_upd),
                "value": self.rows * steps / (t_up + t_step), "unit": UNIT, "upsert_flush_ms_per_step": t_up / steps * 1e3,
                "rank_ms_per_step": t_step / steps * 1e3, "scores_bit_identical_after_updates": bool(np.array_equal(got, want))}

    # --- CPU arm (oracle port): parallel row gather + C oracle scorer + per-request ordering
    def cpu_setup(self):
        from oracle import oracle
        self.ob = oracle.OracleBooster(0, self.model_blob())
        if not hasattr(self, "cat"):
            self.cat = self.catalogue()
        self.cpu_pick = self.picks(0)

    def cpu_run(self, n_requests, threads):
        from oracle import oracle
        rows = n_requests * self.items
        X = oracle.gather_rows(self.cat, self.cpu_pick[:rows], threads)
        sc = self.ob.predictMat(X, rows, self.c["features"], threads=threads)
        oracle.rank_order_batch(sc, (np.arange(n_requests + 1, dtype=np.int32) * self.items).astype(np.int32), threads)
        return rows

    cpu_what = "parallel row gather of the stored scalars + C oracle scorer (OpenMP over rows) + per-request stable ordering"

    def free(self):
        self.state.free(); self.mapping.free(); self.booster.free()


class C5(C2):
    """One mega-request; `value` = items of the request / its latency.  world > 1: mr_group_rank."""
    name = "C5"

    def model_blob(self):
        from metarank_b200 import synth
        return synth.lightgbm_model_text(self.c["trees"], self.c["features"], seed=self.c["model_seed"])

    def picks(self, rank):  # the SAME request on every rank
        return np.random.Generator(np.random.PCG64(self.c["data_seed"] + 1)).choice(self.c["catalogue"], self.rows, replace=False)

    def total_rows(self):
        return self.rows  # strong scaling: the job is one request however many GPUs serve it

    def build(self):
        super().build()
        from metarank_b200 import sharded
        self.group = sharded.Group(self.ctx, self.rank, self.world, self.rows)
        self.group.connect_distributed()

    def step(self, explain=False):
        import torch
        from metarank_b200 import features as F
        if explain:
            if self.d_feat is None:
                self.d_feat = torch.empty(self.rows * self.c["features"], dtype=torch.float64, device="cuda")
            F.rank_device(self.state, None, 1, self.rows, self.d_offs.data_ptr(), self.d_ids.data_ptr(), 0, 0,
                          self.d_feat.data_ptr(), self.stream, max_items=self.rows)
        self.group.rank_device(self.state, self.booster, self.rows, self.d_offs.data_ptr(), self.d_ids.data_ptr(),
                               self.d_out.data_ptr(), self.d_order.data_ptr(), self.stream)

    def e2e_step(self, slot=0):
        return self.group.rank_arrays(self.state, self.booster, self.arrays)

    def latency(self):
        ts = []
        for _ in range(5):
            self.e2e_step()
        for _ in range(50):
            a = time.perf_counter()
            self.e2e_step()
            ts.append(time.perf_counter() - a)
        return {"p50_ms": float(np.percentile(ts, 50) * 1e3), "p99_ms": float(np.percentile(ts, 99) * 1e3),
                "what": f"mr_group_rank on {self.world} GPU(s), one 10 000-item request: id hashes in -> scores + order out, host buffers"}

    def parity(self, n_chk=1):
        from oracle import oracle
        self.step(explain=True)
        self.status()
        got = self.d_out.cpu().numpy()
        got_order = self.d_order.cpu().numpy()
        got_feat = self.d_feat.cpu().numpy().reshape(-1, self.c["features"])
        want_feat = self.cat[self.pick]
        want = oracle.OracleBooster(0, self.blob).predictMat(want_feat, self.rows, self.c["features"], threads=_usable_cores())
        self._parity_ref = (got, got_order)
        return {"features_bit_identical": bool(np.array_equal(got_feat, want_feat, equal_nan=True)),
                "scores_bit_identical": bool(np.array_equal(got, want)),
                "ordering_identical": bool(np.array_equal(got_order, oracle.rank_order(want))), "checked_items": self.rows}

    def free(self):
        self.group.free()
        super().free()


class GenericWorkload(Workload):
    """C3 / C4: requests as RankingEvent dicts -> FeatureMapping.pack_requests -> mr_rank_device."""

    def finish_build(self, feats, model, state, reqs, blob, kind):
        import torch
        import metarank_b200 as mb
        from metarank_b200 import features as F
        self.feats, self.model_names, self.state_dict, self.reqs, self.blob, self.kind = feats, model, state, reqs, blob, kind
        self.mapping = F.FeatureMapping(self.ctx, feats, model)
        self.state = F.DeviceState(self.ctx, self.mapping)
        t0 = time.perf_counter()
        self.state.put(state)
        self.state.flush()
        self.upload_s = time.perf_counter() - t0
        cls = mb.LightGBMBooster if kind == 0 else mb.XGBoostBooster
        self.booster = cls(self.ctx, blob, n_features=self.mapping.dim)
        self.rk = F.Ranker(self.mapping, self.state)
        self.arrays = self.mapping.pack_requests(reqs)
        self.rows = self.arrays["total_items"]
        self.dev = _dev(self.arrays)
        d = self.dev
        self.d_out = torch.empty(self.rows, dtype=torch.float64, device="cuda")
        self.d_order = torch.empty(self.rows, dtype=torch.int32, device="cuda")
        self.stream = torch.cuda.current_stream().cuda_stream
        opt = lambda k: d[k].data_ptr() if k in d else None  # noqa: E731
        self.batch = F.RankBatch(self.arrays["n_requests"], d["offsets"].data_ptr(), d["ids"].data_ptr(), d["users"].data_ptr(),
                                 d["sessions"].data_ptr(), d["req_f64"].data_ptr(), d["req_u64"].data_ptr(), d["req_vec"].data_ptr(),
                                 d["req_vp"].data_ptr(), opt("item_f64"), opt("tok_off"), opt("tok_hash"), opt("tok_w"))
        self.batch.max_items_per_request = self.items
        self.h2d = int(sum(v.nbytes for v in self.arrays.values() if isinstance(v, np.ndarray)))
        self.d2h = self.rows * 12

    def step(self, explain=False):
        import metarank_b200 as mb
        import torch
        if explain and getattr(self, "d_feat", None) is None:
            self.d_feat = torch.empty(self.rows * self.c["features"], dtype=torch.float64, device="cuda")
        mb._capi.check(mb._capi.lib().mr_rank_device(self.state._h, self.booster._h, C.byref(self.batch), C.c_int32(self.rows),
                                                     C.c_void_p(self.d_out.data_ptr()), C.c_void_p(self.d_order.data_ptr()),
                                                     C.c_void_p(self.d_feat.data_ptr()) if explain else None, C.c_void_p(self.stream)))

    def status(self):
        from metarank_b200 import features as F
        F.rank_device_status(self.state, self.stream)

    def prime(self):
        """The device entry point never retries: when the per-request tag tables outgrow the scratch pool the status call
        grows the pool and asks for a resubmit (the host entry point does that loop itself).  Settle it before timing."""
        import metarank_b200 as mb
        for _ in range(8):
            self.step()
            try:
                self.status()
                return
            except mb._capi.MrError as e:
                if "resubmit" not in str(e):
                    raise
        raise RuntimeError("tag-table scratch pool did not settle")

    def e2e_prepare(self, n_slots=1):
        import torch
        # page-locked request / response buffers, one response pair per in-flight call
        self._pins = []
        pinned = {}
        for k, v in self.arrays.items():
            if isinstance(v, np.ndarray) and v.size:
                flat = np.ascontiguousarray(v)
                as_i = flat.view(np.int64) if flat.dtype == np.uint64 else flat
                t = torch.from_numpy(as_i).pin_memory()
                self._pins.append(t)
                pinned[k] = t.numpy().view(flat.dtype).reshape(flat.shape)
            else:
                pinned[k] = v
        self.e2e_arrays = pinned
        n = self.arrays["total_items"]
        self.e2e_out = [(torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory())
                        for _ in range(n_slots)]

    def e2e_step(self, slot=0):
        sc, od = self.e2e_out[slot]
        return self.rk.rank_arrays(self.e2e_arrays, self.booster, want_order=True, out_scores=sc.numpy(), out_order=od.numpy())[:2]

    def parity(self, n_chk=2):
        from oracle import features_oracle as fo, oracle
        self.step()
        self.status()
        sc_d = self.d_out.cpu().numpy()
        od_d = self.d_order.cpu().numpy()
        mapping = fo.FeatureMapping(self.feats, self.model_names)
        ob = oracle.OracleBooster(self.kind, self.blob)
        offs = self.arrays["offsets"]
        feats = self.rk.make_query(self.reqs[:n_chk])
        f_ok = s_ok = o_ok = True
        for r in range(n_chk):
            want = fo.dense_matrix(mapping, self.reqs[r], self.state_dict)
            ws = ob.predictMat(want, *want.shape)
            f_ok &= bool(np.array_equal(feats[r], want, equal_nan=True))
            s_ok &= bool(np.array_equal(sc_d[offs[r]:offs[r + 1]], ws))
            o_ok &= bool(np.array_equal(od_d[offs[r]:offs[r + 1]], oracle.rank_order(ws)))
        self._parity_ref = (sc_d[:offs[n_chk]], od_d[:offs[n_chk]])
        return {"features_bit_identical": f_ok, "scores_bit_identical": s_ok, "ordering_identical": o_ok,
                "checked_items": int(offs[n_chk])}

    def latency(self):
        return None

    def cpu_setup(self):
        from oracle import features_oracle as fo, oracle
        self.make()  # host-side data only
        self.o_mapping = fo.FeatureMapping(self.feats, self.model_names)
        self.ob = oracle.OracleBooster(self.kind, self.blob)

    def cpu_run(self, n_requests, threads):
        from oracle import features_oracle as fo, oracle
        rows = 0
        for r in range(n_requests):
            X = fo.dense_matrix(self.o_mapping, self.reqs[r % len(self.reqs)], self.state_dict)
            sc = self.ob.predictMat(X, *X.shape, threads=threads)
            oracle.rank_order(sc)
            rows += X.shape[0]
        return rows

    cpu_what = ("feature oracle (pure-Python restatement of the extractors, one thread) + C oracle scorer (OpenMP over rows) "
                "+ stable ordering")

    def free(self):
        self.state.free(); self.mapping.free(); self.booster.free()


class C3(GenericWorkload):
    cat16 = True  # the model's one categorical column has 16 categories: small-categorical codes, resolved inside the level loop
    name = "C3"

    def make(self):
        from metarank_b200 import synth
        c = self.c
        self.feats, self.model_names = synth.ranklens_config()
        self.state_dict, item_ids, sessions = synth.ranklens_state(n_items=c["catalogue"], n_sessions=2000, seed=c["data_seed"])
        self.reqs = synth.ranklens_requests(item_ids, sessions, c["requests_per_step"], c["items"], seed=c["data_seed"] + 1 + self.rank)
        self.blob = synth.lightgbm_model_text(c["trees"], c["features"], seed=c["model_seed"], cat_features={7: 16})
        self.kind = 0

    def build(self):
        self.make()
        self.finish_build(self.feats, self.model_names, self.state_dict, self.reqs, self.blob, 0)


class C4(GenericWorkload):
    name = "C4"
    DIM = 384

    def make(self):
        from metarank_b200 import synth
        c = self.c
        rng = np.random.Generator(np.random.PCG64(c["data_seed"]))
        dim, n = self.DIM, c["catalogue"]
        feats = [dict(name="sim", type="field_match", rankingField="ranking.query", itemField="item.title",
                      method=dict(type="bi-encoder", dim=dim), distance="cos")]
        feats += [dict(name=f"n{k}", type="number", scope="item", source=f"metadata.n{k}") for k in range(15)]
        ids = [f"i{k}" for k in range(n)]
        E = rng.standard_normal((n, dim)).astype(np.float32)
        E /= np.linalg.norm(E, axis=1, keepdims=True)
        nums = rng.standard_normal((n, 15))
        state = {}
        for i, it in enumerate(ids):
            state[(("item", it), "sim")] = ("scalar", E[i].astype(np.float64))
            for k in range(15):
                state[(("item", it), f"n{k}")] = ("scalar", float(nums[i, k]))
        rq = np.random.Generator(np.random.PCG64(c["data_seed"] + 1 + self.rank))
        reqs = []
        for r in range(c["requests_per_step"]):
            pick = rq.choice(n, c["items"], replace=False)
            reqs.append(dict(event="ranking", id=f"r{r}", timestamp=0, user=None, session=None, fields=[("query", "q")],
                             embeddings={"sim": rq.standard_normal(dim).astype(np.float32)},
                             items=[dict(id=ids[int(j)], fields=[]) for j in pick]))
        self.feats, self.model_names, self.state_dict, self.reqs = feats, [f["name"] for f in feats], state, reqs
        self.blob = synth.xgboost_model_json(c["trees"], c["features"], depth=6, seed=c["model_seed"])
        self.kind = 1

    def build(self):
        self.make()
        self.finish_build(self.feats, self.model_names, self.state_dict, self.reqs, self.blob, 1)


WORKLOADS = {"C2": C2, "C3": C3, "C4": C4, "C5": C5}


# ======================================================================================= measurement
def device_time(w, steps, warmup, barrier, dist, world):
    """W warm-up steps, then exactly K timed steps between barrier + synchronize; CUDA events; max over ranks."""
    import torch
    stream = torch.cuda.current_stream()
    if hasattr(w, "prime"):
        w.prime()
    for _ in range(warmup):
        w.step()
    w.status()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    e0.record(stream)
    for _ in range(steps):
        w.step()
    e1.record(stream)
    barrier()
    t_end = time.perf_counter()
    total_ms = e0.elapsed_time(e1)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    rank_ms = [total_ms / steps]
    if world > 1:
        allt = torch.zeros(world, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allt, t)
        rank_ms = [float(x) / steps for x in allt.cpu()]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), rank_ms, t_begin, t_end


def kernel_profile(w, steps):
    """Per-kernel device time over `steps` steps from the library's own per-launch CUDA events."""
    import metarank_b200 as mb
    lib = mb._capi.lib()
    mb._capi.check(lib.mr_profile_begin())
    for _ in range(steps):
        w.step()
    buf = C.create_string_buffer(1 << 16)
    n = C.c_size_t()
    mb._capi.check(lib.mr_profile_end(buf, C.c_size_t(len(buf)), C.byref(n)))
    ks = json.loads(buf.value.decode())
    for k in ks:
        k["ms_per_step"] = k.pop("ms") / steps
        k["launches_per_step"] = k.pop("launches") / steps
    return ks


def scorer_roofline(w, kernels, step_ms, sm_mhz):
    """The dominant kernel against the pipe that bounds it (see the module docstring)."""
    import torch
    b = w.booster
    c = w.c
    peak_hbm, sm_max, peak_src = _peaks()
    clk = (sm_mhz or sm_max) * 1e6
    dom = max(kernels, key=lambda k: k["ms_per_step"])
    out = {"kernel": dom["kernel"], "kernel_ms": dom["ms_per_step"], "step_share": dom["ms_per_step"] / step_ms,
           "peak_source": peak_src}
    slim = dom["kernel"].startswith("gbdt_score_slim")
    leaves = dom["kernel"].startswith("gbdt_leaves")
    if not (slim or leaves or dom["kernel"].startswith("gbdt_score_compact")):
        out.update({"bound": "hbm", "achieved": None, "peak": peak_hbm, "unit": "GB/s", "frac": None, "traffic": None,
                    "note": "dominant kernel is not the compact scorer; see `kernels` for its HBM fraction"})
        return out
    # walk statistics on this very batch (first <= 262144 rows)
    rows = min(w.rows, 262144)
    if getattr(w, "d_feat", None) is None:
        w.step(explain=True)
        w.status()
    import metarank_b200 as mb
    lane, warp, wt = C.c_double(), C.c_double(), C.c_double()
    mb._capi.check(mb._capi.lib().mr_model_walk_stats(b._h, C.c_void_p(w.d_feat.data_ptr()), C.c_int32(rows), C.c_int32(c["features"]),
                                                      C.byref(lane), C.byref(warp), C.byref(wt), C.c_void_p(w.stream)))
    scale = w.rows / rows
    levels_per_wt = warp.value / wt.value          # level steps a warp issues per tree (deepest lane)
    dbar = lane.value / (rows * c["trees"])        # mean path per item per tree
    lanes_active = lane.value / (32.0 * warp.value)
    # wavefront floor of the lock-step walk: per level one node load and one code load (1 wavefront each); per tree the leaf
    # value (LDS.64: 1, or 2 once lanes sit on different leaves) and a quarter of a root-table load (LDS.128 per 4 trees);
    # the slim layout also reads the root entry (1) and its leaf value as two half-warp passes (2)
    # (the leaf-slot kernel of the low-latency path walks 8-byte nodes and stores a u16 slot per tree: 1 per warp-tree)
    # with the root table in the kernel's parameter space (<= 1920 trees) level 0 has no node load and no root-table load:
    # 2 per level step - 1 + the leaf value's 2
    root_tab = slim and c["trees"] <= 1920
    wavefronts = (2.0 * warp.value + (1.0 if root_tab else 3.25 if slim else 1.0 if leaves else 1.25) * wt.value) * scale
    t = dom["ms_per_step"] / 1e3
    smem_peak = SM_COUNT * 128.0 * clk / 1e9       # GB/s
    achieved = wavefronts * 128.0 / t / 1e9
    # the other pipe the walk leans on: issue slots (4 schedulers per SM, one warp instruction per clock each).  Instruction
    # floor of the slim kernel from its SASS (tests/test_capi_cpu.py guards the loop): 8 per level step inside the loop (10
    # with small categorical bitsets), and per warp-tree 13.25 with the parameter-space root table (level 0: 2 LDC.64, LOP3,
    # LDS, HSETP2, SEL, IMAD; leaf: BSYNC, BSSY, LOP3, LDS.64, DADD; 1.25 of loop control) — level 0 is outside the loop —
    # or 10.5 with the chunk-resident table (level 0 is a loop iteration)
    issue = None
    if slim:
        cat16 = bool(getattr(w, "cat16", False))
        per_level = 10.0 if cat16 else 8.0
        if root_tab:
            instrs = per_level * (warp.value - wt.value) + (13.25 + (2.0 if cat16 else 0.0)) * wt.value
        else:
            instrs = per_level * warp.value + 10.5 * wt.value
        instrs *= scale
        issue_peak = SM_COUNT * 4.0 * clk          # warp instructions per second
        issue = {"frac": instrs / t / issue_peak, "warp_instructions_floor": instrs, "peak_per_s": issue_peak,
                 "what": "warp instructions the walk needs at least (SASS counts x walk statistics) / kernel time, against SMs x 4 schedulers x SM clock"}
    b_item = 8 * c["features"] + c["trees"] * (dbar * 16 + 8) + 8
    static = {}
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tf):
        try:
            static = json.load(open(tf))
        except Exception:
            static = {}
    smem_frac = achieved / smem_peak
    by_issue = issue is not None and issue["frac"] > smem_frac
    out.update({
        "bound": "issue" if by_issue else "smem",
        "achieved": issue["warp_instructions_floor"] / t / 1e9 if by_issue else achieved,
        "peak": issue["peak_per_s"] / 1e9 if by_issue else smem_peak,
        "unit": "G warp-instructions/s" if by_issue else "GB/s",
        "frac": issue["frac"] if by_issue else smem_frac,
        "smem": {"frac": smem_frac, "achieved_gbs": achieved, "peak_gbs": smem_peak},
        "what": "the busier of the walk's two pipes — issue slots (`issue`) or the shared-memory crossbar (`smem`); both floors come "
                "from walk statistics counted on this batch.  shared-memory crossbar: wavefronts the lock-step walk needs at least (counted on this batch by "
                "mr_model_walk_stats: 2 per warp level step + per warp-tree 1 (4-byte nodes, root table in the parameter space), 3.25 (4-byte nodes) or 1.25 (8-byte nodes)) x 128 B / kernel time, against "
                "SMs x 128 B/clk x SM clock",
        "wavefronts_per_launch_floor": wavefronts, "warp_levels_per_tree": levels_per_wt, "mean_path": dbar,
        "lanes_active_of_32": 32.0 * lanes_active, "sm_clock_mhz": clk / 1e6,
        "issue": issue if issue is not None else {"frac": None, "what": "no instruction model for this kernel; see profiles/ (ncu smsp__issue_active)"},
        "traffic": static.get("dominant_kernel_c2_bytes_per_launch") if w.name == "C2" else None,
        "traffic_source": "static: profiles/traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum of one capture), not measured in this run",
        "ncu_static": {k: static[k] for k in ("smem_wavefronts_per_launch", "l1tex_pipe_pct", "issue_active_pct", "source") if k in static},
        "algorithmic_gbs": b_item * w.rows / t / 1e9, "algorithmic_bytes_per_item": b_item,
        "algorithmic_note": "SURVEY.md 8(d) B_item = 8F + T(16 d + 8) + 8: bytes the walk touches; they come from shared memory "
                            "after one TMA stage per chunk, so this exceeds the HBM peak and is NOT a roofline fraction",
        "hbm": {"peak": peak_hbm, "unit": "GB/s",
                "algorithmic_bytes_per_item": (2 * w.tile_cols + 8) if getattr(w, "tile_cols", None) else None,
                "what": "what the scorer must move through HBM per item: its u16 code tile in, one f64 score out"},
    })
    return out


def kernel_table(w, kernels):
    """Algorithmic HBM bytes per kernel launch for the HBM-bound kernels of the step."""
    peak_hbm, _, _ = _peaks()
    info = w.state.info()
    row_b = int(info.item_row_bytes)
    n = w.rows
    tc = getattr(w, "tile_cols", None)
    alg = {
        "lookup_kernel": n * (8 + 16 + 8 + 4 + 4),                       # id + probe (key, value) + request search + two outputs
        "code_gather_kernel": n * (2 * 2 * (tc or 0)) if tc else None,   # code row read + tile write
        "row_gather_kernel": n * (row_b + 2 * (tc or 0)) if tc else n * (row_b + 8 * w.c["features"]),
        "cosine_kernel": n * (8 * 384 + 16) if w.name == "C4" else None,
        "cosine_f32_kernel": n * (4 * 384 + 8 + 8 + 8) if w.name == "C4" else None,   # f32 row + bSum + row index + the raw cosine out
        "order_small_kernel": n * 12,
        "order_kernel": n * 12,
    }
    out = []
    for k in kernels:
        e = dict(k)
        a = alg.get(k["kernel"])
        if a and k["ms_per_step"] > 0:
            gbs = a / (k["ms_per_step"] / 1e3) / 1e9  # `a` = bytes per step over all launches of this kernel
            e.update({"bound": "hbm", "algorithmic_bytes": a, "achieved_gbs": gbs, "frac_hbm": gbs / peak_hbm})
        elif k["kernel"].startswith("gbdt_"):
            e["bound"] = "smem" if any(x in k["kernel"] for x in ("compact", "slim", "leaves")) else "latency"
        out.append(e)
    return sorted(out, key=lambda e: -e["ms_per_step"])


def cpu_baseline(w, budget_s=12.0):
    """The oracle port on a bounded sample of the same workload, all usable host cores."""
    w.cpu_setup()
    cands = sorted({_usable_cores(), os.cpu_count() or 1})
    n_probe = max(1, min(w.R, 64 if w.name == "C2" else 1))
    best, best_rate = cands[0], 0.0
    for c in cands:
        w.cpu_run(n_probe, c)
        t0 = time.perf_counter()
        rows = w.cpu_run(n_probe, c)
        rate = rows / (time.perf_counter() - t0)
        if rate > best_rate:
            best, best_rate = c, rate
    n_req = int(max(1, min(w.R if w.name != "C5" else 1, best_rate * budget_s / w.items)))
    t0 = time.perf_counter()
    rows = w.cpu_run(n_req, best)
    dt = time.perf_counter() - t0
    return {"value": rows / dt, "unit": UNIT, "cores": best, "kind": "port",
            "sample": f"{n_req} request(s) x {w.items} items: {w.cpu_what}"}


def run_reference(args, rank, world):
    """CPU arm: the oracle port on all host cores, bounded sample per step; same config block as the GPU arm."""
    if rank != 0:
        return
    w = WORKLOADS[args.config](None, 0, 1)
    w.cpu_setup()
    cands = sorted({_usable_cores(), os.cpu_count() or 1})
    n_probe = 64 if args.config == "C2" else 1
    best, best_rate = cands[0], 0.0
    for c in cands:
        w.cpu_run(n_probe, c)
        t0 = time.perf_counter()
        rows = w.cpu_run(n_probe, c)
        rate = rows / (time.perf_counter() - t0)
        if rate > best_rate:
            best, best_rate = c, rate
    total_steps = max(1, args.steps + args.warmup)
    n_req = int(max(1, min(w.R if args.config != "C5" else 1, best_rate * 60.0 / total_steps / w.items)))
    for _ in range(args.warmup):
        w.cpu_run(n_req, best)
    t0 = time.perf_counter()
    rows = 0
    for _ in range(args.steps):
        rows += w.cpu_run(n_req, best)
    dt = time.perf_counter() - t0
    v = rows / dt
    sample = f"{n_req} request(s) x {w.items} items per step (same generator/seed as the GPU arm): {w.cpu_what}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.config == "C5" else "weak",
        "vs_baseline": None, "dtype": "f32" if args.config == "C4" else "f64", "data": "synthetic",
        "config": config_block(args.config, args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": best, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference scorer (ltrlib -> LightGBM / XGBoost JNI) is not installable here (no JVM, no jars, no wheels: "
                "profiles/probe_r2_gbdt_libs.txt); this is the C oracle port of its arithmetic, OpenMP over rows like "
                "LightGBM's predictor, fed by a parallel row gather",
    }), flush=True)


def measure(w, args, rank, world, dist, barrier, full=True):
    """Timed region + (full) kernel profile, roofline, e2e, latency, parity, CPU baseline."""
    import torch
    import metarank_b200 as mb
    lib = mb._capi.lib()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    sampler = ClockSampler(local)
    if rank == 0 and full:
        sampler.start()
    launches0 = lib.mr_kernel_launches()
    total_ms, rank_ms, t_begin, t_end = device_time(w, args.steps, args.warmup, barrier, dist, world)
    gpu_launches = lib.mr_kernel_launches() - launches0
    value = w.total_rows() * args.steps / (total_ms / 1e3)
    res = {"value": value, "ms_per_step": total_ms / args.steps, "rank_ms_per_step": rank_ms, "gpu_launches": int(gpu_launches)}
    if full:
        # keep the same loop running until the clock record spans >= 1.2 s of load (nvidia-smi samples every 20 ms);
        # the repeat count derives from the max-over-ranks time, so every rank runs the same number of (collective) steps
        extra = max(0, int((1200.0 - total_ms) / (total_ms / args.steps)) + 1) if total_ms < 1200.0 else 0
        for _ in range(extra):
            w.step()
        torch.cuda.synchronize()
        barrier()
        t_end = time.perf_counter()
        res["clocks"] = sampler.stop(t_begin, t_end) if rank == 0 else None
        if res["clocks"]:
            res["clocks"]["what"] = "sampled over the timed region and the repeats of the same loop that follow it"
    # e2e: the same metric through the host-buffer API.  Calls are issued the way a server issues them — from worker threads,
    # here two, each with its own response buffers — so one call's copies overlap the other's kernels (the library runs
    # concurrent calls on separate streams); the mega-request path is a collective over the ranks and stays one call at a time.
    import threading
    n_slots = 1 if w.name == "C5" else 2
    w.e2e_prepare(n_slots) if w.name != "C5" else w.e2e_prepare()
    e2e_steps = max(4, min(args.steps, 10))
    e2e_steps += e2e_steps % n_slots
    for _ in range(2):
        sc_h, ord_h = w.e2e_step(0)
    t0 = time.perf_counter()
    for _ in range(3):
        sc_h, ord_h = w.e2e_step(0)
    serial_s = (time.perf_counter() - t0) / 3
    results = [None] * n_slots
    errors = []

    def worker(slot, n):
        try:
            for _ in range(n):
                results[slot] = w.e2e_step(slot)
        except Exception as ex:  # surfaced below: a failed call must not pass as a fast one
            errors.append(ex)

    def run_workers(per_slot):
        if n_slots == 1:
            worker(0, per_slot)
            return
        ths = [threading.Thread(target=worker, args=(k, per_slot)) for k in range(n_slots)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()

    run_workers(3)  # untimed: every worker's stream, scratch and staging buffers exist before the clock starts
    barrier()
    t0 = time.perf_counter()
    run_workers(e2e_steps // n_slots)
    e2e_s = time.perf_counter() - t0
    if errors:
        raise errors[0]
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["parity"] = w.parity()
    got, got_order = w._parity_ref
    n = len(got)
    e2e_ok = all(bool(np.array_equal(r[0][:n], got) and np.array_equal(r[1][:n], got_order)) for r in results)
    res["e2e"] = {"value": w.total_rows() * e2e_steps / float(t.item()), "unit": UNIT, "h2d_bytes_per_step": w.h2d,
                  "d2h_bytes_per_step": w.d2h, "steps": e2e_steps, "parity_ok": e2e_ok, "calls_in_flight": n_slots,
                  "one_call_at_a_time": w.total_rows() / world / serial_s,
                  "what": "host buffers in -> scores + order in host buffers, H2D and D2H copies inside the timer; "
                          f"{n_slots} worker thread(s) per GPU issuing calls (one_call_at_a_time: a single thread, per GPU)"}
    if w.name == "C5":
        res["latency"] = w.latency()  # collective: every rank takes part
    elif rank == 0 and full:
        res["latency"] = w.latency()
    barrier()
    ks = kernel_profile(w, max(2, min(args.steps, 10)))
    w.status()
    if rank == 0:
        try:
            w.tile_cols = int(round(w.booster.codes_bytes(1 << 16) / float(1 << 16) / 2.0)) or None  # u16 codes per item (whole CTA tiles: ask for many rows)
        except Exception:
            w.tile_cols = None
        res["kernels"] = kernel_table(w, ks)
        dom = max(res["kernels"], key=lambda k: k["ms_per_step"])
        walk = None
        if dom["kernel"].startswith(("gbdt_score_slim", "gbdt_score_compact", "gbdt_leaves")):
            try:
                walk = scorer_roofline(w, ks, total_ms / args.steps, (res.get("clocks") or {}).get("sm_mhz"))
            except Exception as ex:  # e.g. walk statistics not available for this model form
                walk = None
                res["roofline_note"] = f"walk statistics unavailable: {type(ex).__name__}: {ex}"[:200]
        if walk is not None:
            res["roofline"] = walk
            for k in res["kernels"]:  # the per-kernel table names the same limiter as the roofline block
                if k["kernel"] == walk.get("kernel") and walk.get("bound"):
                    k["bound"] = walk["bound"]
        else:
            res["roofline"] = {"bound": dom.get("bound"), "kernel": dom["kernel"], "kernel_ms": dom["ms_per_step"],
                               "achieved": dom.get("achieved_gbs"), "peak": _peaks()[0], "unit": "GB/s", "frac": dom.get("frac_hbm"),
                               "traffic": None, "step_share": dom["ms_per_step"] / (total_ms / args.steps)}
        res["cpu_baseline"] = cpu_baseline(w, 12.0 if full else 4.0)
        if w.name == "C2" and full:
            try:
                res["churn"] = w.churn()
            except Exception as ex:
                res["churn"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
        if w.name == "C4":
            try:
                res["query_encoder"] = query_encoder_block(w.ctx, w.arrays["n_requests"])
            except Exception as ex:  # the encoder block must never cost the config's own line
                res["query_encoder"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
        info = w.state.info()
        res["state"] = {"items": int(info.rows[1]), "device_bytes": int(info.device_bytes),
                        "item_row_bytes": int(info.item_row_bytes), "upload_s": w.upload_s}
    return res


def query_encoder_block(ctx, n_queries, seq=16):
    """BASELINE configs[3]'s other half: the bi-encoder QUERY forward (OnnxBiEncoder.embed) for the batch's
    requests — e5-small's shape (12 layers x 384, 12 heads, FFN 1536, 30522-token vocabulary), synthetic weights, seq tokens
    per query.  Tensor-core bound: achieved = dense-layer flops / device time against the measured bf16 matmul peak."""
    import torch
    from metarank_b200 import encoder as E
    from oracle import encoder_oracle as eo
    layers, hidden, inter, heads = 12, 384, 1536, 12
    w = E.synthetic_bert_weights(hidden=hidden, layers=layers, intermediate=inter, seed=4)
    enc = E.OnnxBiEncoder(ctx, E.write_safetensors(w), n_heads=heads)
    rng = np.random.default_rng(11)
    ids = rng.integers(0, 30522, (n_queries, seq))
    lens = rng.integers(3, seq + 1, n_queries)
    lens[0] = seq
    mask = (np.arange(seq)[None, :] < lens[:, None]).astype(np.int64)
    tt = np.zeros_like(ids)
    dev = torch.device("cuda", torch.cuda.current_device())
    d_ids, d_tt, d_mask = (torch.from_numpy(x).to(dev) for x in (ids, tt, mask))
    out = torch.empty(n_queries, hidden, device=dev)
    out64 = torch.empty(n_queries, hidden, device=dev, dtype=torch.float64)
    st = torch.cuda.Stream()  # not the legacy stream: the library replays its CUDA graph of the forward on real streams only
    torch.cuda.synchronize()

    def timed(b, iters):
        for _ in range(3):
            enc.embed_device(d_ids.data_ptr(), d_tt.data_ptr(), d_mask.data_ptr(), b, seq, out.data_ptr(), out64.data_ptr(), st.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(st)
        for _ in range(iters):
            enc.embed_device(d_ids.data_ptr(), d_tt.data_ptr(), d_mask.data_ptr(), b, seq, out.data_ptr(), out64.data_ptr(), st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    ms_batch = timed(n_queries, 5)
    ms_one = timed(1, 50)
    n_chk = min(n_queries, 6)
    got = enc.embed(ids[:n_chk], tt[:n_chk], mask[:n_chk])
    want = eo.embed(w, ids[:n_chk], tt[:n_chk], mask[:n_chk], n_heads=heads)
    cos = lambda a, b: (a.astype(np.float64) * b).sum(-1) / np.sqrt((a.astype(np.float64) ** 2).sum(-1) * (b.astype(np.float64) ** 2).sum(-1))  # noqa: E731
    pair = float(np.abs(cos(got[:-1], got[1:]) - cos(want[:-1], want[1:])).max())
    flops = 2.0 * n_queries * seq * layers * (4 * hidden * hidden + 2 * hidden * inter)
    peak = 1699.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
    except Exception:
        pass
    tf = flops / ms_batch / 1e9
    enc.close()
    return {"what": "bi-encoder query forward for the batch's requests: token ids resident -> f32 + f64 embeddings (mr_encoder_embed_device)",
            "model": "e5-small shape, synthetic weights: 12 layers x 384 hidden, 12 heads, FFN 1536", "queries": int(n_queries), "seq": seq,
            "ms": ms_batch, "queries_per_s": n_queries / ms_batch * 1e3, "one_query_ms": ms_one,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS 8192^3, back to back)", "traffic": None,
                         "flops": "dense layers only: 2 * tokens * layers * (4 H^2 + 2 H I)"},
            "parity": {"vs": "fp32 restatement (oracle/encoder_oracle.py)", "checked_queries": int(n_chk),
                       "max_abs_err": float(np.abs(got - want).max()), "pair_cosine_err": pair, "tolerance_pair_cosine": 1e-3,
                       "ok": bool(pair < 1e-3)}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="C2", choices=sorted(CONFIG))
    ap.add_argument("--no-extras", action="store_true", help="default C2 run: skip the short C3 / C4 / C5 measurements")
    ap.add_argument("--requests-per-step", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)
    if args.requests_per_step > 0:
        CONFIG[args.config]["requests_per_step"] = args.requests_per_step

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
    w = WORKLOADS[args.config](ctx, rank, world)
    w.build()
    res = measure(w, args, rank, world, dist, barrier, full=True)
    w.free()

    extras = {}
    if args.config == "C2" and not args.no_extras:
        short = argparse.Namespace(steps=20, warmup=3)
        for name in (["C3", "C4", "C5"] if world == 1 else ["C5"]):
            try:
                x = WORKLOADS[name](ctx, rank, world)
                x.build()
                r = measure(x, short, rank, world, dist, barrier, full=False)
                x.free()
                if rank == 0:
                    e = {"config": config_block(name, world), "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"],
                         "steps": short.steps, "n_gpus": world, "scaling": "strong" if name == "C5" else "weak",
                         "dtype": "f32" if name == "C4" else "f64", "e2e": r["e2e"], "parity": r["parity"],
                         "kernels": r["kernels"], "roofline": r["roofline"], "cpu_baseline": r["cpu_baseline"],
                         "gpu_launches": r["gpu_launches"]}
                    if r.get("latency"):
                        e["latency"] = r["latency"]
                    if r.get("query_encoder"):
                        e["query_encoder"] = r["query_encoder"]
                    extras[name] = e
            except Exception as ex:  # an extra must never cost the headline line
                if rank == 0:
                    extras[name] = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    if rank == 0:
        out = {
            "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if args.config == "C5" else "weak", "vs_baseline": None,
            "dtype": "f32" if args.config == "C4" else "f64", "data": "synthetic",
            "config": config_block(args.config, world), "e2e": res["e2e"], "gpu_launches": res["gpu_launches"],
            "rank_ms_per_step": res["rank_ms_per_step"], "roofline": res["roofline"], "kernels": res["kernels"],
            "cpu_baseline": res["cpu_baseline"], "clocks": res.get("clocks"), "latency": res.get("latency"),
            "parity": res["parity"], "state": res["state"],
        }
        if res.get("query_encoder"):
            out["query_encoder"] = res["query_encoder"]
        if res.get("churn"):
            out["churn"] = res["churn"]
        if extras:
            out["other_configs"] = extras
        print(json.dumps(out), flush=True)

    if world > 1:
        # the other ranks wait here while rank 0 writes its line: their NCCL teardown messages (NCCL_DEBUG=INFO prints to
        # stdout) must not land in the middle of it
        dist.barrier()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
