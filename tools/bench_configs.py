"""Throughput of the other BASELINE configs on one GPU (not the headline bench):
  C3  ranklens feature set (24 columns: numbers, index string, normalized + field-scoped rate, interacted_with
      over 4 fields, position, 5 diversity features), 1000-item requests, 500-tree LightGBM
  C4  bi-encoder cosine (384-d) + 15 numbers, 256-item requests, 200-tree XGBoost
Prints one JSON line per config: items/s device-side via mr_rank (host buffers), per-kernel launch list is
taken separately with ncu."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import metarank_b200 as mb
from metarank_b200 import features as F, synth
from oracle import features_oracle as fo, oracle

ctx = mb.Context(0)
which = sys.argv[1:] or ["C3", "C4"]


def run(name, feats, model, state, reqs, booster, blob, kind, n_check=2):
    fm = F.FeatureMapping(ctx, feats, model)
    ds = F.DeviceState(ctx, fm)
    t0 = time.perf_counter(); ds.put(state); ds.flush(); t_up = time.perf_counter() - t0
    rk = F.Ranker(fm, ds)
    arrays = fm.pack_requests(reqs)
    N = arrays["total_items"]
    for _ in range(3): sc, od, _ = rk.rank_arrays(arrays, booster, want_order=True)
    K = 10
    t0 = time.perf_counter()
    for _ in range(K): sc, od, _ = rk.rank_arrays(arrays, booster, want_order=True)
    dt = (time.perf_counter() - t0) / K
    # device-resident
    d = {k: torch.from_numpy(np.ascontiguousarray(v).view(np.int64) if v.dtype == np.uint64 else np.ascontiguousarray(v)).cuda()
         for k, v in arrays.items() if isinstance(v, np.ndarray)}
    d_s = torch.empty(N, dtype=torch.float64, device="cuda"); d_o = torch.empty(N, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    b = F.RankBatch(arrays["n_requests"], d["offsets"].data_ptr(), d["ids"].data_ptr(), d["users"].data_ptr(), d["sessions"].data_ptr(),
                    d["req_f64"].data_ptr(), d["req_u64"].data_ptr(), d["req_vec"].data_ptr(), d["req_vp"].data_ptr(),
                    d["item_f64"].data_ptr() if "item_f64" in d else None)
    import ctypes as C
    def dev_step():
        mb._capi.check(mb._capi.lib().mr_rank_device(ds._h, booster._h, C.byref(b), C.c_int32(N), C.c_void_p(d_s.data_ptr()),
                                                     C.c_void_p(d_o.data_ptr()), None, C.c_void_p(st)))
    for _ in range(3): dev_step()
    F.rank_device_status(ds, st)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(K): dev_step()
    e1.record(); torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / K
    # parity on the first requests
    mapping = fo.FeatureMapping(feats, model)
    ob = oracle.OracleBooster(kind, blob)
    offs = arrays["offsets"]; ok = True
    for r in range(n_check):
        want = fo.dense_matrix(mapping, reqs[r], state)
        ws = ob.predictMat(want, *want.shape)
        ok &= bool(np.array_equal(sc[offs[r]:offs[r + 1]], ws) and np.array_equal(od[offs[r]:offs[r + 1]], oracle.rank_order(ws)))
        ok &= bool(np.array_equal(d_s[offs[r]:offs[r + 1]].cpu().numpy(), ws))
    print(json.dumps({"config": name, "requests": len(reqs), "items": N, "cols": fm.dim, "e2e_ms": dt * 1e3,
                      "e2e_items_per_s": N / dt, "device_ms": dev_ms, "device_items_per_s": N / dev_ms * 1e3,
                      "parity_ok": ok, "state_upload_s": t_up}), flush=True)
    ds.free(); fm.free()


if "C3" in which:
    feats, model = synth.ranklens_config()
    state, item_ids, sessions = synth.ranklens_state(n_items=20000, n_sessions=2000, seed=45)
    reqs = synth.ranklens_requests(item_ids, sessions, 256, 1000, seed=46)
    blob = synth.lightgbm_model_text(500, 24, seed=1237, cat_features={7: 16})
    run("C3 ranklens 1000-item requests, 24 cols, 500-tree LightGBM", feats, model, state, reqs,
        mb.LightGBMBooster(ctx, blob, n_features=24), blob, 0)
if "C4" in which:
    rng = np.random.Generator(np.random.PCG64(77)); dim = 384
    feats = [dict(name="sim", type="field_match", rankingField="ranking.query", itemField="item.title",
                  method=dict(type="bi-encoder", dim=dim), distance="cos")]
    feats += [dict(name=f"n{k}", type="number", scope="item", source=f"metadata.n{k}") for k in range(15)]
    model = [f["name"] for f in feats]
    ids = [f"i{k}" for k in range(20000)]
    E = rng.standard_normal((20000, dim)).astype(np.float32); E /= np.linalg.norm(E, axis=1, keepdims=True)
    state = {}
    for i, it in enumerate(ids):
        state[(("item", it), "sim")] = ("scalar", E[i].astype(np.float64))
        for k in range(15): state[(("item", it), f"n{k}")] = ("scalar", float(rng.standard_normal()))
    reqs = []
    for r in range(512):
        pick = rng.choice(20000, 256, replace=False)
        reqs.append(dict(event="ranking", id=f"r{r}", timestamp=0, user=None, session=None, fields=[("query", "q")],
                         embeddings={"sim": rng.standard_normal(dim).astype(np.float32)},
                         items=[dict(id=ids[int(j)], fields=[]) for j in pick]))
    blob = synth.xgboost_model_json(200, 16, depth=6, seed=1238)
    run("C4 cosine(384)+15 numbers, 256-item requests, 200-tree XGBoost", feats, model, state, reqs,
        mb.XGBoostBooster(ctx, blob, n_features=16), blob, 1)
