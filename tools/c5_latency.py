"""BASELINE config #5 on one GPU: where the time of a 10 000-item x 64-feature x 2000-tree request goes, and the
crossover between the tree-parallel (latency) scorer and the thread-per-item (throughput) scorer.
  python tools/c5_latency.py [sweep]"""
import json, os, sys, threading, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # members of a group share one device here: one hardware queue per stream
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import metarank_b200 as mb
from metarank_b200 import features as F, sharded, synth
from oracle import oracle

ctx = mb.Context(0)
stream = torch.cuda.current_stream().cuda_stream


def ev_time(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def sweep(name, blob, nf):
    b = mb.LightGBMBooster(ctx, blob, n_features=nf)
    for rows in (1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072):
        X = torch.from_numpy(synth.feature_matrix(rows, nf, seed=rows)).cuda()
        codes = torch.empty(b.codes_bytes(rows), dtype=torch.uint8, device="cuda")
        out = torch.empty(rows, dtype=torch.float64, device="cuda")
        b.bin_device(X.data_ptr(), rows, nf, codes.data_ptr(), stream)
        res = {}
        for label, lr in (("throughput", 1), ("latency", 1 << 20)):
            b.set_option("latency_rows", lr)
            try:
                res[label] = ev_time(lambda: b.score_codes_device(codes.data_ptr(), rows, out.data_ptr(), stream), n=10)
            except mb.MrError as e:
                res[label] = str(e)[:60]
        print(json.dumps({"sweep": name, "rows": rows, **res}), flush=True)
    b.free()


if "sweep" in sys.argv:
    sweep("C2 model 500 trees x 30", synth.lightgbm_model_text(500, 30, 16, 8, seed=1236), 30)
    sweep("C5 model 2000 trees x 64", synth.lightgbm_model_text(2000, 64, seed=1239), 64)

NF, NT, NI, CAT = 64, 2000, 10_000, 50_000
names = [f"f{j}" for j in range(NF)]
fm = F.FeatureMapping(ctx, [dict(name=n, type="number", scope="item", source=f"metadata.{n}") for n in names], names)
st = F.DeviceState(ctx, fm)
cat = synth.feature_matrix(CAT, NF, seed=47)
ids = (np.arange(1, CAT + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
st.put_packed(F.pack_number_columns(names, ids, cat)); st.flush()
blob = synth.lightgbm_model_text(NT, NF, seed=1239)
booster = mb.LightGBMBooster(ctx, blob, n_features=NF)
pick = np.random.Generator(np.random.PCG64(48)).choice(CAT, NI, replace=False)
rk = F.Ranker(fm, st)
arrays = dict(offsets=np.array([0, NI], dtype=np.int32), ids=ids[pick], users=np.zeros(1, dtype=np.uint64),
              sessions=np.zeros(1, dtype=np.uint64), req_f64=np.zeros((1, 1)), req_u64=np.zeros((1, 1), dtype=np.uint64),
              req_vec=np.zeros((1, 1), dtype=np.float32), req_vp=np.zeros((1, 1), dtype=np.uint8), item_f64=None,
              n_requests=1, total_items=NI)
want = oracle.OracleBooster(0, blob).predictMat(cat[pick], NI, NF, threads=os.cpu_count())
d_ids = torch.from_numpy(ids[pick].view(np.int64)).cuda()
d_offs = torch.tensor([0, NI], dtype=torch.int32, device="cuda")
d_sc = torch.empty(NI, dtype=torch.float64, device="cuda"); d_od = torch.empty(NI, dtype=torch.int32, device="cuda")
for label, lr in (("throughput scorer (round 1 path)", 1), ("latency scorer", 0)):
    booster.set_option("latency_rows", lr)
    for _ in range(3): sc, od, _ = rk.rank_arrays(arrays, booster)
    ts = []
    for _ in range(30):
        t0 = time.perf_counter(); sc, od, _ = rk.rank_arrays(arrays, booster); ts.append(time.perf_counter() - t0)
    dev = ev_time(lambda: F.rank_device(st, booster, 1, NI, d_offs.data_ptr(), d_ids.data_ptr(), d_sc.data_ptr(), d_od.data_ptr(), 0, stream, max_items=NI))
    ok = bool(np.array_equal(sc, want) and np.array_equal(od, oracle.rank_order(want)) and np.array_equal(d_sc.cpu().numpy(), want)
              and np.array_equal(d_od.cpu().numpy(), od))
    print(json.dumps({"c5_one_gpu": label, "mr_rank_p50_ms": float(np.median(ts) * 1e3), "device_ms": dev, "parity": ok}), flush=True)

# the group machinery on ONE device (members share it): what the exchange itself costs
for world in (1, 2, 4, 8):
    members = [sharded.Group(ctx, r, world, NI) for r in range(world)]
    sharded.Group.connect_local(members)
    res = [None] * world
    def go(k, n):
        for _ in range(n): res[k] = members[k].rank_arrays(st, booster, arrays)
    def run(n):
        th = [threading.Thread(target=go, args=(k, n)) for k in range(world)]
        [t.start() for t in th]; [t.join() for t in th]
    run(3)
    t0 = time.perf_counter(); run(20); dt = (time.perf_counter() - t0) / 20
    ok = all(np.array_equal(r[0], want) and np.array_equal(r[1], oracle.rank_order(want)) for r in res)
    print(json.dumps({"c5_group_on_one_device": world, "ms_per_request": dt * 1e3, "parity": bool(ok)}), flush=True)
    [m.free() for m in members]
