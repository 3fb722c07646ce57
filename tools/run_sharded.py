"""BASELINE config #5 on N GPUs: one 10 000-item mega-request, 64 features, 2000-tree ensemble,
item-sharded across the ranks with an NCCL all-gather of the scores (metarank_b200/sharded.py).
Launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_sharded.py"""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import metarank_b200 as mb
from metarank_b200 import features as F, sharded, synth
from oracle import oracle

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = mb.Context(local)
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

def one():
    score_slice, n = sharded.cuda_slice_scorer(rk, booster, arrays)   # every rank: full assembly on device
    return sharded.ShardedScorer(score_slice, ctx.rank_order).rerank(n)

for _ in range(3): scores, order = one()
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
K = 10
for _ in range(K): scores, order = one()
dt = (time.perf_counter() - t0) / K
want = oracle.OracleBooster(0, blob).predictMat(cat[pick], NI, NF, threads=os.cpu_count())
ok = bool(np.array_equal(scores, want) and np.array_equal(order, oracle.rank_order(want)))
if rank == 0:
    print(json.dumps({"config": "C5: 10000-item request x 64 features x 2000 trees", "n_gpus": world, "ms_per_request": dt * 1e3,
                      "items_per_s": NI / dt, "scores_bit_identical": ok}), flush=True)
if world > 1: dist.destroy_process_group()
