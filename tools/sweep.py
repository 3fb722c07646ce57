"""Scratch sweep of kernel variants on the GPU (not part of the product)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import metarank_b200 as mb
from metarank_b200 import synth
from oracle import oracle

ctx = mb.Context(0)
cfgs = [("C2", 0, synth.lightgbm_model_text(500, 30, seed=1236), 30)]
if len(sys.argv) > 1 and sys.argv[1] == "all":
    cfgs += [("C5", 0, synth.lightgbm_model_text(2000, 64, seed=1239), 64),
             ("C4x", 1, synth.xgboost_model_json(200, 16, depth=6, seed=1238), 16)]
rows = 1638400
for name, kind, blob, F in cfgs:
    X = synth.feature_matrix(rows, F, seed=44)
    want = oracle.OracleBooster(kind, blob).predictMat(X[:4096], 4096, F, threads=0)
    dX = torch.from_numpy(X).cuda()
    dO = torch.empty(rows, dtype=torch.float64, device='cuda')
    b = mb.B200Booster(ctx, blob, kind=kind)
    print(name, "mean path", b.mean_path(X[:4096], 4096, F), flush=True)
    st = torch.cuda.current_stream().cuda_stream
    for chunk_kb in (16, 24):
        b.set_option("chunk_kb", chunk_kb)
        if True:
            for threads in [0] + list(range(320, 1025, 32)):
                variant, ilp = 4, 1
                b.set_option("threads", threads); b.set_option("variant", variant);
                try:
                    n_codes = b.codes_bytes(rows)
                    d_codes = torch.empty(n_codes, dtype=torch.uint8, device="cuda")
                    b.bin_device(dX.data_ptr(), rows, F, d_codes.data_ptr(), st)
                    for _ in range(2): b.score_codes_device(d_codes.data_ptr(), rows, dO.data_ptr(), st)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                    e0.record()
                    for _ in range(5): b.score_codes_device(d_codes.data_ptr(), rows, dO.data_ptr(), st)
                    e1.record(); torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / 5
                    ok = np.array_equal(dO[:4096].cpu().numpy(), want)
                    print(json.dumps(dict(cfg=name, chunk_kb=chunk_kb, nchunks=b.info().n_chunks, threads=threads,
                                          ms=round(ms, 3), Mitems_s=round(rows / ms / 1e3, 1), ok=bool(ok))), flush=True)
                except Exception as ex:
                    print("ERR", chunk_kb, threads, ex, flush=True)
    b.free()
