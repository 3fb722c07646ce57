"""Runs the GBDT scoring kernel a few times on one config (target for ncu)."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import metarank_b200 as mb
from metarank_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="C2")
ap.add_argument("--rows", type=int, default=1 << 20)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
c = synth.CONFIGS[a.cfg]
if c["kind"] == "lightgbm":
    kind, blob = 0, synth.lightgbm_model_text(c["trees"], c["features"], c["leaves"], c["max_depth"], seed=1236)
else:
    kind, blob = 1, synth.xgboost_model_json(c["trees"], c["features"], c["depth"], seed=1238)
F = c["features"]
ctx = mb.Context(0)
b = mb.B200Booster(ctx, blob, kind=kind)
for o in a.opt:
    k, v = o.split("=")
    b.set_option(k, int(v))
X = torch.from_numpy(synth.feature_matrix(a.rows, F, seed=44)).cuda()
O = torch.empty(a.rows, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
b.predict_device(X.data_ptr(), a.rows, F, O.data_ptr(), st)
torch.cuda.synchronize()
e0.record()
for _ in range(a.iters):
    b.predict_device(X.data_ptr(), a.rows, F, O.data_ptr(), st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
print(f"{a.cfg} rows={a.rows} {ms:.3f} ms  {a.rows / ms / 1e3:.1f} M items/s", flush=True)
