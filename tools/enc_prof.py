"""One forward of the MiniLM-shaped encoder at a given batch: the target of `ncu -k regex:<kernel>`."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from metarank_b200.booster import Context
from metarank_b200 import encoder as E

B, S = (int(x) for x in sys.argv[1:3]) if len(sys.argv) > 2 else (16384, 16)
ctx = Context(0)
dev = torch.device("cuda:0")
enc = E.OnnxBiEncoder(ctx, E.write_safetensors(E.synthetic_bert_weights(layers=2, seed=1)), n_heads=12)
rng = np.random.default_rng(0)
ids = torch.from_numpy(rng.integers(0, 30522, (B, S))).to(dev)
mask = torch.ones(B, S, dtype=torch.int64, device=dev); tt = torch.zeros(B, S, dtype=torch.int64, device=dev)
out = torch.empty(B, 384, device=dev)
for _ in range(2):
    enc.embed_device(ids.data_ptr(), tt.data_ptr(), mask.data_ptr(), B, S, out.data_ptr(), 0, 0)
torch.cuda.synchronize()
