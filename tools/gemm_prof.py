"""One dense-layer shape, a few launches: the target of `ncu -k regex:encoder_gemm`."""
import sys
import torch
sys.path.insert(0, ".")
from metarank_b200.booster import Context
from metarank_b200 import encoder as E

M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (262144, 1152, 384)
ctx = Context(0)
dev = torch.device("cuda:0")
a = torch.randn(M, K, device=dev).half(); w = torch.randn(N, K, device=dev).half(); b = torch.randn(N, device=dev)
o16 = torch.empty(M, N, device=dev, dtype=torch.half)
for _ in range(3):
    E.gemm_f16_device(ctx, a.data_ptr(), w.data_ptr(), b.data_ptr(), 0, 0, o16.data_ptr(), M, N, K, False, 0)
torch.cuda.synchronize()
