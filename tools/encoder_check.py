"""First-light + timing of the tcgen05 dense layer and the bi-encoder forward on a B200 (run under gpurun)."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from metarank_b200.booster import Context
from metarank_b200 import encoder as E
from oracle import encoder_oracle as eo

ctx = Context(0)
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
res = {"gemm": [], "forward": []}

def gemm_case(M, N, K, bias=True, resid=False, gelu=False, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = (torch.randn(M, K, generator=g) * 0.5).half().to(dev)
    w = (torch.randn(N, K, generator=g) * 0.05).half().to(dev)
    b = torch.randn(N, generator=g).to(dev) if bias else None
    r = torch.randn(M, N, generator=g).to(dev) if resid else None
    o32 = torch.full((M, N), float("nan"), device=dev)
    o16 = torch.full((M, N), float("nan"), device=dev, dtype=torch.half)
    E.gemm_f16_device(ctx, a.data_ptr(), w.data_ptr(), b.data_ptr() if bias else 0, r.data_ptr() if resid else 0, o32.data_ptr(), o16.data_ptr(), M, N, K, gelu)
    torch.cuda.synchronize()
    ref = a.double() @ w.double().T
    if bias: ref = ref + b.double()
    if gelu: ref = torch.nn.functional.gelu(ref)
    if resid: ref = ref + r.double()
    err = (o32.double() - ref).abs().max().item()
    err16 = (o16.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    ok = err < 2e-4 * max(scale, 1) and err16 < 2e-3 * max(scale, 1)
    res["gemm"].append({"M": M, "N": N, "K": K, "gelu": gelu, "resid": resid, "max_err_f32": err, "max_err_f16": err16, "ref_max": scale, "ok": ok})
    print("gemm", M, N, K, "gelu" if gelu else "", "resid" if resid else "", "err", err, err16, "scale", scale, "OK" if ok else "FAIL", flush=True)
    return ok

allok = True
for (M, N, K) in [(128, 128, 64), (128, 128, 384), (1, 384, 384), (16, 1152, 384), (100, 1536, 384), (300, 384, 1536), (257, 64, 128), (4096, 1152, 384), (1000, 192, 448)]:
    allok &= gemm_case(M, N, K)
allok &= gemm_case(200, 1536, 384, gelu=True)
allok &= gemm_case(200, 384, 1536, resid=True)
allok &= gemm_case(77, 384, 384, bias=False)

def gemm_time(M, N, K, iters=20):
    a = torch.randn(M, K, device=dev).half(); w = torch.randn(N, K, device=dev).half(); b = torch.randn(N, device=dev)
    o16 = torch.empty(M, N, device=dev, dtype=torch.half)
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3): E.gemm_f16_device(ctx, a.data_ptr(), w.data_ptr(), b.data_ptr(), 0, 0, o16.data_ptr(), M, N, K, False, s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): E.gemm_f16_device(ctx, a.data_ptr(), w.data_ptr(), b.data_ptr(), 0, 0, o16.data_ptr(), M, N, K, False, s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    for _ in range(3): torch.matmul(a, w.T)
    torch.cuda.synchronize(); t0.record()
    for _ in range(iters): torch.matmul(a, w.T)
    t1.record(); torch.cuda.synchronize()
    cms = t0.elapsed_time(t1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    print(f"gemm time {M}x{N}x{K}: {ms*1000:.1f} us = {tf:.1f} TFLOP/s   (cuBLAS f16: {cms*1000:.1f} us = {2.0*M*N*K/cms/1e9:.1f} TFLOP/s)", flush=True)
    return {"M": M, "N": N, "K": K, "us": ms * 1000, "tflops": tf, "cublas_us": cms * 1000}

res["gemm_time"] = [gemm_time(*s) for s in [(16, 1152, 384), (16, 1536, 384), (16, 384, 1536), (262144, 1152, 384), (262144, 1536, 384), (262144, 384, 1536), (8192, 8192, 8192)]]

# ---- forward vs the fp32 restatement
def cos(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return (a * b).sum(-1) / np.sqrt((a * a).sum(-1) * (b * b).sum(-1))

for (name, kw, heads, B, S) in [("tiny", dict(hidden=128, layers=2, intermediate=256, vocab=1000, max_pos=64, seed=3), 4, 5, 12),
                                ("minilm-l6", dict(seed=1), 12, 8, 24), ("minilm-l6-long", dict(seed=1), 12, 2, 200)]:
    w = E.synthetic_bert_weights(**kw)
    enc = E.OnnxBiEncoder(ctx, E.write_safetensors(w), n_heads=heads)
    rng = np.random.default_rng(7)
    ids = rng.integers(0, w["embeddings.word_embeddings.weight"].shape[0], (B, S))
    lens = rng.integers(1, S + 1, B); lens[0] = S
    mask = (np.arange(S)[None, :] < lens[:, None]).astype(np.int64)
    tt = np.zeros((B, S), dtype=np.int64)
    got = enc.embed(ids, tt, mask)
    want = eo.embed(w, ids, tt, mask, n_heads=heads)
    err = np.abs(got - want).max(); c = cos(got, want)
    # the reference's bar: cosine between two embeddings within 1e-3
    cg = cos(got[:-1], got[1:]); cw = cos(want[:-1], want[1:])
    ok = bool(err < 2e-2 and np.abs(cg - cw).max() < 1e-3 and (1 - c).max() < 1e-4)
    allok &= ok
    res["forward"].append({"case": name, "B": B, "S": S, "max_abs_err": float(err), "min_cos_vs_f32": float(c.min()), "pair_cos_err": float(np.abs(cg - cw).max()), "ok": ok})
    print("forward", name, "max abs err", err, "min cos", c.min(), "pair-cos err", np.abs(cg - cw).max(), "OK" if ok else "FAIL", flush=True)
    if name == "minilm-l6":
        # latency of one query, throughput of a batch
        for (b, s, it) in [(1, 16, 50), (64, 16, 20), (4096, 16, 5), (16384, 16, 3)]:
            i2 = rng.integers(0, 30522, (b, s)); m2 = np.ones((b, s), dtype=np.int64); t2 = np.zeros((b, s), dtype=np.int64)
            di, dm, dt = [torch.from_numpy(x).to(dev) for x in (i2, m2, t2)]
            out = torch.empty(b, 384, device=dev)
            side = torch.cuda.Stream(); torch.cuda.set_stream(side)
            st = side.cuda_stream
            for _ in range(3): enc.embed_device(di.data_ptr(), dt.data_ptr(), dm.data_ptr(), b, s, out.data_ptr(), 0, st)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(it): enc.embed_device(di.data_ptr(), dt.data_ptr(), dm.data_ptr(), b, s, out.data_ptr(), 0, st)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / it
            t0 = time.perf_counter()
            for _ in range(it): enc.embed(i2, t2, m2)
            host_ms = (time.perf_counter() - t0) / it * 1000
            flops = 2.0 * b * s * 6 * (4 * 384 * 384 + 2 * 384 * 1536)
            if b in (1, 16384):
                import ctypes as C
                from metarank_b200 import _capi
                lib = _capi.lib()
                _capi.check(lib.mr_profile_begin())
                for _ in range(3): enc.embed_device(di.data_ptr(), dt.data_ptr(), dm.data_ptr(), b, s, out.data_ptr(), 0, st)
                buf = C.create_string_buffer(1 << 16); n = C.c_size_t()
                _capi.check(lib.mr_profile_end(buf, C.c_size_t(len(buf)), C.byref(n)))
                for k in json.loads(buf.value.decode()):
                    print(f"      {k['kernel']}: {k['launches']} launches, {k['ms'] / k['launches'] * 1000:.1f} us each", flush=True)
            print(f"embed batch {b} x {s} tokens: device {ms:.3f} ms ({b/ms*1000:.0f} queries/s, {flops/ms/1e9:.1f} dense TFLOP/s), host call {host_ms:.3f} ms", flush=True)
            res.setdefault("embed_time", []).append({"batch": b, "seq": s, "device_ms": ms, "host_call_ms": host_ms, "dense_tflops": flops / ms / 1e9})
    enc.close()
res["ok"] = bool(allok)
json.dump(res, open("gpurun_out/encoder_check.json", "w"), indent=1)
print("ALL OK" if allok else "FAILURES")
