"""Bi-encoder query forward on the GPU against the fp32 restatement (oracle/encoder_oracle.py, itself
pinned to transformers.BertModel in tests/test_encoder_cpu.py).  Floating point, so a tolerance, and it is the
reference's own: cosines of embedding pairs within 1e-3 (T/ml/onnx/sbert/OnnxBiencoderTest.scala:22-25); on top of that
every embedding must point the way the fp32 one does (1 - cos < 1e-4) and no component may be off by more than 2e-2.
The dense layer alone is held to an f64 matmul of the same binary16 operands: f32 accumulation, so 2e-4 relative."""
import numpy as np
import pytest
import torch

from metarank_b200 import _capi, encoder as E
from oracle import encoder_oracle as eo

pytestmark = pytest.mark.gpu

PAIR_COS_TOL = 1e-3   # the reference test's +-0.001
DIR_TOL = 1e-4        # 1 - cos(gpu, fp32)
ABS_TOL = 2e-2        # any single component (embeddings are O(1) after LayerNorm)


def _cos(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    return (a * b).sum(-1) / np.sqrt((a * a).sum(-1) * (b * b).sum(-1))


@pytest.mark.parametrize("M,N,K", [(1, 384, 384), (16, 1152, 384), (100, 1536, 384), (128, 128, 64), (257, 64, 128),
                                   (300, 384, 1536), (1000, 192, 448), (4096, 1152, 384)])
def test_dense_layer_matches_f64_matmul(ctx, M, N, K):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M * 7 + N)
    a = (torch.randn(M, K, generator=g) * 0.5).half().to(dev)
    w = (torch.randn(N, K, generator=g) * 0.05).half().to(dev)
    b = torch.randn(N, generator=g).to(dev)
    o32 = torch.full((M, N), float("nan"), device=dev)
    o16 = torch.full((M, N), float("nan"), device=dev, dtype=torch.half)
    E.gemm_f16_device(ctx, a.data_ptr(), w.data_ptr(), b.data_ptr(), 0, o32.data_ptr(), o16.data_ptr(), M, N, K)
    torch.cuda.synchronize()
    ref = a.double() @ w.double().T + b.double()
    scale = max(ref.abs().max().item(), 1.0)
    assert (o32.double() - ref).abs().max().item() < 2e-4 * scale
    assert (o16.double() - ref).abs().max().item() < 2e-3 * scale


def test_dense_layer_epilogues(ctx):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    M, N, K = 200, 384, 1536
    a = (torch.randn(M, K, generator=g) * 0.5).half().to(dev)
    w = (torch.randn(N, K, generator=g) * 0.05).half().to(dev)
    b = torch.randn(N, generator=g).to(dev)
    r = torch.randn(M, N, generator=g).to(dev)
    o32 = torch.empty(M, N, device=dev)
    base = a.double() @ w.double().T
    # exact-erf GELU (BertIntermediate)
    E.gemm_f16_device(ctx, a.data_ptr(), w.data_ptr(), b.data_ptr(), 0, o32.data_ptr(), 0, M, N, K, gelu=True)
    torch.cuda.synchronize()
    ref = torch.nn.functional.gelu(base + b.double())
    assert (o32.double() - ref).abs().max().item() < 2e-4 * max(ref.abs().max().item(), 1.0)
    # bias + residual (BertSelfOutput / BertOutput before their LayerNorm)
    E.gemm_f16_device(ctx, a.data_ptr(), w.data_ptr(), b.data_ptr(), r.data_ptr(), o32.data_ptr(), 0, M, N, K)
    torch.cuda.synchronize()
    ref = base + b.double() + r.double()
    assert (o32.double() - ref).abs().max().item() < 2e-4 * max(ref.abs().max().item(), 1.0)
    with pytest.raises(_capi.MrError):  # the forward never needs both in one layer
        E.gemm_f16_device(ctx, a.data_ptr(), w.data_ptr(), b.data_ptr(), r.data_ptr(), o32.data_ptr(), 0, M, N, K, gelu=True)
    E.gemm_f16_device(ctx, a.data_ptr(), w.data_ptr(), 0, 0, o32.data_ptr(), 0, M, N, K)
    torch.cuda.synchronize()
    assert (o32.double() - base).abs().max().item() < 2e-4 * max(base.abs().max().item(), 1.0)


def test_dense_layer_rejects_untileable_shapes(ctx):
    with pytest.raises(_capi.MrError):
        E.gemm_f16_device(ctx, 1, 1, 0, 0, 0, 0, 16, 100, 64)
    with pytest.raises(_capi.MrError):
        E.gemm_f16_device(ctx, 1, 1, 0, 0, 0, 0, 16, 128, 72)


def _case(ctx, kw, heads, B, S, seed, lens=None, types=False):
    w = E.synthetic_bert_weights(**kw)
    enc = E.OnnxBiEncoder(ctx, E.write_safetensors(w), n_heads=heads)
    rng = np.random.default_rng(seed)
    vocab = w["embeddings.word_embeddings.weight"].shape[0]
    ids = rng.integers(0, vocab, (B, S))
    if lens is None:
        lens = rng.integers(1, S + 1, B)
        lens[0] = S
    mask = (np.arange(S)[None, :] < np.asarray(lens)[:, None]).astype(np.int64)
    tt = rng.integers(0, 2, (B, S)) if types else np.zeros((B, S), dtype=np.int64)
    got = enc.embed(ids, tt, mask)
    want = eo.embed(w, ids, tt, mask, n_heads=heads)
    enc.close()
    return got, want


def _check(got, want):
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.abs(got - want).max() < ABS_TOL
    assert (1 - _cos(got, want)).max() < DIR_TOL
    if len(got) > 1:
        assert np.abs(_cos(got[:-1], got[1:]) - _cos(want[:-1], want[1:])).max() < PAIR_COS_TOL


def test_minilm_shape_forward_matches_fp32(ctx):
    got, want = _case(ctx, dict(seed=1), 12, 8, 24, seed=7)
    _check(got, want)


def test_small_encoder_ragged_batch_and_token_types(ctx):
    got, want = _case(ctx, dict(hidden=128, layers=2, intermediate=256, vocab=1000, max_pos=64, seed=3), 4, 7, 33, seed=2, types=True)
    _check(got, want)


def test_large_batch_of_short_queries(ctx):
    # 1024 queries x 4 heads: the register-resident attention kernel, against the fp32 restatement
    got, want = _case(ctx, dict(hidden=128, layers=2, intermediate=256, vocab=1000, max_pos=64, seed=13), 4, 1024, 9, seed=4)
    _check(got, want)


@pytest.mark.parametrize("S", [16, 8, 3, 1])
def test_tensor_core_attention_for_short_sequences(ctx, S):
    # seq <= 16 with >= 1024 (sequence, head) pairs: attention_mma16_kernel (mma.sync tiles, P rounded to binary16)
    got, want = _case(ctx, dict(hidden=128, layers=2, intermediate=256, vocab=1000, max_pos=64, seed=15), 4, 300, S, seed=S)
    _check(got, want)


def test_head_dimension_64(ctx):
    got, want = _case(ctx, dict(hidden=256, layers=2, intermediate=512, vocab=500, max_pos=32, seed=4), 4, 3, 17, seed=5)
    _check(got, want)


def test_long_sequences_cross_a_row_tile(ctx):
    # 2 x 200 tokens: M = 400 spans four 128-row tiles, keys stride the warp more than once
    got, want = _case(ctx, dict(layers=2, seed=6), 12, 2, 200, seed=9, lens=[200, 131])
    _check(got, want)


def test_single_token_and_empty_mask(ctx):
    # one real token (just [CLS]); and attention_mask all zero: avgpool divides 0.0 by 0 -> NaN, like the reference
    got, want = _case(ctx, dict(hidden=128, layers=1, intermediate=256, vocab=100, max_pos=16, seed=8), 4, 3, 5, seed=1, lens=[5, 1, 0])
    _check(got[:2], want[:2])
    assert np.isnan(got[2]).all() and np.isnan(want[2]).all()


def test_one_query_is_independent_of_its_batch(ctx):
    # padding = true pads to the longest text of the batch; a text's embedding must not depend on its neighbours
    w = E.synthetic_bert_weights(layers=2, seed=12)
    enc = E.OnnxBiEncoder(ctx, E.write_safetensors(w), n_heads=12)
    rng = np.random.default_rng(3)
    ids = rng.integers(0, 30522, (4, 20))
    lens = np.array([20, 9, 14, 3])
    mask = (np.arange(20)[None, :] < lens[:, None]).astype(np.int64)
    tt = np.zeros_like(ids)
    full = enc.embed(ids, tt, mask)
    for b in range(4):
        alone = enc.embed(ids[b:b + 1, :lens[b]], tt[b:b + 1, :lens[b]], mask[b:b + 1, :lens[b]])
        assert np.abs(alone[0] - full[b]).max() < 2e-3
        assert 1 - _cos(alone, full[b:b + 1])[0] < 1e-5
    enc.close()


def test_short_and_general_attention_agree(ctx):
    # seq <= 32 runs the register-resident warp-per-head kernel, anything longer the shared-memory one; one extra masked
    # pad token moves a batch from the first to the second without changing the mathematics
    w = E.synthetic_bert_weights(layers=2, seed=21)
    enc = E.OnnxBiEncoder(ctx, E.write_safetensors(w), n_heads=12)
    rng = np.random.default_rng(8)
    n = 352  # x 12 heads >= 4096 (sequence, head) pairs: the batch size from which the warp-per-head kernel is chosen
    ids = rng.integers(0, 30522, (n, 32))
    lens = rng.integers(1, 33, n)
    lens[:5] = [32, 17, 1, 31, 8]
    mask = (np.arange(32)[None, :] < lens[:, None]).astype(np.int64)
    tt = np.zeros_like(ids)
    short = enc.embed(ids, tt, mask)
    pad = lambda a: np.concatenate([a, np.zeros((n, 1), dtype=np.int64)], axis=1)  # noqa: E731
    general = enc.embed(pad(ids), pad(tt), pad(mask))
    enc.close()
    assert np.abs(short - general).max() < 1e-5


def test_argument_errors(ctx):
    w = E.synthetic_bert_weights(hidden=128, layers=1, intermediate=256, vocab=100, max_pos=16, seed=8)
    blob = E.write_safetensors(w)
    enc = E.OnnxBiEncoder(ctx, blob, n_heads=4)
    ok = np.zeros((1, 4), dtype=np.int64)
    with pytest.raises(_capi.MrError):   # beyond the position table
        enc.embed(np.zeros((1, 17), dtype=np.int64), None, np.ones((1, 17), dtype=np.int64))
    with pytest.raises(_capi.MrError):   # token id outside the vocabulary: ONNX Runtime's Gather fails the run
        enc.embed(np.full((1, 4), 100, dtype=np.int64), ok, np.ones((1, 4), dtype=np.int64))
    assert np.isfinite(enc.embed(ok, ok, np.ones((1, 4), dtype=np.int64))).all()  # and the handle survives it
    enc.close()
    with pytest.raises(_capi.MrError):   # heads must divide hidden into 32 or 64
        E.OnnxBiEncoder(ctx, blob, n_heads=3)
    with pytest.raises(_capi.MrError):
        E.OnnxBiEncoder(ctx, blob[:100], n_heads=4)
    del w["encoder.layer.0.output.dense.bias"]
    with pytest.raises(_capi.MrError):
        E.OnnxBiEncoder(ctx, E.write_safetensors(w), n_heads=4)
