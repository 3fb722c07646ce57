"""Host-side mirror of the reference's bi-encoder (S/ml/onnx/sbert/OnnxBiEncoder.scala:11-36) over libmrgpu.so.

`OnnxBiEncoder.embed(batch: Array[String])` tokenizes with DJL's HuggingFaceTokenizer and runs an ONNX session; the
tokenizer stays on the caller's side here (it is a third-party Rust library in the reference too), so `embed` takes the
three int64 tensors the reference builds from the encodings and returns what `avgpool` returns: one f32 vector per text.
There is no CPU path: without the CUDA library this raises.
"""
from __future__ import annotations

import ctypes as C
import json
import struct

import numpy as np

from . import _capi
from .booster import Context


class OnnxBiEncoder:
    """mr_encoder_* of include/mr_b200.h.  `weights` = bytes of a HuggingFace BertModel `model.safetensors`."""

    def __init__(self, ctx: Context, weights: bytes, n_heads: int = 12, layer_norm_eps: float = 1e-12):
        self._ctx = ctx
        self._h = C.c_void_p()
        buf = (C.c_uint8 * len(weights)).from_buffer_copy(weights)
        _capi.check(_capi.lib().mr_encoder_load(ctx.handle, buf, C.c_size_t(len(weights)), C.c_int32(n_heads),
                                                C.c_double(layer_norm_eps), C.byref(self._h)))
        d, l, t, v = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        _capi.check(_capi.lib().mr_encoder_info(self._h, C.byref(d), C.byref(l), C.byref(t), C.byref(v)))
        self.dim, self.layers, self.max_tokens, self.vocab = d.value, l.value, t.value, v.value

    def embed(self, input_ids, token_type_ids, attention_mask) -> np.ndarray:
        """[batch x seq] int64 each (padded to the longest text, as `padding = true` does) -> [batch x dim] f32."""
        ids = np.ascontiguousarray(input_ids, dtype=np.int64)
        mask = np.ascontiguousarray(attention_mask, dtype=np.int64)
        if ids.ndim != 2 or mask.shape != ids.shape:
            raise ValueError("input_ids and attention_mask must be [batch x seq] of the same shape")
        types = None if token_type_ids is None else np.ascontiguousarray(token_type_ids, dtype=np.int64)
        if types is not None and types.shape != ids.shape:
            raise ValueError("token_type_ids must have the shape of input_ids")
        out = np.empty((ids.shape[0], self.dim), dtype=np.float32)
        p64 = C.POINTER(C.c_int64)
        _capi.check(_capi.lib().mr_encoder_embed(
            self._h, ids.ctypes.data_as(p64), types.ctypes.data_as(p64) if types is not None else None,
            mask.ctypes.data_as(p64), C.c_int32(ids.shape[0]), C.c_int32(ids.shape[1]),
            out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def embed_device(self, d_ids: int, d_types: int, d_mask: int, batch: int, seq: int, d_out: int, d_out_f64: int = 0,
                     stream: int = 0) -> None:
        _capi.check(_capi.lib().mr_encoder_embed_device(
            self._h, C.c_void_p(d_ids), C.c_void_p(d_types), C.c_void_p(d_mask), C.c_int32(batch), C.c_int32(seq),
            C.c_void_p(d_out), C.c_void_p(d_out_f64), C.c_void_p(stream)))

    def close(self) -> None:
        if self._h:
            _capi.check(_capi.lib().mr_encoder_free(self._h))
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gemm_f16_device(ctx: Context, d_a: int, d_w: int, d_bias: int, d_residual: int, d_out_f32: int, d_out_f16: int,
                    m: int, n: int, k: int, gelu: bool = False, stream: int = 0) -> None:
    """mr_encoder_gemm_f16: the forward's dense layer on its own (device pointers)."""
    _capi.check(_capi.lib().mr_encoder_gemm_f16(
        ctx.handle, C.c_void_p(d_a), C.c_void_p(d_w), C.c_void_p(d_bias), C.c_void_p(d_residual), C.c_void_p(d_out_f32),
        C.c_void_p(d_out_f16), C.c_int32(m), C.c_int32(n), C.c_int32(k), C.c_int32(1 if gelu else 0), C.c_void_p(stream)))


def write_safetensors(tensors: dict[str, np.ndarray]) -> bytes:
    """Minimal safetensors writer (F32 / F16 / I64) for synthetic weights: tests and bench.py have no model download."""
    header, blobs, off = {}, [], 0
    for name, a in tensors.items():
        a = np.ascontiguousarray(a)
        dt = {np.dtype(np.float32): "F32", np.dtype(np.float16): "F16", np.dtype(np.int64): "I64"}[a.dtype]
        raw = a.tobytes()
        header[name] = {"dtype": dt, "shape": list(a.shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    return struct.pack("<Q", len(hj)) + hj + b"".join(blobs)


def synthetic_bert_weights(hidden=384, layers=6, intermediate=1536, vocab=30522, max_pos=512, n_types=2, seed=0,
                           std=0.04) -> dict[str, np.ndarray]:
    """Random-init weights with the tensor names of a HuggingFace BertModel (all-MiniLM-L6-v2's shape by default)."""
    rng = np.random.default_rng(seed)
    n = lambda *s: (rng.standard_normal(s) * std).astype(np.float32)
    t = {"embeddings.word_embeddings.weight": n(vocab, hidden), "embeddings.position_embeddings.weight": n(max_pos, hidden),
         "embeddings.token_type_embeddings.weight": n(n_types, hidden),
         "embeddings.LayerNorm.weight": (1 + n(hidden)).astype(np.float32), "embeddings.LayerNorm.bias": n(hidden)}
    for l in range(layers):
        p = f"encoder.layer.{l}."
        for q in ("query", "key", "value"):
            t[p + f"attention.self.{q}.weight"] = n(hidden, hidden)
            t[p + f"attention.self.{q}.bias"] = n(hidden)
        t[p + "attention.output.dense.weight"] = n(hidden, hidden)
        t[p + "attention.output.dense.bias"] = n(hidden)
        t[p + "attention.output.LayerNorm.weight"] = (1 + n(hidden)).astype(np.float32)
        t[p + "attention.output.LayerNorm.bias"] = n(hidden)
        t[p + "intermediate.dense.weight"] = n(intermediate, hidden)
        t[p + "intermediate.dense.bias"] = n(intermediate)
        t[p + "output.dense.weight"] = n(hidden, intermediate)
        t[p + "output.dense.bias"] = n(hidden)
        t[p + "output.LayerNorm.weight"] = (1 + n(hidden)).astype(np.float32)
        t[p + "output.LayerNorm.bias"] = n(hidden)
    return t
