"""TEST INFRASTRUCTURE ONLY — CPU restatement of the bi-encoder forward, never on the product path.

What it restates: OnnxBiEncoder.embed + avgpool (reference S/ml/onnx/sbert/OnnxBiEncoder.scala:13-60).  The graph the
reference runs lives in a third-party artefact, the ONNX export of a HuggingFace BertModel
(sentence-transformers/all-MiniLM-L6-v2, fetched from the HuggingFace hub at run time by S/ml/onnx/sbert/OnnxSession.scala:57-85;
not vendored, and there is no network here), executed by onnxruntime 1.22.0.  So the published algorithm is restated —
BertModel.forward: embeddings + LayerNorm, per layer softmax(QK^T / sqrt(d) + mask) V, output dense + residual + LayerNorm,
erf-GELU feed-forward + residual + LayerNorm — in plain fp32 torch ops, and avgpool exactly as the Scala loop does it
(double sum over the first sum(attention_mask) tokens, divided by the count, narrowed to float).

Pinning: tests/test_encoder_cpu.py holds this restatement to `transformers.BertModel` (an independent implementation of the
same graph) on seeded random weights to 2e-6, and avgpool to a literal transcription of the Scala loop.  The reference's own
known answers for this path (cosine 0.539 / 0.738 +-1e-3, T/ml/onnx/sbert/OnnxBiencoderTest.scala:22-25; 0.7093 / 0.6511 /
0.2450, T/feature/FieldMatchBiencoderFeatureTest.scala:73-75) need the MiniLM weights and its tokenizer, both absent here:
against those vectors PARITY IS UNPINNED.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def _ln(x, g, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def last_hidden_state(w: dict, input_ids, token_type_ids, attention_mask, n_heads: int, eps: float) -> torch.Tensor:
    """BertModel(...)[0] in fp32: [batch x seq x hidden]."""
    W = {k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in w.items()}
    ids = torch.as_tensor(np.asarray(input_ids), dtype=torch.long)
    tt = torch.zeros_like(ids) if token_type_ids is None else torch.as_tensor(np.asarray(token_type_ids), dtype=torch.long)
    mask = torch.as_tensor(np.asarray(attention_mask), dtype=torch.long)
    B, S = ids.shape
    x = W["embeddings.word_embeddings.weight"][ids] + W["embeddings.position_embeddings.weight"][:S][None] \
        + W["embeddings.token_type_embeddings.weight"][tt]
    x = _ln(x, W["embeddings.LayerNorm.weight"], W["embeddings.LayerNorm.bias"], eps)
    H = x.shape[-1]
    D = H // n_heads
    add_mask = (1.0 - mask[:, None, None, :].float()) * torch.finfo(torch.float32).min
    layer = 0
    while f"encoder.layer.{layer}.attention.self.query.weight" in W:
        p = f"encoder.layer.{layer}."
        lin = lambda t, n: t @ W[p + n + ".weight"].T + W[p + n + ".bias"]
        split = lambda t: t.view(B, S, n_heads, D).transpose(1, 2)
        q, k, v = split(lin(x, "attention.self.query")), split(lin(x, "attention.self.key")), split(lin(x, "attention.self.value"))
        s = q @ k.transpose(-1, -2) / math.sqrt(D) + add_mask
        ctx = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, S, H)
        x = _ln(lin(ctx, "attention.output.dense") + x, W[p + "attention.output.LayerNorm.weight"], W[p + "attention.output.LayerNorm.bias"], eps)
        mid = torch.nn.functional.gelu(lin(x, "intermediate.dense"))
        x = _ln(lin(mid, "output.dense") + x, W[p + "output.LayerNorm.weight"], W[p + "output.LayerNorm.bias"], eps)
        layer += 1
    return x


def avgpool(tensor: np.ndarray, token_lengths, dim: int) -> np.ndarray:
    """OnnxBiEncoder.avgpool (:38-60): f64 sum of the first tokenLengths(s) rows / their count -> f32."""
    t = np.asarray(tensor, dtype=np.float32)
    out = np.empty((t.shape[0], dim), dtype=np.float32)
    for s in range(t.shape[0]):
        n = min(int(token_lengths[s]), t.shape[1])
        with np.errstate(invalid="ignore", divide="ignore"):
            out[s] = (t[s, :n, :dim].astype(np.float64).sum(axis=0) / np.float64(n)).astype(np.float32) if n else np.float32("nan")
    return out


def embed(w: dict, input_ids, token_type_ids, attention_mask, n_heads: int = 12, eps: float = 1e-12) -> np.ndarray:
    """OnnxBiEncoder.embed (:13-36) after tokenization."""
    with torch.no_grad():
        h = last_hidden_state(w, input_ids, token_type_ids, attention_mask, n_heads, eps).numpy()
    lengths = np.asarray(attention_mask).sum(axis=1)
    return avgpool(h, lengths, h.shape[-1])
