"""The bi-encoder oracle (oracle/encoder_oracle.py) pinned to an independent implementation of the same graph
(transformers.BertModel, fp32) and to a literal transcription of the reference's avgpool loop
(S/ml/onnx/sbert/OnnxBiEncoder.scala:38-60); the safetensors writer the synthetic weights travel in."""
import json
import struct

import numpy as np
import pytest
import torch

from metarank_b200 import encoder as E
from oracle import encoder_oracle as eo


def _hf(w, heads, vocab, hidden, layers, inter, max_pos):
    from transformers import BertConfig, BertModel

    cfg = BertConfig(vocab_size=vocab, hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads,
                     intermediate_size=inter, max_position_embeddings=max_pos, type_vocab_size=2, layer_norm_eps=1e-12,
                     hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    m = BertModel(cfg, add_pooling_layer=False).eval()
    r = m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    assert not r.unexpected_keys and all("position_ids" in k for k in r.missing_keys), r
    return m


@pytest.mark.parametrize("heads,hidden,layers,inter", [(4, 128, 2, 256), (12, 384, 2, 1536)])
def test_oracle_matches_transformers_bert(heads, hidden, layers, inter):
    pytest.importorskip("transformers")
    w = E.synthetic_bert_weights(hidden=hidden, layers=layers, intermediate=inter, vocab=1000, max_pos=64, seed=3)
    m = _hf(w, heads, 1000, hidden, layers, inter, 64)
    rng = np.random.default_rng(0)
    B, S = 4, 13
    ids = rng.integers(0, 1000, (B, S))
    tt = rng.integers(0, 2, (B, S))
    mask = np.ones((B, S), dtype=np.int64)
    mask[1, 6:] = 0
    mask[2, 3:] = 0
    mask[3, 1:] = 0
    with torch.no_grad():
        ref = m(input_ids=torch.from_numpy(ids), token_type_ids=torch.from_numpy(tt), attention_mask=torch.from_numpy(mask))[0].numpy()
    mine = eo.last_hidden_state(w, ids, tt, mask, heads, 1e-12).numpy()
    # padded positions differ by construction in nothing: both run the same masked softmax
    assert np.abs(ref - mine).max() < 5e-6


def _avgpool_scala(tensor, token_lengths, dim):
    """while-loops of OnnxBiEncoder.avgpool, line for line"""
    result = []
    for s in range(len(tensor)):
        embed = np.zeros(dim, dtype=np.float32)
        for i in range(dim):
            total, cnt = 0.0, 0
            for j in range(len(tensor[s])):
                if j < token_lengths[s]:
                    total += float(tensor[s][j][i])
                    cnt += 1
            with np.errstate(invalid="ignore", divide="ignore"):
                embed[i] = np.float32(np.float64(total) / np.float64(cnt))
        result.append(embed)
    return np.stack(result)


def test_avgpool_matches_reference_loop():
    rng = np.random.default_rng(5)
    t = rng.standard_normal((4, 9, 6)).astype(np.float32)
    lengths = [9, 4, 1, 0]
    got = eo.avgpool(t, lengths, 6)
    want = _avgpool_scala(t, lengths, 6)
    assert np.array_equal(got[:3].view(np.uint32), want[:3].view(np.uint32))
    assert np.isnan(got[3]).all() and np.isnan(want[3]).all()  # 0.0 / 0 -> NaN in the reference too


def test_safetensors_writer_layout():
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    b = np.arange(4, dtype=np.float16)
    blob = E.write_safetensors({"x.weight": a, "y": b})
    (hl,) = struct.unpack("<Q", blob[:8])
    hdr = json.loads(blob[8:8 + hl])
    assert hdr["x.weight"] == {"dtype": "F32", "shape": [2, 3], "data_offsets": [0, 24]}
    assert hdr["y"] == {"dtype": "F16", "shape": [4], "data_offsets": [24, 32]}
    data = blob[8 + hl:]
    assert np.array_equal(np.frombuffer(data[:24], dtype=np.float32).reshape(2, 3), a)
    assert np.array_equal(np.frombuffer(data[24:32], dtype=np.float16), b)


def test_synthetic_weights_have_the_bert_names():
    w = E.synthetic_bert_weights(hidden=64, layers=1, intermediate=128, vocab=50, max_pos=16)
    assert w["encoder.layer.0.attention.self.query.weight"].shape == (64, 64)
    assert w["encoder.layer.0.intermediate.dense.weight"].shape == (128, 64)
    assert w["encoder.layer.0.output.dense.weight"].shape == (64, 128)
    assert w["embeddings.word_embeddings.weight"].shape == (50, 64)
