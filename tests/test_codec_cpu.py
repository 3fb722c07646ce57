"""the reference's binary store format.

* the oracle's restatement (oracle/codec_oracle.py) is PINNED on the reference's golden files for the
  codec family: T/resources/codec/ctv-v{1,2,3}.bin, asserted byte for byte by
  T/fstore/codec/impl/TrainValuesCodecTest.scala:57-86.  The reference checkout holds their git-lfs
  pointers (sha256 + size), reproduced here from the test's own `ctv` value (:17-55);
* the native decoder (metarank_b200/csrc/fv_codec.cpp, host only) must emit, for every FeatureValue, exactly the
  mr_state_upsert record the Python packer builds from the oracle-decoded value.
"""
import datetime
import hashlib
import math
import struct

import numpy as np
import pytest

from oracle import codec_oracle as co

# git-lfs pointers of /root/reference/src/test/resources/codec/ctv-v{1,2,3}.bin: (sha256, size)
GOLDEN = {1: ("1f53dddebb11b830db2f3388b4bdb2d6617066d8b2b038c07c05579ebf66b8af", 207),
          2: ("5a8e2efb805b724eb1f7e5e19a55d44f4df0abef54c570dd2673d99cf4a5bd50", 208),
          3: ("1cd75fda1f292507c2f805e1d8d90782d808c68e89729244bb7f2f6d983020aa", 210)}


def _reference_ctv():
    """TrainValuesCodecTest.scala:17-55"""
    ts = int(datetime.datetime(2022, 11, 17, 15, 32, 0, tzinfo=datetime.timezone.utc).timestamp() * 1000)  # Timestamp.date
    mv = [("single", "f1", 1.0), ("vector", "f2", [1.0], 1), ("category", "f3", "x", 0)]
    return dict(ct=dict(id="e1", ts=ts, user="alice", session="wow", items=["p1", "p2", "p3", "p4"],
                        interactions=[("p2", "click", None), ("p2", "purchase", None)], ranking_fields=[("foo", "bar")]),
                values=[("p1", mv), ("p2", mv), ("p3", mv)])


@pytest.mark.parametrize("version", [1, 2, 3])
def test_oracle_encoder_reproduces_the_reference_golden_files(version):
    b = co.encode_train_values(_reference_ctv(), version)
    sha, size = GOLDEN[version]
    assert len(b) == size and hashlib.sha256(b).hexdigest() == sha


def test_varnum_matches_the_published_layout():
    """VarNum.java:12-83 (Bazel's VarInt): 7 bits per byte, low group first."""
    for v, enc in [(0, "00"), (1, "01"), (127, "7f"), (128, "8001"), (300, "ac02"), (16384, "808001")]:
        o = co.Out(); o.varint(v)
        assert o.bytes().hex() == enc
        o = co.Out(); o.varlong(v)
        assert o.bytes().hex() == enc
    for v in [0, 1, 127, 128, 2 ** 31 - 1, -1, -2 ** 31]:
        o = co.Out(); o.varint(v)
        assert co.In(o.bytes()).varint() == v
    for v in [0, 1, 2 ** 35, 2 ** 63 - 1, -1, -2 ** 63, 1668699120000]:
        o = co.Out(); o.varlong(v)
        assert co.In(o.bytes()).varlong() == v
    o = co.Out(); o.varlong(-1)
    assert len(o.bytes()) == 10  # a negative long takes the full ten groups


def test_write_utf_is_modified_utf8():
    o = co.Out(); o.utf("a\0é€😀")
    assert o.bytes().hex() == "000e" + "61" + "c080" + "c3a9" + "e282ac" + "eda0bd" + "edb880"
    assert co.In(o.bytes()).utf() == "a\0é€😀"


def _sample_values():
    rng = np.random.Generator(np.random.PCG64(5))
    scopes = [("item", "p1"), ("item", "товар-2"), ("user", "alice"), ("session", "s😀"), ("global",),
              ("field", "genre", "comedy"), ("irf", "query", "jeans", "p7"), ("ranking", "r-9")]
    vals = []
    k = 0

    def key():
        nonlocal k
        k += 1
        return scopes[k % len(scopes)], f"feat_{k}"

    for v in [1.5, -0.0, float("nan"), float("inf"), 1e-310, True, False, "blue", "", "naïve\0x",
              ["a", "b", "c"], ("strings", []), [1.0, 2.5, float("nan")], [float(x) for x in rng.normal(size=40)]]:
        vals.append(dict(type="scalar", key=key(), ts=1668699120000 + k, value=v, expire_ms=86400000))
    for n in [0, 1, 127, 128, 2 ** 40, -5]:
        vals.append(dict(type="counter", key=key(), ts=k, value=n, expire_ms=1))
    vals.append(dict(type="pcounter", key=key(), ts=5, expire_ms=7,
                     values=[dict(start=10 * j, end=10 * j + 9, periods=p, value=int(rng.integers(0, 1 << 40)))
                             for j, p in enumerate([1, 7, 30])]))
    vals.append(dict(type="pcounter", key=key(), ts=5, values=[], expire_ms=7))
    vals.append(dict(type="blist", key=key(), ts=9, expire_ms=3, values=[(100 - j, f"p{j}") for j in range(12)]))
    vals.append(dict(type="blist", key=key(), ts=9, expire_ms=3, values=[]))
    vals.append(dict(type="blist", key=key(), ts=9, expire_ms=3, values=[(1, "p1"), (2, 3.0), (3, "p7")]))  # mixed: the ids survive
    vals.append(dict(type="numstats", key=key(), ts=1, min=0.5, max=9.0, quantiles={50: 3.0, 90: 8.0}, expire_ms=2))
    vals.append(dict(type="map", key=key(), ts=1, values={"a": 1.0, "b": "x", "c": ["y"]}, expire_ms=2))
    vals.append(dict(type="frequency", key=key(), ts=1, values={"x": 0.25, "y": 0.75}, expire_ms=2))
    return vals


def _same(a, b):
    if isinstance(a, float) and isinstance(b, float):
        return (math.isnan(a) and math.isnan(b)) or struct.pack(">d", a) == struct.pack(">d", b)
    if isinstance(a, (list, tuple)) and isinstance(b, (list, tuple)):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    if isinstance(a, dict) and isinstance(b, dict):
        return a.keys() == b.keys() and all(_same(a[k], b[k]) for k in a)
    return type(a) == type(b) and a == b


@pytest.mark.parametrize("legacy", [False, True])
def test_feature_value_codec_roundtrips(legacy):
    """The reference's own codec tests are round trips (T/fstore/codec/values/VCodecTest.scala:10-52):
    encode -> decode, delimited, and a stream read until EOF."""
    vals = _sample_values()
    blob = co.encode_delimited(vals, legacy)
    back = co.decode_delimited(blob)
    assert len(back) == len(vals)
    for a, b in zip(vals, back):
        want = dict(a)
        if want["type"] == "scalar" and isinstance(want["value"], tuple):
            want["value"] = list(want["value"][1])
        if legacy:
            want["expire_ms"] = co.DAYS_90_MS
        want["key"] = (tuple(want["key"][0]), want["key"][1])
        if want["type"] == "blist":
            want["values"] = [tuple(x) for x in want["values"]]
        assert _same(want, b), (a, b)
    assert co.decode_delimited(b"") == []                       # "handle eof"
    assert len(co.decode_delimited(blob[:-3])) == len(vals) - 1   # truncated tail ends the stream


@pytest.mark.parametrize("legacy", [False, True])
def test_native_decoder_emits_the_packers_upsert_records(legacy):
    from metarank_b200 import features as F

    vals = _sample_values()
    blob = co.encode_delimited(vals, legacy)
    recs, n, unsupported, consumed = F.transcode_feature_values(blob)
    state = co.to_state(co.decode_delimited(blob))
    assert n == len(vals) and consumed == len(blob)
    assert unsupported == len(vals) - len(state) == 3            # numstats, map, frequency (a mixed list keeps its ids)
    assert recs == F.pack_feature_values(state)
    # a truncated trailing record ends the stream (BinaryVCodec.decodeDelimited -> Right(None))
    recs2, n2, _, consumed2 = F.transcode_feature_values(blob[:-2])
    assert n2 == len(vals) - 1 and consumed2 < len(blob) - 2 and recs.startswith(recs2)
    assert F.transcode_feature_values(b"") == (b"", 0, 0, 0)


def test_native_decoder_rejects_malformed_records():
    from metarank_b200 import _capi, features as F

    good = co.encode_delimited([dict(type="scalar", key=(("item", "p1"), "price"), ts=1, value=1.0, expire_ms=1)])
    body = bytearray(good[4:])
    for mutate, what in [(lambda b: b.__setitem__(0, 99), "fv index"), (lambda b: b.__setitem__(1, 42), "scope")]:
        bad = bytearray(body)
        mutate(bad)
        with pytest.raises(_capi.MrError) as e:
            F.transcode_feature_values(struct.pack(">i", len(bad)) + bytes(bad))
        assert what in str(e.value)
    short = body[:-4]  # frame length is honest but the record ends inside its fields
    with pytest.raises(_capi.MrError) as e:
        F.transcode_feature_values(struct.pack(">i", len(short)) + bytes(short))
    assert "ends inside a field" in str(e.value)
    with pytest.raises(_capi.MrError):
        F.transcode_feature_values(struct.pack(">i", -5) + b"xxxx")


def test_native_decoder_survives_corrupted_streams():
    """Every byte flip / truncation of a valid stream must end in a clean MR_ERR_PARSE or a shorter decode,
    never in a crash, a hang or an out-of-bounds read (lengths are attacker-controlled varints)."""
    from metarank_b200 import _capi, features as F

    rng = np.random.Generator(np.random.PCG64(9))
    base = co.encode_delimited(_sample_values())
    outcomes = {"ok": 0, "err": 0}
    for it in range(1500):
        b = bytearray(base)
        for _ in range(int(rng.integers(1, 4))):
            k = int(rng.integers(0, len(b)))
            b[k] = int(rng.integers(0, 256))
        if rng.random() < 0.3:
            b = b[:int(rng.integers(0, len(b)))]
        try:
            recs, n, uns, consumed = F.transcode_feature_values(bytes(b))
            assert consumed <= len(b) and n >= uns >= 0
            outcomes["ok"] += 1
        except _capi.MrError:
            outcomes["err"] += 1
    assert outcomes["ok"] > 0 and outcomes["err"] > 0
    # a huge list length must not allocate or loop for long: the record ends first
    o = co.Out(); o.byte(7); co.write_key(o, (("item", "p"), "f")); o.varlong(1); o.byte(4); o.varint(2 ** 31 - 1)
    with pytest.raises(_capi.MrError):
        F.transcode_feature_values(struct.pack(">i", len(o.b)) + o.bytes())
