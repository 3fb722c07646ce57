"""CPU restatement of the reference's BINARY store format (test infrastructure only — never imported by
the product path).

What it follows (S = /root/reference/src/main/scala/ai/metarank, J = .../src/main/java/ai/metarank):
  * J/util/VarNum.java:12-84                       varint / varlong (7 bits per byte, low group first)
  * java.io.DataOutput                             writeByte/Boolean/Int/Long/Double big-endian, writeUTF =
                                                   u16 length + MODIFIED UTF-8 (NUL -> C0 80, astral chars as
                                                   two 3-byte surrogates)
  * S/fstore/codec/impl/ScalarCodec.scala:11-49    Scalar: tag 0 SString, 1 SDouble, 2 SBoolean, 3 SStringList,
                                                   4 SDoubleList
  * S/fstore/codec/impl/TimeValueCodec.scala:9-19  TimeValue = varlong ts + Scalar
  * S/fstore/codec/impl/FeatureValueCodec.scala:41-237   FeatureValue (tags 7-13 current, 0-6 legacy without
                                                   the expire field), PeriodicValue, binary Key and Scope
  * S/fstore/codec/impl/{Map,Array,List,Option}Codec.scala   varint count + elements / boolean + element
  * S/fstore/codec/values/BinaryVCodec.scala:45-62 delimited framing: writeInt(len) + bytes
  * S/fstore/codec/impl/TrainValuesCodec.scala:9-98, ClickthroughCodec.scala, ItemValueCodec.scala,
    MValueCodec.scala, FieldCodec.scala, TypedIntCodec.scala, DoubleArrayCodec.scala
                                                   — only to PIN the primitives: the reference keeps golden
    files for this codec (T/resources/codec/ctv-v{1,2,3}.bin, asserted byte for byte in
    T/fstore/codec/impl/TrainValuesCodecTest.scala:57-86).  The checkout holds their git-lfs pointers
    (sha256 + size), which tests/test_codec_cpu.py reproduces from `encode_train_values`.

Values are plain Python: a FeatureValue is a dict {"type", "key": (scope_tuple, feature_name), "ts", ...,
"expire_ms"}; scopes are the tuples used everywhere else in oracle/ ("item", id), ("user", id), ("global",),
("session", id), ("field", name, value), ("irf", name, value, item), ("ranking", id).
"""
from __future__ import annotations

import struct

DAYS_90_MS = 90 * 24 * 3600 * 1000  # the legacy tags' implicit expire (FeatureValueCodec.scala:44 "90.days")


# ---------------------------------------------------------------------------------------------- DataOutput
class Out:
    def __init__(self):
        self.b = bytearray()

    def byte(self, v):
        self.b.append(v & 0xFF)

    def boolean(self, v):
        self.b.append(1 if v else 0)

    def int32(self, v):
        self.b += struct.pack(">i", v)

    def int64(self, v):
        self.b += struct.pack(">q", v)

    def double(self, v):
        self.b += struct.pack(">d", v)

    def utf(self, s: str):
        """DataOutput.writeUTF: modified UTF-8, length in BYTES as u16 (UTFDataFormatException beyond 65535)."""
        enc = bytearray()
        for unit in _utf16_units(s):
            if 0x0001 <= unit <= 0x007F:
                enc.append(unit)
            elif unit <= 0x07FF:  # includes NUL
                enc += bytes((0xC0 | (unit >> 6), 0x80 | (unit & 0x3F)))
            else:
                enc += bytes((0xE0 | (unit >> 12), 0x80 | ((unit >> 6) & 0x3F), 0x80 | (unit & 0x3F)))
        if len(enc) > 65535:
            raise ValueError("encoded string too long for writeUTF")
        self.b += struct.pack(">H", len(enc)) + enc

    def varint(self, v: int):
        """VarNum.putVarInt (VarNum.java:26-36): the int is treated as unsigned 32-bit."""
        v &= 0xFFFFFFFF
        while True:
            bits = v & 0x7F
            v >>= 7
            if v == 0:
                self.b.append(bits)
                return
            self.b.append(bits | 0x80)

    def varlong(self, v: int):
        """VarNum.putVarLong (VarNum.java:12-24).  NOTE the reference's loop condition is `highBits > 0L`
        on a SIGNED long after `>>>= 7`: a negative input still terminates (the shift clears the sign), but
        only because the first shift makes it positive; restated literally."""
        high = v & 0xFFFFFFFFFFFFFFFF
        while True:
            low = high & 0x7F
            high >>= 7  # >>> on the 64-bit pattern
            if high > 0:
                self.b.append(low | 0x80)
            else:
                self.b.append(low)
                break

    def bytes(self) -> bytes:
        return bytes(self.b)


def _utf16_units(s: str):
    for ch in s:
        cp = ord(ch)
        if cp >= 0x10000:
            cp -= 0x10000
            yield 0xD800 | (cp >> 10)
            yield 0xDC00 | (cp & 0x3FF)
        else:
            yield cp


class In:
    def __init__(self, b: bytes, pos: int = 0):
        self.b, self.p = b, pos

    def _take(self, n):
        if self.p + n > len(self.b):
            raise EOFError("truncated")
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def byte(self):  # signed, like DataInput.readByte
        return struct.unpack(">b", self._take(1))[0]

    def boolean(self):
        return self._take(1)[0] != 0

    def int32(self):
        return struct.unpack(">i", self._take(4))[0]

    def int64(self):
        return struct.unpack(">q", self._take(8))[0]

    def double(self):
        return struct.unpack(">d", self._take(8))[0]

    def utf(self) -> str:
        n = struct.unpack(">H", self._take(2))[0]
        raw = self._take(n)
        units, i = [], 0
        while i < n:
            c = raw[i]
            if c < 0x80:
                units.append(c); i += 1
            elif (c >> 5) == 0b110:
                units.append(((c & 0x1F) << 6) | (raw[i + 1] & 0x3F)); i += 2
            elif (c >> 4) == 0b1110:
                units.append(((c & 0x0F) << 12) | ((raw[i + 1] & 0x3F) << 6) | (raw[i + 2] & 0x3F)); i += 3
            else:
                raise ValueError("malformed modified UTF-8")
        return b"".join(struct.pack("<H", u) for u in units).decode("utf-16-le", errors="surrogatepass")

    def varint(self) -> int:
        """VarNum.getVarInt (VarNum.java:39-67) — at most 5 groups contribute, extra continuation bytes are
        skipped; result wraps to a signed 32-bit int."""
        result, shift, n = 0, 0, 0
        while True:
            t = self._take(1)[0]
            n += 1
            if n <= 5:
                result |= (t & 0x7F) << shift
                shift += 7
            if t < 0x80:
                break
        result &= 0xFFFFFFFF
        return result - (1 << 32) if result & 0x80000000 else result

    def varlong(self) -> int:
        """VarNum.getVarLong (VarNum.java:69-83)."""
        idx, value = 0, 0
        while True:
            t = self._take(1)[0]
            value |= (t & 0x7F) << (idx * 7)
            idx += 1
            if t < 0x80:
                break
        value &= 0xFFFFFFFFFFFFFFFF
        return value - (1 << 64) if value & (1 << 63) else value


# ---------------------------------------------------------------------------------------------- Scalar
def write_scalar(o: Out, v):
    """ScalarCodec.write (ScalarCodec.scala:11-27).  Python bool -> SBoolean, number -> SDouble, str -> SString,
    list of str -> SStringList, list of numbers -> SDoubleList (an EMPTY list must be tagged by the caller:
    ("strings", []) / ("doubles", []))."""
    if isinstance(v, tuple) and v and v[0] in ("strings", "doubles"):
        kind, items = v
        o.byte(3 if kind == "strings" else 4)
        o.varint(len(items))
        for x in items:
            o.utf(x) if kind == "strings" else o.double(float(x))
    elif isinstance(v, bool):
        o.byte(2); o.boolean(v)
    elif isinstance(v, (int, float)):
        o.byte(1); o.double(float(v))
    elif isinstance(v, str):
        o.byte(0); o.utf(v)
    elif len(v) and isinstance(v[0], str):
        o.byte(3); o.varint(len(v))
        for x in v:
            o.utf(x)
    else:
        o.byte(4); o.varint(len(v))
        for x in v:
            o.double(float(x))


def read_scalar(i: In):
    """ScalarCodec.read (ScalarCodec.scala:29-47)."""
    tag = i.byte()
    if tag == 0:
        return i.utf()
    if tag == 1:
        return i.double()
    if tag == 2:
        return i.boolean()
    if tag == 3:
        return [i.utf() for _ in range(i.varint())]
    if tag == 4:
        return [i.double() for _ in range(i.varint())]
    raise ValueError(f"cannot decode scalar {tag}")


# ---------------------------------------------------------------------------------------------- Key / Scope
_SCOPE_TAGS = {"user": 0, "item": 1, "global": 2, "session": 3, "field": 4, "irf": 5, "ranking": 6}


def write_key(o: Out, key):
    """FeatureValueCodec.KeyCodec / ScopeCodec (FeatureValueCodec.scala:186-236): scope tag + its strings, then
    the feature name."""
    scope, name = key
    o.byte(_SCOPE_TAGS[scope[0]])
    for part in scope[1:]:
        o.utf(part)
    o.utf(name)


def read_key(i: In):
    tag = i.byte()
    kinds = {v: k for k, v in _SCOPE_TAGS.items()}
    if tag not in kinds:
        raise ValueError(f"cannot parse scope with index {tag}")
    n_parts = {"user": 1, "item": 1, "global": 0, "session": 1, "field": 2, "irf": 3, "ranking": 1}[kinds[tag]]
    scope = (kinds[tag],) + tuple(i.utf() for _ in range(n_parts))
    return scope, i.utf()


# ---------------------------------------------------------------------------------------------- FeatureValue
_TAG = {"scalar": 7, "counter": 8, "numstats": 9, "map": 10, "pcounter": 11, "frequency": 12, "blist": 13}


def write_feature_value(o: Out, fv: dict, legacy: bool = False):
    """FeatureValueCodec.write (FeatureValueCodec.scala:117-170).  legacy=True writes the pre-expire tags 0-6
    that FeatureValueCodec.read still accepts (:42-113, "compat")."""
    t = fv["type"]
    o.byte(_TAG[t] - 7 if legacy else _TAG[t])
    write_key(o, fv["key"])
    o.varlong(fv["ts"])
    if t == "scalar":
        write_scalar(o, fv["value"])
    elif t == "counter":
        o.varlong(fv["value"])
    elif t == "numstats":
        o.double(fv["min"]); o.double(fv["max"])
        o.varint(len(fv["quantiles"]))
        for k, v in fv["quantiles"].items():
            o.varint(k); o.double(v)
    elif t == "map":
        o.varint(len(fv["values"]))
        for k, v in fv["values"].items():
            o.utf(k); write_scalar(o, v)
    elif t == "pcounter":
        o.varint(len(fv["values"]))
        for pv in fv["values"]:  # PeriodicValueCodec :172-184
            o.varlong(pv["start"]); o.varlong(pv["end"]); o.varint(pv["periods"]); o.varlong(pv["value"])
    elif t == "frequency":
        o.varint(len(fv["values"]))
        for k, v in fv["values"].items():
            o.utf(k); o.double(v)
    elif t == "blist":
        o.varint(len(fv["values"]))
        for ts, v in fv["values"]:  # TimeValueCodec
            o.varlong(ts); write_scalar(o, v)
    else:
        raise ValueError(t)
    if not legacy:
        o.varlong(fv.get("expire_ms", DAYS_90_MS))


def read_feature_value(i: In) -> dict:
    """FeatureValueCodec.read (FeatureValueCodec.scala:41-115)."""
    tag = i.byte()
    if not 0 <= tag <= 13:
        raise ValueError(f"cannot decode fv index {tag}")
    legacy = tag < 7
    t = {0: "scalar", 1: "counter", 2: "numstats", 3: "map", 4: "pcounter", 5: "frequency", 6: "blist"}[tag % 7]
    fv = {"type": t, "key": read_key(i), "ts": i.varlong()}
    if t == "scalar":
        fv["value"] = read_scalar(i)
    elif t == "counter":
        fv["value"] = i.varlong()
    elif t == "numstats":
        fv["min"], fv["max"] = i.double(), i.double()
        fv["quantiles"] = {i.varint(): i.double() for _ in range(i.varint())}
    elif t == "map":
        n = i.varint()
        fv["values"] = {}
        for _ in range(n):
            k = i.utf()
            fv["values"][k] = read_scalar(i)
    elif t == "pcounter":
        fv["values"] = [dict(start=i.varlong(), end=i.varlong(), periods=i.varint(), value=i.varlong())
                        for _ in range(i.varint())]
    elif t == "frequency":
        n = i.varint()
        fv["values"] = {}
        for _ in range(n):
            k = i.utf()
            fv["values"][k] = i.double()
    elif t == "blist":
        n = i.varint()
        fv["values"] = []
        for _ in range(n):
            ts = i.varlong()
            fv["values"].append((ts, read_scalar(i)))
    fv["expire_ms"] = DAYS_90_MS if legacy else i.varlong()
    return fv


def encode_delimited(values, legacy: bool = False) -> bytes:
    """BinaryVCodec(compress = false, FeatureValueCodec).encodeDelimited per value (BinaryVCodec.scala:45-50):
    the stream a FileKVStore / export of Persistence.values consists of."""
    out = bytearray()
    for fv in values:
        o = Out()
        write_feature_value(o, fv, legacy)
        out += struct.pack(">i", len(o.b)) + o.b
    return bytes(out)


def decode_delimited(blob: bytes):
    """decodeDelimited until EOF (BinaryVCodec.scala:52-62, VCodecTest "read stream with eof"): a truncated
    tail ends the stream silently, like the reference's Right(None)."""
    out, p = [], 0
    while p + 4 <= len(blob):
        n = struct.unpack(">i", blob[p:p + 4])[0]
        if n < 0 or p + 4 + n > len(blob):
            break
        out.append(read_feature_value(In(blob[p + 4:p + 4 + n])))
        p += 4 + n
    return out


def to_state(values) -> dict:
    """Decoded FeatureValues -> the {(scope, name): (kind, value)} map the oracle's extractors and
    metarank_b200.features.pack_feature_values take.  NumStats / Map / Frequency values have no reader among
    the supported extractors and are dropped (the native loader counts them as skipped)."""
    st = {}
    for fv in values:
        t = fv["type"]
        if t == "scalar":
            st[fv["key"]] = ("scalar", fv["value"])
        elif t == "counter":
            st[fv["key"]] = ("counter", fv["value"])
        elif t == "pcounter":
            st[fv["key"]] = ("pcounter", [pv["value"] for pv in fv["values"]])
        elif t == "blist":
            # the reader keeps the list and collects its SString entries (InteractedWithFeature.scala:117:
            # values.map(_.value).collect { case SString(id) => id }); other scalar kinds are skipped
            st[fv["key"]] = ("blist", [v for _, v in fv["values"] if isinstance(v, str)])
    return st


# ---------------------------------------------------------------------------------------------- TrainValues (pin only)
def _write_field(o: Out, f):
    """FieldCodec.write (FieldCodec.scala:18-41): (name, value) with value str | bool | float | [str] | [float]."""
    name, v = f
    if isinstance(v, str):
        o.byte(0); o.utf(name); o.utf(v)
    elif isinstance(v, bool):
        o.byte(1); o.utf(name); o.boolean(v)
    elif isinstance(v, (int, float)):
        o.byte(2); o.utf(name); o.double(float(v))
    elif len(v) and isinstance(v[0], str):
        o.byte(3); o.utf(name); o.int32(len(v))
        for x in v:
            o.utf(x)
    else:
        o.byte(4); o.utf(name); o.int32(len(v))
        for x in v:
            o.double(float(x))


def _write_mvalue(o: Out, m):
    """MValueCodec.write (MValueCodec.scala:19-34)."""
    if m[0] == "single":
        o.byte(0); o.utf(m[1]); o.double(m[2])
    elif m[0] == "vector":
        o.byte(1); o.utf(m[1]); o.varint(len(m[2]))
        for x in m[2]:
            o.double(x)
        o.varint(m[3])
    else:
        o.byte(2); o.utf(m[1]); o.utf(m[2]); o.varint(m[3])


def encode_train_values(ctv: dict, version: int = 3) -> bytes:
    """TrainValuesCodec.write for a ClickthroughValues (TrainValuesCodec.scala:33-38, 79-95; ClickthroughCodec
    .scala:27-39; TypedIntCodec.scala:18-28).  version 1/2 reproduce the older layouts the reference still reads
    (:16-29): v1 has no sub-type byte, v1/v2 interactions carry no relevancy option."""
    o = Out()
    o.byte(version)
    if version >= 2:
        o.byte(0)  # sub-type 0 = ClickthroughValues
    ct = ctv["ct"]
    o.utf(ct["id"])
    o.varlong(ct["ts"])
    if ct.get("user") is not None:
        o.boolean(True); o.utf(ct["user"])
    else:
        o.boolean(False)
    if ct.get("session") is not None:  # OptionCodec
        o.boolean(True); o.utf(ct["session"])
    else:
        o.boolean(False)
    o.varint(len(ct["items"]))
    for it in ct["items"]:
        o.utf(it)
    o.varint(len(ct["interactions"]))
    for item, tpe, rel in ct["interactions"]:
        o.utf(item); o.utf(tpe)
        if version == 3:
            if rel is not None:
                o.boolean(True); o.int32(rel)
            else:
                o.boolean(False)
    o.varint(len(ct["ranking_fields"]))
    for f in ct["ranking_fields"]:
        _write_field(o, f)
    o.varint(len(ctv["values"]))
    for item, mvalues in ctv["values"]:
        o.utf(item)
        o.varint(len(mvalues))
        for m in mvalues:
            _write_mvalue(o, m)
    return o.bytes()
