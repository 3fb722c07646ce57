// Decoder for the reference's binary FeatureValue stream.  What each piece follows
// (S = src/main/scala/ai/metarank, J = src/main/java/ai/metarank of the reference):
//   framing        BinaryVCodec.encodeDelimited / decodeDelimited, S/fstore/codec/values/BinaryVCodec.scala:45-62
//                  (big-endian i32 length + bytes; a truncated tail ends the stream like Right(None))
//   FeatureValue   S/fstore/codec/impl/FeatureValueCodec.scala:41-115 (tags 7-13, legacy 0-6 without expire)
//   Key / Scope    same file :186-236 (binary form: scope tag, its strings, feature name)
//   Scalar         S/fstore/codec/impl/ScalarCodec.scala:29-47
//   TimeValue      S/fstore/codec/impl/TimeValueCodec.scala:9-13
//   varint/varlong J/util/VarNum.java:39-83
//   strings        java.io.DataInput.readUTF (u16 byte length + modified UTF-8)
// Strings leave this file as mr_hash64 of their STANDARD UTF-8 bytes, the same hash the request side uses.
#include "fv_codec.h"

#include <cstring>
#include <string>

#include "common.h"
#include "schema.h"

namespace mr {
namespace {

struct In {
  const uint8_t *p, *e;
  void need(size_t n) const {
    if ((size_t)(e - p) < n) fail(MR_ERR_PARSE, "feature value record ends inside a field");
  }
  uint8_t u8() { need(1); return *p++; }
  int8_t i8() { return (int8_t)u8(); }
  uint16_t be16() { need(2); uint16_t v = (uint16_t)((p[0] << 8) | p[1]); p += 2; return v; }
  uint64_t be64() {
    need(8);
    uint64_t v = 0;
    for (int k = 0; k < 8; k++) v = (v << 8) | p[k];
    p += 8;
    return v;
  }
  double f64() { const uint64_t b = be64(); double d; memcpy(&d, &b, 8); return d; }
  // VarNum.getVarLong: 7 bits per byte, low group first, continue while the top bit is set
  int64_t varlong() {
    uint64_t v = 0;
    int idx = 0;
    uint8_t b;
    do {
      b = u8();
      if (idx < 10) v |= (uint64_t)(b & 0x7F) << (idx * 7);  // Java's shift count wraps at 64; keep the low 64 bits
      idx++;
    } while (b & 0x80);
    return (int64_t)v;
  }
  // VarNum.getVarInt: at most five groups contribute, further continuation bytes are skipped
  int32_t varint() {
    uint32_t v = 0;
    int n = 0;
    uint8_t b;
    do {
      b = u8();
      if (n < 5) v |= (uint32_t)(b & 0x7F) << (7 * n);
      n++;
    } while (b & 0x80);
    return (int32_t)v;
  }
  // readUTF -> standard UTF-8 (NUL from C0 80, surrogate pairs joined into one 4-byte sequence)
  void utf(std::string &out) {
    const size_t n = be16();
    need(n);
    out.clear();
    const uint8_t *q = p, *qe = p + n;
    p += n;
    uint32_t pending_hi = 0;
    auto emit = [&](uint32_t cp) {
      if (cp < 0x80) out.push_back((char)cp);
      else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
      else if (cp < 0x10000) {
        out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F)));
      } else {
        out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
        out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F)));
      }
    };
    while (q < qe) {
      uint32_t u;
      const uint8_t c = *q;
      if (c < 0x80) { u = c; q += 1; }
      else if ((c >> 5) == 0x6) { if (qe - q < 2) fail(MR_ERR_PARSE, "malformed modified UTF-8"); u = ((c & 0x1Fu) << 6) | (q[1] & 0x3Fu); q += 2; }
      else if ((c >> 4) == 0xE) { if (qe - q < 3) fail(MR_ERR_PARSE, "malformed modified UTF-8"); u = ((c & 0x0Fu) << 12) | ((q[1] & 0x3Fu) << 6) | (q[2] & 0x3Fu); q += 3; }
      else fail(MR_ERR_PARSE, "malformed modified UTF-8");
      if (pending_hi) {
        if (u >= 0xDC00 && u <= 0xDFFF) { emit(0x10000 + ((pending_hi - 0xD800) << 10) + (u - 0xDC00)); pending_hi = 0; continue; }
        emit(pending_hi);  // lone surrogate: kept as its 3-byte form
        pending_hi = 0;
      }
      if (u >= 0xD800 && u <= 0xDBFF) pending_hi = u;
      else emit(u);
    }
    if (pending_hi) emit(pending_hi);
  }
};

struct Out {
  std::vector<uint8_t> &b;
  template <class T> void put(T v) { const size_t o = b.size(); b.resize(o + sizeof(T)); memcpy(b.data() + o, &v, sizeof(T)); }
  void bytes(const void *p, size_t n) { const size_t o = b.size(); b.resize(o + n); if (n) memcpy(b.data() + o, p, n); }
};

uint64_t hash_str(const std::string &s) { return hash64(s.data(), s.size()); }

// Scalar -> upsert (kind, payload).  Returns the Scalar tag.
int scalar(In &in, Out &out, std::string &tmp) {
  const int tag = in.i8();
  switch (tag) {
    case 0: in.utf(tmp); out.put<uint8_t>(1); out.put<uint64_t>(hash_str(tmp)); break;           // SString
    case 1: { const double d = in.f64(); out.put<uint8_t>(0); out.put<double>(d); break; }        // SDouble
    case 2: { const uint8_t v = in.u8(); out.put<uint8_t>(7); out.put<uint8_t>(v ? 1 : 0); break; }  // SBoolean
    case 3: {                                                                                      // SStringList
      const int32_t n = in.varint();
      if (n < 0) fail(MR_ERR_PARSE, "negative list size");
      out.put<uint8_t>(2); out.put<uint32_t>((uint32_t)n);
      for (int32_t k = 0; k < n; k++) { in.utf(tmp); out.put<uint64_t>(hash_str(tmp)); }
      break;
    }
    case 4: {                                                                                      // SDoubleList
      const int32_t n = in.varint();
      if (n < 0) fail(MR_ERR_PARSE, "negative list size");
      out.put<uint8_t>(3); out.put<uint32_t>((uint32_t)n);
      for (int32_t k = 0; k < n; k++) out.put<double>(in.f64());
      break;
    }
    default: fail(MR_ERR_PARSE, "cannot decode scalar %d", tag);
  }
  return tag;
}

// One FeatureValue.  Appends its upsert record to `dst` and returns true, or returns false (dst untouched)
// for a class that has no reader.
bool feature_value(In in, std::vector<uint8_t> &dst) {
  std::vector<uint8_t> rec;
  Out out{rec};
  std::string s0, s1, s2, name, tmp;
  const int tag = in.i8();
  if (tag < 0 || tag > 13) fail(MR_ERR_PARSE, "cannot decode fv index %d", tag);
  const bool legacy = tag < 7;
  const int cls = tag % 7;  // 0 scalar 1 counter 2 numstats 3 map 4 periodic counter 5 frequency 6 bounded list
  // Key: scope then feature name (the upsert record wants the name first)
  const int scope = in.i8();
  int n_parts;
  switch (scope) {
    case 0: case 1: case 3: case 6: n_parts = 1; break;
    case 2: n_parts = 0; break;
    case 4: n_parts = 2; break;
    case 5: n_parts = 3; break;
    default: fail(MR_ERR_PARSE, "cannot parse scope with index %d", scope);
  }
  if (n_parts > 0) in.utf(s0);
  if (n_parts > 1) in.utf(s1);
  if (n_parts > 2) in.utf(s2);
  in.utf(name);
  if (name.size() > 65535) fail(MR_ERR_PARSE, "feature name too long");
  out.put<uint16_t>((uint16_t)name.size());
  out.bytes(name.data(), name.size());
  // reference scope tags (user 0, item 1, global 2, session 3, field 4, irf 5, ranking 6) -> upsert tags
  switch (scope) {
    case 0: out.put<uint8_t>(2); out.put<uint64_t>(hash_str(s0)); break;
    case 1: out.put<uint8_t>(1); out.put<uint64_t>(hash_str(s0)); break;
    case 2: out.put<uint8_t>(0); break;
    case 3: out.put<uint8_t>(3); out.put<uint64_t>(hash_str(s0)); break;
    case 4: out.put<uint8_t>(4); out.put<uint64_t>(hash_str(s1)); break;                                   // field value
    case 5: out.put<uint8_t>(5); out.put<uint64_t>(hash_str(s1)); out.put<uint64_t>(hash_str(s2)); break;  // value, item
    case 6: out.put<uint8_t>(6); out.put<uint64_t>(hash_str(s0)); break;
  }
  (void)in.varlong();  // ts
  bool supported = true;
  switch (cls) {
    case 0: scalar(in, out, tmp); break;
    case 1: out.put<uint8_t>(4); out.put<int64_t>(in.varlong()); break;
    case 2: {  // NumStatsValue: min, max, Map[Int, Double]
      in.f64(); in.f64();
      const int32_t n = in.varint();
      for (int32_t k = 0; k < n; k++) { in.varint(); in.f64(); }
      supported = false;
      break;
    }
    case 3: {  // MapValue: Map[String, Scalar]
      const int32_t n = in.varint();
      std::vector<uint8_t> sink;
      Out so{sink};
      for (int32_t k = 0; k < n; k++) { in.utf(tmp); scalar(in, so, tmp); sink.clear(); }
      supported = false;
      break;
    }
    case 4: {  // PeriodicCounterValue: Array[PeriodicValue(start, end, periods, value)] -> the values, in order
      const int32_t n = in.varint();
      if (n < 0) fail(MR_ERR_PARSE, "negative array size");
      out.put<uint8_t>(5); out.put<uint32_t>((uint32_t)n);
      for (int32_t k = 0; k < n; k++) { in.varlong(); in.varlong(); in.varint(); out.put<int64_t>(in.varlong()); }
      break;
    }
    case 5: {  // FrequencyValue: Map[String, Double]
      const int32_t n = in.varint();
      for (int32_t k = 0; k < n; k++) { in.utf(tmp); in.f64(); }
      supported = false;
      break;
    }
    case 6: {  // BoundedListValue: List[TimeValue(ts, Scalar)].  The reader keeps the list and collects its SString
               // entries only (S/feature/InteractedWithFeature.scala:117: values.map(_.value).collect { case SString(id) => id }):
               // entries of another scalar kind are skipped, the rest of the history survives
      const int32_t n = in.varint();
      if (n < 0) fail(MR_ERR_PARSE, "negative list size");
      std::vector<uint8_t> one, kept;
      Out oo{one};
      uint32_t n_kept = 0;
      for (int32_t k = 0; k < n; k++) {
        in.varlong();
        one.clear();
        if (scalar(in, oo, tmp) == 0) { kept.insert(kept.end(), one.begin() + 1, one.begin() + 9); n_kept++; }  // drop the kind byte
      }
      out.put<uint8_t>(6); out.put<uint32_t>(n_kept);
      out.bytes(kept.data(), kept.size());
      break;
    }
  }
  if (!legacy) (void)in.varlong();  // expire
  if (supported) dst.insert(dst.end(), rec.begin(), rec.end());
  return supported;
}

}  // namespace

void transcode_feature_values(const uint8_t *bytes, size_t len, std::vector<uint8_t> &out, FvStats &st) {
  size_t p = 0;
  while (len - p >= 4) {
    const int32_t n = (int32_t)(((uint32_t)bytes[p] << 24) | ((uint32_t)bytes[p + 1] << 16) | ((uint32_t)bytes[p + 2] << 8) | bytes[p + 3]);
    if (n < 0) fail(MR_ERR_PARSE, "negative record length %d at byte %zu", n, p);
    if ((size_t)n > len - p - 4) break;  // truncated tail: decodeDelimited's Right(None)
    In in{bytes + p + 4, bytes + p + 4 + n};
    st.records++;
    if (!feature_value(in, out)) st.unsupported++;
    p += 4 + (size_t)n;
  }
  st.consumed = p;
}

}  // namespace mr
