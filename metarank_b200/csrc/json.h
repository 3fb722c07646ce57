// Minimal JSON + UBJSON DOM reader (host side).  Used for XGBoost model bytes
// (`XGBoostBooster(bytes)`, reference S/ml/rank/LambdaMARTRanker.scala:230) and the
// feature-schema JSON handed to mr_schema_create.  Numbers keep their source token so
// that float fields can be rounded to binary32 exactly once (strtof on the token).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.h"

namespace mr {

struct JValue {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  double num = 0;      // value as double
  float f32 = 0;       // value rounded once to binary32 (token -> strtof, or the stored f32)
  bool is_int = false; // token had no '.', 'e', 'E' (or UBJSON integer tag)
  int64_t i64 = 0;
  std::string str;
  std::vector<JValue> arr;
  std::vector<std::pair<std::string, JValue>> obj;

  const JValue *get(const char *key) const {
    if (kind != Obj) return nullptr;
    for (auto &kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  const JValue &at(const char *key) const {
    const JValue *v = get(key);
    if (!v) fail(MR_ERR_PARSE, "json: missing key '%s'", key);
    return *v;
  }
  // numbers that XGBoost writes as strings ("num_feature":"30") or as numbers
  int64_t as_int() const {
    if (kind == Num) return is_int ? i64 : (int64_t)num;
    if (kind == Str) return strtoll(str.c_str(), nullptr, 10);
    if (kind == Bool) return b;
    fail(MR_ERR_PARSE, "json: expected integer");
  }
  double as_double() const {
    if (kind == Num) return num;
    if (kind == Str) return strtod(str.c_str(), nullptr);
    fail(MR_ERR_PARSE, "json: expected number");
  }
  float as_f32() const {
    if (kind == Num) return f32;
    if (kind == Str) {
      const char *s = str.c_str();
      while (*s == '[' || *s == ' ') s++;  // XGBoost 3.1 writes base_score as "[5E-1]"
      return strtof(s, nullptr);
    }
    fail(MR_ERR_PARSE, "json: expected float");
  }
};

// The JSON number grammar: -?(0|[1-9][0-9]*)(\.[0-9]+)?([eE][+-]?[0-9]+)?
inline bool json_number_token_ok(const char *buf, size_t n) {
  size_t k = 0;
  auto dig = [&](size_t at) { return at < n && buf[at] >= '0' && buf[at] <= '9'; };
  if (k < n && buf[k] == '-') k++;
  if (!dig(k)) return false;
  if (buf[k] == '0') k++;
  else while (dig(k)) k++;
  if (k < n && buf[k] == '.') { k++; if (!dig(k)) return false; while (dig(k)) k++; }
  if (k < n && (buf[k] == 'e' || buf[k] == 'E')) {
    k++;
    if (k < n && (buf[k] == '+' || buf[k] == '-')) k++;
    if (!dig(k)) return false;
    while (dig(k)) k++;
  }
  return k == n;
}

class JsonParser {
 public:
  // strict: numbers must follow the JSON grammar (request bodies, where the reference's jawn parser rejects
  // '+1', '01', '1.', NaN); model files keep the lenient reader (XGBoost writes NaN / Infinity tokens)
  JsonParser(const uint8_t *p, size_t n, bool strict = false) : p_(p), e_(p + n), strict_(strict) {}
  JValue parse() {
    JValue v = value(0);
    ws();
    return v;
  }
  const uint8_t *pos() const { return p_; }  // after parse(): first byte behind the value and its trailing blanks

 private:
  const uint8_t *p_, *e_;
  bool strict_ = false;
  void ws() {
    while (p_ < e_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) p_++;
  }
  [[noreturn]] void bad(const char *what) { fail(MR_ERR_PARSE, "json: %s", what); }
  JValue value(int depth) {
    if (depth > 64) bad("nesting too deep");
    ws();
    if (p_ >= e_) bad("unexpected end");
    JValue v;
    switch (*p_) {
      case '{': {
        v.kind = JValue::Obj;
        p_++;
        ws();
        if (p_ < e_ && *p_ == '}') { p_++; return v; }
        for (;;) {
          ws();
          if (p_ >= e_ || *p_ != '"') bad("expected object key");
          std::string k = string();
          ws();
          if (p_ >= e_ || *p_ != ':') bad("expected ':'");
          p_++;
          v.obj.emplace_back(std::move(k), value(depth + 1));
          ws();
          if (p_ < e_ && *p_ == ',') { p_++; continue; }
          if (p_ < e_ && *p_ == '}') { p_++; return v; }
          bad("expected ',' or '}'");
        }
      }
      case '[': {
        v.kind = JValue::Arr;
        p_++;
        ws();
        if (p_ < e_ && *p_ == ']') { p_++; return v; }
        for (;;) {
          v.arr.push_back(value(depth + 1));
          ws();
          if (p_ < e_ && *p_ == ',') { p_++; continue; }
          if (p_ < e_ && *p_ == ']') { p_++; return v; }
          bad("expected ',' or ']'");
        }
      }
      case '"':
        v.kind = JValue::Str;
        v.str = string();
        return v;
      case 't':
        lit("true"); v.kind = JValue::Bool; v.b = true; return v;
      case 'f':
        lit("false"); v.kind = JValue::Bool; v.b = false; return v;
      case 'n':
        lit("null"); return v;
      default:
        return number();
    }
  }
  void lit(const char *s) {
    size_t n = strlen(s);
    if ((size_t)(e_ - p_) < n || memcmp(p_, s, n) != 0) bad("bad literal");
    p_ += n;
  }
  JValue number() {
    // JSON numbers plus the non-standard NaN / Infinity tokens XGBoost may emit
    char buf[64];
    size_t n = 0;
    bool is_int = true;
    while (p_ < e_ && n < sizeof(buf) - 1) {
      char c = (char)*p_;
      if ((c >= '0' && c <= '9') || c == '-' || c == '+') {
      } else if (c == '.' || c == 'e' || c == 'E' || c == 'N' || c == 'a' || c == 'I' || c == 'n' ||
                 c == 'f' || c == 'i' || c == 't' || c == 'y') {
        is_int = false;
      } else {
        break;
      }
      buf[n++] = c;
      p_++;
    }
    if (n == 0) bad("unexpected character");
    buf[n] = 0;
    if (strict_ && !json_number_token_ok(buf, n)) bad("bad number");
    JValue v;
    v.kind = JValue::Num;
    v.is_int = is_int;
    if (is_int && n <= 15) {
      // short integers (ids, counts, timestamps): exact in every representation, no libc round trips
      size_t k = (buf[0] == '-' || buf[0] == '+') ? 1 : 0;
      bool plain = k < n;
      int64_t acc = 0;
      for (size_t j = k; j < n && plain; j++) {
        if (buf[j] < '0' || buf[j] > '9') plain = false;
        else acc = acc * 10 + (buf[j] - '0');
      }
      if (plain && !(buf[0] == '-' && acc == 0)) {  // "-0" keeps its sign through strtod
        v.i64 = buf[0] == '-' ? -acc : acc;
        v.num = (double)v.i64;
        v.f32 = (float)v.i64;  // |i64| < 2^53: one rounding, same as strtof on the token
        return v;
      }
    }
    char *end = nullptr;
    v.num = strtod(buf, &end);
    if (end == buf) bad("bad number");
    v.f32 = strtof(buf, nullptr);
    if (is_int) v.i64 = strtoll(buf, nullptr, 10);
    return v;
  }
  std::string string() {
    std::string s;
    p_++;  // opening quote
    while (p_ < e_ && *p_ != '"') {
      if (*p_ == '\\') {
        p_++;
        if (p_ >= e_) bad("bad escape");
        switch (*p_) {
          case 'n': s += '\n'; break;
          case 't': s += '\t'; break;
          case 'r': s += '\r'; break;
          case 'b': s += '\b'; break;
          case 'f': s += '\f'; break;
          case 'u': {
            if (e_ - p_ < 5) bad("bad \\u escape");
            unsigned cp = (unsigned)strtoul(std::string((const char *)p_ + 1, 4).c_str(), nullptr, 16);
            p_ += 4;
            // a UTF-16 surrogate pair written as two escapes is ONE code point (ids are hashed as UTF-8)
            if (cp >= 0xD800 && cp <= 0xDBFF && e_ - p_ >= 7 && p_[1] == '\\' && p_[2] == 'u') {
              const unsigned lo = (unsigned)strtoul(std::string((const char *)p_ + 3, 4).c_str(), nullptr, 16);
              if (lo >= 0xDC00 && lo <= 0xDFFF) {
                cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                p_ += 6;
                s += (char)(0xF0 | (cp >> 18)); s += (char)(0x80 | ((cp >> 12) & 0x3F));
                s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F));
                break;
              }
            }
            if (cp < 0x80) s += (char)cp;
            else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
            else { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: s += (char)*p_;
        }
        p_++;
      } else {
        // bulk-copy the run up to the next quote or escape
        const uint8_t *q = p_;
        while (q < e_ && *q != '"' && *q != '\\') q++;
        s.append((const char *)p_, (size_t)(q - p_));
        p_ = q;
      }
    }
    if (p_ >= e_) bad("unterminated string");
    p_++;
    return s;
  }
};

// UBJSON (draft 12) as written by XGBoost's UBJWriter: big-endian scalars, typed
// arrays `[$<type>#<count>`, object keys as length-prefixed strings without 'S'.
class UbjParser {
 public:
  UbjParser(const uint8_t *p, size_t n) : p_(p), e_(p + n) {}
  JValue parse() { return value(take(), 0); }

 private:
  const uint8_t *p_, *e_;
  uint8_t take() {
    if (p_ >= e_) fail(MR_ERR_PARSE, "ubjson: unexpected end");
    return *p_++;
  }
  template <class T> T be() {
    if ((size_t)(e_ - p_) < sizeof(T)) fail(MR_ERR_PARSE, "ubjson: truncated");
    uint8_t tmp[sizeof(T)];
    for (size_t i = 0; i < sizeof(T); i++) tmp[i] = p_[sizeof(T) - 1 - i];
    p_ += sizeof(T);
    T v;
    memcpy(&v, tmp, sizeof(T));
    return v;
  }
  int64_t integer(uint8_t tag) {
    switch (tag) {
      case 'i': return be<int8_t>();
      case 'U': return be<uint8_t>();
      case 'I': return be<int16_t>();
      case 'l': return be<int32_t>();
      case 'L': return be<int64_t>();
    }
    fail(MR_ERR_PARSE, "ubjson: expected integer tag, got 0x%02x", tag);
  }
  std::string rawstr() {
    int64_t n = integer(take());
    if (n < 0 || n > e_ - p_) fail(MR_ERR_PARSE, "ubjson: bad string length");
    std::string s((const char *)p_, (size_t)n);
    p_ += n;
    return s;
  }
  JValue scalar(uint8_t tag) {
    JValue v;
    switch (tag) {
      case 'Z': return v;
      case 'T': v.kind = JValue::Bool; v.b = true; return v;
      case 'F': v.kind = JValue::Bool; v.b = false; return v;
      case 'i': case 'U': case 'I': case 'l': case 'L':
        v.kind = JValue::Num; v.is_int = true; v.i64 = integer(tag); v.num = (double)v.i64; v.f32 = (float)v.i64;
        return v;
      case 'd': v.kind = JValue::Num; v.f32 = be<float>(); v.num = v.f32; return v;
      case 'D': v.kind = JValue::Num; v.num = be<double>(); v.f32 = (float)v.num; return v;
      case 'C': v.kind = JValue::Str; v.str = std::string(1, (char)take()); return v;
      case 'S': v.kind = JValue::Str; v.str = rawstr(); return v;
    }
    fail(MR_ERR_PARSE, "ubjson: unknown tag 0x%02x", tag);
  }
  JValue value(uint8_t tag, int depth) {
    if (depth > 64) fail(MR_ERR_PARSE, "ubjson: nesting too deep");
    JValue v;
    if (tag == '{') {
      v.kind = JValue::Obj;
      for (;;) {
        if (p_ < e_ && *p_ == '}') { p_++; return v; }
        std::string k = rawstr();
        v.obj.emplace_back(std::move(k), value(take(), depth + 1));
      }
    }
    if (tag == '[') {
      v.kind = JValue::Arr;
      if (p_ < e_ && *p_ == '$') {
        p_++;
        uint8_t ety = take();
        if (take() != '#') fail(MR_ERR_PARSE, "ubjson: typed array without count");
        int64_t n = integer(take());
        if (n < 0) fail(MR_ERR_PARSE, "ubjson: negative count");
        v.arr.reserve((size_t)n);
        for (int64_t i = 0; i < n; i++) v.arr.push_back(scalar(ety));
        return v;
      }
      if (p_ < e_ && *p_ == '#') {
        p_++;
        int64_t n = integer(take());
        for (int64_t i = 0; i < n; i++) v.arr.push_back(value(take(), depth + 1));
        return v;
      }
      for (;;) {
        uint8_t t = take();
        if (t == ']') return v;
        v.arr.push_back(value(t, depth + 1));
      }
    }
    return scalar(tag);
  }
};

}  // namespace mr
