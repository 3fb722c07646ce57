// RankingEvent JSON -> mr_rank_batch arrays.  What each rule follows (S = src/main/scala/ai/metarank):
//   event shape      S/model/Event.scala:44-99 (id, timestamp, user?, session?, fields?, items[{id, relevancy?,
//                    fields?, label?}]; `relevancy` is sugar for a leading NumberField("relevancy", r), :84-93)
//   fields           S/model/Field.scala:36-58 (string | bool | number | string[] | number[]; null / object /
//                    mixed list are decoding failures)
//   timestamp        S/model/Timestamp.scala (long | numeric string | ISO date-time with a zone)
//   request reads    number / word_count on a ranking field (S/feature/NumberFeature.scala:58-69,
//                    WordCountFeature.scala:63-68), string on a ranking field (StringFeature.scala:96-116),
//                    rate scoped ranking.<field> (RateFeature.scala), item_age (ItemAgeFeature.scala:74-86),
//                    local_time (LocalDateTimeFeature.scala:44-80), field_match (FieldMatchFeature.scala:60-68,
//                    FieldMatchBiencoderFeature.scala:84-88), relevancy (RelevancyFeature.scala:36-51), per-item
//                    overrides of number / string (NumberFeature.scala:84-93)
// The Python shim (metarank_b200/features.py pack_requests + rank_api.py decode_ranking_event) is the same logic
// and is what tests/test_request_codec_cpu.py compares this file with, array for array.
#include "request_codec.h"

#include <atomic>
#include <exception>
#include <mutex>
#include <system_error>
#include <thread>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_map>

#include "json.h"

namespace mr {
namespace {

// ---------------------------------------------------------------- decoded event
struct FieldVal {
  enum Kind { Str, Bool, Num, StrList, NumList } kind = Str;
  std::string s;
  bool b = false;
  double d = 0;
  std::vector<std::string> sl;
  std::vector<double> dl;
  bool is_str() const { return kind == Str; }
  bool is_num() const { return kind == Num; }
  bool is_strlist() const { return kind == StrList; }
};
struct Field { std::string name; FieldVal v; };

// days since 1970-01-01 of a proleptic Gregorian date, and back (H. Hinnant's civil algorithms)
int64_t days_from_civil(int64_t y, unsigned m, unsigned d) {
  y -= m <= 2;
  const int64_t era = (y >= 0 ? y : y - 399) / 400;
  const unsigned yoe = (unsigned)(y - era * 400);
  const unsigned doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
  const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return era * 146097 + (int64_t)doe - 719468;
}
void civil_from_days(int64_t z, int64_t &y, unsigned &m, unsigned &d) {
  z += 719468;
  const int64_t era = (z >= 0 ? z : z - 146096) / 146097;
  const unsigned doe = (unsigned)(z - era * 146097);
  const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  y = (int64_t)yoe + era * 400;
  const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  const unsigned mp = (5 * doy + 2) / 153;
  d = doy - (153 * mp + 2) / 5 + 1;
  m = mp < 10 ? mp + 3 : mp - 9;
  y += m <= 2;
}

struct ZonedTime {  // wall-clock fields in the value's own offset + the instant
  int64_t year = 1970;
  unsigned month = 1, day = 1, hour = 0, minute = 0, second = 0;
  int64_t epoch_ms = 0;
  int iso_weekday() const {  // Monday = 1 .. Sunday = 7 (java.time.DayOfWeek.getValue)
    const int64_t days = days_from_civil(year, month, day);
    int64_t wd = (days + 3) % 7;  // 1970-01-01 was a Thursday
    if (wd < 0) wd += 7;
    return (int)wd + 1;
  }
};

// ZonedDateTime.parse(s, ISO_DATE_TIME): yyyy-MM-ddTHH:mm[:ss[.fraction]] then Z | +HH[:MM[:SS]] | +HHMM, an optional
// [Region/Id] suffix is dropped.  No zone -> not parsed (the extractor then yields no value).
bool parse_iso_zoned(const std::string &in, ZonedTime &out) {
  std::string s = in;
  if (!s.empty() && s.back() == ']') {
    const size_t lb = s.find('[');
    if (lb == std::string::npos) return false;
    s.resize(lb);
  }
  auto digits = [&](size_t pos, int n, int &v) -> bool {
    if (pos + n > s.size()) return false;
    v = 0;
    for (int k = 0; k < n; k++) { const char c = s[pos + k]; if (c < '0' || c > '9') return false; v = v * 10 + (c - '0'); }
    return true;
  };
  int Y, M, D, h, mi, se = 0;
  if (!digits(0, 4, Y) || s.size() < 16 || s[4] != '-' || !digits(5, 2, M) || s[7] != '-' || !digits(8, 2, D) || s[10] != 'T' ||
      !digits(11, 2, h) || s[13] != ':' || !digits(14, 2, mi))
    return false;
  size_t p = 16;
  int64_t frac_ms = 0;
  if (p < s.size() && s[p] == ':') {
    if (!digits(p + 1, 2, se)) return false;
    p += 3;
    if (p < s.size() && (s[p] == '.' || s[p] == ',')) {
      p++;
      int nd = 0;
      while (p < s.size() && s[p] >= '0' && s[p] <= '9') { if (nd < 3) frac_ms = frac_ms * 10 + (s[p] - '0'); nd++; p++; }
      if (nd == 0) return false;
      for (; nd < 3; nd++) frac_ms *= 10;
    }
  }
  if (M < 1 || M > 12 || D < 1 || D > 31 || h > 23 || mi > 59 || se > 59) return false;
  int64_t off_s = 0;
  if (p >= s.size()) return false;  // zone required
  if (s[p] == 'Z' || s[p] == 'z') { if (p + 1 != s.size()) return false; }
  else if (s[p] == '+' || s[p] == '-') {
    const int sign = s[p] == '-' ? -1 : 1;
    int oh, om = 0, os = 0;
    if (!digits(p + 1, 2, oh)) return false;
    p += 3;
    if (p < s.size()) {
      if (s[p] == ':') p++;
      if (!digits(p, 2, om)) return false;
      p += 2;
      if (p < s.size()) {
        if (s[p] == ':') p++;
        if (!digits(p, 2, os)) return false;
        p += 2;
      }
    }
    if (p != s.size()) return false;
    off_s = sign * ((int64_t)oh * 3600 + om * 60 + os);
  } else return false;
  out.year = Y; out.month = (unsigned)M; out.day = (unsigned)D; out.hour = (unsigned)h; out.minute = (unsigned)mi; out.second = (unsigned)se;
  const int64_t local_s = days_from_civil(Y, (unsigned)M, (unsigned)D) * 86400 + h * 3600 + mi * 60 + se;
  out.epoch_ms = (local_s - off_s) * 1000 + frac_ms;
  return true;
}

ZonedTime utc_from_epoch_seconds(int64_t sec) {
  ZonedTime z;
  int64_t days = sec / 86400, rem = sec % 86400;
  if (rem < 0) { rem += 86400; days -= 1; }
  civil_from_days(days, z.year, z.month, z.day);
  z.hour = (unsigned)(rem / 3600); z.minute = (unsigned)((rem % 3600) / 60); z.second = (unsigned)(rem % 60);
  z.epoch_ms = sec * 1000;
  return z;
}

int64_t floor_div(int64_t a, int64_t b) { int64_t q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) q--; return q; }

int64_t decode_timestamp(const JValue &v) {
  if (v.kind == JValue::Num) {
    if (v.is_int) return v.i64;
    if (v.num == std::floor(v.num) && std::fabs(v.num) < 9.2e18) return (int64_t)v.num;
    fail(MR_ERR_PARSE, "cannot decode timestamp");
  }
  if (v.kind == JValue::Str) {
    size_t a = 0, b = v.str.size();
    while (a < b && isspace((unsigned char)v.str[a])) a++;
    while (b > a && isspace((unsigned char)v.str[b - 1])) b--;
    const std::string s = v.str.substr(a, b - a);
    bool numeric = !s.empty();
    for (size_t k = 0; k < s.size(); k++) numeric &= (s[k] >= '0' && s[k] <= '9') || (k == 0 && s[k] == '-' && s.size() > 1);
    if (numeric) return strtoll(s.c_str(), nullptr, 10);
    ZonedTime z;
    if (parse_iso_zoned(s, z)) return z.epoch_ms;
  }
  fail(MR_ERR_PARSE, "cannot decode timestamp");
}

struct Item { std::string id; std::vector<Field> fields; };
struct Event {
  std::string id;
  int64_t ts = 0;
  bool has_user = false, has_session = false;
  std::string user, session;
  std::vector<Field> fields;
  std::vector<Item> items;
  JValue embeddings, tokens;  // extensions (kept as small DOMs): caller-side model outputs / analyzers
};

// The body is walked once with a small tokenizer: no DOM (building ~600 JSON values for a 100-item request was
// 3/4 of the decoder's time), strings go straight into their destination, unknown keys are skipped in place.
// Only the rare "embeddings" / "tokens" objects and the timestamp go through the shared JSON reader.
struct Cur {
  const uint8_t *p, *e;
  [[noreturn]] void bad(const char *what) { fail(MR_ERR_PARSE, "json: %s", what); }
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
  char peek() { ws(); return p < e ? (char)*p : '\0'; }
  bool eat(char c) { ws(); if (p < e && *p == (uint8_t)c) { p++; return true; } return false; }
  void expect(char c, const char *what) { if (!eat(c)) bad(what); }
  bool lit(const char *w) {
    const size_t n = strlen(w);
    if ((size_t)(e - p) >= n && memcmp(p, w, n) == 0) { p += n; return true; }
    return false;
  }
  static void utf8(std::string &out, unsigned cp) {
    if (cp < 0x80) out += (char)cp;
    else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
    else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
  }
  unsigned hex4() {
    if (e - p < 4) bad("bad \\u escape");
    unsigned v = 0;
    for (int k = 0; k < 4; k++) {
      const uint8_t c = p[k];
      if (!((c >= '0' && c <= '9') || (c >= 'a' && c <= 'f') || (c >= 'A' && c <= 'F'))) bad("bad \\u escape");  // jawn rejects it too
      v = v * 16 + (c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c - 'A' + 10);
    }
    p += 4;
    return v;
  }
  // a JSON string into `out` (cleared first); the same escape rules as json.h (surrogate pairs joined)
  void str(std::string &out) {
    ws();
    if (p >= e || *p != '"') bad("expected a string");
    p++;
    out.clear();
    for (;;) {
      const uint8_t *q = p;
      while (q < e && *q != '"' && *q != '\\') q++;
      out.append((const char *)p, (size_t)(q - p));
      p = q;
      if (p >= e) bad("unterminated string");
      if (*p == '"') { p++; return; }
      p++;  // backslash
      if (p >= e) bad("unterminated string");
      const char c = (char)*p++;
      switch (c) {
        case 'n': out += '\n'; break;
        case 't': out += '\t'; break;
        case 'r': out += '\r'; break;
        case 'b': out += '\b'; break;
        case 'f': out += '\f'; break;
        case 'u': {
          unsigned cp = hex4();
          if (cp >= 0xD800 && cp <= 0xDBFF && e - p >= 6 && p[0] == '\\' && p[1] == 'u') {
            const uint8_t *save = p;
            p += 2;
            const unsigned lo = hex4();
            if (lo >= 0xDC00 && lo <= 0xDFFF) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            else p = save;
          }
          utf8(out, cp);
          break;
        }
        default: out += c;  // \" \\ \/ and anything else: the character itself
      }
    }
  }
  // a number token (finite JSON numbers only: NaN / Infinity are not JSON here)
  void number(double &d, bool &is_int, int64_t &i64) {
    ws();
    char buf[64];
    size_t n = 0;
    is_int = true;
    while (p < e && n < sizeof(buf) - 1) {
      const char c = (char)*p;
      if ((c >= '0' && c <= '9') || c == '-' || c == '+') {}
      else if (c == '.' || c == 'e' || c == 'E') is_int = false;
      else break;
      buf[n++] = c;
      p++;
    }
    if (n == 0) bad("unexpected character");
    buf[n] = 0;
    {  // the JSON grammar, as jawn enforces it: -?(0|[1-9][0-9]*)(\.[0-9]+)?([eE][+-]?[0-9]+)?  ('+1', '01', '1.', '.5' are not numbers)
      size_t k = 0;
      auto dig = [&](size_t at) { return at < n && buf[at] >= '0' && buf[at] <= '9'; };
      if (buf[k] == '-') k++;
      if (!dig(k)) bad("bad number");
      if (buf[k] == '0') k++;
      else while (dig(k)) k++;
      if (k < n && buf[k] == '.') { k++; if (!dig(k)) bad("bad number"); while (dig(k)) k++; }
      if (k < n && (buf[k] == 'e' || buf[k] == 'E')) {
        k++;
        if (k < n && (buf[k] == '+' || buf[k] == '-')) k++;
        if (!dig(k)) bad("bad number");
        while (dig(k)) k++;
      }
      if (k != n) bad("bad number");
    }
    char *end = nullptr;
    d = strtod(buf, &end);
    if (end == buf || *end != 0) bad("bad number");
    i64 = is_int ? strtoll(buf, nullptr, 10) : 0;
  }
  void skip(int depth = 0) {  // any value, unparsed
    if (depth > 64) bad("nesting too deep");
    switch (peek()) {
      case '{': {
        p++;
        if (eat('}')) return;
        std::string k;
        for (;;) { str(k); expect(':', "expected ':'"); skip(depth + 1); if (eat(',')) continue; expect('}', "expected ',' or '}'"); return; }
      }
      case '[': {
        p++;
        if (eat(']')) return;
        for (;;) { skip(depth + 1); if (eat(',')) continue; expect(']', "expected ',' or ']'"); return; }
      }
      case '"': { std::string s; str(s); return; }
      case 't': if (!lit("true")) bad("bad literal"); return;
      case 'f': if (!lit("false")) bad("bad literal"); return;
      case 'n': if (!lit("null")) bad("bad literal"); return;
      default: { double d; bool ii; int64_t i; number(d, ii, i); return; }
    }
  }
  JValue dom() {  // a whole value through the shared reader (rare keys only)
    ws();
    if (p >= e) bad("unexpected end of input");
    JsonParser jp(p, (size_t)(e - p), /*strict=*/true);
    JValue v = jp.parse();
    p = jp.pos();
    return v;
  }
  template <class F> void object(F &&on_key, const char *what) {  // on_key(key) consumes the value
    if (!eat('{')) fail(MR_ERR_PARSE, "%s", what);
    if (eat('}')) return;
    std::string k;
    for (;;) {
      str(k);
      expect(':', "expected ':'");
      on_key(k);
      if (eat(',')) continue;
      expect('}', "expected ',' or '}'");
      return;
    }
  }
  template <class F> void array(F &&on_elem) {  // caller checked peek() == '['
    expect('[', "expected '['");
    if (eat(']')) return;
    for (;;) {
      on_elem();
      if (eat(',')) continue;
      expect(']', "expected ',' or ']'");
      return;
    }
  }
};

// the `value` of a Field (S/model/Field.scala:36-58)
void stream_field_value(Cur &c, const std::string &name, FieldVal &v) {
  switch (c.peek()) {
    case 'n': fail(MR_ERR_PARSE, "null value in field %s", name.c_str());
    case '{': fail(MR_ERR_PARSE, "cannot decode field %s: got object", name.c_str());
    case '"': v.kind = FieldVal::Str; c.str(v.s); return;
    case 't': if (!c.lit("true")) c.bad("bad literal"); v.kind = FieldVal::Bool; v.b = true; return;
    case 'f': if (!c.lit("false")) c.bad("bad literal"); v.kind = FieldVal::Bool; v.b = false; return;
    case '[': {
      bool all_str = true, all_num = true;
      std::string s;
      c.array([&] {
        const char k = c.peek();
        if (k == '"') { c.str(s); v.sl.push_back(s); all_num = false; }
        else if (k == '{' || k == '[' || k == 't' || k == 'f' || k == 'n') { c.skip(); all_str = all_num = false; }
        else { double d; bool ii; int64_t i; c.number(d, ii, i); v.dl.push_back(d); all_str = false; }
      });
      if (all_str) { v.kind = FieldVal::StrList; v.dl.clear(); }
      else if (all_num) { v.kind = FieldVal::NumList; v.sl.clear(); }
      else fail(MR_ERR_PARSE, "cannot decode field %s: got a mixed list", name.c_str());
      return;
    }
    default: {
      bool ii; int64_t i;
      c.number(v.d, ii, i);  // NaN / Infinity tokens are not numbers here (T/model/FieldTest.scala:15-17)
      v.kind = FieldVal::Num;
      return;
    }
  }
}

void stream_field(Cur &c, Field &f) {
  bool has_name = false, has_value = false;
  // "value" may precede "name": remember where it starts and decode it once the name is known
  const uint8_t *value_at = nullptr;
  c.object([&](const std::string &k) {
    if (k == "name" && !has_name) {
      if (c.peek() != '"') fail(MR_ERR_PARSE, "field needs a string 'name'");
      c.str(f.name);
      has_name = true;
    } else if (k == "value" && !has_value) {
      c.ws();
      value_at = c.p;
      c.skip();
      has_value = true;
    } else {
      c.skip();
    }
  }, "field needs a string 'name'");
  if (!has_name) fail(MR_ERR_PARSE, "field needs a string 'name'");
  if (!has_value) fail(MR_ERR_PARSE, "field value not found");
  Cur vc{value_at, c.e};
  stream_field_value(vc, f.name, f.v);
}

void stream_fields(Cur &c, std::vector<Field> &out) {
  if (c.peek() != '[') { c.skip(); return; }  // not a list: ignored
  c.array([&] { out.emplace_back(); stream_field(c, out.back()); });
}

Event decode_event(Cur &c) {
  Event e;
  bool has_id = false, has_ts = false, has_items = false, seen_user = false, seen_session = false, has_fields = false;
  c.object([&](const std::string &k) {
    // a repeated key overwrites the earlier one, as in circe's JsonObject (and json.loads)
    if (k == "id") {
      if (c.peek() == '"') c.str(e.id);
      else {
        double d; bool ii; int64_t i;
        c.number(d, ii, i);
        if (!ii) fail(MR_ERR_PARSE, "event id must be a string");
        e.id = std::to_string(i);
      }
      has_id = true;
    } else if (k == "timestamp") {
      e.ts = decode_timestamp(c.dom());
      has_ts = true;
    } else if (k == "user" || k == "session") {
      const bool user = k == "user";
      (user ? seen_user : seen_session) = true;
      (user ? e.has_user : e.has_session) = false;
      if (c.peek() == 'n') { if (!c.lit("null")) c.bad("bad literal"); return; }
      if (c.peek() != '"') fail(MR_ERR_PARSE, "'%s' must be a string", k.c_str());
      c.str(user ? e.user : e.session);
      (user ? e.has_user : e.has_session) = true;
    } else if (k == "fields") {
      has_fields = true;
      e.fields.clear();
      stream_fields(c, e.fields);
    } else if (k == "items") {
      has_items = true;
      e.items.clear();
      if (c.peek() != '[') fail(MR_ERR_PARSE, "items must be a non-empty list");
      c.array([&] {
        e.items.emplace_back();
        Item &I = e.items.back();
        bool has_iid = false, seen_rel = false, has_rel = false, has_ifields = false;
        double rel = 0;
        c.object([&](const std::string &ik) {
          if (ik == "id") {
            if (c.peek() != '"') fail(MR_ERR_PARSE, "item needs a string 'id'");
            c.str(I.id);
            has_iid = true;
          } else if (ik == "relevancy") {
            seen_rel = true;
            has_rel = false;
            const char pk = c.peek();
            if (pk == 'n') { if (!c.lit("null")) c.bad("bad literal"); return; }
            if (pk == '"' || pk == '{' || pk == '[' || pk == 't' || pk == 'f') fail(MR_ERR_PARSE, "relevancy must be a number");
            bool ii; int64_t i;
            c.number(rel, ii, i);
            has_rel = true;
          } else if (ik == "fields") {
            has_ifields = true;
            I.fields.clear();
            stream_fields(c, I.fields);
          } else {
            c.skip();  // label and anything else
          }
        }, "item needs a string 'id'");
        if (!has_iid) fail(MR_ERR_PARSE, "item needs a string 'id'");
        if (has_rel) {  // the sugar comes first (Event.scala:84-93)
          Field f;
          f.name = "relevancy"; f.v.kind = FieldVal::Num; f.v.d = rel;
          I.fields.insert(I.fields.begin(), std::move(f));
        }
      });
      if (e.items.empty()) fail(MR_ERR_PARSE, "items must be a non-empty list");  // NonEmptyList
    } else if (k == "embeddings") {
      e.embeddings = c.dom();
    } else if (k == "tokens") {
      e.tokens = c.dom();
    } else {
      c.skip();
    }
  }, "ranking event must be a JSON object");
  if (!has_id) fail(MR_ERR_PARSE, "required field 'id' missing in JSON");
  if (!has_ts) fail(MR_ERR_PARSE, "required field 'timestamp' missing in JSON");
  if (!has_items) fail(MR_ERR_PARSE, "required field 'items' missing in JSON");
  return e;
}

// ---------------------------------------------------------------- per-feature request plan (from the schema JSON)
struct ReqFeature {
  std::string name, type, scope;
  std::string src_event, src_field;           // number / word_count / string / boolean: source (or legacy `field`)
  bool encode_index = false;
  std::vector<std::string> values;
  std::string rate_field;                     // rate scoped ranking.<field>
  std::string lt_field, lt_parse;             // local_time
  std::string rank_field;                     // field_match: ranking field
  std::string mtype, language;
  int n = 0;
  double docs = 0;
  std::unordered_map<std::string, double> termfreq;
  int s_f64 = -1, s_u64 = -1, s_item = -1, s_vec = -1, s_tok = -1, vec_off = 0, vec_dim = 0, dim = 1;
  int item_kind = 0;  // what the per-item loop reads for this feature: 0 nothing, 1 relevancy, 2 number override, 3 string override
};

void split_field(const std::string &s, std::string &ev, std::string &fld) {
  const size_t dot = s.find('.');
  ev = dot == std::string::npos ? s : s.substr(0, dot);
  fld = dot == std::string::npos ? "" : s.substr(dot + 1);
  if (ev == "metadata") ev = "item";
}

int slot_of(const std::vector<std::string> &v, const std::string &n) {
  for (size_t k = 0; k < v.size(); k++) if (v[k] == n) return (int)k;
  return -1;
}

std::vector<ReqFeature> build_features(const Schema &S) {
  JValue doc = JsonParser((const uint8_t *)S.source_json.data(), S.source_json.size()).parse();
  std::unordered_map<std::string, const JValue *> by_name;
  for (auto &o : doc.at("features").arr)
    if (const JValue *n = o.get("name")) by_name[n->str] = &o;
  std::vector<ReqFeature> plan;
  auto sget = [](const JValue &o, const char *k) -> std::string { const JValue *v = o.get(k); return v && v->kind == JValue::Str ? v->str : ""; };
  for (auto &name : S.model_features) {
    auto it = by_name.find(name);
    if (it == by_name.end()) continue;
    const JValue &o = *it->second;
    ReqFeature f;
    f.name = name;
    f.type = sget(o, "type");
    f.scope = sget(o, "scope");
    const std::string src = o.get("source") ? sget(o, "source") : sget(o, "field");
    if (!src.empty()) split_field(src, f.src_event, f.src_field);
    f.encode_index = sget(o, "encode") == "index";
    if (const JValue *vals = o.get("values"))
      if (vals->kind == JValue::Arr) for (auto &x : vals->arr) f.values.push_back(x.str);
    if (f.type == "rate" && f.scope.rfind("ranking.", 0) == 0) f.rate_field = f.scope.substr(8);
    if (f.type == "local_time") { std::string ev; split_field(sget(o, "source"), ev, f.lt_field); f.lt_parse = sget(o, "parse"); }
    if (f.type == "field_match") {
      std::string ev;
      split_field(sget(o, "rankingField"), ev, f.rank_field);
      const JValue &m = o.at("method");
      f.mtype = sget(m, "type");
      f.language = sget(m, "language");
      if (const JValue *n = m.get("n")) f.n = (int)n->as_int();
      if (const JValue *d = m.get("docs")) f.docs = d->as_double();
      if (const JValue *tf = m.get("termfreq"))
        if (tf->kind == JValue::Obj) for (auto &kv : tf->obj) f.termfreq[kv.first] = kv.second.as_double();
    }
    f.s_f64 = slot_of(S.in_req_f64, name);
    f.s_u64 = slot_of(S.in_req_u64, name);
    f.s_item = slot_of(S.in_item_f64, name);
    f.s_tok = slot_of(S.in_req_tok, name);
    for (size_t k = 0; k < S.in_req_vec.size(); k++)
      if (S.in_req_vec[k].feature == name) { f.s_vec = (int)k; f.vec_off = S.in_req_vec[k].offset; f.vec_dim = S.in_req_vec[k].dim; }
    auto c = S.col_of.find(name);
    if (c != S.col_of.end()) f.dim = c->second.second;
    if (f.s_item >= 0)
      f.item_kind = f.type == "relevancy" ? 1 : (f.type == "number" && f.scope != "ranking") ? 2 : (f.type == "string" && f.src_event != "ranking") ? 3 : 0;
    plan.push_back(std::move(f));
  }
  return plan;
}

// StringFeature encoders on request-side values (StringFeature.scala:118-137)
void encode_string(const ReqFeature &f, const std::vector<std::string> &vals, double *dst) {
  auto index_of = [&](const std::string &v) -> int { for (size_t k = 0; k < f.values.size(); k++) if (f.values[k] == v) return (int)k; return -1; };
  if (f.encode_index) {
    const int ix = vals.empty() ? -1 : index_of(vals[0]);
    dst[0] = ix >= 0 ? (double)(ix + 1) : 0.0;
    return;
  }
  for (size_t k = 0; k < f.values.size(); k++) dst[k] = 0.0;
  for (auto &v : vals) { const int ix = index_of(v); if (ix >= 0) dst[ix] = 1.0; }
}

// FieldMatcher.tokenize for the `whitespace` analyzer (Lucene WhitespaceTokenizer: split on whitespace only);
// NgramMatcher.tokenize / TermMatcher.tokenize, then FieldMatcher.unique (sort + dedupe)
std::vector<std::string> match_tokens(const ReqFeature &f, const std::string &text) {
  if (f.language != "whitespace")
    fail(MR_ERR_INVALID_ARG, "feature %s: language '%s' is a Lucene analyzer; the request must carry tokens.%s", f.name.c_str(),
         f.language.c_str(), f.name.c_str());
  std::vector<std::string> terms;
  size_t i = 0;
  auto is_ws = [](unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f); };
  while (i < text.size()) {
    while (i < text.size() && is_ws((unsigned char)text[i])) i++;
    size_t j = i;
    while (j < text.size() && !is_ws((unsigned char)text[j])) j++;
    if (j > i) terms.push_back(text.substr(i, j - i));
    i = j;
  }
  std::vector<std::string> out;
  if (f.mtype == "ngram") {
    for (auto &t : terms) {
      // String.substring works on UTF-16 units; for BMP text a code point is one unit: walk UTF-8 code points
      std::vector<size_t> cp;
      for (size_t k = 0; k < t.size(); k++) if (((unsigned char)t[k] & 0xC0) != 0x80) cp.push_back(k);
      cp.push_back(t.size());
      const int L = (int)cp.size() - 1;
      for (int j = 0; j + f.n <= L; j++) out.push_back(t.substr(cp[j], cp[j + f.n] - cp[j]));
    }
  } else {
    out = terms;
  }
  std::sort(out.begin(), out.end());
  out.erase(std::unique(out.begin(), out.end()), out.end());
  return out;
}

double map_datetime(const std::string &parse, const ZonedTime &z) {  // LocalDateTimeFeature.scala:44-80
  if (parse == "time_of_day") return (double)(z.hour * 3600 + z.minute * 60 + z.second) / 3600.0;
  if (parse == "day_of_week") return (double)z.iso_weekday();
  if (parse == "month_of_year") return (double)z.month;
  if (parse == "year") return (double)z.year;
  return (double)floor_div(z.epoch_ms, 1000);  // "second"
}

}  // namespace

struct RequestPlan {
  std::vector<ReqFeature> features;
};

std::shared_ptr<const RequestPlan> make_request_plan(const Schema &S) {
  auto p = std::make_shared<RequestPlan>();
  p->features = build_features(S);
  return p;
}

// Element boundaries of a top-level JSON array: [begin, end) of every element, found by nesting depth outside strings.
// Only a structural scan — no token is validated; returns false on anything it does not recognise (the caller then
// walks the body sequentially, which also produces the error the reference's parser would).
static bool split_top_level_array(const uint8_t *p, const uint8_t *e, std::vector<std::pair<const uint8_t *, const uint8_t *>> &out,
                                  const uint8_t *&after) {
  auto ws = [&] { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; };
  ws();
  if (p >= e || *p != '[') return false;
  p++;
  ws();
  if (p < e && *p == ']') { after = p + 1; return true; }
  for (;;) {
    ws();
    const uint8_t *b = p;
    int depth = 0;
    for (;;) {
      if (p >= e) return false;
      const uint8_t ch = *p;
      if (ch == '"') {
        p++;
        while (p < e && *p != '"') p += (*p == '\\' && p + 1 < e) ? 2 : 1;
        if (p >= e) return false;
        p++;
      } else if (ch == '{' || ch == '[') { depth++; p++; }
      else if (ch == '}' || ch == ']') {
        if (depth == 0) { if (ch == '}') return false; break; }  // the array's own ']'
        depth--; p++;
      } else if (ch == ',' && depth == 0) break;
      else p++;
    }
    const uint8_t *t = p;
    while (t > b && (t[-1] == ' ' || t[-1] == '\n' || t[-1] == '\t' || t[-1] == '\r')) t--;
    if (t == b) return false;  // empty element: let the sequential parser say what is wrong
    out.emplace_back(b, t);
    if (*p == ',') { p++; continue; }
    after = p + 1;  // past ']'
    return true;
  }
}

// `work` on nt threads (the caller's included).  A thread that cannot be started is simply not there: `work` pulls from
// a shared counter, so the threads that do run finish the job.
template <class F> static void run_on_threads(int nt, F &&work) {
  std::vector<std::thread> pool;
  try {
    for (int t = 1; t < nt; t++) pool.emplace_back(work);
  } catch (const std::system_error &) {
  }
  work();
  for (auto &th : pool) th.join();
}

void decode_requests(const Schema &S, const RequestPlan &rp, const char *json, size_t len, PackedRequests &P) {
  Cur c{(const uint8_t *)json, (const uint8_t *)json + len};
  std::vector<Event> events;
  bool parsed = false;
  // A large batch (an array of events) is parsed by several threads: the elements are independent once their boundaries
  // are known.  Anything unusual — a scan that does not end cleanly, an element that fails or is not consumed whole —
  // drops back to the sequential walk below, so accepted bodies decode to the same batch and rejected ones get the same
  // error either way.
  const char *env_threads = getenv("MR_DECODE_THREADS");  // 1 = always sequential
  const int max_threads = env_threads ? std::max(1, atoi(env_threads)) : (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if (max_threads > 1 && len >= (size_t)(256 << 10) && c.peek() == '[') {
    std::vector<std::pair<const uint8_t *, const uint8_t *>> el;
    const uint8_t *after = nullptr;
    if (split_top_level_array(c.p, c.e, el, after) && el.size() >= 64) {
      const size_t n = el.size();
      const int nt = (int)std::min<size_t>((size_t)max_threads, n / 32);
      events.resize(n);
      std::atomic<size_t> next{0};
      std::atomic<bool> failed{false};
      auto work = [&] {
        for (;;) {
          const size_t i0 = next.fetch_add(16);
          if (i0 >= n || failed.load(std::memory_order_relaxed)) return;
          for (size_t i = i0; i < std::min(n, i0 + 16); i++) {
            try {
              Cur ci{el[i].first, el[i].second};
              events[i] = decode_event(ci);
              if (ci.peek() != '\0') { failed = true; return; }
            } catch (...) {
              failed = true;
              return;
            }
          }
        }
      };
      run_on_threads(nt, work);
      if (!failed) { c.p = after; parsed = true; }
      else events.clear();
    }
  }
  if (!parsed) {
    if (c.peek() == '[') c.array([&] { events.push_back(decode_event(c)); });
    else events.push_back(decode_event(c));
  }
  if (c.peek() != '\0' || c.p != c.e) fail(MR_ERR_PARSE, "json: trailing characters after the request");  // circe's parser fails on them
  const std::vector<ReqFeature> &plan = rp.features;
  const int R = (int)events.size();
  const size_t nrf = S.in_req_f64.size(), nru = S.in_req_u64.size(), nrv = S.in_req_vec.size(), nif = S.in_item_f64.size(), ntk = S.in_req_tok.size();
  const double kNaN = std::nan("");
  P = PackedRequests{};
  P.n_requests = R;
  P.offsets.assign(R + 1, 0);
  for (int r = 0; r < R; r++) P.offsets[r + 1] = P.offsets[r] + (int32_t)events[r].items.size();
  const int N = P.total_items = P.offsets[R];
  P.ids.assign(std::max(N, 1), 0);
  P.users.assign(std::max(R, 1), 0);
  P.sessions.assign(std::max(R, 1), 0);
  P.req_f64.assign((size_t)std::max(R, 1) * std::max<size_t>(nrf, 1), kNaN);
  P.req_u64.assign((size_t)std::max(R, 1) * std::max<size_t>(nru, 1), 0);
  P.req_vec.assign((size_t)std::max(R, 1) * std::max(S.vec_stride, 1), 0.f);
  P.req_vp.assign((size_t)std::max(R, 1) * std::max<size_t>(nrv, 1), 0);
  // the per-item override matrix is N x nif doubles (48 MB for 200 000 items x 30 features): only built when some item of
  // the batch carries fields at all — otherwise mr_rank_batch.item_f64 is NULL anyway
  bool any_item_fields = false;
  for (int r = 0; r < R && !any_item_fields; r++)
    for (auto &it : events[r].items)
      if (!it.fields.empty()) { any_item_fields = true; break; }
  if (any_item_fields || nif == 0) P.item_f64.assign((size_t)std::max(N, 1) * std::max<size_t>(nif, 1), kNaN);
  P.tok_off.assign(1, 0);
  P.item_ids.resize(N);
  P.request_ids.resize(R);
  P.timestamps.resize(R);
  // per request and token slot: the (hash, weight) list; concatenated in request order once every request is packed
  using TokList = std::vector<std::pair<uint64_t, double>>;
  std::vector<std::vector<TokList>> req_tok(ntk ? (size_t)R : 0);
  std::atomic<bool> any_item_f64{false};
  // One request's slice of every output array (disjoint from every other request's, so requests pack in parallel).
  auto pack_request = [&](int r) {
    Event &q = events[r];
    std::vector<TokList> tok((size_t)std::max<size_t>(ntk, 1));
    P.request_ids[r] = q.id;
    P.timestamps[r] = q.ts;
    if (q.has_user) P.users[r] = hash64(q.user.data(), q.user.size());
    if (q.has_session) P.sessions[r] = hash64(q.session.data(), q.session.size());
    auto last = [&](const std::string &n) -> const FieldVal * {  // RankingEvent.fieldsMap: last duplicate wins
      const FieldVal *v = nullptr;
      for (auto &f : q.fields) if (f.name == n) v = &f.v;
      return v;
    };
    auto first = [&](const std::string &n) -> const FieldVal * {
      for (auto &f : q.fields) if (f.name == n) return &f.v;
      return nullptr;
    };
    for (auto &f : plan) {
      if ((f.type == "number" || f.type == "word_count") && f.scope == "ranking" && f.s_f64 >= 0) {
        const FieldVal *v = last(f.src_field);
        double *dst = &P.req_f64[(size_t)r * nrf + f.s_f64];
        if (f.type == "number" && v && v->is_num()) *dst = v->d;
        if (f.type == "word_count" && v && v->is_str()) *dst = (double)token_count(v->s.data(), v->s.size());
      } else if (f.type == "string" && f.src_event == "ranking" && f.s_f64 >= 0) {
        const FieldVal *v = first(f.src_field);
        std::vector<std::string> vals;
        if (v && v->is_str()) vals.push_back(v->s);
        else if (v && v->is_strlist()) vals = v->sl;
        encode_string(f, vals, &P.req_f64[(size_t)r * nrf + f.s_f64]);
      } else if (f.type == "rate" && !f.rate_field.empty() && f.s_u64 >= 0) {
        const FieldVal *v = last(f.rate_field);
        if (v && v->is_str()) P.req_u64[(size_t)r * nru + f.s_u64] = hash64(v->s.data(), v->s.size());
      } else if (f.type == "item_age" && f.s_u64 >= 0) {
        P.req_u64[(size_t)r * nru + f.s_u64] = (uint64_t)q.ts;
      } else if (f.type == "local_time" && f.s_f64 >= 0) {
        ZonedTime z;
        bool ok = false;
        if (f.lt_field == "timestamp") { z = utc_from_epoch_seconds(floor_div(q.ts, 1000)); ok = true; }
        else if (const FieldVal *v = last(f.lt_field)) ok = v->is_str() && parse_iso_zoned(v->s, z);
        if (ok) P.req_f64[(size_t)r * nrf + f.s_f64] = map_datetime(f.lt_parse, z);
      } else if (f.type == "field_match" && f.s_tok >= 0) {
        const FieldVal *qf = last(f.rank_field);
        if (qf && qf->is_str()) {  // only a StringField is tokenized (FieldMatchFeature.scala:62-68)
          std::vector<std::string> toks;
          const JValue *given = q.tokens.get(f.name.c_str());
          if (given && given->kind == JValue::Arr) for (auto &x : given->arr) toks.push_back(x.str);
          else toks = match_tokens(f, qf->s);
          for (auto &tk : toks) {
            double w = 0.0;
            if (f.mtype == "bm25") {  // BM25Matcher.score's termIDF (BM25Matcher.scala:26-27)
              auto tf = f.termfreq.find(tk);
              const double gtf = tf == f.termfreq.end() ? 0.0 : tf->second;
              w = std::log(1.0 + (f.docs - gtf + 0.5) / (gtf + 0.5));
            }
            tok[f.s_tok].push_back({hash64(tk.data(), tk.size()), w});
          }
        }
      } else if (f.type == "field_match" && f.s_vec >= 0) {
        const JValue *emb = q.embeddings.get(f.name.c_str());
        const FieldVal *qf = last(f.rank_field);
        if (emb && emb->kind == JValue::Arr && qf && (qf->is_str() || qf->is_strlist())) {
          if ((int)emb->arr.size() != f.vec_dim)
            fail(MR_ERR_INVALID_ARG, "feature %s: query embedding has %zu values, the schema says %d", f.name.c_str(), emb->arr.size(), f.vec_dim);
          for (int k = 0; k < f.vec_dim; k++) P.req_vec[(size_t)r * S.vec_stride + f.vec_off + k] = (float)emb->arr[k].as_double();
          P.req_vp[(size_t)r * nrv + f.s_vec] = 1;
        }
      }
    }
    if (ntk) req_tok[r] = std::move(tok);
    for (size_t j = 0; j < q.items.size(); j++) {
      Item &it = q.items[j];
      const size_t i = (size_t)P.offsets[r] + j;
      P.ids[i] = hash64(it.id.data(), it.id.size());
      if (it.fields.empty()) { P.item_ids[i] = std::move(it.id); continue; }  // nothing to override with
      for (auto &f : plan) {
        if (f.item_kind == 0) continue;
        double *dst = &P.item_f64[i * nif + f.s_item];
        if (f.item_kind == 1) {
          for (auto &fl : it.fields)
            if (fl.name == "relevancy") { if (fl.v.is_num()) { *dst = fl.v.d; any_item_f64 = true; } break; }  // the FIRST one decides
        } else if (f.item_kind == 2) {
          for (auto &fl : it.fields)
            if (fl.name == f.src_field && fl.v.is_num()) { *dst = fl.v.d; any_item_f64 = true; break; }
        } else if (f.item_kind == 3) {
          for (auto &fl : it.fields)
            if (fl.name == f.src_field && (fl.v.is_str() || fl.v.is_strlist())) {
              std::vector<std::string> vals = fl.v.is_str() ? std::vector<std::string>{fl.v.s} : fl.v.sl;
              encode_string(f, vals, dst);
              any_item_f64 = true;
              break;
            }
        }
      }
      P.item_ids[i] = std::move(it.id);
    }
  };
  const int pack_threads = (max_threads > 1 && R >= 64 && N >= 4096) ? (int)std::min<size_t>((size_t)max_threads, (size_t)R / 16) : 1;
  if (pack_threads <= 1) {
    for (int r = 0; r < R; r++) pack_request(r);
  } else {
    // a request the sequential walk would have failed on FIRST decides the error: keep the failure of the lowest index
    std::atomic<int> next{0};
    std::mutex err_mu;
    int err_at = R;
    std::exception_ptr err;
    auto work = [&] {
      for (;;) {
        const int r0 = next.fetch_add(8);
        if (r0 >= R) return;
        for (int r = r0; r < std::min(R, r0 + 8); r++) {
          try {
            pack_request(r);
          } catch (...) {
            std::lock_guard<std::mutex> g(err_mu);
            if (r < err_at) { err_at = r; err = std::current_exception(); }
          }
        }
      }
    };
    run_on_threads(pack_threads, work);
    if (err) std::rethrow_exception(err);
  }
  for (int r = 0; r < R && ntk; r++)
    for (size_t sl = 0; sl < ntk; sl++) {
      for (auto &hw : req_tok[r][sl]) { P.tok_hash.push_back(hw.first); P.tok_w.push_back(hw.second); }
      P.tok_off.push_back((int32_t)P.tok_hash.size());
    }
  if (ntk) { P.tok_hash.push_back(0); P.tok_w.push_back(0.0); }
  P.has_item_f64 = nif > 0 && any_item_f64.load();
  mr_rank_batch &b = P.batch;
  b.n_requests = R;
  b.item_offsets = P.offsets.data();
  b.item_ids = P.ids.data();
  b.user_ids = P.users.data();
  b.session_ids = P.sessions.data();
  b.req_f64 = P.req_f64.data();
  b.req_u64 = P.req_u64.data();
  b.req_vec = P.req_vec.data();
  b.req_vec_present = P.req_vp.data();
  b.item_f64 = P.has_item_f64 ? P.item_f64.data() : nullptr;
  b.req_tok_offsets = ntk ? P.tok_off.data() : nullptr;
  b.req_tok_hashes = ntk ? P.tok_hash.data() : nullptr;
  b.req_tok_weights = ntk ? P.tok_w.data() : nullptr;
}

}  // namespace mr
