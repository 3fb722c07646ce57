// Parses the Metarank `features:` section (as JSON) + a model's feature list into the
// extractor plan.  See schema.h.  JSON shapes follow the reference's circe decoders:
//   S/model/FeatureSchema.scala:41-80 (polymorphic on "type"), S/model/ScopeType.scala
//   ("item" | "user" | "session" | "global" | "ranking" | "item.<f>" | "ranking.<f>"),
//   S/model/FieldName.scala ("item.x" | "metadata.x" | "ranking.x" | ...), and each
//   extractor's schema decoder.
#include "schema.h"

#include <algorithm>
#include <cstring>

#include "json.h"

namespace mr {

uint64_t hash64(const void *bytes, size_t len) {
  // FNV-1a 64 + a splitmix-style finaliser; 0 is reserved for "absent"
  const uint8_t *p = (const uint8_t *)bytes;
  uint64_t h = 0xCBF29CE484222325ull;
  for (size_t i = 0; i < len; i++) {
    h ^= p[i];
    h *= 0x100000001B3ull;
  }
  h ^= h >> 30; h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 27; h *= 0x94D049BB133111EBull;
  h ^= h >> 31;
  return h ? h : 1;
}

// java: "\\s+".r.split(s).length.  java.util.regex.Pattern.split: a zero-length input
// yields [""] (length 1); a match at index 0 yields a leading empty string (for a non-zero
// -width match); trailing empty strings are removed.  \s = [ \t\n\x0B\f\r].
int32_t token_count(const char *s, size_t len) {
  auto ws = [](unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == 0x0B || c == '\f' || c == '\r'; };
  if (len == 0) return 1;
  int tokens = 0;        // substrings emitted so far (including empty leading one)
  int trailing_empty = 0;
  size_t i = 0, start = 0;
  bool matched_any = false;
  while (i < len) {
    if (ws((unsigned char)s[i])) {
      size_t j = i;
      while (j < len && ws((unsigned char)s[j])) j++;
      // substring [start, i)
      matched_any = true;
      if (i == start) { tokens++; trailing_empty++; }  // empty token (only possible at index 0)
      else { tokens++; trailing_empty = 0; }
      start = j;
      i = j;
    } else {
      i++;
    }
  }
  if (!matched_any) return 1;
  // remainder [start, len)
  if (start < len) { tokens++; trailing_empty = 0; }
  else { tokens++; trailing_empty++; }
  tokens -= trailing_empty;
  return tokens < 0 ? 0 : tokens;
}

int64_t parse_duration_ms(const std::string &s, const std::string &feature) {
  // ai.metarank.util.DurationJson: "([0-9]+)([smhd]{1})"
  size_t i = 0;
  while (i < s.size() && s[i] >= '0' && s[i] <= '9') i++;
  if (i == 0 || i + 1 != s.size()) fail(MR_ERR_PARSE, "feature %s: duration is in wrong format: %s", feature.c_str(), s.c_str());
  const int64_t n = strtoll(s.c_str(), nullptr, 10);
  switch (s[i]) {
    case 's': return n * 1000;
    case 'm': return n * 60000;
    case 'h': return n * 3600000;
    case 'd': return n * 86400000;
  }
  fail(MR_ERR_PARSE, "feature %s: duration is in wrong format: %s", feature.c_str(), s.c_str());
}

namespace {

struct ScopeSpec { int scope; std::string field; };

ScopeSpec parse_scope(const std::string &s, const std::string &feature) {
  if (s == "global") return {SC_GLOBAL, ""};
  if (s == "item") return {SC_ITEM, ""};
  if (s == "user") return {SC_USER, ""};
  if (s == "session") return {SC_SESSION, ""};
  if (s == "ranking") return {SC_RANKING, ""};
  if (s.rfind("item.", 0) == 0 && s.size() > 5) return {SC_FIELD, s.substr(5)};
  if (s.rfind("ranking.", 0) == 0 && s.size() > 8) return {SC_IRF, s.substr(8)};
  fail(MR_ERR_PARSE, "feature %s: scope type %s not supported", feature.c_str(), s.c_str());
}

struct FieldSpec { std::string event, field; };
FieldSpec parse_field(const std::string &s, const std::string &feature) {
  size_t dot = s.find('.');
  if (dot == std::string::npos || dot == 0 || dot + 1 >= s.size())
    fail(MR_ERR_PARSE, "feature %s: cannot decode source field '%s': it should have a format of <type>.<name>", feature.c_str(), s.c_str());
  std::string ev = s.substr(0, dot);
  if (ev == "metadata") ev = "item";
  return {ev, s.substr(dot + 1)};
}

std::string str_of(const JValue &o, const char *key, const std::string &feature, const char *dflt = nullptr) {
  const JValue *v = o.get(key);
  if (!v || v->kind == JValue::Null) {
    if (dflt) return dflt;
    fail(MR_ERR_PARSE, "feature %s: missing field '%s'", feature.c_str(), key);
  }
  if (v->kind != JValue::Str) fail(MR_ERR_PARSE, "feature %s: field '%s' must be a string", feature.c_str(), key);
  return v->str;
}

}  // namespace

Schema parse_schema_json(const char *json, size_t len) {
  JValue doc = JsonParser((const uint8_t *)json, len).parse();
  const JValue &jf = doc.at("features");
  const JValue &jm = doc.at("model_features");
  if (jf.kind != JValue::Arr || jm.kind != JValue::Arr) fail(MR_ERR_PARSE, "schema: 'features' and 'model_features' must be arrays");
  Schema S;
  S.source_json.assign(json, len);
  for (auto &v : jm.arr) {
    if (v.kind != JValue::Str) fail(MR_ERR_PARSE, "schema: model_features must be strings");
    S.model_features.push_back(v.str);
  }

  auto add_slot = [&](int feature, const std::string &name, int table, int kind, int n_words, int p) -> int {
    if (S.slot_by_name.count(name)) fail(MR_ERR_PARSE, "duplicate feature state name '%s'", name.c_str());
    Slot sl;
    sl.name = name; sl.table = table; sl.kind = kind; sl.n_words = n_words; sl.p = p; sl.feature = feature;
    int id = (int)S.slots.size();
    S.slots.push_back(sl);
    S.slot_by_name[name] = id;
    S.features[feature].slots.push_back(id);
    return id;
  };

  // ---- pass 1: decode every configured feature
  std::vector<const JValue *> jdefs;
  for (auto &o : jf.arr) {
    if (o.kind != JValue::Obj) fail(MR_ERR_PARSE, "schema: feature entries must be objects");
    FeatureDef fd;
    fd.name = str_of(o, "name", "?");
    fd.type = str_of(o, "type", fd.name);
    for (auto &other : S.features)
      if (other.name == fd.name) fail(MR_ERR_PARSE, "non-unique feature '%s' is defined more than once", fd.name.c_str());
    S.features.push_back(fd);
    jdefs.push_back(&o);
  }

  // ---- pass 2: model features in model order -> columns, slots, plan
  int col = 0;
  for (const std::string &mf : S.model_features) {
    int fi = -1;
    for (size_t i = 0; i < S.features.size(); i++)
      if (S.features[i].name == mf) fi = (int)i;
    if (fi < 0) continue;  // silently dropped from the descriptor (S/FeatureMapping.scala:66-72)
    FeatureDef &fd = S.features[fi];
    if (fd.col >= 0) fail(MR_ERR_PARSE, "model lists feature '%s' twice", mf.c_str());
    const JValue &o = *jdefs[fi];
    const std::string &t = fd.type, &n = fd.name;
    fd.col = col;
    DFeature d;
    memset(&d, 0, sizeof d);
    d.col = col; d.dim = 1; d.in0 = d.in1 = -1;
    for (int k = 0; k < 4; k++) d.w[k] = d.b[k] = -1;
    auto table_of = [&](int scope) {
      if (scope == SC_RANKING) fail(MR_ERR_UNSUPPORTED, "feature %s: ranking-scoped state is not read on the /rank path", n.c_str());
      return scope;
    };
    auto bind = [&](int k, int slot) { d.w[k] = slot; d.b[k] = slot; };  // resolved to words/bits below

    if (t == "number" || t == "word_count") {
      ScopeSpec sc = parse_scope(str_of(o, "scope", n), n);
      std::string src = o.get("source") ? str_of(o, "source", n) : str_of(o, "field", n);
      parse_field(src, n);
      if (sc.scope == SC_FIELD || sc.scope == SC_IRF) fail(MR_ERR_PARSE, "feature %s: scope %s not supported for %s", n.c_str(), str_of(o, "scope", n).c_str(), t.c_str());
      fd.scope = sc.scope;
      if (sc.scope == SC_RANKING) {
        d.kind = FK_CONST_REQ;
        d.in0 = (int)S.in_req_f64.size();
        S.in_req_f64.push_back(n);
      } else {
        d.kind = FK_NUMBER;
        d.scope = sc.scope;
        bind(0, add_slot(fi, n, table_of(sc.scope), SK_F64, 1, 0));
        if (t == "number") {  // per-item field override (S/feature/NumberFeature.scala:84-93)
          d.in0 = (int)S.in_item_f64.size();
          S.in_item_f64.push_back(n);
        }
      }
    } else if (t == "boolean") {
      ScopeSpec sc = parse_scope(str_of(o, "scope", n), n);
      parse_field(o.get("field") ? str_of(o, "field", n) : str_of(o, "source", n), n);
      if (sc.scope == SC_FIELD || sc.scope == SC_IRF || sc.scope == SC_RANKING)
        fail(MR_ERR_UNSUPPORTED, "feature %s: scope not supported for boolean features", n.c_str());
      d.kind = FK_NUMBER; d.scope = sc.scope; fd.scope = sc.scope;  // reads 1.0 / 0.0 / NaN, no per-item override
      bind(0, add_slot(fi, n, sc.scope, SK_BOOL, 1, 0));
    } else if (t == "vector") {
      ScopeSpec sc = parse_scope(str_of(o, "scope", n), n);
      parse_field(str_of(o, "source", n), n);
      if (sc.scope == SC_FIELD || sc.scope == SC_IRF || sc.scope == SC_RANKING)
        fail(MR_ERR_UNSUPPORTED, "feature %s: scope not supported for vector features", n.c_str());
      int dimv = 0;
      const JValue *red = o.get("reduce");
      if (!red || red->kind == JValue::Null) dimv = 4;  // default reducers: min, max, size, avg
      else if (red->kind == JValue::Arr) {
        for (auto &r : red->arr) {
          if (r.kind != JValue::Str) fail(MR_ERR_PARSE, "feature %s: reducers must be strings", n.c_str());
          static const char *kOne[] = {"first", "last", "min", "max", "avg", "random", "sum", "size", "euclidean_distance"};
          bool one = false;
          for (const char *k : kOne) one |= r.str == k;
          if (one) dimv += 1;
          else if (r.str.rfind("vector", 0) == 0 && r.str.size() > 6 && r.str.find_first_not_of("0123456789", 6) == std::string::npos)
            dimv += atoi(r.str.c_str() + 6);
          else fail(MR_ERR_PARSE, "feature %s: reducer %s is not supported", n.c_str(), r.str.c_str());
        }
      } else fail(MR_ERR_PARSE, "feature %s: 'reduce' must be a list", n.c_str());
      if (dimv <= 0 || dimv > 4096) fail(MR_ERR_PARSE, "feature %s: bad vector dimension %d", n.c_str(), dimv);
      d.kind = FK_VECTOR; d.scope = sc.scope; d.dim = dimv; fd.scope = sc.scope;
      bind(0, add_slot(fi, n, sc.scope, SK_F64VEC, dimv, dimv));
    } else if (t == "item_age") {
      FieldSpec fs = parse_field(str_of(o, "source", n), n);
      if (fs.event != "item") fail(MR_ERR_PARSE, "feature %s: can only work with fields from metadata events", n.c_str());
      d.kind = FK_ITEM_AGE; d.scope = SC_ITEM;
      bind(0, add_slot(fi, n, SC_ITEM, SK_F64, 1, 0));
      d.in1 = (int)S.in_req_u64.size();  // the RankingEvent's timestamp (epoch millis)
      S.in_req_u64.push_back(n);
    } else if (t == "local_time") {
      FieldSpec fs = parse_field(str_of(o, "source", n), n);
      if (fs.event != "ranking") fail(MR_ERR_PARSE, "feature %s: can only work with ranking event fields", n.c_str());
      std::string pm = str_of(o, "parse", n);
      if (pm != "time_of_day" && pm != "day_of_week" && pm != "month_of_year" && pm != "year" && pm != "second")
        fail(MR_ERR_PARSE, "feature %s: parsing method %s is not supported", n.c_str(), pm.c_str());
      d.kind = FK_CONST_REQ;  // the caller evaluates the DateTimeMapper on the request's timestamp / field
      d.in0 = (int)S.in_req_f64.size();
      S.in_req_f64.push_back(n);
    } else if (t == "string") {
      ScopeSpec sc = parse_scope(str_of(o, "scope", n), n);
      std::string src = o.get("source") ? str_of(o, "source", n) : str_of(o, "field", n);
      FieldSpec fs = parse_field(src, n);
      std::string enc = str_of(o, "encode", n, "onehot");
      if (enc != "index" && enc != "onehot") fail(MR_ERR_PARSE, "feature %s: string encoding method %s is not supported", n.c_str(), enc.c_str());
      const JValue &vals = o.at("values");
      if (vals.kind != JValue::Arr || vals.arr.empty()) fail(MR_ERR_PARSE, "feature %s: 'values' must be a non-empty list", n.c_str());
      for (auto &v : vals.arr) {
        if (v.kind != JValue::Str) fail(MR_ERR_PARSE, "feature %s: 'values' must be strings", n.c_str());
        fd.cat_values.push_back(v.str);
        fd.cat_hashes.push_back(hash64(v.str.data(), v.str.size()));
      }
      const bool onehot = enc == "onehot";
      d.dim = onehot ? (int)fd.cat_values.size() : 1;
      if (onehot && d.dim > 64) fail(MR_ERR_UNSUPPORTED, "feature %s: onehot encoding of more than 64 values is not supported", n.c_str());
      d.aux0 = onehot;
      fd.scope = sc.scope;
      if (fs.event == "ranking") {
        d.kind = FK_CONST_REQ;
        d.in0 = (int)S.in_req_f64.size();
        for (int k = 0; k < d.dim; k++) S.in_req_f64.push_back(n);
      } else {
        if (sc.scope == SC_FIELD || sc.scope == SC_IRF || sc.scope == SC_RANKING)
          fail(MR_ERR_UNSUPPORTED, "feature %s: scope not supported for string features", n.c_str());
        d.kind = onehot ? FK_ONEHOT : FK_CATEGORY;
        d.scope = sc.scope;
        bind(0, add_slot(fi, n, table_of(sc.scope), SK_CAT, 1, 0));
        d.in0 = (int)S.in_item_f64.size();  // per-item override, already encoded by the caller
        for (int k = 0; k < d.dim; k++) S.in_item_f64.push_back(n);
      }
    } else if (t == "referer") {
      // RefererFeature (S/feature/RefererFeature.scala:39-110).  The referer URL is parsed when the event is WRITTEN
      // (snowplow's referers.json, on the JVM: writeField :70-90 stores SString(medium) under the user or the session);
      // the rank path only reads that string back and maps it through the fixed table :47-54 — unknown 0, search 1,
      // internal 2, social 3, email 4, paid 5; absent / anything else -> CategoryValue("unknown", 0).  That is the
      // string-index encoder (index + 1, 0 = nil) over the five named mediums, with no request-side override.
      ScopeSpec sc = parse_scope(str_of(o, "scope", n), n);
      (void)parse_field(str_of(o, "source", n), n);  // required by RefererSchema; read by the writer only
      if (sc.scope != SC_USER && sc.scope != SC_SESSION)
        fail(MR_ERR_UNSUPPORTED, "feature %s: referer is kept per user or per session (RefererFeature.writeField), scope %s has no key",
             n.c_str(), str_of(o, "scope", n).c_str());
      for (const char *m : {"search", "internal", "social", "email", "paid"}) {
        fd.cat_values.push_back(m);
        fd.cat_hashes.push_back(hash64(m, strlen(m)));
      }
      d.dim = 1;
      d.aux0 = 0;
      fd.scope = sc.scope;
      d.kind = FK_CATEGORY;
      d.scope = sc.scope;
      bind(0, add_slot(fi, n, table_of(sc.scope), SK_CAT, 1, 0));
      d.in0 = -1;
    } else if (t == "interaction_count") {
      ScopeSpec sc = parse_scope(str_of(o, "scope", n), n);
      if (sc.scope == SC_FIELD || sc.scope == SC_IRF || sc.scope == SC_RANKING)
        fail(MR_ERR_UNSUPPORTED, "feature %s: scope not supported for interaction_count", n.c_str());
      d.kind = FK_COUNT; d.scope = sc.scope; fd.scope = sc.scope;
      bind(0, add_slot(fi, n, sc.scope, SK_COUNTER, 1, 0));
    } else if (t == "window_count") {
      ScopeSpec sc = parse_scope(str_of(o, "scope", n), n);
      if (sc.scope == SC_FIELD || sc.scope == SC_IRF || sc.scope == SC_RANKING)
        fail(MR_ERR_UNSUPPORTED, "feature %s: scope not supported for window_count", n.c_str());
      const JValue &per = o.at("periods");
      if (per.kind != JValue::Arr || per.arr.empty()) fail(MR_ERR_PARSE, "feature %s: 'periods' must be a non-empty list", n.c_str());
      d.kind = FK_WINDOW; d.scope = sc.scope; fd.scope = sc.scope;
      d.dim = (int)per.arr.size();
      {
        const int sl = add_slot(fi, n, sc.scope, SK_PCOUNTER, d.dim, d.dim);
        bind(0, sl);
        S.slots[sl].period_ms = parse_duration_ms(str_of(o, "bucket", n), n);
        for (auto &pv : per.arr) S.slots[sl].ranges.push_back((int)pv.as_int());
      }
    } else if (t == "rate") {
      std::string top = str_of(o, "top", n), bottom = str_of(o, "bottom", n);
      ScopeSpec sc = o.get("scope") && o.get("scope")->kind == JValue::Str ? parse_scope(o.get("scope")->str, n) : ScopeSpec{SC_ITEM, ""};
      if (sc.scope != SC_ITEM && sc.scope != SC_FIELD && sc.scope != SC_IRF)
        fail(MR_ERR_PARSE, "scope %s is not supported for rate feature %s", str_of(o, "scope", n).c_str(), n.c_str());
      const JValue &per = o.at("periods");
      if (per.kind != JValue::Arr || per.arr.empty()) fail(MR_ERR_PARSE, "feature %s: 'periods' must be a non-empty list", n.c_str());
      const int P = (int)per.arr.size();
      d.kind = FK_RATE; d.scope = sc.scope; d.dim = P; fd.scope = sc.scope; fd.scope_field = sc.field;
      bind(0, add_slot(fi, n + "_" + top, sc.scope, SK_PCOUNTER, P, P));
      bind(1, add_slot(fi, n + "_" + bottom, sc.scope, SK_PCOUNTER, P, P));
      bind(2, add_slot(fi, n + "_" + top + "_norm", SC_GLOBAL, SK_PCOUNTER, P, P));
      bind(3, add_slot(fi, n + "_" + bottom + "_norm", SC_GLOBAL, SK_PCOUNTER, P, P));
      {
        const int64_t bucket = parse_duration_ms(str_of(o, "bucket", n), n);
        for (int k = 0; k < 4; k++) {
          Slot &sl = S.slots[d.w[k]];
          sl.period_ms = bucket;
          for (auto &pv : per.arr) sl.ranges.push_back((int)pv.as_int());
        }
      }
      const JValue *norm = o.get("normalize");
      if (norm && norm->kind == JValue::Obj) {
        d.aux0 = 1;
        d.dparam = norm->at("weight").as_double();
      }
      if (sc.scope == SC_FIELD) {
        d.aux1 = add_slot(fi, n + "_field", SC_ITEM, SK_STRID, 1, 0);  // resolved below
        d.uparam = hash64(sc.field.data(), sc.field.size());
      } else if (sc.scope == SC_IRF) {
        d.in1 = (int)S.in_req_u64.size();
        S.in_req_u64.push_back(n);
        d.uparam = hash64(sc.field.data(), sc.field.size());
      }
    } else if (t == "interacted_with") {
      ScopeSpec sc = parse_scope(str_of(o, "scope", n), n);
      if (sc.scope != SC_USER && sc.scope != SC_SESSION) fail(MR_ERR_PARSE, "feature %s: can only be scoped to user/session", n.c_str());
      std::vector<std::string> fields;
      const JValue &jfld = o.at("field");
      if (jfld.kind == JValue::Str) fields.push_back(jfld.str);
      else if (jfld.kind == JValue::Arr) {
        for (auto &v : jfld.arr) {
          if (v.kind != JValue::Str) fail(MR_ERR_PARSE, "feature %s: 'field' entries must be strings", n.c_str());
          fields.push_back(v.str);
        }
      } else fail(MR_ERR_PARSE, "feature %s: 'field' must be a string or a list of strings", n.c_str());
      for (auto &f : fields) {
        FieldSpec fs = parse_field(f, n);
        if (fs.event != "item") fail(MR_ERR_PARSE, "feature %s: can only be applied to item fields", n.c_str());
        f = fs.field;
      }
      // the output order is the iteration order of a Scala immutable Map built from the field
      // list: insertion order up to 4 entries, hash order beyond (SURVEY.md §7 "quirks")
      if (fields.empty() || fields.size() > 4)
        fail(MR_ERR_UNSUPPORTED, "feature %s: interacted_with supports 1..4 fields (Scala Map ordering beyond 4 is hash order)", n.c_str());
      for (size_t a = 0; a < fields.size(); a++)
        for (size_t b2 = a + 1; b2 < fields.size(); b2++)
          if (fields[a] == fields[b2]) fail(MR_ERR_PARSE, "feature %s: duplicate field %s", n.c_str(), fields[a].c_str());
      int vis = add_slot(fi, n + "_interactions", sc.scope, SK_BLIST, 1, 0);
      {
        const JValue *cnt = o.get("count"), *dur = o.get("duration");
        S.slots[vis].list_count = (cnt && cnt->kind == JValue::Num) ? (int)cnt->as_int() : 100;  // getOrElse(100)
        S.slots[vis].list_duration_ms = (dur && dur->kind == JValue::Str) ? parse_duration_ms(dur->str, n) : 86400000;  // 24.hours
      }
      fd.dim = (int)fields.size();
      fd.scope = sc.scope;
      S.needs_prepass = true;
      S.needs_visitor = true;
      for (size_t k = 0; k < fields.size(); k++) {
        DFeature e = d;
        e.kind = FK_INTERACTED; e.scope = sc.scope; e.col = col + (int)k; e.dim = 1;
        e.w[0] = e.b[0] = add_slot(fi, n + "_" + fields[k], SC_ITEM, SK_STRLIST, 1, 0);
        e.w[1] = e.b[1] = vis;
        e.aux0 = S.n_hist++;
        S.plan.push_back(e);
      }
      col += fd.dim;
      S.col_of[n] = {fd.col, fd.dim};
      continue;
    } else if (t == "relevancy") {
      d.kind = FK_RELEVANCY;
      d.in0 = (int)S.in_item_f64.size();
      S.in_item_f64.push_back(n);
    } else if (t == "position") {
      d.kind = FK_POSITION;
      d.dparam = (double)o.at("position").as_int();
    } else if (t == "diversity") {
      FieldSpec fs = parse_field(str_of(o, "source", n), n);
      if (fs.event != "item") fail(MR_ERR_PARSE, "diversity feature '%s' can only accept item fields, but got '%s'", n.c_str(), fs.event.c_str());
      d.kind = FK_DIVERSITY; d.scope = SC_ITEM;
      bind(0, add_slot(fi, n, SC_ITEM, SK_DIVERSITY, 2, 0));
      const JValue *top = o.get("top");
      d.aux0 = (top && top->kind == JValue::Num) ? (int)top->as_int() : 20;  // S/feature/DiversityFeature.scala:164
      d.aux1 = S.n_hist++;
      d.aux2 = S.n_reqagg++;
      S.needs_prepass = true;
    } else if (t == "field_match") {
      const JValue &method = o.at("method");
      std::string mt = str_of(method, "type", n);
      FieldSpec rf = parse_field(str_of(o, "rankingField", n), n), itf = parse_field(str_of(o, "itemField", n), n);
      if (rf.event != "ranking") fail(MR_ERR_PARSE, "feature %s: expected ranking field", n.c_str());
      if (itf.event != "item") fail(MR_ERR_PARSE, "feature %s: expected item field", n.c_str());
      if (mt == "ngram" || mt == "term" || mt == "bm25") {
        // FieldMatchFeature (S/feature/FieldMatchFeature.scala:30-93): the item's token set is stored at write
        // time as SStringList under "<name>_<itemField>"; the query's tokens arrive with the request
        // (MR_IN_REQ_TOKENS).  Tokenisation itself (Lucene analyzers) stays with the caller.
        d.kind = FK_TOKEN_MATCH; d.scope = SC_ITEM;
        int sl = add_slot(fi, n + "_" + itf.field, SC_ITEM, SK_STRLIST, 1, 1 /* sorted + unique */);
        bind(0, sl);
        d.in0 = (int)S.in_req_tok.size();
        S.in_req_tok.push_back(n);
        d.aux0 = mt == "bm25" ? 1 : 0;
        if (mt == "bm25") {
          // BM25Matcher needs TermFreqDic.avgdl (S/feature/matcher/BM25Matcher.scala:19-33); the per-token IDF
          // comes with the request as the token's weight
          if (!method.get("avgdl")) fail(MR_ERR_PARSE, "feature %s: bm25 needs method.avgdl (TermFreqDic.avgdl)", n.c_str());
          d.dparam = method.at("avgdl").as_double();
        }
      } else {
      if (mt != "bi-encoder") fail(MR_ERR_UNSUPPORTED, "feature %s: field_match method %s is not supported on the GPU path", n.c_str(), mt.c_str());
      std::string dist = str_of(o, "distance", n, "cos");
      if (dist != "cos" && dist != "Cos" && dist != "cosine" && dist != "Cosine")
        fail(MR_ERR_UNSUPPORTED, "feature %s: distance '%s' is not supported", n.c_str(), dist.c_str());
      std::string norm = str_of(o, "norm", n, "noop");
      int nm = norm == "noop" ? 0 : norm == "linear" ? 1 : norm == "position" ? 2 : -1;
      if (nm < 0) fail(MR_ERR_PARSE, "feature %s: normalizer %s is not supported", n.c_str(), norm.c_str());
      const int dimv = (int)method.at("dim").as_int();
      if (dimv <= 0) fail(MR_ERR_PARSE, "feature %s: bad embedding dim", n.c_str());
      d.kind = FK_COSINE; d.scope = SC_ITEM;
      int sl = add_slot(fi, n, SC_ITEM, SK_F64LIST, 0, dimv);
      bind(0, sl);
      S.slots[sl].side = (int)S.sides.size();
      S.sides.push_back({SC_ITEM, dimv, sl});
      d.aux0 = dimv; d.aux1 = nm;
      d.aux2 = 0;  // cosine scratch column, assigned below
      d.aux3 = S.n_reqagg++;
      d.in0 = (int)S.in_req_vec.size();
      S.in_req_vec.push_back({n, dimv, S.vec_stride});
      S.vec_stride += dimv;
      S.needs_cosine = true;
      if (nm != 0) S.needs_prepass = true;
      }
    } else {
      fail(MR_ERR_UNSUPPORTED, "feature type %s (feature %s) is not supported on the GPU /rank path", t.c_str(), n.c_str());
    }
    if (d.scope == SC_USER || d.scope == SC_SESSION) S.needs_visitor = true;
    fd.dim = d.dim;
    fd.kind = d.kind;
    if (t == "string" || t == "referer") fd.kind = d.aux0 ? FK_ONEHOT : FK_CATEGORY;  // encoder kind, also when request-scoped
    S.plan.push_back(d);
    S.col_of[n] = {fd.col, fd.dim};
    col += d.dim;
  }
  S.dim = col;

  // ---- row layouts: presence words first, then the slots' payload words
  for (int tb = 0; tb < SC_N_TABLES; tb++) {
    int n = 0;
    for (auto &sl : S.slots) if (sl.table == tb) sl.bit = n++;
    TableLayout &L = S.tables[tb];
    L.n_slots = n;
    L.presence_words = (n + 63) / 64;
    int w = L.presence_words;
    for (auto &sl : S.slots) if (sl.table == tb) { sl.word = w; w += sl.n_words; }
    L.row_words = std::max((w + 3) & ~3, 4);  // 32-byte sectors: a row never straddles an extra one
  }
  // resolve slot ids in the plan to (word, bit)
  int n_cos = 0;
  for (auto &d : S.plan) {
    for (int k = 0; k < 4; k++) {
      if (d.w[k] >= 0) {
        const Slot &sl = S.slots[d.w[k]];
        if (d.kind == FK_COSINE) d.uparam = (uint64_t)sl.side;
        d.w[k] = sl.word;
        d.b[k] = sl.bit;
      }
    }
    if (d.kind == FK_RATE && d.scope == SC_FIELD) {
      const Slot &sl = S.slots[d.aux1];
      d.aux1 = sl.word;
      d.aux2 = sl.bit;
    }
    if (d.kind == FK_COSINE) d.aux2 = n_cos++;
  }
  // fast columns: item-scoped entries whose every column is one row word (see FastCol)
  if (S.tables[SC_ITEM].row_words <= 128 && S.dim <= 65535) {
    for (auto &d : S.plan) {
      d.fast = 0;
      d.pad = 0;
      if (d.scope != SC_ITEM) continue;
      int conv = -1, missing = 0;
      switch (d.kind) {
        case FK_NUMBER: conv = 0; break;
        case FK_VECTOR: conv = 0; break;
        case FK_CATEGORY: conv = 2; missing = 1; break;
        case FK_COUNT: conv = 1; missing = 1; break;
        case FK_WINDOW: conv = 1; break;
        default: break;
      }
      if (conv < 0) continue;
      d.fast = 1;
      for (int k = 0; k < d.dim; k++) {
        FastCol c{};
        c.word = (uint16_t)(d.w[0] + k);
        c.bit = (uint16_t)d.b[0];
        c.conv = (uint8_t)conv;
        c.missing = (uint8_t)missing;
        c.override_slot = (int16_t)((d.kind == FK_NUMBER || d.kind == FK_CATEGORY) && k == 0 ? d.in0 : -1);
        c.col = (uint16_t)(d.col + k);
        S.fast_has_override |= c.override_slot >= 0;
        S.fast_cols.push_back(c);
      }
    }
  } else {
    for (auto &d : S.plan) { d.fast = 0; d.pad = 0; }
  }
  return S;
}

}  // namespace mr
