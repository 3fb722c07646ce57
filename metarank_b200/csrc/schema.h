// FeatureMapping for one model: parsed Metarank feature schemas -> extractor plan, state
// slots and the dense column layout.  Mirrors S/FeatureMapping.scala:56-99 (feature list,
// Schema of state configs, DatasetDescriptor in model-feature order) for the extractors on
// the /rank hot path.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.h"

namespace mr {

enum FeatKind : int32_t {
  FK_NUMBER = 0,      // number, word_count (stored SDouble)     S/feature/NumberFeature.scala, WordCountFeature.scala
  FK_CATEGORY = 1,    // string, encode: index                   S/feature/StringFeature.scala:124-137
  FK_ONEHOT = 2,      // string, encode: onehot                  S/feature/StringFeature.scala:118-122
  FK_COUNT = 3,       // interaction_count                       S/feature/InteractionCountFeature.scala:44-59
  FK_WINDOW = 4,      // window_count                            S/feature/WindowInteractionCountFeature.scala:50-63
  FK_RATE = 5,        // rate (plain + normalized)               S/feature/RateFeature.scala:290-356
  FK_INTERACTED = 6,  // interacted_with, one entry per field    S/feature/InteractedWithFeature.scala:133-164
  FK_RELEVANCY = 7,   // relevancy                               S/feature/RelevancyFeature.scala:36-51
  FK_POSITION = 8,    // position (online: constant)             S/feature/PositionFeature.scala:30-35
  FK_DIVERSITY = 9,   // diversity                               S/feature/DiversityFeature.scala:67-130
  FK_COSINE = 10,     // field_match / bi-encoder                S/feature/FieldMatchBiencoderFeature.scala:80-109
  FK_CONST_REQ = 11,  // number/word_count/string with scope ranking, local_time: one value per request
  FK_VECTOR = 12,     // vector (stored, already reduced SDoubleList)   S/feature/NumVectorFeature.scala:55-70
  FK_ITEM_AGE = 13,   // item_age                                 S/feature/ItemAgeFeature.scala:74-86
  FK_TOKEN_MATCH = 14,  // field_match ngram / term / bm25 over pre-tokenized strings   S/feature/FieldMatchFeature.scala:60-93
};

enum ScopeT : int32_t { SC_GLOBAL = 0, SC_ITEM = 1, SC_USER = 2, SC_SESSION = 3, SC_FIELD = 4, SC_IRF = 5, SC_RANKING = 6, SC_N_TABLES = 6 };

enum SlotKind : int32_t {
  SK_F64 = 0,        // ScalarValue(SDouble)                  1 word
  SK_STRID = 1,      // ScalarValue(SString) as hash          1 word
  SK_CAT = 2,        // SStringList pre-encoded: index+1 / onehot mask   1 word
  SK_COUNTER = 3,    // CounterValue                          1 word
  SK_PCOUNTER = 4,   // PeriodicCounterValue                  P words (+ presence only if length == P)
  SK_STRLIST = 5,    // ScalarValue(SStringList|SString) hashes: {u32 off, u32 len} into the table's pool
                     // (Slot::p == 1: kept sorted by hash and unique, for field_match's token sets)
  SK_BLIST = 6,      // BoundedListValue item hashes:          {u32 off, u32 len}
  SK_F64LIST = 7,    // ScalarValue(SDoubleList) dim doubles in a side array (presence bit only)
  SK_DIVERSITY = 8,  // 2 words: {kind 1 double | 2 strings, payload f64 | {off,len}}
  SK_BOOL = 9,       // ScalarValue(SBoolean) stored as the double it reads as (1.0 / 0.0)   1 word
  SK_F64VEC = 10,    // ScalarValue(SDoubleList) of the feature's dim, inline in the row     dim words
};

struct Slot {
  std::string name;  // Key.feature
  int table = 0;     // ScopeT (< SC_N_TABLES)
  int kind = 0;
  int word = 0;      // first payload word within the row
  int n_words = 1;
  int bit = 0;       // presence bit index
  int p = 0;         // periods (PCOUNTER) / dim (F64LIST)
  int side = -1;     // side-array index (F64LIST)
  int feature = -1;  // owning feature (for CAT encoding)
  // write-path configuration (PeriodicCounterConfig / BoundedListConfig of the owning extractor)
  int64_t period_ms = 0;          // PCOUNTER: bucket size
  std::vector<int> ranges;        // PCOUNTER: sumPeriodRanges start offsets (end offset is 0)
  int list_count = 0;             // BLIST: max entries
  int64_t list_duration_ms = 0;   // BLIST: max age
};

// Device-visible extractor descriptor (POD).
struct DFeature {
  int32_t kind, col, dim, scope;
  int32_t w[4], b[4];  // slot words / presence bits (meaning depends on kind)
  int32_t in0, in1;    // request-input slot indices (-1 none)
  int32_t aux0, aux1, aux2, aux3;
  double dparam;
  uint64_t uparam;
  int32_t fast;        // 1: every column of this entry is produced by row_gather_kernel (see FastCol)
  int32_t pad;
};

// One output column that is a plain function of ONE word of the item's row: item-scoped number /
// boolean / word_count (f64), string-index (i32 -> f64), interaction_count and window_count (i64 -> f64),
// vector (f64).  These are gathered by row_gather_kernel with one coalesced row load per item.
struct FastCol {
  uint16_t word;      // row word holding the value
  uint16_t bit;       // presence bit
  uint8_t conv;       // 0 f64 bits, 1 i64 -> f64, 2 i32 -> f64
  uint8_t missing;    // 0 NaN, 1 0.0
  int16_t override_slot;  // MR_IN_ITEM_F64 slot whose non-NaN value wins, or -1
  uint16_t col;       // output column
  uint16_t pad;
};

struct FeatureDef {  // host-side description of one configured feature
  std::string name, type;
  int kind = 0;
  int dim = 1;
  int scope = SC_ITEM;
  std::string scope_field;  // item.<field> / ranking.<field>
  std::vector<std::string> cat_values;
  std::vector<uint64_t> cat_hashes;
  int col = -1;  // first dense column, -1 when not in the model
  std::vector<int> slots;  // indices into Schema::slots
};

struct TableLayout {
  int n_slots = 0;
  int presence_words = 0;
  int row_words = 0;  // total u64 words per row
};

struct SideArray { int table; int dim; int slot; };

struct Schema {
  std::string source_json;                   // the document handed to mr_schema_create (request decoder re-reads the configs)
  std::vector<FeatureDef> features;          // config order
  std::vector<std::string> model_features;   // model order
  std::vector<Slot> slots;
  std::unordered_map<std::string, int> slot_by_name;
  TableLayout tables[SC_N_TABLES];
  std::vector<SideArray> sides;
  std::vector<DFeature> plan;                // extractor entries in column order
  bool fast_has_override = false;            // some fast column can be overridden by a request-item field
  std::vector<FastCol> fast_cols;            // columns handled by the coalesced row-gather kernel
  int dim = 0;
  // request inputs
  std::vector<std::string> in_req_f64, in_req_u64, in_item_f64;  // owning feature names
  std::vector<std::string> in_req_tok;                           // per-request token lists (field_match ngram/term/bm25)
  struct VecIn { std::string feature; int dim; int offset; };
  std::vector<VecIn> in_req_vec;
  int vec_stride = 0;
  // per-request aggregates
  int n_hist = 0;       // sorted tag multisets (interacted_with fields, diversity strings)
  int n_reqagg = 0;     // per-request scalar aggregates (diversity median / totals, cosine min/max)
  bool needs_prepass = false, needs_cosine = false, needs_visitor = false;
  std::unordered_map<std::string, std::pair<int, int>> col_of;  // model feature -> (offset, dim)
};

Schema parse_schema_json(const char *json, size_t len);

int64_t parse_duration_ms(const std::string &s, const std::string &feature);  // DurationJson: "[0-9]+[smhd]"
uint64_t hash64(const void *bytes, size_t len);
int32_t token_count(const char *s, size_t len);
inline uint64_t hash_combine(uint64_t a, uint64_t b) {
  uint64_t x = a ^ (b + 0x9E3779B97F4A7C15ull + (a << 6) + (a >> 2));
  x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32;
  return x ? x : 1;
}

}  // namespace mr
