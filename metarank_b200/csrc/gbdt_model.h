// Host-side booster representation, parsers and the HBM packing of the tree ensemble.
//
// Replaces ltrlib's `LightGBMBooster(bytes)` / `XGBoostBooster(bytes)` constructors
// (reference S/ml/rank/LambdaMARTRanker.scala:228-232).  The blobs are the public
// LightGBM model text / XGBoost JSON|UBJSON model formats.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "common.h"

namespace mr {

// decision flags shared by host and device (low 8 bits of DNode::ff >> 24)
enum : uint32_t {
  NF_CATEGORICAL = 1u,     // LightGBM kCategoricalMask
  NF_DEFAULT_LEFT = 2u,    // LightGBM kDefaultLeftMask / XGBoost default_left
  NF_MISSING_SHIFT = 2,    // bits 2-3: 0 None, 1 Zero, 2 NaN
  NF_NAN_LEFT = 16u,       // precomputed: where a NaN input goes at this node
};

// One decision tree in the unified form both parsers produce: internal nodes
// 0..n-1 with children >= 0 (internal) or < 0 (leaf ~c), leaves separate.
struct HostTree {
  std::vector<int32_t> feat;
  std::vector<double> thr;      // f64 threshold (LightGBM) or the exact f32 split_condition widened
  std::vector<uint8_t> flags;
  std::vector<int32_t> left, right;
  std::vector<double> leaf;     // f64 leaf (LightGBM) or f32 leaf widened (XGBoost)
  // categorical nodes: thr[i] is unused; cat_begin/cat_n index cat_words
  std::vector<int32_t> cat_begin, cat_n;  // per internal node (0,0 for numerical)
  std::vector<uint32_t> cat_words;
  int depth() const;
};

struct HostModel {
  int kind = 0;            // MR_BOOSTER_*
  int n_features = 0;
  float base_score = 0.f;  // XGBoost only
  std::vector<HostTree> trees;
  bool has_cat = false, has_zero_missing = false;
  int64_t n_internal = 0;
  int max_leaves = 0, max_depth = 0;
  std::vector<uint8_t> blob;  // original bytes for Booster.save()
};

HostModel parse_lightgbm_text(const uint8_t *blob, size_t len);
HostModel parse_xgboost_model(const uint8_t *blob, size_t len);

// Metarank model framing (LambdaMARTPredictor.load). Returns the booster kind and the
// [begin,end) byte range of the booster blob; fills names.
void parse_metarank_frame(const uint8_t *blob, size_t len, std::vector<std::string> &names, int &kind,
                          size_t &begin, size_t &end);

// ---- device layout -------------------------------------------------------------
// A model is a sequence of CHUNKS, each one contiguous, 16-byte aligned byte range that
// a single cp.async.bulk (TMA 1-D bulk copy) stages into shared memory:
//
//   +0   u32 n_trees, u32 reserved[3]
//   +16  u32 tab[2*n_trees]  {byte offset of tree's node array, byte offset of its leaves}
//        (padded to 16 B)
//   ...  per tree: DNode nodes[n_internal]   (16 B each)
//                  Real  leaves[n_leaves]    (f64 LightGBM / f32 XGBoost; padded to 16 B)
//   ...  u32 cat_words[] of the chunk's categorical nodes (padded to 16 B)
//
// DNode (16 B): { f64 thr | f32 thr | {u32 cat_word_off, u32 cat_n_words} ;
//                 u32 ff = feature | flags << 24 ; i16 left ; i16 right }
// children: >= 0 internal node index within the tree, < 0 leaf ~c.
struct ChunkDesc {
  uint32_t byte_off;   // from model base
  uint32_t bytes;      // multiple of 16
  uint32_t n_trees;
  uint32_t first_tree;
};

struct PackedModel {
  std::vector<uint8_t> bytes;
  std::vector<ChunkDesc> chunks;
  uint32_t max_chunk_bytes = 0;
};

// chunk_budget: maximum bytes per chunk (a single tree larger than it gets its own chunk).
PackedModel pack_model(const HostModel &m, size_t chunk_budget);

// ---- binned layout ----------------------------------------------------------------
// Every numerical split `x <= t` (LightGBM) / `x < t` (XGBoost) only asks on which side of t
// the value lies, so a feature value can be replaced, exactly, by its rank among the sorted
// distinct thresholds the model uses on that feature: bin(x) = #{t : t < x} (LightGBM,
// lower_bound) or #{t : t <= x} (XGBoost, upper_bound), and the node test becomes the
// integer compare `bin <= k` with k = the node threshold's index.  The traversal kernel then
// reads 8-byte nodes and 2-byte feature codes instead of 16-byte nodes and 8-byte doubles.
//
// BNode (8 B): { u16 k | cat-table index ; u16 feature(12) | flags(4) << 12 ; i16 left ; i16 right }
// feature codes (u16): 0..65533 bin | category value; 0xFFFF = NaN (or a category that is
// negative / not an int); 0xFFFE = value inside LightGBM's zero band (only when the model has
// missing_type == Zero nodes, which then also need the numeric bin -> such models fall back).
enum : uint32_t { BF_NAN_LEFT = 1u, BF_CATEGORICAL = 8u };
constexpr uint16_t kBinNaN = 0xFFFFu;

// Per-column bucket index that replaces most of the binary search: bucket(x) is a MONOTONE function of
// x, so every threshold in a lower bucket is < x and every threshold in a higher bucket is > x; only
// the thresholds sharing x's bucket (usually 0-2) are compared.  Exact for `<` and `<=` alike.
struct BinMeta {
  double mn, scale;   // bucket(x) = x <= mn ? 0 : min(g - 1, (int)((x - mn) * scale))
  uint32_t g;         // number of buckets (>= 1)
  uint32_t idx_off;   // offset of this column's g entries in `bucket_range`
  uint32_t thr_off;   // offset of this column's thresholds in `thr`
  // Where the code goes in the scorer's tile.  bit 0: categorical column (code = the integer value);
  // bit 1: a NaN is stored as code 0 in the base column (every split of this feature sends NaN left),
  // otherwise as 0xFFFF (never <= k: right); bits 16-31: a second tile column that holds the same codes
  // with NaN -> 0, referenced by the NaN-left nodes of a feature whose nodes disagree (0xFFFF: none).
  // With the direction baked into the column the traversal needs no NaN test at all.
  uint32_t flags;
};
constexpr uint32_t kMetaCat = 1u, kMetaNanLow = 2u, kMetaCat16 = 4u, kMetaNoDup = 0xFFFFu;
// Small-categorical code form (BinnedModel::cat16: every bitset of the model lives in categories 0..15).  A categorical
// column's code is then 0xFFF0 | category, 0xFFE2 for NaN / negative / >= 16: both are binary16 NaN patterns, so the slim
// scorer's numeric compare `code <= k` is false on them whatever the entry holds, and `code & 31` = 16 + category (2 for
// the missing code) is the position, inside a categorical entry, of the node's bitset bit (the bitset sits in the entry's
// high half; bit 2 of an entry is 0 — or 1 in an entry that stores the complement of its set, see pack_slim).  The level loop then resolves a categorical node with two more instructions
// instead of leaving.  Every other consumer turns the code back into the category with cat_of_code().
constexpr uint16_t kCat16Base = 0xFFF0u, kCat16Missing = 0xFFE2u;

struct BinnedModel {
  bool ok = false;                 // false: model cannot be binned exactly -> use the f64/f32 kernel
  std::vector<uint32_t> thr_off;   // [n_features + 1] offsets into thr (numerical) per feature
  std::vector<double> thr;         // sorted distinct thresholds, all features back to back
  std::vector<BinMeta> meta;       // [n_features]
  std::vector<uint32_t> bucket_range;  // per bucket: first threshold index | one-past-last << 16 (within the column)
  std::vector<uint8_t> is_cat;     // [n_features] feature is split categorically
  bool cat16 = false;              // categorical columns carry the small-categorical code form (kMetaCat16)
  int tile_cols = 0;               // columns of the code tile: n_features + duplicated (mixed NaN direction) columns
  PackedModel packed;              // chunks of BNodes (+ leaves, + categorical tables)
};
BinnedModel pack_binned(const HostModel &m, size_t chunk_budget);

// "Compact" form of the binned model for the fast lock-step kernel (<= 1023 tile columns, chunks <= 64 KB).
// Children are BYTE offsets inside the chunk whose low bits say what they point at — bit 0 = leaf (the offset
// then addresses the leaf VALUE), bit 1 = categorical node — so a level needs no address arithmetic and the
// loop's leaf test also catches categorical nodes:
//   CNode (8 B): word0 = tile_column*64 (bits 6..15) | categorical (bit 1) | k << 16
//                word1 = left pointer (16) | right pointer (16)
//   no NaN flag: the direction is a property of the tile column (BinMeta::flags, BinnedModel::tile_cols)
//   a categorical node's k is the 8-byte index (in the chunk) of its {bitset word offset, n words} pair
//   chunk: +0 u32 n_trees, pad; +16 u32 root[n_trees] (pointer to the root node; a single-leaf tree is a dummy
//          split whose children both are its leaf); then per tree its nodes followed by its leaf values (8-byte slots)
// chunk_budget 0 = sized from the code tile (what ~48 resident warps leave of the shared memory).
BinnedModel pack_compact(const HostModel &m, const BinnedModel &binned, size_t chunk_budget);

// "Slim" form of the compact model for the throughput scorer's fast path: 4-byte nodes, so that a warp whose lanes sit
// on different nodes still reads them in ONE shared-memory wavefront (an 8-byte load splits into two half-warp passes).
// A tree is a power-of-two sized and aligned BLOCK of 4-byte entries followed by its 8-byte leaf values:
//   entry 0 = the root, entries (2j, 2j+1) for j >= 1 = the two children of an internal node, side by side;
//   internal entry  bits 31..16  k as a binary16 bit pattern (k <= 0x7BFF: non-negative halves order like integers, and a
//                                NaN code, 0xFFFF, fails every `<=`), compared by ONE HSETP2 against the code's low half
//                   bits 15..s   (tile column pair + 1), i.e. the byte offset of that pair's row in the CTA's code tile
//                                (s = log2(4 * tile_T): the tile is [column pair][item of the CTA] u32, pair p at (p+1) << s)
//                   bits s-1..3  byte offset of the node's child pair inside the block (so blocks are <= 4 * tile_T bytes)
//                   bit 1        which half of the pair's u32 holds this column (the byte offset 2)
//   leaf entry      bit 31 set, bits 15..0 = byte offset of the leaf's 8-byte value slot inside the block
// One level of a walk is LOP3 (code address) -> LDS.U16 -> HSETP2 -> LOP3 (child pair | block base) -> predicated +4 ->
// LDS.32 (next entry) -> sign test -> branch: 8 instructions and 2 wavefronts, against 10 and ~3.4 for the 8-byte layout.
//   chunk: +0 u32 n_trees, pad; +16 per tree {u32 block offset from the chunk start, u32 copy of its root entry} (chunk
//          buffers are 2 KB aligned in shared memory and every block is aligned to its own size; a single-leaf tree is a
//          dummy split onto its leaf, so every walk starts at an internal entry that arrives with the table); then the blocks
//   categorical entry: bit 0 set (the level loop leaves on it with the test that finds a leaf), bits 30..16 = 8-byte index
//                   inside the block of the node's {bitset byte offset, n words}; column and child fields as above
// ok == false (> 31 744 thresholds on a column, too many tile columns, or a tree — with its bitsets — too large for any
// tile_T): the 8-byte compact kernel scores the model.  The tile mapping (BinMeta, tile_cols) is the compact model's.
//   cat16 models (kMetaCat16): a categorical entry is bit 0 set, bits 31..16 = the node's 16-bit bitset, column and child
//                   fields as above, and nothing else in the block; the level loop handles it in place
// root_tab (models of <= kSlimRootTabMax trees): per tree, in tree order, {block offset from the chunk start, root entry,
// left child entry, right child entry}.  The scorer receives it as a kernel PARAMETER (constant bank): level 0 of every
// walk then costs no node load and no root-table load on the shared-memory pipe — the root entry is warp-uniform anyway.
constexpr int kSlimRootTabMax = 1920;  // 16 B per tree inside the 32 764-byte parameter space
struct SlimModel {
  bool ok = false;
  bool cat16 = false;   // categorical nodes are in the in-loop form (needs the compact model's cat16 codes)
  int tile_T = 0;       // items per CTA = threads per CTA: 512, 256 or 128
  int n_pairs = 0;      // column pairs of the tile
  int col_base = 1;     // pair p of the CTA's code tile lives at absolute shared address (p + col_base) * 4 * tile_T
  PackedModel packed;
  std::vector<uint32_t> root_tab;  // 4 words per tree, or empty (too many trees / wide categorical nodes)
};
// max_T: largest tile size to consider (mr_model_selfcheck packs the smaller ones too); warps_per_sm: the residency the
// automatic chunk budget leaves room for;
// min_tile_addr: lowest absolute shared address the code tile may start at (the kernel's dynamic window begins at 1 KB on
// sm_100 and holds three mbarriers first) — the tile's pair p lives at (p + col_base) * 4 * tile_T.
SlimModel pack_slim(const HostModel &m, const BinnedModel &compact, size_t chunk_budget, int max_T = 512, int warps_per_sm = 48,
                    uint32_t min_tile_addr = 1088);
// host-side layout check of pack_slim's output (mr_model_selfcheck): number of (sample, tree) leaf mismatches
size_t slim_pack_selfcheck(const HostModel &m, const BinnedModel &compact, const SlimModel &slim, int n_samples, uint64_t seed);

}  // namespace mr
