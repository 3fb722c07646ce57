// Launch interface of the feature-assembly kernels (assemble_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "binning.cuh"
#include "schema.h"
#include "state.h"

namespace mr {

struct RankArgs {
  DState st;
  const DFeature *plan;  // device copy of Schema::plan
  int n_plan, dim;
  // batch (device pointers)
  int n_requests, total_items;
  const int32_t *item_offsets;
  const uint64_t *item_ids, *user_ids, *session_ids;
  const double *req_f64;
  const uint64_t *req_u64;
  const float *req_vec;
  const uint8_t *req_vec_present;
  const double *item_f64;
  // per-request token lists (field_match ngram/term/bm25): request r, list slot s owns tokens
  // [req_tok_off[r * n_req_tok + s], req_tok_off[r * n_req_tok + s + 1]) - req_tok_base of hashes / weights
  const int32_t *req_tok_off;
  const uint64_t *req_tok_hash;
  const double *req_tok_w;
  int32_t req_tok_base;
  int n_req_f64, n_req_u64, n_req_vec, vec_stride, n_item_f64, n_req_tok;
  // scratch (device)
  int32_t *item_req;      // [total_items] owning request
  uint32_t *item_row;     // [total_items] row in the item table or 0xFFFFFFFF
  uint32_t *visitor_row;  // [n_requests x 2] user row, session row
  double *cos;            // [2 x n_cos x total_items] raw, then normalised
  double *qnorm;          // [n_requests x n_cos] CosineDistance's aSum of the request's query embedding (lookup_kernel)
  double *reqagg;         // [n_requests x n_reqagg x 4]
  uint2 *hist_desc;       // [n_requests x n_hist] {offset, length | unsorted << 31}
  uint64_t *hist_pool;
  uint32_t hist_pool_cap;
  uint32_t *hist_cursor;
  int32_t *error_flag;    // 0 ok | MR_ERR_ARITHMETIC | -1 histogram pool overflow
  int n_hist, n_reqagg, n_cos;
  // output
  double *out_features;   // [total_items x dim] row-major (ltrlib Query.values); may be null when codes != null
  uint16_t *codes;        // optional: u16 rank codes [group of 32 items][column][lane] for the binned scorer
  BinParams bin;          // thresholds of the model the codes are for (valid when codes != null)
  int stage_meta;         // set by launch_assemble: bucket-index headers staged in shared memory
  const FastCol *fast_cols;  // device copy of Schema::fast_cols (row_gather_kernel)
  int n_fast;
  // per-model code rows of the item table (rank_api.cu CodeCache): row r + 1 holds item row r's codes for
  // the fast columns in tile-column order, row 0 the codes of an unknown item.  nullptr: compute from the rows.
  const uint32_t *code_rows;
  int code_row_words;        // u32 words per code row
};

// Enqueues lookup -> cosine -> per-request prepass -> assemble on `stream`.
void launch_assemble(const RankArgs &a, const Schema &schema, cudaStream_t stream);
// Ranker.rerank's sortBy(-score): order[off[r] + k] = index (within the request) of the k-th item.
// (Re)computes code rows: all item rows [0, n_rows) plus the unknown-item row when idx == nullptr, else the
// rows listed in idx.  a.st / a.fast_cols / a.n_fast / a.bin describe the source and the model's binning.
void launch_code_rows(const RankArgs &a, uint32_t *code_rows, int code_row_words, uint32_t n_rows, const uint32_t *d_idx,
                      uint32_t n_idx, cudaStream_t stream);
// max_items_hint: the largest request of the batch when the host knows it (0 = unknown: every size class is launched
// and returns at once where it does not apply).  d_rank_tmp: 3 * total_items ints (8-byte aligned) of scratch for the mega-request path
// (nullptr: stream-ordered allocation when needed).
void launch_rank_order(const double *d_scores, const int32_t *d_item_offsets, int n_requests, int total_items,
                       int32_t *d_order, cudaStream_t stream, int max_items_hint = 0, int32_t *d_rank_tmp = nullptr);

}  // namespace mr
