// Native request decoder: RankingEvent JSON -> the flat arrays of mr_rank_batch.
// Mirrors the decoding rules of the reference's circe codecs (S/model/Event.scala:44-99, S/model/Field.scala:36-58,
// S/model/Timestamp.scala) and the per-extractor reads of the request (the files under S/feature/).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../include/mr_b200.h"
#include "schema.h"

namespace mr {

struct PackedRequests {
  int32_t n_requests = 0, total_items = 0;
  std::vector<int32_t> offsets;
  std::vector<uint64_t> ids, users, sessions;
  std::vector<double> req_f64;
  std::vector<uint64_t> req_u64;
  std::vector<float> req_vec;
  std::vector<uint8_t> req_vp;
  std::vector<double> item_f64;
  bool has_item_f64 = false;
  std::vector<int32_t> tok_off;
  std::vector<uint64_t> tok_hash;
  std::vector<double> tok_w;
  std::vector<int64_t> timestamps;       // per request, epoch millis
  std::vector<std::string> item_ids;     // per item, for the response
  std::vector<std::string> request_ids;  // per request
  mr_rank_batch batch{};                 // views into the vectors above
};

// What the decoder needs to know about each model feature (source fields, encoders, matcher settings, input
// slots), read once from the schema document.
struct RequestPlan;
std::shared_ptr<const RequestPlan> make_request_plan(const Schema &S);

// json: one RankingEvent object or an array of them.  Throws mr::Error (MR_ERR_PARSE for a body the reference's
// decoder rejects, MR_ERR_INVALID_ARG for inputs only this path needs, e.g. missing caller-side tokens).
void decode_requests(const Schema &S, const RequestPlan &plan, const char *json, size_t len, PackedRequests &out);

}  // namespace mr
