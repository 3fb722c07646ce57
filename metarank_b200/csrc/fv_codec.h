// The reference's BINARY store format for Persistence.values: a stream of delimited
// FeatureValue records, transcoded on the host into mr_state_upsert records (include/mr_b200.h).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace mr {

struct FvStats {
  int64_t records = 0;      // FeatureValues decoded
  int64_t unsupported = 0;  // of those: classes no supported extractor reads (NumStats / Map / Frequency,
                            // bounded lists of non-strings) — not transcoded
  size_t consumed = 0;      // bytes of whole records read; < len when the tail is a truncated record
};

// Appends the upsert records to `out`.  Throws mr::Error(MR_ERR_PARSE) on an unknown tag or a record that
// ends before its fields do.
void transcode_feature_values(const uint8_t *bytes, size_t len, std::vector<uint8_t> &out, FvStats &stats);

}  // namespace mr
