// See state.h.  Wire format of mr_state_upsert is documented in include/mr_b200.h; the
// FeatureValue kinds are those of reference S/model/FeatureValue.scala:18-50.
#include "state.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace mr {

uint32_t HostTable::find(uint64_t key) const {
  if (keys.empty()) return UINT32_MAX;
  const size_t mask = keys.size() - 1;
  size_t i = (size_t)mix64(key) & mask;
  for (;;) {
    if (keys[i] == key) return vals[i];
    if (keys[i] == 0) return UINT32_MAX;
    i = (i + 1) & mask;
  }
}

void HostTable::grow_map() {
  const size_t cap = keys.empty() ? 1024 : keys.size() * 2;
  std::vector<uint64_t> nk(cap, 0);
  std::vector<uint32_t> nv(cap, 0);
  for (size_t i = 0; i < keys.size(); i++) {
    if (!keys[i]) continue;
    size_t j = (size_t)mix64(keys[i]) & (cap - 1);
    while (nk[j]) j = (j + 1) & (cap - 1);
    nk[j] = keys[i];
    nv[j] = vals[i];
  }
  keys.swap(nk);
  vals.swap(nv);
  map_dirty = true;
}

uint32_t HostTable::find_or_insert(uint64_t key) {
  if (key == 0) key = 1;
  if (keys.empty() || (n_rows + 1) * 2 > keys.size()) grow_map();
  const size_t mask = keys.size() - 1;
  size_t i = (size_t)mix64(key) & mask;
  for (;;) {
    if (keys[i] == key) return vals[i];
    if (keys[i] == 0) break;
    i = (i + 1) & mask;
  }
  if (n_rows >= (size_t)UINT32_MAX - 1) fail(MR_ERR_UNSUPPORTED, "state table is full");
  const uint32_t row = (uint32_t)n_rows++;
  keys[i] = key;
  vals[i] = row;
  map_dirty = true;
  rows.resize(n_rows * (size_t)row_words, 0);
  return row;
}

StateStore::StateStore(const Schema &s) : schema(s) {
  for (int t = 0; t < SC_N_TABLES; t++) tables[t].row_words = schema.tables[t].row_words;
  if (schema.sides.size() > (size_t)kMaxSides) fail(MR_ERR_UNSUPPORTED, "at most %d embedding features are supported", kMaxSides);
  for (size_t i = 0; i < schema.sides.size(); i++) {
    HostTable &T = tables[schema.sides[i].table];
    T.side_ids.push_back((int)i);
    T.sides.emplace_back();
    T.d_sides.push_back(nullptr);
    T.d_sides_cap.push_back(0);
    T.side_bs.emplace_back();
    T.side_is_f32.push_back(1);
    T.side_dev_mode.push_back(0);
    T.d_sides_f32.push_back(nullptr);
    T.d_sides_f32_cap.push_back(0);
    T.d_side_bs.push_back(nullptr);
    T.d_side_bs_cap.push_back(0);
  }
  // the global scope always has its single row
  tables[SC_GLOBAL].find_or_insert(1);
}

StateStore::~StateStore() {
  for (auto &T : tables) {
    if (T.d_keys) cudaFree(T.d_keys);
    if (T.d_vals) cudaFree(T.d_vals);
    if (T.d_rows) cudaFree(T.d_rows);
    if (T.d_pool) cudaFree(T.d_pool);
    for (auto p : T.d_sides) if (p) cudaFree(p);
    for (auto p : T.d_sides_f32) if (p) cudaFree(p);
    for (auto p : T.d_side_bs) if (p) cudaFree(p);
  }
}

namespace {
struct Reader {
  const uint8_t *p, *e;
  template <class T> T get() {
    if ((size_t)(e - p) < sizeof(T)) fail(MR_ERR_PARSE, "state record truncated");
    T v;
    memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  const uint8_t *bytes(size_t n) {
    if ((size_t)(e - p) < n) fail(MR_ERR_PARSE, "state record truncated");
    const uint8_t *q = p;
    p += n;
    return q;
  }
};
}  // namespace

uint32_t StateStore::row_for(const Slot &sl, int scope, uint64_t id0, uint64_t id1) {
  uint64_t key;
  if (scope == SC_GLOBAL) key = 1;
  else if (scope == SC_FIELD) {
    const FeatureDef &fd = schema.features[sl.feature];
    key = hash_combine(hash64(fd.scope_field.data(), fd.scope_field.size()), id0);
  } else if (scope == SC_IRF) {
    const FeatureDef &fd = schema.features[sl.feature];
    key = hash_combine(hash_combine(hash64(fd.scope_field.data(), fd.scope_field.size()), id0), id1);
  } else key = id0 ? id0 : 1;
  return tables[sl.table].find_or_insert(key);
}

void StateStore::upsert(const uint8_t *buf, size_t len, int64_t *applied, int64_t *skipped) {
  std::unique_lock<std::shared_mutex> g(mu);
  // one record of the packed buffer (include/mr_b200.h "state record")
  struct Rec { std::string name; uint8_t scope = 0, kind = 0; uint64_t id0 = 0, id1 = 0; double f64 = 0; uint64_t u64 = 0; int64_t i64 = 0;
               uint32_t n = 0; const uint8_t *arr = nullptr; };
  auto read_record = [](Reader &r, Rec &c) {
    const uint16_t nl = r.get<uint16_t>();
    c.name.assign((const char *)r.bytes(nl), nl);
    c.scope = r.get<uint8_t>();
    c.id0 = c.id1 = 0;
    switch (c.scope) {
      case SC_GLOBAL: break;
      case SC_ITEM: case SC_USER: case SC_SESSION: case SC_RANKING: case SC_FIELD: c.id0 = r.get<uint64_t>(); break;
      case SC_IRF: c.id0 = r.get<uint64_t>(); c.id1 = r.get<uint64_t>(); break;
      default: fail(MR_ERR_PARSE, "state record: bad scope tag %d", (int)c.scope);
    }
    c.kind = r.get<uint8_t>();
    c.f64 = 0; c.u64 = 0; c.i64 = 0; c.n = 0; c.arr = nullptr;
    switch (c.kind) {
      case 0: c.f64 = r.get<double>(); break;
      case 1: c.u64 = r.get<uint64_t>(); break;
      case 2: case 3: case 5: case 6: c.n = r.get<uint32_t>(); c.arr = r.bytes((size_t)c.n * 8); break;
      case 4: c.i64 = r.get<int64_t>(); break;
      case 7: c.u64 = r.get<uint8_t>(); break;  // SBoolean
      default: fail(MR_ERR_PARSE, "state record: bad value kind %d", (int)c.kind);
    }
  };
  // first pass: the whole buffer is checked (framing, tags, vector / embedding dims) before any table is touched, so a
  // malformed or mismatching record later in the batch cannot leave the earlier ones half applied
  {
    Reader v{buf, buf + len};
    Rec c;
    while (v.p < v.e) {
      read_record(v, c);
      auto it = schema.slot_by_name.find(c.name);
      if (it == schema.slot_by_name.end()) continue;
      const Slot &sl = schema.slots[it->second];
      if (sl.table != (int)c.scope || c.kind != 3) continue;
      if (sl.kind == SK_F64VEC && (int)c.n != sl.p)
        fail(MR_ERR_INVALID_ARG, "vector '%s' has %u values, the feature's dim is %d", c.name.c_str(), c.n, sl.p);
      if (sl.kind == SK_F64LIST && (int)c.n != sl.p)
        fail(MR_ERR_INVALID_ARG, "embedding '%s' has %u values, the schema says dim %d", c.name.c_str(), c.n, sl.p);
    }
  }
  Reader r{buf, buf + len};
  int64_t n_ok = 0, n_skip = 0;
  Rec rec;
  while (r.p < r.e) {
    read_record(r, rec);
    const std::string &name = rec.name;
    const uint8_t scope = rec.scope, kind = rec.kind;
    const uint64_t id0 = rec.id0, id1 = rec.id1, u64 = rec.u64;
    const double f64 = rec.f64;
    const int64_t i64 = rec.i64;
    uint32_t n = rec.n;
    const uint8_t *arr = rec.arr;
    auto it = schema.slot_by_name.find(name);
    if (it == schema.slot_by_name.end()) { n_skip++; continue; }
    const Slot &sl = schema.slots[it->second];
    if (sl.table != (int)scope) { n_skip++; continue; }  // same name under another scope: not what the extractor reads
    HostTable &T = tables[sl.table];
    const uint32_t row = row_for(sl, scope, id0, id1);
    uint64_t *w = T.rows.data() + (size_t)row * T.row_words;
    auto set_present = [&](bool on) {
      if (on) w[sl.bit >> 6] |= 1ull << (sl.bit & 63);
      else w[sl.bit >> 6] &= ~(1ull << (sl.bit & 63));
    };
    // List payloads live in a per-(slot, row) region of the table's pool that is rewritten in place while the new
    // list fits (FeatureValueSink re-emits the same keys continuously: appending every time would grow the pool
    // without bound); a list that outgrows its region moves to a new one of twice the size.
    const RawKey rk{((uint64_t)sl.table << 56) ^ ((uint64_t)it->second << 40) ^ (uint64_t)row};
    auto put_words = [&](uint64_t *dst, const uint64_t *src, uint32_t cnt) {
      auto reg = upsert_region.find(rk);
      if (reg == upsert_region.end() || reg->second.second < cnt) {
        const uint32_t cap = std::max<uint32_t>(cnt, reg == upsert_region.end() ? 0u : 2u * reg->second.second);
        if (T.pool.size() + cap > (size_t)UINT32_MAX) fail(MR_ERR_UNSUPPORTED, "state pool is full");
        const uint32_t off = (uint32_t)T.pool.size();
        T.pool.resize(T.pool.size() + cap, 0);
        if (reg == upsert_region.end()) reg = upsert_region.emplace(rk, std::make_pair(off, cap)).first;
        else reg->second = std::make_pair(off, cap);
      }
      const uint32_t off = reg->second.first;
      if (cnt) memcpy(T.pool.data() + off, src, (size_t)cnt * 8);
      if (off < T.pool_uploaded && cnt) {
        T.pool_dirty_lo = std::min<size_t>(T.pool_dirty_lo, off);
        T.pool_dirty_hi = std::max<size_t>(T.pool_dirty_hi, std::min<size_t>((size_t)off + cnt, T.pool_uploaded));
      }
      *dst = (uint64_t)off | ((uint64_t)cnt << 32);
    };
    auto put_list = [&](uint64_t *dst) { put_words(dst, reinterpret_cast<const uint64_t *>(arr), n); };
    bool ok = true;
    switch (sl.kind) {
      case SK_F64:
        if (kind == 0) { memcpy(&w[sl.word], &f64, 8); set_present(true); }
        else set_present(false);  // a non-SDouble scalar reads as missing (NumberFeature.value's match)
        break;
      case SK_BOOL: {
        // BooleanFeature.value: SBoolean -> 1.0 / 0.0, anything else reads as missing
        if (kind == 7) { const double v = u64 ? 1.0 : 0.0; memcpy(&w[sl.word], &v, 8); set_present(true); }
        else set_present(false);
        break;
      }
      case SK_F64VEC:
        if (kind != 3) { set_present(false); break; }
        if ((int)n != sl.p) fail(MR_ERR_INVALID_ARG, "vector '%s' has %u values, the feature's dim is %d", name.c_str(), n, sl.p);
        memcpy(&w[sl.word], arr, (size_t)n * 8);
        set_present(true);
        break;
      case SK_STRID:
        if (kind == 1) { w[sl.word] = u64; set_present(true); } else set_present(false);
        break;
      case SK_CAT: {
        // StringFeature stores SStringList; the encoder is applied here, once, instead of per request
        if (kind != 2 && kind != 1) { set_present(false); break; }
        const FeatureDef &fd = schema.features[sl.feature];
        std::vector<uint64_t> vals;
        if (kind == 1) vals.push_back(u64);
        else { vals.resize(n); if (n) memcpy(vals.data(), arr, (size_t)n * 8); }
        const bool onehot = fd.kind == FK_ONEHOT;
        uint64_t enc = 0;
        if (!onehot) {
          // IndexCategoricalEncoder: first value -> index + 1, unknown / empty -> 0
          if (!vals.empty())
            for (size_t k = 0; k < fd.cat_hashes.size(); k++)
              if (fd.cat_hashes[k] == vals[0]) { enc = k + 1; break; }
        } else {
          for (uint64_t v : vals)
            for (size_t k = 0; k < fd.cat_hashes.size(); k++)
              if (fd.cat_hashes[k] == v) { enc |= 1ull << k; break; }  // indexOf: first match
        }
        w[sl.word] = enc;
        set_present(true);
        break;
      }
      case SK_COUNTER:
        if (kind == 4) { memcpy(&w[sl.word], &i64, 8); set_present(true); } else set_present(false);
        break;
      case SK_PCOUNTER:
        // readers require values.length == dim (e.g. S/feature/RateFeature.scala:318-319); a value
        // of another length behaves as missing
        if (kind == 5 && (int)n == sl.p) { memcpy(&w[sl.word], arr, (size_t)n * 8); set_present(true); }
        else set_present(false);
        buckets.erase(rk);  // the refreshed value supersedes the day buckets apply_writes kept (see apply_writes)
        break;
      case SK_STRLIST:
        if (kind == 2 && sl.p == 1) {
          // field_match token set: kept sorted by hash and unique so the kernel can binary-search it
          std::vector<uint64_t> v(n);
          if (n) memcpy(v.data(), arr, (size_t)n * 8);
          std::sort(v.begin(), v.end());
          v.erase(std::unique(v.begin(), v.end()), v.end());
          put_words(&w[sl.word], v.data(), (uint32_t)v.size());
          set_present(true);
        } else if (kind == 2) { put_list(&w[sl.word]); set_present(true); }
        else set_present(false);  // InteractedWith / FieldMatch collect only SStringList
        break;
      case SK_BLIST:
        if (kind == 6) { put_list(&w[sl.word]); set_present(true); } else set_present(false);
        lists.erase(rk);
        list_region.erase(rk);
        break;
      case SK_F64LIST: {
        if (kind != 3) { set_present(false); break; }
        if ((int)n != sl.p) fail(MR_ERR_INVALID_ARG, "embedding '%s' has %u values, the schema says dim %d", name.c_str(), n, sl.p);
        int local = -1;
        for (size_t k = 0; k < T.side_ids.size(); k++) if (T.side_ids[k] == sl.side) local = (int)k;
        auto &S = T.sides[local];
        if (S.size() < (size_t)T.n_rows * sl.p) S.resize((size_t)T.n_rows * sl.p, 0.0);
        memcpy(S.data() + (size_t)row * sl.p, arr, (size_t)n * 8);
        {
          auto &BS = T.side_bs[local];
          if (BS.size() < T.n_rows) BS.resize(T.n_rows, 0.0);
          // CosineDistance's bSum (S/ml/onnx/distance/DistanceFunction.scala:22): sequential, product rounded before the add
          volatile double bs = 0.0;
          bool f32 = T.side_is_f32[local] != 0;
          const double *e = S.data() + (size_t)row * sl.p;
          for (int k = 0; k < sl.p; k++) {
            volatile double sq = e[k] * e[k];
            bs = bs + sq;
            f32 = f32 && ((double)(float)e[k] == e[k] || e[k] != e[k]);  // NaN payloads do not matter: NaN * x is NaN either way
          }
          BS[row] = bs;
          T.side_is_f32[local] = f32;
        }
        set_present(true);
        break;
      }
      case SK_DIVERSITY:
        // DiversityFeature keeps whatever scalar the item field had: number | string | string[]
        if (kind == 0) { w[sl.word] = 1; memcpy(&w[sl.word + 1], &f64, 8); set_present(true); }
        else if (kind == 1) { n = 1; arr = (const uint8_t *)&u64; w[sl.word] = 2; put_list(&w[sl.word + 1]); set_present(true); }
        else if (kind == 2) { w[sl.word] = 2; put_list(&w[sl.word + 1]); set_present(true); }
        else if (kind == 3) { w[sl.word] = 3; set_present(true); }  // "other" scalar kinds -> emptyResponse when first
        else set_present(false);
        break;
      default: ok = false;
    }
    if (ok) { T.touch(row); n_ok++; } else n_skip++;
  }
  if (applied) *applied = n_ok;
  if (skipped) *skipped = n_skip;
}

// Incremental flush: a handful of touched rows are packed [row index][row words] on the host, copied once
// and scattered into the table by this kernel — instead of re-uploading the span between the first and
// the last touched row.
__global__ void scatter_rows_kernel(uint64_t *rows, int row_words, const uint32_t *idx, const uint64_t *packed, int n) {
  const int r = blockIdx.x;
  if (r >= n) return;
  uint64_t *dst = rows + (size_t)idx[r] * row_words;
  const uint64_t *src = packed + (size_t)r * row_words;
  for (int w = threadIdx.x; w < row_words; w += blockDim.x) dst[w] = src[w];
}
template <class V>
__global__ void scatter_side_kernel(V *side, int dim, const uint32_t *idx, const V *packed, int n) {
  const int r = blockIdx.x;
  if (r >= n) return;
  V *dst = side + (size_t)idx[r] * dim;
  const V *src = packed + (size_t)r * dim;
  for (int w = threadIdx.x; w < dim; w += blockDim.x) dst[w] = src[w];
}

template <class T> static void ensure_dev(T *&d, size_t &cap, size_t need, size_t keep, int64_t &bytes) {
  if (need <= cap) return;
  size_t ncap = std::max(need, cap * 2);
  T *nd = nullptr;
  MR_CUDA_CHECK(cudaMalloc((void **)&nd, ncap * sizeof(T)));
  if (d && keep) MR_CUDA_CHECK(cudaMemcpy(nd, d, keep * sizeof(T), cudaMemcpyDeviceToDevice));
  if (d) cudaFree(d);
  bytes += (int64_t)(ncap - cap) * (int64_t)sizeof(T);
  d = nd;
  cap = ncap;
}

void StateStore::flush() {
  std::unique_lock<std::shared_mutex> g(mu);
  MR_CUDA_CHECK(cudaSetDevice(device));
  MR_CUDA_CHECK(cudaDeviceSynchronize());  // no kernel may be reading while rows move
  for (auto &T : tables) {
    if (T.map_dirty) {
      size_t cap0 = T.d_cap, cap1 = T.d_cap;
      ensure_dev(T.d_keys, cap0, T.keys.size(), 0, device_bytes);
      ensure_dev(T.d_vals, cap1, T.keys.size(), 0, device_bytes);
      T.d_cap = cap0;
      MR_CUDA_CHECK(cudaMemcpy(T.d_keys, T.keys.data(), T.keys.size() * 8, cudaMemcpyHostToDevice));
      MR_CUDA_CHECK(cudaMemcpy(T.d_vals, T.vals.data(), T.vals.size() * 4, cudaMemcpyHostToDevice));
      T.d_keys_n = T.keys.size();
      T.map_dirty = false;
    }
    if (T.dirty_lo < T.dirty_hi) {
      const size_t had = T.d_rows_cap;
      ensure_dev(T.d_rows, T.d_rows_cap, T.rows.size(), 0, device_bytes);
      bool realloc_any = T.d_rows_cap != had;
      for (size_t s = 0; s < T.sides.size(); s++) {
        const int dim = schema.sides[T.side_ids[s]].dim;
        if (T.sides[s].size() < T.n_rows * (size_t)dim) T.sides[s].resize(T.n_rows * (size_t)dim, 0.0);
        if (T.side_bs[s].size() < T.n_rows) T.side_bs[s].resize(T.n_rows, 0.0);
        // binary32 on the device while every element round-trips (and rows stay 16-byte aligned for the kernel's copies)
        const uint8_t mode = (T.side_is_f32[s] && dim % 4 == 0) ? 1 : 2;
        if (mode != T.side_dev_mode[s]) {  // first upload, or an element that needs all 53 bits arrived: switch representation
          realloc_any = true;
          if (mode == 2 && T.d_sides_f32[s]) { cudaFree(T.d_sides_f32[s]); T.d_sides_f32[s] = nullptr; device_bytes -= (int64_t)T.d_sides_f32_cap[s] * 4; T.d_sides_f32_cap[s] = 0; }
          T.side_dev_mode[s] = mode;
        }
        const size_t had_s = mode == 1 ? T.d_sides_f32_cap[s] : T.d_sides_cap[s], had_b = T.d_side_bs_cap[s];
        if (mode == 1) ensure_dev(T.d_sides_f32[s], T.d_sides_f32_cap[s], T.sides[s].size(), 0, device_bytes);
        else ensure_dev(T.d_sides[s], T.d_sides_cap[s], T.sides[s].size(), 0, device_bytes);
        ensure_dev(T.d_side_bs[s], T.d_side_bs_cap[s], T.side_bs[s].size(), 0, device_bytes);
        realloc_any |= (mode == 1 ? T.d_sides_f32_cap[s] : T.d_sides_cap[s]) != had_s || T.d_side_bs_cap[s] != had_b;
      }
      const size_t n_dirty = T.dirty_rows.size();
      const bool sparse = !realloc_any && n_dirty * 8 < (T.dirty_hi - T.dirty_lo);  // < 1/8 of the span touched
      if (sparse) {
        // scatter path: [idx][packed rows] in one staging blob per table (and per side array)
        uint32_t *d_idx = nullptr;
        MR_CUDA_CHECK(cudaMalloc((void **)&d_idx, n_dirty * 4));
        MR_CUDA_CHECK(cudaMemcpy(d_idx, T.dirty_rows.data(), n_dirty * 4, cudaMemcpyHostToDevice));
        {
          std::vector<uint64_t> packed(n_dirty * (size_t)T.row_words);
          for (size_t k = 0; k < n_dirty; k++)
            memcpy(packed.data() + k * T.row_words, T.rows.data() + (size_t)T.dirty_rows[k] * T.row_words, (size_t)T.row_words * 8);
          uint64_t *d_packed = nullptr;
          MR_CUDA_CHECK(cudaMalloc((void **)&d_packed, packed.size() * 8));
          MR_CUDA_CHECK(cudaMemcpy(d_packed, packed.data(), packed.size() * 8, cudaMemcpyHostToDevice));
          scatter_rows_kernel<<<(unsigned)n_dirty, 64>>>(T.d_rows, T.row_words, d_idx, d_packed, (int)n_dirty);
          MR_CUDA_CHECK(cudaGetLastError());
          MR_CUDA_CHECK(cudaDeviceSynchronize());
          cudaFree(d_packed);
        }
        for (size_t s = 0; s < T.sides.size(); s++) {
          const int dim = schema.sides[T.side_ids[s]].dim;
          auto scatter = [&](auto *d_side, auto zero, int width, auto fetch) {
            using V = decltype(zero);
            std::vector<V> packed(n_dirty * (size_t)width);
            for (size_t k = 0; k < n_dirty; k++)
              for (int w = 0; w < width; w++) packed[k * width + w] = fetch((size_t)T.dirty_rows[k], w);
            V *d_packed = nullptr;
            MR_CUDA_CHECK(cudaMalloc((void **)&d_packed, packed.size() * sizeof(V)));
            MR_CUDA_CHECK(cudaMemcpy(d_packed, packed.data(), packed.size() * sizeof(V), cudaMemcpyHostToDevice));
            scatter_side_kernel<V><<<(unsigned)n_dirty, 128>>>(d_side, width, d_idx, d_packed, (int)n_dirty);
            MR_CUDA_CHECK(cudaGetLastError());
            MR_CUDA_CHECK(cudaDeviceSynchronize());
            cudaFree(d_packed);
          };
          const std::vector<double> &S = T.sides[s], &BS = T.side_bs[s];
          if (T.side_dev_mode[s] == 1) scatter(T.d_sides_f32[s], 0.0f, dim, [&](size_t r, int w) { return (float)S[r * dim + w]; });
          else scatter(T.d_sides[s], 0.0, dim, [&](size_t r, int w) { return S[r * dim + w]; });
          scatter(T.d_side_bs[s], 0.0, 1, [&](size_t r, int) { return BS[r]; });
        }
        cudaFree(d_idx);
      } else {
        size_t lo = T.dirty_lo, hi = T.dirty_hi;
        if (realloc_any) { lo = 0; hi = T.n_rows; }  // reallocated: upload everything
        MR_CUDA_CHECK(cudaMemcpy(T.d_rows + lo * T.row_words, T.rows.data() + lo * T.row_words,
                                 (hi - lo) * T.row_words * 8, cudaMemcpyHostToDevice));
        for (size_t s = 0; s < T.sides.size(); s++) {
          const int dim = schema.sides[T.side_ids[s]].dim;
          if (T.side_dev_mode[s] == 1) {
            std::vector<float> f((hi - lo) * dim);
            const double *src = T.sides[s].data() + lo * dim;
            for (size_t k = 0; k < f.size(); k++) f[k] = (float)src[k];
            MR_CUDA_CHECK(cudaMemcpy(T.d_sides_f32[s] + lo * dim, f.data(), f.size() * 4, cudaMemcpyHostToDevice));
          } else {
            MR_CUDA_CHECK(cudaMemcpy(T.d_sides[s] + lo * dim, T.sides[s].data() + lo * dim, (hi - lo) * dim * 8,
                                     cudaMemcpyHostToDevice));
          }
          MR_CUDA_CHECK(cudaMemcpy(T.d_side_bs[s] + lo, T.side_bs[s].data() + lo, (hi - lo) * 8, cudaMemcpyHostToDevice));
        }
      }
      if (&T == &tables[SC_ITEM]) {
        ItemChange ch;
        ch.epoch = ++item_epoch;
        ch.all = realloc_any || n_dirty * 4 > T.n_rows;
        if (!ch.all) ch.rows = T.dirty_rows;
        item_log.push_back(std::move(ch));
        if (item_log.size() > kItemLogMax) item_log.pop_front();
      }
      for (uint32_t r : T.dirty_rows) T.row_dirty[r] = 0;
      T.dirty_rows.clear();
      T.dirty_lo = SIZE_MAX;
      T.dirty_hi = 0;
    }
    if (T.pool_dirty_lo < T.pool_dirty_hi && T.d_pool) {
      MR_CUDA_CHECK(cudaMemcpy(T.d_pool + T.pool_dirty_lo, T.pool.data() + T.pool_dirty_lo,
                               (T.pool_dirty_hi - T.pool_dirty_lo) * 8, cudaMemcpyHostToDevice));
    }
    T.pool_dirty_lo = SIZE_MAX;
    T.pool_dirty_hi = 0;
    T.d_n_rows = T.n_rows;
    if (T.pool_uploaded < T.pool.size()) {
      const size_t had = T.d_pool_cap;
      ensure_dev(T.d_pool, T.d_pool_cap, T.pool.size(), T.pool_uploaded, device_bytes);
      (void)had;
      MR_CUDA_CHECK(cudaMemcpy(T.d_pool + T.pool_uploaded, T.pool.data() + T.pool_uploaded,
                               (T.pool.size() - T.pool_uploaded) * 8, cudaMemcpyHostToDevice));
      T.pool_uploaded = T.pool.size();
    }
  }
}

DState StateStore::view() const {
  DState v;
  memset(&v, 0, sizeof v);
  for (int t = 0; t < SC_N_TABLES; t++) {
    const HostTable &T = tables[t];
    v.t[t].keys = T.d_keys;
    v.t[t].vals = T.d_vals;
    // the snapshot of the last flush: upserts that are still pending must not change what kernels index
    v.t[t].mask = (T.d_keys && T.d_keys_n) ? (uint32_t)(T.d_keys_n - 1) : 0;
    v.t[t].n_rows = T.d_rows ? (uint32_t)T.d_n_rows : 0;
    v.t[t].rows = T.d_rows;
    v.t[t].pool = T.d_pool;
    v.t[t].row_words = T.row_words;
    for (size_t s = 0; s < T.side_ids.size(); s++) {
      v.side[T.side_ids[s]] = T.side_dev_mode[s] == 2 ? T.d_sides[s] : nullptr;
      v.side_f32[T.side_ids[s]] = T.side_dev_mode[s] == 1 ? T.d_sides_f32[s] : nullptr;
      v.side_bs[T.side_ids[s]] = T.d_side_bs[s];
      v.side_dim[T.side_ids[s]] = schema.sides[T.side_ids[s]].dim;
    }
  }
  return v;
}

// ------------------------------------------------------------------ write path
// Wire format (little-endian), records back to back:
//   u16 name_len, name | u8 scope | scope payload (as in upsert) | u8 op | i64 ts (epoch millis) | payload
//   op 0 Put               payload = u8 kind + value exactly as an upsert record's kind/payload
//   op 1 Increment         payload = i64 inc                      (MemCounter.put)
//   op 2 PeriodicIncrement payload = i64 inc                      (MemPeriodicCounter.put + fromMap)
//   op 3 Append            payload = u64 hash of the appended SString (MemBoundedList.put)
void StateStore::apply_writes(const uint8_t *buf, size_t len, int64_t *applied, int64_t *skipped) {
  int64_t n_ok = 0, n_skip = 0;
  Reader r{buf, buf + len};
  std::vector<uint8_t> put_rec;
  while (r.p < r.e) {
    const uint8_t *rec_begin = r.p;
    const uint16_t nl = r.get<uint16_t>();
    std::string name((const char *)r.bytes(nl), nl);
    const uint8_t scope = r.get<uint8_t>();
    uint64_t id0 = 0, id1 = 0;
    switch (scope) {
      case SC_GLOBAL: break;
      case SC_ITEM: case SC_USER: case SC_SESSION: case SC_RANKING: case SC_FIELD: id0 = r.get<uint64_t>(); break;
      case SC_IRF: id0 = r.get<uint64_t>(); id1 = r.get<uint64_t>(); break;
      default: fail(MR_ERR_PARSE, "write record: bad scope tag %d", (int)scope);
    }
    const uint8_t *after_scope = r.p;
    const uint8_t op = r.get<uint8_t>();
    const int64_t ts = r.get<int64_t>();
    if (op == 0) {
      // Put == KVStore.put of the scalar: re-frame as an upsert record (MemScalarFeature just stores it)
      const uint8_t *val_begin = r.p;
      const uint8_t kind = r.get<uint8_t>();
      switch (kind) {
        case 0: r.bytes(8); break;
        case 1: r.bytes(8); break;
        case 2: case 3: { const uint32_t n = r.get<uint32_t>(); r.bytes((size_t)n * 8); break; }
        case 7: r.bytes(1); break;
        default: fail(MR_ERR_PARSE, "Put of value kind %d is not a scalar", (int)kind);
      }
      put_rec.assign(rec_begin, after_scope);
      put_rec.insert(put_rec.end(), val_begin, r.p);
      int64_t a = 0, s = 0;
      upsert(put_rec.data(), put_rec.size(), &a, &s);
      n_ok += a; n_skip += s;
      continue;
    }
    int64_t inc = 0; uint64_t item = 0;
    if (op == 1 || op == 2) inc = r.get<int64_t>();
    else if (op == 3) item = r.get<uint64_t>();
    else fail(MR_ERR_PARSE, "write record: bad op %d", (int)op);
    auto it = schema.slot_by_name.find(name);
    if (it == schema.slot_by_name.end() || schema.slots[it->second].table != (int)scope) { n_skip++; continue; }
    std::unique_lock<std::shared_mutex> g(mu);
    const Slot &sl = schema.slots[it->second];
    HostTable &T = tables[sl.table];
    const uint32_t row = row_for(sl, scope, id0, id1);
    uint64_t *w = T.rows.data() + (size_t)row * T.row_words;
    const RawKey rk{((uint64_t)sl.table << 56) ^ ((uint64_t)it->second << 40) ^ (uint64_t)row};
    auto set_present = [&] { w[sl.bit >> 6] |= 1ull << (sl.bit & 63); };
    if (op == 1 && sl.kind == SK_COUNTER) {
      int64_t cur = (w[sl.bit >> 6] >> (sl.bit & 63)) & 1 ? (int64_t)w[sl.word] : 0;
      cur += inc;
      memcpy(&w[sl.word], &cur, 8);
      set_present();
    } else if (op == 2 && sl.kind == SK_PCOUNTER) {
      // MemPeriodicCounter.put: bucket = ts.toStartOfPeriod(period) = floor(ts.toDouble / period) * period
      const int64_t bucket = (int64_t)std::floor((double)ts / (double)sl.period_ms) * sl.period_ms;
      const bool was_present = (w[sl.bit >> 6] >> (sl.bit & 63)) & 1;
      if (was_present && buckets.find(rk) == buckets.end())
        fail(MR_ERR_UNSUPPORTED,
             "periodic counter '%s' of this key was loaded as a refreshed value (mr_state_upsert / mr_state_load_feature_values): "
             "a PeriodicCounterValue carries window sums, not the day buckets behind them, so a PeriodicIncrement cannot continue "
             "it without resetting the windows — feed a slot either refreshed values or raw writes, not both",
             name.c_str());
      auto &m = buckets[rk];
      m[bucket] += inc;
      // PeriodicCounterFeature.fromMap: windows anchored at the LAST bucket,
      // [last - period*start, last - period*0 + period], both inclusive
      const int64_t last = m.rbegin()->first;
      for (size_t k = 0; k < sl.ranges.size(); k++) {
        const int64_t start = last - sl.period_ms * sl.ranges[k], end = last + sl.period_ms;
        int64_t sum = 0;
        for (auto bi = m.lower_bound(start); bi != m.end() && bi->first <= end; ++bi) sum += bi->second;
        memcpy(&w[sl.word + k], &sum, 8);
      }
      set_present();
    } else if (op == 3 && sl.kind == SK_BLIST) {
      // MemBoundedList.put: the first write is stored untrimmed; later ones prepend, drop entries
      // older than ts - duration, keep `count`
      auto lit = lists.find(rk);
      if (lit == lists.end() && ((w[sl.bit >> 6] >> (sl.bit & 63)) & 1))
        fail(MR_ERR_UNSUPPORTED,
             "bounded list '%s' of this key was loaded as a refreshed value (mr_state_upsert / mr_state_load_feature_values): "
             "its entries' timestamps are not part of the record, so an Append cannot age them — feed a slot either refreshed "
             "values or raw writes, not both",
             name.c_str());
      if (lit == lists.end()) {
        lists[rk].push_front({ts, item});
        lit = lists.find(rk);
      } else {
        auto &L = lit->second;
        L.push_front({ts, item});
        std::deque<std::pair<int64_t, uint64_t>> kept;
        for (auto &tv : L) {
          if (tv.first >= ts - sl.list_duration_ms) kept.push_back(tv);
          if ((int)kept.size() >= sl.list_count) break;
        }
        L.swap(kept);
      }
      auto &L = lit->second;
      // the list lives in a fixed region of max(count, first-write size) pool entries that is
      // rewritten in place, so a long event stream does not grow the pool
      const size_t cap = std::max<size_t>((size_t)std::max(sl.list_count, 1), L.size());
      auto reg = list_region.find(rk);
      if (reg == list_region.end()) {
        if (T.pool.size() + cap > (size_t)UINT32_MAX) fail(MR_ERR_UNSUPPORTED, "state pool is full");
        reg = list_region.emplace(rk, (uint32_t)T.pool.size()).first;
        T.pool.resize(T.pool.size() + cap, 0);
      }
      const uint32_t off = reg->second;
      size_t k = 0;
      for (auto &tv : L) T.pool[off + k++] = tv.second;
      if (off < T.pool_uploaded) {
        T.pool_dirty_lo = std::min<size_t>(T.pool_dirty_lo, off);
        T.pool_dirty_hi = std::max<size_t>(T.pool_dirty_hi, std::min<size_t>(off + cap, T.pool_uploaded));
      }
      w[sl.word] = (uint64_t)off | ((uint64_t)L.size() << 32);
      set_present();
    } else {
      n_skip++;
      continue;
    }
    T.touch(row);
    n_ok++;
  }
  if (applied) *applied = n_ok;
  if (skipped) *skipped = n_skip;
}

}  // namespace mr
