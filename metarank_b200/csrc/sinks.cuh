// Score sinks: the device side of the mega-request all-gather (rank_api.cu mr_group_rank).
// A scorer's final store goes to the score buffer of every GPU of the group through peer memory (NVLink /
// NVSwitch), and the CTA that finishes last raises this member's flag on every peer — the all-gather is the
// scorer's epilogue, no collective kernel follows it.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "gbdt_kernels.cuh"

namespace mr {

__device__ __forceinline__ void store_score(double *out, const ScoreSinks &s, int item, double v) {
  if (out) out[item] = v;
  for (int g = 0; g < s.n_peer; g++) s.peer[g][s.item_base + item] = v;
}

// Called by EVERY thread of every CTA after its stores.  Release pattern: stores -> system fence -> CTA barrier ->
// one device-scope arrival per CTA; the last arriver fences again and writes the flags with release semantics,
// so a peer that acquires the flag sees every score of this launch.
__device__ __forceinline__ void publish_when_last(const ScoreSinks &s) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y;
    if (atomicAdd(s.done, 1u) == total - 1u) {
      *s.done = 0u;  // the next launch on this stream starts from zero
      __threadfence_system();
      for (int g = 0; g < s.n_peer; g++)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(s.flag[g]), "r"(s.seq) : "memory");
    }
  }
}

}  // namespace mr
