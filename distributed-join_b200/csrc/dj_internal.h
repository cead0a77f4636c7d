// dj_internal.h -- host-side declarations shared between the .cu translation units.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/dj_b200.h"

namespace dj {

constexpr int kMaxPayload = 3;
constexpr int kMaxFanout  = 1024;

// Bucket function of one partition pass.
//   mode 0: (row_hash(key; seed, hash_id)) % F   -- the cuDF-compatible rank partition
//   mode 1: (local_hash(key) >> shift) & (F-1)   -- the join's private radix sub-partition
struct PassDesc {
  int mode;
  uint32_t seed;
  int hash_id;
  int shift;
  int F;     // fan-out per parent bucket (<= kMaxFanout)
  int P;     // number of parent buckets (1 for a top-level pass, <= kMaxFanout)
  int npay;  // payload columns moved with the key (1..kMaxPayload)
};

struct PassBuffers {
  const int64_t* in_key;
  const int64_t* in_pay[kMaxPayload];
  int64_t* out_key;
  int64_t* out_pay[kMaxPayload];
  int64_t nrows;
  const int64_t* d_parent_off;  // [P+1] absolute row offsets of the parents; nullptr when P == 1
  int64_t* d_child_off;         // [P*F+1] out: absolute row offsets of the child buckets
};

size_t pass_workspace_bytes(int P, int F);
int run_partition_pass(const PassDesc& desc, const PassBuffers& buf, void* d_ws, size_t ws_bytes,
                       cudaStream_t stream);

// Local join of radix-partitioned tables: bucket b of the build side is rows
// [d_build_off[b], d_build_off[b+1]) of (bk, bp), likewise for the probe side.
struct JoinBuffers {
  const int64_t* bk;
  const int64_t* bp;
  const int64_t* d_build_off;
  const int64_t* pk;
  const int64_t* pp;
  const int64_t* d_probe_off;
  int nbuckets;
  int64_t* out[4];  // build key, build payload, probe key, probe payload
  int64_t out_capacity;
  int64_t out_base;        // rows already in the output before this call (batched joins)
  int64_t* d_out_count;    // running total (device); the kernel atomically adds to it
  int* d_work_counter;     // zeroed by the caller of run_bucket_join
};
int run_bucket_join(const JoinBuffers& jb, bool swap_output_sides, cudaStream_t stream);

// Radix plan for a local join with `nbuild` build rows.
struct RadixPlan {
  int bits1, bits2;  // fan-out bits of level 1 / level 2 (0 = level not used)
  int nbuckets;
};
RadixPlan make_radix_plan(int64_t nbuild);

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Simple bump allocator over a caller-provided device workspace.
struct Arena {
  char* base;
  size_t size;
  size_t used = 0;
  Arena(void* p, size_t n) : base((char*)p), size(n) {}
  template <typename T>
  T* take(size_t count)
  {
    size_t off = align_up(used, 256);
    size_t end = off + count * sizeof(T);
    if (end > size) return nullptr;
    used = end;
    return (T*)(base + off);
  }
};

}  // namespace dj
