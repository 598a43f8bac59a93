// plan_direct.cuh -- the index plan of a whole-step context built WITHOUT a pass of its own (bucket.cu, bprmf_flash.cu,
// bprmf_step.cu).  The fused forward kernel reads every id of the batch anyway; while it parks them in shared memory it
// also drops each (row, position) pair into the bucket of its row range: fixed-capacity bucket regions, one L2 atomic per
// pair on a line-padded cursor, pairs beyond a region's capacity (skewed batches) into a spill list.  What is left of the
// plan is ONE kernel between the forward and the update: the per-bucket sort (k_bucket_sort_direct), which also lists the
// row heads for k_apply_sorted.  No side stream, no prefetched batch, no count / scan / scatter launches.
#pragma once
#include "common.cuh"

namespace b2r {

constexpr int kPadD = 32;                 // ints per bucket cursor: one 128-byte line each (L2 atomics serialise per line)

struct DirectPlanDev {                    // what the producing kernel needs
    int* cursor;                          // [nb * kPadD]
    uint64_t* region;                     // [nb][cap] pairs (row << 32 | position)
    uint64_t* spill;                      // [spill_cap] pairs that did not fit their region
    int* counters;                        // [0] spill count, [1] big-area cursor, [2] long rows, [3] row heads
    int cap, shift, spill_cap;
};

__device__ __forceinline__ void direct_scatter(const DirectPlanDev& P, uint32_t key, uint32_t pos) {
    const int b = (int)(key >> P.shift);
    const int slot = atomicAdd(&P.cursor[(int64_t)b * kPadD], 1);
    const uint64_t pair = ((uint64_t)key << 32) | (uint64_t)pos;
    if (slot < P.cap) {
        P.region[(int64_t)b * P.cap + slot] = pair;
    } else {
        const int s = atomicAdd(&P.counters[0], 1);
        if (s < P.spill_cap) P.spill[s] = pair;
    }
}

size_t direct_workspace_bytes(int64_t n, int64_t n_rows);
int direct_workspace_init(void* ws, size_t ws_bytes, int64_t n, int64_t n_rows, cudaStream_t s);
DirectPlanDev direct_plan_dev(void* ws, int64_t n, int64_t n_rows);
// standalone producer (the prefetch form of a step context: the NEXT batch's ids on a side stream): drop the pairs of one or
// two id arrays into their plans in ONE launch
int direct_scatter_pair(const int64_t* ids_a, int64_t n_a, int64_t rows_a, void* ws_a, const int64_t* ids_b, int64_t n_b,
                        int64_t rows_b, void* ws_b, int32_t* err_flag, cudaStream_t s);
// sort every bucket of one or two plans (b may be NULL) in ONE launch and list their row heads
int direct_sort_pair(void* ws_a, int64_t n_a, int64_t rows_a, void* ws_b, int64_t n_b, int64_t rows_b, cudaStream_t s);
// b2r_bucket_apply_pair on plans built this way (job.ws = a direct workspace)
int direct_apply_pair(const b2r_apply_job* ja, const b2r_apply_job* jb, int d, int mode, const b2r_optim* opt, cudaStream_t s);

}  // namespace b2r
