// shard_p2p.cu -- the exchange of the row-sharded BPRMF path (BASELINE config 5) done by the kernels themselves over
// peer-mapped memory (torch.distributed._symmetric_memory buffers; NVLink loads / stores), instead of NCCL collectives
// and tensor glue between kernels.  The reference is single-device: nothing here replaces reference code;
// the contract is SURVEY.md 8(e) (row-range shards, score routing).
//
//   b2r_route_count / b2r_route_scatter   a rank's (sample, candidate) ids -> their owners.  A STABLE partition (pairs of
//       one destination stay ordered by sample, then candidate -- the owner's per-sample reductions depend on it and it
//       makes every sum's order fixed): per-sample counts, per-CTA totals, and a scatter that writes each pair's local row
//       and query index STRAIGHT INTO THE OWNER'S receive arrays through its peer pointer, pads the unused tail of each
//       destination's region with -1, and remembers where every pair went (slot_of) for the scores that come back.
//   b2r_serve_rows       an owner answers row requests by storing the rows into EVERY rank's copy of a replicated block
//       (the user vectors every owner scores against) -- the all-gather folded into the gather.
//   b2r_scatter_f32_to_peers / b2r_scatter_rows_to_peers   values / rows to (owner, slot) through peer pointers
//       (the loss gradient g to the item-row owners, the dQ rows to the user-row owners).
//   b2r_sum_rows_from_peers   out[r] = sum over ranks, in rank order, of peer[rank][offset + r]: the reduce-scatter of the
//       owners' dQ partials as P2P loads with a fixed summation order (deterministic).
// Ordering between ranks is by signal-pad barriers issued from the host side (shard.py); no kernel here spins on remote
// memory.
#include "common.cuh"

namespace b2r {

constexpr int kMaxW = 16;

struct PeerPtrs {
    void* p[kMaxW];
};

// ---- routing -------------------------------------------------------------------------------------------------------
// counts[b * W + o] = number of candidates of sample b owned by rank o; cta_tot[blk * W + o] = the CTA's totals.
// One warp per sample; owner = id / rows_per (ids are range-checked: out-of-range -> row 0 of rank 0, counted).
__global__ void __launch_bounds__(256)
k_route_count(const int64_t* __restrict__ ids, int B, int C, int W, int64_t rows_per, int64_t n_rows,
              int* __restrict__ counts, int* __restrict__ cta_tot, int32_t* err_flag) {
    __shared__ int tot[kMaxW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < kMaxW) tot[threadIdx.x] = 0;
    __syncthreads();
    const int b = blockIdx.x * 8 + warp;
    if (b < B) {
        int mine[kMaxW];
#pragma unroll
        for (int o = 0; o < kMaxW; ++o) mine[o] = 0;
        for (int c = lane; c < C; c += 32) {
            const int64_t id = checked_id(ids[(int64_t)b * C + c], n_rows, err_flag);
            const int o = (int)(id / rows_per);
#pragma unroll
            for (int k = 0; k < kMaxW; ++k) mine[k] += (k == o) ? 1 : 0;
        }
#pragma unroll
        for (int o = 0; o < kMaxW; ++o) {
            if (o < W) {
                const int v = __reduce_add_sync(B2R_FULL_MASK, mine[o]);
                if (lane == 0) {
                    counts[b * W + o] = v;
                    atomicAdd(&tot[o], v);
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < W) cta_tot[blockIdx.x * W + threadIdx.x] = tot[threadIdx.x];
}

// Scatter with the slots the counts imply.  dst_rows[o] / dst_q[o] point at the REGION OF THIS RANK inside rank o's
// receive arrays (cap entries each); query index = q_base + b.  slot_of[b * C + c] = o * cap + slot.
// The last CTA of the grid order (not of time) pads: every CTA pads nothing; a separate tail pass does (below).
__global__ void __launch_bounds__(256)
k_route_scatter(const int64_t* __restrict__ ids, int B, int C, int W, int64_t rows_per, int64_t n_rows,
                const int* __restrict__ counts, const int* __restrict__ cta_tot, PeerPtrs dst_rows, PeerPtrs dst_q,
                int64_t q_base, int cap, int* __restrict__ slot_of, int* __restrict__ dest_total,
                int* __restrict__ overflow) {
    __shared__ int base[kMaxW];              // slots taken by the CTAs before this one
    __shared__ int wcnt[8][kMaxW];           // this CTA's per-warp (per-sample) counts
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < kMaxW) base[threadIdx.x] = 0;
    __syncthreads();
    // prefix over the preceding CTAs' totals: W columns, blockIdx.x rows
    for (int i = threadIdx.x; i < (int)blockIdx.x * W; i += 256) atomicAdd(&base[i % W], cta_tot[i]);
    const int b = blockIdx.x * 8 + warp;
    if (lane < W) wcnt[warp][lane] = (b < B) ? counts[b * W + lane] : 0;
    __syncthreads();
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x < W) {       // grand totals (for the tail padding and the owners)
        int t = base[threadIdx.x];
        for (int w2 = 0; w2 < 8; ++w2) t += wcnt[w2][threadIdx.x];
        dest_total[threadIdx.x] = t;
        if (t > cap) atomicAdd(overflow, 1);
    }
    if (b >= B) return;
    int run[kMaxW];                           // next free slot per destination for this sample
#pragma unroll
    for (int o = 0; o < kMaxW; ++o) {
        int s = 0;
        if (o < W) {
            s = base[o];
            for (int w2 = 0; w2 < warp; ++w2) s += wcnt[w2][o];
        }
        run[o] = s;
    }
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        int64_t id = 0;
        int o = -1;
        if (c < C) {
            id = ids[(int64_t)b * C + c];
            if (id < 0 || id >= n_rows) id = 0;
            o = (int)(id / rows_per);
        }
#pragma unroll
        for (int k = 0; k < kMaxW; ++k) {
            if (k < W) {
                const unsigned m = __ballot_sync(B2R_FULL_MASK, o == k);
                if (o == k) {
                    const int slot = run[k] + __popc(m & ((1u << lane) - 1u));
                    if (slot < cap) {
                        reinterpret_cast<int64_t*>(dst_rows.p[k])[slot] = id - (int64_t)k * rows_per;
                        reinterpret_cast<int64_t*>(dst_q.p[k])[slot] = q_base + b;
                    }
                    slot_of[(int64_t)b * C + c] = k * cap + (slot < cap ? slot : cap - 1);
                }
                run[k] += __popc(m);
            }
        }
    }
}

// unused tail of every destination region: rows = -1 (the owners' kernels skip such slots)
__global__ void __launch_bounds__(256)
k_route_pad(PeerPtrs dst_rows, const int* __restrict__ dest_total, int W, int cap) {
    for (int o = 0; o < W; ++o) {
        const int t = min(dest_total[o], cap);
        int64_t* r = reinterpret_cast<int64_t*>(dst_rows.p[o]);
        for (int i = t + blockIdx.x * 256 + threadIdx.x; i < cap; i += gridDim.x * 256) r[i] = -1;
    }
}

// ---- owners answer row requests into every rank's replicated block ---------------------------------------------------
// req_rows[e] (local row, < 0: unused), req_q[e] (row of the replicated block to fill); dst.p[r] = rank r's block.
template <int LPR>
__global__ void __launch_bounds__(256)
k_serve_rows(const float* __restrict__ T, int64_t n_t, const int64_t* __restrict__ req_rows,
             const int64_t* __restrict__ req_q, int64_t n, PeerPtrs dst, int W, int32_t* err_flag) {
    constexpr int D = LPR * 4;
    const int sub = threadIdx.x % LPR, grp = threadIdx.x / LPR;
    for (int64_t e = (int64_t)blockIdx.x * (256 / LPR) + grp; e < n; e += (int64_t)gridDim.x * (256 / LPR)) {
        int64_t r = req_rows[e];
        if (r < 0) continue;
        r = checked_id(r, n_t, sub == 0 ? err_flag : nullptr);
        const float4 v = ld4(T + r * D + sub * 4);
        const int64_t q = req_q[e];
        for (int k = 0; k < W; ++k) st4(reinterpret_cast<float*>(dst.p[k]) + q * D + sub * 4, v);
    }
}

// ---- values / rows to (owner, slot) -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_scatter_f32_to_peers(const float* __restrict__ val, const int* __restrict__ slot_of, int64_t n, PeerPtrs dst, int cap,
                       float scale) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int so = slot_of[i];
        reinterpret_cast<float*>(dst.p[so / cap])[so % cap] = val[i] * scale;
    }
}

template <int LPR>
__global__ void __launch_bounds__(256)
k_scatter_rows_to_peers(const float* __restrict__ src, const int* __restrict__ slot_of, int64_t n, PeerPtrs dst, int cap) {
    constexpr int D = LPR * 4;
    const int sub = threadIdx.x % LPR, grp = threadIdx.x / LPR;
    for (int64_t i = (int64_t)blockIdx.x * (256 / LPR) + grp; i < n; i += (int64_t)gridDim.x * (256 / LPR)) {
        const int so = slot_of[i];
        st4(reinterpret_cast<float*>(dst.p[so / cap]) + (int64_t)(so % cap) * D + sub * 4, ld4(src + i * D + sub * 4));
    }
}

// ---- out[i] = sum_k peer[k][offset + i], k ascending (float4 granularity) ---------------------------------------------
__global__ void __launch_bounds__(256)
k_sum_rows_from_peers(PeerPtrs src, int W, int64_t offset4, float4* __restrict__ out, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        float4 acc = reinterpret_cast<const float4*>(src.p[0])[offset4 + i];
        for (int k = 1; k < W; ++k) {
            const float4 v = reinterpret_cast<const float4*>(src.p[k])[offset4 + i];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        out[i] = acc;
    }
}

static int fill_ptrs(PeerPtrs* out, const void* const* tab, int W) {
    if (W < 1 || W > kMaxW || tab == nullptr) return -1;
    for (int k = 0; k < kMaxW; ++k) out->p[k] = k < W ? const_cast<void*>(tab[k]) : nullptr;
    return 0;
}

}  // namespace b2r

using namespace b2r;

extern "C" size_t b2r_route_workspace_bytes(int B, int W) {
    if (B <= 0 || W < 1 || W > kMaxW) return 0;
    const size_t ctas = (size_t)(B + 7) / 8;
    return align_up((size_t)B * W * 4, 256) + align_up(ctas * W * 4, 256) + 256;
}

// Route the ids [B, C] of one table: counts + stable scatter + tail padding.  peer_rows / peer_q: HOST arrays of W device
// pointers (rank o's receive region for THIS rank, cap int64 each).  slot_of [B*C] int32 out; dest_total [W] int32 out
// (device); overflow: device counter bumped when a destination needed more than cap slots.
extern "C" int b2r_route_ids(const int64_t* ids, int B, int C, int W, int64_t rows_per, int64_t n_rows,
                             const void* const* peer_rows, const void* const* peer_q, int64_t q_base, int cap,
                             int* slot_of, int* dest_total, int* overflow, void* ws, size_t ws_bytes, int32_t* err_flag,
                             b2r_stream_t stream) {
    B2R_REQUIRE(ids && slot_of && dest_total && overflow && ws, B2R_E_BADARG, "b2r_route_ids: null pointer");
    B2R_REQUIRE(B > 0 && C > 0 && cap > 0 && rows_per > 0, B2R_E_BADARG, "b2r_route_ids: bad sizes");
    PeerPtrs pr, pq;
    B2R_REQUIRE(fill_ptrs(&pr, peer_rows, W) == 0 && fill_ptrs(&pq, peer_q, W) == 0, B2R_E_BADARG,
                "b2r_route_ids: W=%d (1..%d) or null pointer table", W, kMaxW);
    B2R_REQUIRE((n_rows + rows_per - 1) / rows_per <= W, B2R_E_BADARG, "b2r_route_ids: rows_per too small for W");
    B2R_REQUIRE(ws_bytes >= b2r_route_workspace_bytes(B, W), B2R_E_WORKSPACE, "b2r_route_ids: workspace too small");
    cudaStream_t s = as_stream(stream);
    const int ctas = (B + 7) / 8;
    int* counts = static_cast<int*>(ws);
    int* cta_tot = reinterpret_cast<int*>(static_cast<char*>(ws) + align_up((size_t)B * W * 4, 256));
    k_route_count<<<ctas, 256, 0, s>>>(ids, B, C, W, rows_per, n_rows, counts, cta_tot, err_flag);
    B2R_LAUNCH_OK("k_route_count");
    k_route_scatter<<<ctas, 256, 0, s>>>(ids, B, C, W, rows_per, n_rows, counts, cta_tot, pr, pq, q_base, cap, slot_of,
                                         dest_total, overflow);
    B2R_LAUNCH_OK("k_route_scatter");
    int pad_grid = (cap / 4 + 255) / 256;
    if (pad_grid > sm_count() * 2) pad_grid = sm_count() * 2;
    if (pad_grid < 1) pad_grid = 1;
    k_route_pad<<<pad_grid, 256, 0, s>>>(pr, dest_total, W, cap);
    B2R_LAUNCH_OK("k_route_pad");
    return 0;
}

extern "C" int b2r_serve_rows(const float* T, int64_t n_t, const int64_t* req_rows, const int64_t* req_q, int64_t n,
                              const void* const* peer_dst, int W, int d, int32_t* err_flag, b2r_stream_t stream) {
    B2R_REQUIRE(T && req_rows && req_q, B2R_E_BADARG, "b2r_serve_rows: null pointer");
    B2R_REQUIRE(d == 32 || d == 64 || d == 128, B2R_E_UNSUPPORTED, "b2r_serve_rows: d=%d", d);
    PeerPtrs pd;
    B2R_REQUIRE(fill_ptrs(&pd, peer_dst, W) == 0, B2R_E_BADARG, "b2r_serve_rows: pointer table");
    if (n <= 0) return 0;
    cudaStream_t s = as_stream(stream);
    const int gpc = 256 / (d / 4);
    int64_t grid = (n + gpc - 1) / gpc;
    if (grid > sm_count() * 16) grid = sm_count() * 16;
    if (d == 32) k_serve_rows<8><<<(int)grid, 256, 0, s>>>(T, n_t, req_rows, req_q, n, pd, W, err_flag);
    else if (d == 64) k_serve_rows<16><<<(int)grid, 256, 0, s>>>(T, n_t, req_rows, req_q, n, pd, W, err_flag);
    else k_serve_rows<32><<<(int)grid, 256, 0, s>>>(T, n_t, req_rows, req_q, n, pd, W, err_flag);
    B2R_LAUNCH_OK("k_serve_rows");
    return 0;
}

extern "C" int b2r_scatter_f32_to_peers(const float* val, const int* slot_of, int64_t n, const void* const* peer_dst, int W,
                                        int cap, float scale, b2r_stream_t stream) {
    B2R_REQUIRE(val && slot_of && cap > 0, B2R_E_BADARG, "b2r_scatter_f32_to_peers: bad argument");
    PeerPtrs pd;
    B2R_REQUIRE(fill_ptrs(&pd, peer_dst, W) == 0, B2R_E_BADARG, "b2r_scatter_f32_to_peers: pointer table");
    if (n <= 0) return 0;
    int64_t grid = (n + 255) / 256;
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    k_scatter_f32_to_peers<<<(int)grid, 256, 0, as_stream(stream)>>>(val, slot_of, n, pd, cap, scale);
    B2R_LAUNCH_OK("k_scatter_f32_to_peers");
    return 0;
}

extern "C" int b2r_scatter_rows_to_peers(const float* src, const int* slot_of, int64_t n, const void* const* peer_dst, int W,
                                         int cap, int d, b2r_stream_t stream) {
    B2R_REQUIRE(src && slot_of && cap > 0, B2R_E_BADARG, "b2r_scatter_rows_to_peers: bad argument");
    B2R_REQUIRE(d == 32 || d == 64 || d == 128, B2R_E_UNSUPPORTED, "b2r_scatter_rows_to_peers: d=%d", d);
    PeerPtrs pd;
    B2R_REQUIRE(fill_ptrs(&pd, peer_dst, W) == 0, B2R_E_BADARG, "b2r_scatter_rows_to_peers: pointer table");
    if (n <= 0) return 0;
    cudaStream_t s = as_stream(stream);
    const int gpc = 256 / (d / 4);
    int64_t grid = (n + gpc - 1) / gpc;
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    if (d == 32) k_scatter_rows_to_peers<8><<<(int)grid, 256, 0, s>>>(src, slot_of, n, pd, cap);
    else if (d == 64) k_scatter_rows_to_peers<16><<<(int)grid, 256, 0, s>>>(src, slot_of, n, pd, cap);
    else k_scatter_rows_to_peers<32><<<(int)grid, 256, 0, s>>>(src, slot_of, n, pd, cap);
    B2R_LAUNCH_OK("k_scatter_rows_to_peers");
    return 0;
}

extern "C" int b2r_sum_rows_from_peers(const void* const* peer_src, int W, int64_t offset_floats, float* out,
                                       int64_t n_floats, b2r_stream_t stream) {
    B2R_REQUIRE(out && n_floats >= 0 && n_floats % 4 == 0 && offset_floats % 4 == 0, B2R_E_BADARG,
                "b2r_sum_rows_from_peers: sizes must be multiples of 4 floats");
    PeerPtrs ps;
    B2R_REQUIRE(fill_ptrs(&ps, peer_src, W) == 0, B2R_E_BADARG, "b2r_sum_rows_from_peers: pointer table");
    if (n_floats == 0) return 0;
    const int64_t n4 = n_floats / 4;
    int64_t grid = (n4 + 255) / 256;
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    k_sum_rows_from_peers<<<(int)grid, 256, 0, as_stream(stream)>>>(ps, W, offset_floats / 4, reinterpret_cast<float4*>(out), n4);
    B2R_LAUNCH_OK("k_sum_rows_from_peers");
    return 0;
}
