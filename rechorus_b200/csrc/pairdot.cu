// pairdot.cu -- score a flat list of (query row, table row) pairs: out[e] = <Q[qidx[e]], T[rows[e]]>.
// Used by the row-sharded table path (config 5): a shard owner receives the (sample, local row) pairs routed to it
// by the all-to-all and scores them against the all-gathered user vectors, returning 4-byte scores instead of
// 4d-byte rows.  Same lane mapping as rowdot.cu (d/4 lanes per row, RCH pairs in flight per lane group); the
// query rows (world*B x d, a few MB) stay in L2.  rows[e] < 0 marks an unused slot of the fixed-capacity
// exchange buffer: its score is 0.
#include "common.cuh"

namespace b2r {

template <int LPR, int RCH>
__global__ void __launch_bounds__(256)
k_pairdot_fwd(const float* __restrict__ Q, const int64_t* __restrict__ qidx, int64_t n_q, const float* __restrict__ T,
              const int64_t* __restrict__ rows, int64_t n_t, float* __restrict__ out, int64_t n, int32_t* err_flag) {
    static_assert(RCH <= LPR, "ids of a chunk are loaded one per lane");
    constexpr int D = LPR * 4;
    constexpr int GPC = 256 / LPR;
    constexpr int GPW = 32 / LPR;
    const int sub = threadIdx.x % LPR;
    const int grp = threadIdx.x / LPR;
    const int64_t nchunks = (n + RCH - 1) / RCH;
    const int64_t warp_first = (int64_t)blockIdx.x * GPC + (grp / GPW) * GPW;
    for (int64_t wbase = warp_first; wbase < nchunks; wbase += (int64_t)gridDim.x * GPC) {
        const int64_t ch = wbase + (grp % GPW);
        const int64_t e0 = ch * RCH;
        const int nr = (ch < nchunks) ? (int)min((int64_t)RCH, n - e0) : 0;
        int64_t my_row = -1, my_q = 0;
        if (sub < nr) {
            my_row = rows[e0 + sub];
            if (my_row >= 0) {
                my_row = checked_id(my_row, n_t, err_flag);
                my_q = checked_id(qidx[e0 + sub], n_q, err_flag);
            }
        }
        float4 r[RCH], q[RCH];
#pragma unroll
        for (int k = 0; k < RCH; ++k) {
            const int64_t rk = __shfl_sync(B2R_FULL_MASK, my_row, k, LPR);
            const int64_t qk = __shfl_sync(B2R_FULL_MASK, my_q, k, LPR);
            if (k < nr && rk >= 0) {
                r[k] = ld_row4(T + rk * D + sub * 4);
                q[k] = ld4(Q + qk * D + sub * 4);
            } else {
                r[k] = q[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        float mine = 0.f;
#pragma unroll
        for (int k = 0; k < RCH; ++k) {
            const float v = group_sum<LPR>(dot4(q[k], r[k]));
            if (sub == k) mine = v;
        }
        if (sub < nr) out[e0 + sub] = mine;
    }
}

__global__ void __launch_bounds__(256)
k_pairdot_fwd_generic(const float* __restrict__ Q, const int64_t* __restrict__ qidx, int64_t n_q,
                      const float* __restrict__ T, const int64_t* __restrict__ rows, int64_t n_t,
                      float* __restrict__ out, int64_t n, int d, int32_t* err_flag) {
    const int lane = threadIdx.x & 31;
    const int d4 = d >> 2;
    for (int64_t e = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); e < n; e += (int64_t)gridDim.x * 8) {
        int64_t rk = rows[e];
        float v = 0.f;
        if (rk >= 0) {
            rk = checked_id(rk, n_t, lane == 0 ? err_flag : nullptr);
            const int64_t qk = checked_id(qidx[e], n_q, lane == 0 ? err_flag : nullptr);
            for (int k = lane; k < d4; k += 32) v += dot4(ld4(Q + qk * d + k * 4), ld_row4(T + rk * d + k * 4));
        }
        v = warp_sum(v);
        if (lane == 0) out[e] = v;
    }
}

// out[key[e], :] = sum over the run of consecutive valid elements sharing key[e] of coef[e] * T[rows[e], :]
// (rows[e] < 0 = unused slot).  The shard owner's half of dQ = sum_c g * I[id]: the pairs a rank receives from one
// source arrive grouped by sample (the sender bucketed them with a stable partition), so every (source, sample) is one
// contiguous run and no sort is needed.  A lane group scans a block of LPR consecutive elements at a time (one
// element's metadata per lane, coalesced), finds the runs that START in its block with a ballot, and walks each of
// them in element order -- metadata travels by shuffle, only the table rows are loaded in the walk.  Every output row
// is written at most once (fixed summation order); rows without elements keep what the caller put there (zeros).
template <int LPR>
__global__ void __launch_bounds__(256)
k_pair_runs_sum(const int64_t* __restrict__ key, const int64_t* __restrict__ rows, const float* __restrict__ coef,
                const float* __restrict__ T, int64_t n_t, float* __restrict__ out, int64_t n_out, int64_t n) {
    constexpr int D = LPR * 4;
    constexpr int GPC = 256 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = threadIdx.x % LPR, grp = threadIdx.x / LPR;
    const int goff = lane - sub;                                           // first lane of this group in its warp
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << goff);
    const int64_t nblocks = (n + LPR - 1) / LPR;
    for (int64_t blk = (int64_t)blockIdx.x * GPC + grp; blk < nblocks; blk += (int64_t)gridDim.x * GPC) {
        const int64_t e0 = blk * LPR;
        int64_t b_row = -1, b_key = -1;
        float b_coef = 0.f;
        if (e0 + sub < n) {
            b_row = rows[e0 + sub];
            b_key = key[e0 + sub];
            b_coef = coef[e0 + sub];
        }
        // head of a run: valid, and the previous element is invalid or has another key
        int64_t p_row = __shfl_up_sync(gmask, b_row, 1, LPR), p_key = __shfl_up_sync(gmask, b_key, 1, LPR);
        if (sub == 0) {
            p_row = -1;
            if (e0 > 0) {
                p_row = rows[e0 - 1];
                p_key = key[e0 - 1];
            }
        }
        const bool is_head = b_row >= 0 && b_key >= 0 && b_key < n_out && (p_row < 0 || p_key != b_key);
        unsigned heads = (__ballot_sync(gmask, is_head) >> goff) & ((LPR == 32) ? 0xffffffffu : ((1u << LPR) - 1u));
        while (heads) {
            const int h = __ffs(heads) - 1;
            heads &= heads - 1;
            const int64_t k = __shfl_sync(gmask, b_key, h, LPR);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            int64_t c_row = b_row, c_key = b_key;                          // metadata of the block being walked
            float c_coef = b_coef;
            int start = h;
            int64_t next = e0 + LPR;
            for (;;) {
                const bool mine = (sub >= start) && c_row >= 0 && c_key == k;
                unsigned in_run = (__ballot_sync(gmask, mine) >> goff) & ((LPR == 32) ? 0xffffffffu : ((1u << LPR) - 1u));
                in_run >>= start;
                const int len = (~in_run == 0u) ? 32 - start : __ffs(~in_run) - 1;   // consecutive members from `start`
                int t = start;
                for (; t + 8 <= start + len; t += 8) {                      // 8 row loads in flight, added in order
                    float4 x[8];
                    float c[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int64_t r = __shfl_sync(gmask, c_row, t + u, LPR);
                        c[u] = __shfl_sync(gmask, c_coef, t + u, LPR);
                        x[u] = ld_row4(T + (r < n_t ? r : 0) * D + sub * 4);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) fma4(acc, c[u], x[u]);
                }
                for (; t + 4 <= start + len; t += 4) {                      // 4 row loads in flight, added in order
                    float4 x[4];
                    float c[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int64_t r = __shfl_sync(gmask, c_row, t + u, LPR);
                        c[u] = __shfl_sync(gmask, c_coef, t + u, LPR);
                        x[u] = ld_row4(T + (r < n_t ? r : 0) * D + sub * 4);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) fma4(acc, c[u], x[u]);
                }
                for (; t < start + len; ++t) {
                    const int64_t r = __shfl_sync(gmask, c_row, t, LPR);
                    const float c = __shfl_sync(gmask, c_coef, t, LPR);
                    fma4(acc, c, ld_row4(T + (r < n_t ? r : 0) * D + sub * 4));
                }
                if (start + len < LPR || next >= n) break;                 // run ended inside this block / input ended
                c_row = -1;
                c_key = -1;
                c_coef = 0.f;
                if (next + sub < n) {
                    c_row = rows[next + sub];
                    c_key = key[next + sub];
                    c_coef = coef[next + sub];
                }
                start = 0;
                next += LPR;
            }
            st4(out + k * D + sub * 4, acc);
        }
    }
}

}  // namespace b2r

using namespace b2r;

extern "C" int b2r_pair_runs_sum(const int64_t* key, const int64_t* rows, const float* coef, const float* T, int64_t n_t,
                                 float* out, int64_t n_out, int64_t n, int d, b2r_stream_t stream) {
    B2R_REQUIRE(key && rows && coef && T && out, B2R_E_BADARG, "b2r_pair_runs_sum: null pointer");
    B2R_REQUIRE(d == 32 || d == 64 || d == 128, B2R_E_UNSUPPORTED, "b2r_pair_runs_sum: d=%d (have 32, 64, 128)", d);
    if (n <= 0) return 0;
    cudaStream_t s = as_stream(stream);
    const int64_t cap = (int64_t)sm_count() * 16;
#define B2R_PR(LPR)                                                                                    \
    do {                                                                                               \
        const int64_t need = (n + 256 / LPR - 1) / (256 / LPR);                                        \
        k_pair_runs_sum<LPR><<<(int)(need < cap ? need : cap), 256, 0, s>>>(key, rows, coef, T, n_t, out, n_out, n); \
    } while (0)
    if (d == 32) B2R_PR(8); else if (d == 64) B2R_PR(16); else B2R_PR(32);
#undef B2R_PR
    B2R_LAUNCH_OK("k_pair_runs_sum");
    return 0;
}

extern "C" int b2r_pairdot_fwd(const float* Q, const int64_t* qidx, int64_t n_q, const float* T, const int64_t* rows,
                               int64_t n_t, float* out, int64_t n, int d, int32_t* err_flag, b2r_stream_t stream) {
    B2R_REQUIRE(Q && qidx && T && rows && out, B2R_E_BADARG, "b2r_pairdot_fwd: null pointer");
    B2R_REQUIRE(n >= 0 && d > 0 && d % 4 == 0, B2R_E_BADARG, "b2r_pairdot_fwd: bad n or d");
    B2R_REQUIRE(aligned16(Q) && aligned16(T), B2R_E_BADARG, "b2r_pairdot_fwd: 16-byte alignment");
    if (n == 0) return 0;
    cudaStream_t s = as_stream(stream);
    constexpr int RCH = 4;
    const int64_t nchunks = (n + RCH - 1) / RCH;
    const int64_t cap = (int64_t)sm_count() * 16;
#define B2R_PD(LPR)                                                                                    \
    do {                                                                                               \
        const int64_t need = (nchunks + 256 / LPR - 1) / (256 / LPR);                                  \
        k_pairdot_fwd<LPR, RCH><<<(int)(need < cap ? need : cap), 256, 0, s>>>(Q, qidx, n_q, T, rows, n_t, out, n, \
                                                                              err_flag);               \
    } while (0)
    if (d == 32) B2R_PD(8);
    else if (d == 64) B2R_PD(16);
    else if (d == 128) B2R_PD(32);
    else {
        const int64_t need = (n + 7) / 8;
        k_pairdot_fwd_generic<<<(int)(need < cap ? need : cap), 256, 0, s>>>(Q, qidx, n_q, T, rows, n_t, out, n, d,
                                                                            err_flag);
    }
#undef B2R_PD
    B2R_LAUNCH_OK("k_pairdot_fwd");
    return 0;
}
