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
// contiguous run and no sort is needed; a lane group owns the run whose first element it lands on and walks it in
// order (4 independent partial sums, combined in a fixed order).  Every output row is written at most once; rows
// without elements keep what the caller put there (zeros).
template <int LPR>
__global__ void __launch_bounds__(256)
k_pair_runs_sum(const int64_t* __restrict__ key, const int64_t* __restrict__ rows, const float* __restrict__ coef,
                const float* __restrict__ T, int64_t n_t, float* __restrict__ out, int64_t n_out, int64_t n) {
    constexpr int D = LPR * 4;
    constexpr int GPC = 256 / LPR;
    const int sub = threadIdx.x % LPR, grp = threadIdx.x / LPR;
    for (int64_t e = (int64_t)blockIdx.x * GPC + grp; e < n; e += (int64_t)gridDim.x * GPC) {
        if (rows[e] < 0) continue;
        const int64_t k = key[e];
        if (e > 0 && rows[e - 1] >= 0 && key[e - 1] == k) continue;          // not the first element of its run
        if (k < 0 || k >= n_out) continue;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
        int64_t j = e;
        for (;;) {
            int64_t r[4];
            float c[4];
            int cnt = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                r[u] = -1;
                c[u] = 0.f;
                if (cnt == u && j < n && rows[j] >= 0 && key[j] == k) {
                    r[u] = rows[j] < n_t ? rows[j] : 0;
                    c[u] = coef[j];
                    ++j;
                    ++cnt;
                }
            }
            if (cnt == 0) break;
            if (r[0] >= 0) fma4(a0, c[0], ld_row4(T + r[0] * D + sub * 4));
            if (r[1] >= 0) fma4(a1, c[1], ld_row4(T + r[1] * D + sub * 4));
            if (r[2] >= 0) fma4(a2, c[2], ld_row4(T + r[2] * D + sub * 4));
            if (r[3] >= 0) fma4(a3, c[3], ld_row4(T + r[3] * D + sub * 4));
            if (cnt < 4) break;
        }
        a0.x = (a0.x + a1.x) + (a2.x + a3.x);
        a0.y = (a0.y + a1.y) + (a2.y + a3.y);
        a0.z = (a0.z + a1.z) + (a2.z + a3.z);
        a0.w = (a0.w + a1.w) + (a2.w + a3.w);
        st4(out + k * D + sub * 4, a0);
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
