// rowdot.cu -- gather + dot scoring and its query-side backward (K1 and half of K2).
//
// Layout: an embedding row of d floats is read by LPR = d/4 lanes, one 128-bit load per lane, so a warp
// covers 32/LPR rows per load instruction and every 128-byte line of the row is fetched whole.  Each
// lane group keeps RCH rows in flight (RCH independent 128-bit loads per lane) before it reduces: with 32
// resident warps per SM that is 128 KB of outstanding row data per SM, past what HBM latency x bandwidth
// needs (~35 KB/SM).  The path is HBM-bound integer-indexed gather: no shared-memory reuse exists except
// the query row, which lives in registers.
#include "common.cuh"

namespace b2r {

constexpr int kThreads = 256;
constexpr int kRowsInFlight = 8;   // RCH

// ---------------------------------------------------------------------------------------------------
// forward: pred[b,c] = <Q[qid[b]], T[ids[b,c]]>
// work item = (sample b, chunk of RCH consecutive candidates); one lane group per item
// ---------------------------------------------------------------------------------------------------
template <int LPR, int RCH>
__global__ void __launch_bounds__(kThreads)
k_rowdot_fwd(const float* __restrict__ Q, const int64_t* __restrict__ qid, int64_t n_q,
             const float* __restrict__ T, const int64_t* __restrict__ ids, int64_t n_t,
             float* __restrict__ pred, int B, int C, int nchunk, int32_t* err_flag) {
    static_assert(RCH <= LPR, "ids of a chunk are loaded one per lane");
    constexpr int D = LPR * 4;
    constexpr int GPC = kThreads / LPR;            // lane groups per CTA
    constexpr int GPW = 32 / LPR;                  // lane groups per warp
    const int sub = threadIdx.x % LPR;
    const int grp = threadIdx.x / LPR;
    const int64_t total = (int64_t)B * nchunk;
    // warp-uniform trip count: the shuffles below need every lane of the warp
    const int64_t warp_first = (int64_t)blockIdx.x * GPC + (grp / GPW) * GPW;
    for (int64_t wbase = warp_first; wbase < total; wbase += (int64_t)gridDim.x * GPC) {
        const int64_t item = wbase + (grp % GPW);
        const bool active = item < total;
        const int b = active ? (int)(item / nchunk) : 0;
        const int c0 = active ? (int)(item % nchunk) * RCH : 0;
        const int nr = active ? min(RCH, C - c0) : 0;

        int64_t qrow = b;
        if (qid != nullptr) qrow = checked_id(qid[b], n_q, sub == 0 && active ? err_flag : nullptr);
        const float4 q = ld4(Q + qrow * D + sub * 4);

        int64_t my_id = 0;
        if (sub < nr) my_id = checked_id(ids[(int64_t)b * C + c0 + sub], n_t, err_flag);

        float4 r[RCH];
#pragma unroll
        for (int k = 0; k < RCH; ++k) {
            const int64_t id_k = __shfl_sync(B2R_FULL_MASK, my_id, k, LPR);
            r[k] = (k < nr) ? ld_row4(T + id_k * D + sub * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float mine = 0.f;
#pragma unroll
        for (int k = 0; k < RCH; ++k) {
            const float s = group_sum<LPR>(dot4(q, r[k]));
            if (sub == k) mine = s;
        }
        if (sub < nr) pred[(int64_t)b * C + c0 + sub] = mine;   // RCH consecutive floats per group
    }
}

// any d % 4 == 0: one warp per (b,c), lanes stride over the 128-bit chunks of the row
__global__ void __launch_bounds__(kThreads)
k_rowdot_fwd_generic(const float* __restrict__ Q, const int64_t* __restrict__ qid, int64_t n_q,
                     const float* __restrict__ T, const int64_t* __restrict__ ids, int64_t n_t,
                     float* __restrict__ pred, int B, int C, int d, int32_t* err_flag) {
    const int lane = threadIdx.x & 31;
    const int64_t total = (int64_t)B * C;
    const int d4 = d >> 2;
    for (int64_t r = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); r < total;
         r += (int64_t)gridDim.x * (kThreads / 32)) {
        const int b = (int)(r / C);
        int64_t qrow = b;
        if (qid != nullptr) qrow = checked_id(qid[b], n_q, lane == 0 ? err_flag : nullptr);
        const int64_t id = checked_id(ids[r], n_t, lane == 0 ? err_flag : nullptr);
        float s = 0.f;
        for (int k = lane; k < d4; k += 32) s += dot4(ld4(Q + qrow * d + k * 4), ld_row4(T + id * d + k * 4));
        s = warp_sum(s);
        if (lane == 0) pred[r] = s;
    }
}

// ---------------------------------------------------------------------------------------------------
// backward w.r.t. the query: dQ[b,:] = sum_c g[b,c] * T[ids[b,c],:]
// GPS lane groups share one sample (chunks round-robin), partial sums meet in shared memory and are
// added in group order -> the result does not depend on scheduling.
// ---------------------------------------------------------------------------------------------------
template <int LPR, int RCH>
__global__ void __launch_bounds__(kThreads)
k_rowdot_bwd_query(const float* __restrict__ g, const float* __restrict__ T,
                   const int64_t* __restrict__ ids, int64_t n_t, float* __restrict__ dQ,
                   int B, int C, int nchunk, int GPS) {
    static_assert(RCH <= LPR, "ids of a chunk are loaded one per lane");
    constexpr int D = LPR * 4;
    constexpr int GPC = kThreads / LPR;
    __shared__ float4 part[GPC][LPR];
    const int sub = threadIdx.x % LPR;
    const int grp = threadIdx.x / LPR;
    const int SPB = GPC / GPS;                      // samples per CTA pass
    const int j = grp % GPS;                        // this group's slot within its sample
    const int trips = (nchunk + GPS - 1) / GPS;     // uniform over the CTA
    for (int64_t sbase = (int64_t)blockIdx.x * SPB; sbase < B; sbase += (int64_t)gridDim.x * SPB) {
        const int64_t b = sbase + grp / GPS;
        const bool have = b < B;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int it = 0; it < trips; ++it) {
            const int ch = it * GPS + j;
            const int c0 = ch * RCH;
            const int nr = (have && ch < nchunk) ? min(RCH, C - c0) : 0;
            int64_t my_id = 0;
            float my_g = 0.f;
            if (sub < nr) {
                my_id = checked_id(ids[b * C + c0 + sub], n_t, nullptr);
                my_g = g[b * C + c0 + sub];
            }
            float4 r[RCH];
#pragma unroll
            for (int k = 0; k < RCH; ++k) {
                const int64_t id_k = __shfl_sync(B2R_FULL_MASK, my_id, k, LPR);
                r[k] = (k < nr) ? ld_row4(T + id_k * D + sub * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int k = 0; k < RCH; ++k) {
                const float gk = __shfl_sync(B2R_FULL_MASK, my_g, k, LPR);
                fma4(acc, gk, r[k]);
            }
        }
        part[grp][sub] = acc;
        __syncthreads();
        if (j == 0 && have) {
            float4 tot = part[grp][sub];
            for (int t = 1; t < GPS; ++t) {
                const float4 x = part[grp + t][sub];
                tot.x += x.x; tot.y += x.y; tot.z += x.z; tot.w += x.w;
            }
            st4(dQ + b * D + sub * 4, tot);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads)
k_rowdot_bwd_query_generic(const float* __restrict__ g, const float* __restrict__ T,
                           const int64_t* __restrict__ ids, int64_t n_t, float* __restrict__ dQ,
                           int B, int C, int d) {
    const int lane = threadIdx.x & 31;
    const int d4 = d >> 2;
    for (int64_t b = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); b < B;
         b += (int64_t)gridDim.x * (kThreads / 32)) {
        for (int k = lane; k < d4; k += 32) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c = 0; c < C; ++c) {
                const int64_t id = checked_id(ids[b * C + c], n_t, nullptr);
                fma4(acc, g[b * C + c], ld_row4(T + id * d + k * 4));
            }
            st4(dQ + b * d + k * 4, acc);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// plain gather out[r,:] = T[ids[r],:]
// ---------------------------------------------------------------------------------------------------
template <int LPR, int RCH>
__global__ void __launch_bounds__(kThreads)
k_gather_rows(const float* __restrict__ T, const int64_t* __restrict__ ids, int64_t n_t,
              float* __restrict__ out, int64_t out_ld, int64_t n, int ids_div, int32_t* err_flag) {
    static_assert(RCH <= LPR, "ids of a chunk are loaded one per lane");
    constexpr int D = LPR * 4;
    constexpr int GPC = kThreads / LPR;
    constexpr int GPW = 32 / LPR;
    const int sub = threadIdx.x % LPR;
    const int grp = threadIdx.x / LPR;
    const int64_t nchunks = (n + RCH - 1) / RCH;
    const int64_t warp_first = (int64_t)blockIdx.x * GPC + (grp / GPW) * GPW;
    for (int64_t wbase = warp_first; wbase < nchunks; wbase += (int64_t)gridDim.x * GPC) {
        const int64_t ch = wbase + (grp % GPW);
        const int64_t r0 = ch * RCH;
        const int nr = (ch < nchunks) ? (int)min((int64_t)RCH, n - r0) : 0;
        int64_t my_id = 0;
        if (sub < nr) my_id = checked_id(ids[(r0 + sub) / ids_div], n_t, err_flag);
        float4 r[RCH];
#pragma unroll
        for (int k = 0; k < RCH; ++k) {
            const int64_t id_k = __shfl_sync(B2R_FULL_MASK, my_id, k, LPR);
            if (k < nr) r[k] = ld_row4(T + id_k * D + sub * 4);
        }
#pragma unroll
        for (int k = 0; k < RCH; ++k)
            if (k < nr) st4(out + (r0 + k) * out_ld + sub * 4, r[k]);
    }
}

__global__ void __launch_bounds__(kThreads)
k_gather_rows_generic(const float* __restrict__ T, const int64_t* __restrict__ ids, int64_t n_t,
                      float* __restrict__ out, int64_t out_ld, int64_t n, int d, int ids_div, int32_t* err_flag) {
    const int lane = threadIdx.x & 31;
    const int d4 = d >> 2;
    for (int64_t r = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); r < n;
         r += (int64_t)gridDim.x * (kThreads / 32)) {
        const int64_t id = checked_id(ids[r / ids_div], n_t, lane == 0 ? err_flag : nullptr);
        for (int k = lane; k < d4; k += 32) st4(out + r * out_ld + k * 4, ld_row4(T + id * d + k * 4));
    }
}

static int grid_for(int64_t ctas_needed) {
    const int64_t cap = (int64_t)sm_count() * 16;   // several resident CTAs per SM, grid-stride beyond that
    int64_t g = ctas_needed < cap ? ctas_needed : cap;
    return (int)(g < 1 ? 1 : g);
}

static int pow2_at_least(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

}  // namespace b2r

using namespace b2r;

extern "C" int b2r_rowdot_fwd(const float* Q, const int64_t* qid, int64_t n_q, const float* T,
                              const int64_t* ids, int64_t n_t, float* pred, int B, int C, int d,
                              int32_t* err_flag, b2r_stream_t stream) {
    B2R_REQUIRE(Q && T && ids && pred, B2R_E_BADARG, "b2r_rowdot_fwd: null pointer");
    B2R_REQUIRE(B >= 0 && C >= 0 && d > 0 && d % 4 == 0, B2R_E_BADARG,
                "b2r_rowdot_fwd: need B,C >= 0 and d %% 4 == 0 (B=%d C=%d d=%d)", B, C, d);
    B2R_REQUIRE(aligned16(Q) && aligned16(T), B2R_E_BADARG, "b2r_rowdot_fwd: tables must be 16-byte aligned");
    if (B == 0 || C == 0) return 0;
    cudaStream_t s = as_stream(stream);
    constexpr int RCH = kRowsInFlight;
    const int nchunk = (C + RCH - 1) / RCH;
    const int64_t items = (int64_t)B * nchunk;
#define B2R_FWD(LPR)                                                                                   \
    k_rowdot_fwd<LPR, RCH><<<grid_for((items + kThreads / LPR - 1) / (kThreads / LPR)), kThreads, 0, s>>>( \
        Q, qid, n_q, T, ids, n_t, pred, B, C, nchunk, err_flag)
    if (d == 32) B2R_FWD(8);
    else if (d == 64) B2R_FWD(16);
    else if (d == 128) B2R_FWD(32);
    else
        k_rowdot_fwd_generic<<<grid_for(((int64_t)B * C + 7) / 8), kThreads, 0, s>>>(Q, qid, n_q, T, ids, n_t,
                                                                                      pred, B, C, d, err_flag);
#undef B2R_FWD
    B2R_LAUNCH_OK("k_rowdot_fwd");
    return 0;
}

extern "C" int b2r_rowdot_bwd_query(const float* g, const float* T, const int64_t* ids, int64_t n_t,
                                    float* dQ, int B, int C, int d, b2r_stream_t stream) {
    B2R_REQUIRE(g && T && ids && dQ, B2R_E_BADARG, "b2r_rowdot_bwd_query: null pointer");
    B2R_REQUIRE(B >= 0 && C >= 0 && d > 0 && d % 4 == 0, B2R_E_BADARG,
                "b2r_rowdot_bwd_query: need B,C >= 0 and d %% 4 == 0 (B=%d C=%d d=%d)", B, C, d);
    B2R_REQUIRE(aligned16(T) && aligned16(dQ), B2R_E_BADARG, "b2r_rowdot_bwd_query: 16-byte alignment");
    if (B == 0) return 0;
    cudaStream_t s = as_stream(stream);
    if (C == 0) {
        B2R_CUDA_OK(cudaMemsetAsync(dQ, 0, (size_t)B * d * sizeof(float), s));
        return 0;
    }
    constexpr int RCH = kRowsInFlight;
    const int nchunk = (C + RCH - 1) / RCH;
#define B2R_BWDQ(LPR)                                                                                  \
    do {                                                                                               \
        constexpr int GPC = kThreads / LPR;                                                            \
        const int GPS = min(GPC, pow2_at_least(nchunk));                                               \
        const int SPB = GPC / GPS;                                                                     \
        k_rowdot_bwd_query<LPR, RCH><<<grid_for(((int64_t)B + SPB - 1) / SPB), kThreads, 0, s>>>(      \
            g, T, ids, n_t, dQ, B, C, nchunk, GPS);                                                    \
    } while (0)
    if (d == 32) B2R_BWDQ(8);
    else if (d == 64) B2R_BWDQ(16);
    else if (d == 128) B2R_BWDQ(32);
    else
        k_rowdot_bwd_query_generic<<<grid_for(((int64_t)B + 7) / 8), kThreads, 0, s>>>(g, T, ids, n_t, dQ, B, C, d);
#undef B2R_BWDQ
    B2R_LAUNCH_OK("k_rowdot_bwd_query");
    return 0;
}

extern "C" int b2r_gather_rows(const float* T, const int64_t* ids, int64_t n_t, float* out, int64_t n,
                               int d, int32_t* err_flag, b2r_stream_t stream) {
    return b2r_gather_rows_strided(T, ids, n_t, out, d, n, d, 1, err_flag, stream);
}

extern "C" int b2r_gather_rows_strided(const float* T, const int64_t* ids, int64_t n_t, float* out, int out_ld,
                                       int64_t n, int d, int ids_div, int32_t* err_flag, b2r_stream_t stream) {
    B2R_REQUIRE(T && ids && out, B2R_E_BADARG, "b2r_gather_rows: null pointer");
    B2R_REQUIRE(n >= 0 && d > 0 && d % 4 == 0, B2R_E_BADARG, "b2r_gather_rows: need n >= 0, d %% 4 == 0");
    B2R_REQUIRE(out_ld >= d && out_ld % 4 == 0 && ids_div >= 1, B2R_E_BADARG,
                "b2r_gather_rows: out_ld=%d must be >= d and a multiple of 4, ids_div=%d >= 1", out_ld, ids_div);
    B2R_REQUIRE(aligned16(T) && aligned16(out), B2R_E_BADARG, "b2r_gather_rows: 16-byte alignment");
    if (n == 0) return 0;
    cudaStream_t s = as_stream(stream);
    constexpr int RCH = kRowsInFlight;
    const int64_t nchunks = (n + RCH - 1) / RCH;
#define B2R_GATHER(LPR)                                                                                \
    k_gather_rows<LPR, RCH><<<grid_for((nchunks + kThreads / LPR - 1) / (kThreads / LPR)), kThreads, 0, s>>>( \
        T, ids, n_t, out, out_ld, n, ids_div, err_flag)
    if (d == 32) B2R_GATHER(8);
    else if (d == 64) B2R_GATHER(16);
    else if (d == 128) B2R_GATHER(32);
    else
        k_gather_rows_generic<<<grid_for((n + 7) / 8), kThreads, 0, s>>>(T, ids, n_t, out, out_ld, n, d, ids_div,
                                                                         err_flag);
#undef B2R_GATHER
    B2R_LAUNCH_OK("k_gather_rows");
    return 0;
}
