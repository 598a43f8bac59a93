// bprmf_fused_ws.cu -- warp-per-sample variant with the whole sample staged in shared memory (B2R_FUSED=ws; A/B only:
// 39.0 us alone at config 2 vs 43.6 us for the CTA-per-sample kernel, superseded by bprmf_flash.cu).
//
// Same contract as k_bprmf_fused (bprmf_fused.cu): gather the user row and the C candidate rows of a sample, score
// them, evaluate the BPR loss (models/BaseModel.py:182-185) and its closed-form gradient, reduce
// dQ = sum_c g_c * I[id_c] -- every candidate row read from HBM once.  What differs is the decomposition, chosen after
// ncu showed the CTA-per-sample kernel issue-bound (22 M warp instructions for 105 MB, 2 barriers and ~50 shuffles per
// lane per sample) rather than HBM-bound:
//
//   * ONE WARP owns a sample.  No __syncthreads, no cross-group exchange through shared memory for the statistics;
//     warps are fully independent, 8 samples in flight per SM.
//   * A row (d = 64, 256 B) is read by 4 lanes, 64 B each (chunks i*4+sub, i = 0..3, so every 128-bit request of the
//     4 lanes covers 64 contiguous bytes); lane group g of 8 takes candidates c = g, g+8, ...  (up to kK = 13 -> C <= 104).
//   * The whole sample (C x 256 B <= 26 KB) is staged in the warp's private shared-memory region with cp.async
//     (16 B per lane, nothing held in registers while in flight) and read from there twice (score, gradient); each lane
//     reads back only what it copied itself.  Chunk slots are XOR-swizzled with the row parity so the two rows a
//     quarter warp touches fall into different banks.
//   * The 4-lane dot products are reduced with a transposed butterfly (12 shuffles for 13 rows) that leaves lane `sub`
//     with the scores of rows 4*sub .. 4*sub+3 of its group: the exp / sigmoid work is spread over all 32 lanes
//     without redundancy.
//   * dQ partials of the 8 groups meet in the (by then dead) stage region and are added in group order (deterministic).
//
#include "common.cuh"

namespace b2r {

constexpr int kK = 13;                      // rows per lane group
constexpr int kWsD = 64;                    // floats per row
constexpr int kWsRowBytes = kWsD * 4;
constexpr int kWsStageBytes = 8 * kK * kWsRowBytes;   // 26,624 B: one sample
constexpr int kWsWarps = 8;

template <bool PRED, bool QOUT>
__global__ void __launch_bounds__(kWsWarps * 32, 1)    // shared memory allows one CTA per SM: all 255 registers are free
k_bprmf_fused_ws(const float* __restrict__ U, const int64_t* __restrict__ uid, int64_t n_users,
                 const float* __restrict__ T, const int64_t* __restrict__ ids, int64_t n_t,
                 float* __restrict__ pred, float* __restrict__ gout, float* __restrict__ row_loss,
                 float* __restrict__ dQ, float* __restrict__ qout, int B, int C, int32_t* err_flag,
                 float* __restrict__ loss_out, unsigned int* __restrict__ done_counter) {
    extern __shared__ __align__(16) unsigned char ws_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane >> 2, sub = lane & 3;
    unsigned char* stage = ws_smem + warp * kWsStageBytes;
    const float invB = 1.f / (float)B;
    const bool upper = (sub & 2) != 0, odd = (sub & 1) != 0;
    // this lane's four 16-byte slots inside every row of its group: chunk i*4+sub, XOR-swizzled with the row parity
    // (rows of a group are c = grp + 8k, so the parity is the group's); row k of the group sits k * 2048 bytes further
    unsigned char* lane_slot[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        lane_slot[i] = stage + grp * kWsRowBytes + (((i * 4 + sub) ^ ((grp & 1) << 2)) * 16);

#pragma unroll 1
    for (int64_t b = (int64_t)blockIdx.x * kWsWarps + warp; b < B; b += (int64_t)gridDim.x * kWsWarps) {
        // ---- ids: lane l holds candidates l, l+32, l+64, l+96 (0 beyond C) ------------------------------------
        uint32_t idr[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int c = lane + 32 * t;
            idr[t] = (c < C) ? (uint32_t)checked_id(ids[b * C + c], n_t, err_flag) : 0u;
        }
        const int64_t qrow = checked_id(uid[b], n_users, lane == 0 ? err_flag : nullptr);
        // ---- request the candidate rows: candidate c = grp + 8k sits in lane c & 31, register c >> 5 = k >> 2 ---
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            const int c = grp + 8 * k;
            const uint32_t id_c = __shfl_sync(B2R_FULL_MASK, idr[k >> 2], c & 31);
            const float* src = T + (size_t)id_c * kWsD;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int chunk = i * 4 + sub;
                unsigned char* dst = lane_slot[i] + k * (8 * kWsRowBytes);
                if (c < C) {
                    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(dst);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + chunk * 4) : "memory");
                } else {
                    *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);   // padding rows score 0, weigh 0
                }
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        // the user row: chunks i*4+sub of the 256-byte row (every group reads the same row; L1 serves the repeats)
        float4 q[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) q[i] = ld4(U + qrow * kWsD + (i * 4 + sub) * 4);
        if (QOUT && grp == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) st4(qout + b * kWsD + (i * 4 + sub) * 4, q[i]);
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");

        // ---- scores: partial dot over this lane's 64 bytes, then the transposed butterfly over the 4 lanes -------
        float x[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float a = 0.f;
            if (k < kK) {                               // compile-time; no branch on c: padding rows are zero-filled
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 r = *reinterpret_cast<const float4*>(lane_slot[i] + k * (8 * kWsRowBytes));
                    a = fmaf(q[i].x, r.x, fmaf(q[i].y, r.y, fmaf(q[i].z, r.z, fmaf(q[i].w, r.w, a))));
                }
            }
            x[k] = a;
        }
        float y8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float send = upper ? x[j] : x[j + 8];
            const float keep = upper ? x[j + 8] : x[j];
            y8[j] = keep + __shfl_xor_sync(B2R_FULL_MASK, send, 2);
        }
        float xs[4];                                    // scores of rows k = 4*sub + j of this group
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float send = odd ? y8[j] : y8[j + 4];
            const float keep = odd ? y8[j + 4] : y8[j];
            xs[j] = keep + __shfl_xor_sync(B2R_FULL_MASK, send, 1);
        }
        const float p = __shfl_sync(B2R_FULL_MASK, xs[0], 0);      // candidate 0 = group 0, row 0 -> lane 0, xs[0]

        // ---- loss statistics, every lane on its own <= 4 candidates ------------------------------------------
        bool valid[4], neg[4];
        float mloc = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = 4 * sub + j;
            const int c = grp + 8 * k;
            valid[j] = (k < kK) && (c < C);
            neg[j] = valid[j] && (c >= 1);
            if (neg[j]) mloc = fmaxf(mloc, xs[j]);
        }
        const float M = warp_max(mloc);
        float e[4], sg[4], Z = 0.f, A = 0.f, Dd = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            e[j] = 0.f;
            sg[j] = 0.f;
            if (neg[j]) {
                e[j] = __expf(xs[j] - M);
                sg[j] = __fdividef(1.f, 1.f + __expf(xs[j] - p));
            }
            const float es = e[j] * sg[j];
            Z += e[j];
            A += es;
            Dd = fmaf(es, 1.f - sg[j], Dd);
        }
        Z = warp_sum(Z);
        A = warp_sum(A);
        Dd = warp_sum(Dd);
        const float S = (C > 1) ? __fdividef(A, Z) : 0.f;
        const bool inside = (S >= 1e-8f) && (S <= 1.f - 1e-8f);
        const float dS = inside ? -__fdividef(invB, S) : 0.f;
        const float invZ = (C > 1) ? __fdividef(1.f, Z) : 0.f;
        float gq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = grp + 8 * (4 * sub + j);
            float g = 0.f;
            if (valid[j]) {
                g = (c == 0) ? dS * Dd * invZ : dS * (e[j] * invZ) * ((sg[j] - S) - sg[j] * (1.f - sg[j]));
                gout[b * C + c] = g;
                if (PRED) pred[b * C + c] = xs[j];
            }
            gq[j] = g;
        }
        if (lane == 0) row_loss[b] = -logf(fminf(fmaxf(S, 1e-8f), 1.f - 1e-8f));

        // ---- dQ: g of row k lives in lane k >> 2 of the group, register k & 3 -----------------------------------
        float4 acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            const float gk = __shfl_sync(B2R_FULL_MASK, gq[k & 3], (lane & ~3) | (k >> 2));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 r = *reinterpret_cast<const float4*>(lane_slot[i] + k * (8 * kWsRowBytes));
                fma4(acc[i], gk, r);
            }
        }
        __syncwarp();                                   // every lane is done with the rows: the region is reused
        float4* part = reinterpret_cast<float4*>(stage);            // [8 groups][16 chunks]
#pragma unroll
        for (int i = 0; i < 4; ++i) part[grp * 16 + i * 4 + sub] = acc[i];
        __syncwarp();
        const float2* pf = reinterpret_cast<const float2*>(stage);  // 32 float2 per group
        float2 tot = make_float2(0.f, 0.f);
#pragma unroll
        for (int g2 = 0; g2 < 8; ++g2) {
            const float2 t2 = pf[g2 * 32 + lane];
            tot.x += t2.x;
            tot.y += t2.y;
        }
        reinterpret_cast<float2*>(dQ + b * kWsD)[lane] = tot;
        __syncwarp();                                   // before the next sample's copies land in the region
    }

    // mean of the per-sample losses by the last CTA to finish (fixed summation order -> deterministic)
    if (loss_out != nullptr) {
        __shared__ bool last;
        __shared__ float red[kWsWarps * 32];
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
        __syncthreads();
        if (last) {
            __threadfence();
            float a = 0.f;
            for (int i = threadIdx.x; i < B; i += kWsWarps * 32) a += __ldcg(row_loss + i);
            red[threadIdx.x] = a;
            __syncthreads();
            for (int o = kWsWarps * 16; o > 0; o >>= 1) {
                if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                loss_out[0] = red[0] / (float)B;
                *done_counter = 0u;
            }
        }
    }
}

}  // namespace b2r

using namespace b2r;

// returns B2R_E_UNSUPPORTED when the shape is outside the candidate's class (d = 64, C <= 104)
int b2r_bprmf_fused_ws_launch(const float* U, const int64_t* uid, int64_t n_users, const float* I, const int64_t* iid,
                              int64_t n_items, float* pred, float* grad_pred, float* row_loss, float* dQ, float* qout,
                              int B, int C, int d, int32_t* err_flag, float* loss_out, unsigned int* done_counter,
                              b2r_stream_t stream) {
    if (d != kWsD || C > 8 * kK) return set_error(B2R_E_UNSUPPORTED, "fused_ws: d=%d C=%d", d, C);
    const int smem = kWsWarps * kWsStageBytes;
    int64_t need = ((int64_t)B + kWsWarps - 1) / kWsWarps;
    const int64_t cap = sm_count();                              // one CTA per SM (shared memory)
    const int grid = (int)(need < cap ? need : cap);
#define B2R_WS(PRED, QOUT)                                                                                         \
    do {                                                                                                           \
        static bool attr_done = false;                                                                             \
        if (!attr_done) {                                                                                          \
            B2R_CUDA_OK(cudaFuncSetAttribute(k_bprmf_fused_ws<PRED, QOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             smem));                                                               \
            attr_done = true;                                                                                      \
        }                                                                                                          \
        k_bprmf_fused_ws<PRED, QOUT><<<grid, kWsWarps * 32, smem, as_stream(stream)>>>(                            \
            U, uid, n_users, I, iid, n_items, pred, grad_pred, row_loss, dQ, qout, B, C, err_flag, loss_out,       \
            done_counter);                                                                                         \
    } while (0)
    if (pred != nullptr && qout != nullptr) B2R_WS(true, true);
    else if (pred != nullptr) B2R_WS(true, false);
    else if (qout != nullptr) B2R_WS(false, true);
    else B2R_WS(false, false);
#undef B2R_WS
    B2R_LAUNCH_OK("k_bprmf_fused_ws");
    return 0;
}
