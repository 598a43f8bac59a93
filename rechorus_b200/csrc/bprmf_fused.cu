// bprmf_fused.cu -- K1+K2a in one pass: gather the user row and the (1+K) candidate rows, score them, evaluate
// the BPR loss and its closed-form gradient, and reduce dQ = sum_c g_c * I[id_c] -- with every candidate row
// read from HBM exactly once and kept in registers between the scoring and the gradient phase.
//
// Work decomposition: a sample is owned by GPS lane groups (GPS a power of two <= groups per CTA); group j takes
// candidates c = j, j+GPS, j+2*GPS ... (at most RPG of them, all loaded before the first reduction -> RPG
// independent 128-bit loads in flight per lane).  Scores meet in shared memory, every group then derives the
// softmax/sigmoid statistics of its sample redundantly (C <= 256 values, cheaper than another barrier),
// computes g for its own rows, accumulates g*row in registers, and the GPS partial dQ vectors are summed in
// group order (deterministic).  replaces: BPRMF.py:39-42 forward, BaseModel.py:182-185 loss, and the
// mul/sum + loss half of loss.backward() (BaseRunner.py:205).
#include <stdlib.h>

#include "common.cuh"

namespace b2r {

__device__ __forceinline__ float sigmoidf_f(float x) { return 1.f / (1.f + expf(-x)); }

template <int LPR>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(B2R_FULL_MASK, v, o));
    return v;
}

// SPI samples per lane-group slot are processed together per CTA pass: all their candidate rows are requested
// before the first reduction (SPI*RPG independent 128-bit loads per lane) and every barrier is shared.
// Phases per pass:  rows -> scores (shared memory)  |barrier|  one WARP per sample: softmax/sigmoid statistics and
// the gradient g of every candidate, each transcendental evaluated once  |barrier|  lane groups: acc = sum g*row
// |barrier|  ordered combine of the GPS partials -> dQ.
template <int LPR, int RPG, int SPI>
__global__ void __launch_bounds__(256, (RPG * SPI <= 8) ? 3 : 2)
k_bprmf_fused(const float* __restrict__ U, const int64_t* __restrict__ uid, int64_t n_users,
              const float* __restrict__ T, const int64_t* __restrict__ ids, int64_t n_t,
              float* __restrict__ pred, float* __restrict__ gout, float* __restrict__ row_loss,
              float* __restrict__ dQ, int B, int C, int GPS, int32_t* err_flag) {
    static_assert(RPG <= LPR, "ids of a group's rows are loaded one per lane");
    constexpr int D = LPR * 4;
    constexpr int GPC = 256 / LPR;
    constexpr int CPL = (GPC * RPG + 31) / 32;          // candidates per lane in the statistics warp (C <= GPC*RPG)
    __shared__ float sp[SPI][GPC * RPG];                // scores, then gradients, of the samples this CTA holds
    __shared__ float4 part[SPI][GPC][LPR];              // partial dQ per group
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = threadIdx.x % LPR;
    const int grp = threadIdx.x / LPR;
    const int SPB = GPC / GPS;                          // sample slots per pass (x SPI samples each)
    const int j = grp % GPS;
    const int slot = grp / GPS;
    const float invB = 1.f / (float)B;
    const int c_mine = j + GPS * sub;                   // the candidate lane `sub` speaks for
    for (int64_t sbase = (int64_t)blockIdx.x * SPB * SPI; sbase < B; sbase += (int64_t)gridDim.x * SPB * SPI) {
        int64_t bs[SPI];
        bool have[SPI];
        float4 q[SPI];
        float4 r[SPI][RPG];
#pragma unroll
        for (int s = 0; s < SPI; ++s) {
            bs[s] = sbase + (int64_t)s * SPB + slot;
            have[s] = bs[s] < B;
            const bool mine_ok = have[s] && sub < RPG && c_mine < C;
            int64_t qrow = 0;
            if (have[s]) qrow = checked_id(uid[bs[s]], n_users, sub == 0 && j == 0 ? err_flag : nullptr);
            q[s] = ld4(U + qrow * D + sub * 4);
            int64_t my_id = 0;
            if (mine_ok) my_id = checked_id(ids[bs[s] * C + c_mine], n_t, err_flag);
#pragma unroll
            for (int k = 0; k < RPG; ++k) {
                const int64_t id_k = __shfl_sync(B2R_FULL_MASK, my_id, k, LPR);
                const bool ok = have[s] && (j + GPS * k) < C;
                r[s][k] = ok ? ld_row4(T + id_k * D + sub * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int s = 0; s < SPI; ++s) {
            float mine = 0.f;
#pragma unroll
            for (int k = 0; k < RPG; ++k) {
                const float v = group_sum<LPR>(dot4(q[s], r[s][k]));
                if (sub == k) mine = v;
            }
            if (have[s] && sub < RPG && c_mine < C) {
                sp[s][slot * (GPS * RPG) + c_mine] = mine;
                if (pred != nullptr) pred[bs[s] * C + c_mine] = mine;
            }
        }
        __syncthreads();
        // ---- statistics + gradient: warp w owns sample slots w, w+8, ... of this pass ----------------------
        for (int ss = warp; ss < SPB * SPI; ss += 8) {
            const int s = ss / SPB, sl = ss % SPB;
            const int64_t b = sbase + (int64_t)s * SPB + sl;
            if (b >= B) continue;                           // warp-uniform
            float* myp = sp[s] + sl * (GPS * RPG);
            const float p = myp[0];
            float x[CPL], e[CPL], sg[CPL];
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int c = 1 + lane + 32 * i;
                x[i] = (c < C) ? myp[c] : -INFINITY;
                mx = fmaxf(mx, x[i]);
            }
            mx = warp_max(mx);
            float Z = 0.f, A = 0.f, Dp = 0.f;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int c = 1 + lane + 32 * i;
                e[i] = 0.f;
                sg[i] = 0.f;
                if (c < C) {
                    e[i] = expf(x[i] - mx);
                    sg[i] = sigmoidf_f(p - x[i]);
                    Z += e[i];
                    A = fmaf(e[i], sg[i], A);
                    Dp = fmaf(e[i] * sg[i], 1.f - sg[i], Dp);
                }
            }
            Z = warp_sum(Z);
            A = warp_sum(A);
            Dp = warp_sum(Dp);
            const float S = (C > 1) ? A / Z : 0.f;
            const bool inside = (S >= 1e-8f) && (S <= 1.f - 1e-8f);
            const float Sc = fminf(fmaxf(S, 1e-8f), 1.f - 1e-8f);
            const float dS = inside ? -invB / S : 0.f;
            const float invZ = (C > 1) ? 1.f / Z : 0.f;
            __syncwarp();                                   // all lanes have read their scores: overwrite with g
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int c = 1 + lane + 32 * i;
                if (c < C) {
                    const float g = dS * (e[i] * invZ) * ((sg[i] - S) - sg[i] * (1.f - sg[i]));
                    myp[c] = g;
                    gout[b * C + c] = g;
                }
            }
            if (lane == 0) {
                const float g0 = dS * Dp * invZ;
                myp[0] = g0;
                gout[b * C] = g0;
                row_loss[b] = -logf(Sc);
            }
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < SPI; ++s) {
            const float* myg = sp[s] + slot * (GPS * RPG);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < RPG; ++k) {
                const int c = j + GPS * k;
                const float gk = (have[s] && c < C) ? myg[c] : 0.f;
                fma4(acc, gk, r[s][k]);
            }
            part[s][grp][sub] = acc;
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < SPI; ++s) {
            if (j == 0 && have[s]) {
                float4 tot = part[s][grp][sub];
                for (int t = 1; t < GPS; ++t) {
                    const float4 y = part[s][grp + t][sub];
                    tot.x += y.x; tot.y += y.y; tot.z += y.z; tot.w += y.w;
                }
                st4(dQ + bs[s] * D + sub * 4, tot);
            }
        }
        __syncthreads();
    }
}

static int pow2ceil(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

// pick rows-per-group and groups-per-sample: smallest padding GPS*RPG - C, ties -> more rows in flight
static bool pick_shape(int C, int GPC, int max_rpg, int* RPG, int* GPS) {
    int best_waste = 1 << 30;
    bool found = false;
    for (int rpg = max_rpg; rpg >= 2; rpg >>= 1) {
        const int gps = pow2ceil((C + rpg - 1) / rpg);
        if (gps > GPC) continue;
        const int waste = gps * rpg - C;
        if (waste < best_waste) {
            best_waste = waste;
            *RPG = rpg;
            *GPS = gps;
            found = true;
        }
    }
    return found;
}

}  // namespace b2r

using namespace b2r;

// returns B2R_E_UNSUPPORTED (without touching the error string semantics) when the shape has no fused variant
extern "C" int b2r_bprmf_fused_fwd_bwd(const float* U, const int64_t* uid, int64_t n_users, const float* I,
                                       const int64_t* iid, int64_t n_items, float* pred, float* grad_pred,
                                       float* row_loss, float* dQ, int B, int C, int d, int32_t* err_flag,
                                       b2r_stream_t stream) {
    B2R_REQUIRE(U && uid && I && iid && grad_pred && row_loss && dQ, B2R_E_BADARG, "b2r_bprmf_fused_fwd_bwd: null pointer");
    B2R_REQUIRE(B > 0 && C > 0, B2R_E_BADARG, "b2r_bprmf_fused_fwd_bwd: B=%d C=%d", B, C);
    B2R_REQUIRE(aligned16(U) && aligned16(I) && aligned16(dQ), B2R_E_BADARG, "b2r_bprmf_fused_fwd_bwd: alignment");
    cudaStream_t s = as_stream(stream);
    int RPG = 0, GPS = 0;
    const int GPC = d == 32 ? 32 : (d == 64 ? 16 : (d == 128 ? 8 : 0));
    if (GPC == 0 || !pick_shape(C, GPC, d == 128 ? 16 : 8, &RPG, &GPS))
        return set_error(B2R_E_UNSUPPORTED, "b2r_bprmf_fused_fwd_bwd: no fused variant for d=%d C=%d", d, C);
    static int spi_env = -1;                // tuning knob B2R_FUSED_SPI = 1 | 2 (read once)
    if (spi_env < 0) {
        const char* e = getenv("B2R_FUSED_SPI");
        spi_env = e ? atoi(e) : 2;
        if (spi_env != 1 && spi_env != 2) spi_env = 2;
    }
    const int SPI = (RPG <= 8) ? spi_env : 1;
    const int SPB = GPC / GPS;
    const int64_t need = ((int64_t)B + (int64_t)SPB * SPI - 1) / ((int64_t)SPB * SPI);
    const int64_t cap = (int64_t)sm_count() * 8;
    const int grid = (int)(need < cap ? need : cap);
#define B2R_FUSED(LPR, R)                                                                              \
    do {                                                                                               \
        if (SPI == 2)                                                                                  \
            k_bprmf_fused<LPR, R, 2><<<grid, 256, 0, s>>>(U, uid, n_users, I, iid, n_items, pred, grad_pred, row_loss, \
                                                          dQ, B, C, GPS, err_flag);                    \
        else                                                                                           \
            k_bprmf_fused<LPR, R, 1><<<grid, 256, 0, s>>>(U, uid, n_users, I, iid, n_items, pred, grad_pred, row_loss, \
                                                          dQ, B, C, GPS, err_flag);                    \
    } while (0)
#define B2R_FUSED1(LPR, R)                                                                             \
    k_bprmf_fused<LPR, R, 1><<<grid, 256, 0, s>>>(U, uid, n_users, I, iid, n_items, pred, grad_pred, row_loss, dQ, B, C, \
                                                  GPS, err_flag)
    if (d == 32) {
        if (RPG == 2) B2R_FUSED(8, 2); else if (RPG == 4) B2R_FUSED(8, 4); else B2R_FUSED(8, 8);
    } else if (d == 64) {
        if (RPG == 2) B2R_FUSED(16, 2); else if (RPG == 4) B2R_FUSED(16, 4); else B2R_FUSED(16, 8);
    } else {
        if (RPG == 2) B2R_FUSED(32, 2); else if (RPG == 4) B2R_FUSED(32, 4);
        else if (RPG == 8) B2R_FUSED(32, 8); else B2R_FUSED1(32, 16);
    }
#undef B2R_FUSED
#undef B2R_FUSED1
    B2R_LAUNCH_OK("k_bprmf_fused");
    return 0;
}
