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

// A CTA walks passes p = blockIdx.x, += gridDim.x; a pass covers SPB = GPC/GPS samples.  The candidate rows of
// pass p+1 are requested (second register set) before pass p enters its barrier phases, so HBM latency overlaps
// the loss/gradient arithmetic and the three barriers of the current pass.  Phases of a pass:
//   rows -> scores (shared memory)  |barrier|  one WARP per sample: softmax/sigmoid statistics and the gradient g of
//   every candidate, each transcendental evaluated once  |barrier|  lane groups: acc = sum g*row  |barrier|
//   ordered combine of the GPS partials -> dQ.   Shared buffers alternate by pass parity (no trailing barrier).
template <int LPR, int RPG>
struct FusedPass {
    static constexpr int D = LPR * 4;
    static constexpr int GPC = 256 / LPR;
    static constexpr int CPL = (GPC * RPG + 31) / 32;      // candidates per lane in the statistics warp

    const float* U; const int64_t* uid; int64_t n_users;
    const float* T; const int64_t* ids; int64_t n_t;
    float* pred; float* gout; float* row_loss; float* dQ;
    int B, C, GPS; int32_t* err_flag;
    int sub, grp, lane, warp, SPB, j, slot, c_mine;
    float invB;

    __device__ __forceinline__ void load(int64_t pass, float4 (&r)[RPG], float4& q) const {
        const int64_t b = pass * SPB + slot;
        const bool have = b < B;
        const bool mine_ok = have && sub < RPG && c_mine < C;
        int64_t qrow = 0;
        if (have) qrow = checked_id(uid[b], n_users, sub == 0 && j == 0 ? err_flag : nullptr);
        q = ld4(U + qrow * D + sub * 4);
        int64_t my_id = 0;
        if (mine_ok) my_id = checked_id(ids[b * C + c_mine], n_t, err_flag);
#pragma unroll
        for (int k = 0; k < RPG; ++k) {
            const int64_t id_k = __shfl_sync(B2R_FULL_MASK, my_id, k, LPR);
            const bool ok = have && (j + GPS * k) < C;
            r[k] = ok ? ld_row4(T + id_k * D + sub * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }

    __device__ __forceinline__ void compute(int64_t pass, const float4 (&r)[RPG], const float4& q, float* sp,
                                            float4 (*part)[LPR]) const {
        const int64_t b0 = pass * SPB;
        const int64_t b = b0 + slot;
        const bool have = b < B;
        float mine = 0.f;
#pragma unroll
        for (int k = 0; k < RPG; ++k) {
            const float v = group_sum<LPR>(dot4(q, r[k]));
            if (sub == k) mine = v;
        }
        if (have && sub < RPG && c_mine < C) {
            sp[slot * (GPS * RPG) + c_mine] = mine;
            if (pred != nullptr) pred[b * C + c_mine] = mine;
        }
        __syncthreads();
        for (int sl = warp; sl < SPB; sl += 8) {                 // warp-uniform
            const int64_t bb = b0 + sl;
            if (bb >= B) continue;
            float* myp = sp + sl * (GPS * RPG);
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
            __syncwarp();                                        // every lane has read its scores: overwrite with g
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int c = 1 + lane + 32 * i;
                if (c < C) {
                    const float g = dS * (e[i] * invZ) * ((sg[i] - S) - sg[i] * (1.f - sg[i]));
                    myp[c] = g;
                    gout[bb * C + c] = g;
                }
            }
            if (lane == 0) {
                const float g0 = dS * Dp * invZ;
                myp[0] = g0;
                gout[bb * C] = g0;
                row_loss[bb] = -logf(Sc);
            }
        }
        __syncthreads();
        const float* myg = sp + slot * (GPS * RPG);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < RPG; ++k) {
            const int c = j + GPS * k;
            const float gk = (have && c < C) ? myg[c] : 0.f;
            fma4(acc, gk, r[k]);
        }
        part[grp][sub] = acc;
        __syncthreads();
        if (j == 0 && have) {
            float4 tot = part[grp][sub];
            for (int t = 1; t < GPS; ++t) {
                const float4 y = part[grp + t][sub];
                tot.x += y.x; tot.y += y.y; tot.z += y.z; tot.w += y.w;
            }
            st4(dQ + b * D + sub * 4, tot);
        }
    }
};

template <int LPR, int RPG>
__global__ void __launch_bounds__(256, 2)
k_bprmf_fused(const float* __restrict__ U, const int64_t* __restrict__ uid, int64_t n_users,
              const float* __restrict__ T, const int64_t* __restrict__ ids, int64_t n_t,
              float* __restrict__ pred, float* __restrict__ gout, float* __restrict__ row_loss,
              float* __restrict__ dQ, int B, int C, int GPS, int32_t* err_flag) {
    static_assert(RPG <= LPR, "ids of a group's rows are loaded one per lane");
    using P = FusedPass<LPR, RPG>;
    __shared__ float sp[2][P::GPC * RPG];                 // scores, then gradients (alternating by pass parity)
    __shared__ float4 part[2][P::GPC][LPR];               // partial dQ per group
    P f;
    f.U = U; f.uid = uid; f.n_users = n_users; f.T = T; f.ids = ids; f.n_t = n_t;
    f.pred = pred; f.gout = gout; f.row_loss = row_loss; f.dQ = dQ; f.B = B; f.C = C; f.GPS = GPS; f.err_flag = err_flag;
    f.lane = threadIdx.x & 31; f.warp = threadIdx.x >> 5;
    f.sub = threadIdx.x % LPR; f.grp = threadIdx.x / LPR;
    f.SPB = P::GPC / GPS; f.j = f.grp % GPS; f.slot = f.grp / GPS; f.c_mine = f.j + GPS * f.sub;
    f.invB = 1.f / (float)B;
    const int64_t npass = ((int64_t)B + f.SPB - 1) / f.SPB;
    float4 ra[RPG], rb[RPG], qa, qb;
    int64_t p = blockIdx.x;
    if (p < npass) f.load(p, ra, qa);
    for (; p < npass; p += 2 * (int64_t)gridDim.x) {
        const int64_t p1 = p + gridDim.x;
        if (p1 < npass) f.load(p1, rb, qb);               // next pass's rows are in flight during this pass
        f.compute(p, ra, qa, sp[0], part[0]);
        if (p1 < npass) {
            const int64_t p2 = p1 + gridDim.x;
            if (p2 < npass) f.load(p2, ra, qa);
            f.compute(p1, rb, qb, sp[1], part[1]);
        }
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
    if (GPC == 0 || !pick_shape(C, GPC, 8, &RPG, &GPS))
        return set_error(B2R_E_UNSUPPORTED, "b2r_bprmf_fused_fwd_bwd: no fused variant for d=%d C=%d", d, C);
    const int SPB = GPC / GPS;
    const int64_t need = ((int64_t)B + SPB - 1) / SPB;          // passes
    const int64_t cap = (int64_t)sm_count() * 2;                // persistent: 2 resident CTAs per SM, each pipelined
    const int grid = (int)(need < cap ? need : cap);
#define B2R_FUSED(LPR, R)                                                                              \
    k_bprmf_fused<LPR, R><<<grid, 256, 0, s>>>(U, uid, n_users, I, iid, n_items, pred, grad_pred, row_loss, dQ, B, C, \
                                               GPS, err_flag)
    if (d == 32) {
        if (RPG == 2) B2R_FUSED(8, 2); else if (RPG == 4) B2R_FUSED(8, 4); else B2R_FUSED(8, 8);
    } else if (d == 64) {
        if (RPG == 2) B2R_FUSED(16, 2); else if (RPG == 4) B2R_FUSED(16, 4); else B2R_FUSED(16, 8);
    } else {
        if (RPG == 2) B2R_FUSED(32, 2); else if (RPG == 4) B2R_FUSED(32, 4); else B2R_FUSED(32, 8);
    }
#undef B2R_FUSED
    B2R_LAUNCH_OK("k_bprmf_fused");
    return 0;
}
