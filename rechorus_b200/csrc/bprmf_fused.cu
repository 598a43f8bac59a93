// bprmf_fused.cu -- K1+K2a in one pass: gather the user row and the (1+K) candidate rows, score them, evaluate
// the BPR loss and its closed-form gradient, and reduce dQ = sum_c g_c * I[id_c] -- with every candidate row
// read from HBM exactly once: it is staged in shared memory by cp.async one pass ahead of its use and read from
// there for the score and again for the gradient.
//
// Work decomposition: a sample is owned by GPS lane groups (GPS a power of two <= groups per CTA); group j takes
// candidates c = j, j+GPS, j+2*GPS ... (at most RPG of them).  Every group derives the positive's score itself, the
// groups' softmax/sigmoid statistics meet in shared memory (online-softmax combine), each lane then has g for its own
// candidate, the group accumulates g*row, and the GPS partial dQ vectors are summed in group order (deterministic).
// replaces: BPRMF.py:39-42 forward, BaseModel.py:182-185 loss, and the mul/sum + loss half of loss.backward()
// (BaseRunner.py:205).
#include <stdlib.h>

#include "common.cuh"
#include "plan_direct.cuh"

namespace b2r {

__device__ __forceinline__ float sigmoidf_f(float x) { return 1.f / (1.f + expf(-x)); }

template <int LPR>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(B2R_FULL_MASK, v, o));
    return v;
}

// sum each of the RPG per-lane values over the LPR lanes of a group with RPG + log2(LPR/RPG) shuffles instead of
// RPG * log2(LPR): at every step the lanes split the live values in two halves and exchange the half they give up.
// Afterwards lane `sub` holds the total of value sub / (LPR/RPG)  (lanes of one block of LPR/RPG hold copies).
template <int LPR, int RPG>
__device__ __forceinline__ float group_sum_multi(float (&v)[RPG], int sub) {
    int n = RPG;
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
        if (n > 1) {
            const bool upper = (sub & o) != 0;
            n >>= 1;
#pragma unroll
            for (int i = 0; i < RPG / 2; ++i) {
                if (i < n) {
                    const float send = upper ? v[i] : v[i + n];
                    const float keep = upper ? v[i + n] : v[i];
                    v[i] = keep + __shfl_xor_sync(B2R_FULL_MASK, send, o);
                }
            }
        } else {
            v[0] += __shfl_xor_sync(B2R_FULL_MASK, v[0], o);
        }
    }
    return v[0];
}

// One pass = SPB = GPC/GPS samples.  Per pass, per lane group (LPR lanes, RPG candidate rows + the positive's row):
//   dots        the group's RPG scores and, redundantly in every group, the positive's score p (its row is an L1/L2
//               hit after the first group) -- so no cross-group exchange is needed before the sigmoid;
//   local stats m_g = max of the group's negatives, e = exp(x - m_g), s = sigmoid(p - x), group sums Z_g, A_g, D_g
//               -> 4 floats per group to shared memory                                           |barrier 1|
//   combine     lane t reads group t's 4 floats; M = max m_t, rescale by exp(m_t - M) (online softmax), group sums
//               -> S, dS, 1/Z; gradient g of the own candidate; acc = sum_k g_k * row_k -> shared |barrier 2|
//   reduce      4*LPR threads per sample add the GPS partial dQ vectors in group order (deterministic).
// Shared buffers alternate by pass parity, so no third barrier is needed.  Transcendentals use the fast
// hardware forms (ex2/rcp based, <= 2 ulp): deviations are ~1e-7 relative on g, far inside the 1e-5 bar.
template <int LPR, int RPG>
struct FusedPass {
    static constexpr int D = LPR * 4;
    static constexpr int GPC = 256 / LPR;
    static constexpr int RS = LPR / RPG;                  // lanes per "speaker" block after group_sum_multi

    const float* U; const int64_t* uid; int64_t n_users;
    const float* T; const int64_t* ids; int64_t n_t;
    float* pred; float* gout; float* row_loss; float* dQ; float* qout;
    int B, C, GPS; int32_t* err_flag;
    int sub, grp, SPB, j, slot;
    int c_load;          // candidate whose id this lane loads (row `sub` of the group), valid if sub < RPG
    int c_mine;          // candidate whose score/gradient this lane speaks for (row sub / RS), if sub % RS == 0
    float invB;

    struct Ids {
        int64_t qrow;
        uint32_t my_id, pos_id;
    };

    __device__ __forceinline__ Ids load_ids(int64_t pass) const {
        const int64_t b = pass * SPB + slot;
        Ids r{0, 0u, 0u};
        if (b < B) {
            r.qrow = checked_id(uid[b], n_users, sub == 0 && j == 0 ? err_flag : nullptr);
            if (sub < RPG && c_load < C) r.my_id = (uint32_t)checked_id(ids[b * C + c_load], n_t, err_flag);
            r.pos_id = (uint32_t)checked_id(ids[b * C], n_t, nullptr);
        }
        return r;
    }

    // request a pass's rows: the RPG candidate rows go to this thread's slots of a shared-memory stage with cp.async
    // (16 B per lane, nothing held in registers while in flight), the user row and the positive's row to registers
    __device__ __forceinline__ void issue_rows(int64_t pass, const Ids& id, float4* stage, float4& rp, float4& q) const {
        const bool have = pass * SPB + slot < B;
        q = ld4(U + id.qrow * D + sub * 4);
        rp = have ? ld4(T + (size_t)id.pos_id * D + sub * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < RPG; ++k) {
            const uint32_t id_k = __shfl_sync(B2R_FULL_MASK, id.my_id, k, LPR);
            const bool ok = have && (j + GPS * k) < C;
            float4* dst = stage + k * 256 + threadIdx.x;
            if (ok) {
                const uint32_t sa = (uint32_t)__cvta_generic_to_shared(dst);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(T + (size_t)id_k * D + sub * 4)
                             : "memory");
            } else {
                *dst = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }

    // STOP (compile-time bisect knob): 1 = scores only, 2 = + loss statistics, 3 = everything (the product)
    template <int STOP>
    __device__ __forceinline__ void compute(int64_t pass, const float4* stage, const float4& rp, const float4& q,
                                            float4* sstat, float4 (*part)[LPR]) const {
        const float4* r = stage + threadIdx.x;                      // row k of this lane's group: r[k * 256]
        const int64_t b = pass * SPB + slot;
        const bool have = b < B;
        const bool mine_ok = have && (sub % RS) == 0 && c_mine < C;
        float d[RPG];
#pragma unroll
        for (int k = 0; k < RPG; ++k) d[k] = dot4(q, r[k * 256]);
        const float p = group_sum<LPR>(dot4(q, rp));                 // positive's score, known to every group
        const float x = group_sum_multi<LPR, RPG>(d, sub);           // score of candidate c_mine (copies in RS lanes)
        if (mine_ok && pred != nullptr) pred[b * C + c_mine] = x;
        if (have && j == 0 && qout != nullptr) st4(qout + b * D + sub * 4, q);   // the sample's user row, for dI
        if (STOP == 1) {
            if (mine_ok) gout[b * C + c_mine] = x;
            return;
        }
        // ---- group-local statistics over this group's negatives ----------------------------------------
        const bool is_neg = mine_ok && c_mine > 0;
        const float mg = group_max<LPR>(is_neg ? x : -INFINITY);
        float e = 0.f, sg = 0.f;
        if (is_neg) {
            e = __expf(x - mg);
            sg = __fdividef(1.f, 1.f + __expf(x - p));
        }
        const float es = e * sg;
        const float Zg = group_sum<LPR>(e);
        const float Ag = group_sum<LPR>(es);
        const float Dg = group_sum<LPR>(es * (1.f - sg));
        if (sub == 0) sstat[slot * GPS + j] = make_float4(mg, Zg, Ag, Dg);
        __syncthreads();
        // ---- combine the GPS groups of the sample (online-softmax rescale), lanes in parallel --------------
        float m_t = -INFINITY, Z = 0.f, A = 0.f, Dp = 0.f;
        for (int t = sub; t < GPS; t += LPR) {                       // one iteration unless GPS > LPR
            const float4 st = sstat[slot * GPS + t];
            if (st.x > m_t) {                                        // rescale what was accumulated so far
                const float sc = __expf(m_t - st.x);
                Z *= sc; A *= sc; Dp *= sc;
                m_t = st.x;
            }
            const float sc2 = (st.x == -INFINITY) ? 0.f : __expf(st.x - m_t);
            Z = fmaf(st.y, sc2, Z); A = fmaf(st.z, sc2, A); Dp = fmaf(st.w, sc2, Dp);
        }
        const float M = group_max<LPR>(m_t);
        const float scl = (m_t == -INFINITY) ? 0.f : __expf(m_t - M);
        Z = group_sum<LPR>(Z * scl);
        A = group_sum<LPR>(A * scl);
        Dp = group_sum<LPR>(Dp * scl);
        const float S = (C > 1) ? __fdividef(A, Z) : 0.f;
        const bool inside = (S >= 1e-8f) && (S <= 1.f - 1e-8f);
        const float dS = inside ? -__fdividef(invB, S) : 0.f;
        const float invZ = (C > 1) ? __fdividef(1.f, Z) : 0.f;
        float gmine = 0.f;
        if (mine_ok) {
            const float ef = is_neg ? e * __expf(mg - M) : 0.f;      // e relative to the sample max
            gmine = (c_mine == 0) ? dS * Dp * invZ : dS * (ef * invZ) * ((sg - S) - sg * (1.f - sg));
            gout[b * C + c_mine] = gmine;
        }
        if (have && j == 0 && sub == 0) row_loss[b] = -logf(fminf(fmaxf(S, 1e-8f), 1.f - 1e-8f));
        if (STOP == 2) return;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < RPG; ++k) {
            const float gk = __shfl_sync(B2R_FULL_MASK, gmine, k * RS, LPR);
            fma4(acc, gk, r[k * 256]);
        }
        part[grp][sub] = acc;
        __syncthreads();
        {
            const float* pf = reinterpret_cast<const float*>(part);
            for (int o = threadIdx.x; o < SPB * D; o += 256) {
                const int sl = o / D, col = o % D;
                const int64_t bb = pass * SPB + sl;
                if (bb < B) {
                    float tot = 0.f;
                    for (int t = 0; t < GPS; ++t) tot += pf[(sl * GPS + t) * D + col];
                    dQ[bb * D + col] = tot;
                }
            }
        }
    }
};

template <int LPR, int RPG, int STOP>
__global__ void __launch_bounds__(256, (RPG <= 4) ? 4 : 3)
k_bprmf_fused(const float* __restrict__ U, const int64_t* __restrict__ uid, int64_t n_users,
              const float* __restrict__ T, const int64_t* __restrict__ ids, int64_t n_t,
              float* __restrict__ pred, float* __restrict__ gout, float* __restrict__ row_loss,
              float* __restrict__ dQ, float* __restrict__ qout, int B, int C, int GPS, int32_t* err_flag,
              float* __restrict__ loss_out, unsigned int* __restrict__ done_counter) {
    static_assert(RPG <= LPR, "ids of a group's rows are loaded one per lane");
    using P = FusedPass<LPR, RPG>;
    __shared__ float4 sstat[2][P::GPC];                   // (max, Z, A, D) per group, alternating by pass parity
    __shared__ float4 part[2][P::GPC][LPR];               // partial dQ per group
    P f;
    f.U = U; f.uid = uid; f.n_users = n_users; f.T = T; f.ids = ids; f.n_t = n_t;
    f.pred = pred; f.gout = gout; f.row_loss = row_loss; f.dQ = dQ; f.qout = qout; f.B = B; f.C = C; f.GPS = GPS; f.err_flag = err_flag;
    f.sub = threadIdx.x % LPR; f.grp = threadIdx.x / LPR;
    f.SPB = P::GPC / GPS; f.j = f.grp % GPS; f.slot = f.grp / GPS;
    f.c_load = f.j + GPS * f.sub;
    f.c_mine = f.j + GPS * (f.sub / P::RS);
    f.invB = 1.f / (float)B;
    const int64_t npass = ((int64_t)B + f.SPB - 1) / f.SPB;
    // Software pipeline over this CTA's passes: while pass p is reduced out of shared-memory stage `par`, the rows of
    // the next pass are landing in the other stage and the ids of the pass after that are in flight.
    extern __shared__ __align__(16) float4 stages[];      // [2][RPG][256]
    int64_t p = blockIdx.x;
    float4 q, rp;
    typename P::Ids idn{0, 0u, 0u};
    if (p < npass) {
        const typename P::Ids id0 = f.load_ids(p);
        f.issue_rows(p, id0, stages, rp, q);
        if (p + gridDim.x < npass) idn = f.load_ids(p + gridDim.x);
    }
    int par = 0;
    for (; p < npass; p += gridDim.x, par ^= 1) {
        const int64_t pn = p + gridDim.x;
        float4 qn = q, rpn = rp;
        if (pn < npass) {
            f.issue_rows(pn, idn, stages + (par ^ 1) * RPG * 256, rpn, qn);
            if (pn + gridDim.x < npass) idn = f.load_ids(pn + gridDim.x);
            asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        }
        f.template compute<STOP>(p, stages + par * RPG * 256, rp, q, sstat[par], part[par]);
        q = qn;
        rp = rpn;
    }
    // mean of the per-sample losses by the last CTA to finish (fixed summation order -> deterministic)
    if (loss_out != nullptr) {
        __shared__ bool last;
        __shared__ float red[256];
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
        __syncthreads();
        if (last) {
            __threadfence();
            float a = 0.f;
            for (int i = threadIdx.x; i < B; i += 256) a += __ldcg(row_loss + i);
            red[threadIdx.x] = a;
            __syncthreads();
            for (int o = 128; o > 0; o >>= 1) {
                if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                loss_out[0] = red[0] / (float)B;
                *done_counter = 0u;                      // ready for the next launch
            }
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

int b2r_bprmf_fused_ws_launch(const float* U, const int64_t* uid, int64_t n_users, const float* I, const int64_t* iid,
                              int64_t n_items, float* pred, float* grad_pred, float* row_loss, float* dQ, float* qout,
                              int B, int C, int d, int32_t* err_flag, float* loss_out, unsigned int* done_counter,
                              b2r_stream_t stream);

int b2r_bprmf_flash_launch(const float* U, const int64_t* uid, int64_t n_users, const float* I, const int64_t* iid,
                           int64_t n_items, float* pred, float* grad_pred, float* row_loss, float* dQ, float* qout,
                           int B, int C, int d, int32_t* err_flag, float* loss_out, unsigned int* done_counter,
                           const b2r::DirectPlanDev* plan_i, const b2r::DirectPlanDev* plan_u, b2r_stream_t stream);

// returns B2R_E_UNSUPPORTED (without touching the error string semantics) when the shape has no fused variant
static int fused_launch(const float* U, const int64_t* uid, int64_t n_users, const float* I, const int64_t* iid,
                        int64_t n_items, float* pred, float* grad_pred, float* row_loss, float* dQ, float* qout, int B,
                        int C, int d, int32_t* err_flag, float* loss_out, unsigned int* done_counter, b2r_stream_t stream);

extern "C" int b2r_bprmf_fused_fwd_bwd(const float* U, const int64_t* uid, int64_t n_users, const float* I,
                                       const int64_t* iid, int64_t n_items, float* pred, float* grad_pred,
                                       float* row_loss, float* dQ, int B, int C, int d, int32_t* err_flag,
                                       b2r_stream_t stream) {
    return fused_launch(U, uid, n_users, I, iid, n_items, pred, grad_pred, row_loss, dQ, nullptr, B, C, d, err_flag, nullptr,
                        nullptr, stream);
}

// same, and the mean loss is produced by the kernel itself (done_counter: a zero-initialised device word the kernel
// leaves zero again) and the gathered user rows are kept (qout [B, d]: the item-side gradient's source, so that the
// user table may be updated while the item table still is); used by the step context
int b2r_bprmf_fused_fwd_bwd_loss(const float* U, const int64_t* uid, int64_t n_users, const float* I, const int64_t* iid,
                                 int64_t n_items, float* grad_pred, float* row_loss, float* dQ, float* qout, int B, int C,
                                 int d, int32_t* err_flag, float* loss_out, unsigned int* done_counter,
                                 b2r_stream_t stream) {
    return fused_launch(U, uid, n_users, I, iid, n_items, nullptr, grad_pred, row_loss, dQ, qout, B, C, d, err_flag, loss_out,
                        done_counter, stream);
}

static int fused_launch(const float* U, const int64_t* uid, int64_t n_users, const float* I, const int64_t* iid,
                        int64_t n_items, float* pred, float* grad_pred, float* row_loss, float* dQ, float* qout, int B,
                        int C, int d, int32_t* err_flag, float* loss_out, unsigned int* done_counter, b2r_stream_t stream) {
    B2R_REQUIRE(U && uid && I && iid && grad_pred && row_loss && dQ, B2R_E_BADARG, "b2r_bprmf_fused_fwd_bwd: null pointer");
    B2R_REQUIRE(B > 0 && C > 0, B2R_E_BADARG, "b2r_bprmf_fused_fwd_bwd: B=%d C=%d", B, C);
    B2R_REQUIRE(aligned16(U) && aligned16(I) && aligned16(dQ), B2R_E_BADARG, "b2r_bprmf_fused_fwd_bwd: alignment");
    // Default: the streaming (flash) kernel of bprmf_flash.cu.  B2R_FUSED=v6 selects the CTA-per-sample kernel below,
    // B2R_FUSED=ws the whole-sample-in-shared-memory warp kernel (bprmf_fused_ws.cu) -- kept for A/B measurements.
    static const int which = [] {
        const char* e = getenv("B2R_FUSED");
        if (e && e[0] == 'v') return 1;
        if (e && e[0] == 'w') return 2;
        return 0;
    }();
    if (which == 0) {
        const int rc = b2r_bprmf_flash_launch(U, uid, n_users, I, iid, n_items, pred, grad_pred, row_loss, dQ, qout, B, C,
                                              d, err_flag, loss_out, done_counter, nullptr, nullptr, stream);
        if (rc != B2R_E_UNSUPPORTED) return rc;
    }
    if (which == 2 && d == 64 && C <= 104)
        return b2r_bprmf_fused_ws_launch(U, uid, n_users, I, iid, n_items, pred, grad_pred, row_loss, dQ, qout, B, C, d,
                                         err_flag, loss_out, done_counter, stream);
    cudaStream_t s = as_stream(stream);
    int RPG = 0, GPS = 0;
    const int GPC = d == 32 ? 32 : (d == 64 ? 16 : (d == 128 ? 8 : 0));
    if (GPC == 0 || !pick_shape(C, GPC, 8, &RPG, &GPS))
        return set_error(B2R_E_UNSUPPORTED, "b2r_bprmf_fused_fwd_bwd: no fused variant for d=%d C=%d", d, C);
    const int SPB = GPC / GPS;
    const int64_t need = ((int64_t)B + SPB - 1) / SPB;          // passes
    const int64_t cap2 = (int64_t)sm_count() * 3;               // persistent: exactly the 3 resident CTAs per SM (measured best)
    const int grid2 = (int)(need < cap2 ? need : cap2);
#define B2R_FUSED(LPR, R)                                                                                          \
    do {                                                                                                           \
        constexpr int smem = 2 * R * 256 * 16;                                                                     \
        static bool attr_done = false;                                                                             \
        if (!attr_done) {                                                                                          \
            B2R_CUDA_OK(cudaFuncSetAttribute(k_bprmf_fused<LPR, R, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
            attr_done = true;                                                                                      \
        }                                                                                                          \
        k_bprmf_fused<LPR, R, 3><<<grid2, 256, smem, s>>>(U, uid, n_users, I, iid, n_items, pred, grad_pred, row_loss, dQ, \
                                                          qout, B, C, GPS, err_flag, loss_out, done_counter);      \
    } while (0)
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
