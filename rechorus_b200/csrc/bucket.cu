// bucket.cu -- the hot-path form of the row-sparse embedding backward + optimizer (K2b):
//   1. k_bucket_count / k_bucket_scan / k_bucket_scatter: partition the batch's (row id, position) pairs into
//      buckets of 2^shift consecutive table rows (~200 pairs each) -- two passes over the ids with L2 atomics,
//      ~10x cheaper than a full device radix sort of the ids;
//   2. k_bucket_apply: one CTA per bucket loads its pairs into shared memory, sorts them by (row, position)
//      with a bitonic network, finds the run heads, and its lane groups then own one unique row each: walk the
//      row's contributions in ascending position (deterministic, whatever order the atomics of step 1 produced),
//      and apply SGD/Adam/Adagrad to w (m, v) in place -- or emit / accumulate the gradient row.
// Rows with many contributions are reduced by the whole CTA (fixed assignment + fixed combine order), buckets
// larger than the shared-memory capacity are walked row by row in position ranges -- both keep the result
// independent of scheduling.  Replaces ATen embedding_dense_backward + the dense grad zero-fill + the embedding
// part of optimizer.step() (helpers/BaseRunner.py:193,205,206).
#include <stdlib.h>

#include "common.cuh"

namespace b2r {

struct BSrc {
    const float* src;
    const float* coef;
    const int64_t* src_id;
    int64_t n;
    int32_t div;
    int32_t ld;
};

constexpr int kCap = 2048;        // pairs a CTA sorts in shared memory (16 KB)
constexpr int kLong = 48;         // rows with at least this many contributions are reduced by the whole CTA
constexpr int kBT = 256;
constexpr int kPad = 32;          // ints per bucket counter: one 128-byte line each (L2 atomics serialise per line)

__device__ __forceinline__ void b_contribution(const BSrc& s0, const BSrc& s1, uint32_t p, const float*& base,
                                               int& ld, int64_t& row, float& c) {
    const bool first = (int64_t)p < s0.n;
    const int64_t pp = first ? (int64_t)p : (int64_t)p - s0.n;
    const int div = first ? s0.div : s1.div;
    const int64_t* sid = first ? s0.src_id : s1.src_id;
    const float* cf = first ? s0.coef : s1.coef;
    int64_t r = (div == 1) ? pp : (int64_t)((uint32_t)pp / (uint32_t)div);
    if (sid != nullptr) r = sid[r];
    row = r;
    c = (cf != nullptr) ? cf[pp] : 1.f;
    base = first ? s0.src : s1.src;
    ld = first ? s0.ld : s1.ld;
}

__device__ __forceinline__ void b_optim(const b2r_optim& o, float4& w, float4& m, float4& v, const float4& gin) {
    float* wp = &w.x;
    float* mp = &m.x;
    float* vp = &v.x;
    const float* gp = &gin.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float g = fmaf(o.weight_decay, wp[i], gp[i]);
        if (o.kind == 0) {
            wp[i] = fmaf(-o.lr, g, wp[i]);
        } else if (o.kind == 1) {
            mp[i] = fmaf(o.beta1, mp[i], (1.f - o.beta1) * g);
            vp[i] = fmaf(o.beta2, vp[i], (1.f - o.beta2) * g * g);
            const float denom = sqrtf(vp[i]) / sqrtf(o.bc2) + o.eps;
            wp[i] = wp[i] - (o.lr / o.bc1) * (mp[i] / denom);
        } else {
            vp[i] = fmaf(g, g, vp[i]);
            wp[i] = wp[i] - o.lr * g / (sqrtf(vp[i]) + o.eps);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// partition
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBT)
k_bucket_count(const int64_t* __restrict__ ids, int64_t n, int64_t n_rows, int shift, int64_t ignore_id,
               int64_t ignore_n, int* __restrict__ count, int32_t* err_flag) {
    for (int64_t i = (int64_t)blockIdx.x * kBT + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBT) {
        const int64_t id = ids[i];
        if (i < ignore_n && id == ignore_id) continue;
        const int64_t key = checked_id(id, n_rows, err_flag);
        atomicAdd(&count[(key >> shift) * kPad], 1);
    }
}

// single CTA: off = exclusive scan(count); cursor = off; count = 0 (ready for the next step)
__global__ void __launch_bounds__(1024)
k_bucket_scan(int* __restrict__ count, int* __restrict__ cursor, int* __restrict__ off, int nb) {
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? count[(int64_t)i * kPad] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(B2R_FULL_MASK, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int s = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(B2R_FULL_MASK, s, o);
                if (lane >= o) s += y;
            }
            wsum[lane] = s;
        }
        __syncthreads();
        const int excl = carry + (warp > 0 ? wsum[warp - 1] : 0) + x - v;
        if (i < nb) {
            off[i] = excl;
            cursor[(int64_t)i * kPad] = excl;
            count[(int64_t)i * kPad] = 0;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry += wsum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) off[nb] = carry;
}

__global__ void __launch_bounds__(kBT)
k_bucket_scatter(const int64_t* __restrict__ ids, int64_t n, int64_t n_rows, int shift, int64_t ignore_id,
                 int64_t ignore_n, int* __restrict__ cursor, uint64_t* __restrict__ pairs) {
    for (int64_t i = (int64_t)blockIdx.x * kBT + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBT) {
        const int64_t id = ids[i];
        if (i < ignore_n && id == ignore_id) continue;
        const int64_t key = (id < 0 || id >= n_rows) ? 0 : id;
        const int slot = atomicAdd(&cursor[(key >> shift) * kPad], 1);
        pairs[slot] = ((uint64_t)key << 32) | (uint64_t)(uint32_t)i;
    }
}

// ---------------------------------------------------------------------------------------------------
// per-bucket sort + segment reduce + optimizer
// ---------------------------------------------------------------------------------------------------
template <int LPR>
struct RowIO {
    static constexpr int D = LPR * 4;
    // MODE 0: rows -> grad_rows[out_base + u], uniq_rows; MODE 1: dense += ; MODE 2: optimizer in place
    template <int MODE>
    __device__ static __forceinline__ void finish(int64_t row, const float4& acc, int sub, float4 w, float4 m, float4 v,
                                                  float* W, float* M, float* V, float* dense, const OptK& opt) {
        if (MODE == 1) {
            w.x += acc.x; w.y += acc.y; w.z += acc.z; w.w += acc.w;
            st4(dense + row * D + sub * 4, w);
        } else {
            optk_update4(opt, w, m, v, acc);
            st4(W + row * D + sub * 4, w);
            if (opt.kind == 1) st4(M + row * (opt.state_ld ? opt.state_ld : D) + sub * 4, m);
            if (opt.kind != 0) st4(V + row * (opt.state_ld ? opt.state_ld : D) + sub * 4, v);
        }
    }
};

__device__ __forceinline__ void bitonic_sort_smem(uint64_t* s, int P) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += kBT) {
                const int i = ((t / j) * (j << 1)) + (t % j);
                const int l = i + j;
                const bool up = (i & k) == 0;
                const uint64_t a = s[i], b = s[l];
                if ((a > b) == up) {
                    s[i] = b;
                    s[l] = a;
                }
            }
            __syncthreads();
        }
    }
}

// kRB = unique rows a lane group keeps in flight
template <int LPR, int MODE, int kRB>
__global__ void __launch_bounds__(kBT, (kRB <= 2) ? 3 : 2)
k_bucket_apply(const uint64_t* __restrict__ pairs, const int* __restrict__ off, int nb, BSrc s0, BSrc s1,
               float* __restrict__ W, float* __restrict__ M, float* __restrict__ V, float* __restrict__ dense,
               OptK opt) {
    constexpr int D = LPR * 4;
    constexpr int GPC = kBT / LPR;
    __shared__ uint64_t s[kCap];
    __shared__ unsigned short heads[kCap];
    __shared__ unsigned short longs[64];
    __shared__ float4 part[GPC][LPR];
    __shared__ int wsum[kBT / 32];
    __shared__ int sh_nu, sh_nlong, sh_next;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int sub = tid % LPR, grp = tid / LPR;

    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const int beg = off[b];
        const int cnt = off[b + 1] - beg;
        if (cnt == 0 || cnt > kCap) continue;      // oversize buckets: k_bucket_apply_big
        {
            // ---- load + sort ------------------------------------------------------------------
            int P = 32;
            while (P < cnt) P <<= 1;
            for (int i = tid; i < P; i += kBT) s[i] = i < cnt ? pairs[beg + i] : ~0ull;
            if (tid == 0) sh_nlong = 0;
            __syncthreads();
            bitonic_sort_smem(s, P);
            // ---- run heads -> heads[0..nu) ----------------------------------------------------
            const int E = (P + kBT - 1) / kBT;                 // consecutive elements per thread
            const int i0 = tid * E;
            int local = 0;
            for (int e = 0; e < E; ++e) {
                const int i = i0 + e;
                if (i < cnt && (i == 0 || (uint32_t)(s[i] >> 32) != (uint32_t)(s[i - 1] >> 32))) ++local;
            }
            int x = local;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(B2R_FULL_MASK, x, o);
                if (lane >= o) x += y;
            }
            if (lane == 31) wsum[warp] = x;
            __syncthreads();
            int wbase = 0;
            for (int wi = 0; wi < warp; ++wi) wbase += wsum[wi];
            int dst = wbase + x - local;
            for (int e = 0; e < E; ++e) {
                const int i = i0 + e;
                if (i < cnt && (i == 0 || (uint32_t)(s[i] >> 32) != (uint32_t)(s[i - 1] >> 32))) heads[dst++] = (unsigned short)i;
            }
            if (tid == kBT - 1) sh_nu = wbase + x;
            __syncthreads();
            const int nu = sh_nu;
            // ---- short rows: kRB consecutive unique rows per lane group, all their loads in flight together ----
            for (int u0 = grp * kRB; u0 < nu; u0 += GPC * kRB) {
                int j0[kRB], j1[kRB];
                int64_t row[kRB];
                float4 w[kRB], m[kRB], v[kRB], acc[kRB];
                int maxlen = 0;
#pragma unroll
                for (int k = 0; k < kRB; ++k) {
                    const int u = u0 + k;
                    j0[k] = j1[k] = 0;
                    row[k] = -1;
                    if (u < nu) {
                        j0[k] = heads[u];
                        j1[k] = (u + 1 < nu) ? heads[u + 1] : cnt;
                        if (j1[k] - j0[k] >= kLong) {           // deferred to the cooperative path
                            if (sub == 0) {
                                const int q = atomicAdd(&sh_nlong, 1);
                                if (q < 64) longs[q] = (unsigned short)u;
                            }
                            j1[k] = j0[k];
                        } else {
                            row[k] = (int64_t)(s[j0[k]] >> 32);
                            maxlen = max(maxlen, j1[k] - j0[k]);
                        }
                    }
                    acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row[k] >= 0) {
                        if (MODE == 2) {
                            w[k] = ld4(W + row[k] * D + sub * 4);
                            if (opt.kind == 1) m[k] = ld4(M + row[k] * (opt.state_ld ? opt.state_ld : D) + sub * 4);
                            if (opt.kind != 0) v[k] = ld4(V + row[k] * (opt.state_ld ? opt.state_ld : D) + sub * 4);
                        } else {
                            w[k] = ld4(dense + row[k] * D + sub * 4);
                        }
                    }
                }
                for (int t = 0; t < maxlen; ++t) {
#pragma unroll
                    for (int k = 0; k < kRB; ++k) {
                        if (j0[k] + t < j1[k]) {
                            const float* base;
                            int ld;
                            int64_t r;
                            float c;
                            b_contribution(s0, s1, (uint32_t)s[j0[k] + t], base, ld, r, c);
                            fma4(acc[k], c, ld4(base + r * ld + sub * 4));
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < kRB; ++k)
                    if (row[k] >= 0) RowIO<LPR>::template finish<MODE>(row[k], acc[k], sub, w[k], m[k], v[k], W, M, V, dense, opt);
            }
            __syncthreads();
            // ---- long rows: the whole CTA reduces one row at a time ----------------------------
            const int nlong = min(sh_nlong, 64);
            if (nlong > 0) {
                // deterministic order of the long rows regardless of which group found them first
                if (tid == 0) {
                    for (int a = 1; a < nlong; ++a) {
                        const unsigned short key = longs[a];
                        int c2 = a - 1;
                        while (c2 >= 0 && longs[c2] > key) { longs[c2 + 1] = longs[c2]; --c2; }
                        longs[c2 + 1] = key;
                    }
                }
                __syncthreads();
                for (int q = 0; q < nlong; ++q) {
                    const int u = longs[q];
                    const int j0 = heads[u];
                    const int j1 = (u + 1 < nu) ? heads[u + 1] : cnt;
                    const int64_t row = (int64_t)(s[j0] >> 32);
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int j = j0 + grp; j < j1; j += GPC) {
                        const float* base;
                        int ld;
                        int64_t r;
                        float c;
                        b_contribution(s0, s1, (uint32_t)s[j], base, ld, r, c);
                        fma4(acc, c, ld4(base + r * ld + sub * 4));
                    }
                    part[grp][sub] = acc;
                    __syncthreads();
                    if (grp == 0) {
                        float4 tot = part[0][sub];
                        for (int g2 = 1; g2 < GPC; ++g2) {
                            const float4 y = part[g2][sub];
                            tot.x += y.x; tot.y += y.y; tot.z += y.z; tot.w += y.w;
                        }
                        float4 w, m, v;
                        if (MODE == 2) {
                            w = ld4(W + row * D + sub * 4);
                            if (opt.kind == 1) m = ld4(M + row * (opt.state_ld ? opt.state_ld : D) + sub * 4);
                            if (opt.kind != 0) v = ld4(V + row * (opt.state_ld ? opt.state_ld : D) + sub * 4);
                        } else {
                            w = ld4(dense + row * D + sub * 4);
                        }
                        RowIO<LPR>::template finish<MODE>(row, tot, sub, w, m, v, W, M, V, dense, opt);
                    }
                    __syncthreads();
                }
                // more than 64 long rows in one bucket: the rest (never recorded) are handled below by rescanning
                if (sh_nlong > 64) {
                    for (int u = 0; u < nu; ++u) {
                        const int j0 = heads[u];
                        const int j1 = (u + 1 < nu) ? heads[u + 1] : cnt;
                        if (j1 - j0 < kLong) continue;
                        bool seen = false;
                        for (int q = 0; q < 64; ++q) seen |= (longs[q] == u);
                        if (seen) continue;
                        const int64_t row = (int64_t)(s[j0] >> 32);
                        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int j = j0 + grp; j < j1; j += GPC) {
                            const float* base;
                            int ld;
                            int64_t r;
                            float c;
                            b_contribution(s0, s1, (uint32_t)s[j], base, ld, r, c);
                            fma4(acc, c, ld4(base + r * ld + sub * 4));
                        }
                        part[grp][sub] = acc;
                        __syncthreads();
                        if (grp == 0) {
                            float4 tot = part[0][sub];
                            for (int g2 = 1; g2 < GPC; ++g2) {
                                const float4 y = part[g2][sub];
                                tot.x += y.x; tot.y += y.y; tot.z += y.z; tot.w += y.w;
                            }
                            float4 w, m, v;
                            if (MODE == 2) {
                                w = ld4(W + row * D + sub * 4);
                                if (opt.kind == 1) m = ld4(M + row * (opt.state_ld ? opt.state_ld : D) + sub * 4);
                                if (opt.kind != 0) v = ld4(V + row * (opt.state_ld ? opt.state_ld : D) + sub * 4);
                            } else {
                                w = ld4(dense + row * D + sub * 4);
                            }
                            RowIO<LPR>::template finish<MODE>(row, tot, sub, w, m, v, W, M, V, dense, opt);
                        }
                        __syncthreads();
                    }
                }
            }
            __syncthreads();
        }
    }
}

// buckets that do not fit the shared-memory sort (cnt > kCap): rare (hot rows / tiny tables); kept out of the
// main kernel so that one stays lean in registers
template <int LPR, int MODE>
__global__ void __launch_bounds__(kBT)
k_bucket_apply_big(const uint64_t* __restrict__ pairs, const int* __restrict__ off, int nb, BSrc s0, BSrc s1,
                   float* __restrict__ W, float* __restrict__ M, float* __restrict__ V, float* __restrict__ dense,
                   OptK opt) {
    constexpr int D = LPR * 4;
    constexpr int GPC = kBT / LPR;
    __shared__ uint64_t s[kCap];
    __shared__ float4 part[GPC][LPR];
    __shared__ int wsum[kBT / 32];
    __shared__ int sh_next;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int sub = tid % LPR, grp = tid / LPR;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const int beg = off[b];
        const int cnt = off[b + 1] - beg;
        if (cnt <= kCap) continue;
        {
            // ---- oversize bucket: row by row (ascending), each row in ascending position ranges that fit ----
            uint32_t last_key = 0;
            bool first_round = true;
            for (;;) {
                // next row id: smallest key > last_key (or any key in the first round)
                uint32_t kmin = 0xffffffffu;
                for (int i = tid; i < cnt; i += kBT) {
                    const uint32_t k = (uint32_t)(pairs[beg + i] >> 32);
                    if ((first_round || k > last_key) && k < kmin) kmin = k;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) kmin = min(kmin, __shfl_xor_sync(B2R_FULL_MASK, kmin, o));
                if (lane == 0) wsum[warp] = (int)kmin;
                __syncthreads();
                uint32_t key = 0xffffffffu;
                for (int wi = 0; wi < kBT / 32; ++wi) key = min(key, (uint32_t)wsum[wi]);
                __syncthreads();
                if (key == 0xffffffffu) break;
                first_round = false;
                last_key = key;
                const int64_t row = (int64_t)key;
                float4 total = make_float4(0.f, 0.f, 0.f, 0.f);      // carried by group 0 across ranges
                uint32_t lo = 0;                                      // positions >= lo still to do
                bool more = true;
                while (more) {
                    // find a range [lo, hi) of positions holding at most kCap occurrences of this key
                    uint64_t span = 0x100000000ull - lo;
                    for (;;) {
                        if (tid == 0) sh_next = 0;
                        __syncthreads();
                        const uint64_t hi = (uint64_t)lo + span;
                        int c = 0;
                        for (int i = tid; i < cnt; i += kBT) {
                            const uint64_t pr = pairs[beg + i];
                            const uint32_t p = (uint32_t)pr;
                            if ((uint32_t)(pr >> 32) == key && p >= lo && (uint64_t)p < hi) ++c;
                        }
                        if (c) atomicAdd(&sh_next, c);
                        __syncthreads();
                        const int tot_c = sh_next;
                        __syncthreads();
                        if (tot_c <= kCap) break;
                        span = (span + 1) >> 1;
                    }
                    const uint64_t hi = (uint64_t)lo + span;
                    // gather the range's positions, sort them, reduce cooperatively in fixed order
                    if (tid == 0) sh_next = 0;
                    __syncthreads();
                    for (int i = tid; i < cnt; i += kBT) {
                        const uint64_t pr = pairs[beg + i];
                        const uint32_t p = (uint32_t)pr;
                        if ((uint32_t)(pr >> 32) == key && p >= lo && (uint64_t)p < hi) s[atomicAdd(&sh_next, 1)] = pr;
                    }
                    __syncthreads();
                    const int m_here = sh_next;
                    int P = 32;
                    while (P < m_here) P <<= 1;
                    for (int i = m_here + tid; i < P; i += kBT) s[i] = ~0ull;
                    __syncthreads();
                    bitonic_sort_smem(s, P);
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int j = grp; j < m_here; j += GPC) {
                        const float* base;
                        int ld;
                        int64_t r;
                        float c;
                        b_contribution(s0, s1, (uint32_t)s[j], base, ld, r, c);
                        fma4(acc, c, ld4(base + r * ld + sub * 4));
                    }
                    part[grp][sub] = acc;
                    __syncthreads();
                    if (grp == 0) {
                        for (int g2 = 0; g2 < GPC; ++g2) {
                            const float4 y = part[g2][sub];
                            total.x += y.x; total.y += y.y; total.z += y.z; total.w += y.w;
                        }
                    }
                    __syncthreads();
                    more = hi < 0x100000000ull;
                    lo = (uint32_t)hi;
                }
                if (grp == 0) {
                    float4 w, m, v;
                    if (MODE == 2) {
                        w = ld4(W + row * D + sub * 4);
                        if (opt.kind == 1) m = ld4(M + row * (opt.state_ld ? opt.state_ld : D) + sub * 4);
                        if (opt.kind != 0) v = ld4(V + row * (opt.state_ld ? opt.state_ld : D) + sub * 4);
                    } else {
                        w = ld4(dense + row * D + sub * 4);
                    }
                    RowIO<LPR>::template finish<MODE>(row, total, sub, w, m, v, W, M, V, dense, opt);
                }
                __syncthreads();
            }
        }
    }
}

struct BucketGeom {
    int shift, nb;
};

static BucketGeom bucket_geom(int64_t n, int64_t n_rows) {
    int64_t target = n / 192;                      // ~192 pairs per bucket ...
    if (target < 256) target = n / 32;             // ... but at least a few hundred buckets for small batches
    if (target < 1) target = 1;
    if (target > 32768) target = 32768;
    int64_t rows_per = (n_rows + target - 1) / target;
    int shift = 0;
    while (((int64_t)1 << shift) < rows_per) ++shift;
    BucketGeom g;
    g.shift = shift;
    g.nb = (int)((n_rows + ((int64_t)1 << shift) - 1) >> shift);
    return g;
}

struct BucketLayout {
    size_t count, cursor, off, pairs, total;
};

static BucketLayout bucket_layout(int64_t n, int64_t n_rows) {
    const BucketGeom g = bucket_geom(n, n_rows);
    BucketLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t r = o;
        o += align_up(bytes, 256);
        return r;
    };
    L.count = take((size_t)g.nb * 4 * kPad);
    L.cursor = take((size_t)g.nb * 4 * kPad);
    L.off = take((size_t)(g.nb + 1) * 4);
    L.pairs = take((size_t)n * 8);
    L.total = o;
    return L;
}

static BSrc to_bsrc(const b2r_grad_source* s, int d) {
    BSrc r{nullptr, nullptr, nullptr, 0, 1, d};
    if (s) {
        r.src = s->src;
        r.coef = s->coef;
        r.src_id = s->src_id;
        r.n = s->n;
        r.div = s->div < 1 ? 1 : s->div;
        r.ld = s->ld > 0 ? s->ld : d;
    }
    return r;
}

}  // namespace b2r

using namespace b2r;

extern "C" size_t b2r_bucket_workspace_bytes(int64_t n, int64_t n_rows) {
    if (n <= 0 || n > 0x7fffffff || n_rows <= 0 || n_rows >= 0xffffffffLL) return 0;
    return bucket_layout(n, n_rows).total;
}

extern "C" int b2r_bucket_workspace_init(void* ws, size_t ws_bytes, int64_t n, int64_t n_rows, b2r_stream_t stream) {
    B2R_REQUIRE(ws, B2R_E_BADARG, "b2r_bucket_workspace_init: null pointer");
    const size_t need = b2r_bucket_workspace_bytes(n, n_rows);
    B2R_REQUIRE(need != 0 && ws_bytes >= need, B2R_E_WORKSPACE, "b2r_bucket_workspace_init: workspace %zu < %zu", ws_bytes,
                need);
    const BucketLayout L = bucket_layout(n, n_rows);
    // the bucket counters must be zero on entry to b2r_bucket_partition; the scan leaves them zero again
    B2R_CUDA_OK(cudaMemsetAsync(static_cast<char*>(ws) + L.count, 0, L.cursor - L.count, as_stream(stream)));
    return 0;
}

extern "C" int b2r_bucket_partition(const int64_t* ids, int64_t n, int64_t n_rows, int64_t ignore_id, int64_t ignore_n,
                                    void* ws, size_t ws_bytes, int32_t* err_flag, b2r_stream_t stream) {
    B2R_REQUIRE(ids && ws, B2R_E_BADARG, "b2r_bucket_partition: null pointer");
    const size_t need = b2r_bucket_workspace_bytes(n, n_rows);
    B2R_REQUIRE(need != 0, B2R_E_UNSUPPORTED, "b2r_bucket_partition: n=%lld n_rows=%lld unsupported", (long long)n,
                (long long)n_rows);
    B2R_REQUIRE(ws_bytes >= need, B2R_E_WORKSPACE, "b2r_bucket_partition: workspace %zu < %zu", ws_bytes, need);
    B2R_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255u) == 0, B2R_E_BADARG, "workspace must be 256-byte aligned");
    cudaStream_t s = as_stream(stream);
    const BucketGeom g = bucket_geom(n, n_rows);
    const BucketLayout L = bucket_layout(n, n_rows);
    char* base = static_cast<char*>(ws);
    int* count = reinterpret_cast<int*>(base + L.count);
    int* cursor = reinterpret_cast<int*>(base + L.cursor);
    int* off = reinterpret_cast<int*>(base + L.off);
    uint64_t* pairs = reinterpret_cast<uint64_t*>(base + L.pairs);
    int grid = (int)((n + kBT * 4 - 1) / (kBT * 4));
    const int cap = sm_count() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    k_bucket_count<<<grid, kBT, 0, s>>>(ids, n, n_rows, g.shift, ignore_id, ignore_n, count, err_flag);
    B2R_LAUNCH_OK("k_bucket_count");
    k_bucket_scan<<<1, 1024, 0, s>>>(count, cursor, off, g.nb);
    B2R_LAUNCH_OK("k_bucket_scan");
    k_bucket_scatter<<<grid, kBT, 0, s>>>(ids, n, n_rows, g.shift, ignore_id, ignore_n, cursor, pairs);
    B2R_LAUNCH_OK("k_bucket_scatter");
    return 0;
}

extern "C" int b2r_bucket_apply(const void* ws, int64_t n, int64_t n_rows, int d, const b2r_grad_source* s0,
                                const b2r_grad_source* s1, int mode, float* dense, float* W, float* m, float* v,
                                const b2r_optim* opt, b2r_stream_t stream) {
    B2R_REQUIRE(ws && s0 && s0->src, B2R_E_BADARG, "b2r_bucket_apply: null pointer");
    B2R_REQUIRE(s0->n + (s1 ? s1->n : 0) == n, B2R_E_BADARG, "b2r_bucket_apply: sources cover %lld of %lld positions",
                (long long)(s0->n + (s1 ? s1->n : 0)), (long long)n);
    B2R_REQUIRE(d == 32 || d == 64 || d == 128, B2R_E_UNSUPPORTED, "b2r_bucket_apply: d=%d (have 32, 64, 128)", d);
    b2r_optim o{};
    if (mode == 1) {
        B2R_REQUIRE(dense, B2R_E_BADARG, "b2r_bucket_apply: mode 1 needs dense");
    } else if (mode == 2) {
        B2R_REQUIRE(W && opt, B2R_E_BADARG, "b2r_bucket_apply: mode 2 needs W and opt");
        o = *opt;
        B2R_REQUIRE(o.kind >= 0 && o.kind <= 2, B2R_E_BADARG, "b2r_bucket_apply: optimizer kind %d", o.kind);
        B2R_REQUIRE(o.kind != 1 || (m && v), B2R_E_BADARG, "b2r_bucket_apply: Adam needs m and v");
        B2R_REQUIRE(o.kind != 2 || v, B2R_E_BADARG, "b2r_bucket_apply: Adagrad needs v");
    } else {
        return set_error(B2R_E_BADARG, "b2r_bucket_apply: mode %d (1 = dense +=, 2 = optimizer)", mode);
    }
    cudaStream_t s = as_stream(stream);
    const BucketGeom g = bucket_geom(n, n_rows);
    const BucketLayout L = bucket_layout(n, n_rows);
    const char* base = static_cast<const char*>(ws);
    const int* off = reinterpret_cast<const int*>(base + L.off);
    const uint64_t* pairs = reinterpret_cast<const uint64_t*>(base + L.pairs);
    const BSrc a = to_bsrc(s0, d), b = to_bsrc(s1, d);
    const int cap = sm_count() * 8;
    const int grid = g.nb < cap ? g.nb : cap;
    const int big_grid = g.nb < sm_count() ? g.nb : sm_count();
    const OptK ok = make_optk(o);
    static int rb = -1;                    // tuning knob B2R_BUCKET_RB = 1 | 2 | 4 (read once)
    if (rb < 0) {
        const char* e = getenv("B2R_BUCKET_RB");
        rb = e ? atoi(e) : 2;
        if (rb != 1 && rb != 2 && rb != 4) rb = 2;
    }
#define B2R_BK(LPR, MODE)                                                                              \
    do {                                                                                               \
        if (rb == 1) k_bucket_apply<LPR, MODE, 1><<<grid, kBT, 0, s>>>(pairs, off, g.nb, a, b, W, m, v, dense, ok); \
        else if (rb == 2) k_bucket_apply<LPR, MODE, 2><<<grid, kBT, 0, s>>>(pairs, off, g.nb, a, b, W, m, v, dense, ok); \
        else k_bucket_apply<LPR, MODE, 4><<<grid, kBT, 0, s>>>(pairs, off, g.nb, a, b, W, m, v, dense, ok); \
        k_bucket_apply_big<LPR, MODE><<<big_grid, kBT, 0, s>>>(pairs, off, g.nb, a, b, W, m, v, dense, ok); \
    } while (0)
    if (mode == 1) {
        if (d == 32) B2R_BK(8, 1); else if (d == 64) B2R_BK(16, 1); else B2R_BK(32, 1);
    } else {
        if (d == 32) B2R_BK(8, 2); else if (d == 64) B2R_BK(16, 2); else B2R_BK(32, 2);
    }
#undef B2R_BK
    B2R_LAUNCH_OK("k_bucket_apply");
    return 0;
}
