// linear_dw_tc.cu -- the weight gradient of nn.Linear on the 5th-generation tensor cores (tcgen05.mma kind::tf32,
// accumulator in TMEM): dW[n, k] = sum_m dY[m, n] * X[m, k], dbias[n] = sum_m dY[m, n], the autograd of
// utils/layers.py:26-28,106-107 and NeuMF.py:70 behind loss.backward() (helpers/BaseRunner.py:205).
//
// Why this one first: in a SASRec step (config 4) the weight gradients are contractions over M = B*L = 204,800 rows with
// a 64 x 64 result -- on the CUDA cores they were 27 % of the step (k_gemm<1,1>, 241 us each), more than the attention.
// As a tensor-core problem they are an outer-product accumulation: D[128 x (K+16)] += A[128 x 32] * B[(K+16) x 32]^T per
// 32-row slab, where A's rows are the N output features (rows N..127 stay zero), B's rows the K input features plus one
// row of ones (its column of D is dbias), and the MMA's reduction dimension is the batch-row index m.  The operands are
// therefore the TRANSPOSES of the row-major activations; they are transposed on the way into shared memory (each lane owns
// one batch row of a slab and writes its elements down a column of the K-major SWIZZLE_128B tile: 32 lanes, 32 distinct
// banks), split x = hi + lo (both TF32-exact) as in linear_tc.cu, and four products accumulate in one TMEM tile across the
// CTA's whole row range.  One tcgen05.ld epilogue per CTA writes a partial [N x (K+16)] tile; a second kernel adds the
// CTAs' partials in CTA order (fixed order -> deterministic).
#include "common.cuh"

namespace b2r {

constexpr int DW_THREADS = 128;
constexpr int DW_MROWS = 128;        // UMMA M (A's rows: N valid + zero rows)
constexpr int DW_SLABS = 2;          // 32-row slabs staged per iteration

__device__ __forceinline__ uint32_t dw_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t dw_desc(uint32_t smem_addr) {        // K-major, SWIZZLE_128B (see linear_tc.cu)
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)64 << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ float dw_rn_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

// element (feature row `row`, reduction column `col` in [0, 32)) of a 128-byte-row swizzled slab
__device__ __forceinline__ uint32_t dw_off(int row, int col) {
    return (uint32_t)row * 128u + (uint32_t)((((col >> 2) ^ (row & 7)) << 4) + ((col & 3) << 2));
}

__global__ void __launch_bounds__(DW_THREADS)
k_linear_dw_tc(const float* __restrict__ dY, int lddy, const float* __restrict__ relu_out, const float* __restrict__ X,
               int ldx, float* __restrict__ part, int M, int N, int K, int rows_per_cta, int tmem_cols) {
    extern __shared__ __align__(1024) unsigned char dw_raw[];
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base_sh;
    const int NB = K + 16;                                      // B's rows: K features, one row of ones, 15 zero rows
    const size_t a_slab = (size_t)DW_MROWS * 128, b_slab = (size_t)NB * 128;
    char* A_hi = reinterpret_cast<char*>(dw_raw) + ((1024u - (dw_smem_u32(dw_raw) & 1023u)) & 1023u);
    char* A_lo = A_hi + DW_SLABS * a_slab;
    char* B_hi = A_lo + DW_SLABS * a_slab;
    char* B_lo = B_hi + DW_SLABS * b_slab;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dw_smem_u32(&tmem_base_sh)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dw_smem_u32(&mma_bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // zero every operand tile once: A's rows N..127, B's rows K+1..K+15 and the lo half of the ones row never change
    {
        const size_t total16 = (2 * DW_SLABS * (a_slab + b_slab)) / 16;
        float4* z = reinterpret_cast<float4*>(A_hi);
        for (size_t i = tid; i < total16; i += DW_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    for (int e = tid; e < DW_SLABS * 32; e += DW_THREADS)       // the row of ones (hi = 1, lo = 0): row K of B
        *reinterpret_cast<float*>(B_hi + (e / 32) * b_slab + dw_off(K, e % 32)) = 1.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_sh;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(DW_MROWS >> 4) << 24);

    const int m_begin = blockIdx.x * rows_per_cta;
    const int m_end = min(M, m_begin + rows_per_cta);
    uint32_t parity = 0, acc = 0;
    for (int m0 = m_begin; m0 < m_end; m0 += DW_SLABS * 32) {
        // warp w stages slab (w % DW_SLABS) of dY (w < DW_SLABS) or X (w >= DW_SLABS): lane = batch row of the slab
        {
            const bool isA = warp < DW_SLABS;
            const int s = warp % DW_SLABS;
            const int m = m0 + s * 32 + lane;
            const bool live = m < m_end;
            const float* src = isA ? dY + (size_t)m * lddy : X + (size_t)m * ldx;
            const float* msk = (isA && relu_out != nullptr) ? relu_out + (size_t)m * lddy : nullptr;
            char* hi = (isA ? A_hi + s * a_slab : B_hi + s * b_slab);
            char* lo = (isA ? A_lo + s * a_slab : B_lo + s * b_slab);
            const int nf = isA ? N : K;
            // 8 x 128-bit loads (32 features) of the lane's batch row in flight before the first use: the loop is bound by
            // load latency otherwise (one CTA stages only 64 rows per MMA round)
            for (int f0 = 0; f0 < nf; f0 += 32) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (live && f0 + 4 * u < nf) v[u] = ld4(src + f0 + 4 * u);
                }
                if (msk != nullptr) {                                // ReLU backward: dY counts where the saved output > 0
                    float4 mk[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        mk[u] = make_float4(1.f, 1.f, 1.f, 1.f);
                        if (live && f0 + 4 * u < nf) mk[u] = ld4(msk + f0 + 4 * u);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        v[u].x = mk[u].x > 0.f ? v[u].x : 0.f;
                        v[u].y = mk[u].y > 0.f ? v[u].y : 0.f;
                        v[u].z = mk[u].z > 0.f ? v[u].z : 0.f;
                        v[u].w = mk[u].w > 0.f ? v[u].w : 0.f;
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (f0 + 4 * u < nf) {
                        const float pv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float h = dw_rn_tf32(pv[i]);
                            const uint32_t off = dw_off(f0 + 4 * u + i, lane);
                            *reinterpret_cast<float*>(hi + off) = h;
                            *reinterpret_cast<float*>(lo + off) = dw_rn_tf32(pv[i] - h);
                        }
                    }
                }
            }
            if (!isA) {                                              // batch rows beyond the range must not count in dbias
                *reinterpret_cast<float*>(hi + dw_off(K, lane)) = live ? 1.f : 0.f;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int prod = 0; prod < 4; ++prod) {
                const char* Ab = (prod < 2) ? A_hi : A_lo;           // hi*hi, hi*lo, lo*hi, lo*lo
                const char* Bb = (prod & 1) ? B_lo : B_hi;
                for (int s = 0; s < DW_SLABS; ++s) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {                    // 4 MMAs of 8 batch rows (32 bytes) per slab
                        const uint64_t ad = dw_desc(dw_smem_u32(Ab + s * a_slab) + k * 32);
                        const uint64_t bd = dw_desc(dw_smem_u32(Bb + s * b_slab) + k * 32);
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                            "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                            : "memory");
                        acc = 1;
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             dw_smem_u32(&mma_bar))
                         : "memory");
        }
        // bounded wait for the MMAs to have read the tiles (a wrong descriptor must trap, not hang)
        {
            const uint32_t addr = dw_smem_u32(&mma_bar);
            uint32_t done = 0;
            for (uint32_t spin = 0; spin < (1u << 24) && !done; ++spin) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                    : "=r"(done)
                    : "r"(addr), "r"(parity)
                    : "memory");
            }
            if (!done) __trap();
        }
        parity ^= 1;
        acc = 1;                                                     // uniform across threads (only thread 0 uses it)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // epilogue: TMEM lane = output feature n (warps 0..N/32-1 hold the valid lanes), columns 0..K = dW row, column K = dbias
    float* out = part + (size_t)blockIdx.x * N * (K + 1);
    if (m_begin < m_end) {
        const int n = warp * 32 + lane;
        for (int c0 = 0; c0 < NB; c0 += 16) {
            uint32_t r[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (n < N) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (c0 + i <= K) out[(size_t)n * (K + 1) + c0 + i] = __uint_as_float(r[i]);
            }
        }
    } else {
        for (int i = tid; i < N * (K + 1); i += DW_THREADS) out[i] = 0.f;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
    }
}

// dW[n, k] = sum_c part[c][n][k], dbias[n] = sum_c part[c][n][K]  (c ascending)
__global__ void __launch_bounds__(256)
k_linear_dw_reduce(const float* __restrict__ part, int ctas, int N, int K, float* __restrict__ dW, float* __restrict__ dbias) {
    const int total = N * (K + 1);
    for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
        float a = 0.f;
        for (int c = 0; c < ctas; ++c) a += part[(size_t)c * total + i];
        const int n = i / (K + 1), k = i % (K + 1);
        if (k < K) dW[(size_t)n * K + k] = a;
        else if (dbias != nullptr) dbias[n] = a;
    }
}

static int dw_ctas(int64_t M) {
    int64_t slabs = (M + DW_SLABS * 32 - 1) / (DW_SLABS * 32);
    int64_t c = (int64_t)sm_count() * 2;
    if (c > slabs) c = slabs;
    return (int)(c < 1 ? 1 : c);
}

}  // namespace b2r

using namespace b2r;

extern "C" size_t b2r_linear_bwd_weight_tc_workspace_bytes(int64_t M, int N, int K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    return align_up((size_t)dw_ctas(M) * N * (K + 1) * 4, 256);
}

// returns B2R_E_UNSUPPORTED for shapes outside the tensor-core kernel's class (callers fall back to b2r_linear_bwd_weight)
extern "C" int b2r_linear_bwd_weight_tc(const float* dY, int lddy, const float* relu_out, const float* X, int ldx, float* dW,
                                        float* dbias, int64_t M, int N, int K, void* ws, size_t ws_bytes,
                                        b2r_stream_t stream) {
    B2R_REQUIRE(dY && X && dW && ws, B2R_E_BADARG, "b2r_linear_bwd_weight_tc: null pointer");
    if (!(N % 4 == 0 && N >= 4 && N <= 128 && K % 16 == 0 && K >= 16 && K <= 240 && M > 0 && M <= 0x7fffffff &&
          lddy % 4 == 0 && ldx % 4 == 0 && lddy >= N && ldx >= K && aligned16(dY) && aligned16(X) &&
          (relu_out == nullptr || aligned16(relu_out))))
        return set_error(B2R_E_UNSUPPORTED, "b2r_linear_bwd_weight_tc: shape M=%lld N=%d K=%d outside the kernel's class",
                         (long long)M, N, K);
    B2R_REQUIRE(ws_bytes >= b2r_linear_bwd_weight_tc_workspace_bytes(M, N, K), B2R_E_WORKSPACE,
                "b2r_linear_bwd_weight_tc: workspace too small");
    const size_t smem = (size_t)2 * DW_SLABS * DW_MROWS * 128 + (size_t)2 * DW_SLABS * (K + 16) * 128 + 1024;
    if (smem > 200 * 1024) return set_error(B2R_E_UNSUPPORTED, "b2r_linear_bwd_weight_tc: %zu B of shared memory needed", smem);
    const int ctas = dw_ctas(M);
    const int64_t slabs = (M + DW_SLABS * 32 - 1) / (DW_SLABS * 32);
    const int rows_per_cta = (int)((slabs + ctas - 1) / ctas) * DW_SLABS * 32;
    int cols = 32;
    while (cols < K + 16) cols <<= 1;
    cudaStream_t s = as_stream(stream);
    static size_t attr = 0;
    if (smem > attr) {
        B2R_CUDA_OK(cudaFuncSetAttribute(k_linear_dw_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    float* part = static_cast<float*>(ws);
    k_linear_dw_tc<<<ctas, DW_THREADS, smem, s>>>(dY, lddy, relu_out, X, ldx, part, (int)M, N, K, rows_per_cta, cols);
    B2R_LAUNCH_OK("k_linear_dw_tc");
    const int total = N * (K + 1);
    k_linear_dw_reduce<<<(total + 255) / 256, 256, 0, s>>>(part, ctas, N, K, dW, dbias);
    B2R_LAUNCH_OK("k_linear_dw_reduce");
    return 0;
}
