// loss.cu -- BPR loss with softmax-weighted negatives and its closed-form gradient in one pass
// (replaces the ~9 eager kernels + autograd nodes of models/BaseModel.py:182-185).
//
//   p = pred[b,0]; n_j = pred[b,j] (j>=1); w = softmax(n); s_j = sigmoid(p - n_j); S = sum_j w_j s_j
//   loss = -(1/B) sum_b log(clamp(S_b, 1e-8, 1-1e-8))
//   dS/dp = sum_j w_j s_j (1 - s_j);   dS/dn_j = -w_j s_j (1 - s_j) + w_j (s_j - S)   (weights NOT detached)
//   dloss/dS = -1/(B S) inside the clamp window, 0 outside.
// The reference shifts the softmax by the max over the whole negative block (a scalar that cancels); the
// kernel shifts by the row max, which is the same function without the cross-row underflow hazard.
// The tensor is tiny ([B,C] floats): the kernel is launch/latency-bound, one warp per sample.
#include "common.cuh"

namespace b2r {

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256)
k_bpr_loss_rows(const float* __restrict__ pred, float* __restrict__ grad, float* __restrict__ row_loss,
                int B, int C) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const float* x = pred + (int64_t)b * C;
    const float p = x[0];
    float mx = -INFINITY;
    for (int c = 1 + lane; c < C; c += 32) mx = fmaxf(mx, x[c]);
    mx = warp_max(mx);
    float Z = 0.f, A = 0.f, Dp = 0.f;     // sum e, sum e*s, sum e*s*(1-s)
    for (int c = 1 + lane; c < C; c += 32) {
        const float n = x[c];
        const float e = expf(n - mx);
        const float s = sigmoidf_acc(p - n);
        Z += e;
        A = fmaf(e, s, A);
        Dp = fmaf(e * s, 1.f - s, Dp);
    }
    Z = warp_sum(Z);
    A = warp_sum(A);
    Dp = warp_sum(Dp);
    const float S = (C > 1) ? A / Z : 0.f;
    const bool inside = (S >= 1e-8f) && (S <= 1.f - 1e-8f);
    const float Sc = fminf(fmaxf(S, 1e-8f), 1.f - 1e-8f);
    if (lane == 0) row_loss[b] = -logf(Sc);
    if (grad == nullptr) return;
    float* gx = grad + (int64_t)b * C;
    const float dS = inside ? -1.f / ((float)B * S) : 0.f;
    const float invZ = (C > 1) ? 1.f / Z : 0.f;
    if (lane == 0) gx[0] = dS * Dp * invZ;
    for (int c = 1 + lane; c < C; c += 32) {
        const float n = x[c];
        const float w = expf(n - mx) * invZ;
        const float s = sigmoidf_acc(p - n);
        gx[c] = dS * w * ((s - S) - s * (1.f - s));
    }
}

// fixed-order mean of row_loss -> loss_out (single CTA: thread t owns rows t, t+1024, ...; tree in smem)
__global__ void __launch_bounds__(1024)
k_mean_rows(const float* __restrict__ row_loss, float* __restrict__ out, int B) {
    __shared__ float sm[1024];
    float a = 0.f;
    for (int i = threadIdx.x; i < B; i += 1024) a += row_loss[i];
    sm[threadIdx.x] = a;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sm[0] / (float)B;
}

int launch_mean_rows(const float* row_loss, float* out, int B, cudaStream_t s) {
    k_mean_rows<<<1, 1024, 0, s>>>(row_loss, out, B);
    B2R_LAUNCH_OK("k_mean_rows");
    return 0;
}

}  // namespace b2r

using namespace b2r;

extern "C" int b2r_bpr_loss(const float* pred, float* loss_out, float* grad_pred, float* row_ws, int B, int C,
                            b2r_stream_t stream) {
    B2R_REQUIRE(pred && loss_out && row_ws, B2R_E_BADARG, "b2r_bpr_loss: null pointer");
    B2R_REQUIRE(B > 0 && C > 0, B2R_E_BADARG, "b2r_bpr_loss: need B > 0 and C > 0 (B=%d C=%d)", B, C);
    cudaStream_t s = as_stream(stream);
    k_bpr_loss_rows<<<(B + 7) / 8, 256, 0, s>>>(pred, grad_pred, row_ws, B, C);
    B2R_LAUNCH_OK("k_bpr_loss_rows");
    return launch_mean_rows(row_ws, loss_out, B, s);
}
