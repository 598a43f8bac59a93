// listwise.cu -- the list-wise ranking losses of the impression models: value and closed-form
// gradient of models/BaseImpressionModel.py:44-128 (ImpressionModel.loss) in two launches.
//
// Input: prediction [B, Cn] and target [B, Cn] int64 (1 = clicked, 0 = shown, -1 = padding); the first `max_pos` columns
// are the positive slots, the rest the negative slots (BaseImpressionModel.py:52-58).  P = valid positive columns,
// N = valid negative columns of a row, x = its predictions.
//   kind 0 "BPR" (reweight between the sigmoid and the log, :82-85), 1 "BPR...after" (:73-75), 2 "BPR...before" (:76-78),
//        `hard` = positives weighted towards LOW scores (:66-68);   pw = softmax(+-x) over P, nw = softmax(x) over N
//        0: L = -log sum_i pw_i sum_j nw_j sigmoid(x_i - x_j)
//        1: L = sum_i pw_i sum_j nw_j softplus(-(x_i - x_j))
//        2: L = sum_{i in P} softplus(-pw_i (x_i - sum_j nw_j x_j)) + (Cn - |P|) ln 2
//   kind 3 "listnet" (:88-97), 4 "softmaxCE" (:99-110), 5 "attention_rank" (:112-128): cross-entropy forms, each row
//        weighted by have_neg_b * B / sum_b have_neg_b (have_neg = column max_pos is not padding).
// The batch loss is the mean over rows.  ("BPR...simple", :79-81, returns a vector in the reference and cannot be
// back-propagated there; it is not provided.)  Softmax shifts use the row maximum (the reference subtracts a batch-wide
// maximum, which cancels).  One warp per row, row staged in shared memory; every sum has a fixed order.
#include "common.cuh"

namespace b2r {

constexpr int kLwWarps = 4;

__device__ __forceinline__ float lw_softplus(float z) {                 // log(1 + e^z), stable
    return fmaxf(z, 0.f) + log1pf(expf(-fabsf(z)));
}
__device__ __forceinline__ float lw_sigmoid(float z) { return 1.f / (1.f + expf(-z)); }

__device__ __forceinline__ float warp_sum_f(float v) { return warp_sum(v); }

// softmax weights over the columns selected by sel (flag array: 1 = member): w[k] = e^{sgn*x[k] - max} / Z, 0 elsewhere
__device__ __forceinline__ void lw_softmax(const float* x, const float* sel, float* w, int Cn, int lane, float sgn) {
    float mx = -INFINITY;
    for (int k = lane; k < Cn; k += 32)
        if (sel[k] != 0.f) mx = fmaxf(mx, sgn * x[k]);
    mx = warp_max(mx);
    float z = 0.f;
    for (int k = lane; k < Cn; k += 32) {
        const float e = sel[k] != 0.f ? expf(sgn * x[k] - mx) : 0.f;
        w[k] = e;
        z += e;
    }
    z = warp_sum_f(z);
    const float iz = 1.f / z;                                            // no member: 0 * inf = NaN, like the reference
    __syncwarp();
    for (int k = lane; k < Cn; k += 32) w[k] = w[k] * iz;
    __syncwarp();
}

__global__ void __launch_bounds__(kLwWarps * 32)
k_listwise_rows(const float* __restrict__ pred, const int64_t* __restrict__ target, int B, int Cn, int max_pos, int kind,
                int hard, float* __restrict__ row_loss, float* __restrict__ have_neg, float* __restrict__ grad) {
    extern __shared__ float lw_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* x = lw_smem + (size_t)warp * 6 * Cn;
    float* vp = x + Cn;        // 1 on valid positive columns
    float* vn = vp + Cn;       // 1 on valid negative columns
    float* pw = vn + Cn;       // weights / probabilities
    float* nw = pw + Cn;
    float* gr = nw + Cn;
    const int b = blockIdx.x * kLwWarps + warp;
    if (b >= B) return;
    const float* xp = pred + (int64_t)b * Cn;
    const int64_t* tp = target + (int64_t)b * Cn;
    float n_valid = 0.f, pos_len = 0.f;
    for (int k = lane; k < Cn; k += 32) {
        const int64_t t = tp[k];
        const bool v = t != -1;
        x[k] = xp[k];
        vp[k] = (v && k < max_pos) ? 1.f : 0.f;
        vn[k] = (v && k >= max_pos) ? 1.f : 0.f;
        gr[k] = 0.f;
        n_valid += v ? 1.f : 0.f;
        pos_len += (t == 1) ? 1.f : 0.f;
    }
    __syncwarp();
    const float hn = (max_pos < Cn && tp[max_pos] != -1) ? 1.f : 0.f;
    float loss = 0.f;
    if (kind <= 2) {
        lw_softmax(x, vp, pw, Cn, lane, hard ? -1.f : 1.f);
        lw_softmax(x, vn, nw, Cn, lane, 1.f);
        const float s = hard ? -1.f : 1.f;
        if (kind == 0 || kind == 1) {
            // A_i = sum_j nw_j f_ij (lanes over j, serial over i); total = sum_i pw_i A_i
            float total = 0.f;
            for (int i = 0; i < max_pos && i < Cn; ++i) {
                if (vp[i] == 0.f) continue;
                float a = 0.f;
                for (int j = max_pos + lane; j < Cn; j += 32) {
                    if (vn[j] == 0.f) continue;
                    const float dlt = x[i] - x[j];
                    a += nw[j] * (kind == 0 ? lw_sigmoid(dlt) : lw_softplus(-dlt));
                }
                a = warp_sum_f(a);
                total += pw[i] * a;
                if (lane == 0) gr[i] = a;                               // park A_i in the gradient slot
            }
            __syncwarp();
            loss = kind == 0 ? -logf(total) : total;
            const float outer = kind == 0 ? -1.f / total : 1.f;          // dL/dS (kind 0) or 1 (kind 1)
            // positives: pw_i * sum_j nw_j f'_ij + s * pw_i (A_i - total);  f' = sig(1-sig) (0) or -(1-sig) (1)
            for (int i = 0; i < max_pos && i < Cn; ++i) {
                if (vp[i] == 0.f) continue;
                float a = 0.f;
                for (int j = max_pos + lane; j < Cn; j += 32) {
                    if (vn[j] == 0.f) continue;
                    const float sg = lw_sigmoid(x[i] - x[j]);
                    a += nw[j] * (kind == 0 ? sg * (1.f - sg) : -(1.f - sg));
                }
                a = warp_sum_f(a);
                const float Ai = gr[i];
                __syncwarp();
                if (lane == 0) gr[i] = outer * (pw[i] * a + s * pw[i] * (Ai - total));
            }
            __syncwarp();
            // negatives: lanes over j, serial over i
            for (int j = max_pos + lane; j < Cn; j += 32) {
                if (vn[j] == 0.f) continue;
                float dsum = 0.f, Bj = 0.f;
                for (int i = 0; i < max_pos; ++i) {
                    if (vp[i] == 0.f) continue;
                    const float dlt = x[i] - x[j];
                    const float sg = lw_sigmoid(dlt);
                    dsum += pw[i] * (kind == 0 ? sg * (1.f - sg) : -(1.f - sg));
                    Bj += pw[i] * (kind == 0 ? sg : lw_softplus(-dlt));
                }
                gr[j] = outer * (-nw[j] * dsum + nw[j] * (Bj - total));
            }
        } else {
            float m = 0.f, npos = 0.f;
            for (int j = lane; j < Cn; j += 32) {
                m += nw[j] * (vn[j] != 0.f ? x[j] : 0.f);
                npos += vp[j];
            }
            m = warp_sum_f(m);
            npos = warp_sum_f(npos);
            float G = 0.f, T = 0.f, l = 0.f;
            for (int i = lane; i < Cn; i += 32) {
                if (vp[i] == 0.f) continue;
                const float D = x[i] - m, z = pw[i] * D;
                const float t = lw_sigmoid(-z);
                l += lw_softplus(-z);
                G += t * pw[i] * D;
                T += t * pw[i];
            }
            G = warp_sum_f(G);
            T = warp_sum_f(T);
            loss = warp_sum_f(l) + ((float)Cn - npos) * 0.69314718055994531f;
            for (int k = lane; k < Cn; k += 32) {
                if (vp[k] != 0.f) {
                    const float D = x[k] - m, t = lw_sigmoid(-pw[k] * D);
                    gr[k] = -t * pw[k] - s * pw[k] * (t * D - G);
                } else if (vn[k] != 0.f) {
                    gr[k] = T * nw[k] * (1.f + x[k] - m);
                }
            }
        }
    } else {
        // cross-entropy forms.  valid = vp + vn; all = every column (listnet's prediction softmax includes the padding)
        float* valid = nw;                                               // reuse: 1 on valid columns
        for (int k = lane; k < Cn; k += 32) valid[k] = vp[k] + vn[k];
        __syncwarp();
        if (kind == 4) {
            lw_softmax(x, valid, pw, Cn, lane, 1.f);
            float l = 0.f, npv = 0.f;
            for (int k = lane; k < Cn; k += 32) {
                if (vp[k] != 0.f) {
                    l -= logf(pw[k]);
                    npv += 1.f;
                }
            }
            l = warp_sum_f(l);
            npv = warp_sum_f(npv);
            pos_len = warp_sum_f(pos_len);
            loss = l / pos_len;
            for (int k = lane; k < Cn; k += 32)
                if (valid[k] != 0.f) gr[k] = (npv * pw[k] - vp[k]) / pos_len;
        } else {
            // target softmax over the valid columns: weights e^{t_k}; tw kept in vp[] after this point
            float zt = 0.f;
            for (int k = lane; k < Cn; k += 32) zt += valid[k] != 0.f ? expf((float)tp[k] - 1.f) : 0.f;
            zt = warp_sum_f(zt);
            __syncwarp();
            for (int k = lane; k < Cn; k += 32) vp[k] = valid[k] != 0.f ? expf((float)tp[k] - 1.f) / zt : 0.f;
            __syncwarp();
            if (kind == 3) {
                for (int k = lane; k < Cn; k += 32) vn[k] = 1.f;         // softmax of the predictions over ALL columns
                __syncwarp();
                lw_softmax(x, vn, pw, Cn, lane, 1.f);
                float l = 0.f;
                for (int k = lane; k < Cn; k += 32)
                    if (valid[k] != 0.f) l -= vp[k] * logf(pw[k]);
                loss = warp_sum_f(l);
                for (int k = lane; k < Cn; k += 32) gr[k] = pw[k] - vp[k];
            } else {
                lw_softmax(x, valid, pw, Cn, lane, 1.f);
                float l = 0.f, cp = 0.f;
                for (int k = lane; k < Cn; k += 32) {
                    if (valid[k] == 0.f) continue;
                    l -= vp[k] * logf(pw[k]);
                    if (pw[k] != 1.f) {
                        l -= (1.f - vp[k]) * logf(1.f - pw[k]);
                        cp += (1.f - vp[k]) / (1.f - pw[k]) * pw[k];
                    }
                }
                loss = warp_sum_f(l);
                cp = warp_sum_f(cp);
                for (int k = lane; k < Cn; k += 32) {
                    if (valid[k] == 0.f) continue;
                    const float c = pw[k] != 1.f ? (1.f - vp[k]) / (1.f - pw[k]) : 0.f;
                    gr[k] = (pw[k] - vp[k]) + (c * pw[k] - pw[k] * cp);
                }
            }
        }
    }
    __syncwarp();
    if (lane == 0) {
        row_loss[b] = loss;
        have_neg[b] = hn;
    }
    if (grad != nullptr)
        for (int k = lane; k < Cn; k += 32) grad[(int64_t)b * Cn + k] = gr[k];
}

// loss = mean_b row_loss_b * f_b, grad *= f_b / B, f_b = 1 (BPR kinds) or have_neg_b * B / sum have_neg (CE kinds);
// single CTA for the scalars (fixed order), then every CTA scales its slice.
__global__ void __launch_bounds__(256)
k_listwise_finish(const float* __restrict__ row_loss, const float* __restrict__ have_neg, int B, int Cn, int kind,
                  float* __restrict__ loss_out, float* __restrict__ grad) {
    __shared__ float red[256];
    __shared__ float s_hn;
    float a = 0.f;
    for (int i = threadIdx.x; i < B; i += 256) a += have_neg[i];
    red[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) s_hn = red[0];
    __syncthreads();
    const float hn_sum = s_hn;
    const float fB = (float)B;
    if (blockIdx.x == 0) {
        float l = 0.f;
        for (int i = threadIdx.x; i < B; i += 256) l += kind <= 2 ? row_loss[i] : row_loss[i] * have_neg[i] / hn_sum * fB;
        __syncthreads();
        red[threadIdx.x] = l;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) loss_out[0] = red[0] / fB;
    }
    if (grad != nullptr) {
        const int64_t n = (int64_t)B * Cn;
        for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
            const int b = (int)(i / Cn);
            const float f = kind <= 2 ? 1.f / fB : have_neg[b] / hn_sum;       // f_b / B
            grad[i] *= f;
        }
    }
}

}  // namespace b2r

using namespace b2r;

extern "C" size_t b2r_listwise_workspace_bytes(int B) { return B > 0 ? align_up((size_t)B * 8, 256) : 0; }

extern "C" int b2r_listwise_loss(const float* pred, const int64_t* target, int B, int Cn, int max_pos, int kind, int hard,
                                 float* loss_out, float* grad_pred, void* ws, size_t ws_bytes, b2r_stream_t stream) {
    B2R_REQUIRE(pred && target && loss_out && ws, B2R_E_BADARG, "b2r_listwise_loss: null pointer");
    B2R_REQUIRE(B > 0 && Cn > 0 && max_pos > 0 && max_pos < Cn, B2R_E_BADARG,
                "b2r_listwise_loss: B=%d Cn=%d max_pos=%d (the reference reads column max_pos: needs max_pos < Cn)", B, Cn, max_pos);
    B2R_REQUIRE(kind >= 0 && kind <= 5, B2R_E_BADARG, "b2r_listwise_loss: kind %d", kind);
    B2R_REQUIRE(Cn <= 1024, B2R_E_UNSUPPORTED, "b2r_listwise_loss: Cn=%d > 1024", Cn);
    B2R_REQUIRE(ws_bytes >= b2r_listwise_workspace_bytes(B), B2R_E_WORKSPACE, "b2r_listwise_loss: workspace too small");
    cudaStream_t s = as_stream(stream);
    float* row_loss = static_cast<float*>(ws);
    float* have_neg = row_loss + B;
    const int smem = kLwWarps * 6 * Cn * 4;
    static int attr = 0;
    if (smem > attr && smem > 48 * 1024) {
        B2R_CUDA_OK(cudaFuncSetAttribute(k_listwise_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = smem;
    }
    k_listwise_rows<<<(B + kLwWarps - 1) / kLwWarps, kLwWarps * 32, smem, s>>>(pred, target, B, Cn, max_pos, kind, hard,
                                                                              row_loss, have_neg, grad_pred);
    B2R_LAUNCH_OK("k_listwise_rows");
    int64_t grid = grad_pred ? ((int64_t)B * Cn + 255) / 256 : 1;
    if (grid > sm_count() * 4) grid = sm_count() * 4;
    k_listwise_finish<<<(int)grid, 256, 0, s>>>(row_loss, have_neg, B, Cn, kind, loss_out, grad_pred);
    B2R_LAUNCH_OK("k_listwise_finish");
    return 0;
}
