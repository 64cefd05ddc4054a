// A13 fused RAdam over a flat buffer, A12 loss heads (weighted cross-entropy, MSE), N3 retrieval ranks.
#include "common.cuh"

namespace mvn {
namespace {

__global__ void __launch_bounds__(256) radam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                    long long n, float lr, float beta1, float beta2, float eps, float wd, float bc1,
                                                    float sqrt_bc2, float rect) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float pi = p[i];
        float gi = g[i];
        if (wd != 0.f) gi = fmaf(wd, pi, gi);                       // coupled L2 (decoupled_weight_decay=False)
        const float mi = m[i] + (gi - m[i]) * (1.0f - beta1);       // lerp_
        const float vi = fmaf(v[i], beta2, gi * gi * (1.0f - beta2));
        m[i] = mi; v[i] = vi;
        const float mhat = mi / bc1;
        float step = mhat * lr;
        if (rect >= 0.f) step *= (sqrt_bc2 / (sqrtf(vi) + eps)) * rect;
        p[i] = pi - step;
    }
}

// Device-side step bookkeeping for CUDA-graph replays (kernel arguments are frozen in a graph, the step count is not):
// ++*step, then the step-dependent RAdam scalars in double precision -> scal = {bc1, sqrt(bc2), rect (<0: not rectified)}.
__global__ void radam_tick_kernel(uint32_t* __restrict__ step, float* __restrict__ scal, double beta1, double beta2) {
    const uint32_t st = *step + 1u;
    *step = st;
    const double b1t = pow(beta1, (double)st), b2t = pow(beta2, (double)st);
    const double bc1 = 1.0 - b1t, bc2 = 1.0 - b2t;
    const double rho_inf = 2.0 / (1.0 - beta2) - 1.0;
    const double rho_t = rho_inf - 2.0 * (double)st * b2t / bc2;
    double rect = -1.0;
    if (rho_t > 5.0) rect = sqrt((rho_t - 4.0) * (rho_t - 2.0) * rho_inf / ((rho_inf - 4.0) * (rho_inf - 2.0) * rho_t));
    scal[0] = (float)bc1; scal[1] = (float)sqrt(bc2); scal[2] = (float)rect;
}
__global__ void __launch_bounds__(256) radam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                        long long n, float lr, float beta1, float beta2, float eps, float wd,
                                                        const float* __restrict__ scal) {
    const float bc1 = scal[0], sqrt_bc2 = scal[1], rect = scal[2];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float pi = p[i];
        float gi = g[i];
        if (wd != 0.f) gi = fmaf(wd, pi, gi);
        const float mi = m[i] + (gi - m[i]) * (1.0f - beta1);
        const float vi = fmaf(v[i], beta2, gi * gi * (1.0f - beta2));
        m[i] = mi; v[i] = vi;
        const float mhat = mi / bc1;
        float step = mhat * lr;
        if (rect >= 0.f) step *= (sqrt_bc2 / (sqrtf(vi) + eps)) * rect;
        p[i] = pi - step;
    }
}

// one warp per sample: numerically stable log-softmax at the label
__global__ void ce_rows_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ cw, int B, int C,
                               float* __restrict__ wnll, float* __restrict__ wsum) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= B) return;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, logits[(size_t)row * C + c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(logits[(size_t)row * C + c] - mx);
    s = warp_sum(s);
    if (lane == 0) {
        const int y = (int)labels[row];
        const float w = cw ? cw[y] : 1.0f;
        wnll[row] = w * (mx + logf(s) - logits[(size_t)row * C + y]);
        wsum[row] = w;
    }
}
__global__ void __launch_bounds__(1024) ce_finish_kernel(const float* __restrict__ wnll, const float* __restrict__ wsum, int B, float* __restrict__ loss) {
    __shared__ float r0[32], r1[32];
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < B; i += 1024) { a += wnll[i]; b += wsum[i]; }
    a = warp_sum(a); b = warp_sum(b);
    if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = a; r1[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x < 32) {
        a = warp_sum(r0[threadIdx.x]); b = warp_sum(r1[threadIdx.x]);
        if (threadIdx.x == 0) { loss[0] = a / b; loss[1] = b; }
    }
}
__global__ void ce_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ cw, int B, int C,
                              const float* __restrict__ wtotal, const float* __restrict__ grad_out, float* __restrict__ dlogits) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= B) return;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, logits[(size_t)row * C + c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(logits[(size_t)row * C + c] - mx);
    s = warp_sum(s);
    const int y = (int)labels[row];
    const float f = (cw ? cw[y] : 1.0f) / wtotal[0] * (grad_out ? *grad_out : 1.0f);
    for (int c = lane; c < C; c += 32) {
        const float pr = expf(logits[(size_t)row * C + c] - mx) / s;
        dlogits[(size_t)row * C + c] = f * (pr - (c == y ? 1.0f : 0.0f));
    }
}

__global__ void __launch_bounds__(1024) mse_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, float* __restrict__ loss) {
    __shared__ float red[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += 1024) { const float d = a[i] - b[i]; s = fmaf(d, d, s); }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) { s = warp_sum(red[threadIdx.x]); if (threadIdx.x == 0) loss[0] = s / (float)n; }
}
__global__ void mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, const float* __restrict__ grad_out, float* __restrict__ da) {
    const float f = 2.0f / (float)n * (grad_out ? *grad_out : 1.0f);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) da[i] = f * (a[i] - b[i]);
}

// masked-light-curve objective (src/models_pretraining.py:206-226): nn.MSELoss over the positions selected by a bool mask,
// loss = sum_i m_i (a_i - b_i)^2 / sum_i m_i  (an empty selection gives 0/0 = NaN like the mean of an empty tensor).
// buf[0] = loss, buf[1] = number of selected positions.
__global__ void __launch_bounds__(1024) masked_mse_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const unsigned char* __restrict__ m,
                                                              int n, float* __restrict__ buf) {
    __shared__ float red[32];
    __shared__ int redc[32];
    float s = 0.f;
    int cnt = 0;
    for (int i = threadIdx.x; i < n; i += 1024)
        if (m[i]) { const float d = a[i] - b[i]; s = fmaf(d, d, s); ++cnt; }
    s = warp_sum(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = s; redc[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x < 32) {
        s = warp_sum(red[threadIdx.x]);
        cnt = redc[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (threadIdx.x == 0) { buf[0] = s / (float)cnt; buf[1] = (float)cnt; }
    }
}
__global__ void masked_mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const unsigned char* __restrict__ m, int n,
                                      const float* __restrict__ buf, const float* __restrict__ grad_out, float* __restrict__ da) {
    const float f = 2.0f / buf[1] * (grad_out ? *grad_out : 1.0f);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) da[i] = m[i] ? f * (a[i] - b[i]) : 0.f;
}

// retrieval rank: one CTA per source j (rows of e2); counts i with cos(e1_i,e2_j) > cos(e1_j,e2_j)
__global__ void __launch_bounds__(256) ranks_kernel(const float* __restrict__ e1, const float* __restrict__ e2, int N, int D, int32_t* __restrict__ ranks) {
    extern __shared__ float src[];          // e2_j normalised
    __shared__ float red[8];
    __shared__ int cnt;
    const int j = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float s = 0.f;
    for (int d = tid; d < D; d += 256) { const float v = e2[(size_t)j * D + d]; src[d] = v; s = fmaf(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) red[wid] = s;
    if (tid == 0) cnt = 0;
    __syncthreads();
    float n2 = 0.f;
    for (int w = 0; w < 8; ++w) n2 += red[w];
    const float inv2 = 1.0f / fmaxf(sqrtf(n2), 1e-12f);
    auto cosine = [&](int i) {
        float dot = 0.f, nn = 0.f;
        for (int d = lane; d < D; d += 32) { const float v = e1[(size_t)i * D + d]; dot = fmaf(v, src[d], dot); nn = fmaf(v, v, nn); }
        dot = warp_sum(dot); nn = warp_sum(nn);
        return dot * inv2 / fmaxf(sqrtf(nn), 1e-12f);
    };
    const float self = cosine(j);
    int c = 0;
    for (int i = wid; i < N; i += 8) c += cosine(i) > self;
    if (lane == 0 && c) atomicAdd(&cnt, c);
    __syncthreads();
    if (tid == 0) ranks[j] = cnt;
}

}  // namespace
}  // namespace mvn

using namespace mvn;

extern "C" int mvn_radam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, float bias_correction1, float sqrt_bias_correction2, float rect,
                              void* stream) {
    MVN_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n > 0, "radam_step: bad arguments");
    ProfScope prof(PROF_OPTIM, (cudaStream_t)stream);
    const long long blocks = (n + 255) / 256;
    radam_kernel<<<(int)(blocks < 2368 ? blocks : 2368), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                                       weight_decay, bias_correction1, sqrt_bias_correction2, rect);
    MVN_LAUNCH_CHECK();
    return 0;
}

extern "C" int mvn_radam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                                  float beta2, float eps, float weight_decay, uint32_t* step_dev, float* scalars_dev, void* stream) {
    MVN_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n > 0 && scalars_dev, "radam_step_dev: bad arguments");
    ProfScope prof(PROF_OPTIM, (cudaStream_t)stream);
    if (step_dev) {     // first segment of a step: advance the counter and refresh the scalars; later segments pass NULL and reuse them
        radam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev, scalars_dev, (double)beta1, (double)beta2);
        MVN_LAUNCH_CHECK();
    }
    const long long blocks = (n + 255) / 256;
    radam_dev_kernel<<<(int)(blocks < 2368 ? blocks : 2368), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                                           weight_decay, scalars_dev);
    MVN_LAUNCH_CHECK();
    return 0;
}

// loss buffer layout: [0] weighted mean NLL, [1] total weight (read by the backward), [2 .. 2+2B) scratch
extern "C" int mvn_weighted_ce_fwd(const float* logits, const int64_t* labels, const float* class_w, int B, int C, float* loss, void* stream) {
    MVN_CHECK_ARG(logits && labels && loss && B > 0 && C > 0, "weighted_ce_fwd: bad arguments");
    // scratch: loss[2 .. 2+2B)
    float* wnll = loss + 2; float* wsum = wnll + B;
    cudaStream_t st = (cudaStream_t)stream;
    ce_rows_kernel<<<cdiv(B * 32, 256), 256, 0, st>>>(logits, labels, class_w, B, C, wnll, wsum);
    MVN_LAUNCH_CHECK();
    ce_finish_kernel<<<1, 1024, 0, st>>>(wnll, wsum, B, loss);
    MVN_LAUNCH_CHECK();
    return 0;
}
extern "C" int mvn_weighted_ce_bwd(const float* logits, const int64_t* labels, const float* class_w, int B, int C, const float* loss_buf,
                                    const float* grad_out, float* dlogits, void* stream) {
    MVN_CHECK_ARG(logits && labels && dlogits && loss_buf && B > 0 && C > 0, "weighted_ce_bwd: bad arguments");
    ce_bwd_kernel<<<cdiv(B * 32, 256), 256, 0, (cudaStream_t)stream>>>(logits, labels, class_w, B, C, loss_buf + 1, grad_out, dlogits);
    MVN_LAUNCH_CHECK();
    return 0;
}
extern "C" int mvn_mse_fwd(const float* pred, const float* target, int n, float* loss, void* stream) {
    MVN_CHECK_ARG(pred && target && loss && n > 0, "mse_fwd: bad arguments");
    mse_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pred, target, n, loss);
    MVN_LAUNCH_CHECK();
    return 0;
}
extern "C" int mvn_masked_mse_fwd(const float* pred, const float* target, const unsigned char* mask, int n, float* loss_buf, void* stream) {
    MVN_CHECK_ARG(pred && target && mask && loss_buf && n > 0, "masked_mse_fwd: bad arguments");
    masked_mse_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pred, target, mask, n, loss_buf);
    MVN_LAUNCH_CHECK();
    return 0;
}
extern "C" int mvn_masked_mse_bwd(const float* pred, const float* target, const unsigned char* mask, int n, const float* loss_buf,
                                  const float* grad_out, float* dpred, void* stream) {
    MVN_CHECK_ARG(pred && target && mask && loss_buf && dpred && n > 0, "masked_mse_bwd: bad arguments");
    masked_mse_bwd_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(pred, target, mask, n, loss_buf, grad_out, dpred);
    MVN_LAUNCH_CHECK();
    return 0;
}
extern "C" int mvn_mse_bwd(const float* pred, const float* target, int n, const float* grad_out, float* dpred, void* stream) {
    MVN_CHECK_ARG(pred && target && dpred && n > 0, "mse_bwd: bad arguments");
    mse_bwd_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(pred, target, n, grad_out, dpred);
    MVN_LAUNCH_CHECK();
    return 0;
}
// meta modality input (src/models_multimodal.py:295-304): row b = [ class_emb[cls_b] | redshift_b repeated `half` times ]
__global__ void meta_input_fwd_kernel(const float* __restrict__ emb, const int64_t* __restrict__ cls, const float* __restrict__ red, int B, int half,
                                      int n_classes, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 2 * half) return;
    const int b = i / (2 * half), d = i % (2 * half);
    long long c = cls[b];
    c = c < 0 ? 0 : (c >= n_classes ? n_classes - 1 : c);
    out[i] = d < half ? emb[c * half + d] : red[b];
}
// d class_emb[c, d] = sum over the batch rows with cls == c (fixed order: deterministic)
__global__ void meta_input_bwd_kernel(const float* __restrict__ dout, const int64_t* __restrict__ cls, int B, int half, int n_classes,
                                      float* __restrict__ demb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_classes * half) return;
    const int c = i / half, d = i % half;
    float s = 0.f;
    for (int b = 0; b < B; ++b)
        if (cls[b] == c) s += dout[(size_t)b * 2 * half + d];
    demb[i] = s;
}
extern "C" int mvn_meta_input_fwd(const float* class_emb, const int64_t* cls, const float* redshift, int B, int half, int n_classes, float* out, void* stream) {
    MVN_CHECK_ARG(class_emb && cls && redshift && out && B > 0 && half > 0 && n_classes > 0, "meta_input_fwd: bad arguments");
    meta_input_fwd_kernel<<<cdiv(B * 2 * half, 256), 256, 0, (cudaStream_t)stream>>>(class_emb, cls, redshift, B, half, n_classes, out);
    MVN_LAUNCH_CHECK();
    return 0;
}
extern "C" int mvn_meta_input_bwd(const float* dout, const int64_t* cls, int B, int half, int n_classes, float* dclass_emb, void* stream) {
    MVN_CHECK_ARG(dout && cls && dclass_emb && B > 0 && half > 0 && n_classes > 0, "meta_input_bwd: bad arguments");
    meta_input_bwd_kernel<<<cdiv(n_classes * half, 128), 128, 0, (cudaStream_t)stream>>>(dout, cls, B, half, n_classes, dclass_emb);
    MVN_LAUNCH_CHECK();
    return 0;
}

// retrieval curve: counts[t] = #{j : rank_j < k_thr[t]}  (get_ROC_data's "idx in idx_sorted[:int(threshold * N)]" summed over sources)
__global__ void __launch_bounds__(256) rank_curve_kernel(const int32_t* __restrict__ ranks, int N, const int32_t* __restrict__ k_thr, int n_thr,
                                                         int32_t* __restrict__ counts) {
    extern __shared__ int32_t sk[];          // thresholds | per-CTA counts
    int32_t* sc = sk + n_thr;
    for (int i = threadIdx.x; i < n_thr; i += blockDim.x) { sk[i] = k_thr[i]; sc[i] = 0; }
    __syncthreads();
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
        const int r = ranks[j];
        for (int t = 0; t < n_thr; ++t)
            if (r < sk[t]) atomicAdd(&sc[t], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_thr; i += blockDim.x)
        if (sc[i]) atomicAdd(&counts[i], sc[i]);      // integer atomics: order-independent, deterministic
}
extern "C" int mvn_retrieval_curve(const int32_t* ranks, int N, const int32_t* k_thr, int n_thr, int32_t* counts, void* stream) {
    MVN_CHECK_ARG(ranks && k_thr && counts && N > 0 && n_thr > 0 && n_thr <= 4096, "retrieval_curve: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    MVN_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_thr * sizeof(int32_t), st));
    const int grid = cdiv(N, 256) < 4 * num_sms() ? cdiv(N, 256) : 4 * num_sms();
    rank_curve_kernel<<<grid, 256, 2 * (size_t)n_thr * sizeof(int32_t), st>>>(ranks, N, k_thr, n_thr, counts);
    MVN_LAUNCH_CHECK();
    return 0;
}
extern "C" int mvn_retrieval_ranks(const float* e1, const float* e2, int N, int D, int32_t* ranks, void* stream) {
    MVN_CHECK_ARG(e1 && e2 && ranks && N > 0 && D > 0 && D <= 8192, "retrieval_ranks: bad arguments");
    ranks_kernel<<<N, 256, D * sizeof(float), (cudaStream_t)stream>>>(e1, e2, N, D, ranks);
    MVN_LAUNCH_CHECK();
    return 0;
}
