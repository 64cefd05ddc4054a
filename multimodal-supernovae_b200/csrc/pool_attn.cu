// A6, agg="attn": nn.MultiheadAttention(emb, heads) with ONE learnable query per sequence over the zero-padded token
// tensor and NO key mask (src/transformer_utils.py:241-247).  kv is [B*T, 2E] = k|v (in_proj already applied, so padded
// rows carry k = b_k, v = b_v exactly like the reference); q is the projected query [E], identical for every sequence.
// One CTA per (sequence, head): scores over the T keys, block softmax, weighted sum of values.  Saves the probabilities.
#include "common.cuh"

namespace mvn {
namespace {

constexpr int QP_THREADS = 128;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < QP_THREADS / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
    return r;
}

template <int HD>
__global__ void __launch_bounds__(QP_THREADS) query_pool_fwd_kernel(const float* __restrict__ q, const float* __restrict__ kv, int T, int E, int H,
                                                                    float scale, float* __restrict__ out, float* __restrict__ probs) {
    __shared__ float red[QP_THREADS / 32];
    __shared__ float acc_s[QP_THREADS / 32][HD];
    const int b = blockIdx.x / H, h = blockIdx.x % H, tid = threadIdx.x;
    float qh[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) qh[d] = q[h * HD + d] * scale;
    const float* base = kv + (size_t)b * T * 2 * E + h * HD;
    float* pr = probs + ((size_t)b * H + h) * T;
    float mx = -INFINITY;
    for (int t = tid; t < T; t += QP_THREADS) {
        const float* kp = base + (size_t)t * 2 * E;
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
            const float4 k4 = *reinterpret_cast<const float4*>(kp + d);
            s = fmaf(qh[d], k4.x, fmaf(qh[d + 1], k4.y, fmaf(qh[d + 2], k4.z, fmaf(qh[d + 3], k4.w, s))));
        }
        pr[t] = s;
        mx = fmaxf(mx, s);
    }
    mx = block_reduce(mx, red, true);
    float sum = 0.f;
    for (int t = tid; t < T; t += QP_THREADS) { const float p = expf(pr[t] - mx); pr[t] = p; sum += p; }
    sum = block_reduce(sum, red, false);
    const float inv = 1.0f / sum;
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
    for (int t = tid; t < T; t += QP_THREADS) {
        const float p = pr[t] * inv;
        pr[t] = p;
        const float* vp = base + (size_t)t * 2 * E + E;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
            const float4 v4 = *reinterpret_cast<const float4*>(vp + d);
            o[d] = fmaf(p, v4.x, o[d]); o[d + 1] = fmaf(p, v4.y, o[d + 1]); o[d + 2] = fmaf(p, v4.z, o[d + 2]); o[d + 3] = fmaf(p, v4.w, o[d + 3]);
        }
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = warp_sum(o[d]);
    if ((tid & 31) == 0) {
#pragma unroll
        for (int d = 0; d < HD; ++d) acc_s[tid >> 5][d] = o[d];
    }
    __syncthreads();
    if (tid < HD) {
        float r = 0.f;
#pragma unroll
        for (int w = 0; w < QP_THREADS / 32; ++w) r += acc_s[w][tid];
        out[(size_t)b * E + h * HD + tid] = r;
    }
}

// dkv rows: dk_t = scale * ds_t * q_h, dv_t = p_t * do_h with ds_t = p_t (do_h . v_t - D), D = sum_t p_t (do_h . v_t);
// dq_part[b, h*HD + d] = scale * sum_t ds_t k_t[d]   (summed over b by rows_sum_kernel)
template <int HD>
__global__ void __launch_bounds__(QP_THREADS) query_pool_bwd_kernel(const float* __restrict__ q, const float* __restrict__ kv,
                                                                    const float* __restrict__ probs, const float* __restrict__ dout, int T, int E, int H,
                                                                    float scale, float* __restrict__ dkv, float* __restrict__ dq_part) {
    __shared__ float red[QP_THREADS / 32];
    __shared__ float acc_s[QP_THREADS / 32][HD];
    const int b = blockIdx.x / H, h = blockIdx.x % H, tid = threadIdx.x;
    float qh[HD], go[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) { qh[d] = q[h * HD + d]; go[d] = dout[(size_t)b * E + h * HD + d]; }
    const float* base = kv + (size_t)b * T * 2 * E + h * HD;
    float* dbase = dkv + (size_t)b * T * 2 * E + h * HD;
    const float* pr = probs + ((size_t)b * H + h) * T;
    float Dp = 0.f;
    for (int t = tid; t < T; t += QP_THREADS) {
        const float* vp = base + (size_t)t * 2 * E + E;
        float dp = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) dp = fmaf(go[d], vp[d], dp);
        Dp = fmaf(pr[t], dp, Dp);
    }
    Dp = block_reduce(Dp, red, false);
    float dq[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) dq[d] = 0.f;
    for (int t = tid; t < T; t += QP_THREADS) {
        const float* kp = base + (size_t)t * 2 * E;
        const float p = pr[t];
        float dp = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) dp = fmaf(go[d], kp[E + d], dp);
        const float ds = p * (dp - Dp) * scale;
        float* dk = dbase + (size_t)t * 2 * E;
#pragma unroll
        for (int d = 0; d < HD; ++d) { dk[d] = ds * qh[d]; dk[E + d] = p * go[d]; dq[d] = fmaf(ds, kp[d], dq[d]); }
    }
#pragma unroll
    for (int d = 0; d < HD; ++d) dq[d] = warp_sum(dq[d]);
    if ((tid & 31) == 0) {
#pragma unroll
        for (int d = 0; d < HD; ++d) acc_s[tid >> 5][d] = dq[d];
    }
    __syncthreads();
    if (tid < HD) {
        float r = 0.f;
#pragma unroll
        for (int w = 0; w < QP_THREADS / 32; ++w) r += acc_s[w][tid];
        dq_part[(size_t)b * E + h * HD + tid] = r;
    }
}

__global__ void rows_sum_kernel(const float* __restrict__ in, int B, int E, float* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += in[(size_t)b * E + e];
    out[e] = s;
}

}  // namespace
}  // namespace mvn

using namespace mvn;

extern "C" int mvn_query_pool_fwd(const float* q, const float* kv, int B, int T, int E, int H, float* out, float* probs, void* stream) {
    MVN_CHECK_ARG(q && kv && out && probs && B > 0 && T > 0 && E > 0 && H > 0 && E % H == 0, "query_pool_fwd: bad arguments");
    MVN_CHECK_ARG(aligned16(kv), "query_pool_fwd: kv must be 16-byte aligned");
    const int hd = E / H;
    const float scale = 1.0f / sqrtf((float)hd);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(PROF_ROW, st);
    switch (hd) {
        case 8: query_pool_fwd_kernel<8><<<B * H, QP_THREADS, 0, st>>>(q, kv, T, E, H, scale, out, probs); break;
        case 16: query_pool_fwd_kernel<16><<<B * H, QP_THREADS, 0, st>>>(q, kv, T, E, H, scale, out, probs); break;
        case 32: query_pool_fwd_kernel<32><<<B * H, QP_THREADS, 0, st>>>(q, kv, T, E, H, scale, out, probs); break;
        case 64: query_pool_fwd_kernel<64><<<B * H, QP_THREADS, 0, st>>>(q, kv, T, E, H, scale, out, probs); break;
        default: MVN_UNSUPPORTED(false, "query_pool: head dim %d not in {8,16,32,64}", hd);
    }
    MVN_LAUNCH_CHECK();
    return 0;
}

extern "C" int mvn_query_pool_bwd(const float* q, const float* kv, const float* probs, const float* dout, int B, int T, int E, int H,
                                  float* dkv, float* dq, void* workspace, size_t workspace_bytes, void* stream) {
    MVN_CHECK_ARG(q && kv && probs && dout && dkv && dq && workspace && B > 0 && T > 0 && E > 0 && H > 0 && E % H == 0, "query_pool_bwd: bad arguments");
    if (workspace_bytes < (size_t)B * E * sizeof(float)) { set_error("query_pool_bwd: workspace %zu < %zu", workspace_bytes, (size_t)B * E * 4); return MVN_E_WORKSPACE; }
    const int hd = E / H;
    const float scale = 1.0f / sqrtf((float)hd);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(PROF_ROW, st);
    float* part = (float*)workspace;
    switch (hd) {
        case 8: query_pool_bwd_kernel<8><<<B * H, QP_THREADS, 0, st>>>(q, kv, probs, dout, T, E, H, scale, dkv, part); break;
        case 16: query_pool_bwd_kernel<16><<<B * H, QP_THREADS, 0, st>>>(q, kv, probs, dout, T, E, H, scale, dkv, part); break;
        case 32: query_pool_bwd_kernel<32><<<B * H, QP_THREADS, 0, st>>>(q, kv, probs, dout, T, E, H, scale, dkv, part); break;
        case 64: query_pool_bwd_kernel<64><<<B * H, QP_THREADS, 0, st>>>(q, kv, probs, dout, T, E, H, scale, dkv, part); break;
        default: MVN_UNSUPPORTED(false, "query_pool: head dim %d not in {8,16,32,64}", hd);
    }
    MVN_LAUNCH_CHECK();
    rows_sum_kernel<<<cdiv(E, 128), 128, 0, st>>>(part, B, E, dq);
    MVN_LAUNCH_CHECK();
    return 0;
}
