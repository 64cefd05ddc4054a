// A3: padding-masked multi-head self-attention over the packed token stream, fp32 tier.
// One CTA per (sequence, head); one thread per query row; keys/values streamed through shared memory in chunks
// with an online softmax, so the T x T scores never exist in memory.  Backward recomputes probabilities from the
// saved log-sum-exp (two passes: dQ by query-thread, dK/dV by key-thread; no atomics, deterministic).
#include "common.cuh"

namespace mvn {
namespace {

constexpr int ATT_THREADS = 128;
constexpr int CHUNK = 128;
constexpr float MASK_FILL = -1e7f;     // src/transformer_utils.py:77

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ cu,
                                                               const uint8_t* __restrict__ keyvalid, float* __restrict__ out,
                                                               float* __restrict__ lse, int E, int H, float scale) {
    __shared__ __align__(16) float Ks[CHUNK][HD];
    __shared__ __align__(16) float Vs[CHUNK][HD];
    __shared__ uint8_t ok[CHUNK];
    __shared__ int nvalid_s;
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int r0 = cu[b], n = cu[b + 1] - r0;
    if (n <= 0) return;
    const int tid = threadIdx.x;
    const size_t ld = 3 * (size_t)E;
    if (tid == 0) nvalid_s = 0;
    __syncthreads();
    {
        int c = 0;
        for (int j = tid; j < n; j += ATT_THREADS) c += keyvalid ? (keyvalid[r0 + j] != 0) : 1;
        if (c) atomicAdd(&nvalid_s, c);
    }
    __syncthreads();
    const bool uniform = nvalid_s == 0;      // every key masked: reference softmax is uniform over all T keys

    for (int q0 = 0; q0 < n; q0 += ATT_THREADS) {
        const int qi = q0 + tid;
        const bool has_q = qi < n;
        float q[HD], o[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) { q[d] = 0.f; o[d] = 0.f; }
        if (has_q) {
            const float* qp = qkv + (size_t)(r0 + qi) * ld + h * HD;
#pragma unroll
            for (int d = 0; d < HD; d += 4) { const float4 v = *reinterpret_cast<const float4*>(qp + d); q[d] = v.x; q[d + 1] = v.y; q[d + 2] = v.z; q[d + 3] = v.w; }
        }
        float mrun = -INFINITY, lrun = 0.f;
        for (int k0 = 0; k0 < n; k0 += CHUNK) {
            const int kn = min(CHUNK, n - k0);
            __syncthreads();
            for (int idx = tid; idx < kn * (HD / 4); idx += ATT_THREADS) {
                const int j = idx / (HD / 4), d = (idx % (HD / 4)) * 4;
                const float* base = qkv + (size_t)(r0 + k0 + j) * ld + h * HD + d;
                *reinterpret_cast<float4*>(&Ks[j][d]) = *reinterpret_cast<const float4*>(base + E);
                *reinterpret_cast<float4*>(&Vs[j][d]) = *reinterpret_cast<const float4*>(base + 2 * E);
            }
            for (int j = tid; j < kn; j += ATT_THREADS) ok[j] = uniform ? 1 : (keyvalid ? keyvalid[r0 + k0 + j] : 1);
            __syncthreads();
            if (has_q) {
                for (int j0 = 0; j0 < kn; j0 += 8) {
                    float s[8];
                    float gmax = -INFINITY;
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int j = j0 + jj;
                        float acc = -INFINITY;
                        if (j < kn && ok[j]) {
                            acc = 0.f;
#pragma unroll
                            for (int d = 0; d < HD; ++d) acc = fmaf(q[d], Ks[j][d], acc);
                            acc = uniform ? MASK_FILL : acc * scale;
                        }
                        s[jj] = acc;
                        gmax = fmaxf(gmax, acc);
                    }
                    if (gmax == -INFINITY) continue;
                    const float mnew = fmaxf(mrun, gmax);
                    const float corr = expf(mrun - mnew);        // exp(-inf)=0 on the first group
                    lrun *= corr;
#pragma unroll
                    for (int d = 0; d < HD; ++d) o[d] *= corr;
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int j = j0 + jj;
                        if (s[jj] == -INFINITY) continue;
                        const float p = expf(s[jj] - mnew);
                        lrun += p;
#pragma unroll
                        for (int d = 0; d < HD; ++d) o[d] = fmaf(p, Vs[j][d], o[d]);
                    }
                    mrun = mnew;
                }
            }
        }
        if (has_q) {
            const float inv = 1.0f / lrun;
            float* op = out + (size_t)(r0 + qi) * E + h * HD;
#pragma unroll
            for (int d = 0; d < HD; d += 4) *reinterpret_cast<float4*>(op + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
            lse[(size_t)(r0 + qi) * H + h] = mrun + logf(lrun);
        }
    }
}

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ cu,
                                                               const uint8_t* __restrict__ keyvalid, const float* __restrict__ out,
                                                               const float* __restrict__ lse, const float* __restrict__ dout,
                                                               float* __restrict__ dqkv, int E, int H, float scale) {
    // shared staging reused by both passes: A = keys (pass 1) / queries (pass 2), Bv = values / dO
    __shared__ __align__(16) float As[CHUNK][HD];
    __shared__ __align__(16) float Bs[CHUNK][HD];
    __shared__ float aux0[CHUNK];      // pass 2: lse_i
    __shared__ float aux1[CHUNK];      // pass 2: D_i
    __shared__ uint8_t ok[CHUNK];
    __shared__ int nvalid_s;
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int r0 = cu[b], n = cu[b + 1] - r0;
    if (n <= 0) return;
    const int tid = threadIdx.x;
    const size_t ld = 3 * (size_t)E;
    if (tid == 0) nvalid_s = 0;
    __syncthreads();
    {
        int c = 0;
        for (int j = tid; j < n; j += ATT_THREADS) c += keyvalid ? (keyvalid[r0 + j] != 0) : 1;
        if (c) atomicAdd(&nvalid_s, c);
    }
    __syncthreads();
    const bool uniform = nvalid_s == 0;
    const float inv_n = 1.0f / (float)n;

    // ---- pass 1: dQ_i = scale * sum_j P_ij (dO_i.V_j - D_i) K_j ------------------------------------
    for (int q0 = 0; q0 < n; q0 += ATT_THREADS) {
        const int qi = q0 + tid;
        const bool has_q = qi < n;
        float q[HD], go[HD], dq[HD];
        float Di = 0.f, lse_i = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) { q[d] = 0.f; go[d] = 0.f; dq[d] = 0.f; }
        if (has_q) {
            const float* qp = qkv + (size_t)(r0 + qi) * ld + h * HD;
            const float* gp = dout + (size_t)(r0 + qi) * E + h * HD;
            const float* op = out + (size_t)(r0 + qi) * E + h * HD;
#pragma unroll
            for (int d = 0; d < HD; ++d) { q[d] = qp[d]; go[d] = gp[d]; Di = fmaf(gp[d], op[d], Di); }
            lse_i = lse[(size_t)(r0 + qi) * H + h];
        }
        if (!uniform) {
            for (int k0 = 0; k0 < n; k0 += CHUNK) {
                const int kn = min(CHUNK, n - k0);
                __syncthreads();
                for (int idx = tid; idx < kn * (HD / 4); idx += ATT_THREADS) {
                    const int j = idx / (HD / 4), d = (idx % (HD / 4)) * 4;
                    const float* base = qkv + (size_t)(r0 + k0 + j) * ld + h * HD + d;
                    *reinterpret_cast<float4*>(&As[j][d]) = *reinterpret_cast<const float4*>(base + E);
                    *reinterpret_cast<float4*>(&Bs[j][d]) = *reinterpret_cast<const float4*>(base + 2 * E);
                }
                for (int j = tid; j < kn; j += ATT_THREADS) ok[j] = keyvalid ? keyvalid[r0 + k0 + j] : 1;
                __syncthreads();
                if (has_q) {
                    for (int j = 0; j < kn; ++j) {
                        if (!ok[j]) continue;
                        float sc = 0.f, dp = 0.f;
#pragma unroll
                        for (int d = 0; d < HD; ++d) { sc = fmaf(q[d], As[j][d], sc); dp = fmaf(go[d], Bs[j][d], dp); }
                        const float p = expf(sc * scale - lse_i);
                        const float ds = p * (dp - Di);
#pragma unroll
                        for (int d = 0; d < HD; ++d) dq[d] = fmaf(ds, As[j][d], dq[d]);
                    }
                }
            }
        }
        if (has_q) {
            float* dqp = dqkv + (size_t)(r0 + qi) * ld + h * HD;
#pragma unroll
            for (int d = 0; d < HD; ++d) dqp[d] = dq[d] * scale;      // zero when every key is masked (constant scores)
        }
    }

    // ---- pass 2: dV_j = sum_i P_ij dO_i ; dK_j = scale * sum_i P_ij (dO_i.V_j - D_i) Q_i ------------
    for (int k0 = 0; k0 < n; k0 += ATT_THREADS) {
        const int kj = k0 + tid;
        const bool has_k = kj < n;
        const bool kvalid = has_k && (uniform || !keyvalid || keyvalid[r0 + kj]);
        float k[HD], v[HD], dk[HD], dv[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) { k[d] = 0.f; v[d] = 0.f; dk[d] = 0.f; dv[d] = 0.f; }
        if (has_k) {
            const float* kp = qkv + (size_t)(r0 + kj) * ld + E + h * HD;
#pragma unroll
            for (int d = 0; d < HD; ++d) { k[d] = kp[d]; v[d] = kp[E + d]; }
        }
        for (int q0 = 0; q0 < n; q0 += CHUNK) {
            const int qn = min(CHUNK, n - q0);
            __syncthreads();
            for (int idx = tid; idx < qn * (HD / 4); idx += ATT_THREADS) {
                const int i = idx / (HD / 4), d = (idx % (HD / 4)) * 4;
                *reinterpret_cast<float4*>(&As[i][d]) = *reinterpret_cast<const float4*>(qkv + (size_t)(r0 + q0 + i) * ld + h * HD + d);
                *reinterpret_cast<float4*>(&Bs[i][d]) = *reinterpret_cast<const float4*>(dout + (size_t)(r0 + q0 + i) * E + h * HD + d);
            }
            for (int i = tid; i < qn; i += ATT_THREADS) {
                const float* gp = dout + (size_t)(r0 + q0 + i) * E + h * HD;
                const float* op = out + (size_t)(r0 + q0 + i) * E + h * HD;
                float Di = 0.f;
#pragma unroll
                for (int d = 0; d < HD; ++d) Di = fmaf(gp[d], op[d], Di);
                aux0[i] = lse[(size_t)(r0 + q0 + i) * H + h];
                aux1[i] = Di;
            }
            __syncthreads();
            if (kvalid) {
                for (int i = 0; i < qn; ++i) {
                    float p, ds = 0.f;
                    if (uniform) {
                        p = inv_n;
                    } else {
                        float sc = 0.f, dp = 0.f;
#pragma unroll
                        for (int d = 0; d < HD; ++d) { sc = fmaf(As[i][d], k[d], sc); dp = fmaf(Bs[i][d], v[d], dp); }
                        p = expf(sc * scale - aux0[i]);
                        ds = p * (dp - aux1[i]);
                    }
#pragma unroll
                    for (int d = 0; d < HD; ++d) { dv[d] = fmaf(p, Bs[i][d], dv[d]); dk[d] = fmaf(ds, As[i][d], dk[d]); }
                }
            }
        }
        if (has_k) {
            float* dkp = dqkv + (size_t)(r0 + kj) * ld + E + h * HD;
#pragma unroll
            for (int d = 0; d < HD; ++d) { dkp[d] = dk[d] * scale; dkp[E + d] = dv[d]; }
        }
    }
}

}  // namespace
}  // namespace mvn

namespace mvn {
int launch_attention_fwd_tc(const float* qkv, const int32_t* cu, float* out, float* lse, int B, int E, int H, float scale, cudaStream_t st);
int launch_attention_bwd_tc(const float* qkv, const int32_t* cu, const float* out, const float* lse, const float* dout, float* dqkv,
                            int B, int E, int H, float scale, cudaStream_t st);
}  // namespace mvn

using namespace mvn;

extern "C" int mvn_attention_fwd(const float* qkv, const int32_t* cu_seqlens, const uint8_t* keyvalid, float* out, float* lse,
                                 int B, int E, int H, float scale, int prec, void* stream) {
    MVN_CHECK_ARG(qkv && cu_seqlens && out && lse && B > 0 && E > 0 && H > 0 && E % H == 0, "attention_fwd: bad arguments (E=%d H=%d)", E, H);
    MVN_CHECK_ARG(aligned16(qkv) && aligned16(out), "attention_fwd: buffers must be 16-byte aligned");
    const int hd = E / H;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(PROF_ATTN_FWD, st);
    if (prec >= 1 && keyvalid == nullptr) {       // packed stream, every key live: tensor-core kernel
        const int r = launch_attention_fwd_tc(qkv, cu_seqlens, out, lse, B, E, H, scale, st);
        if (r != MVN_E_UNSUPPORTED) { if (r == 0) count_tier(TIER_MMA); return r; }
    }
    count_tier(TIER_FFMA);
    const int grid = B * H;
    switch (hd) {
        case 4: attn_fwd_kernel<4><<<grid, ATT_THREADS, 0, st>>>(qkv, cu_seqlens, keyvalid, out, lse, E, H, scale); break;
        case 8: attn_fwd_kernel<8><<<grid, ATT_THREADS, 0, st>>>(qkv, cu_seqlens, keyvalid, out, lse, E, H, scale); break;
        case 16: attn_fwd_kernel<16><<<grid, ATT_THREADS, 0, st>>>(qkv, cu_seqlens, keyvalid, out, lse, E, H, scale); break;
        case 32: attn_fwd_kernel<32><<<grid, ATT_THREADS, 0, st>>>(qkv, cu_seqlens, keyvalid, out, lse, E, H, scale); break;
        default: MVN_UNSUPPORTED(false, "attention: head dim %d not in {4,8,16,32}", hd);
    }
    MVN_LAUNCH_CHECK();
    return 0;
}

extern "C" int mvn_attention_bwd(const float* qkv, const int32_t* cu_seqlens, const uint8_t* keyvalid, const float* out,
                                 const float* lse, const float* dout, float* dqkv, int B, int E, int H, float scale, int prec,
                                 void* stream) {
    MVN_CHECK_ARG(qkv && cu_seqlens && out && lse && dout && dqkv && B > 0 && E > 0 && H > 0 && E % H == 0, "attention_bwd: bad arguments");
    MVN_CHECK_ARG(aligned16(qkv) && aligned16(dout) && aligned16(dqkv), "attention_bwd: buffers must be 16-byte aligned");
    const int hd = E / H;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(PROF_ATTN_BWD, st);
    if (prec >= 1 && keyvalid == nullptr) {
        const int r = launch_attention_bwd_tc(qkv, cu_seqlens, out, lse, dout, dqkv, B, E, H, scale, st);
        if (r != MVN_E_UNSUPPORTED) { if (r == 0) count_tier(TIER_MMA); return r; }
    }
    count_tier(TIER_FFMA);
    const int grid = B * H;
    switch (hd) {
        case 4: attn_bwd_kernel<4><<<grid, ATT_THREADS, 0, st>>>(qkv, cu_seqlens, keyvalid, out, lse, dout, dqkv, E, H, scale); break;
        case 8: attn_bwd_kernel<8><<<grid, ATT_THREADS, 0, st>>>(qkv, cu_seqlens, keyvalid, out, lse, dout, dqkv, E, H, scale); break;
        case 16: attn_bwd_kernel<16><<<grid, ATT_THREADS, 0, st>>>(qkv, cu_seqlens, keyvalid, out, lse, dout, dqkv, E, H, scale); break;
        case 32: attn_bwd_kernel<32><<<grid, ATT_THREADS, 0, st>>>(qkv, cu_seqlens, keyvalid, out, lse, dout, dqkv, E, H, scale); break;
        default: MVN_UNSUPPORTED(false, "attention: head dim %d not in {4,8,16,32}", hd);
    }
    MVN_LAUNCH_CHECK();
    return 0;
}
