// Shared helpers for libmaven_sm100.so (sm_100a only).
#pragma once
#include <stdlib.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/maven_sm100.h"

namespace mvn {

void set_error(const char* fmt, ...);

#define MVN_CHECK_ARG(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            mvn::set_error(__VA_ARGS__);         \
            return MVN_E_BADARG;                 \
        }                                        \
    } while (0)

#define MVN_UNSUPPORTED(cond, ...)               \
    do {                                         \
        if (!(cond)) {                           \
            mvn::set_error(__VA_ARGS__);         \
            return MVN_E_UNSUPPORTED;            \
        }                                        \
    } while (0)

#define MVN_CUDA(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            mvn::set_error("%s -> %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (int)_e;                                                         \
        }                                                                           \
    } while (0)

#define MVN_LAUNCH_CHECK()                                                          \
    do {                                                                            \
        mvn::note_launch();                                                         \
        cudaError_t _e = cudaGetLastError();                                        \
        if (_e != cudaSuccess) {                                                    \
            mvn::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (int)_e;                                                         \
        }                                                                           \
    } while (0)

#define MVN_TRY(expr)            \
    do {                         \
        int _r = (expr);         \
        if (_r != 0) return _r;  \
    } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int num_sms();
void note_launch();

// per-kernel-class device timing (bench.py's roofline leg): CUDA events on the launching stream
enum ProfClass { PROF_GEMM = 0, PROF_WGRAD = 1, PROF_ATTN_FWD = 2, PROF_ATTN_BWD = 3, PROF_ROW = 4, PROF_LOSS = 5, PROF_OPTIM = 6, PROF_CONV = 7, PROF_FUSED_FWD = 8, PROF_FUSED_BWD = 9, PROF_NCLASS = 10 };
// which arithmetic tier a GEMM-class launch actually ran on (mvn_tier_count): tests assert that tensor-core shapes did not fall back
enum Tier { TIER_FFMA = 0, TIER_TC = 1, TIER_MMA = 2, TIER_FUSED = 3, TIER_N = 4 };
void count_tier(int tier);
struct ProfScope {
    int cls; cudaStream_t st; cudaEvent_t stop; bool on;
    ProfScope(int cls_, cudaStream_t st_);
    ~ProfScope();
};

// number of row-slabs every weight-gradient style reduction is split into (partials are [kSlabs][...])
constexpr int kSlabs = 148;      // one slab per B200 SM: every slab-partitioned kernel (weight gradients, LayerNorm backward, statistics) fills the GPU

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int W>
__device__ __forceinline__ float group_sum(float v) {   // sum over aligned groups of W lanes
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int W>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------------
// Every heavy kernel calls pdl_trigger() first thing: once all of its CTAs have started, a dependent kernel launched
// with the programmatic-serialization attribute may take over SMs as they drain and run its prologue (barrier init,
// TMEM allocation, weight staging) under this kernel's tail.  Such a dependent must call pdl_wait() before it touches
// anything its predecessor wrote.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();      // env MVN_PDL=0 switches the attribute off (A/B measurements)
// Launch `k` as a programmatic dependent of its predecessor in the stream when MVN_PDL_NEW=1 (its CTAs are scheduled as the
// predecessor drains); a plain launch otherwise.  `k` MUST call pdl_wait() before its first global access.  Used by the attention,
// LayerNorm-backward and partial-reduction kernels.  Measured on B200 (bench.py, CUDA-graph replay): making these kernels dependents
// changed C2 by -0.7 %, C3 by 0 % and C4 / C5 run with dependents off altogether (api.cu), so the attribute is OFF by default here --
// only the tcgen05 GEMM / fused feed-forward kernels, whose weight-staging prologue is long, launch as dependents.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_dependent(void (*k)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool on = getenv("MVN_PDL_NEW") && getenv("MVN_PDL_NEW")[0] == '1';
    cfg.attrs = at; cfg.numAttrs = (on && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k, static_cast<KArgs>(args)...);
}

// ---- dropout (nn.Dropout in Transformer / TransformerBlock, src/transformer_utils.py:112,115,147) ---------------------
// Counter-based: the keep/drop decision of element `idx` at dropout site `site` is a pure function of (seed, site, idx),
// so the backward regenerates the forward's mask instead of storing it.  thresh == 0 switches it off.
struct DropCfg {
    uint32_t thresh = 0;     // P(drop) = thresh / 2^32
    float scale = 1.0f;      // 1 / (1 - p) on kept elements
    uint32_t s0 = 0, s1 = 0; // seed words (already mixed with the site id)
    const uint32_t* ctr = nullptr;   // optional device step counter mixed into every row key (mvn_set_step_counter): lets a
                                     // CUDA-graph replay, whose kernel arguments are frozen, draw a fresh mask every step
};
// Two-level hash: a strong per-row key (computed once per row) and a cheap per-element mix of (rowkey, col).
__device__ __forceinline__ uint32_t drop_rowkey(const DropCfg& d, uint32_t row) {
    uint32_t h = row ^ d.s0;
    if (d.ctr) h += __ldg(d.ctr) * 0x9E3779B9u;
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return h ^ d.s1;
}
__device__ __forceinline__ float drop_scale(const DropCfg& d, uint32_t rowkey, uint32_t col) {
    uint32_t h = rowkey + col * 0x9E3779B1u;
    h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12; h *= 0x297a2d39u; h ^= h >> 15;
    return h >= d.thresh ? d.scale : 0.f;
}
const uint32_t* step_counter();     // api.cu: device pointer registered with mvn_set_step_counter (nullptr = none)
static inline DropCfg make_drop(float p, uint64_t seed, uint32_t site) {
    DropCfg d;
    if (!(p > 0.f)) return d;
    d.ctr = step_counter();
    const double t = (double)p * 4294967296.0;
    d.thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
    if (d.thresh == 0) d.thresh = 1;
    d.scale = 1.0f / (1.0f - p);
    d.s0 = (uint32_t)seed ^ (site * 0x85EBCA77u + 0x165667B1u);
    d.s1 = (uint32_t)(seed >> 32) ^ (site * 0xC2B2AE3Du);
    return d;
}

// ---- internal launchers shared between the per-op C entry points and the fused encoder -------------
struct GemmEpilogue {
    const float* bias = nullptr;     // [N]
    const float* addend = nullptr;   // [M,N] added before the activation / mask
    const float* act_src = nullptr;  // [M,N] source for dact
    int act = MVN_ACT_NONE;          // forward activation
    int dact = 0;                    // 0 none, 1 relu mask (act_src>0), 2 gelu'(act_src)
    // LayerNorm epilogue (N == tile width): Y = LN(acc+bias+addend)*gamma+beta
    const float* gamma = nullptr;
    const float* beta = nullptr;
    float* xhat = nullptr;
    float* rstd = nullptr;
    float eps = 1e-5f;
    DropCfg drop;                    // applied to Y after the LayerNorm affine (xhat stays pre-dropout)
};
// C[M,N] = A[M,K] * op(B):  b_is_nk: B stored [N,K] (nn.Linear weight) else B stored [K,N].
int launch_gemm(const float* A, const float* Bm, float* C, const int32_t* n_rows_dev, int M_cap, int N, int K,
                bool b_is_nk, const GemmEpilogue& ep, int prec, cudaStream_t st);
// partial[s*pstride + woff + n*K+k] = sum over slab-s rows of dY[m,n]*X[m,k];  optional bias partial at boff.
int launch_wgrad_partials(const float* dY, const float* X, const int32_t* n_rows_dev, int M_cap, int N, int K,
                          float* partial, size_t pstride, size_t woff, long long boff, int prec, cudaStream_t st);
// out[i] = (accumulate? out[i]:0) + sum_s partial[s*pstride + i], i < n
int launch_reduce_partials(const float* partial, size_t pstride, size_t n, float* out, int accumulate, cudaStream_t st);
// attention_tc.cu: work-item order of the tensor-core attention kernels (longest sequence first); nullptr = index order
void set_attention_order(const int32_t* order);
int launch_seq_order(const int32_t* cu, int B, int32_t* order, cudaStream_t st);
int launch_reduce_partials_n(const float* partial, size_t pstride, size_t n, int nslabs, float* out, int accumulate, cudaStream_t st);
// nl independent ranges in one launch: out[l * out_lstride + i] (+)= sum_s partial[s * pstride + l * in_lstride + i], i < n
int launch_reduce_partials_2d(const float* partial, size_t pstride, size_t n, int nslabs, float* out, int accumulate, int nl, size_t in_lstride,
                              size_t out_lstride, cudaStream_t st);
int launch_ln_bwd(const float* dY, const float* xhat, const float* rstd, const float* gamma, float* dZ,
                  const int32_t* n_rows_dev, int M_cap, int E, float* partial, size_t pstride, size_t goff, size_t boff,
                  cudaStream_t st, const DropCfg& drop = DropCfg());     // drop: dY is the gradient AFTER the dropout that followed this LayerNorm
int launch_embed_bwd_partials(const float* x, const int32_t* tok_src, const float* dout, const int32_t* n_rows_dev,
                              int M_cap, int T, int E, int nband, float* partial, size_t pstride, size_t off, cudaStream_t st,
                              const DropCfg& drop = DropCfg());
// fused feed-forward half of a block (ffn_fused.cu); partial offsets are floats inside one slab
bool ffn_fused_supported(int E, int ff_mult);
int launch_ffn_fused_fwd(const float* X, const float* W1, const float* b1, const float* W2, const float* b2, const float* gamma,
                         const float* beta, float* Y, float* xhat, float* rstd, const int32_t* n_rows_dev, int M_cap, int E, float eps,
                         const DropCfg& drop, cudaStream_t st);
int launch_ffn_fused_bwd(const float* dY, const float* xhat, const float* rstd, const float* X, const float* W1, const float* b1,
                         const float* W2, const float* gamma, float* dX, const int32_t* n_rows_dev, int M_cap, int E, const DropCfg& drop,
                         float* partial, size_t pstride, size_t o_w1, size_t o_b1, size_t o_w2, size_t o_b2, size_t o_g, size_t o_b,
                         float* partial2, cudaStream_t st, size_t pstride2 = 0);
size_t ffn_fused_slab_floats(int E);        // one FFN-only slab of the second slab set (w1 | b1 | w2 | b2 | gamma | beta)
int ffn_fused_bwd_slab_sets(int E);         // 2 when the backward of this width runs two CTAs per SM
int launch_embed_fwd(const float* x, const float* t, const int32_t* cu_seqlens, const int32_t* tok_src, const float* div_term,
                     const float* w, const float* b, const float* band_emb, int B, int T, int E, int nband, float* out,
                     cudaStream_t st, const DropCfg& drop);

}  // namespace mvn
