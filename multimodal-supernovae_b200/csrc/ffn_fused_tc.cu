// Fused feed-forward FORWARD on tcgen05 (fused tier):  x2 = dropout( LN( x1 + W2 relu(W1 x1 + b1) + b2 ) * gamma + beta )
// for one 128-token tile per pipeline step, with the hidden activation h living only in TMEM and shared memory:
//
//   warp 0  TMA producer : x1 tile [128 x E] as K-major 128-byte-swizzled boxes (A operand of GEMM 1, and the residual)
//   warp 1  MMA issuer   : per 32-column chunk c of h:  hacc[c%2] = x1 W1[c]^T   (TMEM, N = 32, K = E)
//                          one chunk behind:             out += h[c-1] W2[:, c-1]^T (TMEM, N = E, K = 32)
//   warps 2-5 epilogue   : hacc -> registers (one thread per token row) -> + b1, ReLU, round to TF32 -> shared memory as the K-major
//                          A operand of GEMM 2 (thread-written 128-byte swizzle, fence.proxy.async);  after the last chunk:
//                          out + b2 + x1 (residual read back from the TMA boxes) -> LayerNorm -> x2, xhat2, rstd2 through a per-warp
//                          swizzled box so that every global store is a full 128-byte line.
// W1 [4E, E] and W2 [E, 4E] are staged once per CTA as K-major swizzled blocks, rounded to nearest TF32; W1 carries the truncation
// compensation of gemm_tc.cu (x1 reaches the tensor core as raw fp32 bits), W2 does not (h is rounded here).
// The same operand patterns are exercised by gemm_tc.cu (weight staging, LayerNorm epilogue) and clip_loss_tc.cu (epilogue-written A
// operand feeding a second MMA).  The backward of the pair stays on warp MMAs (ffn_fused.cu explains why).
#include <stdlib.h>

#include "tc_common.cuh"

namespace mvn {
namespace {
using namespace tc;

constexpr int FT_M = 128;
constexpr int FT_BOX = FT_M * 128;                 // [128 rows x 32 fp32]
constexpr int FT_THREADS = 192;
constexpr int FT_WBOX = 32 * 128;                  // per-warp [32 x 32] output box

template <int E>
struct FtGeom {
    static constexpr int F = 4 * E, KC1 = E / 32, NCH = F / 32;
    static constexpr int OFF_W1 = 0;                               // KC1 blocks of [F rows x 128 B]
    static constexpr int OFF_W2 = OFF_W1 + F * E * 4;              // NCH blocks of [E rows x 128 B]
    static constexpr int OFF_X = OFF_W2 + E * F * 4;               // KC1 boxes
    static constexpr int OFF_H = OFF_X + KC1 * FT_BOX;             // 2 boxes
    static constexpr int OFF_OUT = OFF_H + 2 * FT_BOX;             // 4 warps x 1 box
    static constexpr int OFF_VEC = OFF_OUT + 4 * FT_WBOX;          // b1[F] b2[E] gamma[E] beta[E]
    static constexpr int OFF_BAR = OFF_VEC + (F + 3 * E) * 4;
    static constexpr int OFF_TMEM = OFF_BAR + 16 * 8;
    static constexpr int SMEM = OFF_TMEM + 16 + 1024;
    static constexpr int TMEM_COLS = 128;                          // hacc 2 x 32 | out E (<= 64)
};

struct FtArgs {
    const float *W1, *b1, *W2, *b2, *gamma, *beta, *X;
    float *Y, *xhat, *rstd;
    const int32_t* n_rows_dev;
    int M_cap;
    float eps, wscale;
    DropCfg drop;
};

__device__ __forceinline__ void box_row_write(uint8_t* box, int r, const float* v) {
    float4* p = reinterpret_cast<float4*>(box + r * 128);
#pragma unroll
    for (int j = 0; j < 8; ++j) p[j ^ (r & 7)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void box_row_read(const uint8_t* box, int r, float* v) {
    const float4* p = reinterpret_cast<const float4*>(box + r * 128);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 t = p[j ^ (r & 7)];
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
}
// the warp's [32 x 32] box -> global rows row0 .. row0+31 (< rows), columns c0 .. c0+31 of a [*, ld] matrix, 128-byte coalesced
__device__ __forceinline__ void box_store_rows(const uint8_t* box, float* __restrict__ dst, int row0, int rows, int ld, int c0, int lane) {
    const int cq = lane & 7, rr = lane >> 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = rr + 4 * i;
        if (row0 + r < rows)
            *reinterpret_cast<float4*>(dst + (size_t)(row0 + r) * ld + c0 + cq * 4) = *reinterpret_cast<const float4*>(box + r * 128 + ((cq ^ (r & 7)) << 4));
    }
}

template <int E>
__global__ void __launch_bounds__(FT_THREADS, 1) tc_ffn_fwd_kernel(const __grid_constant__ CUtensorMap tmapX, const FtArgs a) {
    using G = FtGeom<E>;
    constexpr int F = G::F, KC1 = G::KC1, NCH = G::NCH, EC = E / 32;
    extern __shared__ uint8_t smem_raw[];
    pdl_trigger();
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* vec = reinterpret_cast<float*>(smem + G::OFF_VEC);
    const uint32_t bar0 = sbase + G::OFF_BAR;
    const uint32_t x_full = bar0, x_empty = bar0 + 8u, out_full = bar0 + 16u, out_empty = bar0 + 24u;
    auto hacc_full = [&](int b) { return bar0 + 8u * (4 + b); };
    auto hacc_empty = [&](int b) { return bar0 + 8u * (6 + b); };
    auto hs_full = [&](int b) { return bar0 + 8u * (8 + b); };
    auto hs_empty = [&](int b) { return bar0 + 8u * (10 + b); };
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + G::OFF_TMEM);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmapX);
        mbar_init(x_full, 1); mbar_init(x_empty, 4); mbar_init(out_full, 1); mbar_init(out_empty, 4);
        for (int b = 0; b < 2; ++b) { mbar_init(hacc_full(b), 1); mbar_init(hacc_empty(b), 4); mbar_init(hs_full(b), 4); mbar_init(hs_empty(b), 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(sbase + G::OFF_TMEM, G::TMEM_COLS);
    // weights -> K-major swizzled blocks: element (n, k) of an [N, K] nn.Linear weight lives in block k/32 at sw128_off(n, k%32)
    {
        float* W1s = reinterpret_cast<float*>(smem + G::OFF_W1);
        float* W2s = reinterpret_cast<float*>(smem + G::OFF_W2);
        for (int i = threadIdx.x; i < F * E / 4; i += FT_THREADS) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(a.W1) + i);
            const int n = i / (E / 4), k = (i % (E / 4)) * 4;
            *reinterpret_cast<float4*>(W1s + (size_t)(k >> 5) * F * 32 + sw128_off(n, k & 31)) =
                make_float4(to_tf32(w.x * a.wscale), to_tf32(w.y * a.wscale), to_tf32(w.z * a.wscale), to_tf32(w.w * a.wscale));
        }
        for (int i = threadIdx.x; i < E * F / 4; i += FT_THREADS) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(a.W2) + i);
            const int n = i / (F / 4), k = (i % (F / 4)) * 4;
            *reinterpret_cast<float4*>(W2s + (size_t)(k >> 5) * E * 32 + sw128_off(n, k & 31)) = make_float4(to_tf32(w.x), to_tf32(w.y), to_tf32(w.z), to_tf32(w.w));
        }
        for (int i = threadIdx.x; i < F; i += FT_THREADS) vec[i] = a.b1 ? a.b1[i] : 0.f;
        for (int i = threadIdx.x; i < E; i += FT_THREADS) { vec[F + i] = a.b2 ? a.b2[i] : 0.f; vec[F + E + i] = a.gamma[i]; vec[F + 2 * E + i] = a.beta[i]; }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_h = tmem_base, tmem_o = tmem_base + 64;
    pdl_wait();
    int rows = a.n_rows_dev ? min(__ldg(a.n_rows_dev), a.M_cap) : a.M_cap;
    rows = __reduce_min_sync(0xffffffffu, rows);
    const int ntiles = (rows + FT_M - 1) / FT_M;

    if (warp == 0) {
        // ===== TMA producer: one x1 tile in flight (the tile is also the residual, so it is released by the final epilogue) =====
        uint32_t ph = 1;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ph ^= 1) {
            mbar_wait(x_empty, ph);
            if (elect_one()) {
                mbar_expect_tx(x_full, KC1 * FT_BOX);
                for (int kc = 0; kc < KC1; ++kc) tma_load_2d(sbase + G::OFF_X + kc * FT_BOX, &tmapX, x_full, kc * 32, tile * FT_M);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc1 = idesc_tf32(FT_M, 32, 0, 0), idesc2 = idesc_tf32(FT_M, E, 0, 0);
        const uint64_t xdesc0 = smem_desc_sw128(sbase + G::OFF_X, 0, 1024);
        const uint64_t w1desc0 = smem_desc_sw128(sbase + G::OFF_W1, 0, 1024);
        const uint64_t w2desc0 = smem_desc_sw128(sbase + G::OFF_W2, 0, 1024);
        const uint64_t hdesc0 = smem_desc_sw128(sbase + G::OFF_H, 0, 1024);
        uint32_t xph = 0, oph = 1, it = 0;          // it: running chunk counter (parities of the two-deep chunk buffers)
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, xph ^= 1, oph ^= 1) {
            mbar_wait(x_full, xph);
            mbar_wait(out_empty, oph);              // the previous tile's output accumulator has been drained
            tc_fence_after();
            for (int c = 0; c <= NCH; ++c) {
                if (c < NCH) {                      // GEMM 1 of chunk c
                    const uint32_t b = (it + c) & 1u, par = (((it + c) >> 1) & 1u) ^ 1u;
                    mbar_wait(hacc_empty(b), par);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int kc = 0; kc < KC1; ++kc)
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                umma_tf32(tmem_h + b * 32, xdesc0 + (uint32_t)(kc * (FT_BOX >> 4)) + 2u * j,
                                          w1desc0 + (uint32_t)((kc * F * 128 + c * 32 * 128) >> 4) + 2u * j, idesc1, (uint32_t)(kc | j));
                        umma_commit(hacc_full(b));
                    }
                    __syncwarp();
                }
                if (c > 0) {                        // GEMM 2 of chunk c-1
                    const uint32_t b = (it + c - 1) & 1u, par = ((it + c - 1) >> 1) & 1u;
                    mbar_wait(hs_full(b), par);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            umma_tf32(tmem_o, hdesc0 + (uint32_t)(b * (FT_BOX >> 4)) + 2u * j, w2desc0 + (uint32_t)(((c - 1) * E * 128) >> 4) + 2u * j, idesc2,
                                      (uint32_t)((c - 1) | j));
                        umma_commit(hs_empty(b));
                        if (c == NCH) umma_commit(out_full);
                    }
                    __syncwarp();
                }
            }
            it += NCH;
        }
    } else {
        // ===== epilogue warps: TMEM lanes 32*(warp%4) .. +31 =====
        const int quad = warp & 3, r = quad * 32 + lane;
        uint8_t* mybox = smem + G::OFF_OUT + (warp - 2) * FT_WBOX;
        const float* b1s = vec; const float* b2s = vec + F; const float* gs = vec + F + E; const float* bs = vec + F + 2 * E;
        uint32_t it = 0, oph = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, oph ^= 1) {
            const int row0 = tile * FT_M + quad * 32, row = row0 + lane;
            mbar_wait(x_full, oph);                                // this thread reads the TMA-written tile itself (residual) further down
            for (int c = 0; c < NCH; ++c) {
                const uint32_t b = (it + c) & 1u, par = ((it + c) >> 1) & 1u;
                mbar_wait(hacc_full(b), par);
                tc_fence_after();
                float v[32];
                tmem_ld32(tmem_h + b * 32 + ((uint32_t)(quad * 32) << 16), v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(hacc_empty(b));
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = to_tf32(fmaxf(v[j] + b1s[c * 32 + j], 0.f));
                mbar_wait(hs_empty(b), par ^ 1u);                  // GEMM 2 of the chunk that used this box two chunks ago has read it
                box_row_write(smem + G::OFF_H + b * FT_BOX, r, v);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(hs_full(b));
            }
            it += NCH;
            // ---- out + b2 + x1 -> LayerNorm ----
            mbar_wait(out_full, oph);
            tc_fence_after();
            float z[EC][32];
#pragma unroll
            for (int ec = 0; ec < EC; ++ec) {
                tmem_ld32(tmem_o + ec * 32 + ((uint32_t)(quad * 32) << 16), z[ec]);
                float xr[32];
                box_row_read(smem + G::OFF_X + ec * FT_BOX, r, xr);
#pragma unroll
                for (int j = 0; j < 32; ++j) z[ec][j] += b2s[ec * 32 + j] + xr[j];
            }
            tc_fence_before();
            fence_proxy_async();          // the residual rows were read through the generic proxy: order them before the next TMA write of the x tile (see gemm_tc.cu::aux_row)
            __syncwarp();
            if (lane == 0) { mbar_arrive(out_empty); mbar_arrive(x_empty); }     // accumulator drained, residual tile read: next tile may load
            float sum = 0.f;
#pragma unroll
            for (int ec = 0; ec < EC; ++ec)
#pragma unroll
                for (int j = 0; j < 32; ++j) sum += z[ec][j];
            const float mean = sum * (1.0f / (float)E);
            float sq = 0.f;
#pragma unroll
            for (int ec = 0; ec < EC; ++ec)
#pragma unroll
                for (int j = 0; j < 32; ++j) { z[ec][j] -= mean; sq = fmaf(z[ec][j], z[ec][j], sq); }
            const float rs = rsqrtf(sq * (1.0f / (float)E) + a.eps);
            if (a.rstd && row < rows) a.rstd[row] = rs;
            const uint32_t rk = a.drop.thresh ? drop_rowkey(a.drop, (uint32_t)row) : 0u;
#pragma unroll
            for (int ec = 0; ec < EC; ++ec) {
#pragma unroll
                for (int j = 0; j < 32; ++j) z[ec][j] *= rs;
                if (a.xhat) {
                    __syncwarp();
                    box_row_write(mybox, lane, z[ec]);
                    __syncwarp();
                    box_store_rows(mybox, a.xhat, row0, rows, E, ec * 32, lane);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    z[ec][j] = fmaf(z[ec][j], gs[ec * 32 + j], bs[ec * 32 + j]);
                    if (a.drop.thresh) z[ec][j] *= drop_scale(a.drop, rk, (uint32_t)(ec * 32 + j));
                }
                __syncwarp();
                box_row_write(mybox, lane, z[ec]);
                __syncwarp();
                box_store_rows(mybox, a.Y, row0, rows, E, ec * 32, lane);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, G::TMEM_COLS);
}

template <int E>
int launch_t(const CUtensorMap& tx, const FtArgs& a, cudaStream_t st) {
    using G = FtGeom<E>;
    static_assert(G::SMEM <= 232448, "tile set exceeds the shared memory of one CTA");
    static bool configured = false;
    if (!configured) {
        MVN_CUDA(cudaFuncSetAttribute(tc_ffn_fwd_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM));
        configured = true;
    }
    const int tiles = cdiv(a.M_cap, FT_M);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(tiles < num_sms() ? tiles : num_sms()); cfg.blockDim = dim3(FT_THREADS); cfg.dynamicSmemBytes = G::SMEM; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    MVN_CUDA(cudaLaunchKernelEx(&cfg, tc_ffn_fwd_kernel<E>, tx, a));
    MVN_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// MVN_E_UNSUPPORTED: shape outside this kernel (caller uses the warp-MMA forward of ffn_fused.cu)
int launch_ffn_fused_fwd_tc(const float* X, const float* W1, const float* b1, const float* W2, const float* b2, const float* gamma, const float* beta,
                            float* Y, float* xhat, float* rstd, const int32_t* n_rows_dev, int M_cap, int E, float eps, float wscale, const DropCfg& drop,
                            cudaStream_t st) {
    if (!(E == 32 || E == 64) || M_cap < FT_M) return MVN_E_UNSUPPORTED;
    const CUtensorMap* tx = get_tmap_2d(X, M_cap, E, FT_M, false);
    if (!tx) return MVN_E_BADARG;
    FtArgs a;
    a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.gamma = gamma; a.beta = beta; a.X = X; a.Y = Y; a.xhat = xhat; a.rstd = rstd;
    a.n_rows_dev = n_rows_dev; a.M_cap = M_cap; a.eps = eps; a.wscale = wscale; a.drop = drop;
    return E == 64 ? launch_t<64>(*tx, a, st) : launch_t<32>(*tx, a, st);
}

}  // namespace mvn
