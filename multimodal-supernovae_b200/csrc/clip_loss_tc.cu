// A10/A11 on tcgen05 (prec >= 1): the two streamed kernels of clip_loss.cu with the similarity tiles on the 5th-generation
// tensor cores.  D = 128 is a textbook UMMA shape: a 128-row tile of R against 128 (forward) / 64 (backward) columns of C is one
// accumulator in TMEM; both operands arrive by TMA as K-major [rows x 32 fp32] boxes with the 128-byte swizzle.
//
//   forward  (tc_lse_kernel):   Z = R C^T in TMEM (double buffered) -> epilogue warps: one thread per row, online log-sum-exp in
//                               base 2 (one MUFU.EX2 per logit) -> partial (max, sum) per column split.
//   backward (tc_grad_kernel):  Z tile -> epilogue turns it into G = (exp(z - lse_R) + exp(z - lse_C) - 2 I) / 2N, written to shared
//                               memory as the K-major A operand of a second MMA  dR += G C  whose accumulator [128 x 128] stays in
//                               TMEM for the CTA's whole column range.  C is loaded twice per tile (K-major for Z, and with the
//                               32-byte-atom swizzle, i.e. MN-major, for G C): the tensor core takes 32-bit MN-major operands in
//                               that layout only (see gemm_tc.cu, weight gradients).
// The N x N logits never leave the SM.  Precision: the MMA truncates both fp32 operands to TF32; the mean shrink of the products
// (2 x 3.52e-4, see gemm_tc.cu::trunc_comp) is folded into the logit scale / the gradient scale, the remainder is zero-mean.
#include <stdlib.h>

#include "tc_common.cuh"

namespace mvn {
namespace {
using namespace tc;

constexpr int LT_M = 128;                         // rows of R per CTA
constexpr int LT_BOX = LT_M * 128;                // one [128 rows x 32 fp32] box
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float TRUNC2 = 1.0f + 2.0f * 3.52e-4f;  // both MMA operands truncated

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void box_row_write(uint8_t* box, int r, const float* v) {      // row r of a [128 x 32 fp32] SW128 box
    float4* p = reinterpret_cast<float4*>(box + r * 128);
#pragma unroll
    for (int j = 0; j < 8; ++j) p[j ^ (r & 7)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
constexpr int LF_NC = 128;                        // columns of C per tile
constexpr int LF_STAGES = 2;
constexpr int LF_OFF_R = 0;
constexpr int LF_OFF_C = 4 * LT_BOX;              // 65536
constexpr int LF_OFF_BAR = LF_OFF_C + LF_STAGES * 4 * LT_BOX;      // 196608
constexpr int LF_OFF_TMEM = LF_OFF_BAR + 16 * 8;
constexpr int LF_SMEM = LF_OFF_TMEM + 16 + 1024;
constexpr int LF_EPI = 8;                         // epilogue warps: two per TMEM lane quadrant, each takes half of the tile's columns
constexpr int LF_THREADS = 32 * (2 + LF_EPI);

struct LseArgs {
    int nr, N, tiles_per_split;
    const float *logit_scale, *logit_bias;
    float *pm, *pl;
};

__global__ void __launch_bounds__(LF_THREADS, 1) tc_lse_kernel(const __grid_constant__ CUtensorMap tmapR, const __grid_constant__ CUtensorMap tmapC,
                                                               const LseArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = sbase + LF_OFF_BAR;
    const uint32_t r_full = bar0;
    auto full_bar = [&](int s) { return bar0 + 8u * (1 + s); };
    auto empty_bar = [&](int s) { return bar0 + 8u * (3 + s); };
    auto tfull_bar = [&](int b) { return bar0 + 8u * (5 + b); };
    auto tempty_bar = [&](int b) { return bar0 + 8u * (7 + b); };
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + LF_OFF_TMEM);
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmapR);
        tma_prefetch_desc(&tmapC);
        mbar_init(r_full, 1);
        for (int s = 0; s < LF_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), LF_EPI); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(sbase + LF_OFF_TMEM, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int m0 = blockIdx.x * LT_M;
    const int ct0 = blockIdx.y * a.tiles_per_split;
    const int ct1 = min(ct0 + a.tiles_per_split, (a.N + LF_NC - 1) / LF_NC);

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            mbar_expect_tx(r_full, 4 * LT_BOX);
            for (int kb = 0; kb < 4; ++kb) tma_load_2d(sbase + LF_OFF_R + kb * LT_BOX, &tmapR, r_full, kb * 32, m0);
        }
        __syncwarp();
        uint32_t s = 0, ph = 1;
        for (int ct = ct0; ct < ct1; ++ct) {
            mbar_wait(empty_bar(s), ph);
            if (elect_one()) {
                mbar_expect_tx(full_bar(s), 4 * LT_BOX);
                for (int kb = 0; kb < 4; ++kb) tma_load_2d(sbase + LF_OFF_C + (s * 4 + kb) * LT_BOX, &tmapC, full_bar(s), kb * 32, ct * LF_NC);
            }
            __syncwarp();
            if (++s == LF_STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = idesc_tf32(LT_M, LF_NC, 0, 0);
        const uint64_t adesc0 = smem_desc_sw128(sbase + LF_OFF_R, 0, 1024);
        const uint64_t bdesc0 = smem_desc_sw128(sbase + LF_OFF_C, 0, 1024);
        mbar_wait(r_full, 0);
        uint32_t s = 0, ph = 0, buf = 0, tph = 1;
        for (int ct = ct0; ct < ct1; ++ct) {
            mbar_wait(tempty_bar(buf), tph);
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        umma_tf32(tmem_base + buf * LF_NC, adesc0 + (uint32_t)(kb * (LT_BOX >> 4)) + 2u * j,
                                  bdesc0 + (uint32_t)((s * 4 + kb) * (LT_BOX >> 4)) + 2u * j, idesc, (uint32_t)(kb | j));
                umma_commit(empty_bar(s));
                umma_commit(tfull_bar(buf));
            }
            __syncwarp();
            if (++s == LF_STAGES) { s = 0; ph ^= 1; }
            if ((buf ^= 1) == 0) tph ^= 1;
        }
    } else {
        // ===== epilogue: one thread per (row, column half), online log-sum-exp in base 2; the two halves of a row are two
        // partial (max, sum) pairs that lse_finish_kernel merges like column splits =====
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int row = m0 + quad * 32 + lane;
        const float s2 = expf(__ldg(a.logit_scale)) * TRUNC2 * LOG2E, b2 = __ldg(a.logit_bias) * LOG2E;
        float m_run = -INFINITY, l_run = 0.f;
        uint32_t it = 0;
        for (int ct = ct0; ct < ct1; ++ct, ++it) {
            const int buf = it & 1;
            mbar_wait(tfull_bar(buf), (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + buf * LF_NC + half * (LF_NC / 2) + ((uint32_t)(quad * 32) << 16);
            const int c_base = ct * LF_NC + half * (LF_NC / 2);
#pragma unroll 1
            for (int c0 = 0; c0 < LF_NC / 2; c0 += 32) {
                float v[32];
                tmem_ld32(taddr + c0, v);
                const int lim = a.N - (c_base + c0);                    // columns >= N (zero rows of C from the TMA fill) are excluded
                float tm = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    v[j] = j < lim ? fmaf(v[j], s2, b2) : -INFINITY;
                    tm = fmaxf(tm, v[j]);
                }
                if (tm > m_run) { l_run *= ex2f(m_run - tm); m_run = tm; }
                if (m_run > -INFINITY) {                                // a column half that lies entirely past N contributes nothing
                    float ps = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) ps += ex2f(v[j] - m_run);
                    l_run += ps;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
        if (row < a.nr) {
            const size_t slot = ((size_t)blockIdx.y * 2 + half) * a.nr + row;
            a.pm[slot] = m_run > -INFINITY ? m_run * LN2 : -1e30f;      // natural-log units; an empty half merges as exp(-1e30 - m) * 0 = 0
            a.pl[slot] = l_run;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------------
constexpr int LB_NC = 64;                         // columns of C per tile
constexpr int LB_CBOX = LB_NC * 128;              // one [64 rows x 32 fp32] box (8 KB)
constexpr int LB_STAGES = 2;
constexpr int LB_STAGE_BYTES = 8 * LB_CBOX;       // 4 K-major boxes + 4 MN-major (atom32) boxes = 64 KB
constexpr int LB_OFF_R = 0;
constexpr int LB_OFF_C = 4 * LT_BOX;                                  // 65536
constexpr int LB_OFF_G = LB_OFF_C + LB_STAGES * LB_STAGE_BYTES;       // 196608: 2 boxes [128 x 32] = 32 KB
constexpr int LB_OFF_LSE = LB_OFF_G + 2 * LT_BOX;                     // 229376: lse_C of the tile, double buffered
constexpr int LB_OFF_RED = LB_OFF_LSE + 2 * LB_NC * 4;
constexpr int LB_OFF_BAR = LB_OFF_RED + 64;
constexpr int LB_OFF_TMEM = LB_OFF_BAR + 16 * 8;
constexpr int LB_SMEM = LB_OFF_TMEM + 16 + 1024;
constexpr int LB_EPI = 8;                         // epilogue warps: two per TMEM lane quadrant, each takes 32 of the tile's 64 columns
constexpr int LB_THREADS = 32 * (2 + LB_EPI);
static_assert(LB_SMEM <= 232448, "backward tile set exceeds the 227 KB of shared memory per CTA");

struct GradArgs {
    int nr, N, tiles_per_split, row_offset;
    const float *logit_scale, *logit_bias, *lse_R, *lse_C;
    float *dR_part, *dls_part;
};

__global__ void __launch_bounds__(LB_THREADS, 1) tc_grad_kernel(const __grid_constant__ CUtensorMap tmapR, const __grid_constant__ CUtensorMap tmapCk,
                                                                const __grid_constant__ CUtensorMap tmapCm, const GradArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = sbase + LB_OFF_BAR;
    const uint32_t r_full = bar0, g_full = bar0 + 8u, g_empty = bar0 + 16u, d_full = bar0 + 24u;
    auto full_bar = [&](int s) { return bar0 + 8u * (4 + s); };
    auto empty_bar = [&](int s) { return bar0 + 8u * (6 + s); };
    auto tfull_bar = [&](int b) { return bar0 + 8u * (8 + b); };
    auto tempty_bar = [&](int b) { return bar0 + 8u * (10 + b); };
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + LB_OFF_TMEM);
    float* lsec = reinterpret_cast<float*>(smem + LB_OFF_LSE);
    float* red = reinterpret_cast<float*>(smem + LB_OFF_RED);
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmapR);
        tma_prefetch_desc(&tmapCk);
        tma_prefetch_desc(&tmapCm);
        mbar_init(r_full, 1); mbar_init(g_full, LB_EPI); mbar_init(g_empty, 1); mbar_init(d_full, 1);
        for (int s = 0; s < LB_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), LB_EPI); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(sbase + LB_OFF_TMEM, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_z = tmem_base, tmem_d = tmem_base + 128;          // Z: 2 x 64 columns, dR: 128 columns
    const int m0 = blockIdx.x * LT_M;
    const int ct0 = blockIdx.y * a.tiles_per_split;
    const int ct1 = min(ct0 + a.tiles_per_split, (a.N + LB_NC - 1) / LB_NC);
    const int ntile = ct1 - ct0;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            mbar_expect_tx(r_full, 4 * LT_BOX);
            for (int kb = 0; kb < 4; ++kb) tma_load_2d(sbase + LB_OFF_R + kb * LT_BOX, &tmapR, r_full, kb * 32, m0);
        }
        __syncwarp();
        uint32_t s = 0, ph = 1;
        for (int ct = ct0; ct < ct1; ++ct) {
            mbar_wait(empty_bar(s), ph);
            if (elect_one()) {
                mbar_expect_tx(full_bar(s), LB_STAGE_BYTES);
                const uint32_t dst = sbase + LB_OFF_C + s * LB_STAGE_BYTES;
                for (int kb = 0; kb < 4; ++kb) tma_load_2d(dst + kb * LB_CBOX, &tmapCk, full_bar(s), kb * 32, ct * LB_NC);
                for (int nb = 0; nb < 4; ++nb) tma_load_2d(dst + (4 + nb) * LB_CBOX, &tmapCm, full_bar(s), nb * 32, ct * LB_NC);
            }
            __syncwarp();
            if (++s == LB_STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: Z_i = R C_i^T, then (one tile behind) dR += G_{i-1} C_{i-1} =====
        const uint32_t idesc_z = idesc_tf32(LT_M, LB_NC, 0, 0);
        const uint32_t idesc_d = idesc_tf32(LT_M, 128, 0, 1);               // A = G K-major, B = C MN-major
        const uint64_t rdesc0 = smem_desc_sw128(sbase + LB_OFF_R, 0, 1024);
        const uint64_t gdesc0 = smem_desc_sw128(sbase + LB_OFF_G, 0, 1024);
        mbar_wait(r_full, 0);
        uint32_t tph = 1, gph = 0;
        auto issue_d = [&](int i) {                                           // second MMA of tile i (stage i % 2)
            const uint32_t s = (uint32_t)i & 1u;
            mbar_wait(g_full, gph);
            gph ^= 1;
            tc_fence_after();
            if (elect_one()) {
                const uint64_t cm0 = smem_desc_sw128_mn32(sbase + LB_OFF_C + s * LB_STAGE_BYTES + 4 * LB_CBOX, LB_CBOX, 512);
#pragma unroll
                for (int ks = 0; ks < LB_NC / 8; ++ks)                        // 8 columns of the tile = 8 rows of the MN-major boxes = 1024 B
                    umma_tf32(tmem_d, gdesc0 + (uint32_t)((ks >> 2) * (LT_BOX >> 4)) + 2u * (ks & 3), cm0 + 64u * ks, idesc_d, (uint32_t)(i | ks));
                umma_commit(g_empty);
                umma_commit(empty_bar(s));
            }
            __syncwarp();
        };
        for (int i = 0; i < ntile; ++i) {
            const uint32_t s = (uint32_t)i & 1u, buf = (uint32_t)i & 1u;
            mbar_wait(tempty_bar(buf), tph);
            mbar_wait(full_bar(s), (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t ck0 = smem_desc_sw128(sbase + LB_OFF_C + s * LB_STAGE_BYTES, 0, 1024);
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        umma_tf32(tmem_z + buf * LB_NC, rdesc0 + (uint32_t)(kb * (LT_BOX >> 4)) + 2u * j, ck0 + (uint32_t)(kb * (LB_CBOX >> 4)) + 2u * j,
                                  idesc_z, (uint32_t)(kb | j));
                umma_commit(tfull_bar(buf));
            }
            __syncwarp();
            if (buf == 1) tph ^= 1;
            if (i > 0) issue_d(i - 1);
        }
        if (ntile > 0) {
            issue_d(ntile - 1);
            if (elect_one()) umma_commit(d_full);
            __syncwarp();
        }
    } else {
        // ===== epilogue: Z -> G (shared memory, K-major), finally dR from TMEM; one thread per (row, 32-column half of the tile) =====
        const int quad = warp & 3, half = (warp - 2) >> 2, et = (warp - 2) * 32 + lane;      // et: 0..255 among the epilogue threads
        const int r = quad * 32 + lane, row = m0 + r;
        const float s_nat = expf(__ldg(a.logit_scale)) * TRUNC2, b_nat = __ldg(a.logit_bias);
        const float s2 = s_nat * LOG2E, b2 = b_nat * LOG2E;
        const float lr2 = row < a.nr ? __ldg(a.lse_R + row) * LOG2E : 0.f;
        const float inv2n = 0.5f / (float)a.N;
        float dls = 0.f;
        uint8_t* gbox = smem + LB_OFF_G + half * LT_BOX;
        for (int i = 0; i < ntile; ++i) {
            const int buf = i & 1, c_base = (ct0 + i) * LB_NC;
            if (et < LB_NC) lsec[buf * LB_NC + et] = (c_base + et < a.N) ? __ldg(a.lse_C + c_base + et) * LOG2E : INFINITY;   // +inf: exp -> 0
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(tfull_bar(buf), (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            mbar_wait(g_empty, (uint32_t)(i & 1) ^ 1u);                        // the second MMA of tile i-1 has read the G boxes
            float v[32];
            tmem_ld32(tmem_z + buf * LB_NC + 32 * half + ((uint32_t)(quad * 32) << 16), v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int c = c_base + 32 * half + j;
                const float z2 = fmaf(v[j], s2, b2);
                float g = ex2f(z2 - lr2) + ex2f(z2 - lsec[buf * LB_NC + 32 * half + j]);
                if (row + a.row_offset == c) g -= 2.0f;
                g = (row < a.nr && c < a.N) ? g * inv2n : 0.f;
                dls = fmaf(g, z2 - b2, dls);                                  // (z - b) in base-2 units; rescaled by ln 2 at the end
                v[j] = g;
            }
            box_row_write(gbox, r, v);
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { mbar_arrive(tempty_bar(buf)); mbar_arrive(g_full); }
        }
        // dR of this CTA's column range: this thread's row, columns 64*half .. +63
        if (ntile > 0) {
            mbar_wait_sleep(d_full, 0);
            tc_fence_after();
        }
        const uint32_t taddr = tmem_d + 64 * half + ((uint32_t)(quad * 32) << 16);
        float* o = a.dR_part + ((size_t)blockIdx.y * a.nr + row) * 128 + 64 * half;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
            float v[32];
            if (ntile > 0) {
                tmem_ld32(taddr + c0, v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            if (row < a.nr) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(o + c0 + j) = make_float4(v[j] * TRUNC2, v[j + 1] * TRUNC2, v[j + 2] * TRUNC2, v[j + 3] * TRUNC2);
            }
        }
        if (a.dls_part) {
            dls = warp_sum(dls * LN2);
            if (lane == 0) red[warp - 2] = dls;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (et == 0) {
                float t_ = 0.f;
#pragma unroll
                for (int w = 0; w < LB_EPI; ++w) t_ += red[w];
                a.dls_part[blockIdx.y * gridDim.x + blockIdx.x] = t_;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

}  // namespace

// column split: as many CTAs as ~2 per SM; tiles_per_split in units of `nc` columns
void tc_loss_split(int n, int N, int nc, int* row_tiles, int* nsplit, int* tiles_per_split) {
    const int rt = cdiv(n, LT_M), ctiles = cdiv(N, nc);
    int want = cdiv(2 * num_sms(), rt);
    if (want < 1) want = 1;
    if (want > ctiles) want = ctiles;
    const int tps = cdiv(ctiles, want);
    *row_tiles = rt; *tiles_per_split = tps; *nsplit = cdiv(ctiles, tps);
}

bool tc_loss_supported(int n, int N, int D) { return D == 128 && n >= 1 && N >= 1; }

int launch_lse_tc(const float* R, const float* C, int nr, int N, const float* logit_scale, const float* logit_bias, int tiles_per_split, int row_tiles,
                  int nsplit, float* pm, float* pl, cudaStream_t st) {
    const CUtensorMap* tr = get_tmap_2d(R, nr, 128, LT_M, false);
    const CUtensorMap* tcm = get_tmap_2d(C, N, 128, LF_NC, false);
    if (!tr || !tcm) return MVN_E_BADARG;
    static bool configured = false;
    if (!configured) {
        MVN_CUDA(cudaFuncSetAttribute(tc_lse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LF_SMEM));
        configured = true;
    }
    LseArgs a;
    a.nr = nr; a.N = N; a.tiles_per_split = tiles_per_split; a.logit_scale = logit_scale; a.logit_bias = logit_bias; a.pm = pm; a.pl = pl;
    tc_lse_kernel<<<dim3(row_tiles, nsplit), LF_THREADS, LF_SMEM, st>>>(*tr, *tcm, a);
    MVN_LAUNCH_CHECK();
    return 0;
}

int launch_grad_tc(const float* R, const float* C, int nr, int N, const float* logit_scale, const float* logit_bias, const float* lse_R,
                   const float* lse_C, int row_offset, int tiles_per_split, int row_tiles, int nsplit, float* dR_part, float* dls_part, cudaStream_t st) {
    const CUtensorMap* tr = get_tmap_2d(R, nr, 128, LT_M, false);
    const CUtensorMap* tck = get_tmap_2d(C, N, 128, LB_NC, false);
    const CUtensorMap* tcm = get_tmap_2d(C, N, 128, LB_NC, true);
    if (!tr || !tck || !tcm) return MVN_E_BADARG;
    static bool configured = false;
    if (!configured) {
        MVN_CUDA(cudaFuncSetAttribute(tc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LB_SMEM));
        configured = true;
    }
    GradArgs a;
    a.nr = nr; a.N = N; a.tiles_per_split = tiles_per_split; a.row_offset = row_offset; a.logit_scale = logit_scale; a.logit_bias = logit_bias;
    a.lse_R = lse_R; a.lse_C = lse_C; a.dR_part = dR_part; a.dls_part = dls_part;
    tc_grad_kernel<<<dim3(row_tiles, nsplit), LB_THREADS, LB_SMEM, st>>>(*tr, *tck, *tcm, a);
    MVN_LAUNCH_CHECK();
    return 0;
}

}  // namespace mvn
