// A8, tensor-core tier (prec >= 1): the patch embedding of ConvMixer (src/models_multimodal.py:52-56, Conv2d(C, dim, p, stride p))
// as an IMPLICIT GEMM straight from the image -- no im2col buffer -- with GELU and the BatchNorm partial sums fused into the
// epilogue, and its weight gradient as the transposed implicit GEMM.
//
//   forward :  u[r, n] = sum_k img[b, c, py*p+i, px*p+j] W[n, k],  r = (b, py, px), k = (c, i, j);  a = GELU(u);  sum a, sum a^2
//   backward:  dW[n, k] = sum_r dU[r, n] img[...]          (the image is the only large operand: read once per pass, 43.2 KB/sample)
//
// Warp-level mma.sync m16n8k8, TF32 operands (round to nearest), fp32 accumulate.  The A operand of the forward (B operand of
// the backward) is gathered from the NCHW image through a k -> offset table; neighbouring k are neighbouring pixels, so the
// gathers are 16-/32-byte runs that L1 merges.  The image bytes are the roofline: 2 x 43.2 KB per sample per training step.
#include "common.cuh"

namespace mvn {
namespace {

constexpr int PC_WARPS = 8;
constexpr int PC_THREADS = PC_WARPS * 32;
constexpr int PC_MAXK = 512;

__device__ __forceinline__ float tf32r(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
__device__ __forceinline__ void mma_tf32(float* c, const float* a, float b0, float b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
          "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

struct PatchGeom { int B, C, H, W, p, Hp, Wp, P, R, Kp, Kpad; };

__device__ __forceinline__ size_t patch_base(const PatchGeom& g, int r) {      // offset of pixel (c=0, i=0, j=0) of patch r
    const int b = r / g.P, q = r % g.P, py = q / g.Wp, px = q % g.Wp;
    return (size_t)b * g.C * g.H * g.W + (size_t)py * g.p * g.W + (size_t)px * g.p;
}

// u, a: [R, DIM];  stat_part: [gridDim.x][2*DIM] doubles (nullptr: no statistics).  Persistent over 128-row blocks.
template <int DIM>
__global__ void __launch_bounds__(PC_THREADS) patch_conv_fwd_kernel(const float* __restrict__ img, const float* __restrict__ Wt, const PatchGeom g,
                                                                    float* __restrict__ u, float* __restrict__ a, double* __restrict__ stat_part) {
    constexpr int NT = DIM / 8;
    extern __shared__ __align__(16) float smem_pc[];
    const int PW = g.Kpad + 4 + ((36 - (g.Kpad + 4) % 32) % 32 + 32) % 32;       // row stride == 4 (mod 32): B-fragment reads hit banks 4g + t
    float* Ws = smem_pc;                                   // [DIM][PW]
    int* koff = reinterpret_cast<int*>(Ws + DIM * PW);     // [Kpad]
    double* red = reinterpret_cast<double*>(koff + PC_MAXK);   // [PC_WARPS][2][DIM]
    for (int i = threadIdx.x; i < DIM * g.Kpad; i += PC_THREADS) {
        const int n = i / g.Kpad, k = i % g.Kpad;
        Ws[n * PW + k] = k < g.Kp ? tf32r(__ldg(Wt + (size_t)n * g.Kp + k)) : 0.f;
    }
    for (int k = threadIdx.x; k < g.Kpad; k += PC_THREADS) {
        const int kk = k < g.Kp ? k : 0;
        const int c = kk / (g.p * g.p), i = (kk / g.p) % g.p, j = kk % g.p;
        koff[k] = (c * g.H + i) * g.W + j;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
    double s1[NT][2], s2[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { s1[nt][0] = s1[nt][1] = 0.0; s2[nt][0] = s2[nt][1] = 0.0; }
    const int nblk = (g.R + 15) / 16;
    for (int blk = warp * gridDim.x + blockIdx.x; blk < nblk; blk += gridDim.x * PC_WARPS) {
        const int r_lo = blk * 16 + gq, r_hi = r_lo + 8;
        const bool v_lo = r_lo < g.R, v_hi = r_hi < g.R;
        const float* p_lo = img + patch_base(g, v_lo ? r_lo : 0);
        const float* p_hi = img + patch_base(g, v_hi ? r_hi : 0);
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll 6
        for (int ks = 0; ks < g.Kpad / 8; ++ks) {                 // 6 k-steps = 24 independent gathers in flight per lane
            const int k0 = 8 * ks + t, k1 = k0 + 4;
            const int o0 = koff[k0], o1 = koff[k1];
            float af[4];
            af[0] = (v_lo && k0 < g.Kp) ? tf32r(__ldg(p_lo + o0)) : 0.f;
            af[1] = (v_hi && k0 < g.Kp) ? tf32r(__ldg(p_hi + o0)) : 0.f;
            af[2] = (v_lo && k1 < g.Kp) ? tf32r(__ldg(p_lo + o1)) : 0.f;
            af[3] = (v_hi && k1 < g.Kp) ? tf32r(__ldg(p_hi + o1)) : 0.f;
            const float* wp = Ws + gq * PW + k0;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[nt], af, wp[8 * nt * PW], wp[8 * nt * PW + 4]);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int col = 8 * nt + 2 * t;
            const float a0 = gelu_erf(acc[nt][0]), a1 = gelu_erf(acc[nt][1]), a2 = gelu_erf(acc[nt][2]), a3 = gelu_erf(acc[nt][3]);
            if (v_lo) {
                *reinterpret_cast<float2*>(u + (size_t)r_lo * DIM + col) = make_float2(acc[nt][0], acc[nt][1]);
                *reinterpret_cast<float2*>(a + (size_t)r_lo * DIM + col) = make_float2(a0, a1);
                s1[nt][0] += a0; s1[nt][1] += a1; s2[nt][0] += (double)a0 * a0; s2[nt][1] += (double)a1 * a1;
            }
            if (v_hi) {
                *reinterpret_cast<float2*>(u + (size_t)r_hi * DIM + col) = make_float2(acc[nt][2], acc[nt][3]);
                *reinterpret_cast<float2*>(a + (size_t)r_hi * DIM + col) = make_float2(a2, a3);
                s1[nt][0] += a2; s1[nt][1] += a3; s2[nt][0] += (double)a2 * a2; s2[nt][1] += (double)a3 * a3;
            }
        }
    }
    if (!stat_part) return;
    // column sums: over the 8 row groups of the warp (shuffles), then over the warps (shared memory), fixed order
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double v1 = s1[nt][e], v2 = s2[nt][e];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) { v1 += __shfl_xor_sync(0xffffffffu, v1, o); v2 += __shfl_xor_sync(0xffffffffu, v2, o); }
            if (gq == 0) { red[(warp * 2 + 0) * DIM + 8 * nt + 2 * t + e] = v1; red[(warp * 2 + 1) * DIM + 8 * nt + 2 * t + e] = v2; }
        }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * DIM; i += PC_THREADS) {
        const int sec = i / DIM, c = i % DIM;
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < PC_WARPS; ++w) s += red[(w * 2 + sec) * DIM + c];
        stat_part[(size_t)blockIdx.x * 2 * DIM + i] = s;
    }
}

// dW partial of this CTA's rows: partial[blockIdx.x * pstride + woff + n*Kp + k].  Warp w owns the k tiles w, w+8, ...
template <int DIM>
__global__ void __launch_bounds__(PC_THREADS) patch_conv_wgrad_kernel(const float* __restrict__ img, const float* __restrict__ dU, const PatchGeom g,
                                                                      float* __restrict__ partial, size_t pstride, size_t woff) {
    constexpr int MT = DIM / 16, RB = 64, PS = DIM + 8, MAXNT = (PC_MAXK / 8 + PC_WARPS - 1) / PC_WARPS;
    __shared__ __align__(16) float dUs[RB * PS];       // A^T operand rows (tokens = patches): reads (rows = t, cols = g) -> banks 8t + g
    __shared__ int koff[PC_MAXK];
    __shared__ size_t rbase[RB];
    for (int k = threadIdx.x; k < g.Kpad; k += PC_THREADS) {
        const int kk = k < g.Kp ? k : 0;
        const int c = kk / (g.p * g.p), i = (kk / g.p) % g.p, j = kk % g.p;
        koff[k] = (c * g.H + i) * g.W + j;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
    const int ntile = g.Kpad / 8;
    float acc[MAXNT][MT][4];
#pragma unroll
    for (int i = 0; i < MAXNT; ++i)
#pragma unroll
        for (int m = 0; m < MT; ++m) acc[i][m][0] = acc[i][m][1] = acc[i][m][2] = acc[i][m][3] = 0.f;
    const int nblk = (g.R + RB - 1) / RB;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const int row0 = blk * RB;
        __syncthreads();
        for (int i = threadIdx.x; i < RB * (DIM / 4); i += PC_THREADS) {
            const int r = i / (DIM / 4), c4 = (i % (DIM / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < g.R) v = __ldg(reinterpret_cast<const float4*>(dU + (size_t)(row0 + r) * DIM + c4));
            *reinterpret_cast<float4*>(dUs + r * PS + c4) = make_float4(tf32r(v.x), tf32r(v.y), tf32r(v.z), tf32r(v.w));
        }
        for (int r = threadIdx.x; r < RB; r += PC_THREADS) rbase[r] = patch_base(g, row0 + r < g.R ? row0 + r : 0);
        __syncthreads();
#pragma unroll 2
        for (int ks = 0; ks < RB / 8; ++ks) {
            const int r0 = 8 * ks + t, r1 = r0 + 4;
            float af[MT][4];
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                af[m][0] = dUs[r0 * PS + 16 * m + gq]; af[m][1] = dUs[r0 * PS + 16 * m + gq + 8];
                af[m][2] = dUs[r1 * PS + 16 * m + gq]; af[m][3] = dUs[r1 * PS + 16 * m + gq + 8];
            }
            const bool v0 = row0 + r0 < g.R, v1 = row0 + r1 < g.R;
            const float* p0 = img + rbase[r0];
            const float* p1 = img + rbase[r1];
#pragma unroll
            for (int i = 0; i < MAXNT; ++i) {
                const int nt = warp + i * PC_WARPS;
                if (nt < ntile) {
                    const int k = 8 * nt + gq;
                    const int o = koff[k];
                    const float b0 = (v0 && k < g.Kp) ? tf32r(__ldg(p0 + o)) : 0.f;
                    const float b1 = (v1 && k < g.Kp) ? tf32r(__ldg(p1 + o)) : 0.f;
#pragma unroll
                    for (int m = 0; m < MT; ++m) mma_tf32(acc[i][m], af[m], b0, b1);
                }
            }
        }
    }
    float* slab = partial + (size_t)blockIdx.x * pstride + woff;
#pragma unroll
    for (int i = 0; i < MAXNT; ++i) {
        const int nt = warp + i * PC_WARPS;
        if (nt >= ntile) continue;
        const int k = 8 * nt + 2 * t;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const int n_lo = 16 * m + gq, n_hi = n_lo + 8;
            if (k < g.Kp) { slab[(size_t)n_lo * g.Kp + k] = acc[i][m][0]; slab[(size_t)n_hi * g.Kp + k] = acc[i][m][2]; }
            if (k + 1 < g.Kp) { slab[(size_t)n_lo * g.Kp + k + 1] = acc[i][m][1]; slab[(size_t)n_hi * g.Kp + k + 1] = acc[i][m][3]; }
        }
    }
}

PatchGeom make_geom(const mvn_conv_cfg& c) {
    PatchGeom g;
    g.B = c.B; g.C = c.C; g.H = c.H; g.W = c.W; g.p = c.patch_size; g.Hp = c.H / c.patch_size; g.Wp = c.W / c.patch_size;
    g.P = g.Hp * g.Wp; g.R = c.B * g.P; g.Kp = c.C * c.patch_size * c.patch_size; g.Kpad = (g.Kp + 7) & ~7;
    return g;
}
int fwd_smem_bytes(const PatchGeom& g, int dim) {
    const int PW = g.Kpad + 4 + ((36 - (g.Kpad + 4) % 32) % 32 + 32) % 32;
    return (dim * PW + PC_MAXK) * 4 + PC_WARPS * 2 * dim * 8 + 16;
}

}  // namespace

bool patch_conv_tc_supported(const mvn_conv_cfg& c) {
    const int Kp = c.C * c.patch_size * c.patch_size;
    return (c.dim == 32 || c.dim == 64) && Kp <= PC_MAXK && fwd_smem_bytes(make_geom(c), c.dim) <= 200 * 1024;
}

int launch_patch_conv_fwd_tc(const mvn_conv_cfg& c, const float* img, const float* Wt, float* u, float* a, double* stat_part, cudaStream_t st) {
    const PatchGeom g = make_geom(c);
    const int smem = fwd_smem_bytes(g, c.dim);
    if (c.dim == 32) {
        static bool cfg32 = false;
        if (!cfg32) { MVN_CUDA(cudaFuncSetAttribute(patch_conv_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); cfg32 = true; }
        patch_conv_fwd_kernel<32><<<kSlabs, PC_THREADS, smem, st>>>(img, Wt, g, u, a, stat_part);
    } else {
        static bool cfg64 = false;
        if (!cfg64) { MVN_CUDA(cudaFuncSetAttribute(patch_conv_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); cfg64 = true; }
        patch_conv_fwd_kernel<64><<<kSlabs, PC_THREADS, smem, st>>>(img, Wt, g, u, a, stat_part);
    }
    MVN_LAUNCH_CHECK();
    count_tier(TIER_MMA);
    return 0;
}

int launch_patch_conv_wgrad_tc(const mvn_conv_cfg& c, const float* img, const float* dU, float* partial, size_t pstride, size_t woff, cudaStream_t st) {
    const PatchGeom g = make_geom(c);
    if (c.dim == 32) patch_conv_wgrad_kernel<32><<<kSlabs, PC_THREADS, 0, st>>>(img, dU, g, partial, pstride, woff);
    else patch_conv_wgrad_kernel<64><<<kSlabs, PC_THREADS, 0, st>>>(img, dU, g, partial, pstride, woff);
    MVN_LAUNCH_CHECK();
    count_tier(TIER_MMA);
    return 0;
}

}  // namespace mvn
