// Whole sequence encoder (A1-A7): TransformerWithTimeEmbeddings.forward + <modality>_projection + L2 norm,
// forward and backward, as a fixed launch sequence over one caller-owned workspace.  No host reads of device
// data: the live token count stays on the device (cu_seqlens[B]), so the sequence is CUDA-graph capturable.
#include <stdlib.h>

#include "common.cuh"

namespace mvn {
namespace {

struct ParamOff {          // offsets in floats into the flat parameter / gradient buffer
    size_t emb_w, emb_b, band, layer0, layer_stride, proj_w, proj_b, mproj_w, mproj_b, total;
    // within a layer
    size_t wqkv, wu, bu, g1, b1n, w1, b1, w2, b2, g2, b2n;
};

ParamOff param_offsets(const mvn_seq_cfg& c) {
    ParamOff o;
    const size_t E = c.E, F = (size_t)c.ff_mult * c.E;
    size_t p = 0;
    o.emb_w = p; p += E;
    o.emb_b = p; p += E;
    o.band = p; if (c.nband > 1) p += (size_t)c.nband * E;
    o.layer0 = p;
    size_t q = 0;
    o.wqkv = q; q += 3 * E * E;
    o.wu = q; q += E * E;
    o.bu = q; q += E;
    o.g1 = q; q += E;
    o.b1n = q; q += E;
    o.w1 = q; q += F * E;
    o.b1 = q; q += F;
    o.w2 = q; q += E * F;
    o.b2 = q; q += E;
    o.g2 = q; q += E;
    o.b2n = q; q += E;
    o.layer_stride = q;
    p += q * (size_t)c.depth;
    if (c.agg != MVN_AGG_NONE) {
        o.proj_w = p; p += (size_t)c.n_out * E;
        o.proj_b = p; p += c.n_out;
        o.mproj_w = p; p += (size_t)c.enc_dim * c.n_out;      // enc_dim == 0: no <modality>_projection
        o.mproj_b = p; p += c.enc_dim;
    } else {
        o.proj_w = o.proj_b = o.mproj_w = o.mproj_b = p;
    }
    o.total = p;
    return o;
}

struct LayerBuf { float *qkv, *att, *lse, *xhat1, *rstd1, *x1, *h, *xhat2, *rstd2, *x2; };

struct Workspace {
    int32_t *cu, *tok_src, *order; uint8_t* keyvalid;
    float* x0;
    float *pooled, *p1, *p2, *ynorm, *norm; int32_t* argmax;
    float *dX, *dA, *dz, *dqkv, *dh, *d_p2, *d_p1, *d_pooled;
    float* partial; size_t pstride;
    float* partial2;        // second (FFN-only) slab set of the fused backward, when it runs two CTAs per SM
    char* layer_base; size_t layer_bytes;
    size_t bytes;
    size_t M, E, F, H;
    bool fuse_ffn;          // prec 2: h / dh are never materialised

    LayerBuf layer(int l) const {
        char* p = layer_base + (size_t)l * layer_bytes;
        LayerBuf b;
        auto take = [&](size_t nfloat) { float* r = (float*)p; p += align_up(nfloat * sizeof(float), 256); return r; };
        b.qkv = take(M * 3 * E); b.att = take(M * E); b.lse = take(M * H);
        b.xhat1 = take(M * E); b.rstd1 = take(M); b.x1 = take(M * E);
        b.h = fuse_ffn ? nullptr : take(M * F); b.xhat2 = take(M * E); b.rstd2 = take(M); b.x2 = take(M * E);
        return b;
    }
};

Workspace carve(const mvn_seq_cfg& c, void* base) {
    Workspace w;
    const size_t M = (size_t)c.B * c.T, E = c.E, F = (size_t)c.ff_mult * c.E, H = c.H, B = c.B;
    w.M = M; w.E = E; w.F = F; w.H = H;
    w.fuse_ffn = c.prec == 2 && ffn_fused_supported(c.E, c.ff_mult);
    char* p = (char*)base;
    auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes, 256); return r; };
    w.cu = (int32_t*)take((B + 1) * 4);
    w.tok_src = (int32_t*)take(M * 4);
    w.order = (int32_t*)take(B * 4);
    w.keyvalid = (uint8_t*)take(M);
    w.x0 = (float*)take(M * E * 4);
    w.pooled = (float*)take(B * E * 4);
    w.argmax = (int32_t*)take(B * E * 4);
    w.p1 = (float*)take(B * (size_t)c.n_out * 4);
    const size_t Dmax = (size_t)(c.enc_dim > c.n_out ? c.enc_dim : c.n_out);
    w.p2 = (float*)take(B * Dmax * 4);
    w.ynorm = (float*)take(B * Dmax * 4);
    w.norm = (float*)take(B * 4);
    w.dX = (float*)take(M * E * 4);
    w.dA = (float*)take(M * E * 4);
    w.dz = (float*)take(M * E * 4);
    w.dqkv = (float*)take(M * 3 * E * 4);
    w.dh = w.fuse_ffn ? nullptr : (float*)take(M * F * 4);
    w.d_p2 = (float*)take(B * Dmax * 4);
    w.d_p1 = (float*)take(B * (size_t)c.n_out * 4);
    w.d_pooled = (float*)take(B * E * 4);
    const ParamOff o = param_offsets(c);
    size_t head = (size_t)c.n_out * E + c.n_out + (size_t)c.enc_dim * c.n_out + c.enc_dim;
    size_t emb = (size_t)(2 + c.nband) * E;
    // Weight-gradient partials: a slab holds ALL layers ([slab][layer][parameter]), so the backward reduces every layer's gradients in
    // one launch at its end (one launch per layer was 18 + 13 small kernels per C4 step on the critical chain); the head / embedding
    // partials reuse the front of the same slabs before / after the layers.
    const size_t nl = c.depth > 0 ? (size_t)c.depth : 1;
    w.pstride = o.layer_stride * nl;
    if (head > w.pstride) w.pstride = head;
    if (emb > w.pstride) w.pstride = emb;
    w.partial = (float*)take((size_t)kSlabs * w.pstride * 4);
    w.partial2 = (w.fuse_ffn && ffn_fused_bwd_slab_sets(c.E) == 2) ? (float*)take((size_t)kSlabs * ffn_fused_slab_floats(c.E) * nl * 4) : nullptr;
    // per-layer saved activations
    {
        size_t lb = 0;
        auto add = [&](size_t nfloat) { lb += align_up(nfloat * sizeof(float), 256); };
        add(M * 3 * E); add(M * E); add(M * H); add(M * E); add(M); add(M * E); if (!w.fuse_ffn) add(M * F); add(M * E); add(M); add(M * E);
        w.layer_bytes = lb;
    }
    w.layer_base = p;
    p += w.layer_bytes * (size_t)c.depth;
    w.bytes = (size_t)(p - (char*)base);
    return w;
}

int check_cfg(const mvn_seq_cfg* c) {
    MVN_CHECK_ARG(c != nullptr, "seq_encoder: null cfg");
    MVN_CHECK_ARG(c->B > 0 && c->T > 0 && c->E > 0 && c->H > 0 && c->depth >= 0 && c->nband >= 1, "seq_encoder: non-positive dims");
    MVN_CHECK_ARG((long long)c->B * c->T < (1ll << 31), "seq_encoder: B*T overflows int32");
    MVN_UNSUPPORTED(c->E == 16 || c->E == 32 || c->E == 64 || c->E == 128, "seq_encoder: emb=%d not in {16,32,64,128}", c->E);
    MVN_UNSUPPORTED(c->E % c->H == 0, "seq_encoder: emb %d not divisible by heads %d", c->E, c->H);
    const int hd = c->E / c->H;
    MVN_UNSUPPORTED(hd == 4 || hd == 8 || hd == 16 || hd == 32, "seq_encoder: head dim %d not in {4,8,16,32}", hd);
    MVN_UNSUPPORTED(c->nband <= 4 && c->T % c->nband == 0, "seq_encoder: nband=%d must be <=4 and divide T=%d", c->nband, c->T);
    MVN_UNSUPPORTED(c->agg == MVN_AGG_MEAN || c->agg == MVN_AGG_MAX || c->agg == MVN_AGG_NONE, "seq_encoder: agg=%d unsupported", c->agg);
    MVN_UNSUPPORTED(c->ff_mult >= 1 && c->ff_mult <= 8, "seq_encoder: ff_mult=%d unsupported", c->ff_mult);
    MVN_CHECK_ARG(c->dropout_p >= 0.0f && c->dropout_p < 1.0f, "seq_encoder: dropout_p=%g outside [0,1)", (double)c->dropout_p);
    if (c->agg != MVN_AGG_NONE) MVN_CHECK_ARG(c->n_out > 0 && c->enc_dim >= 0, "seq_encoder: n_out must be positive, enc_dim >= 0");
    return 0;
}

}  // namespace
}  // namespace mvn

using namespace mvn;

extern "C" size_t mvn_seq_param_count(const mvn_seq_cfg* cfg) {
    if (check_cfg(cfg) != 0) return 0;
    return param_offsets(*cfg).total;
}

// tensor-core attention launches take their (sequence, head) work items longest sequence first (attention_tc.cu::seq_order_kernel);
// the order is computed once per forward and kept in the workspace for the backward.  MVN_ATTN_ORDER=0: index order (A/B).
static bool attention_ordered(const mvn_seq_cfg& c) {
    static const bool lpt = !(getenv("MVN_ATTN_ORDER") && getenv("MVN_ATTN_ORDER")[0] == '0');
    return lpt && c.prec >= 1 && (size_t)c.B * 4 <= 160 * 1024;
}

extern "C" size_t mvn_seq_workspace_bytes(const mvn_seq_cfg* cfg) {
    if (check_cfg(cfg) != 0) return 0;
    return carve(*cfg, nullptr).bytes + 256;
}

extern "C" int mvn_seq_encoder_fwd(const mvn_seq_cfg* cfg, const float* params, const float* div_term, const float* x, const float* t,
                                   const uint8_t* mask, float* out, void* workspace, size_t workspace_bytes, void* stream) {
    MVN_TRY(check_cfg(cfg));
    MVN_CHECK_ARG(params && div_term && x && t && out && workspace, "seq_encoder_fwd: null pointer");
    MVN_CHECK_ARG(aligned16(params) && aligned16(workspace) && aligned16(out), "seq_encoder_fwd: params/workspace/out must be 16-byte aligned");
    const mvn_seq_cfg& c = *cfg;
    void* wsbase = (void*)align_up((size_t)workspace, 256);
    const Workspace w = carve(c, wsbase);
    if ((char*)wsbase + w.bytes > (char*)workspace + workspace_bytes) { set_error("seq_encoder_fwd: workspace %zu < %zu", workspace_bytes, w.bytes + 256); return MVN_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    const ParamOff o = param_offsets(c);
    const int M = c.B * c.T, E = c.E, F = c.ff_mult * c.E;
    const int32_t* nrows = w.cu + c.B;
    const float scale = 1.0f / sqrtf((float)E);
    const int gp = c.prec >= 1 ? 1 : 0;                                   // precision of the per-op GEMM / attention launches
    const bool fuse_ffn = c.prec == 2 && ffn_fused_supported(E, c.ff_mult);  // prec 2: h stays on chip (ffn_fused.cu)

    MVN_TRY(mvn_pack_plan(mask, c.B, c.T, 1, w.cu, w.tok_src, w.keyvalid, st));
    const bool ordered = attention_ordered(c);
    if (ordered) MVN_TRY(launch_seq_order(w.cu, c.B, w.order, st));
    // dropout sites (src/transformer_utils.py:147,112,115): 0 = transformer input, 1+2l = after norm1, 2+2l = after norm2 of layer l
    MVN_TRY(launch_embed_fwd(x, t, w.cu, w.tok_src, div_term, params + o.emb_w, params + o.emb_b, c.nband > 1 ? params + o.band : nullptr,
                             c.B, c.T, E, c.nband, w.x0, st, make_drop(c.dropout_p, c.seed, 0)));
    const float* xin = w.x0;
    for (int l = 0; l < c.depth; ++l) {
        const float* P = params + o.layer0 + (size_t)l * o.layer_stride;
        const LayerBuf lb = w.layer(l);
        GemmEpilogue e0;
        MVN_TRY(launch_gemm(xin, P + o.wqkv, lb.qkv, nrows, M, 3 * E, E, true, e0, gp, st));
        set_attention_order(ordered ? w.order : nullptr);
        const int ra = mvn_attention_fwd(lb.qkv, w.cu, nullptr, lb.att, lb.lse, c.B, E, c.H, scale, gp, st);
        set_attention_order(nullptr);
        MVN_TRY(ra);
        GemmEpilogue e1;
        e1.bias = P + o.bu; e1.addend = xin; e1.gamma = P + o.g1; e1.beta = P + o.b1n; e1.xhat = lb.xhat1; e1.rstd = lb.rstd1; e1.eps = c.ln_eps;
        e1.drop = make_drop(c.dropout_p, c.seed, 1 + 2 * l);
        MVN_TRY(launch_gemm(lb.att, P + o.wu, lb.x1, nrows, M, E, E, true, e1, gp, st));
        if (fuse_ffn) {
            MVN_TRY(launch_ffn_fused_fwd(lb.x1, P + o.w1, P + o.b1, P + o.w2, P + o.b2, P + o.g2, P + o.b2n, lb.x2, lb.xhat2, lb.rstd2, nrows, M, E,
                                         c.ln_eps, make_drop(c.dropout_p, c.seed, 2 + 2 * l), st));
            xin = lb.x2;
            continue;
        }
        GemmEpilogue e2;
        e2.bias = P + o.b1; e2.act = MVN_ACT_RELU;
        MVN_TRY(launch_gemm(lb.x1, P + o.w1, lb.h, nrows, M, F, E, true, e2, gp, st));
        GemmEpilogue e3;
        e3.bias = P + o.b2; e3.addend = lb.x1; e3.gamma = P + o.g2; e3.beta = P + o.b2n; e3.xhat = lb.xhat2; e3.rstd = lb.rstd2; e3.eps = c.ln_eps;
        e3.drop = make_drop(c.dropout_p, c.seed, 2 + 2 * l);
        MVN_TRY(launch_gemm(lb.h, P + o.w2, lb.x2, nrows, M, E, F, true, e3, gp, st));
        xin = lb.x2;
    }
    if (c.agg == MVN_AGG_NONE) return mvn_unpack_rows(xin, w.tok_src, nullptr, nrows, M, E, out, st);

    MVN_TRY(mvn_pool_fwd(xin, w.cu, nullptr, c.B, c.T, E, c.agg, w.pooled, w.argmax, st));
    GemmEpilogue ep;
    ep.bias = params + o.proj_b;
    MVN_TRY(launch_gemm(w.pooled, params + o.proj_w, w.p1, nullptr, c.B, c.n_out, E, true, ep, 0, st));
    // final feature width: enc_dim when the <modality>_projection is part of the call, else n_out
    const int D = c.enc_dim > 0 ? c.enc_dim : c.n_out;
    const float* feat = w.p1;
    if (c.enc_dim > 0) {
        GemmEpilogue em;
        em.bias = params + o.mproj_b;
        MVN_TRY(launch_gemm(w.p1, params + o.mproj_w, w.p2, nullptr, c.B, c.enc_dim, c.n_out, true, em, 0, st));
        feat = w.p2;
    }
    if (c.normalize) {
        MVN_TRY(mvn_l2norm_fwd(feat, w.ynorm, w.norm, c.B, D, st));
        feat = w.ynorm;
    }
    MVN_CUDA(cudaMemcpyAsync(out, feat, (size_t)c.B * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int mvn_seq_encoder_bwd(const mvn_seq_cfg* cfg, const float* params, const float* x, const float* dout, float* grads,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    MVN_TRY(check_cfg(cfg));
    MVN_CHECK_ARG(params && x && dout && grads && workspace, "seq_encoder_bwd: null pointer");
    MVN_CHECK_ARG(aligned16(params) && aligned16(grads) && aligned16(dout), "seq_encoder_bwd: params/grads/dout must be 16-byte aligned");
    const mvn_seq_cfg& c = *cfg;
    void* wsbase = (void*)align_up((size_t)workspace, 256);
    const Workspace w = carve(c, wsbase);
    if ((char*)wsbase + w.bytes > (char*)workspace + workspace_bytes) { set_error("seq_encoder_bwd: workspace too small"); return MVN_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    const ParamOff o = param_offsets(c);
    const int M = c.B * c.T, E = c.E, F = c.ff_mult * c.E;
    const int32_t* nrows = w.cu + c.B;
    const float scale = 1.0f / sqrtf((float)E);
    const int gp = c.prec >= 1 ? 1 : 0;
    const bool fuse_ffn = c.prec == 2 && ffn_fused_supported(E, c.ff_mult);
    const bool ordered = attention_ordered(c);
    float* part = w.partial;
    const size_t ps = w.pstride;

    const float* xlast = c.depth > 0 ? w.layer(c.depth - 1).x2 : w.x0;
    (void)xlast;
    if (c.agg == MVN_AGG_NONE) {
        MVN_TRY(mvn_pack_rows(dout, w.tok_src, nullptr, nrows, M, E, w.dX, st));
    } else {
        const int D = c.enc_dim > 0 ? c.enc_dim : c.n_out;
        const float* dfeat = dout;
        if (c.normalize) {
            MVN_TRY(mvn_l2norm_bwd(dout, w.ynorm, w.norm, w.d_p2, c.B, D, st));
            dfeat = w.d_p2;
        }
        // head partial layout: [proj_w | proj_b | mproj_w | mproj_b] exactly as in the flat parameter buffer
        const size_t h_pw = 0, h_pb = (size_t)c.n_out * E, h_mw = h_pb + c.n_out, h_mb = h_mw + (size_t)c.enc_dim * c.n_out;
        GemmEpilogue e0;
        const float* d_p1 = dfeat;
        if (c.enc_dim > 0) {
            MVN_TRY(launch_wgrad_partials(dfeat, w.p1, nullptr, c.B, c.enc_dim, c.n_out, part, ps, h_mw, (long long)h_mb, 0, st));
            MVN_TRY(launch_gemm(dfeat, params + o.mproj_w, w.d_p1, nullptr, c.B, c.n_out, c.enc_dim, false, e0, 0, st));
            d_p1 = w.d_p1;
        }
        MVN_TRY(launch_wgrad_partials(d_p1, w.pooled, nullptr, c.B, c.n_out, E, part, ps, h_pw, (long long)h_pb, 0, st));
        MVN_TRY(launch_gemm(d_p1, params + o.proj_w, w.d_pooled, nullptr, c.B, E, c.n_out, false, e0, 0, st));
        MVN_TRY(launch_reduce_partials(part, ps, h_mb + c.enc_dim, grads + o.proj_w, 0, st));
        MVN_TRY(mvn_pool_bwd(w.d_pooled, w.cu, nullptr, w.argmax, c.B, c.T, E, c.agg, w.dX, st));
    }

    for (int l = c.depth - 1; l >= 0; --l) {
        const float* P = params + o.layer0 + (size_t)l * o.layer_stride;
        const LayerBuf lb = w.layer(l);
        const float* xin = l > 0 ? w.layer(l - 1).x2 : w.x0;
        float* lpart = part + (size_t)l * o.layer_stride;                          // this layer's region inside every slab
        float* lpart2 = w.partial2 ? w.partial2 + (size_t)l * ffn_fused_slab_floats(E) : nullptr;
        if (fuse_ffn) {
            // norm2 backward + ff.2 / ff.0 input and weight gradients in one kernel; h is recomputed from x1 on chip
            MVN_TRY(launch_ffn_fused_bwd(w.dX, lb.xhat2, lb.rstd2, lb.x1, P + o.w1, P + o.b1, P + o.w2, P + o.g2, w.dA, nrows, M, E,
                                         make_drop(c.dropout_p, c.seed, 2 + 2 * l), lpart, ps, o.w1, o.b1, o.w2, o.b2, o.g2, o.b2n, lpart2, st,
                                         (size_t)c.depth * ffn_fused_slab_floats(E)));
        } else {
        // norm2 backward: dX (grad of x2) -> dz2 in w.dz
        MVN_TRY(launch_ln_bwd(w.dX, lb.xhat2, lb.rstd2, P + o.g2, w.dz, nrows, M, E, lpart, ps, o.g2, o.b2n, st, make_drop(c.dropout_p, c.seed, 2 + 2 * l)));
        // ff.2: dW2 = dz2^T h ; dh = (dz2 W2) * relu'(h)
        MVN_TRY(launch_wgrad_partials(w.dz, lb.h, nrows, M, E, F, lpart, ps, o.w2, (long long)o.b2, gp, st));
        GemmEpilogue eh;
        eh.act_src = lb.h; eh.dact = 1;
        MVN_TRY(launch_gemm(w.dz, P + o.w2, w.dh, nrows, M, F, E, false, eh, gp, st));
        // ff.0: dW1 = dh^T x1 ; dx1 = dh W1 + dz2 (residual)
        MVN_TRY(launch_wgrad_partials(w.dh, lb.x1, nrows, M, F, E, lpart, ps, o.w1, (long long)o.b1, gp, st));
        GemmEpilogue e1;
        e1.addend = w.dz;
        MVN_TRY(launch_gemm(w.dh, P + o.w1, w.dA, nrows, M, E, F, false, e1, gp, st));
        }
        // norm1 backward: dA -> dz1 in w.dz
        MVN_TRY(launch_ln_bwd(w.dA, lb.xhat1, lb.rstd1, P + o.g1, w.dz, nrows, M, E, lpart, ps, o.g1, o.b1n, st, make_drop(c.dropout_p, c.seed, 1 + 2 * l)));
        // unifyheads: dWu = dz1^T att ; datt = dz1 Wu  (into w.dA)
        MVN_TRY(launch_wgrad_partials(w.dz, lb.att, nrows, M, E, E, lpart, ps, o.wu, (long long)o.bu, gp, st));
        GemmEpilogue e2;
        MVN_TRY(launch_gemm(w.dz, P + o.wu, w.dA, nrows, M, E, E, false, e2, gp, st));
        set_attention_order(ordered ? w.order : nullptr);
        const int ra = mvn_attention_bwd(lb.qkv, w.cu, nullptr, lb.att, lb.lse, w.dA, w.dqkv, c.B, E, c.H, scale, gp, st);
        set_attention_order(nullptr);
        MVN_TRY(ra);
        // q/k/v projections: dWqkv = dqkv^T xin ; dxin = dqkv Wqkv + dz1 (residual) -> w.dX
        MVN_TRY(launch_wgrad_partials(w.dqkv, xin, nrows, M, 3 * E, E, lpart, ps, o.wqkv, -1, gp, st));
        GemmEpilogue e3;
        e3.addend = w.dz;
        MVN_TRY(launch_gemm(w.dqkv, P + o.wqkv, w.dX, nrows, M, E, 3 * E, false, e3, gp, st));
    }
    if (c.depth > 0) {
        // every layer's parameter gradients in one reduction, then the second slab set (feed-forward gradients of the CTAs >= kSlabs)
        MVN_TRY(launch_reduce_partials(part, ps, (size_t)c.depth * o.layer_stride, grads + o.layer0, 0, st));
        if (fuse_ffn && w.partial2)
            MVN_TRY(launch_reduce_partials_2d(w.partial2, (size_t)c.depth * ffn_fused_slab_floats(E), ffn_fused_slab_floats(E), kSlabs, grads + o.layer0 + o.w1, 1,
                                              c.depth, ffn_fused_slab_floats(E), o.layer_stride, st));
    }
    // embedding_mag / band_emb
    MVN_TRY(launch_embed_bwd_partials(x, w.tok_src, w.dX, nrows, M, c.T, E, c.nband, part, ps, 0, st, make_drop(c.dropout_p, c.seed, 0)));
    MVN_TRY(launch_reduce_partials(part, ps, (size_t)(2 + (c.nband > 1 ? c.nband : 0)) * E, grads + o.emb_w, 0, st));
    return 0;
}
