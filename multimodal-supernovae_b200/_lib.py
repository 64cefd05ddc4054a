"""ctypes binding of libmaven_sm100.so (C ABI: include/maven_sm100.h).  Loading is lazy so that host-side
logic (configs, parameter layouts) is importable without the library; any compute call fails loudly if the
library is missing -- there is no fallback path."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MVN_LIB_PATH") or os.path.join(_HERE, "libmaven_sm100.so")      # MVN_LIB_PATH: A/B builds of the same library (scripts/)

MVN_ACT_NONE, MVN_ACT_RELU, MVN_ACT_GELU = 0, 1, 2
MVN_AGG_MEAN, MVN_AGG_MAX, MVN_AGG_NONE = 0, 1, 2


class SeqCfg(Structure):
    _fields_ = [("B", c_int32), ("T", c_int32), ("E", c_int32), ("H", c_int32), ("depth", c_int32), ("nband", c_int32),
                ("n_out", c_int32), ("enc_dim", c_int32), ("agg", c_int32), ("normalize", c_int32), ("prec", c_int32),
                ("ff_mult", c_int32), ("ln_eps", c_float), ("dropout_p", c_float), ("seed", c_uint64)]


class ConvCfg(Structure):
    _fields_ = [("B", c_int32), ("C", c_int32), ("H", c_int32), ("W", c_int32), ("dim", c_int32), ("depth", c_int32),
                ("kernel_size", c_int32), ("patch_size", c_int32), ("n_out", c_int32), ("enc_dim", c_int32), ("hidden", c_int32),
                ("normalize", c_int32), ("training", c_int32), ("prec", c_int32), ("bn_eps", c_float), ("bn_momentum", c_float),
                ("global_count", c_int64), ("dropout_p", c_float), ("seed", c_uint64)]


P = c_void_p
_SIGS = {
    "mvn_last_error": (c_char_p, []),
    "mvn_abi_version": (c_int, []),
    "mvn_num_sms": (c_int, []),
    "mvn_num_slabs": (c_int, []),
    "mvn_launch_count": (ctypes.c_longlong, []),
    "mvn_tier_count": (ctypes.c_longlong, [c_int]),
    "mvn_tier_reset": (None, []),
    "mvn_set_pdl": (None, [c_int]),
    "mvn_prof_enable": (None, [ctypes.c_uint]),
    "mvn_prof_read": (c_int, [c_int, POINTER(c_double), POINTER(ctypes.c_longlong)]),
    "mvn_pack_plan": (c_int, [P, c_int, c_int, c_int, P, P, P, P]),
    "mvn_embed_fwd": (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, P, P]),
    "mvn_embed_bwd": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P, P, P, P, c_size_t, P]),
    "mvn_linear_fwd": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "mvn_linear_res_ln_fwd": (c_int, [P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_float, c_int, P]),
    "mvn_linear_bwd_input": (c_int, [P, P, P, P, P, c_int, P, c_int, c_int, c_int, c_int, P]),
    "mvn_linear_bwd_weight": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, P, c_size_t, c_int, P]),
    "mvn_linear_bwd_weight_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mvn_ffn_fused_fwd": (c_int, [P, P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_float, c_float, c_uint64, c_int, P]),
    "mvn_ffn_fused_bwd_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mvn_ffn_fused_bwd": (c_int, [P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_float, c_uint64, c_int, P, c_size_t, P]),
    "mvn_relu_bwd": (c_int, [P, P, c_int64, P, P]),
    "mvn_layernorm_bwd": (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, P, c_size_t, P]),
    "mvn_attention_fwd": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_float, c_int, P]),
    "mvn_attention_bwd": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, c_float, c_int, P]),
    "mvn_pool_fwd": (c_int, [P, P, P, c_int, c_int, c_int, c_int, P, P, P]),
    "mvn_pool_bwd": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P, P]),
    "mvn_unpack_rows": (c_int, [P, P, P, P, c_int, c_int, P, P]),
    "mvn_pack_rows": (c_int, [P, P, P, P, c_int, c_int, P, P]),
    "mvn_attn_pool_saved_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mvn_attn_pool_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mvn_attn_pool_fwd": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, P, P, c_size_t, P]),
    "mvn_attn_pool_bwd": (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, c_size_t, P]),
    "mvn_query_pool_fwd": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P, P]),
    "mvn_query_pool_bwd": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P, P, P, c_size_t, P]),
    "mvn_l2norm_fwd": (c_int, [P, P, P, c_int, c_int, P]),
    "mvn_l2norm_bwd": (c_int, [P, P, P, P, c_int, c_int, P]),
    "mvn_dropout_scale": (c_int, [c_uint64, c_int, c_float, c_int, c_int, P, P]),
    "mvn_dropout_apply": (c_int, [P, P, c_int, c_int, c_uint64, c_int, c_float, P]),
    "mvn_augment_seq": (c_int, [P, P, P, c_float, c_int64, c_uint64, P, P]),
    "mvn_image_noise_range_workspace_bytes": (c_size_t, []),
    "mvn_image_noise_range": (c_int, [P, c_int, c_int64, c_float, P, P, c_size_t, P]),
    "mvn_augment_images": (c_int, [P, c_int, P, P, P, c_uint64, c_int, c_int, c_int, c_int, P, P]),
    "mvn_seq_param_count": (c_size_t, [POINTER(SeqCfg)]),
    "mvn_seq_workspace_bytes": (c_size_t, [POINTER(SeqCfg)]),
    "mvn_seq_encoder_fwd": (c_int, [POINTER(SeqCfg), P, P, P, P, P, P, P, c_size_t, P]),
    "mvn_seq_encoder_bwd": (c_int, [POINTER(SeqCfg), P, P, P, P, P, c_size_t, P]),
    "mvn_clip_loss_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mvn_clip_loss_fwd": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P, P, P, P, P, P, c_size_t, c_int, P]),
    "mvn_clip_loss_bwd": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, P, c_size_t, c_int, P]),
    "mvn_conv_param_count": (c_size_t, [POINTER(ConvCfg)]),
    "mvn_conv_workspace_bytes": (c_size_t, [POINTER(ConvCfg)]),
    "mvn_conv_num_bn": (c_int, [POINTER(ConvCfg)]),
    "mvn_convmixer_fwd_stage": (c_int, [POINTER(ConvCfg), c_int, P, P, P, P, P, P, c_size_t, P]),
    "mvn_convmixer_bwd_stage": (c_int, [POINTER(ConvCfg), c_int, P, P, P, P, P, P, P, c_size_t, P]),
    "mvn_weighted_ce_fwd": (c_int, [P, P, P, c_int, c_int, P, P]),
    "mvn_weighted_ce_bwd": (c_int, [P, P, P, c_int, c_int, P, P, P, P]),
    "mvn_mse_fwd": (c_int, [P, P, c_int, P, P]),
    "mvn_mse_bwd": (c_int, [P, P, c_int, P, P, P]),
    "mvn_masked_mse_fwd": (c_int, [P, P, P, c_int, P, P]),
    "mvn_masked_mse_bwd": (c_int, [P, P, P, c_int, P, P, P, P]),
    "mvn_meta_input_fwd": (c_int, [P, P, P, c_int, c_int, c_int, P, P]),
    "mvn_meta_input_bwd": (c_int, [P, P, c_int, c_int, c_int, P, P]),
    "mvn_radam_step": (c_int, [P, P, P, P, c_int64, c_float, c_float, c_float, c_float, c_float, c_float, c_float, c_float, P]),
    "mvn_radam_step_dev": (c_int, [P, P, P, P, c_int64, c_float, c_float, c_float, c_float, c_float, P, P, P]),
    "mvn_set_step_counter": (None, [P]),
    "mvn_retrieval_ranks": (c_int, [P, P, c_int, c_int, P, P]),
    "mvn_retrieval_curve": (c_int, [P, c_int, P, c_int, P, P]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                               "(maven_b200 has no CPU or PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().mvn_last_error().decode(errors="replace")
        raise RuntimeError(f"libmaven_sm100 {what} failed (code {rc}): {msg}")
