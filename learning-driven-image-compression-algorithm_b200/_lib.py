"""ctypes binding of libldic_b200.so (the C ABI declared in include/ldic.h).

There is deliberately NO fallback: if the library is missing or the device is
not sm_100, every op raises.  Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

_LIB = None

LDIC_CONV_S2_5x5_P12 = 0
LDIC_CONV_S2_5x5_P2 = 1
LDIC_CONV_S1_3x3_P1 = 2
LDIC_CONV_1x1 = 3
LDIC_DECONV_GS_5x5 = 4
LDIC_DECONV_HS_5x5 = 5
LDIC_DECONV_S1_3x3 = 6
LDIC_DECONV_GS_5x5_MERGED = 7
LDIC_CTX_CONV1 = 8
LDIC_CTX_CONV2 = 9
LDIC_CTX_CONV3 = 10
LDIC_CTX_FC = 11
LDIC_CONV_FIRST_5x5S2 = 12
ACT_NONE, ACT_RELU, ACT_LEAKY02, ACT_GDN, ACT_IGDN, ACT_LEAKY001 = 0, 1, 2, 3, 4, 5
LDIC_EINVAL_RC = -1


class LdicError(RuntimeError):
    pass


class LikelihoodArgs(C.Structure):
    _fields_ = [
        ("v", C.c_void_p), ("v_rs", C.c_longlong), ("v_off", C.c_longlong),
        ("mu", C.c_void_p), ("mu_rs", C.c_longlong), ("mu_off", C.c_longlong), ("mu_mode", C.c_int),
        ("sigma", C.c_void_p), ("sigma_rs", C.c_longlong), ("sigma_off", C.c_longlong), ("sigma_mode", C.c_int),
        ("sigma_period", C.c_int),
        ("rows", C.c_longlong), ("cols", C.c_longlong),
        ("quant", C.c_int), ("form", C.c_int), ("sigma_is_log", C.c_int),
        ("lik_bound", C.c_float), ("scale_bound", C.c_float),
        ("v_hat", C.c_void_p), ("v_hat_rs", C.c_longlong), ("v_hat_off", C.c_longlong),
        ("v_hat_bf16", C.c_void_p), ("vb_rs", C.c_longlong), ("vb_off", C.c_longlong),
        ("lik", C.c_void_p), ("sum_ln_out", C.c_void_p), ("workspace", C.c_void_p),
    ]


class RansArgs(C.Structure):
    _fields_ = [
        ("v", C.c_void_p), ("v_rs", C.c_longlong), ("v_off", C.c_longlong),
        ("mu", C.c_void_p), ("mu_rs", C.c_longlong), ("mu_off", C.c_longlong), ("mu_mode", C.c_int),
        ("sigma", C.c_void_p), ("sigma_rs", C.c_longlong), ("sigma_off", C.c_longlong), ("sigma_mode", C.c_int),
        ("sigma_period", C.c_int),
        ("rows", C.c_longlong), ("cols", C.c_longlong), ("rows_per_segment", C.c_longlong),
        ("quant", C.c_int), ("sigma_is_log", C.c_int), ("scale_bound", C.c_float), ("streams", C.c_int),
        ("col_groups", C.c_int),
    ]


class ConvDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int),
                ("Cout", C.c_int), ("Cin_pad", C.c_int), ("Cout_pad", C.c_int), ("act", C.c_int),
                ("out_f32", C.c_int), ("aux0", C.c_int), ("aux1", C.c_int), ("sm_limit", C.c_int), ("precision", C.c_int)]


class ConvTail(C.Structure):
    _fields_ = [("x_nchw", C.c_void_p), ("w", C.c_void_p), ("x_tilde_nchw", C.c_void_p), ("sq_err", C.c_void_p),
                ("H", C.c_int), ("W", C.c_int), ("x_is_u8", C.c_int), ("tanh_out", C.c_int)]


class SyntaxArgs(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("B", "h", "w", "N", "M")] +
                [(n, C.c_void_p) for n in (
                    "y", "h2",
                    "sm_down0_w", "sm_down0_b", "sm_down1_w", "sm_down1_b", "sm_conv_w", "sm_conv_b",
                    "ps_down0_w", "ps_down0_b", "ps_down1_w", "ps_down1_b", "ps_fc_w", "ps_fc_b",
                    "cg_w0", "cg_b0", "cg_w1", "cg_b1", "cg_w2", "cg_b2",
                    "sm_ds1", "sm_ds2", "ps_ds0", "ps_ds1", "pool_part",
                    "z3", "z3_round", "mu", "sigma", "conv_w", "z3_round_in")])


_SIGS = {
    "ldic_version": (C.c_int, []),
    "ldic_last_error": (C.c_char_p, []),
    "ldic_launch_count": (C.c_longlong, []),
    "ldic_check_device": (C.c_int, [C.c_int]),
    "ldic_set_tuning": (C.c_int, [C.c_char_p, C.c_int]),
    "ldic_conv_plan_cache_size": (C.c_int, []),
    "ldic_conv_plan_cache_clear": (None, []),
    "ldic_u8_to_f32_pm1": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ldic_lower_bound": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ldic_lower_bound_bwd": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ldic_nonneg_reparam": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ldic_gdn_prepare": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ldic_gdn_nchw_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_void_p]),
    "ldic_likelihood_workspace_bytes": (C.c_size_t, []),
    "ldic_round_likelihood_bpp": (C.c_int, [C.POINTER(LikelihoodArgs), C.c_void_p]),
    "ldic_mse_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]),
    "ldic_rd_pack_metrics": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldic_rd_finish_metrics": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "ldic_syntax_conv_mse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldic_nchw_f32_to_nhwc_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_void_p]),
    "ldic_nhwc_to_nchw_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p]),
    "ldic_latent_prep": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldic_ctx_pack_input": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]),
    "ldic_im2col_5x5s2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ldic_im2col_5x5s2_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ldic_conv_weight_elems": (C.c_longlong, [C.POINTER(ConvDesc)]),
    "ldic_conv_n_cols": (C.c_int, [C.POINTER(ConvDesc)]),
    "ldic_conv_bias_elems": (C.c_int, [C.POINTER(ConvDesc)]),
    "ldic_conv_out_dims": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(C.c_int)]),
    "ldic_conv_pack_weights": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "ldic_conv_out_shape": (None, [C.POINTER(ConvDesc), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ldic_conv_forward": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "ldic_conv_forward_residual": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    "ldic_syntax_workspace_elems": (C.c_longlong, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ldic_syntax_branch": (C.c_int, [C.POINTER(SyntaxArgs), C.c_void_p]),
    "ldic_tritplane_workspace_bytes": (C.c_size_t, []),
    "ldic_tritplane_likelihood": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_float,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldic_rans_max_bytes": (C.c_size_t, [C.c_longlong, C.c_int]),
    "ldic_rans_workspace_bytes": (C.c_size_t, [C.c_longlong, C.c_longlong, C.c_int]),
    "ldic_rans_encode": (C.c_int, [C.POINTER(RansArgs), C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldic_rans_decode": (C.c_int, [C.POINTER(RansArgs), C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong,
                                  C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldic_rans_decode_begin": (C.c_int, [C.POINTER(RansArgs), C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "ldic_rans_decode_ranges": (C.c_int, [C.POINTER(RansArgs), C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_longlong, C.c_longlong,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldic_rans_phi_table": (C.POINTER(C.c_uint), [C.POINTER(C.c_int)]),
    "ldic_window_attention_bias": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ldic_window_attention_core": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ldic_window_attention_core_table": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ldic_residual_nhwc_to_nchw_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_int, C.c_void_p]),
    "ldic_gate_residual_nhwc_to_nchw_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                     C.c_int, C.c_int, C.c_void_p]),
    "ldic_debug_last_timeout": (C.c_int, [C.POINTER(C.c_ulonglong)]),
    "ldic_conv_forward_fused_tail": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.POINTER(ConvTail), C.c_void_p]),
    "ldic_conv_forward_f32_reference_kernel": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                                                         C.c_void_p, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS.keys())


def lib_path() -> str:
    # LDIC_LIB_PATH: A/B aid (load another build of the same C ABI); the default is the in-tree library
    return os.environ.get("LDIC_LIB_PATH") or _build.LIB_PATH


def load(build_if_missing: bool = False):
    """Loads (once) and returns the ctypes handle.  Raises LdicError when the
    shared library is absent -- the product path never falls back to torch/CPU."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        if build_if_missing:
            _build.build()
        else:
            raise LdicError(f"{path} not found: run `python __graft_entry__.py build` (nvcc, sm_100a). "
                            "There is no CPU/torch fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        if not hasattr(lib, name) and os.environ.get("LDIC_LIB_PATH"):
            continue                                  # older A/B build without the newest entry points
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().ldic_last_error()
        raise LdicError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
