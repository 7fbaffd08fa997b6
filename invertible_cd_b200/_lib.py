"""ctypes binding of libicd_b200.so — the C ABI declared in include/icd_b200.h.

The product path has no CPU or PyTorch fallback: if the shared library is missing or a call fails,
a RuntimeError is raised (loudly), never a silent re-route.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ICD_LIB_PATH") or os.path.join(_HERE, "libicd_b200.so")   # ICD_LIB_PATH: debug builds

# every symbol include/icd_b200.h declares (tests assert the library exports all of them)
EXPORTED_SYMBOLS = [
    "icd_last_error", "icd_device_info", "icd_abi_version", "icd_set_pdl", "icd_gemm", "icd_gemm_pick_bn", "icd_attention", "icd_attention_ex",
    "icd_groupnorm", "icd_groupnorm_launches", "icd_layernorm", "icd_softmax", "icd_softmax_causal", "icd_act", "icd_embed_tokens", "icd_upsample2x", "icd_im2col_s2", "icd_im2col_s2_pad", "icd_latent_to_nhwc",
    "icd_timestep_embedding", "icd_guidance_embedding", "icd_silu", "icd_add", "icd_consistency_update",
    # fp32 validation path (ABI 3)
    "icd_sgemm_f32", "icd_groupnorm_f32", "icd_layernorm_f32", "icd_softmax_f32", "icd_silu_f32", "icd_geglu_f32",
    "icd_upsample2x_f32", "icd_im2col_s2_f32", "icd_nchw_to_nhwc_f32", "icd_nhwc_to_nchw_f32",
    "icd_sincos_embedding_f32",
]


class IcdGemm(C.Structure):
    """Mirror of `struct IcdGemm` (include/icd_b200.h). Field order and types must match exactly."""
    _fields_ = [
        ("a0", C.c_void_p), ("a1", C.c_void_p), ("a_mode", C.c_int), ("K0", C.c_int), ("K1", C.c_int),
        ("a0_ld", C.c_longlong), ("a1_ld", C.c_longlong), ("a_z1_stride", C.c_longlong),
        ("a_z2_stride", C.c_longlong), ("ZA1", C.c_int), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("b", C.c_void_p), ("b_ld", C.c_longlong), ("b_z1_stride", C.c_longlong), ("b_z2_stride", C.c_longlong),
        ("ZB1", C.c_int), ("b_mn_major", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("Z", C.c_int),
        ("alpha", C.c_float), ("bias", C.c_void_p), ("rowvec", C.c_void_p), ("rows_per_img", C.c_int),
        ("ldv", C.c_int), ("residual", C.c_void_p), ("ldr", C.c_longlong), ("res_zstride", C.c_longlong),
        ("out", C.c_void_p), ("ldc", C.c_longlong), ("out_z1_stride", C.c_longlong), ("out_z2_stride", C.c_longlong), ("out_imgstride", C.c_longlong),
        ("out_fp32", C.c_int), ("out_mode", C.c_int), ("geglu", C.c_int), ("force_bn", C.c_int), ("force_bm", C.c_int), ("force_splits", C.c_int),
        ("ws", C.c_void_p), ("ws_bytes", C.c_longlong),
        ("upd_x", C.c_void_p), ("upd_out", C.c_void_p),
        ("alpha_t", C.c_float), ("sigma_t", C.c_float), ("alpha_s", C.c_float), ("sigma_s", C.c_float),
        ("exp_stats", C.c_void_p),
    ]


class IcdSgemm(C.Structure):
    """Mirror of `struct IcdSgemm` (include/icd_b200.h): the fp32 validation contraction."""
    _fields_ = [
        ("a0", C.c_void_p), ("a1", C.c_void_p), ("C0", C.c_int), ("C1", C.c_int), ("a0_ld", C.c_longlong),
        ("a1_ld", C.c_longlong), ("conv", C.c_int), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("w_tap_ld", C.c_int), ("b", C.c_void_p), ("b_ld", C.c_longlong), ("b_kn", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("Z", C.c_int), ("ZH", C.c_int),
        ("a_zb", C.c_longlong), ("a_zh", C.c_longlong), ("b_zb", C.c_longlong), ("b_zh", C.c_longlong),
        ("c_zb", C.c_longlong), ("c_zh", C.c_longlong), ("r_zb", C.c_longlong), ("r_zh", C.c_longlong),
        ("alpha", C.c_float), ("bias", C.c_void_p), ("rowvec", C.c_void_p), ("rows_per_img", C.c_int),
        ("ldv", C.c_longlong), ("residual", C.c_void_p), ("ldr", C.c_longlong), ("out", C.c_void_p),
        ("ldc", C.c_longlong), ("vec", C.c_int),
    ]


_lib = None


def load():
    """Load the shared library (once). Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build the sm_100a extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C invertible_cd_b200/csrc). "
            "There is no CPU/PyTorch fallback for the iCD hot path.")
    lib = C.CDLL(LIB_PATH)
    lib.icd_last_error.restype = C.c_char_p
    lib.icd_abi_version.restype = C.c_int
    lib.icd_set_pdl.argtypes = [C.c_int]
    lib.icd_device_info.argtypes = [C.POINTER(C.c_int)] * 3
    lib.icd_gemm.argtypes = [C.POINTER(IcdGemm), C.c_void_p]
    lib.icd_gemm_pick_bn.argtypes = [C.c_int] * 6
    lib.icd_attention.argtypes = [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_longlong] * 4 + [C.c_float, C.c_void_p,
                                                                                         C.c_longlong, C.c_void_p]
    lib.icd_attention_ex.argtypes = [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_longlong] * 4 + [C.c_float, C.c_void_p,
                                                                                            C.c_longlong, C.c_void_p,
                                                                                            C.c_void_p]
    lib.icd_groupnorm.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.icd_groupnorm_launches.argtypes = [C.c_int] * 3
    lib.icd_layernorm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                  C.c_void_p]
    lib.icd_softmax.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_longlong, C.c_void_p]
    lib.icd_softmax_causal.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_longlong, C.c_int, C.c_void_p]
    lib.icd_act.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]
    lib.icd_embed_tokens.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int,
                                     C.c_int, C.c_void_p]
    lib.icd_upsample2x.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.icd_im2col_s2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.icd_im2col_s2_pad.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.icd_latent_to_nhwc.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.icd_timestep_embedding.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.icd_guidance_embedding.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.icd_silu.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
    lib.icd_add.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
    lib.icd_consistency_update.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    P, I, L, F = C.c_void_p, C.c_int, C.c_longlong, C.c_float
    lib.icd_sgemm_f32.argtypes = [C.POINTER(IcdSgemm), P]
    lib.icd_groupnorm_f32.argtypes = [P, I, P, I, P, I, I, I, F, P, P, I, P]
    lib.icd_layernorm_f32.argtypes = [P, P, I, I, F, P, P, P]
    lib.icd_softmax_f32.argtypes = [P, L, I, L, P]
    lib.icd_silu_f32.argtypes = [P, P, L, P]
    lib.icd_geglu_f32.argtypes = [P, P, L, I, I, P]
    lib.icd_upsample2x_f32.argtypes = [P, P, I, I, I, I, P]
    lib.icd_im2col_s2_f32.argtypes = [P, P, I, I, I, I, P]
    lib.icd_nchw_to_nhwc_f32.argtypes = [P, P, I, I, I, I, P]
    lib.icd_nhwc_to_nchw_f32.argtypes = [P, L, P, I, I, I, P]
    lib.icd_sincos_embedding_f32.argtypes = [P, P, P, I, I, F, I, P]
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)
        if name != "icd_last_error":
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().icd_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libicd_b200 {what} failed: {msg}")
