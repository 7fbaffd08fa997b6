"""fp32 validation ops: the API of `ops.py` over the `icd_*_f32` kernels (include/icd_b200.h, ABI 3).

`B200UNet(..., precision='fp32')` — what `load_models(dtype='fp32')` builds, the dtype the reference runs SD1.5
editing in (running/sd1.5/launch_editing_iCD_sd1.5.sh:38) — runs the same executor code with this module in place of
`ops`: same packed layouts (q|k|v concat, all-layer K|V, GEGLU interleave, conv tap order), fp32 tensors throughout,
FMA-pipe contractions with fp32 accumulation. Attention always takes the explicit route of the reference's patched
forward (scores -> softmax -> [controller] -> P.V, utils/p2p.py:335-338). A correctness mode, not a fast path.
"""
import ctypes as C

import torch

from . import _lib
from . import ops as _ops16

DTYPE = torch.float32
GEGLU_BN = 256


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _f32(t, name):
    if t.dtype != torch.float32 or not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA fp32 tensor, got {t.dtype} on {t.device}")
    if t.device.index != torch.cuda.current_device():
        raise ValueError(f"{name}: tensor lives on {t.device} but the current device is cuda:"
                         f"{torch.cuda.current_device()}")


def _count(n=1):
    _ops16._count(n)


def sgemm_raw(**kw):
    g = _lib.IcdSgemm()
    g.alpha, g.Z, g.ZH = 1.0, 1, 1
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            v = v.data_ptr()
        setattr(g, k, v if v is not None else 0)
    _lib.check(_lib.load().icd_sgemm_f32(C.byref(g), _stream()), "icd_sgemm_f32")
    _count()
    _ops16._work("gemm", 2.0 * g.M * g.N * (g.C0 + g.C1) * (9 if g.conv else 1) * g.Z)


def geglu_act(x, F, out=None):
    M = x.shape[0]
    if out is None:
        out = torch.empty((M, F), device=x.device, dtype=DTYPE)
    _lib.check(_lib.load().icd_geglu_f32(_ptr(x), _ptr(out), M, F, GEGLU_BN, _stream()), "icd_geglu_f32")
    _count()
    return out


def linear(a, w, bias=None, out=None, residual=None, a1=None, rowvec=None, rows_per_img=0, geglu=False,
           out_fp32=True, alpha=1.0, force_bn=0, force_bm=0, force_splits=0):
    """out[M, N] = alpha * [a | a1] @ w.T (+ bias) (+ rowvec per image) (+ residual); geglu: h * gelu(g) of the
    interleaved projection. fp32 in, fp32 out (`out_fp32`, `force_*` accepted for signature parity with ops.linear)."""
    _f32(a, "a"); _f32(w, "w")
    M, K0 = a.shape
    K1 = a1.shape[1] if a1 is not None else 0
    N = w.shape[0]
    if geglu:
        proj = linear(a, w, bias=bias, a1=a1, alpha=alpha)
        return geglu_act(proj, N // 2, out=out)
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=DTYPE)
    sgemm_raw(a0=a, a1=a1, C0=K0, C1=K1, a0_ld=a.stride(0), a1_ld=a1.stride(0) if a1 is not None else 0, conv=0,
              b=w, b_ld=w.stride(0), b_kn=0, M=M, N=N, alpha=alpha, bias=bias, rowvec=rowvec,
              rows_per_img=rows_per_img if rowvec is not None else 0,
              ldv=rowvec.stride(0) if rowvec is not None else 0, residual=residual,
              ldr=residual.stride(0) if residual is not None else 0, out=out, ldc=out.stride(0))
    return out


def nhwc_to_nchw(x, B, Cc, HW, out):
    _lib.check(_lib.load().icd_nhwc_to_nchw_f32(_ptr(x), x.stride(0), _ptr(out), B, Cc, HW, _stream()),
               "icd_nhwc_to_nchw_f32")
    _count()
    return out


def conv3x3(x0, w, B, H, W, bias=None, x1=None, rowvec=None, residual=None, out=None, out_fp32=True,
            nchw_out=None, upd_x=None, upd_out=None, upd_coefs=None, force_bn=0, force_bm=0, force_splits=0):
    """3x3 / pad 1 / stride 1 convolution, implicit gather. w: [Cout, 9 * cin_pad] packed (ky, kx, cin) as for the
    fp16 path. `nchw_out` / `upd_*`: transposed fp32 store and the consistency update (separate small kernels here)."""
    _f32(x0, "x0"); _f32(w, "w")
    M = B * H * W
    C0 = x0.shape[1]
    C1 = x1.shape[1] if x1 is not None else 0
    N = w.shape[0]
    res = torch.empty((M, N), device=x0.device, dtype=DTYPE) if (out is None or nchw_out is not None) else out
    sgemm_raw(a0=x0, a1=x1, C0=C0, C1=C1, a0_ld=x0.stride(0), a1_ld=x1.stride(0) if x1 is not None else 0, conv=1,
              B=B, H=H, W=W, w_tap_ld=w.shape[1] // 9, b=w, b_ld=w.stride(0), b_kn=0, M=M, N=N, bias=bias,
              rowvec=rowvec, rows_per_img=H * W if rowvec is not None else 0,
              ldv=rowvec.stride(0) if rowvec is not None else 0, residual=residual,
              ldr=residual.stride(0) if residual is not None else 0, out=res, ldc=res.stride(0))
    if nchw_out is None:
        return res
    nhwc_to_nchw(res, B, N, H * W, nchw_out)
    if upd_x is not None:
        a_t, s_t, a_s, s_s = upd_coefs
        full = lambda v: torch.full((B,), float(v), device=x0.device, dtype=DTYPE)
        _ops16.consistency_update(nchw_out, upd_x, full(a_t), full(s_t), full(a_s), full(s_s), out=upd_out)
    return nchw_out


def attn_scores(q, k, B, H, Nq, Nk, D, scale, out):
    """out[b*H+h, q, :Nk] = scale * Q.K^T; out: fp32 [B*H, Nq, ld]."""
    sgemm_raw(a0=q, C0=D, a0_ld=q.stride(0), conv=0, b=k, b_ld=k.stride(0), b_kn=0, M=Nq, N=Nk, Z=B * H, ZH=H,
              a_zb=Nq * q.stride(0), a_zh=D, b_zb=Nk * k.stride(0), b_zh=D, c_zb=H * out.stride(0),
              c_zh=out.stride(0), alpha=scale, out=out, ldc=out.stride(1))
    return out


def softmax_(x, cols, causal_period=0):
    if causal_period:
        raise NotImplementedError("causal softmax is only needed by the fp16 CLIP text encoder")
    ld = x.stride(-2)
    rows = x.numel() // x.shape[-1]
    _lib.check(_lib.load().icd_softmax_f32(_ptr(x), rows, cols, ld, _stream()), "icd_softmax_f32")
    _count()
    return x


def attn_pv(probs, v, B, H, Nq, Nk, D, out):
    """out[b, q, h*D:(h+1)*D] = probs[b*H+h, q, :Nk] @ V[b, :, h, :]."""
    sgemm_raw(a0=probs, C0=Nk, a0_ld=probs.stride(1), conv=0, b=v, b_ld=v.stride(0), b_kn=1, M=Nq, N=D, Z=B * H,
              ZH=H, a_zb=H * probs.stride(0), a_zh=probs.stride(0), b_zb=Nk * v.stride(0), b_zh=D,
              c_zb=Nq * out.stride(0), c_zh=D, out=out, ldc=out.stride(0))
    return out


def attention(q, k, v, B, H, Nq, Nk, D, scale, out=None, probs_out=None, stats_out=None):
    """softmax(scale * Q.K^T) . V with the probabilities materialised in fp32 (one batch chunk at a time so that the
    transient stays below ~2 GB). `probs_out`: fp32 [B*H, Nq, >= Nk] receives them."""
    _f32(q, "q"); _f32(k, "k"); _f32(v, "v")
    if stats_out is not None:
        raise NotImplementedError("softmax statistics export is an fp16-path optimisation")
    if out is None:
        out = torch.empty((B * Nq, H * D), device=q.device, dtype=DTYPE)
    if probs_out is not None:
        attn_scores(q, k, B, H, Nq, Nk, D, scale, probs_out)
        softmax_(probs_out, Nk)
        return attn_pv(probs_out, v, B, H, Nq, Nk, D, out)
    ld = (Nk + 3) // 4 * 4
    per_b = H * Nq * ld * 4
    step = max(1, min(B, (2 << 30) // per_b))
    for b0 in range(0, B, step):
        nb = min(step, B - b0)
        probs = torch.empty((nb * H, Nq, ld), device=q.device, dtype=DTYPE)
        attn_scores(q[b0 * Nq:(b0 + nb) * Nq], k[b0 * Nk:(b0 + nb) * Nk], nb, H, Nq, Nk, D, scale, probs)
        softmax_(probs, Nk)
        attn_pv(probs, v[b0 * Nk:(b0 + nb) * Nk], nb, H, Nq, Nk, D, out[b0 * Nq:(b0 + nb) * Nq])
    _ops16._work("attention", 4.0 * B * H * Nq * Nk * D)
    return out


def groupnorm(x0, B, HW, gamma, beta, eps, silu, ws, x1=None, out=None, groups=32):
    _f32(x0, "x0")
    C0 = x0.shape[1]
    C1 = x1.shape[1] if x1 is not None else 0
    if not x0.is_contiguous() or (x1 is not None and not x1.is_contiguous()):
        raise ValueError("groupnorm (fp32): contiguous NHWC inputs expected")
    if out is None:
        out = torch.empty((B * HW, C0 + C1), device=x0.device, dtype=DTYPE)
    _lib.check(_lib.load().icd_groupnorm_f32(_ptr(x0), C0, _ptr(x1), C1, _ptr(out), B, HW, groups, float(eps),
                                             _ptr(gamma), _ptr(beta), int(silu), _stream()), "icd_groupnorm_f32")
    _count()
    return out


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    _f32(x, "x")
    rows, Cc = x.shape
    if out is None:
        out = torch.empty_like(x)
    _lib.check(_lib.load().icd_layernorm_f32(_ptr(x), _ptr(out), rows, Cc, float(eps), _ptr(gamma), _ptr(beta),
                                             _stream()), "icd_layernorm_f32")
    _count()
    return out


def upsample2x(x, B, H, W, out=None):
    Cc = x.shape[1]
    if out is None:
        out = torch.empty((B * 4 * H * W, Cc), device=x.device, dtype=DTYPE)
    _lib.check(_lib.load().icd_upsample2x_f32(_ptr(x), _ptr(out), B, H, W, Cc, _stream()), "icd_upsample2x_f32")
    _count()
    return out


def im2col_s2(x, B, H, W, out=None, pad=1):
    if pad != 1:
        raise NotImplementedError("fp32 im2col: U-Net Downsample2D (pad 1) only")
    Cc = x.shape[1]
    if out is None:
        out = torch.empty((B * (H // 2) * (W // 2), 9 * Cc), device=x.device, dtype=DTYPE)
    _lib.check(_lib.load().icd_im2col_s2_f32(_ptr(x), _ptr(out), B, H, W, Cc, _stream()), "icd_im2col_s2_f32")
    _count()
    return out


def latent_to_nhwc(x, cpad=8, out=None):
    B, Cc, H, W = x.shape
    if x.dtype != torch.float32:
        raise ValueError("latent_to_nhwc expects fp32 NCHW")
    if out is None:
        out = torch.empty((B * H * W, cpad), device=x.device, dtype=DTYPE)
    _lib.check(_lib.load().icd_nchw_to_nhwc_f32(_ptr(x), _ptr(out), B, Cc, H * W, cpad, _stream()),
               "icd_nchw_to_nhwc_f32")
    _count()
    return out


def _sincos(v, freqs, dim, scale, sin_first, out):
    n = v.numel()
    if out is None:
        out = torch.empty((n, dim), device=v.device, dtype=DTYPE)
    _lib.check(_lib.load().icd_sincos_embedding_f32(_ptr(v), _ptr(freqs), _ptr(out), n, dim, float(scale),
                                                    int(sin_first), _stream()), "icd_sincos_embedding_f32")
    _count()
    return out


def timestep_embedding(t, freqs, dim, out=None):
    return _sincos(t, freqs, dim, 1.0, False, out)


def guidance_embedding(w, freqs, dim, out=None):
    return _sincos(w, freqs, dim, 1000.0, True, out)


def silu(x, out=None):
    if out is None:
        out = torch.empty_like(x)
    _lib.check(_lib.load().icd_silu_f32(_ptr(x), _ptr(out), x.numel(), _stream()), "icd_silu_f32")
    _count()
    return out
