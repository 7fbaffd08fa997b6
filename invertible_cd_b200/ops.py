"""Tensor-level wrappers over the C ABI (include/icd_b200.h).

PyTorch is used only for device memory and streams; every op below launches a hand-written sm_100a kernel
from libicd_b200.so on torch's current CUDA stream (so the calls can be captured into a CUDA graph).
Layout conventions: activations fp16 channels-last ([B, H*W, C] token matrices), weights fp16 [N, K].
"""
import ctypes as C

import torch

from . import _lib

launch_count = 0  # number of kernel launches issued through this module (bench.py reports it)
profile = None    # when a list: every tensor-core launch appends (kind, flops, start_event, end_event)
shape_log = None  # when a list: every gemm / attention launch appends a dict describing its problem
work = None       # when a dict: algorithmic work per kernel family ("gemm"/"attention": FLOP; "groupnorm"/"layernorm": bytes)


def _work(kind, amount):
    if work is not None:
        work[kind] = work.get(kind, 0.0) + amount


def _prof_begin():
    if profile is None:
        return None
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


def _prof_end(ev, kind, work):
    if ev is None:
        return
    end = torch.cuda.Event(enable_timing=True)
    end.record()
    profile.append((kind, work, ev, end))


def _stream():
    """torch's current stream of the CURRENT device; the callers (B200UNet.forward, graphs.run) make the tensors'
    device current first, and `_f16` rejects tensors that live elsewhere."""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


_debug_sync = __import__("os").environ.get("ICD_DEBUG_SYNC", "0") != "0"


def _count(n=1):
    global launch_count
    launch_count += n
    if _debug_sync and not torch.cuda.is_current_stream_capturing():
        torch.cuda.synchronize()      # debugging aid: surface an asynchronous kernel fault at the op that caused it


def _f16(t, name):
    if t.dtype != torch.float16 or not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA fp16 tensor, got {t.dtype} on {t.device}")
    if t.device.index != torch.cuda.current_device():
        raise ValueError(f"{name}: tensor lives on {t.device} but the current device is cuda:"
                         f"{torch.cuda.current_device()} (wrap the call in torch.cuda.device(t.device))")


_splitk_ws = {}
SPLITK_WS_BYTES = 64 << 20


def _workspace(device):
    """fp32 split-K scratch (one per device, stream-ordered reuse). Allocated on first use — i.e. during the eager
    warm-up, never inside a CUDA-graph capture."""
    key = str(device)
    if key not in _splitk_ws:
        _splitk_ws[key] = torch.empty(SPLITK_WS_BYTES // 4, device=device, dtype=torch.float32)
    return _splitk_ws[key]


def gemm_raw(**kw):
    """Fill an IcdGemm from keyword fields (tensors are converted to device pointers) and launch."""
    g = _lib.IcdGemm()
    g.alpha = 1.0
    if "ws" not in kw and isinstance(kw.get("out"), torch.Tensor):
        ws = _workspace(kw["out"].device)
        kw["ws"], kw["ws_bytes"] = ws, ws.numel() * 4
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            v = v.data_ptr()
        setattr(g, k, v if v is not None else 0)
    if shape_log is not None:
        shape_log.append({"kind": "gemm_tc", "M": g.M, "N": g.N, "K": g.K * (9 if g.a_mode == 1 else 1), "Z": g.Z,
                          "conv": g.a_mode, "geglu": g.geglu, "mn": g.b_mn_major,
                          "bn": pick_bn(g.M, g.N, g.Z, bool(g.geglu), bool(g.b_mn_major), g.force_bn),
                          "flops": 2.0 * g.M * g.N * g.K * (9 if g.a_mode == 1 else 1) * g.Z})
    ev = _prof_begin()
    _lib.check(_lib.load().icd_gemm(C.byref(g), _stream()), "icd_gemm")
    _count()
    _work("gemm", 2.0 * g.M * g.N * g.K * (9 if g.a_mode == 1 else 1) * g.Z)
    if ev is not None:
        k_total = g.K * (9 if g.a_mode == 1 else 1)
        _prof_end(ev, "gemm_tc", 2.0 * g.M * g.N * k_total * g.Z)


def pick_bn(M, N, Z=1, geglu=False, b_mn_major=False, force_bn=0):
    return _lib.load().icd_gemm_pick_bn(M, N, Z, int(geglu), int(b_mn_major), force_bn)


def linear(a, w, bias=None, out=None, residual=None, a1=None, rowvec=None, rows_per_img=0, geglu=False,
           out_fp32=False, alpha=1.0, force_bn=0, force_bm=0, force_splits=0):
    """out[M, N] = epilogue(alpha * [a | a1] @ w.T).  a: [M, K0] (row stride allowed), w: [N, K0+K1]."""
    _f16(a, "a"); _f16(w, "w")
    M, K0 = a.shape
    K1 = a1.shape[1] if a1 is not None else 0
    N = w.shape[0]
    n_out = N // 2 if geglu else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=torch.float32 if out_fp32 else torch.float16)
    gemm_raw(a0=a, a1=a1, a_mode=0, K0=K0, K1=K1, a0_ld=a.stride(0), a1_ld=a1.stride(0) if a1 is not None else 0,
             ZA1=1, b=w, b_ld=w.stride(0), ZB1=1, M=M, N=N, K=K0 + K1, Z=1, alpha=alpha, bias=bias, rowvec=rowvec,
             rows_per_img=rows_per_img, ldv=rowvec.stride(0) if rowvec is not None else 0, residual=residual,
             ldr=residual.stride(0) if residual is not None else 0, out=out, ldc=out.stride(0),
             out_fp32=int(out.dtype == torch.float32), out_mode=0, geglu=int(geglu), force_bn=force_bn,
             force_bm=force_bm, force_splits=force_splits)
    return out


def conv3x3(x0, w, B, H, W, bias=None, x1=None, rowvec=None, residual=None, out=None, out_fp32=False,
            nchw_out=None, upd_x=None, upd_out=None, upd_coefs=None, force_bn=0, force_bm=0, force_splits=0):
    """3x3 / pad 1 / stride 1 convolution as implicit GEMM.
    x0: [B*H*W, C0] (NHWC), optional x1: [B*H*W, C1] concatenated along channels; w: [Cout, 9*(C0+C1)] packed
    (ky, kx, cin). `nchw_out`: fp32 [B, Cout, H, W] transposed store (conv_out); with upd_* the consistency
    update is fused into the same epilogue."""
    _f16(x0, "x0"); _f16(w, "w")
    M = B * H * W
    C0 = x0.shape[1]
    C1 = x1.shape[1] if x1 is not None else 0
    N = w.shape[0]
    kw = dict(a0=x0, a1=x1, a_mode=1, K0=C0, K1=C1, a0_ld=x0.stride(0), a1_ld=x1.stride(0) if x1 is not None else 0,
              B=B, H=H, W=W, b=w, b_ld=w.stride(0), ZB1=1, M=M, N=N, K=C0 + C1, Z=1, alpha=1.0, bias=bias,
              rowvec=rowvec, rows_per_img=H * W, ldv=rowvec.stride(0) if rowvec is not None else 0,
              residual=residual, ldr=residual.stride(0) if residual is not None else 0, force_bn=force_bn,
              force_bm=force_bm, force_splits=force_splits)
    if nchw_out is not None:
        kw.update(out=nchw_out, ldc=H * W, out_imgstride=N * H * W, out_fp32=1, out_mode=1)
        if upd_x is not None:
            a_t, s_t, a_s, s_s = upd_coefs
            kw.update(upd_x=upd_x, upd_out=upd_out, alpha_t=a_t, sigma_t=s_t, alpha_s=a_s, sigma_s=s_s)
        gemm_raw(**kw)
        return nchw_out
    if out is None:
        out = torch.empty((M, N), device=x0.device, dtype=torch.float32 if out_fp32 else torch.float16)
    kw.update(out=out, ldc=out.stride(0), out_fp32=int(out.dtype == torch.float32), out_mode=0)
    gemm_raw(**kw)
    return out


def attn_scores(q, k, B, H, Nq, Nk, D, scale, out):
    """Explicit path, step 1 (utils/p2p.py:335 get_attention_scores): out[b*H+h, q, :Nk] = scale * Q.K^T (fp16).
    q: [B*Nq, H*D], k: [B*Nk, H*D]; out: [B*H, Nq, ldp] with ldp >= Nk, ldp % 8 == 0."""
    gemm_raw(a0=q, a_mode=0, K0=D, a0_ld=q.stride(0), a_z1_stride=D, a_z2_stride=Nq * q.stride(0), ZA1=H,
             b=k, b_ld=k.stride(0), b_z1_stride=D, b_z2_stride=Nk * k.stride(0), ZB1=H, M=Nq, N=Nk, K=D, Z=B * H,
             alpha=scale, out=out, ldc=out.stride(1), out_z1_stride=out.stride(0), out_z2_stride=H * out.stride(0), out_fp32=0,
             out_mode=0)
    return out


def attn_pv(probs, v, B, H, Nq, Nk, D, out):
    """Explicit path, step 3 (utils/p2p.py:338 torch.bmm(attention_probs, value)):
    out[b, q, h*D:(h+1)*D] = probs[b*H+h, q, :Nk] @ V[b, :, h, :].  V is consumed MN-major (no transpose)."""
    gemm_raw(a0=probs, a_mode=0, K0=Nk, a0_ld=probs.stride(1), a_z1_stride=probs.stride(0),
             a_z2_stride=H * probs.stride(0), ZA1=H,
             b=v, b_ld=v.stride(0), b_z1_stride=D, b_z2_stride=Nk * v.stride(0), ZB1=H, b_mn_major=1,
             M=Nq, N=D, K=Nk, Z=B * H, alpha=1.0, out=out, ldc=out.stride(0), out_z1_stride=D, out_z2_stride=Nq * out.stride(0),
             out_fp32=0, out_mode=0,
             force_bn=64 if D <= 64 else (128 if D <= 128 else 256))
    return out


def attn_probs_from_stats(q, k, B, H, Nq, Nk, D, scale, stats, out):
    """Normalised attention probabilities in one pass over the scores: out[b*H+h, q, :] =
    exp2(scale*log2(e) * Q.K^T - stats[.., 0]) * stats[.., 1] with the online-softmax statistics `attention(...,
    stats_out=stats)` produced (AttentionStore capture of self-attention maps, utils/p2p.py:145-149, without a separate
    softmax pass). out: fp16 [B*H, Nq, Nk], Nk % 8 == 0."""
    gemm_raw(a0=q, a_mode=0, K0=D, a0_ld=q.stride(0), a_z1_stride=D, a_z2_stride=Nq * q.stride(0), ZA1=H,
             b=k, b_ld=k.stride(0), b_z1_stride=D, b_z2_stride=Nk * k.stride(0), ZB1=H, M=Nq, N=Nk, K=D, Z=B * H,
             alpha=scale * 1.4426950408889634, out=out, ldc=out.stride(1), out_z1_stride=out.stride(0),
             out_z2_stride=H * out.stride(0), out_fp32=0, out_mode=0, exp_stats=stats)
    return out


def attention(q, k, v, B, H, Nq, Nk, D, scale, out=None, probs_out=None, stats_out=None):
    """Fused attention core. q: [B*Nq, H*D]; k, v: [B*Nk, H*D] (row strides allowed); out: [B*Nq, H*D].
    `stats_out`: optional fp32 [B*H, Nq, 2] receiving (reference max * scale * log2 e, 1 / row sum) per query row."""
    _f16(q, "q"); _f16(k, "k"); _f16(v, "v")
    if out is None:
        out = torch.empty((B * Nq, H * D), device=q.device, dtype=torch.float16)
    if shape_log is not None:
        shape_log.append({"kind": "attention_tc", "B": B, "H": H, "Nq": Nq, "Nk": Nk, "D": D,
                          "probs": probs_out is not None, "flops": 4.0 * B * H * Nq * Nk * D})
    ev = _prof_begin()
    _lib.check(_lib.load().icd_attention_ex(_ptr(q), _ptr(k), _ptr(v), _ptr(out), B, H, Nq, Nk, D, q.stride(0),
                                            k.stride(0), v.stride(0), out.stride(0), float(scale), _ptr(probs_out),
                                            probs_out.stride(1) if probs_out is not None else 0, _ptr(stats_out),
                                            _stream()),
               "icd_attention")
    _count()
    _work("attention", 4.0 * B * H * Nq * Nk * D)
    _prof_end(ev, "attention_tc", 4.0 * B * H * Nq * Nk * D)
    return out


_gn_plan_cache = {}


def _gn_launches(B, HW, Cc):
    key = (B, HW, Cc)
    if key not in _gn_plan_cache:
        _gn_plan_cache[key] = int(_lib.load().icd_groupnorm_launches(B, HW, Cc))
    return _gn_plan_cache[key]


def groupnorm(x0, B, HW, gamma, beta, eps, silu, ws, x1=None, out=None, groups=32):
    _f16(x0, "x0")
    C0 = x0.shape[1]
    C1 = x1.shape[1] if x1 is not None else 0
    if out is None:
        out = torch.empty((B * HW, C0 + C1), device=x0.device, dtype=torch.float16)
    if ws.numel() < B * 64 * 64:
        raise ValueError("groupnorm workspace too small (need B*4096 floats)")
    ev = _prof_begin()
    _lib.check(_lib.load().icd_groupnorm(_ptr(x0), C0, _ptr(x1), C1, _ptr(out), B, HW, groups, float(eps),
                                         _ptr(gamma), _ptr(beta), int(silu), _ptr(ws), _stream()), "icd_groupnorm")
    _count(_gn_launches(B, HW, C0 + C1))
    _work("groupnorm", 2.0 * B * HW * (C0 + C1) * 2)
    _prof_end(ev, "groupnorm", 2.0 * B * HW * (C0 + C1) * 2)   # algorithmic bytes: read once + write once
    return out


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    _f16(x, "x")
    rows, Cc = x.shape
    if out is None:
        out = torch.empty_like(x)
    _lib.check(_lib.load().icd_layernorm(_ptr(x), _ptr(out), rows, Cc, float(eps), _ptr(gamma), _ptr(beta),
                                         _stream()), "icd_layernorm")
    _count()
    _work("layernorm", 4.0 * rows * Cc)
    return out


def softmax_(x, cols, causal_period=0):
    """In-place softmax over the first `cols` entries of the last dim of a [.., rows, ld] fp16 tensor.
    `causal_period` T > 0: row r only sees keys 0..(r % T) (the rest of the row becomes 0)."""
    ld = x.stride(-2)
    rows = x.numel() // x.shape[-1]
    _lib.check(_lib.load().icd_softmax_causal(_ptr(x), rows, cols, ld, causal_period, _stream()), "icd_softmax")
    _count()
    return x


def upsample2x(x, B, H, W, out=None):
    Cc = x.shape[1]
    if out is None:
        out = torch.empty((B * 4 * H * W, Cc), device=x.device, dtype=torch.float16)
    _lib.check(_lib.load().icd_upsample2x(_ptr(x), _ptr(out), B, H, W, Cc, _stream()), "icd_upsample2x")
    _count()
    return out


def im2col_s2(x, B, H, W, out=None, pad=1):
    """Operand of a stride-2 3x3 conv. pad=1: U-Net Downsample2D; pad=0: VAE encoder Downsample2D (zeros bottom/right)."""
    Cc = x.shape[1]
    if out is None:
        out = torch.empty((B * (H // 2) * (W // 2), 9 * Cc), device=x.device, dtype=torch.float16)
    _lib.check(_lib.load().icd_im2col_s2_pad(_ptr(x), _ptr(out), B, H, W, Cc, pad, _stream()), "icd_im2col_s2")
    _count()
    return out


def latent_to_nhwc(x, cpad=8, out=None):
    B, Cc, H, W = x.shape
    if x.dtype != torch.float32:
        raise ValueError("latent_to_nhwc expects fp32 NCHW")
    if out is None:
        out = torch.empty((B * H * W, cpad), device=x.device, dtype=torch.float16)
    _lib.check(_lib.load().icd_latent_to_nhwc(_ptr(x), _ptr(out), B, Cc, H * W, cpad, _stream()),
               "icd_latent_to_nhwc")
    _count()
    return out


def timestep_embedding(t, freqs, dim, out=None):
    n = t.numel()
    if out is None:
        out = torch.empty((n, dim), device=t.device, dtype=torch.float16)
    _lib.check(_lib.load().icd_timestep_embedding(_ptr(t), _ptr(freqs), _ptr(out), n, dim, _stream()),
               "icd_timestep_embedding")
    _count()
    return out


def guidance_embedding(w, freqs, dim, out=None):
    n = w.numel()
    if out is None:
        out = torch.empty((n, dim), device=w.device, dtype=torch.float16)
    _lib.check(_lib.load().icd_guidance_embedding(_ptr(w), _ptr(freqs), _ptr(out), n, dim, _stream()),
               "icd_guidance_embedding")
    _count()
    return out


def silu(x, out=None):
    if out is None:
        out = torch.empty_like(x)
    _lib.check(_lib.load().icd_silu(_ptr(x), _ptr(out), x.numel(), _stream()), "icd_silu")
    _count()
    return out


ACT_SILU, ACT_QUICK_GELU, ACT_GELU = 0, 1, 2


def act(x, kind, out=None):
    if out is None:
        out = torch.empty_like(x)
    _lib.check(_lib.load().icd_act(_ptr(x), _ptr(out), x.numel(), kind, _stream()), "icd_act")
    _count()
    return out


def embed_tokens(ids, tok, pos, out=None):
    """ids: int64 [B, T] (device) -> fp16 [B*T, C] = tok[ids] + pos[t]."""
    B, T = ids.shape
    Cc = tok.shape[1]
    if out is None:
        out = torch.empty((B * T, Cc), device=tok.device, dtype=torch.float16)
    _lib.check(_lib.load().icd_embed_tokens(_ptr(ids), _ptr(tok), _ptr(pos), _ptr(out), B * T, T, Cc, tok.shape[0],
                                            _stream()), "icd_embed_tokens")
    _count()
    return out


def add(a, b, out=None):
    if out is None:
        out = torch.empty_like(a)
    _lib.check(_lib.load().icd_add(_ptr(a), _ptr(b), _ptr(out), a.numel(), _stream()), "icd_add")
    _count()
    return out


def consistency_update(eps, x, alpha_t, sigma_t, alpha_s, sigma_s, out=None):
    """predicted_origin (utils/generation.py:136-155); eps/x fp32 [B, ...]; coefficient tensors fp32 [B]."""
    B = x.shape[0]
    if out is None:
        out = torch.empty_like(x)
    _lib.check(_lib.load().icd_consistency_update(_ptr(eps), _ptr(x), _ptr(out), x.numel() // B, B, _ptr(alpha_t),
                                                  _ptr(sigma_t), _ptr(alpha_s), _ptr(sigma_s), _stream()),
               "icd_consistency_update")
    _count()
    return out
