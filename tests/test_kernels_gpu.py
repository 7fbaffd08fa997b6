"""Kernel-level parity (GPU): every sm_100a kernel behind the C ABI vs a plain PyTorch fp32 reference of the same
op on identical fp16 inputs. Tolerances (stated per test): outputs are fp16, accumulation is fp32, so the bar is
rtol=1e-3 plus an absolute term of ~1 fp16 ulp of the output scale (atol given per case)."""
import math

import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from invertible_cd_b200 import ops as _ops
    return _ops


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to("cuda", torch.float16)


def _close(got, ref, rtol=1e-3, atol=2e-3, what=""):
    got = got.float()
    ref = ref.float().to(got.device)
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = (err > tol).sum().item()
    assert bad == 0, f"{what}: {bad}/{err.numel()} mismatches, max err {err.max().item():.4e}, ref max {ref.abs().max().item():.3e}"


# --------------------------------------------------------------------------- GEMM / linear
@pytest.mark.parametrize("M,K,N,bn", [(300, 320, 320, 0), (128, 64, 64, 64), (1000, 1280, 640, 128),
                                      (257, 320, 320, 160), (4096, 640, 1280, 256), (8, 1280, 2560, 0),
                                      (513, 328, 72, 0)])
def test_linear_bias_residual(ops, M, K, N, bn):
    a, w = _rand(M, K, seed=1), _rand(N, K, scale=K ** -0.5, seed=2)
    bias = torch.randn(N, device="cuda")
    res = _rand(M, N, seed=3)
    out = ops.linear(a, w, bias=bias, residual=res, force_bn=bn)
    ref = a.float() @ w.float().t() + bias + res.float()
    torch.cuda.synchronize()
    _close(out, ref, what=f"linear {M}x{K}x{N} bn={bn}")


@pytest.mark.parametrize("M,K,N,bn", [(1000, 1280, 640, 128), (4096, 640, 1280, 256), (700, 320, 320, 160),
                                      (8192, 2560, 640, 0), (300, 192, 64, 64)])
def test_linear_256row_tiles(ops, M, K, N, bn):
    """BM=256 variant: two 128-row MMA halves share each B tile (halves the L2->SM operand traffic)."""
    a, w = _rand(M, K, seed=1), _rand(N, K, scale=K ** -0.5, seed=2)
    bias = torch.randn(N, device="cuda")
    res = _rand(M, N, seed=3)
    out = ops.linear(a, w, bias=bias, residual=res, force_bn=bn, force_bm=256)
    _close(out, a.float() @ w.float().t() + bias + res.float(), what=f"linear bm256 {M}x{K}x{N} bn={bn}")
    out32 = ops.linear(a, w, bias=bias, out_fp32=True, force_bn=bn, force_bm=256)     # direct epilogue
    _close(out32, a.float() @ w.float().t() + bias, rtol=1e-4, atol=1e-3, what="bm256 fp32")


@pytest.mark.parametrize("M,K,N,geglu,bias", [(2048, 1280, 2560, True, True), (4096, 1024, 5120, True, False),
                                              (4096, 1280, 3840, False, False), (4096, 1024, 4096, False, True),
                                              (8192, 2048, 3840, False, True)])
def test_linear_two_sm_tiles(ops, M, K, N, geglu, bias):
    """Shapes icd_gemm routes to the 2-SM kernel (one 256x256 tile per CTA pair, tcgen05.mma.cta_group::2): GEGLU
    projections with K >= 1024 and fat-N plain projections. Several tiles per pair, both accumulator buffers."""
    from invertible_cd_b200.packing import pack_geglu
    a, w = _rand(M, K, seed=11), _rand(N, K, scale=K ** -0.5, seed=12)
    b = torch.randn(N, device="cuda") if bias else None
    y = a.float() @ w.float().t() + (b if bias else 0)
    if geglu:
        wp, bp = pack_geglu(w, b if bias else torch.zeros(N, device="cuda"), 256)
        out = ops.linear(a, wp, bias=bp if bias else None, geglu=True, force_bn=256)
        ref = y[:, :N // 2] * F.gelu(y[:, N // 2:])
    else:
        out = ops.linear(a, w, bias=b)
        ref = y
    _close(out, ref, rtol=2e-3, atol=4e-3, what=f"2-SM {M}x{K}x{N} geglu={geglu}")


@pytest.mark.parametrize("splits,bm", [(3, 0), (4, 256), (0, 0)])
def test_split_k(ops, splits, bm):
    """Few-tile, deep-K problems (the 8x8-level convs): K split over CTAs, fp32 partials + fused reduce epilogue."""
    from invertible_cd_b200.packing import pack_conv3x3
    B, H, W, Cin, Cout = 8, 8, 8, 320, 1280
    x = _rand(B * H * W, Cin, seed=70)
    w = _rand(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=71)
    bias, rv = torch.randn(Cout, device="cuda"), torch.randn(B, Cout, device="cuda")
    res = _rand(B * H * W, Cout, seed=72)
    out = ops.conv3x3(x, pack_conv3x3(w), B, H, W, bias=bias, rowvec=rv, residual=res, force_splits=splits,
                      force_bm=bm)
    ref = _conv_ref(x, w, B, H, W, bias) + rv.repeat_interleave(H * W, 0) + res.float()
    _close(out, ref, what=f"split-K conv splits={splits}")
    a, wl = _rand(300, 4096, seed=73), _rand(320, 4096, scale=1 / 64, seed=74)
    o2 = ops.linear(a, wl, bias=torch.ones(320, device="cuda"), force_splits=splits or 2, alpha=0.5)
    _close(o2, 0.5 * (a.float() @ wl.float().t()) + 1.0, what="split-K linear")


def test_linear_fp32_out_and_alpha(ops):
    a, w = _rand(200, 192, seed=4), _rand(96, 192, scale=0.1, seed=5)
    out = ops.linear(a, w, out_fp32=True, alpha=0.25)
    ref = 0.25 * (a.float() @ w.float().t())
    _close(out, ref, rtol=1e-4, atol=1e-4, what="fp32 out")


def test_linear_virtual_concat(ops):
    a0, a1 = _rand(700, 128, seed=6), _rand(700, 192, seed=7)
    w = _rand(320, 320, scale=320 ** -0.5, seed=8)
    out = ops.linear(a0, w, a1=a1)
    ref = torch.cat([a0, a1], 1).float() @ w.float().t()
    _close(out, ref, what="concat linear")


def test_linear_strided_views(ops):
    big = _rand(400, 512, seed=9)
    a = big[:, 128:384]                      # row stride 512, K = 256
    w = _rand(160, 256, scale=1 / 16, seed=10)
    outbuf = torch.zeros(400, 512, device="cuda", dtype=torch.float16)
    ops.linear(a, w, out=outbuf[:, 64:224])
    _close(outbuf[:, 64:224], a.float() @ w.float().t(), what="strided")
    assert outbuf[:, :64].abs().max() == 0 and outbuf[:, 224:].abs().max() == 0


@pytest.mark.parametrize("bn", [128, 256])
def test_linear_geglu(ops, bn):
    from invertible_cd_b200.packing import pack_geglu
    M, K, Fh = 333, 320, 1280
    a = _rand(M, K, seed=11)
    w = _rand(2 * Fh, K, scale=K ** -0.5, seed=12)
    b = torch.randn(2 * Fh, device="cuda") * 0.1
    wp, bp = pack_geglu(w, b, bn)
    out = ops.linear(a, wp, bias=bp, geglu=True, force_bn=bn)
    y = a.float() @ w.float().t() + b
    ref = y[:, :Fh] * F.gelu(y[:, Fh:])
    assert out.shape == (M, Fh)
    _close(out, ref, what=f"geglu bn={bn}")


def test_linear_rowvec(ops):
    B, HW, K, N = 3, 64, 128, 192
    a, w = _rand(B * HW, K, seed=13), _rand(N, K, scale=K ** -0.5, seed=14)
    rv = torch.randn(B, N, device="cuda")
    out = ops.linear(a, w, rowvec=rv, rows_per_img=HW)
    ref = a.float() @ w.float().t() + rv.repeat_interleave(HW, 0)
    _close(out, ref, what="rowvec")


# --------------------------------------------------------------------------- implicit-GEMM 3x3 conv
def _conv_ref(x_nhwc, w, B, H, W, bias=None):
    x = x_nhwc.float().reshape(B, H, W, -1).permute(0, 3, 1, 2)
    y = F.conv2d(x, w.float(), bias, padding=1)
    return y.permute(0, 2, 3, 1).reshape(B * H * W, -1)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 16, 16, 64, 128), (3, 8, 8, 128, 64), (1, 32, 32, 320, 320),
                                            (2, 64, 64, 64, 64), (1, 128, 128, 64, 32), (5, 4, 4, 64, 96)])
def test_conv3x3(ops, B, H, W, Cin, Cout):
    from invertible_cd_b200.packing import pack_conv3x3
    x = _rand(B * H * W, Cin, seed=20)
    w = _rand(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=21)
    bias = torch.randn(Cout, device="cuda")
    out = ops.conv3x3(x, pack_conv3x3(w), B, H, W, bias=bias)
    _close(out, _conv_ref(x, w, B, H, W, bias), what=f"conv {B}x{H}x{W} {Cin}->{Cout}")


@pytest.mark.parametrize("B,H,W,Cin,Cout,bn", [(2, 32, 32, 128, 320, 160), (3, 8, 8, 128, 256, 256),
                                               (1, 64, 64, 64, 128, 128), (5, 4, 4, 64, 96, 0)])
def test_conv3x3_256row_tiles(ops, B, H, W, Cin, Cout, bn):
    from invertible_cd_b200.packing import pack_conv3x3
    x = _rand(B * H * W, Cin, seed=20)
    w = _rand(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=21)
    bias = torch.randn(Cout, device="cuda")
    res = _rand(B * H * W, Cout, seed=22)
    out = ops.conv3x3(x, pack_conv3x3(w), B, H, W, bias=bias, residual=res, force_bn=bn, force_bm=256)
    _close(out, _conv_ref(x, w, B, H, W, bias) + res.float(), what=f"conv bm256 {B}x{H}x{W} {Cin}->{Cout}")


def test_conv3x3_concat_rowvec_residual(ops):
    from invertible_cd_b200.packing import pack_conv3x3
    B, H, W, C0, C1, Cout = 2, 16, 16, 128, 64, 160
    x0, x1 = _rand(B * H * W, C0, seed=22), _rand(B * H * W, C1, seed=23)
    w = _rand(Cout, C0 + C1, 3, 3, scale=(9 * (C0 + C1)) ** -0.5, seed=24)
    bias = torch.randn(Cout, device="cuda")
    rv = torch.randn(B, Cout, device="cuda")
    res = _rand(B * H * W, Cout, seed=25)
    out = ops.conv3x3(x0, pack_conv3x3(w), B, H, W, bias=bias, x1=x1, rowvec=rv, residual=res)
    ref = _conv_ref(torch.cat([x0, x1], 1), w, B, H, W, bias) + rv.repeat_interleave(H * W, 0) + res.float()
    _close(out, ref, what="conv concat")


def test_conv_in_padded_channels(ops):
    from invertible_cd_b200.packing import pack_conv3x3
    B, H, W = 2, 64, 64
    lat = torch.randn(B, 4, H, W, device="cuda")
    x = ops.latent_to_nhwc(lat, cpad=8)
    assert x.shape == (B * H * W, 8)
    torch.testing.assert_close(x[:, :4].float().reshape(B, H, W, 4).permute(0, 3, 1, 2), lat.half().float())
    assert x[:, 4:].abs().max() == 0
    w = _rand(320, 4, 3, 3, scale=1 / 6, seed=26)
    bias = torch.randn(320, device="cuda")
    out = ops.conv3x3(x, pack_conv3x3(w), B, H, W, bias=bias)
    ref = F.conv2d(lat.half().float(), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(B * H * W, 320)
    _close(out, ref, what="conv_in")


def test_conv_out_nchw_fused_update(ops):
    from invertible_cd_b200.packing import pack_conv3x3
    B, H, W, Cin = 2, 64, 64, 320
    x = _rand(B * H * W, Cin, seed=27)
    w = _rand(4, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=28)
    bias = torch.randn(4, device="cuda")
    x_t = torch.randn(B, 4, H, W, device="cuda")
    eps_out = torch.empty(B, 4, H, W, device="cuda")
    x_s = torch.empty_like(x_t)
    coefs = (0.0682648867, 0.997667253, 0.210238948, 0.977650046)   # t=999 -> s=779
    ops.conv3x3(x, pack_conv3x3(w), B, H, W, bias=bias, nchw_out=eps_out, upd_x=x_t, upd_out=x_s, upd_coefs=coefs)
    ref_eps = F.conv2d(x.float().reshape(B, H, W, Cin).permute(0, 3, 1, 2), w.float(), bias, padding=1)
    _close(eps_out, ref_eps, rtol=1e-3, atol=1e-3, what="conv_out eps")
    a_t, s_t, a_s, s_s = coefs
    ref_x = a_s * ((x_t - s_t * eps_out) / a_t) + s_s * eps_out
    torch.testing.assert_close(x_s, ref_x, rtol=1e-5, atol=1e-5)


def test_downsample_conv_via_im2col(ops):
    from invertible_cd_b200.packing import pack_conv3x3
    B, H, W, Cc = 2, 32, 32, 64
    x = _rand(B * H * W, Cc, seed=29)
    w = _rand(128, Cc, 3, 3, scale=(9 * Cc) ** -0.5, seed=30)
    cols = ops.im2col_s2(x, B, H, W)
    out = ops.linear(cols, pack_conv3x3(w))
    ref = F.conv2d(x.float().reshape(B, H, W, Cc).permute(0, 3, 1, 2), w.float(), None, stride=2, padding=1)
    _close(out, ref.permute(0, 2, 3, 1).reshape(-1, 128), what="downsample")


def test_upsample2x(ops):
    B, H, W, Cc = 2, 8, 16, 64
    x = _rand(B * H * W, Cc, seed=31)
    y = ops.upsample2x(x, B, H, W)
    ref = F.interpolate(x.float().reshape(B, H, W, Cc).permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    assert torch.equal(y.float().reshape(B, 2 * H, 2 * W, Cc).permute(0, 3, 1, 2), ref)


# --------------------------------------------------------------------------- norms / softmax / elementwise
@pytest.mark.parametrize("B,HW,C0,C1,silu", [(2, 4096, 320, 0, True), (3, 256, 1280, 1280, True),
                                             (2, 1024, 640, 320, False), (1, 64, 2560, 0, True),
                                             (8, 1024, 640, 0, True),       # single-pass kernel, 4 CTAs per SM
                                             (8, 4096, 640, 320, True),     # too large to stay on chip: two kernels
                                             (5, 200, 320, 0, False)])      # ragged pixel chunks
def test_groupnorm(ops, B, HW, C0, C1, silu):
    x0 = _rand(B * HW, C0, seed=40) + 0.5
    x1 = _rand(B * HW, C1, scale=2.0, seed=41) if C1 else None
    Cc = C0 + C1
    gamma, beta = torch.randn(Cc, device="cuda"), torch.randn(Cc, device="cuda")
    ws = torch.empty(B * 4096, device="cuda")
    eps = 1e-5
    y = ops.groupnorm(x0, B, HW, gamma, beta, eps, silu, ws, x1=x1)
    x = x0 if x1 is None else torch.cat([x0, x1], 1)
    xr = x.float().reshape(B, HW, Cc).permute(0, 2, 1)
    ref = F.group_norm(xr, 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    _close(y, ref.permute(0, 2, 1).reshape(B * HW, Cc), what="groupnorm")
    # the single-pass kernel synchronises the CTAs of an image through counters that the last CTA resets:
    # back-to-back launches and a CUDA-graph replay must give the same bits
    y2 = ops.groupnorm(x0, B, HW, gamma, beta, eps, silu, ws, x1=x1)
    y3 = ops.groupnorm(x0, B, HW, gamma, beta, eps, silu, ws, x1=x1)
    assert torch.equal(y, y2) and torch.equal(y, y3)
    if os.environ.get("PYTORCH_NO_CUDA_MEMORY_CACHING"):
        return      # compute-sanitizer runs (tools/gpu_sanitizer.sh): no allocation is possible inside a capture
    out = torch.empty_like(y)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ops.groupnorm(x0, B, HW, gamma, beta, eps, silu, ws, x1=x1, out=out)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ops.groupnorm(x0, B, HW, gamma, beta, eps, silu, ws, x1=x1, out=out)
        ops.groupnorm(out, B, HW, gamma, beta, eps, silu, ws, out=torch.empty_like(out)) if C1 == 0 else None
    for _ in range(3):
        out.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, y)


@pytest.mark.parametrize("B,HW,C0,C1", [(8, 4096, 640, 320), (4, 16384, 320, 0), (2, 16384, 128, 0)])
def test_groupnorm_workspace_contract_two_kernel_path(ops, B, HW, C0, C1):
    """The two-kernel fallback (tensors too large for the single-pass kernel, or ICD_GN_FUSED=0) must stay inside the
    documented workspace of B * 4096 floats: canary words right behind it survive (round-1 ADVICE: the statistics
    kernel used up to 148 chunks per image against a 64-chunk workspace). (2, 16384, 128): 4 channels per group, the
    VAE's 128-channel level."""
    x0 = _rand(B * HW, C0, seed=50) + 0.25
    x1 = _rand(B * HW, C1, seed=51) if C1 else None
    Cc = C0 + C1
    gamma, beta = torch.randn(Cc, device="cuda"), torch.randn(Cc, device="cuda")
    buf = torch.full((B * 4096 + 65536,), 12345.0, device="cuda")
    y = ops.groupnorm(x0, B, HW, gamma, beta, 1e-5, True, buf[:B * 4096], x1=x1)
    torch.cuda.synchronize()
    assert bool((buf[B * 4096:] == 12345.0).all()), "GroupNorm wrote past its B*4096-float workspace"
    x = x0 if x1 is None else torch.cat([x0, x1], 1)
    ref = F.silu(F.group_norm(x.float().reshape(B, HW, Cc).permute(0, 2, 1), 32, gamma, beta, 1e-5))
    _close(y, ref.permute(0, 2, 1).reshape(B * HW, Cc), what="groupnorm-2k")


@pytest.mark.parametrize("rows,Cc", [(1000, 320), (77, 640), (4096, 1280)])
def test_layernorm(ops, rows, Cc):
    x = _rand(rows, Cc, seed=42) * 3 + 1
    gamma, beta = torch.randn(Cc, device="cuda"), torch.randn(Cc, device="cuda")
    y = ops.layernorm(x, gamma, beta, 1e-5)
    _close(y, F.layer_norm(x.float(), (Cc,), gamma, beta, 1e-5), what="layernorm")


def test_softmax_padded_rows(ops):
    x = _rand(16, 256, 80, seed=43) * 4
    ref = torch.softmax(x[..., :77].float(), -1)
    ops.softmax_(x, 77)
    _close(x[..., :77], ref, rtol=2e-3, atol=1e-4, what="softmax")


def test_embeddings_silu_add_update(ops):
    dev = "cuda"
    # Timesteps(320): exponent = -ln(10000) * arange(160) / 160  (diffusers get_timestep_embedding, freq_shift=0)
    half = 160
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half)
    t = torch.tensor([999.0, 259.0])
    emb = t[:, None] * freqs[None, :]
    ref = torch.cat([torch.cos(emb), torch.sin(emb)], -1)
    got = ops.timestep_embedding(t.to(dev), freqs.to(dev), 320)
    _close(got, ref, rtol=1e-3, atol=1e-3, what="timestep emb")
    # SURVEY Appendix B known answers
    assert abs(got[0, 0].item() - 0.9996498) < 1e-3 and abs(got[0, 160].item() - (-0.0264608)) < 1e-3
    # guidance_scale_embedding (utils/generation.py:96-122)
    w = torch.tensor([0.0, 7.0, 19.0])
    half = 256
    e = torch.log(torch.tensor(10000.0)) / (half - 1)
    fr = torch.exp(torch.arange(half, dtype=torch.float32) * -e)
    arg = (w * 1000.0)[:, None] * fr[None, :]
    refw = torch.cat([torch.sin(arg), torch.cos(arg)], 1)
    gotw = ops.guidance_embedding(w.to(dev), fr.to(dev), 512)
    _close(gotw, refw, rtol=1e-3, atol=1.5e-3, what="w emb")
    x = _rand(1000, seed=44) * 3
    _close(ops.silu(x), F.silu(x.float()), what="silu")
    y = _rand(1000, seed=45)
    _close(ops.add(x, y), x.float() + y.float(), what="add")
    B = 3
    epsn, xt = torch.randn(B, 4, 64, 64, device=dev), torch.randn(B, 4, 64, 64, device=dev)
    a_t = torch.tensor([0.0682648867, 0.210238948, 0.811772883], device=dev)
    s_t = torch.tensor([0.997667253, 0.977650046, 0.583973229], device=dev)
    a_s = torch.tensor([0.210238948, 0.501575053, 1.0], device=dev)
    s_s = torch.tensor([0.977650046, 0.865114152, 0.0], device=dev)
    got = ops.consistency_update(epsn, xt, a_t, s_t, a_s, s_s)
    v = lambda z: z.view(B, 1, 1, 1)
    ref = v(a_s) * ((xt - v(s_t) * epsn) / v(a_t)) + v(s_s) * epsn
    torch.testing.assert_close(got, ref, rtol=1e-6, atol=1e-6)


# --------------------------------------------------------------------------- attention
def _attn_ref(q, k, v, B, H, Nq, Nk, D, scale):
    qf = q.float().reshape(B, Nq, H, D).permute(0, 2, 1, 3)
    kf = k.float().reshape(B, Nk, H, D).permute(0, 2, 1, 3)
    vf = v.float().reshape(B, Nk, H, D).permute(0, 2, 1, 3)
    p = torch.softmax(qf @ kf.transpose(-1, -2) * scale, -1)
    o = (p @ vf).permute(0, 2, 1, 3).reshape(B * Nq, H * D)
    return o, p.reshape(B * H, Nq, Nk)


@pytest.mark.parametrize("B,H,Nq,Nk,D", [(2, 8, 256, 77, 40), (1, 8, 1024, 77, 80), (2, 5, 64, 77, 160),
                                         (2, 10, 320, 77, 64)])
def test_attention_explicit_path(ops, B, H, Nq, Nk, D):
    """scores GEMM (alpha=scale) -> in-place softmax -> P.V GEMM with MN-major V: the p2p-controller route."""
    q, k, v = _rand(B * Nq, H * D, seed=50), _rand(B * Nk, H * D, seed=51), _rand(B * Nk, H * D, seed=52)
    scale = D ** -0.5
    ldp = (Nk + 7) // 8 * 8
    probs = torch.zeros(B * H, Nq, ldp, device="cuda", dtype=torch.float16)
    ops.attn_scores(q, k, B, H, Nq, Nk, D, scale, probs)
    ops.softmax_(probs, Nk)
    out = torch.empty(B * Nq, H * D, device="cuda", dtype=torch.float16)
    ops.attn_pv(probs, v, B, H, Nq, Nk, D, out)
    ref_o, ref_p = _attn_ref(q, k, v, B, H, Nq, Nk, D, scale)
    _close(probs[..., :Nk], ref_p, rtol=2e-3, atol=2e-4, what="explicit probs")
    _close(out, ref_o, rtol=2e-3, atol=2e-3, what="explicit out")


@pytest.mark.parametrize("B,H,Nq,Nk,D", [(2, 8, 256, 256, 40), (1, 8, 4096, 4096, 40), (2, 8, 1024, 1024, 80),
                                         (2, 8, 64, 64, 160), (1, 10, 1024, 1024, 64), (2, 8, 200, 333, 64),
                                         (1, 8, 256, 256, 160)])
def test_attention_fused_self(ops, B, H, Nq, Nk, D):
    q, k, v = _rand(B * Nq, H * D, seed=60), _rand(B * Nk, H * D, seed=61), _rand(B * Nk, H * D, seed=62)
    scale = D ** -0.5
    out = ops.attention(q, k, v, B, H, Nq, Nk, D, scale)
    ref_o, _ = _attn_ref(q, k, v, B, H, Nq, Nk, D, scale)
    _close(out, ref_o, rtol=2e-3, atol=2e-3, what=f"fused attn {Nq}x{Nk} d{D}")


@pytest.mark.parametrize("B,H,Nq,Nk,D,peaky", [(2, 8, 1024, 1024, 80, False), (2, 8, 256, 256, 160, False),
                                               (1, 8, 64, 64, 160, False), (2, 10, 1024, 1024, 64, True),
                                               (3, 8, 200, 136, 40, True)])
def test_attention_stats_and_probs_from_stats(ops, B, H, Nq, Nk, D, peaky):
    """Read-only capture of larger maps (self-attention, N_q <= 1024; utils/p2p.py:145-149): the fused kernel also writes
    its online-softmax statistics, and one GEMM pass over Q.K^T with the softmax-from-statistics epilogue emits the
    normalised probabilities. `peaky`: scores with a large spread, so the lazily-updated reference maximum moves."""
    q, k, v = _rand(B * Nq, H * D, seed=70), _rand(B * Nk, H * D, seed=71), _rand(B * Nk, H * D, seed=72)
    if peaky:
        q = (q.float() * 6).half()
    scale = D ** -0.5
    stats = torch.full((B * H, Nq, 2), float("nan"), device="cuda")
    out = ops.attention(q, k, v, B, H, Nq, Nk, D, scale, stats_out=stats)
    probs = torch.empty((B * H, Nq, Nk), device="cuda", dtype=torch.float16)
    ops.attn_probs_from_stats(q, k, B, H, Nq, Nk, D, scale, stats, probs)
    ref_o, ref_p = _attn_ref(q, k, v, B, H, Nq, Nk, D, scale)
    assert torch.isfinite(stats).all()
    _close(out, ref_o, rtol=2e-3, atol=2e-3, what="fused attn with stats")
    _close(probs, ref_p, rtol=4e-3, atol=2e-4, what="probs from stats")
    torch.testing.assert_close(probs.float().sum(-1), torch.ones(B * H, Nq, device="cuda"), rtol=0, atol=4e-3)


@pytest.mark.parametrize("B,H,Nq,D", [(2, 8, 1024, 80), (2, 8, 4096, 40), (1, 20, 1024, 64), (2, 8, 64, 160)])
def test_attention_fused_cross_with_capture(ops, B, H, Nq, D):
    Nk = 77
    q = _rand(B * Nq, H * D, seed=63)
    kv = _rand(B * Nk, 2 * H * D, seed=64)          # K and V as column slices of one projection output
    k, v = kv[:, :H * D], kv[:, H * D:]
    scale = D ** -0.5
    # garbage-initialised capture buffer: the kernel must write every padded row completely (pad columns = 0)
    probs = torch.full((B * H, Nq, 80), float("nan"), device="cuda", dtype=torch.float16)
    out = ops.attention(q, k, v, B, H, Nq, Nk, D, scale, probs_out=probs)
    ref_o, ref_p = _attn_ref(q, k, v, B, H, Nq, Nk, D, scale)
    _close(out, ref_o, rtol=2e-3, atol=2e-3, what="fused cross out")
    _close(probs[..., :Nk], ref_p, rtol=2e-3, atol=2e-4, what="fused cross probs")
    assert probs[..., Nk:].abs().max() == 0


@pytest.mark.parametrize("B,H,Nq,Nk,D,stats", [(8, 8, 4096, 77, 40, True),     # 7-tile runs per CTA, last run shorter
                                               (10, 8, 1100, 77, 64, False),   # 3-tile runs, partial last query tile
                                               (4, 20, 1024, 77, 64, False), (8, 8, 1024, 77, 80, True),
                                               (8, 8, 256, 77, 160, False), (2, 8, 64, 77, 160, True),
                                               (3, 5, 500, 80, 40, False), (3, 5, 130, 9, 64, True)])
def test_attention_short_context_kernel(ops, B, H, Nq, Nk, D, stats):
    """attention_smallkv_kernel (N_kv <= 80: the text context): one CTA walks a run of query tiles with K/V resident,
    exact one-pass softmax, probabilities normalised before P.V. Without capture; optional statistics export."""
    q = _rand(B * Nq, H * D, seed=73) * 2
    kv = _rand(B * Nk, 2 * H * D, seed=74) * 2
    k, v = kv[:, :H * D], kv[:, H * D:]
    scale = D ** -0.5
    st = torch.full((B * H, Nq, 2), float("nan"), device="cuda") if stats else None
    out = ops.attention(q, k, v, B, H, Nq, Nk, D, scale, stats_out=st)
    ref_o, ref_p = _attn_ref(q, k, v, B, H, Nq, Nk, D, scale)
    _close(out, ref_o, rtol=2e-3, atol=2e-3, what="short-context out")
    if stats:
        sc = (q.float().reshape(B, Nq, H, D).permute(0, 2, 1, 3) @
              k.float().reshape(B, Nk, H, D).permute(0, 2, 3, 1)) * scale
        m = sc.max(-1).values.reshape(B * H, Nq)
        l = torch.exp(sc - sc.max(-1, keepdim=True).values).sum(-1).reshape(B * H, Nq)
        torch.testing.assert_close(st[..., 0], m * 1.4426950408889634, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(st[..., 1], 1.0 / l, rtol=2e-3, atol=1e-6)


@pytest.mark.parametrize("Nq,Nk,ldp", [(200, 64, 64), (200, 128, 128), (77, 100, 104), (130, 13, 16), (96, 77, 77),
                                       (300, 77, 144)])
def test_attention_capture_ragged(ops, Nq, Nk, ldp):
    """Capture write-out on ragged shapes: partial query tiles, 1 and 2 key tiles, padded / over-wide / unaligned
    (scalar fallback) row strides. Large scores force the lazy rescale between the two key tiles."""
    B, H, D = 2, 3, 40
    q = _rand(B * Nq, H * D, seed=70) * 4
    k, v = _rand(B * Nk, H * D, seed=71) * 4, _rand(B * Nk, H * D, seed=72)
    scale = D ** -0.5
    probs = torch.zeros((B * H, Nq, ldp), device="cuda", dtype=torch.float16)
    out = ops.attention(q, k, v, B, H, Nq, Nk, D, scale, probs_out=probs)
    ref_o, ref_p = _attn_ref(q, k, v, B, H, Nq, Nk, D, scale)
    _close(out, ref_o, rtol=2e-3, atol=2e-3, what="ragged capture out")
    _close(probs[..., :Nk], ref_p, rtol=3e-3, atol=3e-4, what="ragged capture probs")
    if ldp > Nk:
        assert probs[..., Nk:].abs().max() == 0
