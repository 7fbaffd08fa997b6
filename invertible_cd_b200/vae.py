"""B200 VAE executor: the drop-in for `diffusers.AutoencoderKL` on the iCD path — the step after the
loop (`latent2image`, utils/generation.py:255-264,527-533; utils/generation_sdxl.py:465-466) and the step before the
inversion (`image2latent`, utils/generation.py:266-284; Img2Img `prepare_latents`, utils/generation_sdxl.py:273).

Runs on the same sm_100a kernels as the U-Net (ops.py -> libicd_b200.so): implicit-GEMM 3x3 convs over NHWC fp16
activations (image rows up to 1024 pixels wide: one 128-pixel row segment per M tile), GroupNorm(32, eps 1e-6)+SiLU
(4 channels per group at the 128-channel level), the single-head 512-wide mid-block attention through the
explicit scores-GEMM -> softmax -> P.V-GEMM path (head dim 512 exceeds the fused kernel's tensor-memory budget), the
encoder's asymmetrically padded stride-2 convs through the im2col gather. No PyTorch-eager or CPU fallback.

Numerics: fp16 operands, fp32 accumulation / statistics / softmax — also where the reference forces the SDXL VAE to
fp32 (utils/generation_sdxl.py:465): with the stock SDXL VAE weights fp16 activations overflow inside the decoder, so
real SDXL checkpoints need the fp16-safe VAE weights (`sdxl-vae-fp16-fix`); synthetic weights are unaffected.
"""
from types import SimpleNamespace

import numpy as np
import torch

from . import ops
from .packing import pack_conv3x3, pack_linear

_ATTN_PROBS_BUDGET = 2 << 30          # bytes of materialised probabilities per attention chunk


class VAEOutput(dict):
    """`["sample"]` / `.sample` / `["latent_dist"]` / `.latent_dist` access like diffusers' output classes."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


class DiagonalGaussianDistribution:
    """moments (B, 2*latent, h, w) -> mean / logvar (clamped to [-30, 20]) / std; `sample(generator)` draws the noise
    with the given (CPU) generator in the moments' dtype like diffusers' randn_tensor."""

    def __init__(self, parameters):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None):
        dev = generator.device if generator is not None else self.parameters.device
        noise = torch.randn(self.mean.shape, generator=generator, device=dev, dtype=self.parameters.dtype)
        return self.mean + self.std * noise.to(self.parameters.device)

    def mode(self):
        return self.mean


class VaeImageProcessor:
    """`postprocess(image, output_type, do_denormalize)` of diffusers' VaeImageProcessor, as used at
    utils/generation_sdxl.py:468."""

    @staticmethod
    def postprocess(image, output_type="pil", do_denormalize=None):
        if do_denormalize is None:
            do_denormalize = [True] * image.shape[0]
        image = torch.stack([(im / 2 + 0.5).clamp(0, 1) if d else im for im, d in zip(image, do_denormalize)])
        if output_type == "pt":
            return image
        arr = image.detach().cpu().permute(0, 2, 3, 1).float().numpy()
        if output_type == "np":
            return arr
        from PIL import Image
        return [Image.fromarray((a * 255).round().astype("uint8")) for a in arr]


def vae_param_shapes(cfg):
    """key -> shape for every AutoencoderKL parameter, diffusers state-dict naming."""
    P = {}
    boc, L = list(cfg.block_out_channels), cfg.latent_channels

    def conv(name, cin, cout, k):
        P[name + ".weight"] = (cout, cin, k, k)
        P[name + ".bias"] = (cout,)

    def norm(name, c):
        P[name + ".weight"] = (c,)
        P[name + ".bias"] = (c,)

    def lin(name, cin, cout):
        P[name + ".weight"] = (cout, cin)
        P[name + ".bias"] = (cout,)

    def res(name, cin, cout):
        norm(name + ".norm1", cin)
        conv(name + ".conv1", cin, cout, 3)
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".conv_shortcut", cin, cout, 1)

    def mid(name, c):
        norm(name + ".attentions.0.group_norm", c)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(f"{name}.attentions.0.{n}", c, c)
        res(name + ".resnets.0", c, c)
        res(name + ".resnets.1", c, c)

    conv("encoder.conv_in", cfg.in_channels, boc[0], 3)
    out = boc[0]
    for i, c in enumerate(boc):
        cin, out = out, c
        for j in range(cfg.layers_per_block):
            res(f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else out, out)
        if i != len(boc) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", out, out, 3)
    mid("encoder.mid_block", boc[-1])
    norm("encoder.conv_norm_out", boc[-1])
    conv("encoder.conv_out", boc[-1], 2 * L, 3)
    conv("decoder.conv_in", L, boc[-1], 3)
    mid("decoder.mid_block", boc[-1])
    rev = boc[::-1]
    out = rev[0]
    for i, c in enumerate(rev):
        prev, out = out, c
        for j in range(cfg.layers_per_block + 1):
            res(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else out, out)
        if i != len(boc) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", out, out, 3)
    norm("decoder.conv_norm_out", boc[0])
    conv("decoder.conv_out", boc[0], cfg.out_channels, 3)
    conv("quant_conv", 2 * L, 2 * L, 1)
    conv("post_quant_conv", L, L, 1)
    return P


def vae_config(**kw):
    base = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                layers_per_block=2, norm_num_groups=32, scaling_factor=0.18215, sample_size=512)
    base.update(kw)
    return SimpleNamespace(**base)


def synthetic_vae_state_dict(cfg, seed=0, device="cpu", dtype=torch.float16):
    """Random-init AutoencoderKL weights (PyTorch default init statistics), there being no network for checkpoints."""
    import math
    g = torch.Generator(device=device).manual_seed(seed)
    shapes = vae_param_shapes(cfg)
    sd = {}
    for key, shape in shapes.items():
        if "norm" in key and len(shape) == 1:
            sd[key] = (torch.ones if key.endswith(".weight") else torch.zeros)(shape, device=device, dtype=dtype)
            continue
        wshape = shape if key.endswith(".weight") else shapes[key[:-len(".bias")] + ".weight"]
        bound = 1.0 / math.sqrt(math.prod(wshape[1:]))
        sd[key] = torch.empty(shape, device=device, dtype=torch.float32).uniform_(-bound, bound, generator=g).to(dtype)
    return sd


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


class B200VAE:
    """Packed weights + encode / decode of one AutoencoderKL."""

    def __init__(self, config, state_dict, device="cuda"):
        self.config = config if not isinstance(config, dict) else SimpleNamespace(**config)
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dtype = torch.float32          # dtype callers are told to hand inputs over in (results are fp32)
        self._pack(state_dict)

    def to(self, *args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.dtype):
                self.dtype = a
        return self

    def eval(self):
        return self

    # ------------------------------------------------------------------ packing
    def _pack(self, sd):
        cfg, dev = self.config, self.device
        if cfg.norm_num_groups != 32:
            raise ValueError("B200VAE: GroupNorm with 32 groups only")
        g = lambda k: sd[k]

        def conv(p):
            return SimpleNamespace(w=pack_conv3x3(g(p + ".weight")).to(dev), b=_f32(g(p + ".bias"), dev))

        def lin(p):
            return SimpleNamespace(w=pack_linear(g(p + ".weight")).to(dev), b=_f32(g(p + ".bias"), dev))

        def norm(p):
            return SimpleNamespace(g=_f32(g(p + ".weight"), dev), b=_f32(g(p + ".bias"), dev))

        def res(p, cin, cout):
            return SimpleNamespace(cin=cin, cout=cout, norm1=norm(p + ".norm1"), conv1=conv(p + ".conv1"),
                                   norm2=norm(p + ".norm2"), conv2=conv(p + ".conv2"),
                                   shortcut=lin(p + ".conv_shortcut") if cin != cout else None)

        def mid(p, c):
            a = p + ".attentions.0"
            return SimpleNamespace(
                C=c, res0=res(p + ".resnets.0", c, c), res1=res(p + ".resnets.1", c, c), norm=norm(a + ".group_norm"),
                qkv=SimpleNamespace(w=torch.cat([pack_linear(g(f"{a}.to_{n}.weight")) for n in "qkv"], 0).to(dev),
                                    b=torch.cat([_f32(g(f"{a}.to_{n}.bias"), dev) for n in "qkv"], 0)),
                out=lin(a + ".to_out.0"))

        def small_1x1(p, cin, cout, kpad, npad):
            """1x1 conv over few channels: zero-padded to [npad, kpad] so the operand rows stay 16-byte aligned."""
            w = torch.zeros(npad, kpad, dtype=torch.float16, device=dev)
            w[:cout, :cin] = pack_linear(g(p + ".weight")).to(dev)
            b = torch.zeros(npad, dtype=torch.float32, device=dev)
            b[:cout] = _f32(g(p + ".bias"), dev)
            return SimpleNamespace(w=w, b=b)

        boc, L = list(cfg.block_out_channels), cfg.latent_channels
        if 2 * L > 8:
            raise ValueError("B200VAE: latent_channels > 4 unsupported")
        enc = SimpleNamespace(conv_in=conv("encoder.conv_in"), down=[])
        out = boc[0]
        for i, c in enumerate(boc):
            cin, out = out, c
            blk = SimpleNamespace(resnets=[res(f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else out, out)
                                           for j in range(cfg.layers_per_block)], down=None)
            if i != len(boc) - 1:
                blk.down = conv(f"encoder.down_blocks.{i}.downsamplers.0.conv")
            enc.down.append(blk)
        enc.mid = mid("encoder.mid_block", boc[-1])
        enc.norm_out = norm("encoder.conv_norm_out")
        # conv_out: 512 -> 2L (= 8) channels, kept as an [M, 8] fp16 matrix for the 1x1 quant_conv that follows
        enc.conv_out = conv("encoder.conv_out")
        self.enc = enc
        self.quant = small_1x1("quant_conv", 2 * L, 2 * L, 8, 8)
        self.post_quant = small_1x1("post_quant_conv", L, L, 8, 8)
        dec = SimpleNamespace(conv_in=conv("decoder.conv_in"), mid=mid("decoder.mid_block", boc[-1]), up=[])
        rev = boc[::-1]
        out = rev[0]
        for i, c in enumerate(rev):
            prev, out = out, c
            blk = SimpleNamespace(resnets=[res(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else out, out)
                                           for j in range(cfg.layers_per_block + 1)], up=None)
            if i != len(boc) - 1:
                blk.up = conv(f"decoder.up_blocks.{i}.upsamplers.0.conv")
            dec.up.append(blk)
        dec.norm_out = norm("decoder.conv_norm_out")
        dec.conv_out = conv("decoder.conv_out")
        self.dec = dec

    # ------------------------------------------------------------------ blocks
    def _res(self, r, x, B, H, W):
        h = ops.groupnorm(x, B, H * W, r.norm1.g, r.norm1.b, 1e-6, True, self._ws)
        h = ops.conv3x3(h, r.conv1.w, B, H, W, bias=r.conv1.b)
        h = ops.groupnorm(h, B, H * W, r.norm2.g, r.norm2.b, 1e-6, True, self._ws)
        sc = x if r.shortcut is None else ops.linear(x, r.shortcut.w, bias=r.shortcut.b)
        return ops.conv3x3(h, r.conv2.w, B, H, W, bias=r.conv2.b, residual=sc)

    def _mid(self, m, x, B, H, W):
        x = self._res(m.res0, x, B, H, W)
        N, C = H * W, m.C
        h = ops.groupnorm(x, B, N, m.norm.g, m.norm.b, 1e-6, False, self._ws)
        qkv = ops.linear(h, m.qkv.w, bias=m.qkv.b)
        a = torch.empty((B * N, C), device=x.device, dtype=torch.float16)
        ldp = (N + 7) // 8 * 8
        per = max(1, min(B, _ATTN_PROBS_BUDGET // (N * ldp * 2)))
        for b0 in range(0, B, per):          # single head of width C: explicit probabilities, a few images at a time
            nb = min(per, B - b0)
            rows = slice(b0 * N, (b0 + nb) * N)
            probs = torch.zeros((nb, N, ldp), device=x.device, dtype=torch.float16)
            ops.attn_scores(qkv[rows, :C], qkv[rows, C:2 * C], nb, 1, N, N, C, C ** -0.5, probs)
            ops.softmax_(probs, N)
            ops.attn_pv(probs, qkv[rows, 2 * C:], nb, 1, N, N, C, a[rows])
            del probs
        x = ops.linear(a, m.out.w, bias=m.out.b, residual=x)
        return self._res(m.res1, x, B, H, W)

    # ------------------------------------------------------------------ public API
    @torch.no_grad()
    def decode(self, z, return_dict=True, generator=None):
        """z: (B, 4, h, w) latents (already divided by the scaling factor) -> image (B, 3, 8h, 8w) fp32."""
        with torch.cuda.device(self.device):
            dev, d = self.device, self.dec
            B, _, H, W = z.shape
            self._ws = torch.empty(B * 4096, device=dev, dtype=torch.float32)
            x = ops.latent_to_nhwc(z.to(device=dev, dtype=torch.float32).contiguous(), cpad=8)
            x = ops.linear(x, self.post_quant.w, bias=self.post_quant.b)              # 1x1 post_quant_conv, [M, 8]
            x = ops.conv3x3(x, d.conv_in.w, B, H, W, bias=d.conv_in.b)
            x = self._mid(d.mid, x, B, H, W)
            for blk in d.up:
                for r in blk.resnets:
                    x = self._res(r, x, B, H, W)
                if blk.up is not None:
                    x = ops.conv3x3(ops.upsample2x(x, B, H, W), blk.up.w, B, 2 * H, 2 * W, bias=blk.up.b)
                    H, W = 2 * H, 2 * W
            x = ops.groupnorm(x, B, H * W, d.norm_out.g, d.norm_out.b, 1e-6, True, self._ws)
            img = torch.empty((B, self.config.out_channels, H, W), device=dev, dtype=torch.float32)
            ops.conv3x3(x, d.conv_out.w, B, H, W, bias=d.conv_out.b, nchw_out=img)
            self._ws = None
        return VAEOutput(sample=img) if return_dict else (img,)

    @torch.no_grad()
    def encode(self, x, return_dict=True):
        """x: (B, 3, H, W) image in [-1, 1] -> latent_dist over (B, 4, H/8, W/8) (moments fp32)."""
        with torch.cuda.device(self.device):
            dev, e = self.device, self.enc
            B, _, H, W = x.shape
            self._ws = torch.empty(B * 4096, device=dev, dtype=torch.float32)
            h = ops.latent_to_nhwc(x.to(device=dev, dtype=torch.float32).contiguous(), cpad=8)
            h = ops.conv3x3(h, e.conv_in.w, B, H, W, bias=e.conv_in.b)
            for blk in e.down:
                for r in blk.resnets:
                    h = self._res(r, h, B, H, W)
                if blk.down is not None:
                    h = ops.linear(ops.im2col_s2(h, B, H, W, pad=0), blk.down.w, bias=blk.down.b)
                    H, W = H // 2, W // 2
            h = self._mid(e.mid, h, B, H, W)
            h = ops.groupnorm(h, B, H * W, e.norm_out.g, e.norm_out.b, 1e-6, True, self._ws)
            h = ops.conv3x3(h, e.conv_out.w, B, H, W, bias=e.conv_out.b)               # [M, 2L] fp16
            L2 = 2 * self.config.latent_channels
            m = ops.linear(h, self.quant.w, bias=self.quant.b, out_fp32=True)          # 1x1 quant_conv -> [M, 8] fp32
            moments = m[:, :L2].reshape(B, H, W, L2).permute(0, 3, 1, 2).contiguous()
            self._ws = None
        dist = DiagonalGaussianDistribution(moments)
        return VAEOutput(latent_dist=dist) if return_dict else (dist,)


def latents_to_uint8(image):
    """(image / 2 + 0.5).clamp(0, 1) -> uint8 NHWC numpy (utils/generation.py:529-532)."""
    image = (image / 2 + 0.5).clamp(0, 1)
    return (image.cpu().permute(0, 2, 3, 1).numpy() * 255).astype(np.uint8)
