"""TEST INFRASTRUCTURE (oracle) — CPU restatement of diffusers-0.25.1 `AutoencoderKL`, the VAE the reference calls at
utils/generation.py:255-284,527-533 (`vae.decode(latents)['sample']`, `vae.encode(image)['latent_dist'].mean`) and
utils/generation_sdxl.py:273,465-466 (`prepare_latents` -> `vae.encode(...).latent_dist.sample(generator)`,
`vae.decode(latents / scaling_factor, return_dict=False)[0]` in fp32).

diffusers is a third-party dependency that is neither vendored in /root/reference nor installed here, so
this follows the published architecture: Encoder (conv_in, 4 DownEncoderBlock2D with 2 ResnetBlock2D each and an
asymmetrically padded stride-2 conv, UNetMidBlock2D with one single-head attention, GroupNorm/SiLU/conv_out to
2 x latent channels), quant_conv / post_quant_conv (1x1), Decoder (conv_in, the same mid block, 4 UpDecoderBlock2D with
3 ResnetBlock2D each and nearest-2x + conv upsamplers, GroupNorm/SiLU/conv_out). GroupNorm: 32 groups, eps 1e-6.
State-dict key names are the diffusers ones, so real `vae/diffusion_pytorch_model.safetensors` files load strictly.

Pin: the parameter count of the SD1.5 / SDXL VAE config is the public 83,653,863 (tests/test_host_cpu.py); beyond that
parity is UNPINNED like the U-Net oracle (no diffusers here, no reference tests). Only tests/, __graft_entry__.smoke()
and bench.py's baseline legs may import this module.
"""
from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class VAEConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215          # SDXL: 0.13025
    sample_size: int = 512

    def __getitem__(self, k):
        return getattr(self, k)


def sd15_vae_config(**kw):
    return VAEConfig(**kw)


def sdxl_vae_config(**kw):
    base = dict(scaling_factor=0.13025, sample_size=1024)
    base.update(kw)
    return VAEConfig(**base)


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    """Single-head spatial self-attention of the VAE mid block (heads = C / attention_head_dim = 1, bias everywhere,
    GroupNorm on the input, residual connection)."""

    def __init__(self, channels, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, channels, eps=1e-6)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        res = x
        t = self.group_norm(x).view(b, c, h * w).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        o = F.scaled_dot_product_attention(q.unsqueeze(1), k.unsqueeze(1), v.unsqueeze(1)).squeeze(1)
        o = self.to_out[0](o)
        return o.transpose(-1, -2).reshape(b, c, h, w) + res


class MidBlock(nn.Module):
    def __init__(self, ch, groups):
        super().__init__()
        self.attentions = nn.ModuleList([Attention(ch, groups)])
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, groups), ResnetBlock2D(ch, ch, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _ConvHolder(nn.Module):
    def __init__(self, conv):
        super().__init__()
        self.conv = conv


class DownEncoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.downsamplers = nn.ModuleList([_ConvHolder(nn.Conv2d(cout, cout, 3, stride=2, padding=0))]) if add_down \
            else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0].conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))
        return x


class UpDecoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_up):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.upsamplers = nn.ModuleList([_ConvHolder(nn.Conv2d(cout, cout, 3, padding=1))]) if add_up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0].conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return x


class Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        boc, g = list(cfg.block_out_channels), cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out = boc[0]
        for i, c in enumerate(boc):
            cin, out = out, c
            self.down_blocks.append(DownEncoderBlock2D(cin, out, cfg.layers_per_block, g, i != len(boc) - 1))
        self.mid_block = MidBlock(boc[-1], g)
        self.conv_norm_out = nn.GroupNorm(g, boc[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[-1], 2 * cfg.latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class Decoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        boc, g = list(cfg.block_out_channels), cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.latent_channels, boc[-1], 3, padding=1)
        self.mid_block = MidBlock(boc[-1], g)
        self.up_blocks = nn.ModuleList()
        rev = boc[::-1]
        out = rev[0]
        for i, c in enumerate(rev):
            prev, out = out, c
            self.up_blocks.append(UpDecoderBlock2D(prev, out, cfg.layers_per_block + 1, g, i != len(boc) - 1))
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
    def __init__(self, parameters):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, dtype=self.parameters.dtype).to(self.parameters.device)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


class _Out(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


class AutoencoderKL(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        self.config = cfg
        self.encoder = Encoder(cfg)
        self.decoder = Decoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)

    @property
    def dtype(self):
        return self.quant_conv.weight.dtype

    def encode(self, x, return_dict=True):
        dist = DiagonalGaussianDistribution(self.quant_conv(self.encoder(x)))
        return _Out(latent_dist=dist) if return_dict else (dist,)

    def decode(self, z, return_dict=True):
        img = self.decoder(self.post_quant_conv(z))
        return _Out(sample=img) if return_dict else (img,)
