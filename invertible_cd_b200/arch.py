"""U-Net architecture description for the iCD models: configs and the parameter inventory (diffusers key names).

The numerical engine the reference drives is diffusers-0.25.1 `UNet2DConditionModel`; this
module only knows *what tensors exist and their shapes* so that checkpoints can be validated/packed and synthetic
(random-init) models of the exact published shapes can be built for benchmarking without any weights on disk.
"""
import math
from collections import OrderedDict
from dataclasses import dataclass
from typing import Optional, Tuple

import torch


@dataclass
class UNetConfig:
    sample_size: int = 64
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    down_block_types: Tuple[str, ...] = ("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",)
    up_block_types: Tuple[str, ...] = ("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 1, 1, 1)
    attention_head_dim: Tuple[int, ...] = (8, 8, 8, 8)     # diffusers quirk: number of heads per level
    cross_attention_dim: int = 768
    use_linear_projection: bool = False
    norm_num_groups: int = 32
    time_cond_proj_dim: Optional[int] = 512                # utils/loading.py:50,102
    addition_embed_type: Optional[str] = None
    addition_time_embed_dim: Optional[int] = None
    projection_class_embeddings_input_dim: Optional[int] = None

    def __getitem__(self, k):
        return getattr(self, k)


def sd15_config(**kw):
    return UNetConfig(**kw)


def sdxl_config(**kw):
    base = dict(sample_size=128, block_out_channels=(320, 640, 1280),
                down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
                up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"),
                transformer_layers_per_block=(1, 2, 10), attention_head_dim=(5, 10, 20), cross_attention_dim=2048,
                use_linear_projection=True, addition_embed_type="text_time", addition_time_embed_dim=256,
                projection_class_embeddings_input_dim=2816)
    base.update(kw)
    return UNetConfig(**base)


def small_sd15_config(**kw):
    """SD1.5 topology at reduced width/depth whose every shape is supported by the kernels (head dim 64,
    >= 8 channels per GroupNorm group): used by the whole-forward parity tests so the CPU oracle runs in seconds."""
    base = dict(sample_size=16, block_out_channels=(256, 512, 512, 512), attention_head_dim=(4, 8, 8, 8),
                cross_attention_dim=128)
    base.update(kw)
    return UNetConfig(**base)


def small_sdxl_config(**kw):
    base = dict(sample_size=16, block_out_channels=(256, 512, 512),
                down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
                up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"),
                transformer_layers_per_block=(1, 1, 2), attention_head_dim=(4, 8, 8), cross_attention_dim=128,
                use_linear_projection=True, addition_embed_type="text_time", addition_time_embed_dim=32,
                projection_class_embeddings_input_dim=6 * 32 + 64)
    base.update(kw)
    return UNetConfig(**base)


NAMED_CONFIGS = {"sd15": sd15_config, "sdxl": sdxl_config, "small_sd15": small_sd15_config,
                 "small_sdxl": small_sdxl_config}


def unet_param_shapes(cfg):
    """OrderedDict key -> shape for every parameter of the U-Net, in diffusers state-dict naming."""
    P = OrderedDict()
    boc = list(cfg.block_out_channels)
    temb = boc[0] * 4

    def conv(name, cin, cout, k):
        P[name + ".weight"] = (cout, cin, k, k)
        P[name + ".bias"] = (cout,)

    def lin(name, cin, cout, bias=True):
        P[name + ".weight"] = (cout, cin)
        if bias:
            P[name + ".bias"] = (cout,)

    def norm(name, c):
        P[name + ".weight"] = (c,)
        P[name + ".bias"] = (c,)

    def res(name, cin, cout):
        norm(name + ".norm1", cin)
        conv(name + ".conv1", cin, cout, 3)
        lin(name + ".time_emb_proj", temb, cout)
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".conv_shortcut", cin, cout, 1)

    def tfm(name, c, depth):
        norm(name + ".norm", c)
        if cfg.use_linear_projection:
            lin(name + ".proj_in", c, c)
        else:
            conv(name + ".proj_in", c, c, 1)
        for k in range(depth):
            b = f"{name}.transformer_blocks.{k}"
            norm(b + ".norm1", c)
            for n in ("to_q", "to_k", "to_v"):
                lin(f"{b}.attn1.{n}", c, c, bias=False)
            lin(b + ".attn1.to_out.0", c, c)
            norm(b + ".norm2", c)
            lin(b + ".attn2.to_q", c, c, bias=False)
            lin(b + ".attn2.to_k", cfg.cross_attention_dim, c, bias=False)
            lin(b + ".attn2.to_v", cfg.cross_attention_dim, c, bias=False)
            lin(b + ".attn2.to_out.0", c, c)
            norm(b + ".norm3", c)
            lin(b + ".ff.net.0.proj", c, 8 * c)
            lin(b + ".ff.net.2", 4 * c, c)
        if cfg.use_linear_projection:
            lin(name + ".proj_out", c, c)
        else:
            conv(name + ".proj_out", c, c, 1)

    conv("conv_in", cfg.in_channels, boc[0], 3)
    lin("time_embedding.linear_1", boc[0], temb)
    if cfg.time_cond_proj_dim:
        lin("time_embedding.cond_proj", cfg.time_cond_proj_dim, boc[0], bias=False)
    lin("time_embedding.linear_2", temb, temb)
    if cfg.addition_embed_type == "text_time":
        lin("add_embedding.linear_1", cfg.projection_class_embeddings_input_dim, temb)
        lin("add_embedding.linear_2", temb, temb)
    depth = list(cfg.transformer_layers_per_block)
    out_ch = boc[0]
    for i, t in enumerate(cfg.down_block_types):
        in_ch, out_ch = out_ch, boc[i]
        for j in range(cfg.layers_per_block):
            res(f"down_blocks.{i}.resnets.{j}", in_ch if j == 0 else out_ch, out_ch)
            if t == "CrossAttnDownBlock2D":
                tfm(f"down_blocks.{i}.attentions.{j}", out_ch, depth[i])
        if i != len(boc) - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv", out_ch, out_ch, 3)
    res("mid_block.resnets.0", boc[-1], boc[-1])
    tfm("mid_block.attentions.0", boc[-1], depth[-1])
    res("mid_block.resnets.1", boc[-1], boc[-1])
    rboc, rdepth = boc[::-1], depth[::-1]
    out_ch = rboc[0]
    for i, t in enumerate(cfg.up_block_types):
        prev, out_ch = out_ch, rboc[i]
        in_ch = rboc[min(i + 1, len(boc) - 1)]
        for j in range(cfg.layers_per_block + 1):
            skip = in_ch if j == cfg.layers_per_block else out_ch
            res(f"up_blocks.{i}.resnets.{j}", (prev if j == 0 else out_ch) + skip, out_ch)
            if t == "CrossAttnUpBlock2D":
                tfm(f"up_blocks.{i}.attentions.{j}", out_ch, rdepth[i])
        if i != len(boc) - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv", out_ch, out_ch, 3)
    norm("conv_norm_out", boc[0])
    conv("conv_out", boc[0], cfg.out_channels, 3)
    return P


def count_params(cfg):
    return sum(math.prod(s) for s in unet_param_shapes(cfg).values())


LORA_TARGETS = ("to_q", "to_k", "to_v", "to_out.0", "proj_in", "proj_out", "ff.net.0.proj", "ff.net.2", "conv1",
                "conv2", "conv_shortcut", "downsamplers.0.conv", "upsamplers.0.conv", "time_emb_proj")
# training/train_icd_sd15_lora.py:617-632


def lora_target_modules(cfg):
    """Module paths (without '.weight') that carry a LoRA adapter in the released iCD checkpoints."""
    mods = []
    for key in unet_param_shapes(cfg):
        if key.endswith(".weight"):
            mod = key[:-len(".weight")]
            if any(mod.endswith(t) for t in LORA_TARGETS) and "time_embedding" not in mod and "add_embedding" not in mod:
                mods.append(mod)
    return mods


def synthetic_state_dict(cfg, seed=0, device="cpu", dtype=torch.float16):
    """Random-init weights of the exact architecture (PyTorch default init statistics: U(+-1/sqrt(fan_in)) for
    conv/linear weights and biases, ones/zeros for norms). There is no network for real checkpoints."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = OrderedDict()
    shapes = unet_param_shapes(cfg)
    for key, shape in shapes.items():
        is_norm = (".norm" in key or key.startswith("conv_norm_out")) and len(shape) == 1
        if is_norm:
            sd[key] = (torch.ones if key.endswith(".weight") else torch.zeros)(shape, device=device, dtype=dtype)
            continue
        wshape = shape if key.endswith(".weight") else shapes[key[:-len(".bias")] + ".weight"]
        bound = 1.0 / math.sqrt(math.prod(wshape[1:]))
        t = torch.empty(shape, device=device, dtype=torch.float32).uniform_(-bound, bound, generator=g)
        sd[key] = t.to(dtype)
    return sd


def synthetic_lora(cfg, r=64, seed=1, device="cpu", std=0.02):
    """Random LoRA adapter in the peft checkpoint key format of utils/loading.py:13-15."""
    g = torch.Generator(device=device).manual_seed(seed)
    shapes = unet_param_shapes(cfg)
    out = OrderedDict()
    for mod in lora_target_modules(cfg):
        ws = shapes[mod + ".weight"]
        a_shape = (r, ws[1]) + tuple(ws[2:])
        b_shape = (ws[0], r) + ((1, 1) if len(ws) == 4 else ())
        base = "unet.base_model.model." + mod
        out[base + ".lora_A.weight"] = torch.randn(a_shape, generator=g, device=device) * std
        out[base + ".lora_B.weight"] = torch.randn(b_shape, generator=g, device=device) * std
    return out
