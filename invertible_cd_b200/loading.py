"""Model loading for the iCD path: teacher + reverse/forward consistency students with LoRA fused at load.

Public surface of utils/loading.py (`load_models`, `load_models_xl`, `get_module_kohya_state_dict`) with the same
signatures and return order.  Differences forced by this build: there is no diffusers here, so a "pipeline" is a
light `ICDPipeline` container exposing exactly the attributes the loop touches (`.unet .vae
.tokenizer .text_encoder .scheduler .device .dtype`), `.unet` is a `B200UNet`, and the LoRA adapters are fused
into the packed fp16 weights by `fuse_lora` below (W += alpha/r * B.A, alpha = 8) instead of going
through `load_lora_weights` + `fuse_lora` of diffusers (utils/loading.py:66-71).
`model_id` may be
  * a local directory in diffusers layout (`unet/config.json` + `unet/diffusion_pytorch_model.safetensors`;
    optional `text_encoder/`, `tokenizer/` loaded with transformers when present), or
  * "synthetic:<sd15|sdxl|small_sd15|small_sdxl>[:seed]" — random-init weights of the exact architecture (there
    is no network access for real checkpoints on the benchmark boxes).
"""
import json
import os
from collections import OrderedDict

import torch

from . import arch
from .schedulers import DDIMScheduler, DDPMScheduler
from .unet import B200UNet
from .vae import B200VAE, VaeImageProcessor, synthetic_vae_state_dict, vae_config

LORA_ALPHA = 8  # utils/loading.py:19-21 (peft default lora_alpha with LoraConfig(r=...))


class ICDPipeline:
    """Attribute bag standing in for StableDiffusion(XL)(Img2Img)Pipeline on the iCD path."""

    def __init__(self, unet, scheduler, vae=None, tokenizer=None, text_encoder=None, device="cuda",
                 dtype=torch.float32, tokenizer_2=None, text_encoder_2=None):
        self.unet, self.scheduler, self.vae = unet, scheduler, vae
        self.tokenizer, self.text_encoder = tokenizer, text_encoder
        self.tokenizer_2, self.text_encoder_2 = tokenizer_2, text_encoder_2
        self.device, self.dtype = torch.device(device), dtype
        self.vae_scale_factor = 8
        self.image_processor = VaeImageProcessor()      # .postprocess(...), utils/generation_sdxl.py:468

    @property
    def _execution_device(self):
        return self.device

    def to(self, *args, **kwargs):
        return self

    def clone_with_unet(self, unet):
        return ICDPipeline(unet, self.scheduler, self.vae, self.tokenizer, self.text_encoder, self.device, self.dtype,
                           self.tokenizer_2, self.text_encoder_2)


def get_module_kohya_state_dict(module, prefix: str, dtype: torch.dtype, adapter_name: str = "default"):
    """peft LoRA keys -> kohya keys with alpha = 8 for every adapted module (utils/loading.py:10-23)."""
    out = {}
    for peft_key, weight in module.items():
        key = peft_key.replace("unet.base_model.model", prefix)
        key = key.replace("lora_A", "lora_down").replace("lora_B", "lora_up")
        key = key.replace(".", "_", key.count(".") - 2)
        out[key] = weight.to(dtype)
        if "lora_down" in key:
            out[f'{key.split(".")[0]}.alpha'] = torch.tensor(LORA_ALPHA).to(dtype)
    return out


def fuse_lora(state_dict, lora_weights, r=64, lora_dtype=torch.float16, alpha=LORA_ALPHA):
    """Return a copy of `state_dict` with W <- W + (alpha/r) * B.A for every adapted module.
    `lora_weights`: peft-format dict (`unet.base_model.model.<module>.lora_{A,B}.weight`). A and B are first cast to
    `lora_dtype` (fp16 for SD1.5, utils/loading.py:68,82; fp32 for SDXL, :122,141), the product is formed in fp32
    and the sum is cast back to the weight dtype. Conv adapters: mm(B.flatten(1), A.flatten(1)).reshape(W.shape)."""
    fused = OrderedDict(state_dict)
    prefix = "unet.base_model.model."
    mods = sorted({k[len(prefix):].rsplit(".lora_", 1)[0] for k in lora_weights if k.startswith(prefix)})
    if not mods and lora_weights:
        raise ValueError("fuse_lora: no peft-format keys ('unet.base_model.model.*.lora_A.weight') found")
    for mod in mods:
        wkey = mod + ".weight"
        if wkey not in fused:
            raise KeyError(f"LoRA adapter for unknown module {mod}")
        A = lora_weights[f"{prefix}{mod}.lora_A.weight"].to(lora_dtype).float()
        Bm = lora_weights[f"{prefix}{mod}.lora_B.weight"].to(lora_dtype).float()
        rank = A.shape[0]
        W = fused[wkey]
        delta = torch.mm(Bm.flatten(1).to(W.device), A.flatten(1).to(W.device)).reshape(W.shape)
        fused[wkey] = (W.float() + (alpha / rank) * delta).to(W.dtype)
    return fused


# ---------------------------------------------------------------------------------------------- sources
def _parse_synthetic(model_id):
    parts = model_id.split(":")
    name = parts[1] if len(parts) > 1 else "sd15"
    seed = int(parts[2]) if len(parts) > 2 else 0
    if name not in arch.NAMED_CONFIGS:
        raise ValueError(f"unknown synthetic model '{name}' (have {sorted(arch.NAMED_CONFIGS)})")
    return name, seed


def _load_tensor_file(path):
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    return torch.load(path, map_location="cpu")


def _gen_device(device, cfg=None):
    """Full-size synthetic models (0.86 / 2.6 G parameters) are generated directly on the GPU when there is one;
    small test models always on the CPU so that the same seed gives the same weights everywhere."""
    d = torch.device(device)
    big = cfg is None or arch.count_params(cfg) > 5e8
    return d if big and d.type == "cuda" and torch.cuda.is_available() else torch.device("cpu")


def _unet_source(model_id, w_embed_dim, is_xl, device="cpu"):
    """-> (config, state_dict, text parts dict)."""
    if isinstance(model_id, str) and model_id.startswith("synthetic"):
        name, seed = _parse_synthetic(model_id)
        cfg = arch.NAMED_CONFIGS[name](time_cond_proj_dim=w_embed_dim if w_embed_dim > 0 else None)
        return cfg, arch.synthetic_state_dict(cfg, seed=seed, device=_gen_device(device, cfg)), {}
    if not os.path.isdir(model_id):
        raise FileNotFoundError(f"model_id '{model_id}' is neither a local diffusers directory nor 'synthetic:*' "
                                "(no network access: hub ids cannot be resolved)")
    with open(os.path.join(model_id, "unet", "config.json")) as f:
        raw = json.load(f)
    fields = {k: raw[k] for k in arch.UNetConfig.__dataclass_fields__ if k in raw}
    n_levels = len(raw["block_out_channels"])
    for k in ("attention_head_dim", "transformer_layers_per_block"):
        v = fields.get(k, 1)
        fields[k] = tuple(v) if isinstance(v, (list, tuple)) else (v,) * n_levels
    for k in ("block_out_channels", "down_block_types", "up_block_types"):
        fields[k] = tuple(fields[k])
    fields["time_cond_proj_dim"] = w_embed_dim if w_embed_dim > 0 else None
    cfg = arch.UNetConfig(**fields)
    sd = None
    for fname in ("diffusion_pytorch_model.fp16.safetensors", "diffusion_pytorch_model.safetensors",
                  "diffusion_pytorch_model.bin"):
        p = os.path.join(model_id, "unet", fname)
        if os.path.exists(p):
            sd = _load_tensor_file(p)
            break
    if sd is None:
        raise FileNotFoundError(f"no U-Net weights under {model_id}/unet")
    if cfg.time_cond_proj_dim and "time_embedding.cond_proj.weight" not in sd:
        # from_pretrained(time_cond_proj_dim=...) creates the layer with default init; the teacher .pt overwrites it
        sd["time_embedding.cond_proj.weight"] = torch.zeros(cfg.block_out_channels[0], cfg.time_cond_proj_dim)
    return cfg, sd, _text_source(model_id, device)


def _vae_source(model_id, device, is_xl):
    """-> B200VAE or None. Synthetic models get a random-init AutoencoderKL of the published shape (83.65 M parameters;
    scaling factor 0.18215 / 0.13025); local diffusers directories load `vae/config.json` + weights when present."""
    if isinstance(model_id, str) and model_id.startswith("synthetic"):
        cfg = vae_config(scaling_factor=0.13025 if is_xl else 0.18215, sample_size=1024 if is_xl else 512)
        return B200VAE(cfg, synthetic_vae_state_dict(cfg, seed=7), device)
    vdir = os.path.join(model_id, "vae")
    if not os.path.isdir(vdir):
        return None
    with open(os.path.join(vdir, "config.json")) as f:
        raw = json.load(f)
    base = vars(vae_config())
    cfg = vae_config(**{k: (tuple(raw[k]) if isinstance(raw[k], list) else raw[k]) for k in base if k in raw})
    for fname in ("diffusion_pytorch_model.fp16.safetensors", "diffusion_pytorch_model.safetensors",
                  "diffusion_pytorch_model.bin"):
        p = os.path.join(vdir, fname)
        if os.path.exists(p):
            return B200VAE(cfg, _load_tensor_file(p), device)
    return None


def _text_source(model_id, device):
    """Tokenizers (host side: transformers' CLIPTokenizer) and CLIP text encoders (B200CLIPTextModel on the sm_100a
    kernels) of a local diffusers directory: `tokenizer[_2]/`, `text_encoder[_2]/{config.json, model.safetensors}`."""
    from .text_encoder import B200CLIPTextModel, clip_text_config
    text = {}
    for suffix in ("", "_2"):
        tdir, edir = os.path.join(model_id, "tokenizer" + suffix), os.path.join(model_id, "text_encoder" + suffix)
        if os.path.isdir(tdir):
            try:
                from transformers import CLIPTokenizer
                text["tokenizer" + suffix] = CLIPTokenizer.from_pretrained(tdir)
            except Exception as e:
                print(f"[loading] tokenizer{suffix} not loaded: {e}")
        if os.path.isdir(edir):
            with open(os.path.join(edir, "config.json")) as f:
                raw = json.load(f)
            base = vars(clip_text_config())
            ccfg = clip_text_config(**{k: raw[k] for k in base if k in raw})
            if "WithProjection" not in str(raw.get("architectures", "")):
                ccfg.projection_dim = None
            for fname in ("model.fp16.safetensors", "model.safetensors", "pytorch_model.bin"):
                p = os.path.join(edir, fname)
                if os.path.exists(p):
                    text["text_encoder" + suffix] = B200CLIPTextModel(ccfg, _load_tensor_file(p), device)
                    break
    return text


def _validate(cfg, sd):
    shapes = arch.unet_param_shapes(cfg)
    missing = [k for k in shapes if k not in sd]
    bad = [k for k in shapes if k in sd and tuple(sd[k].shape) != tuple(shapes[k])]
    if missing or bad:
        raise RuntimeError(f"U-Net state dict mismatch: {len(missing)} missing (e.g. {missing[:3]}), "
                           f"{len(bad)} wrong shape (e.g. {bad[:3]})")


def _lora_source(spec, cfg, r, device="cpu"):
    if spec is None:
        return None
    if isinstance(spec, dict):
        return spec
    if isinstance(spec, str) and spec.startswith("synthetic"):
        parts = spec.split(":")
        return arch.synthetic_lora(cfg, r=r, seed=int(parts[1]) if len(parts) > 1 else 1,
                                   device=_gen_device(device, cfg))
    return _load_tensor_file(spec)


def _adapter_mode(adapters):
    mode = adapters if adapters is not None else os.environ.get("ICD_LORA_ADAPTERS", "resident")
    if mode not in ("resident", "swap"):
        raise ValueError(f"adapters must be 'resident' or 'swap', got {mode!r}")
    return mode


def _students_swap(teacher_pipe, shared, cfg, ckpts, r, device, lora_dtype):
    """One packed U-Net for the teacher and both students: the checkpoints become hot-swappable adapters of `shared`
    (B200UNet.add_adapter); every pipeline's `.unet` is an AdapterView that activates its adapter on use."""
    teacher_pipe.unet = shared.adapter_view(None)
    out = []
    for name, ckpt in ckpts:
        if ckpt is None:
            out.append(None)
            continue
        print(f'{name} CD is loading from {ckpt if isinstance(ckpt, str) else "<state dict>"} (hot-swappable adapter)')
        shared.add_adapter(name.lower(), _lora_source(ckpt, cfg, r, device), lora_dtype=lora_dtype)
        out.append(teacher_pipe.clone_with_unet(shared.adapter_view(name.lower())))
    return out


# ---------------------------------------------------------------------------------------------- public API
def load_models(model_id, device, reverse_checkpoint, forward_checkpoint, r=64, w_embed_dim=0,
                teacher_checkpoint=None, dtype='fp32', adapters=None):
    """-> (ldm_stable, reverse_cons_model, forward_cons_model), as utils/loading.py:27-90.
    `dtype`: 'fp16' = the tcgen05 path (fp16 operands / activations, fp32 accumulation); 'fp32' = fp32 latents AND the
    fp32 validation kernels (ops_f32: fp32 operands, activations and accumulation on the FMA pipe — the reference's
    fp32 editing mode, running/sd1.5/launch_editing_iCD_sd1.5.sh:38; an SD1.5 row-forward takes 111 ms against 8.8 ms in fp16, eager). ICD_FP32_KERNELS=0 keeps
    the fp16 kernels under fp32 latents (the round-1 behaviour).
    `adapters` (extension; default from ICD_LORA_ADAPTERS, else 'resident'): 'resident' keeps three packed U-Nets as
    the reference keeps three pipelines; 'swap' keeps ONE plus the low-rank factors and re-fuses on the GPU when a
    different model is called (unet.B200UNet.set_adapter)."""
    mode = _adapter_mode(adapters)
    tdtype = torch.float32 if dtype == 'fp32' else torch.float16
    precision = "fp32" if dtype == 'fp32' and os.environ.get("ICD_FP32_KERNELS", "1") != "0" else "fp16"
    scheduler = DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
                              set_alpha_to_one=False)
    cfg, sd, text = _unet_source(model_id, w_embed_dim, is_xl=False, device=device)
    if w_embed_dim > 0:
        print(f'Forward CD is initialized with guidance embedding, dim {w_embed_dim}')
        if teacher_checkpoint is not None:
            print(f'Embedded model is loading from {teacher_checkpoint}')
            sd = teacher_checkpoint if isinstance(teacher_checkpoint, dict) else _load_tensor_file(teacher_checkpoint)
        elif not str(model_id).startswith("synthetic"):
            print('PROVIDE TEACHER')
    _validate(cfg, sd)
    text_encoder = text.get("text_encoder")
    ldm_stable = ICDPipeline(B200UNet(cfg, sd, device, precision), scheduler, _vae_source(model_id, device, False),
                             text.get("tokenizer"), text_encoder, device, tdtype)
    if mode == "swap":
        rev, fwd = _students_swap(ldm_stable, ldm_stable.unet, cfg,
                                  (("Reverse", reverse_checkpoint), ("Forward", forward_checkpoint)), r, device,
                                  torch.float16)
        return ldm_stable, rev, fwd
    students = []
    for name, ckpt in (("Reverse", reverse_checkpoint), ("Forward", forward_checkpoint)):
        if ckpt is None:
            students.append(None)
            continue
        print(f'{name} CD is loading from {ckpt if isinstance(ckpt, str) else "<state dict>"}')
        fused = fuse_lora(sd, _lora_source(ckpt, cfg, r, device), r=r, lora_dtype=torch.float16)
        students.append(ldm_stable.clone_with_unet(B200UNet(cfg, fused, device, precision)))
    return ldm_stable, students[0], students[1]


def load_models_xl(model_id, reverse_checkpoint, forward_checkpoint, teacher_checkpoint, device="cuda", r=64,
                   adapters=None):
    """-> (stable_pipe, pipe, forw_pipe), as utils/loading.py:93-147 (fp16 base, LoRA kept fp32 for the fuse).
    `adapters`: see load_models (in 'swap' mode the factors are rounded to the fp16 operands of the fuse GEMM)."""
    mode = _adapter_mode(adapters)
    cfg, sd, text = _unet_source(model_id, 512, is_xl=True, device=device)
    if teacher_checkpoint is not None:
        sd = teacher_checkpoint if isinstance(teacher_checkpoint, dict) else _load_tensor_file(teacher_checkpoint)
    sd = OrderedDict((k, v.to(torch.float16)) for k, v in sd.items())
    _validate(cfg, sd)
    scheduler = DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear")
    scheduler.num_train_timesteps = 1000
    stable_pipe = ICDPipeline(B200UNet(cfg, sd, device), scheduler, _vae_source(model_id, device, True),
                              text.get("tokenizer"), text.get("text_encoder"), device, torch.float16,
                              text.get("tokenizer_2"), text.get("text_encoder_2"))
    if mode == "swap":
        rev, fwd = _students_swap(stable_pipe, stable_pipe.unet, cfg,
                                  (("Reverse", reverse_checkpoint), ("Forward", forward_checkpoint)), r, device,
                                  torch.float32)
        return stable_pipe, rev, fwd
    pipes = []
    for name, ckpt in (("Reverse", reverse_checkpoint), ("Forward", forward_checkpoint)):
        print(f'{name} CD is loading from {ckpt if isinstance(ckpt, str) else "<state dict>"}')
        fused = fuse_lora(sd, _lora_source(ckpt, cfg, r, device), r=r, lora_dtype=torch.float32)
        pipes.append(stable_pipe.clone_with_unet(B200UNet(cfg, fused, device)))
    return stable_pipe, pipes[0], pipes[1]


def load_benchmark(path_to_prompts, path_to_images=None):
    """CSV reader of utils/loading.py:151-175 (kept for interface completeness; not on the accelerated path)."""
    import pandas as pd
    files = pd.read_csv(path_to_prompts)
    if path_to_images is None:
        return list(files['caption']), list(files['file_name'])
    return [(f"{path_to_images}/{row['file_name']}", {'before': row['old_caption'], 'after': row['edited_caption']},
             row['blended_words']) for _, row in files.reset_index().iterrows()]
