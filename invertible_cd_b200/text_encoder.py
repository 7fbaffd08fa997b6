"""B200 CLIP text encoder: the drop-in for `transformers.CLIPTextModel` / `CLIPTextModelWithProjection` on the iCD path
 — the step before the loop: `self.model.text_encoder(input_ids)[0]` (utils/generation.py:286-303) and
`text_encoder(ids, output_hidden_states=True)` -> `[0]` / `.hidden_states[-2]` (utils/generation_sdxl.py:9-46).

Runs on the same sm_100a kernels as the U-Net (ops.py -> libicd_b200.so): token + position embedding gather, pre-LN
transformer layers (LayerNorm, fused Q|K|V projection, causal attention over the 77 tokens through the
scores-GEMM -> causally-masked softmax -> P.V-GEMM path, output projection with the residual in the GEMM epilogue, MLP
with quick_gelu (CLIP-L) or exact GELU (OpenCLIP bigG)), final LayerNorm, EOS pooling and the optional text projection.
State-dict keys are transformers' (`text_model.embeddings.token_embedding.weight`, `text_model.encoder.layers.{i}.…`),
so `text_encoder/model.safetensors` of a diffusers directory loads as is. Tokenisation stays on the host
(transformers' CLIPTokenizer when a local `tokenizer/` directory exists): strings are not GPU work.
"""
from types import SimpleNamespace

import torch

from . import ops
from .packing import pack_linear

_ACTS = {"quick_gelu": ops.ACT_QUICK_GELU, "gelu": ops.ACT_GELU, "silu": ops.ACT_SILU}


class CLIPTextOutput(tuple):
    """Tuple-like (as transformers' ModelOutput): `[0]` = last_hidden_state (CLIPTextModel) or text_embeds
    (CLIPTextModelWithProjection); attributes `last_hidden_state`, `pooler_output`, `text_embeds`, `hidden_states`."""

    def __new__(cls, first, **fields):
        self = super().__new__(cls, (first,))
        self.__dict__.update(fields)
        return self


def clip_text_config(**kw):
    """Defaults = CLIP ViT-L/14 text tower (SD1.5 `text_encoder`, SDXL `text_encoder`)."""
    base = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5, eos_token_id=2,
                projection_dim=None)
    base.update(kw)
    return SimpleNamespace(**base)


def clip_bigg_text_config(**kw):
    """OpenCLIP ViT-bigG/14 text tower with projection (SDXL `text_encoder_2`)."""
    base = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=20,
                hidden_act="gelu", projection_dim=1280)
    base.update(kw)
    return clip_text_config(**base)


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


class B200CLIPTextModel:
    def __init__(self, config, state_dict, device="cuda"):
        self.config = config if not isinstance(config, dict) else SimpleNamespace(**config)
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dtype = torch.float16
        cfg, dev = self.config, self.device
        if cfg.hidden_act not in _ACTS:
            raise ValueError(f"B200CLIPTextModel: unsupported activation {cfg.hidden_act}")
        if (cfg.hidden_size // cfg.num_attention_heads) % 8 != 0:
            raise ValueError("B200CLIPTextModel: head dim must be a multiple of 8")
        g = lambda k: state_dict[k]
        tm = "text_model."
        self.tok = g(tm + "embeddings.token_embedding.weight").to(dev, torch.float16).contiguous()
        self.pos = g(tm + "embeddings.position_embedding.weight").to(dev, torch.float16).contiguous()

        def lin(p):
            return SimpleNamespace(w=pack_linear(g(p + ".weight")).to(dev), b=_f32(g(p + ".bias"), dev))

        def norm(p):
            return SimpleNamespace(g=_f32(g(p + ".weight"), dev), b=_f32(g(p + ".bias"), dev))

        self.layers = []
        for i in range(cfg.num_hidden_layers):
            p = f"{tm}encoder.layers.{i}."
            a = p + "self_attn."
            self.layers.append(SimpleNamespace(
                ln1=norm(p + "layer_norm1"), ln2=norm(p + "layer_norm2"),
                qkv=SimpleNamespace(w=torch.cat([pack_linear(g(f"{a}{n}_proj.weight")) for n in "qkv"], 0).to(dev),
                                    b=torch.cat([_f32(g(f"{a}{n}_proj.bias"), dev) for n in "qkv"], 0)),
                out=lin(a + "out_proj"), fc1=lin(p + "mlp.fc1"), fc2=lin(p + "mlp.fc2")))
        self.final_ln = norm(tm + "final_layer_norm")
        self.proj = (pack_linear(g("text_projection.weight")).to(dev)
                     if getattr(cfg, "projection_dim", None) and "text_projection.weight" in state_dict else None)

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    @torch.no_grad()
    def __call__(self, input_ids, attention_mask=None, output_hidden_states=False, return_dict=True):
        with torch.cuda.device(self.device):
            cfg, dev = self.config, self.device
            ids = input_ids.to(device=dev, dtype=torch.int64).contiguous()
            B, T = ids.shape
            C, H = cfg.hidden_size, cfg.num_attention_heads
            d = C // H
            act = _ACTS[cfg.hidden_act]
            x = ops.embed_tokens(ids, self.tok, self.pos)
            hidden = [x]
            ldp = (T + 7) // 8 * 8
            for L in self.layers:
                n = ops.layernorm(x, L.ln1.g, L.ln1.b, cfg.layer_norm_eps)
                qkv = ops.linear(n, L.qkv.w, bias=L.qkv.b)
                probs = torch.zeros((B * H, T, ldp), device=dev, dtype=torch.float16)
                ops.attn_scores(qkv[:, :C], qkv[:, C:2 * C], B, H, T, T, d, d ** -0.5, probs)
                ops.softmax_(probs, T, causal_period=T)
                a = torch.empty((B * T, C), device=dev, dtype=torch.float16)
                ops.attn_pv(probs, qkv[:, 2 * C:], B, H, T, T, d, a)
                x = ops.linear(a, L.out.w, bias=L.out.b, residual=x)
                n = ops.layernorm(x, L.ln2.g, L.ln2.b, cfg.layer_norm_eps)
                h = ops.act(ops.linear(n, L.fc1.w, bias=L.fc1.b), act)
                x = ops.linear(h, L.fc2.w, bias=L.fc2.b, residual=x)
                hidden.append(x)
            last = ops.layernorm(x, self.final_ln.g, self.final_ln.b, cfg.layer_norm_eps).reshape(B, T, C)
            # pooled = the hidden state at the EOS token: transformers takes argmax(input_ids) when eos_token_id == 2
            # (the original CLIP vocabulary, where EOS is the largest id), else the first occurrence of eos_token_id
            if getattr(cfg, "eos_token_id", 2) == 2:
                eos = ids.argmax(dim=-1)
            else:
                eos = (ids == cfg.eos_token_id).int().argmax(dim=-1)
            pooled = last[torch.arange(B, device=dev), eos]
            hs = tuple(h.reshape(B, T, C) for h in hidden) if output_hidden_states else None
            if self.proj is not None:
                text_embeds = ops.linear(pooled.contiguous(), self.proj)
                return CLIPTextOutput(text_embeds, text_embeds=text_embeds, last_hidden_state=last,
                                      hidden_states=hs)
            return CLIPTextOutput(last, last_hidden_state=last, pooler_output=pooled, hidden_states=hs)


def clip_param_shapes(cfg):
    C, F = cfg.hidden_size, cfg.intermediate_size
    P = {"text_model.embeddings.token_embedding.weight": (cfg.vocab_size, C),
         "text_model.embeddings.position_embedding.weight": (cfg.max_position_embeddings, C)}
    for i in range(cfg.num_hidden_layers):
        p = f"text_model.encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            P[f"{p}self_attn.{n}.weight"] = (C, C)
            P[f"{p}self_attn.{n}.bias"] = (C,)
        for n in ("layer_norm1", "layer_norm2"):
            P[f"{p}{n}.weight"] = (C,)
            P[f"{p}{n}.bias"] = (C,)
        P[p + "mlp.fc1.weight"], P[p + "mlp.fc1.bias"] = (F, C), (F,)
        P[p + "mlp.fc2.weight"], P[p + "mlp.fc2.bias"] = (C, F), (C,)
    P["text_model.final_layer_norm.weight"] = (C,)
    P["text_model.final_layer_norm.bias"] = (C,)
    if getattr(cfg, "projection_dim", None):
        P["text_projection.weight"] = (cfg.projection_dim, C)
    return P


def synthetic_clip_state_dict(cfg, seed=0, device="cpu", dtype=torch.float16):
    """Random-init CLIP text weights (N(0, 0.02) matrices, unit LayerNorms), there being no network for checkpoints."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for key, shape in clip_param_shapes(cfg).items():
        if "layer_norm" in key:
            sd[key] = (torch.ones if key.endswith("weight") else torch.zeros)(shape, device=device, dtype=dtype)
        elif key.endswith(".bias"):
            sd[key] = (torch.randn(shape, generator=g, device=device) * 0.02).to(dtype)
        else:
            sd[key] = (torch.randn(shape, generator=g, device=device) * 0.02).to(dtype)
    return sd
