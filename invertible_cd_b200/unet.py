"""B200 U-Net executor: the drop-in for `diffusers.UNet2DConditionModel.forward` on the iCD path.

Replaces the call the reference makes at utils/generation.py:208,241-244 and utils/generation_sdxl.py:288-295,
445-453:   unet(sample, t, encoder_hidden_states=ctx, timestep_cond=w_emb, added_cond_kwargs=...)["sample"] / [0].

Host code is Python (as the reference's is); all arithmetic runs in hand-written sm_100a kernels reached through
the C ABI (invertible_cd_b200/ops.py -> libicd_b200.so).  There is no PyTorch-eager or CPU fallback.

Data layout in HBM
  activations  fp16 channels-last token matrices [rows*H*W, C]  (an NHWC image IS the transformer token matrix,
               so ResNet blocks and transformer blocks hand tensors to each other without a transpose)
  weights      fp16 [N, K] K-contiguous; 3x3 convs [Cout, (ky,kx,cin)]; q|k|v and k|v projections concatenated;
               GEGLU rows interleaved per 256-wide N tile; all 22/17 time_emb_proj matrices concatenated into one;
               the to_k|to_v matrices of EVERY cross-attention layer concatenated into one (the text context is the
               same for all of them: one GEMM per forward instead of 16 / 70 launch-bound ones)
  latents      fp32 NCHW at the boundary (what the reference's loop carries)
"""
import math
from types import SimpleNamespace

import torch

from . import ops, ops_f32
from . import packing

GEGLU_BN = 256


class UNetOutput(dict):
    """`["sample"]`, `.sample` and `[0]` access like diffusers' UNet2DConditionOutput."""

    def __getitem__(self, k):
        return dict.__getitem__(self, "sample" if k == 0 else k)

    @property
    def sample(self):
        return dict.__getitem__(self, "sample")


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def _capturing():
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


class B200UNet:
    """Packed weights + forward of one U-Net (teacher, forward-consistency or reverse-consistency model)."""

    def __init__(self, config, state_dict, device="cuda", precision="fp16"):
        """`precision`: 'fp16' = the tcgen05 path (fp16 operands and activations, fp32 accumulation);
        'fp32' = the validation path (ops_f32: fp32 everywhere, FMA-pipe contractions) that
        `load_models(dtype='fp32')` selects, as the reference's fp32 mode (utils/loading.py:38-41)."""
        if precision not in ("fp16", "fp32"):
            raise ValueError(f"precision must be 'fp16' or 'fp32', got {precision!r}")
        self.precision = precision
        self.ops = ops_f32 if precision == "fp32" else ops
        self.act_dtype = torch.float32 if precision == "fp32" else torch.float16
        self.config = config if not isinstance(config, dict) else SimpleNamespace(**config)
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.controller = None          # set by p2p.register_attention_control
        self.attn_places = []           # execution-ordered ('down'|'mid'|'up') per Attention module
        self._pack(state_dict)
        self._freq_cache = {}

    supports_cond_only = True       # rows may be the conditional half only (host logic in generation.py)

    # ------------------------------------------------------------------ reference-visible attributes
    @property
    def dtype(self):
        return self.act_dtype           # compute dtype of the kernels (utils/generation.py:241 casts inputs to it)

    @property
    def in_channels(self):
        return self.config.in_channels

    @property
    def num_attention_layers(self):
        return len(self.attn_places)

    def named_children(self):
        return iter(())

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    # ------------------------------------------------------------------ packing
    def _pack(self, sd):
        cfg, dev, wd = self.config, self.device, self.act_dtype
        g = lambda k: sd[k]
        pack_conv3x3 = lambda w: packing.pack_conv3x3(w, wd)
        pack_linear = lambda w: packing.pack_linear(w, wd)
        pack_geglu = lambda w, b, bn, return_perm=False: packing.pack_geglu(w, b, bn, return_perm, wd)
        boc = list(cfg.block_out_channels)
        self.temb_ch = boc[0] * 4
        temb_w, temb_b = [], []
        self._temb_off = 0
        kv_w = []
        self._kv_off = 0
        # diffusers module path -> (kind, getter of the packed tensor, first row, rows, row permutation): where a LoRA
        # delta B.A of that module lands in the packed weights (adapters below)
        self._wmap = {}

        def reg(mod, kind, getter, row0, nrows, perm=None):
            self._wmap[mod] = (kind, getter, row0, nrows, perm)

        def conv(prefix):
            ns = SimpleNamespace(w=pack_conv3x3(g(prefix + ".weight")).to(dev), b=_f32(g(prefix + ".bias"), dev))
            reg(prefix, "conv", lambda: ns.w, 0, ns.w.shape[0])
            return ns

        def lin(prefix, bias=True):
            ns = SimpleNamespace(w=pack_linear(g(prefix + ".weight")).to(dev),
                                 b=_f32(g(prefix + ".bias"), dev) if bias else None)
            reg(prefix, "linear", lambda: ns.w, 0, ns.w.shape[0])
            return ns

        def norm(prefix):
            return SimpleNamespace(g=_f32(g(prefix + ".weight"), dev), b=_f32(g(prefix + ".bias"), dev))

        def res(prefix, cin, cout):
            r = SimpleNamespace(cin=cin, cout=cout, norm1=norm(prefix + ".norm1"), conv1=conv(prefix + ".conv1"),
                                norm2=norm(prefix + ".norm2"), conv2=conv(prefix + ".conv2"), temb_off=self._temb_off,
                                shortcut=lin(prefix + ".conv_shortcut") if cin != cout else None)
            temb_w.append(g(prefix + ".time_emb_proj.weight"))
            temb_b.append(g(prefix + ".time_emb_proj.bias"))
            reg(prefix + ".time_emb_proj", "linear", lambda: self.temb_all.w, self._temb_off, cout)
            self._temb_off += cout
            return r

        def attn_block(prefix, C, heads, place):
            a1, a2 = prefix + ".attn1", prefix + ".attn2"
            w1, b1, perm = pack_geglu(g(prefix + ".ff.net.0.proj.weight").to(dev),
                                      g(prefix + ".ff.net.0.proj.bias").to(dev), GEGLU_BN, return_perm=True)
            blk = SimpleNamespace(
                ln1=norm(prefix + ".norm1"), ln2=norm(prefix + ".norm2"), ln3=norm(prefix + ".norm3"),
                qkv=torch.cat([pack_linear(g(a1 + f".to_{n}.weight")) for n in "qkv"], 0).to(dev),
                out1=lin(a1 + ".to_out.0"),
                q2=pack_linear(g(a2 + ".to_q.weight")).to(dev),
                kv_off=self._kv_off,     # column offset of this layer's [K | V] in the all-layers projection
                out2=lin(a2 + ".to_out.0"),
                ff1=SimpleNamespace(w=w1, b=b1), ff2=lin(prefix + ".ff.net.2"))
            kv_w.extend(pack_linear(g(a2 + f".to_{n}.weight")) for n in "kv")
            for i, n in enumerate("qkv"):
                reg(a1 + f".to_{n}", "linear", lambda: blk.qkv, i * C, C)
            reg(a2 + ".to_q", "linear", lambda: blk.q2, 0, C)
            reg(a2 + ".to_k", "linear", lambda: self.kv_all, self._kv_off, C)
            reg(a2 + ".to_v", "linear", lambda: self.kv_all, self._kv_off + C, C)
            reg(prefix + ".ff.net.0.proj", "linear", lambda: blk.ff1.w, 0, w1.shape[0], perm)
            self._kv_off += 2 * C
            self.attn_places += [place, place]
            return blk

        def tfm(prefix, C, heads, depth, place):
            return SimpleNamespace(C=C, heads=heads, d=C // heads, norm=norm(prefix + ".norm"),
                                   proj_in=lin(prefix + ".proj_in"), proj_out=lin(prefix + ".proj_out"),
                                   blocks=[attn_block(f"{prefix}.transformer_blocks.{k}", C, heads, place)
                                           for k in range(depth)])

        heads, depth = list(cfg.attention_head_dim), list(cfg.transformer_layers_per_block)
        lpb = cfg.layers_per_block
        self.conv_in = conv("conv_in")
        te = "time_embedding"
        self.time_lin1, self.time_lin2 = lin(te + ".linear_1"), lin(te + ".linear_2")
        self.cond_proj = (pack_linear(g(te + ".cond_proj.weight")).to(dev)
                          if getattr(cfg, "time_cond_proj_dim", None) else None)
        self.is_xl = getattr(cfg, "addition_embed_type", None) == "text_time"
        if self.is_xl:
            self.add_lin1, self.add_lin2 = lin("add_embedding.linear_1"), lin("add_embedding.linear_2")

        self.down = []
        out_ch = boc[0]
        for i, t in enumerate(cfg.down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            has_attn = t == "CrossAttnDownBlock2D"
            blk = SimpleNamespace(resnets=[], attns=[] if has_attn else None, down=None)
            for j in range(lpb):
                blk.resnets.append(res(f"down_blocks.{i}.resnets.{j}", in_ch if j == 0 else out_ch, out_ch))
                if has_attn:
                    blk.attns.append(tfm(f"down_blocks.{i}.attentions.{j}", out_ch, heads[i], depth[i], "down"))
            if i != len(boc) - 1:
                blk.down = conv(f"down_blocks.{i}.downsamplers.0.conv")
            self.down.append(blk)
        self.mid = SimpleNamespace(
            res0=res("mid_block.resnets.0", boc[-1], boc[-1]),
            attn=tfm("mid_block.attentions.0", boc[-1], heads[-1], depth[-1], "mid"),
            res1=res("mid_block.resnets.1", boc[-1], boc[-1]))
        self.up = []
        rboc, rheads, rdepth = boc[::-1], heads[::-1], depth[::-1]
        out_ch = rboc[0]
        for i, t in enumerate(cfg.up_block_types):
            prev, out_ch = out_ch, rboc[i]
            in_ch = rboc[min(i + 1, len(boc) - 1)]
            has_attn = t == "CrossAttnUpBlock2D"
            blk = SimpleNamespace(resnets=[], attns=[] if has_attn else None, up=None)
            for j in range(lpb + 1):
                skip = in_ch if j == lpb else out_ch
                rin = prev if j == 0 else out_ch
                blk.resnets.append(res(f"up_blocks.{i}.resnets.{j}", rin + skip, out_ch))
                if has_attn:
                    blk.attns.append(tfm(f"up_blocks.{i}.attentions.{j}", out_ch, rheads[i], rdepth[i], "up"))
            if i != len(boc) - 1:
                blk.up = conv(f"up_blocks.{i}.upsamplers.0.conv")
            self.up.append(blk)
        self.norm_out = norm("conv_norm_out")
        self.conv_out = conv("conv_out")
        # one GEMM for the to_k / to_v projections of every cross-attention layer (attn2, utils/p2p.py:331-333)
        self.kv_all = torch.cat(kv_w, 0).to(dev)
        # one GEMM for every ResnetBlock2D.time_emb_proj of the network
        self.temb_all = SimpleNamespace(w=torch.cat([pack_linear(w) for w in temb_w], 0).to(dev),
                                        b=torch.cat([_f32(b, dev) for b in temb_b], 0))

    # ------------------------------------------------------------------ hot-swappable LoRA adapters
    # The reference keeps three full U-Nets resident (teacher + two deep copies with a LoRA fused in,
    # utils/loading.py:63-88,116-146). Alternative offered here: ONE packed U-Net, the base copy of the
    # adapted matrices, and per adapter the low-rank factors pre-arranged in the packed layout. Switching adapters
    # re-materialises  W = W_base + (alpha/r) * B.A  on the GPU: one tcgen05 GEMM per adapted module (M = C_out,
    # N = packed C_in, K = r) whose epilogue adds the base weights (the residual is accumulated on the tensor core in
    # fp32) and rounds to fp16 once — the same arithmetic as `loading.fuse_lora` up to fp32 summation order. The working
    # tensors are updated in place, so captured CUDA graphs (which hold their addresses) stay valid.
    def add_adapter(self, name, lora_weights, lora_dtype=torch.float16, alpha=8):
        """Register a peft-format LoRA state dict (`unet.base_model.model.<module>.lora_{A,B}.weight`,
        utils/loading.py:10-23) under `name`. A and B are first cast to `lora_dtype` as the reference does (fp16 for
        SD1.5, utils/loading.py:68; fp32 for SDXL, :122), then rounded to the fp16 operands of the fuse GEMM."""
        prefix = "unet.base_model.model."
        mods = sorted({k[len(prefix):].rsplit(".lora_", 1)[0] for k in lora_weights if k.startswith(prefix)})
        if not mods:
            raise ValueError("add_adapter: no peft-format keys ('unet.base_model.model.*.lora_A.weight') found")
        if not hasattr(self, "_adapters"):
            self._adapters, self._base_copy, self._active = {}, {}, None
        dev, packed = self.device, {}
        for mod in mods:
            if mod not in self._wmap:
                raise KeyError(f"LoRA adapter for unknown module {mod}")
            kind, getter, row0, nrows, perm = self._wmap[mod]
            A = lora_weights[f"{prefix}{mod}.lora_A.weight"].to(lora_dtype).float().to(dev)
            Bm = lora_weights[f"{prefix}{mod}.lora_B.weight"].to(lora_dtype).float().to(dev).flatten(1)
            rank = A.shape[0]
            if Bm.shape[0] != nrows:
                raise ValueError(f"LoRA B of {mod}: {Bm.shape[0]} rows, expected {nrows}")
            if perm is not None:
                Bm = Bm[perm]
            Bp = (Bm * (alpha / rank)).to(self.act_dtype).contiguous()                # [C_out, r]
            Ap = (packing.pack_conv3x3 if kind == "conv" else packing.pack_linear)(A, self.act_dtype)   # [r, packed C_in]
            if Ap.shape[1] != getter().shape[1]:
                raise ValueError(f"LoRA A of {mod}: {Ap.shape[1]} packed columns, expected {getter().shape[1]}")
            packed[mod] = (Bp, Ap.t().contiguous())                                   # GEMM "weight" operand [N, K=r]
            w = getter()
            if w.data_ptr() not in self._base_copy:
                self._base_copy[w.data_ptr()] = w.clone()
        self._adapters[name] = packed

    def set_adapter(self, name):
        """Make adapter `name` (None = the base model) the one the packed weights contain. Launches only; no sync."""
        if not hasattr(self, "_adapters"):
            if name is None:
                return
            raise KeyError(f"no adapter '{name}' (none registered)")
        if name == self._active:
            return
        if name is not None and name not in self._adapters:
            raise KeyError(f"no adapter '{name}' (have {sorted(self._adapters)})")
        if _capturing():
            raise RuntimeError("set_adapter inside a CUDA-graph capture: activate the adapter before capturing")
        new = self._adapters[name] if name is not None else {}
        old = self._adapters[self._active] if self._active is not None else {}
        with torch.cuda.device(self.device):
            for mod in sorted(set(new) | set(old)):
                _, getter, row0, nrows, _ = self._wmap[mod]
                w = getter()
                work, base = w[row0:row0 + nrows], self._base_copy[w.data_ptr()][row0:row0 + nrows]
                if mod in new:
                    Bp, At = new[mod]
                    self.ops.linear(Bp, At, residual=base, out=work)
                else:
                    work.copy_(base)
        self._active = name

    @property
    def active_adapter(self):
        return getattr(self, "_active", None)

    def adapter_view(self, name):
        """A U-Net-like handle that runs this executor with adapter `name` active (see AdapterView)."""
        return AdapterView(self, name)

    # ------------------------------------------------------------------ embeddings
    def _freqs(self, kind, dim):
        key = (kind, dim)
        if key not in self._freq_cache:
            half = dim // 2
            if kind == "t":     # diffusers get_timestep_embedding(freq_shift=0), fp32 op order
                f = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half)
            else:               # guidance_scale_embedding, utils/generation.py:114-116 (fp32 op order)
                e = torch.log(torch.tensor(10000.0)) / (half - 1)
                f = torch.exp(torch.arange(half, dtype=torch.float32) * -e)
            self._freq_cache[key] = f.to(self.device)
        return self._freq_cache[key]

    def cached_vector(self, values):
        """Small fp32 device vector, uploaded once per distinct value tuple (keeps the step free of H2D copies)."""
        key = ("vec",) + tuple(float(v) for v in values)
        if key not in self._freq_cache:
            self._freq_cache[key] = torch.tensor([float(v) for v in values], dtype=torch.float32, device=self.device)
        return self._freq_cache[key]

    def guidance_embedding(self, w, dim=512):
        """w: fp32 [rows] device tensor -> fp16 [rows, dim] (utils/generation.py:96-122)."""
        with torch.cuda.device(self.device):
            return self.ops.guidance_embedding(w, self._freqs("w", dim), dim)

    def _time_embed(self, rows, timestep, timestep_cond, added):
        dev, ops = self.device, self.ops
        if torch.is_tensor(timestep):
            t = timestep.to(device=dev, dtype=torch.float32).reshape(-1)
            if t.numel() == 1:
                t = t.expand(rows)
            t = t.contiguous()
        else:
            t = torch.full((rows,), float(timestep), device=dev, dtype=torch.float32)
        c0 = self.config.block_out_channels[0]
        t_emb = ops.timestep_embedding(t, self._freqs("t", c0), c0)
        if timestep_cond is not None:
            if self.cond_proj is None:
                raise ValueError("timestep_cond given but the model has no time_cond_proj_dim")
            cond = timestep_cond.to(device=dev, dtype=self.act_dtype).contiguous()
            t_emb = ops.linear(cond, self.cond_proj, residual=t_emb)        # t_emb + cond_proj(w_emb)
        h = ops.silu(ops.linear(t_emb, self.time_lin1.w, bias=self.time_lin1.b))
        emb = ops.linear(h, self.time_lin2.w, bias=self.time_lin2.b)
        if self.is_xl:
            text_embeds = added["text_embeds"].to(device=dev, dtype=self.act_dtype)
            time_ids = added["time_ids"].to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
            d = self.config.addition_time_embed_dim
            tid = ops.timestep_embedding(time_ids, self._freqs("t", d), d).reshape(rows, -1)
            add_in = torch.cat([text_embeds, tid], dim=-1).contiguous()
            h = ops.silu(ops.linear(add_in, self.add_lin1.w, bias=self.add_lin1.b))
            emb = ops.linear(h, self.add_lin2.w, bias=self.add_lin2.b, residual=emb)   # emb + aug_emb
        # every ResnetBlock2D consumes time_emb_proj(silu(emb)): one GEMM for all of them, fp32 out
        return ops.linear(ops.silu(emb), self.temb_all.w, bias=self.temb_all.b, out_fp32=True)

    # ------------------------------------------------------------------ blocks
    def _res(self, r, x0, x1, B, H, W):
        """ResnetBlock2D over channels-last tokens; (x0 | x1) is the virtual channel concat of the skip."""
        ws, ops = self._gn_ws, self.ops
        h = ops.groupnorm(x0, B, H * W, r.norm1.g, r.norm1.b, 1e-5, True, ws, x1=x1)
        temb = self._temb[:, r.temb_off:r.temb_off + r.cout]
        h = ops.conv3x3(h, r.conv1.w, B, H, W, bias=r.conv1.b, rowvec=temb)
        h = ops.groupnorm(h, B, H * W, r.norm2.g, r.norm2.b, 1e-5, True, ws)
        if r.shortcut is not None:
            sc = ops.linear(x0, r.shortcut.w, bias=r.shortcut.b, a1=x1)
        else:
            sc = x0
        return ops.conv3x3(h, r.conv2.w, B, H, W, bias=r.conv2.b, residual=sc)

    def _attention(self, q, k, v, B, heads, Nq, Nk, d, is_cross, place):
        """Attention core + the p2p controller protocol (utils/p2p.py:335-338)."""
        scale = d ** -0.5
        ctrl, ops, fp32 = self.controller, self.ops, self.precision == "fp32"
        req = "none" if ctrl is None else ctrl.probs_request(is_cross, place, Nq, Nk)
        if req == "none":
            out = ops.attention(q, k, v, B, heads, Nq, Nk, d, scale)
            if ctrl is not None:
                ctrl.layer_skipped(is_cross, place)
            return out
        ldp = (Nk + 7) // 8 * 8
        if req == "read" and Nk <= 128:
            # the kernel writes every padded row completely (pad columns are zero): no memset needed
            probs = torch.empty((B * heads, Nq, ldp), device=q.device, dtype=self.act_dtype)
            out = ops.attention(q, k, v, B, heads, Nq, Nk, d, scale, probs_out=probs)
            ctrl.call_rows(probs[..., :Nk], is_cross, place, self._cond_only)
            return out
        if req == "read" and Nk % 8 == 0 and not fp32:
            # larger read-only maps (self-attention, N_q <= 1024): the fused kernel produces the output and the
            # online-softmax statistics; one more pass over Q.K^T writes the normalised probabilities directly
            stats = torch.empty((B * heads, Nq, 2), device=q.device, dtype=torch.float32)
            out = ops.attention(q, k, v, B, heads, Nq, Nk, d, scale, stats_out=stats)
            probs = torch.empty((B * heads, Nq, Nk), device=q.device, dtype=torch.float16)
            ops.attn_probs_from_stats(q, k, B, heads, Nq, Nk, d, scale, stats, probs)
            ctrl.call_rows(probs, is_cross, place, self._cond_only)
            return out
        # explicit probabilities: scores GEMM -> softmax -> controller (may edit in place) -> P.V GEMM
        probs = torch.zeros((B * heads, Nq, ldp), device=q.device, dtype=self.act_dtype)
        ops.attn_scores(q, k, B, heads, Nq, Nk, d, scale, probs)
        ops.softmax_(probs, Nk)
        view = probs[..., :Nk]
        edited = ctrl.call_rows(view, is_cross, place, self._cond_only)
        if edited.data_ptr() != view.data_ptr() or edited.stride() != view.stride():
            view.copy_(edited)           # controller returned a new tensor: bring it back into the padded buffer
        out = torch.empty((B * Nq, heads * d), device=q.device, dtype=self.act_dtype)
        ops.attn_pv(probs, v, B, heads, Nq, Nk, d, out)
        return out

    def _tfm(self, t, x, ctx, B, HW, place):
        C, heads, d, ops = t.C, t.heads, t.d, self.ops
        n_ctx = ctx.shape[0] // B
        h = ops.groupnorm(x, B, HW, t.norm.g, t.norm.b, 1e-6, False, self._gn_ws)
        h = ops.linear(h, t.proj_in.w, bias=t.proj_in.b)
        for blk in t.blocks:
            n = ops.layernorm(h, blk.ln1.g, blk.ln1.b)
            qkv = ops.linear(n, blk.qkv)
            a = self._attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, heads, HW, HW, d, False, place)
            h = ops.linear(a, blk.out1.w, bias=blk.out1.b, residual=h)
            n = ops.layernorm(h, blk.ln2.g, blk.ln2.b)
            q = ops.linear(n, blk.q2)
            kv = self._kv[:, blk.kv_off:blk.kv_off + 2 * C]
            a = self._attention(q, kv[:, :C], kv[:, C:], B, heads, HW, n_ctx, d, True, place)
            h = ops.linear(a, blk.out2.w, bias=blk.out2.b, residual=h)
            n = ops.layernorm(h, blk.ln3.g, blk.ln3.b)
            gg = ops.linear(n, blk.ff1.w, bias=blk.ff1.b, geglu=True, force_bn=GEGLU_BN)
            h = ops.linear(gg, blk.ff2.w, bias=blk.ff2.b, residual=h)
        return ops.linear(h, t.proj_out.w, bias=t.proj_out.b, residual=x)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, *args, **kwargs):
        """See `_forward`. Runs with this U-Net's device current: every launch goes to torch's current stream OF
        THAT DEVICE (`load_models(model_id, 'cuda:1', ...)` works without a prior torch.cuda.set_device)."""
        with torch.cuda.device(self.device):
            return self._forward(*args, **kwargs)

    def _forward(self, sample, timestep, encoder_hidden_states=None, timestep_cond=None, added_cond_kwargs=None,
                 cross_attention_kwargs=None, return_dict=True, cond_only=False, update=None):
        """sample: [rows, 4, H, W] (any float dtype; computed in fp16) -> eps fp32 [rows, 4, H, W].
        `cond_only`: the rows are the conditional half only (controller sees them all).
        `update`: optional (x_t fp32 NCHW, alpha_t, sigma_t, alpha_s, sigma_s): fuse predicted_origin into the
        conv_out epilogue; the next latent is returned as out["next_sample"]."""
        dev, ops = self.device, self.ops
        rows, _, H, W = sample.shape
        lat = sample.to(device=dev, dtype=torch.float32).contiguous()
        ctx = encoder_hidden_states.to(device=dev, dtype=self.act_dtype)
        if ctx.shape[0] != rows:
            raise ValueError(f"encoder_hidden_states rows {ctx.shape[0]} != sample rows {rows}")
        ctx = ctx.reshape(rows * ctx.shape[1], ctx.shape[2]).contiguous()
        self._cond_only = cond_only
        self._gn_ws = torch.empty(rows * 4096, device=dev, dtype=torch.float32)
        self._temb = self._time_embed(rows, timestep, timestep_cond, added_cond_kwargs)
        self._kv = ops.linear(ctx, self.kv_all)     # [rows*77, sum over layers of 2C]: K | V of every attn2

        x = ops.conv3x3(ops.latent_to_nhwc(lat, cpad=8), self.conv_in.w, rows, H, W, bias=self.conv_in.b)
        skips = [(x, H, W)]
        for blk in self.down:
            for j, r in enumerate(blk.resnets):
                x = self._res(r, x, None, rows, H, W)
                if blk.attns is not None:
                    x = self._tfm(blk.attns[j], x, ctx, rows, H * W, "down")
                skips.append((x, H, W))
            if blk.down is not None:
                x = ops.linear(ops.im2col_s2(x, rows, H, W), blk.down.w, bias=blk.down.b)
                H, W = H // 2, W // 2
                skips.append((x, H, W))
        x = self._res(self.mid.res0, x, None, rows, H, W)
        x = self._tfm(self.mid.attn, x, ctx, rows, H * W, "mid")
        x = self._res(self.mid.res1, x, None, rows, H, W)
        for blk in self.up:
            for j, r in enumerate(blk.resnets):
                skip, sh, sw = skips.pop()
                assert (sh, sw) == (H, W)
                x = self._res(r, x, skip, rows, H, W)
                if blk.attns is not None:
                    x = self._tfm(blk.attns[j], x, ctx, rows, H * W, "up")
            if blk.up is not None:
                x = ops.conv3x3(ops.upsample2x(x, rows, H, W), blk.up.w, rows, 2 * H, 2 * W, bias=blk.up.b)
                H, W = 2 * H, 2 * W
        x = ops.groupnorm(x, rows, H * W, self.norm_out.g, self.norm_out.b, 1e-5, True, self._gn_ws)
        eps = torch.empty((rows, self.config.out_channels, H, W), device=dev, dtype=torch.float32)
        nxt = None
        if update is not None:
            x_t, a_t, s_t, a_s, s_s = update
            nxt = torch.empty_like(eps)
            ops.conv3x3(x, self.conv_out.w, rows, H, W, bias=self.conv_out.b, nchw_out=eps,
                        upd_x=x_t, upd_out=nxt, upd_coefs=(a_t, s_t, a_s, s_s))
        else:
            ops.conv3x3(x, self.conv_out.w, rows, H, W, bias=self.conv_out.b, nchw_out=eps)
        self._temb = self._kv = None
        if not return_dict:
            return (eps,) if nxt is None else (eps, nxt)
        out = UNetOutput(sample=eps)
        if nxt is not None:
            out["next_sample"] = nxt
        return out

    __call__ = forward


class AdapterView:
    """`pipeline.unet` of one model (teacher / forward / reverse consistency student) when several adapters share a
    single B200UNet (`loading.load_models(..., adapters="swap")`). Every call activates its adapter first (a no-op
    when it already is the active one), then runs the shared executor with this view's controller. Owns its own
    CUDA-graph cache (graphs.py keys the cache by owner object); graphs captured for one view replay correctly for it
    because the packed weights are rewritten in place."""

    def __init__(self, shared, name):
        self.__dict__["_shared"] = shared
        self.__dict__["_name"] = name
        self.__dict__["controller"] = None

    supports_cond_only = True

    def activate(self):
        self._shared.set_adapter(self._name)

    def __getattr__(self, k):           # config, device, dtype, guidance_embedding, cached_vector, attn_places, ...
        return getattr(self.__dict__["_shared"], k)

    def __setattr__(self, k, v):
        if k in ("controller", "_icd_graphs"):
            self.__dict__[k] = v
        else:
            setattr(self._shared, k, v)

    def forward(self, *args, **kwargs):
        shared = self._shared
        if not _capturing():
            self.activate()
        elif shared.active_adapter != self._name:
            raise RuntimeError("AdapterView called inside a CUDA-graph capture with another adapter active")
        saved = shared.controller
        shared.controller = self.controller
        try:
            return shared.forward(*args, **kwargs)
        finally:
            shared.controller = saved

    __call__ = forward

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def named_children(self):
        return iter(())
