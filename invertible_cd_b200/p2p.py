"""Prompt-to-prompt attention control for the B200 U-Net — the plugin surface of utils/p2p.py.

Same public names, call protocol and module-level knobs as the reference (`NUM_DDIM_STEPS`, `tokenizer`,
`device`, `LOW_RESOURCE`, `MAX_NUM_WORDS`, utils/p2p.py:9-13): a controller is called once per Attention layer in
execution order with the probabilities `(rows*heads, N_q, N_kv)` and may edit them in place; `step_callback` runs
once per sampling step.  Controllers stay Python (they are stateful, aliasing, order-dependent host logic); what
changes is *who produces the probabilities*: the reference monkey-patches every diffusers `Attention.forward`
with an explicit softmax (utils/p2p.py:291-386) and so materialises every map, including the five 4096x4096 maps
per U-Net row that no controller keeps; here `register_attention_control` hands the controller to the CUDA
executor, which asks it per layer what it needs (`probs_request`) and
    'none' -> runs the fused tcgen05 attention kernel and only advances the controller's layer counter,
    'read' -> has the kernel write the normalised cross-attention map in the same pass (AttentionStore capture),
    'edit' -> runs scores-GEMM -> softmax -> controller -> P.V-GEMM with the probabilities materialised.
Custom controllers that subclass `AttentionControl` and only override `forward` get 'edit' everywhere, i.e. the
reference's semantics.
"""
import abc
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as nnf

from . import seq_aligner

MAX_NUM_WORDS = 77
LOW_RESOURCE = False
NUM_DDIM_STEPS = 50
device = "cuda"
tokenizer = None

_STORE_MAX_QUERIES = 32 ** 2      # utils/p2p.py:147 "avoid memory overhead"


# ------------------------------------------------------------------------------------------------ local blend
class LocalBlend:
    """Latent-space blending from the accumulated 16x16 cross-attention maps (utils/p2p.py:18-70)."""

    def __init__(self, prompts: List[str], words, substruct_words=None, start_blend=0.2, th=(.3, .3)):
        self.alpha_layers = self._word_mask(prompts, words).to(device)
        self.substruct_layers = (self._word_mask(prompts, substruct_words).to(device)
                                 if substruct_words is not None else None)
        self.start_blend = int(start_blend * NUM_DDIM_STEPS)
        self.counter = 0
        self.th = th

    @staticmethod
    def _word_mask(prompts, words):
        mask = torch.zeros(len(prompts), 1, 1, 1, 1, MAX_NUM_WORDS)
        for row, (prompt, ws) in enumerate(zip(prompts, words)):
            for word in ([ws] if type(ws) is str else ws):
                mask[row, :, :, :, :, get_word_inds(prompt, word, tokenizer)] = 1
        return mask

    def get_mask(self, maps, alpha, use_pool, x_t):
        k = 1
        m = (maps * alpha).sum(-1).mean(1)
        if use_pool:
            m = nnf.max_pool2d(m, (2 * k + 1, 2 * k + 1), (1, 1), padding=(k, k))
        m = nnf.interpolate(m, size=(x_t.shape[2:]))
        m = m / m.max(2, keepdims=True)[0].max(3, keepdims=True)[0]
        m = m.gt(self.th[1 - int(use_pool)])
        return m[:1] + m                     # bool '+' == logical OR with the source row's mask

    def __call__(self, x_t, attention_store):
        self.counter += 1
        if self.counter <= self.start_blend:
            return x_t
        picked = attention_store["down_cross"][2:4] + attention_store["up_cross"][:3]
        n = self.alpha_layers.shape[0]
        maps = torch.cat([m.reshape(n, -1, 1, 16, 16, MAX_NUM_WORDS) for m in picked], dim=1)
        mask = self.get_mask(maps, self.alpha_layers, True, x_t)
        if self.substruct_layers is not None:
            mask = mask * ~self.get_mask(maps, self.substruct_layers, False, x_t)
        mask = mask.float()
        return x_t[:1] + mask * (x_t - x_t[:1])


# ------------------------------------------------------------------------------------------------ controllers
class EmptyControl:

    def step_callback(self, x_t):
        return x_t

    def between_steps(self):
        return

    def __call__(self, attn, is_cross: bool, place_in_unet: str):
        return attn

    # executor protocol
    def probs_request(self, is_cross, place_in_unet, n_query, n_key):
        return "none"

    def layer_skipped(self, is_cross, place_in_unet):
        return

    def call_rows(self, attn, is_cross, place_in_unet, cond_only):
        return attn


class AttentionControl(abc.ABC):

    def __init__(self):
        self.cur_step = 0
        self.num_att_layers = -1
        self.cur_att_layer = 0

    def step_callback(self, x_t):
        return x_t

    def between_steps(self):
        return

    @property
    def num_uncond_att_layers(self):
        return self.num_att_layers if LOW_RESOURCE else 0

    @abc.abstractmethod
    def forward(self, attn, is_cross: bool, place_in_unet: str):
        raise NotImplementedError

    def _advance(self):
        self.cur_att_layer += 1
        if self.cur_att_layer == self.num_att_layers + self.num_uncond_att_layers:
            self.cur_att_layer = 0
            self.cur_step += 1
            self.between_steps()

    def __call__(self, attn, is_cross: bool, place_in_unet: str):
        """Reference protocol (utils/p2p.py:101-113): rows are [uncond..., cond...]; only the cond half is shown
        to `forward`, whose result is written back in place."""
        if self.cur_att_layer >= self.num_uncond_att_layers:
            if LOW_RESOURCE:
                attn = self.forward(attn, is_cross, place_in_unet)
            else:
                h = attn.shape[0]
                attn[h // 2:] = self.forward(attn[h // 2:], is_cross, place_in_unet)
        self._advance()
        return attn

    def reset(self):
        self.cur_step = 0
        self.cur_att_layer = 0

    # ---- executor protocol -----------------------------------------------------------------------
    def probs_request(self, is_cross, place_in_unet, n_query, n_key):
        """What this controller needs from the layer about to run: 'none' | 'read' | 'edit'."""
        return "edit"

    def layer_skipped(self, is_cross, place_in_unet):
        """The layer ran fused (no probabilities materialised): keep the layer/step bookkeeping in sync."""
        self._advance()

    def call_rows(self, attn, is_cross, place_in_unet, cond_only):
        """`cond_only`: the U-Net ran only the conditional rows (w-embedded models discard the uncond half,
        utils/generation.py:245-251), so every row goes to `forward`."""
        if not cond_only:
            return self(attn, is_cross, place_in_unet)
        if self.cur_att_layer >= self.num_uncond_att_layers:
            attn = self.forward(attn, is_cross, place_in_unet)
        self._advance()
        return attn


class SpatialReplace(EmptyControl):

    def __init__(self, stop_inject: float):
        super().__init__()
        self.stop_inject = int((1 - stop_inject) * NUM_DDIM_STEPS)
        self.cur_step = 0

    def step_callback(self, x_t):
        if self.cur_step < self.stop_inject:
            x_t = x_t[:1].expand(x_t.shape[0], *x_t.shape[1:])
        return x_t


class AttentionStore(AttentionControl):

    # Set to False to skip materialising the self-attention maps (N_q <= 32^2): nothing in the reference consumes
    # them (LocalBlend reads five cross maps only, utils/p2p.py:35-37) and they are 11x the volume of the cross
    # maps. Default True = reference behaviour (utils/p2p.py:145-149).
    capture_self = True

    def __init__(self):
        super().__init__()
        self.step_store = self.get_empty_store()
        self.attention_store = {}

    @staticmethod
    def get_empty_store():
        return {f"{place}_{kind}": [] for kind in ("cross", "self") for place in ("down", "mid", "up")}

    def forward(self, attn, is_cross: bool, place_in_unet: str):
        if attn.shape[1] <= _STORE_MAX_QUERIES:
            self.step_store[f"{place_in_unet}_{'cross' if is_cross else 'self'}"].append(attn)
        return attn

    def between_steps(self):
        if not self.attention_store:
            self.attention_store = self.step_store
        else:
            # in-place accumulation as utils/p2p.py:153-156; on the GPU all maps of the step go through ONE
            # multi-tensor add instead of one launch per map (same fp16 additions, same results)
            dst = [m for key, maps in self.attention_store.items() for m in maps]
            src = [self.step_store[key][i] for key, maps in self.attention_store.items() for i in range(len(maps))]
            if dst and all(t.is_cuda for t in dst):
                torch._foreach_add_(dst, src)
            else:
                for d, s_ in zip(dst, src):
                    d += s_
        self.step_store = self.get_empty_store()

    def get_average_attention(self):
        return {key: [m / self.cur_step for m in maps] for key, maps in self.attention_store.items()}

    def reset(self):
        super().reset()
        self.step_store = self.get_empty_store()
        self.attention_store = {}

    def probs_request(self, is_cross, place_in_unet, n_query, n_key):
        if type(self).forward is not AttentionStore.forward and not isinstance(self, AttentionControlEdit):
            return "edit"       # user subclass with its own forward: reference semantics
        if not is_cross and not self.capture_self:
            return "none"
        return "read" if n_query <= _STORE_MAX_QUERIES else "none"


class AttentionControlEdit(AttentionStore, abc.ABC):

    def __init__(self, prompts, num_steps: int,
                 cross_replace_steps: Union[float, Tuple[float, float], Dict[str, Tuple[float, float]]],
                 self_replace_steps: Union[float, Tuple[float, float]], local_blend: Optional[LocalBlend]):
        super().__init__()
        self.batch_size = len(prompts)
        alpha = get_time_words_attention_alpha(prompts, num_steps, cross_replace_steps, tokenizer)
        self._cross_active = [bool(a.any()) for a in alpha]     # host-side, avoids a device sync per layer
        self.cross_replace_alpha = alpha.to(device)
        if type(self_replace_steps) is float:
            self_replace_steps = 0, self_replace_steps
        self.num_self_replace = int(num_steps * self_replace_steps[0]), int(num_steps * self_replace_steps[1])
        self.local_blend = local_blend

    def step_callback(self, x_t):
        if self.local_blend is not None:
            x_t = self.local_blend(x_t, self.attention_store)
        return x_t

    def replace_self_attention(self, attn_base, att_replace, place_in_unet):
        if att_replace.shape[2] <= _STORE_MAX_QUERIES:
            return attn_base.unsqueeze(0).expand(att_replace.shape[0], *attn_base.shape)
        return att_replace

    @abc.abstractmethod
    def replace_cross_attention(self, attn_base, att_replace):
        raise NotImplementedError

    def _self_window(self):
        return self.num_self_replace[0] <= self.cur_step < self.num_self_replace[1]

    def forward(self, attn, is_cross: bool, place_in_unet: str):
        super().forward(attn, is_cross, place_in_unet)      # stores a VIEW: later edits show up in the store
        if is_cross or self._self_window():
            heads = attn.shape[0] // self.batch_size
            attn = attn.reshape(self.batch_size, heads, *attn.shape[1:])
            base, repl = attn[0], attn[1:]
            if is_cross:
                a = self.cross_replace_alpha[self.cur_step]
                attn[1:] = self.replace_cross_attention(base, repl) * a + (1 - a) * repl
            else:
                attn[1:] = self.replace_self_attention(base, repl, place_in_unet)
            attn = attn.reshape(self.batch_size * heads, *attn.shape[2:])
        return attn

    def probs_request(self, is_cross, place_in_unet, n_query, n_key):
        small = n_query <= _STORE_MAX_QUERIES
        if is_cross:
            step = min(self.cur_step, len(self._cross_active) - 1)
            if self._cross_active[step]:
                return "edit"
            return "read" if small else "none"      # alpha == 0: the edit is the identity
        if small:
            return "edit" if self._self_window() else "read"
        return "none"                               # large self maps are neither stored nor replaced


class AttentionReplace(AttentionControlEdit):

    def __init__(self, prompts, num_steps: int, cross_replace_steps: float, self_replace_steps: float,
                 local_blend: Optional[LocalBlend] = None):
        super().__init__(prompts, num_steps, cross_replace_steps, self_replace_steps, local_blend)
        self.mapper = seq_aligner.get_replacement_mapper(prompts, tokenizer).to(device)

    def replace_cross_attention(self, attn_base, att_replace):
        return torch.einsum('hpw,bwn->bhpn', attn_base, self.mapper.to(attn_base.dtype))


class AttentionRefine(AttentionControlEdit):

    def __init__(self, prompts, num_steps: int, cross_replace_steps: float, self_replace_steps: float,
                 local_blend: Optional[LocalBlend] = None):
        super().__init__(prompts, num_steps, cross_replace_steps, self_replace_steps, local_blend)
        mapper, alphas = seq_aligner.get_refinement_mapper(prompts, tokenizer)
        self.mapper = mapper.to(device)
        self.alphas = alphas.to(device).reshape(alphas.shape[0], 1, 1, alphas.shape[1])

    def replace_cross_attention(self, attn_base, att_replace):
        gathered = attn_base[:, :, self.mapper].permute(2, 0, 1, 3)
        return gathered * self.alphas + att_replace * (1 - self.alphas)


class AttentionReweight(AttentionControlEdit):

    def __init__(self, prompts, num_steps: int, cross_replace_steps: float, self_replace_steps: float, equalizer,
                 local_blend: Optional[LocalBlend] = None, controller: Optional[AttentionControlEdit] = None):
        super().__init__(prompts, num_steps, cross_replace_steps, self_replace_steps, local_blend)
        self.equalizer = equalizer.to(device)
        self.prev_controller = controller
        self.attn = []

    def replace_cross_attention(self, attn_base, att_replace):
        if self.prev_controller is not None:
            attn_base = self.prev_controller.replace_cross_attention(attn_base, att_replace)
        return attn_base[None, :, :, :] * self.equalizer[:, None, None, :]


# ------------------------------------------------------------------------------------------------ factories
def make_controller(prompts: List[str], is_replace_controller: bool, cross_replace_steps: Dict[str, float],
                    self_replace_steps: float, blend_words=None, equilizer_params=None) -> AttentionControlEdit:
    lb = None if blend_words is None else LocalBlend(prompts, blend_words, start_blend=0.0, th=(0.3, 0.3))
    cls = AttentionReplace if is_replace_controller else AttentionRefine
    controller = cls(prompts, NUM_DDIM_STEPS, cross_replace_steps=cross_replace_steps,
                     self_replace_steps=self_replace_steps, local_blend=lb)
    if equilizer_params is not None:
        eq = get_equalizer(prompts[1], equilizer_params["words"], equilizer_params["values"])
        controller = AttentionReweight(prompts, NUM_DDIM_STEPS, cross_replace_steps=cross_replace_steps,
                                       self_replace_steps=self_replace_steps, equalizer=eq, local_blend=lb,
                                       controller=controller)
    return controller


def register_attention_control(model, controller):
    """Attach `controller` to `model.unet` (utils/p2p.py:291-386 patches 32 Attention.forward methods instead).
    `controller=None` detaches (the reference installs a pass-through DummyController)."""
    unet = model.unet
    if not hasattr(unet, "num_attention_layers"):
        raise TypeError("register_attention_control expects a pipeline whose .unet is an invertible_cd_b200 "
                        f"B200UNet, got {type(unet).__name__}")
    unet.controller = controller
    if controller is not None:
        controller.num_att_layers = unet.num_attention_layers


def get_equalizer(text: str, word_select: Union[int, Tuple[int, ...]],
                  values: Union[List[float], Tuple[float, ...]]):
    if type(word_select) is int or type(word_select) is str:
        word_select = (word_select,)
    equalizer = torch.ones(1, 77)
    for word, val in zip(word_select, values):
        equalizer[:, get_word_inds(text, word, tokenizer)] = val
    return equalizer


def update_alpha_time_word(alpha, bounds: Union[float, Tuple[float, float]], prompt_ind: int,
                           word_inds: Optional[torch.Tensor] = None):
    if type(bounds) is float:
        bounds = 0, bounds
    start, end = int(bounds[0] * alpha.shape[0]), int(bounds[1] * alpha.shape[0])
    if word_inds is None:
        word_inds = torch.arange(alpha.shape[2])
    alpha[:start, prompt_ind, word_inds] = 0
    alpha[start:end, prompt_ind, word_inds] = 1
    alpha[end:, prompt_ind, word_inds] = 0
    return alpha


def get_time_words_attention_alpha(prompts, num_steps,
                                   cross_replace_steps: Union[float, Dict[str, Tuple[float, float]]],
                                   tokenizer, max_num_words=77):
    if type(cross_replace_steps) is not dict:
        cross_replace_steps = {"default_": cross_replace_steps}
    if "default_" not in cross_replace_steps:
        cross_replace_steps["default_"] = (0., 1.)
    n_edit = len(prompts) - 1
    alpha = torch.zeros(num_steps + 1, n_edit, max_num_words)
    for i in range(n_edit):
        alpha = update_alpha_time_word(alpha, cross_replace_steps["default_"], i)
    for word, bounds in cross_replace_steps.items():
        if word == "default_":
            continue
        for i in range(n_edit):
            ind = get_word_inds(prompts[i + 1], word, tokenizer)
            if len(ind) > 0:
                alpha = update_alpha_time_word(alpha, bounds, i, ind)
    return alpha.reshape(num_steps + 1, n_edit, 1, 1, max_num_words)


def get_word_inds(text: str, word_place, tokenizer):
    return seq_aligner.get_word_inds(text, word_place, tokenizer)
