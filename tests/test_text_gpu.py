"""CLIP text encoder parity (GPU): B200CLIPTextModel (sm_100a kernels through the C ABI) vs
transformers' own CLIPTextModel / CLIPTextModelWithProjection — the third-party classes the reference itself calls
(utils/generation.py:286-303, utils/generation_sdxl.py:9-46), random-init, fp32 on the CPU, same weights and token ids.
Then the reference's call sites: Generator.init_prompt with prompt strings, generation_sdxl.encode_prompt."""
import json
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

# (rel-L2, max |err| / max|ref|) gates = 2x measured (profiles/r2_parity_report.jsonl)
# measured: rel-L2 0.92e-3 .. 1.09e-3, max 0.97e-3 .. 2.08e-3
GATES = {"clip_l_last": (2.2e-3, 3.4e-3), "clip_l_penultimate": (2.1e-3, 4.2e-3), "clip_l_pooled": (2.1e-3, 2e-3),
         "clip_g_text_embeds": (2e-3, 2e-3), "clip_g_penultimate": (1.9e-3, 4e-3)}


def _report(name, got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    diff = (got - ref).abs()
    rec = {"name": name, "shape": list(ref.shape), "max_abs": diff.max().item(), "ref_absmax": ref.abs().max().item(),
           "rel_l2": ((got - ref).norm() / ref.norm()).item(),
           "viol_rtol1e-3_atol1e-4": int((diff > 1e-4 + 1e-3 * ref.abs()).sum().item()), "numel": ref.numel()}
    rec["max_rel_to_absmax"] = rec["max_abs"] / max(rec["ref_absmax"], 1e-30)
    print("[parity]", json.dumps(rec))
    try:
        with open(os.path.join(ROOT, "gpurun_out", "parity_report.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
    assert torch.isfinite(got).all(), name
    assert rec["rel_l2"] <= GATES[name][0] and rec["max_rel_to_absmax"] <= GATES[name][1], rec


def _ids(B, vocab, eos, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1, eos - 1, (B, 77), generator=g)
    ids[:, 0] = eos - 1                                   # BOS
    for b in range(B):
        n = int(torch.randint(3, 70, (1,), generator=g))
        ids[b, n:] = eos                                  # EOS, then padded with EOS like CLIPTokenizer does
    return ids


def _hf(cfg_kw, with_projection, seed):
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection
    torch.manual_seed(seed)
    hcfg = CLIPTextConfig(**cfg_kw)
    m = (CLIPTextModelWithProjection if with_projection else CLIPTextModel)(hcfg).eval()
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(p.half().float())
    return m


def test_clip_l_matches_transformers():
    from invertible_cd_b200.text_encoder import B200CLIPTextModel, clip_text_config
    kw = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
              max_position_embeddings=77, hidden_act="quick_gelu", eos_token_id=2, bos_token_id=0, pad_token_id=1)
    hf = _hf(kw, False, 0)
    mine = B200CLIPTextModel(clip_text_config(), hf.state_dict(), "cuda")
    ids = _ids(3, 49408, 49407, 1)
    with torch.no_grad():
        ref = hf(ids, output_hidden_states=True)
    out = mine(ids.cuda(), output_hidden_states=True)
    torch.cuda.synchronize()
    assert len(out.hidden_states) == len(ref.hidden_states) == 13
    _report("clip_l_last", out[0], ref[0])
    _report("clip_l_penultimate", out.hidden_states[-2], ref.hidden_states[-2])
    _report("clip_l_pooled", out.pooler_output, ref.pooler_output)


def test_clip_bigg_with_projection_matches_transformers():
    """OpenCLIP bigG text tower shape at reduced depth (8 of 32 layers: the CPU reference stays in seconds)."""
    from invertible_cd_b200.text_encoder import B200CLIPTextModel, clip_bigg_text_config
    kw = dict(vocab_size=49408, hidden_size=1280, intermediate_size=5120, num_hidden_layers=8, num_attention_heads=20,
              max_position_embeddings=77, hidden_act="gelu", projection_dim=1280, eos_token_id=2, bos_token_id=0,
              pad_token_id=1)
    hf = _hf(kw, True, 1)
    mine = B200CLIPTextModel(clip_bigg_text_config(num_hidden_layers=8), hf.state_dict(), "cuda")
    ids = _ids(2, 49408, 49407, 2)
    with torch.no_grad():
        ref = hf(ids, output_hidden_states=True)
    out = mine(ids.cuda(), output_hidden_states=True)
    torch.cuda.synchronize()
    _report("clip_g_text_embeds", out[0], ref[0])
    _report("clip_g_penultimate", out.hidden_states[-2], ref.hidden_states[-2])


def test_init_prompt_and_encode_prompt_with_strings():
    """The reference's call sites with prompt STRINGS: Generator.init_prompt (utils/generation.py:286-303) builds the
    [uncond ; cond] context through tokenizer + text_encoder; generation_sdxl.encode_prompt (:9-46) concatenates the
    penultimate hidden states of two encoders and keeps the pooled output of the second."""
    from invertible_cd_b200 import arch, generation, generation_sdxl
    from invertible_cd_b200.loading import ICDPipeline
    from invertible_cd_b200.schedulers import DDIMScheduler, DDPMScheduler
    from invertible_cd_b200.text_encoder import (B200CLIPTextModel, clip_text_config, synthetic_clip_state_dict)
    from invertible_cd_b200.unet import B200UNet
    from toy_tokenizer import ToyCallableTokenizer
    c1 = clip_text_config(hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                          vocab_size=4096, eos_token_id=2)
    c2 = clip_text_config(hidden_size=64, intermediate_size=256, num_hidden_layers=3, num_attention_heads=1,
                          vocab_size=4096, eos_token_id=2, hidden_act="gelu", projection_dim=64)
    e1 = B200CLIPTextModel(c1, synthetic_clip_state_dict(c1, 1), "cuda")
    e2 = B200CLIPTextModel(c2, synthetic_clip_state_dict(c2, 2), "cuda")
    tok = ToyCallableTokenizer()
    prompts = ["a photo of a house on a mountain", "a painting of a lighthouse at winter evening"]
    # SD1.5 path
    ucfg = arch.small_sd15_config(cross_attention_dim=128)
    pipe = ICDPipeline(B200UNet(ucfg, arch.synthetic_state_dict(ucfg, seed=0), "cuda"), DDIMScheduler(), None, tok, e1,
                       device="cuda")
    solver = generation.Generator(model=pipe, n_steps=50, noise_scheduler=DDPMScheduler(), forward_cons_model=pipe,
                                  reverse_cons_model=pipe, reverse_timesteps=[259, 519, 779, 999],
                                  forward_timesteps=[19, 259, 519, 779])
    solver.init_prompt(prompts)
    assert solver.context.shape == (4, 77, 128) and torch.isfinite(solver.context).all()
    assert torch.equal(solver.context[0], solver.context[1])          # the "" embedding, expanded
    ids = tok(prompts).input_ids
    assert torch.equal(solver.context[2:], e1(ids.cuda())[0])
    lat = solver.cons_generation(torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(0)).cuda(),
                                 guidance_scale=7.0, w_embed_dim=512)[-1]
    assert torch.isfinite(lat).all()
    # SDXL path
    embeds, pooled = generation_sdxl.encode_prompt(prompts, [e1, e2], [tok, tok], 0.0, is_train=False)
    assert embeds.shape == (2, 77, 128 + 64) and pooled.shape == (2, 64)
    o2 = e2(ids.cuda(), output_hidden_states=True)
    assert torch.equal(embeds[..., 128:], o2.hidden_states[-2]) and torch.equal(pooled, o2[0])
