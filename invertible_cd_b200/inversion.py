"""Inversion driver — the surface of utils/inversion.py::invert (cons and DDIM branches).

Null-text inversion (`do_nti`, utils/inversion.py:11-48) needs autograd through the U-Net w.r.t. the unconditional
embeddings; the sm_100a path is inference-only, so that branch raises. NPI (`do_npi`) is just
a different `uncond_embeddings` list and is supported.
"""
import torch

from .generation import load_512
from .p2p import register_attention_control


def invert(solver, stop_step, is_cons_inversion=False, inv_guidance_scale=1, nti_guidance_scale=8,
           dynamic_guidance=False, tau1=0.4, tau2=0.6, w_embed_dim=0, image_path=None, prompt='',
           offsets=(0, 0, 0, 0), do_nti=False, do_npi=False, num_inner_steps=10, early_stop_epsilon=1e-5, seed=0):
    solver.init_prompt(prompt)
    uncond_embeddings, cond_embeddings = solver.context.chunk(2)
    register_attention_control(solver.model, None)      # teacher only; the students are untouched (§3.2)
    if torch.is_tensor(image_path):
        image_gt = image_path                           # pre-encoded latent(s): no VAE on this path (§8f)
    elif isinstance(image_path, list):
        image_gt = [load_512(path, *offsets) for path in image_path]
    else:
        image_gt = load_512(image_path, *offsets)
    if is_cons_inversion:
        image_rec, latents = solver.cons_inversion(image_gt, w_embed_dim=w_embed_dim,
                                                   guidance_scale=inv_guidance_scale, seed=seed)
    else:
        image_rec, latents = solver.ddim_inversion(image_gt, n_steps=stop_step, guidance_scale=inv_guidance_scale,
                                                   dynamic_guidance=dynamic_guidance, tau1=tau1, tau2=tau2,
                                                   w_embed_dim=w_embed_dim)
    if do_nti:
        raise NotImplementedError("null-text optimisation needs backward through the U-Net; the B200 path is "
                                  "inference-only (SURVEY §8f)")
    uncond = [cond_embeddings] * solver.n_steps if do_npi else None
    return (image_gt, image_rec), latents[-1], uncond
