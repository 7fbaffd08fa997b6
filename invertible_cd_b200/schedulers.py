"""Minimal scheduler objects exposing what the iCD loop reads from diffusers schedulers:
`alphas_cumprod`, `final_alpha_cumprod`, `config.{prediction_type,num_train_timesteps}`, `set_timesteps`,
`timesteps`, `num_inference_steps`, and `add_noise` for the DDPM noise scheduler (utils/generation.py:427)."""
from types import SimpleNamespace

import numpy as np
import torch


class DDIMScheduler:
    """DDIMScheduler(beta_start=.00085, beta_end=.012, 'scaled_linear', clip_sample=False, set_alpha_to_one=False)
    as constructed at utils/loading.py:39-40 (SDXL's from_pretrained scheduler has the same betas)."""

    def __init__(self, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000,
                 clip_sample=False, set_alpha_to_one=False, steps_offset=0, prediction_type="epsilon"):
        if beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(beta_schedule)
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, prediction_type=prediction_type,
                                      steps_offset=steps_offset, beta_start=beta_start, beta_end=beta_end,
                                      beta_schedule=beta_schedule, clip_sample=clip_sample,
                                      set_alpha_to_one=set_alpha_to_one, timestep_spacing="leading")
        self.num_train_timesteps = num_train_timesteps
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))
        self.init_noise_sigma = 1.0

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        step_ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device) if device is not None else torch.from_numpy(ts)

    def add_noise(self, original_samples, noise, timesteps):
        acp = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sa = (acp[timesteps] ** 0.5).flatten()
        so = ((1 - acp[timesteps]) ** 0.5).flatten()
        while sa.dim() < original_samples.dim():
            sa, so = sa.unsqueeze(-1), so.unsqueeze(-1)
        return sa * original_samples + so * noise


class DDPMScheduler(DDIMScheduler):
    """Only `add_noise` is used by the path (Generator.cons_inversion)."""
