"""VAE parity (GPU): B200VAE (sm_100a kernels through the C ABI) vs the CPU fp32 oracle restatement of
diffusers' AutoencoderKL (oracle/vae_oracle.py), same weights and inputs; then `runner(return_type='image')` and
`invert(image=...)` end to end. Gates <= 2x the error measured on the B200 (profiles/r2_parity_report.jsonl)."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (max rel-L2, max |err| / max|ref|); measured: decode 1.72e-3..1.78e-3 / 2.0e-3..2.7e-3, encode 1.25e-3..1.28e-3 / 1.5e-3..1.7e-3
GATES = {"vae_decode_small": (3.6e-3, 5.5e-3), "vae_encode_small": (2.6e-3, 3.4e-3), "vae_decode_512": (3.6e-3, 5.5e-3),
         "vae_encode_512": (2.6e-3, 3.4e-3), "vae_decode_256_b3": (3.6e-3, 5.5e-3)}


def _report(name, got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    diff = (got - ref).abs()
    rec = {"name": name, "shape": list(ref.shape), "max_abs": diff.max().item(), "ref_absmax": ref.abs().max().item(),
           "rel_l2": ((got - ref).norm() / ref.norm()).item(),
           "viol_rtol1e-3_atol1e-4": int((diff > 1e-4 + 1e-3 * ref.abs()).sum().item()), "numel": ref.numel()}
    rec["max_rel_to_absmax"] = rec["max_abs"] / max(rec["ref_absmax"], 1e-30)
    print("[parity]", json.dumps(rec))
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_report.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
    assert torch.isfinite(got).all(), name
    assert rec["rel_l2"] <= GATES[name][0] and rec["max_rel_to_absmax"] <= GATES[name][1], rec


@pytest.fixture(scope="module")
def vaes():
    from invertible_cd_b200.vae import B200VAE, synthetic_vae_state_dict, vae_config
    from oracle import vae_oracle as V
    cfg = vae_config()
    sd = synthetic_vae_state_dict(cfg, seed=3)
    with torch.device("meta"):
        oracle = V.AutoencoderKL(V.sd15_vae_config())
    oracle.load_state_dict({k: v.float() for k, v in sd.items()}, strict=True, assign=True)
    return B200VAE(cfg, sd, "cuda"), oracle.eval()


@pytest.mark.parametrize("name,B,S", [("vae_decode_small", 2, 16), ("vae_decode_256_b3", 3, 32), ("vae_decode_512", 1, 64)])
def test_decode_matches_oracle(vaes, name, B, S):
    vae, oracle = vaes
    z = torch.randn(B, 4, S, S, generator=torch.Generator().manual_seed(S)) * 3.0
    out = vae.decode(z.cuda())["sample"]
    torch.cuda.synchronize()
    assert out.shape == (B, 3, 8 * S, 8 * S) and out.dtype == torch.float32
    with torch.no_grad():
        ref = oracle.decode(z)["sample"]
    _report(name, out, ref)
    assert vae.decode(z.cuda(), return_dict=False)[0].shape == out.shape


@pytest.mark.parametrize("name,B,S", [("vae_encode_small", 2, 128), ("vae_encode_512", 1, 512)])
def test_encode_matches_oracle(vaes, name, B, S):
    vae, oracle = vaes
    x = torch.rand(B, 3, S, S, generator=torch.Generator().manual_seed(S)) * 2 - 1
    dist = vae.encode(x.cuda())["latent_dist"]
    torch.cuda.synchronize()
    with torch.no_grad():
        rdist = oracle.encode(x)["latent_dist"]
    assert dist.mean.shape == (B, 4, S // 8, S // 8)
    _report(name, dist.parameters, rdist.parameters)
    # sample(generator): same noise on both sides (CPU generator, the moments' dtype)
    a = dist.sample(torch.Generator().manual_seed(5)).cpu()
    b = rdist.sample(torch.Generator().manual_seed(5))
    assert ((a - b).norm() / b.norm()).item() <= 2 * GATES[name][0]


def test_runner_returns_images_and_invert_accepts_images(tmp_path):
    """runner(return_type='image') -> uint8 (B, 512, 512, 3) like utils/generation.py:62,527-533; invert() from a
    512x512 uint8 image (utils/generation.py:266-284 via cons_inversion) round-trips through the VAE encoder."""
    import numpy as np
    from invertible_cd_b200 import generation, inversion, loading, p2p
    from invertible_cd_b200.schedulers import DDPMScheduler
    ldm, rev, fwd = loading.load_models("synthetic:small_sd15:0", "cuda", "synthetic:1", "synthetic:2", r=8,
                                        w_embed_dim=512, dtype="fp16")
    assert ldm.vae is not None and rev.vae is ldm.vae
    # the small synthetic U-Net works on 64x64 latents when asked to (sample_size is only a default)
    solver = generation.Generator(model=ldm, n_steps=50, noise_scheduler=DDPMScheduler(), forward_cons_model=fwd,
                                  reverse_cons_model=rev, reverse_timesteps=[259, 519, 779, 999],
                                  forward_timesteps=[19, 259, 519, 779])
    ctx = torch.randn(2, 77, 128, generator=torch.Generator().manual_seed(1)).half().float()
    images, x_T = generation.runner(model=rev, prompt=ctx, controller=p2p.AttentionStore(), solver=solver,
                                    is_cons_forward=True, guidance_scale=19.0, return_type="image", tau1=0.8, tau2=0.8,
                                    w_embed_dim=512, generator=torch.Generator().manual_seed(2))
    assert images.shape == (2, 512, 512, 3) and images.dtype == np.uint8 and x_T.shape == (1, 4, 64, 64)
    from PIL import Image
    path = str(tmp_path / "gen.png")
    Image.fromarray(images[0]).save(path)
    (gt, rec), latent, _ = inversion.invert(solver, stop_step=50, is_cons_inversion=True, inv_guidance_scale=0.0,
                                            w_embed_dim=512, image_path=path, prompt=ctx[:1], seed=3)
    assert gt.shape == (512, 512, 3) and np.array_equal(gt, images[0])
    assert latent.shape == (1, 4, 64, 64) and torch.isfinite(latent).all()
    assert rec.shape == (512, 512, 3) and rec.dtype == np.uint8
