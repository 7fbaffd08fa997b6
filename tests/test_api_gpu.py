"""Public-API parity on the GPU: load_models / Generator / runner / invert / p2p controllers / SDXL samplers over the
B200 U-Net vs the same host code over the CPU oracle U-Net (the host code itself is pinned to the reference's output in
tests/test_host_cpu.py). Tolerance: relative L2 error of the K-step latents <= 2x the value measured on the B200
(fp16 noise floor through K U-Net passes: 2.4e-4 .. 1.7e-3), stated per test."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


def _oracle_pipe(cfg, sd, xl=False):
    from invertible_cd_b200.loading import ICDPipeline
    from invertible_cd_b200.schedulers import DDIMScheduler
    from oracle import unet_oracle as O
    mk = O.tiny_sdxl_config if xl else O.tiny_sd15_config
    ocfg = mk(sample_size=cfg.sample_size, block_out_channels=cfg.block_out_channels,
              attention_head_dim=cfg.attention_head_dim, cross_attention_dim=cfg.cross_attention_dim,
              transformer_layers_per_block=cfg.transformer_layers_per_block)
    m = O.UNet2DConditionModel(ocfg).eval()
    m.load_state_dict({k: v.float().cpu() for k, v in sd.items()}, strict=True)
    sch = DDIMScheduler()
    sch.num_train_timesteps = 1000
    return ICDPipeline(m, sch, device="cpu")


def _solver(pipe, rev, fwd):
    from invertible_cd_b200 import generation
    from invertible_cd_b200.schedulers import DDPMScheduler
    return generation.Generator(model=pipe, n_steps=50, noise_scheduler=DDPMScheduler(), forward_cons_model=fwd,
                                reverse_cons_model=rev, reverse_timesteps=[259, 519, 779, 999],
                                forward_timesteps=[19, 259, 519, 779])


@pytest.fixture(scope="module")
def sd15_models():
    from invertible_cd_b200 import arch, loading
    cfg = arch.small_sd15_config()
    base = arch.synthetic_state_dict(cfg, seed=0)
    lora_r, lora_f = arch.synthetic_lora(cfg, r=8, seed=1), arch.synthetic_lora(cfg, r=8, seed=2)
    ldm, rev, fwd = loading.load_models("synthetic:small_sd15:0", "cuda", lora_r, lora_f, r=8, w_embed_dim=512,
                                        dtype="fp16")
    o_rev = _oracle_pipe(cfg, loading.fuse_lora(base, lora_r, r=8))
    o_fwd = _oracle_pipe(cfg, loading.fuse_lora(base, lora_f, r=8))
    return cfg, ldm, rev, fwd, o_rev, o_fwd


def test_runner_generation_with_attention_store(sd15_models):
    from invertible_cd_b200 import generation, p2p
    cfg, ldm, rev, fwd, o_rev, o_fwd = sd15_models
    g = torch.Generator().manual_seed(5)
    ctx = torch.randn(3, 77, cfg.cross_attention_dim, generator=g).half().float()
    x_T = torch.randn(1, 4, 64, 64, generator=g)
    solver = _solver(ldm, rev, fwd)
    store = p2p.AttentionStore()
    lat, x_back = generation.runner(model=rev, prompt=ctx, controller=store, solver=solver, is_cons_forward=True,
                                    guidance_scale=19.0, latent=x_T, return_type="latent", tau1=0.8, tau2=0.8,
                                    w_embed_dim=512)
    assert lat.shape == (3, 4, 64, 64) and lat.dtype == torch.float32 and x_back is x_T
    assert store.cur_step == 4 and len(store.attention_store["down_cross"]) == 4   # 64x64 input: 4096-query level not stored
    o_solver = _solver(o_rev, o_rev, o_fwd)
    o_solver.init_prompt(ctx)
    ref = o_solver.cons_generation(x_T.expand(3, 4, 64, 64), guidance_scale=19.0, w_embed_dim=512,
                                   dynamic_guidance=True, tau1=0.8, tau2=0.8)[-1]
    err = _rel(lat, ref)
    print("runner 4-step rel-L2:", err)
    assert err <= 5e-4                                   # measured 2.41e-4


def test_invert_then_edit_with_refine_controller(sd15_models):
    """Config 3 shape: forward-consistency inversion (w=0, forward LoRA) then a 4-step reverse edit of 2 prompts with
    an AttentionRefine + LocalBlend controller (4 U-Net rows in the reference == 2 conditional rows here)."""
    from invertible_cd_b200 import generation, inversion, p2p
    from oracle import unet_oracle as O
    from toy_tokenizer import ToyTokenizer
    cfg, ldm, rev, fwd, o_rev, o_fwd = sd15_models
    g = torch.Generator().manual_seed(9)
    ctx = torch.randn(2, 77, cfg.cross_attention_dim, generator=g).half().float()
    image_latent = torch.randn(1, 4, 64, 64, generator=g) * 0.5
    solver = _solver(ldm, rev, fwd)
    (_, _), x_inv, uncond = inversion.invert(solver, stop_step=50, is_cons_inversion=True, inv_guidance_scale=0.0,
                                             w_embed_dim=512, image_path=image_latent.cuda(), prompt=ctx[:1], seed=3)
    assert uncond is None and x_inv.shape == (1, 4, 64, 64)
    o_solver = _solver(o_rev, o_rev, o_fwd)
    o_solver.init_prompt(ctx[:1])
    _, o_inv = o_solver.cons_inversion(image_latent, guidance_scale=0.0, w_embed_dim=512, seed=3)
    err_inv = _rel(x_inv, o_inv[0])
    print("inversion rel-L2:", err_inv)
    assert err_inv <= 3.2e-3                             # measured 1.57e-3

    prompts = ["a photo of a house on a mountain", "a photo of a house on a mountain at winter evening"]

    def controller(dev):
        p2p.tokenizer, p2p.device, p2p.NUM_DDIM_STEPS = ToyTokenizer(), dev, 4
        return p2p.make_controller(prompts, False, {"default_": 0.3}, 0.6, None, None)

    ctrl = controller("cuda")
    lat, _ = generation.runner(model=rev, prompt=ctx, controller=ctrl, solver=solver, is_cons_forward=True,
                               guidance_scale=19.0, latent=o_inv[0], return_type="latent", tau1=0.8, tau2=0.8,
                               w_embed_dim=512)
    o_ctrl = controller("cpu")
    O.register_attention_control(o_rev.unet, o_ctrl)
    o_solver.init_prompt(ctx)
    ref = o_solver.cons_generation(o_inv[0].expand(2, 4, 64, 64), guidance_scale=19.0, w_embed_dim=512,
                                   dynamic_guidance=True, tau1=0.8, tau2=0.8, controller=o_ctrl)[-1]
    O.register_attention_control(o_rev.unet, None)
    err = _rel(lat, ref)
    print("edit rel-L2:", err, "steps", ctrl.cur_step, o_ctrl.cur_step)
    assert ctrl.cur_step == o_ctrl.cur_step == 4
    assert err <= 3.5e-3                                 # measured 1.71e-3


def test_cuda_graph_replay_equals_eager(sd15_models):
    from invertible_cd_b200 import p2p
    cfg, ldm, rev, fwd, _, _ = sd15_models
    solver = _solver(ldm, rev, fwd)
    g = torch.Generator().manual_seed(13)
    ctx = torch.randn(2, 77, cfg.cross_attention_dim, generator=g).half().cuda()
    lat = torch.randn(2, 4, 64, 64, generator=g).cuda()

    def loop():
        store = p2p.AttentionStore()
        p2p.register_attention_control(rev, store)
        solver.context = torch.cat([ctx, ctx]).float()
        return solver.cons_generation(lat, guidance_scale=7.0, w_embed_dim=512, dynamic_guidance=False,
                                      controller=store)[-1], store

    eager, _ = loop()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out, store = loop()
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager)                      # deterministic kernels: bit-identical replay
    assert store.cur_step == 4 and torch.isfinite(store.attention_store["up_cross"][0]).all()
    p2p.register_attention_control(rev, None)


def test_sdxl_sample_and_inverse(sd15_models):
    from invertible_cd_b200 import arch, generation_sdxl, loading
    cfg = arch.small_sdxl_config()
    base = arch.synthetic_state_dict(cfg, seed=0)
    lr, lf = arch.synthetic_lora(cfg, r=8, seed=1), arch.synthetic_lora(cfg, r=8, seed=2)
    stable, pipe, fpipe = loading.load_models_xl("synthetic:small_sdxl:0", lr, lf, None, device="cuda", r=8)
    base16 = {k: v.half() for k, v in base.items()}
    o_pipe = _oracle_pipe(cfg, loading.fuse_lora(base16, lr, r=8, lora_dtype=torch.float32), xl=True)
    o_fpipe = _oracle_pipe(cfg, loading.fuse_lora(base16, lf, r=8, lora_dtype=torch.float32), xl=True)
    g = torch.Generator().manual_seed(17)
    B = 2

    def emb(dev, dtype):
        gg = torch.Generator().manual_seed(23)
        return {"prompt_embeds": torch.randn(B, 77, cfg.cross_attention_dim, generator=gg).half().to(dev, dtype),
                "text_embeds": torch.randn(B, 64, generator=gg).half().to(dev, dtype),
                "time_ids": torch.tensor([[1024., 1024., 0., 0., 1024., 1024.]] * B).to(dev, dtype)}

    def src(dev, dtype):
        gg = torch.Generator().manual_seed(29)
        return {"prompt_embeds": torch.randn(B, 77, cfg.cross_attention_dim, generator=gg).half().to(dev, dtype),
                "text_embeds": torch.zeros(B, 64).to(dev, dtype), "time_ids": torch.zeros(B, 6).to(dev, dtype)}

    lat0 = torch.randn(B, 4, 16, 16, generator=g)
    o_pipe.unet.float(); o_pipe.dtype = torch.float32
    kw = dict(num_inference_steps=3, timesteps=[339, 699, 999], guidance_scale=19.0, is_sdxl=True, return_latent=True,
              use_dynamic_guidance=True, tau1=0.8, tau2=0.8)
    _, got = generation_sdxl.sample_deterministic(pipe, emb("cuda", torch.float16), latents=lat0.cuda(),
                                                  amplify_prompt=src("cuda", torch.float16), **kw)
    _, ref = generation_sdxl.sample_deterministic(o_pipe, emb("cpu", torch.float32), latents=lat0,
                                                  amplify_prompt=src("cpu", torch.float32), **kw)
    err = _rel(got, ref)
    print("sdxl reverse 3-step rel-L2:", err)
    assert got.dtype == torch.float16 and err <= 1e-3    # measured 4.8e-4
    img_lat = torch.randn(B, 4, 16, 16, generator=g) * 0.3
    kwi = dict(num_inference_steps=3, timesteps=[19, 339, 699], guidance_scale=0.0, is_sdxl=True, seed=4)
    goti, start = generation_sdxl.inverse_sample_deterministic(fpipe, img_lat.cuda(), emb("cuda", torch.float16),
                                                               return_start_latent=True, **kwi)
    # the start latent = add_noise(image latent, noise at t=19) with the noise drawn IN THE PIPELINE DTYPE on the CPU
    # generator, as diffusers' prepare_latents / randn_tensor do (fp16 here; an fp32 draw gives other values)
    noise = torch.randn(img_lat.shape, generator=torch.Generator().manual_seed(4), dtype=torch.float16)
    exp_start = fpipe.scheduler.add_noise(img_lat.half(), noise, torch.tensor([19]))
    torch.testing.assert_close(start.cpu().float(), exp_start.float(), atol=2e-3, rtol=2e-3)
    # the oracle pipeline is fp32 and would draw fp32 noise: run its loop from the same start latent instead
    e = emb("cpu", torch.float32)
    al, sg = generation_sdxl._schedule_tables(o_fpipe, "cpu")
    refi = generation_sdxl._loop(o_fpipe, start.float().cpu(), [19, 339, 699], [339, 699, 999], e["prompt_embeds"], None,
                                 {k: e[k] for k in ("text_embeds", "time_ids")}, [0.0] * B, False, 0.0, 0.0,
                                 torch.float32, al, sg)
    erri = _rel(goti, refi)
    print("sdxl forward 3-step rel-L2:", erri)
    assert erri <= 3.5e-3                                # measured 1.74e-3


def test_library_graph_cache_equals_eager(sd15_models):
    """The product path replays the K-step loop from the library's own CUDA-graph cache (graphs.py): results, the
    AttentionStore state and the step counters must equal the eager run bit for bit, on the capturing call and on
    later replays with different inputs; edit controllers keep running eagerly."""
    from invertible_cd_b200 import generation, graphs, inversion, p2p
    cfg, ldm, rev, fwd, _, _ = sd15_models
    solver = _solver(ldm, rev, fwd)
    g = torch.Generator().manual_seed(21)

    def run(seed, use_graphs, self_maps):
        prev = graphs.set_enabled(use_graphs)
        try:
            gg = torch.Generator().manual_seed(seed)
            ctx = torch.randn(2, 77, cfg.cross_attention_dim, generator=gg).half().float()
            x_T = torch.randn(1, 4, 64, 64, generator=gg)
            store = p2p.AttentionStore()
            store.capture_self = self_maps
            lat, _ = generation.runner(model=rev, prompt=ctx, controller=store, solver=solver, is_cons_forward=True,
                                       guidance_scale=19.0, latent=x_T, return_type="latent", tau1=0.8, tau2=0.8,
                                       w_embed_dim=512)
            torch.cuda.synchronize()
            return lat, store
        finally:
            graphs.set_enabled(prev)

    graphs.clear(rev.unet)
    before = dict(graphs.stats)
    for self_maps in (False, True):
        for seed in (1, 2, 3):                     # seed 1 captures, 2 and 3 replay with new inputs
            lat_g, st_g = run(seed, True, self_maps)
            lat_e, st_e = run(seed, False, self_maps)
            assert torch.equal(lat_g, lat_e)
            assert st_g.cur_step == st_e.cur_step == 4 and st_g.cur_att_layer == st_e.cur_att_layer == 0
            assert set(st_g.attention_store) == set(st_e.attention_store)
            for k in st_e.attention_store:
                assert len(st_g.attention_store[k]) == len(st_e.attention_store[k]), k
                for a, b in zip(st_g.attention_store[k], st_e.attention_store[k]):
                    assert torch.equal(a, b), k
    assert graphs.stats["captures"] - before["captures"] == 2          # one graph per controller kind
    assert graphs.stats["replays"] - before["replays"] == 6
    # forward-consistency inversion (no controller on the forward model) goes through the cache as well
    img = torch.randn(1, 4, 64, 64, generator=g) * 0.5
    ctx = torch.randn(1, 77, cfg.cross_attention_dim, generator=g).half().float()
    outs = []
    for use in (True, True, False):
        prev = graphs.set_enabled(use)
        (_, _), x_inv, _ = inversion.invert(solver, stop_step=50, is_cons_inversion=True, inv_guidance_scale=0.0,
                                            w_embed_dim=512, image_path=img.cuda(), prompt=ctx, seed=3)
        graphs.set_enabled(prev)
        outs.append(x_inv.clone())
    assert torch.equal(outs[0], outs[2]) and torch.equal(outs[1], outs[2])
    p2p.register_attention_control(rev, None)


def test_edit_controllers_replay_from_the_graph_cache(sd15_models):
    """BASELINE configs[2] is launch-bound when the edit loop runs eagerly (~350 launches per U-Net forward plus the
    controller's tensor ops, issued from Python). Fresh AttentionRefine / AttentionReplace / AttentionReweight
    controllers (with LocalBlend) are captured too: their per-edit tensors are inputs of the graph, so the NEXT edit —
    another controller object, other prompts — replays it. Results, stored maps and counters must equal the eager
    run bit for bit."""
    from invertible_cd_b200 import generation, graphs, p2p
    from toy_tokenizer import ToyTokenizer
    cfg, ldm, rev, fwd, _, _ = sd15_models
    solver = _solver(ldm, rev, fwd)
    p2p.tokenizer, p2p.device, p2p.NUM_DDIM_STEPS = ToyTokenizer(), "cuda", 4
    cases = {
        "refine": [(["a photo of a house on a mountain", "a photo of a house on a mountain at winter evening"],
                    dict(is_replace_controller=False, blend_words=(("mountain",), ("mountain",)))),
                   (["a cat sits on a sofa", "a cat sits on a sofa under a lamp"],
                    dict(is_replace_controller=False, blend_words=(("sofa",), ("sofa",))))],
        "replace": [(["a photo of a house", "a photo of a castle"], dict(is_replace_controller=True)),
                    (["a small dog runs fast", "a small cat runs fast"], dict(is_replace_controller=True))],
        "reweight": [(["a photo of a snowy house", "a photo of a snowy house"],
                      dict(is_replace_controller=False, equilizer_params={"words": ("snowy",), "values": (3.0,)})),
                     (["a very tall tree", "a very tall tree"],
                      dict(is_replace_controller=False, equilizer_params={"words": ("tall",), "values": (0.5,)}))],
    }

    def run(prompts, kw, seed, use_graphs):
        prev = graphs.set_enabled(use_graphs)
        try:
            g = torch.Generator().manual_seed(seed)
            ctx = torch.randn(2, 77, cfg.cross_attention_dim, generator=g).half().float()
            x_T = torch.randn(1, 4, 64, 64, generator=g)
            ctrl = p2p.make_controller(prompts, kw["is_replace_controller"], {"default_": 0.4}, 0.6,
                                       kw.get("blend_words"), kw.get("equilizer_params"))
            lat, _ = generation.runner(model=rev, prompt=ctx, controller=ctrl, solver=solver, is_cons_forward=True,
                                       guidance_scale=19.0, latent=x_T, return_type="latent", tau1=0.8, tau2=0.8,
                                       w_embed_dim=512)
            torch.cuda.synchronize()
            return lat, ctrl
        finally:
            graphs.set_enabled(prev)

    graphs.clear(rev.unet)
    before = dict(graphs.stats)
    for name, variants in cases.items():
        for i, (prompts, kw) in enumerate(variants + variants[:1]):      # capture, replay (other prompts), replay
            lat_g, c_g = run(prompts, kw, 100 + i, True)
            lat_e, c_e = run(prompts, kw, 100 + i, False)
            assert torch.equal(lat_g, lat_e), (name, i)
            assert c_g.cur_step == c_e.cur_step == 4 and c_g.cur_att_layer == c_e.cur_att_layer == 0
            if c_e.local_blend is not None:
                assert c_g.local_blend.counter == c_e.local_blend.counter == 4
            assert set(c_g.attention_store) == set(c_e.attention_store)
            for k in c_e.attention_store:
                assert len(c_g.attention_store[k]) == len(c_e.attention_store[k]), (name, k)
                for a, b in zip(c_g.attention_store[k], c_e.attention_store[k]):
                    assert torch.equal(a, b), (name, k)
    assert graphs.stats["captures"] - before["captures"] == 3           # one graph per controller kind
    assert graphs.stats["replays"] - before["replays"] == 9
    p2p.register_attention_control(rev, None)
