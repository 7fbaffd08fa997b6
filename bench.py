#!/usr/bin/env python
"""Benchmark of the iCD hot path (BASELINE.json metric: 4-step iCD latents/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload sd15|sdxl]

One "step" = one complete K_icd-step reverse-consistency generation of the per-GPU batch (K_icd U-Net row-forwards
+ fused consistency updates + AttentionStore cross-map capture), i.e. BASELINE.json configs[1]:
iCD-SD1.5, t = 999->779->519->259->0, batch 8 per GPU, 512^2 (4x64x64 latents), w_embed_dim 512, guidance 19,
LoRA r=64 fused at load, random weights / synthetic latents+context (no network for checkpoints).

  value     whole-job latents/s with inputs resident in HBM; the whole loop is one CUDA-graph replay per step
  e2e       the same metric through the public API (generation.runner) with HOST (pinned) inputs: H2D of the latent
            and the context and D2H of the finished latents inside the timed region, eager launches
  roofline  dominant kernel (gemm_tc: every conv/linear) — algorithmic FLOPs / CUDA-event time, vs measured bf16 peak
  cpu_baseline / --impl reference: the oracle restatement of the reference's diffusers path (the reference itself
            cannot run here: diffusers is not installed) driven by the reference's own doubled-batch procedure on
            the host cores, fp32, bounded sample of 1 prompt per step.
Multi-GPU: one process per GPU (torchrun), batch-axis sharding only, one all-gather of the finished latents per
step; time = max over ranks.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    "sd15": dict(model="synthetic:sd15:0", per_gpu_batch=8, latent=64, ctx_dim=768, xl=False,
                 reverse=[259, 519, 779, 999], forward=[19, 259, 519, 779], guidance=19.0,
                 flop_per_row_forward=803.27e9,
                 name="iCD-SD1.5 4-step reverse generation t=[999,779,519,259]->0, 512^2, w_embed 512, w=19, LoRA r=64"),
    "sdxl": dict(model="synthetic:sdxl:0", per_gpu_batch=4, latent=128, ctx_dim=2048, xl=True,
                 reverse=[249, 499, 699, 999], forward=[19, 249, 499, 699], guidance=7.0,
                 flop_per_row_forward=6761.24e9,
                 name="iCD-SDXL 4-step reverse generation t=[999,699,499,249]->0, 1024^2, w_embed 512, w=7, LoRA r=64"),
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    def __init__(self, dev_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(dev_index),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                sm.append(float(f[0]))
                mx = float(f[1])
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     f[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons)}
        return out


# ------------------------------------------------------------------------------------------------ reference arm
def _oracle_pipeline(wl, threads):
    """Oracle U-Net (CPU, fp32) of the workload's architecture with cheap random weights."""
    from invertible_cd_b200.loading import ICDPipeline
    from invertible_cd_b200.schedulers import DDIMScheduler
    from oracle import unet_oracle as O
    torch.set_num_threads(threads)
    cfg = O.sdxl_config() if wl["xl"] else O.sd15_config()
    with torch.device("meta"):
        model = O.UNet2DConditionModel(cfg)
    model = model.to_empty(device="cpu")
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 1:
                p.fill_(1.0 if name.endswith("weight") else 0.0)
            else:
                p.normal_(0.0, 0.02, generator=g)
    return ICDPipeline(model.eval(), DDIMScheduler(), device="cpu"), cfg


def run_reference_loop(wl, steps, warmup, threads):
    """The reference's procedure (utils/generation.py:373-412: doubled batch, CPU-built w-embedding,
    predicted_origin) over the oracle U-Net on the host cores; each step = 1 prompt, full K-step loop."""
    from invertible_cd_b200 import generation
    from invertible_cd_b200.schedulers import DDPMScheduler
    if wl["xl"]:
        return _run_reference_loop_xl(wl, steps, warmup, threads)
    pipe, cfg = _oracle_pipeline(wl, threads)
    solver = generation.Generator(model=pipe, n_steps=50, noise_scheduler=DDPMScheduler(), forward_cons_model=pipe,
                                  reverse_cons_model=pipe, reverse_timesteps=list(wl["reverse"]),
                                  forward_timesteps=list(wl["forward"]))
    g = torch.Generator().manual_seed(1)
    S = wl["latent"]
    times = []
    for i in range(warmup + steps):
        ctx = torch.randn(1, 77, wl["ctx_dim"], generator=g)
        lat = torch.randn(1, 4, S, S, generator=g)
        t0 = time.perf_counter()
        solver.init_prompt(ctx)
        solver.cons_generation(lat, guidance_scale=wl["guidance"], w_embed_dim=512, dynamic_guidance=False)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return 1.0 / statistics.mean(times), statistics.mean(times) * 1e3


def _run_reference_loop_xl(wl, steps, warmup, threads):
    from invertible_cd_b200 import generation_sdxl
    pipe, cfg = _oracle_pipeline(wl, threads)
    pipe.scheduler.num_train_timesteps = 1000
    g = torch.Generator().manual_seed(1)
    S = wl["latent"]
    times = []
    for i in range(warmup + steps):
        emb = {"prompt_embeds": torch.randn(1, 77, wl["ctx_dim"], generator=g),
               "text_embeds": torch.randn(1, 1280, generator=g),
               "time_ids": torch.tensor([[1024., 1024., 0., 0., 1024., 1024.]])}
        lat = torch.randn(1, 4, S, S, generator=g)
        t0 = time.perf_counter()
        pipe.dtype = torch.float32
        generation_sdxl.sample_deterministic(pipe, emb, latents=lat, num_inference_steps=4,
                                             timesteps=list(wl["reverse"]), guidance_scale=wl["guidance"],
                                             is_sdxl=True, return_latent=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return 1.0 / statistics.mean(times), statistics.mean(times) * 1e3


# ------------------------------------------------------------------------------------------------ our arm
def build_ours(wl, device):
    from invertible_cd_b200 import generation, loading
    from invertible_cd_b200.schedulers import DDPMScheduler
    if wl["xl"]:
        stable, rev, fwd = loading.load_models_xl(wl["model"], "synthetic:1", "synthetic:2", None, device=device)
        return stable, rev, None
    ldm, rev, fwd = loading.load_models(wl["model"], device, "synthetic:1", None, r=64, w_embed_dim=512,
                                        dtype="fp16")
    solver = generation.Generator(model=ldm, n_steps=50, noise_scheduler=DDPMScheduler(), forward_cons_model=fwd,
                                  reverse_cons_model=rev, reverse_timesteps=list(wl["reverse"]),
                                  forward_timesteps=list(wl["forward"]))
    return ldm, rev, solver


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sd15", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-step", action="store_true",
                    help="run ONE eager step inside an NVTX range 'icd_step' (for ncu) and dump its launch shapes")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    base = {"metric": "4-step iCD latents/sec", "unit": "latents/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "data": "synthetic", "dtype": "fp16",
            "config": {"workload": wl["name"], "per_gpu_batch": wl["per_gpu_batch"],
                       "controller": "AttentionStore (cross maps fused into the attention kernel)",
                       "l2": "weights (1.7 GB SD1.5 / 5.1 GB SDXL per model) exceed the 126 MB L2: no flush needed",
                       "parallelism": f"dp{args.gpus} (batch axis, one all-gather of latents per step)"}}

    if args.impl == "reference":
        if rank != 0:
            return
        value, ms = run_reference_loop(wl, args.steps, max(1, min(args.warmup, 1)), cores)
        line = dict(base)
        line.update({"impl": "reference", "value": value, "ms_per_step": ms, "dtype": "f32", "n_gpus": args.gpus,
                     "cpu_baseline": {"value": value, "unit": "latents/s", "cores": cores, "kind": "port",
                                      "sample": "1 prompt per step, full K-step loop, doubled U-Net batch (reference "
                                                "procedure), fp32 oracle U-Net"},
                     "e2e": {"value": value, "unit": "latents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "gpu_launches": 0})
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ ours
    import torch.distributed as dist
    from invertible_cd_b200 import dist_utils, generation, generation_sdxl, ops, p2p
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device: the iCD path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    if world > 1:
        dist_utils.init("nccl")
    ldm, rev, solver = build_ours(wl, device)
    B, S = wl["per_gpu_batch"], wl["latent"]
    g = torch.Generator().manual_seed(100 + rank)
    host_lat = torch.randn(B, 4, S, S, generator=g).pin_memory()
    host_ctx = torch.randn(B, 77, wl["ctx_dim"], generator=g).half().pin_memory()
    static_lat = host_lat.to(device)
    static_ctx = host_ctx.to(device)
    n_total = B * world
    added = None
    if wl["xl"]:
        added = {"prompt_embeds": static_ctx, "text_embeds": torch.randn(B, 1280, generator=g).half().to(device),
                 "time_ids": torch.tensor([[1024., 1024., 0., 0., 1024., 1024.]] * B).to(device)}

    def loop(lat, ctx, gather=True):
        """The hot path on resident inputs: K_icd U-Net forwards + fused updates + AttentionStore capture."""
        if wl["xl"]:
            emb = dict(added)
            emb["prompt_embeds"] = ctx
            out = generation_sdxl.sample_deterministic(rev, emb, latents=lat, num_inference_steps=4,
                                                       timesteps=list(wl["reverse"]), guidance_scale=wl["guidance"],
                                                       is_sdxl=True, return_latent=True)[1]
        else:
            store = p2p.AttentionStore()
            store.capture_self = False
            p2p.register_attention_control(rev, store)
            solver.context = torch.cat([ctx, ctx])       # [uncond ; cond] layout of init_prompt; uncond rows unused
            out = solver.cons_generation(lat, guidance_scale=wl["guidance"], w_embed_dim=512, dynamic_guidance=False,
                                         controller=store)[-1]
        if gather and world > 1:
            out = dist_utils.gather_latents(out, n_total, B)
        return out

    if args.profile_step:
        for _ in range(2):
            loop(static_lat, static_ctx)
        torch.cuda.synchronize()
        ops.shape_log = []
        torch.cuda.nvtx.range_push("icd_step")
        loop(static_lat, static_ctx)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"step_shapes_{args.workload}.json"), "w") as f:
            json.dump(ops.shape_log, f)
        return

    # eager warm-up (fills descriptor / constant caches), then count launches of one step
    for _ in range(2):
        loop(static_lat, static_ctx)
    torch.cuda.synchronize()
    n0 = ops.launch_count
    loop(static_lat, static_ctx)
    launches_per_step = ops.launch_count - n0
    torch.cuda.synchronize()

    # capture the whole step (local part) into one CUDA graph
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        loop(static_lat, static_ctx, gather=False)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        graph_out = loop(static_lat, static_ctx, gather=False)

    def graphed_step():
        graph.replay()
        if world > 1:
            return dist_utils.gather_latents(graph_out, n_total, B)
        return graph_out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        graphed_step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        graphed_step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler is not None else None
    if world > 1:
        t = torch.tensor([ms_total], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = t.item()
    ms_per_step = ms_total / args.steps
    value = n_total / (ms_per_step / 1e3)

    # ------------------------------------------------------------------ e2e through the public API, host buffers
    def e2e_step():
        if wl["xl"]:
            emb = dict(added)
            emb["prompt_embeds"] = host_ctx.to(device, non_blocking=True)
            lat = host_lat.to(device, non_blocking=True)
            out = generation_sdxl.sample_deterministic(rev, emb, latents=lat, num_inference_steps=4,
                                                       timesteps=list(wl["reverse"]), guidance_scale=wl["guidance"],
                                                       is_sdxl=True, return_latent=True)[1]
        else:
            store = p2p.AttentionStore()
            store.capture_self = False
            out, _ = generation.runner(model=rev, prompt=host_ctx, controller=store, solver=solver,
                                       is_cons_forward=True, guidance_scale=wl["guidance"], latent=host_lat[:1],
                                       return_type="latent", tau1=1.0, tau2=1.0, w_embed_dim=512)
        if world > 1:
            out = dist_utils.gather_latents(out, n_total, B)
        return out.to("cpu", non_blocking=False)

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    ev0.record()
    for _ in range(e2e_steps):
        res = e2e_step()
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = t.item()
    e2e_value = n_total / (e2e_ms / e2e_steps / 1e3)
    h2d = (host_lat[:1].numel() * 4 if not wl["xl"] else host_lat.numel() * 4) + host_ctx.numel() * 2
    d2h = res.numel() * res.element_size()

    if rank != 0:
        return

    # ------------------------------------------------------------------ roofline of the dominant kernel (live events)
    ops.profile = []
    loop(static_lat, static_ctx, gather=False)
    torch.cuda.synchronize()
    agg = {}
    for kind, work, a, b in ops.profile:
        d = agg.setdefault(kind, [0.0, 0.0, 0])
        d[0] += work
        d[1] += a.elapsed_time(b)
        d[2] += 1
    ops.profile = None
    peaks, peak_src = _peaks()
    kernels = {}
    for kind, (work, ms, n) in agg.items():
        if kind == "groupnorm":
            kernels[kind] = {"launch_pairs": n, "ms": ms, "achieved_GBps": work / (ms * 1e-3) / 1e9,
                             "frac_of_hbm_peak": work / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
        else:
            kernels[kind] = {"launches": n, "ms": ms, "achieved_TFLOPs": work / (ms * 1e-3) / 1e12,
                             "frac_of_bf16_sustained": work / (ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"]}
    gw, gms, gn = agg["gemm_tc"]
    peak_tf = peaks["bf16_tflops_sustained"]
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r1_h_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if args.workload in tj:
            traffic = tj[args.workload]["gemm_tc_dram_bytes_per_launch"]
            traffic_src = "profiles/r1_h_traffic.json (ncu dram__bytes_read+write, average over the step's gemm_tc launches)"
    roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel (all convs + linears)",
                "achieved": gw / (gms * 1e-3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": gw / (gms * 1e-3) / 1e12 / peak_tf, "traffic": traffic, "traffic_unit": "bytes/launch",
                "traffic_source": traffic_src, "achieved_flop_per_launch": gw / gn,
                "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)",
                "launches_per_step": gn, "ms_per_step_in_kernel": gms,
                "share_of_eager_step": gms / sum(v[1] for v in agg.values()),
                "whole_step_frac": (wl["flop_per_row_forward"] * B * 4) / (ms_per_step * 1e-3) / 1e12 / peak_tf,
                "note": "per-launch times taken eagerly with CUDA events around each launch (includes launch gaps)"}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        v, ms = run_reference_loop(wl, 1, 0, cores) if not wl["xl"] else (None, None)
        if v is not None:
            cpu_baseline = {"value": v, "unit": "latents/s", "cores": cores, "kind": "port",
                            "sample": "1 prompt, full 4-step loop, doubled U-Net batch (reference procedure), fp32 "
                                      f"oracle U-Net, {ms / 1e3:.1f} s"}

    line = dict(base)
    line.update({"value": value, "ms_per_step": ms_per_step, "clocks": clocks,
                 "e2e": {"value": e2e_value, "unit": "latents/s", "h2d_bytes_per_step": h2d,
                         "d2h_bytes_per_step": d2h, "mode": "eager launches through generation.runner / "
                                                            "sample_deterministic, pinned host inputs"},
                 "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
                 "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline})
    print(json.dumps(line))


if __name__ == "__main__":
    try:
        main()
    finally:
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()
