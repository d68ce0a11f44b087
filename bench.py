#!/usr/bin/env python
"""bench.py — the hot path's headline benchmark (contract in the task brief).

    python bench.py --gpus N --steps K --warmup W            # own arm (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N ...             # the UNMODIFIED reference on the host cores, rank 0 only

Headline workload (BASELINE.json configs[2], the configuration the metric is quoted on; it fits one GPU):
SDXL LoRA rank 32, bf16, 1024 px (4x128x128 latents), batch 4 per GPU, min-SNR-5 weighting, AdamW,
global-norm clip 1.0, data parallel (weak scaling).  The wrapped network is an attention-only SDXL-topology STAND-IN
(onetrainer_b200/harness/standin.py: the 722 hooked leaves at their true shapes, ResNet blocks replaced by strided
convolutions) — the UNet itself is outside the path (BASELINE.json north_star).  A step is ONE pass of the hot path over
one synthetic batch: fused noise/timestep/add-noise -> the 722 hooked LoRA leaves forward -> fused loss fwd+bwd ->
leaves backward -> [gradient exchange] -> grad-norm -> fused clip+AdamW -> bf16 shadow refresh.

One JSON line on stdout (rank 0).  `value`: device-timed throughput with inputs resident in HBM; `e2e`: the same loop
through the public API with HOST batches (pinned H2D copy + D2H loss read per step).  Beside the headline the line carries
compact sub-records for the other BASELINE.json configurations (`configs`), the step with the reference's default gradient
checkpointing (`checkpointing`), the unmodified reference on the same GPU (`gpu_reference`), and at N > 1 a self-check of
the gradient exchange (`exchange_check`).  `--no-extras` prints the headline only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

WORKLOAD = "sdxl_lora_r32_bf16_1024px_b4_minsnr"
METRIC = "sdxl_lora_train_steps_per_sec"
UNIT = "steps/s (batch-4 steps, summed over ranks)"
BATCH, RANK, LATENT = 4, 32, 128
LORA_PARAMS = 85_114_880          # 722 leaves of the SDXL topology at rank 32 (checked against the built model)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["own", "reference"], default="own")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline + roofline only (no sub-records for the other configs)")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured CUDA graph")
    return ap.parse_args()


def config_dict(world: int) -> dict:
    """The workload description both arms print (the reference arm must carry the own arm's config)."""
    return {"workload": WORKLOAD, "network": "attention-only SDXL stand-in (722 hooked leaves at true shapes; harness/standin.py)",
            "per_gpu_batch": BATCH, "global_batch": BATCH * world, "lora_rank": RANK, "hooked_leaves": 722, "lora_params": LORA_PARAMS,
            "parallelism": f"dp{world}",
            "l2": "inputs larger than L2 (4.4 GB of frozen weights + 0.68 GB LoRA/optimizer state stream every step)"}


def measured_peaks() -> tuple[dict, str]:
    f = REPO / "MEASURED_PEAKS.json"
    if f.exists():
        return json.loads(f.read_text()), "measured"
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# the reference on the host cores
# ---------------------------------------------------------------------------------------------------
CPU_SAMPLE_LATENT = 32        # the bounded sample: one image at 32x32 latents = 1/64 of the tokens of a batch-4 1024-px step


class CpuReference:
    """The UNMODIFIED reference (baseline/_ref, driven by baseline/ref_arm.py) on the host cores, fp32 as in BASELINE.json
    configs[0]: its LoRAModuleWrapper hooked on the SDXL stand-in, its noise / add-noise / loss mixins, nn.utils
    clip_grad_norm_, its patched AdamW over all 1444 LoRA tensors.

    A full batch-4 1024-px step takes this path many minutes on a CPU (measured in the build container, 8 cores: 37 s
    for ONE image at 64x64 latents), so one timed "step" is a BOUNDED SAMPLE: the whole step — every one of the 722
    leaves forward and backward through the whole network, loss, clip, optimizer over all 85.1 M parameters — on one
    image at 32x32 latents, i.e. 1/64 of the tokens of the headline step.  The headline-equivalent step time is
    64 x (forward+backward of the sample) + (clip + optimizer of the sample, which does not depend on the batch): a
    LOWER bound for the reference (its attention cost grows quadratically with the tokens per image, and its measured
    forward+backward grows 7x from 32x32 to 64x64 latents), i.e. the extrapolation favours the CPU arm."""

    def __init__(self, threads: int | None = None):
        import torch
        from baseline import ref_arm
        from onetrainer_b200.harness import trainer
        self.torch, self.ref_arm = torch, ref_arm
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        R = ref_arm.load()
        self.cfg = ref_arm.reference_config(train_device="cpu", batch_size=1, lora_rank=RANK, lora_alpha=1.0,
                                            loss_weight_fn=R.LossWeight.MIN_SNR_GAMMA, clip_grad_norm=1.0)
        t0 = time.perf_counter()
        self.model, self.setup = ref_arm.build_reference_model("sdxl", self.cfg, "cpu", train_dtype=torch.float32,
                                                               stochastic_rounding=True, lora_up_std=0.02)
        self.build_seconds = time.perf_counter() - t0
        assert sum(p.numel() for p in self.model.parameters) == LORA_PARAMS
        self.batch = trainer.synthetic_batch("sdxl", 1, "cpu", latent_hw=CPU_SAMPLE_LATENT, rank=0)
        self.token_fraction = (1 * CPU_SAMPLE_LATENT * CPU_SAMPLE_LATENT) / (BATCH * LATENT * LATENT)

    def step(self) -> dict:
        tm = {}
        t0 = time.perf_counter()
        loss = self.ref_arm.reference_train_step(self.setup, self.model, self.batch, self.cfg, timings=tm)
        tm["sample_seconds"] = time.perf_counter() - t0
        tm["loss"] = loss
        return tm

    def summarize(self, samples: list[dict]) -> dict:
        fb = statistics.median(s["forward_backward"] for s in samples)
        opt = statistics.median(s["clip_optimizer_ema"] for s in samples)
        sample = statistics.median(s["sample_seconds"] for s in samples)
        full = fb / self.token_fraction + opt
        return {"value": 1.0 / full, "unit": UNIT, "cores": self.threads, "kind": "reference",
                "sample": f"the unmodified reference (baseline/_ref: LoRAModuleWrapper, noise/add-noise/loss mixins, patch_adamw; fp32, torch CPU, "
                          f"{self.threads} threads) running its WHOLE step on the SDXL stand-in — all 722 leaves forward+backward, min-SNR loss, "
                          f"clip_grad_norm_, AdamW over all {LORA_PARAMS} parameters — on ONE image at {CPU_SAMPLE_LATENT}x{CPU_SAMPLE_LATENT} latents "
                          f"(1/{round(1 / self.token_fraction)} of the tokens of the batch-{BATCH} {LATENT}x{LATENT} step)",
                "sample_ms_per_step": sample * 1e3, "sample_steps": len(samples),
                "sample_parts_ms": {"forward_backward": fb * 1e3, "clip_optimizer": opt * 1e3},
                "extrapolation": {"ms_per_headline_step": full * 1e3, "rule": "forward_backward / token_fraction + clip_optimizer",
                                  "token_fraction": self.token_fraction,
                                  "bias": "lower bound for the reference: its per-token cost grows with resolution (attention)"},
                "model_build_seconds": self.build_seconds}


def reference_arm(args) -> None:
    """`--impl reference`: the reference's own CPU implementation of the path on the host cores, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from baseline import ref_arm
    if not ref_arm.available():
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref is missing (tools/make_baseline_ref.py needs /root/reference)"}))
        return
    t0 = time.perf_counter()
    ref = CpuReference()
    for _ in range(args.warmup):
        ref.step()
    t1 = time.perf_counter()
    samples = [ref.step() for _ in range(max(args.steps, 1))]
    timed = time.perf_counter() - t1
    cpu = ref.summarize(samples)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": timed / max(args.steps, 1) * 1e3,
        "ms_per_step_is": "the bounded sample actually timed (see cpu_baseline.sample); `value` is the headline-equivalent rate from cpu_baseline.extrapolation",
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_dict(args.gpus), "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "loss": samples[-1]["loss"], "wall_seconds": time.perf_counter() - t0,
    }))


# ---------------------------------------------------------------------------------------------------
# own arm: helpers
# ---------------------------------------------------------------------------------------------------
def leaf_flops(wrapper, recompute: bool = False) -> float:
    """Algorithmic FLOPs of the fused LoRA GEMMs in one step: per leaf 4MKN + 6Mr(K+N); a leaf whose
    input takes no gradient (context K/V) has no dX GEMM, here as in the reference's autograd: 2MKN + 4Mr(K+N) + 2MrN.
    `recompute`: one more forward of every leaf (gradient checkpointing)."""
    total = 0.0
    for m in wrapper.lora_modules.values():
        M = getattr(m, "_last_rows", 0)
        w = m.orig_module.weight
        N, K, r = w.shape[0], w[0].numel(), m.rank
        fwd = 2.0 * M * K * N + 2.0 * M * r * (K + N)
        total += fwd + 2.0 * M * r * (K + N)                              # forward, dA, dB
        total += fwd if getattr(m, "_last_input_grad", True) else 2.0 * M * r * N
        if recompute:
            total += fwd
    return total


def _graph_time(fn, dev, reps: int = 3) -> float:
    """Device seconds of one call of `fn` (a launch sequence), captured once and replayed so that CUDA events bracket
    device time only (eager launches of 10-30 us kernels are host-bound and would under-report the kernel)."""
    import torch
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e3 / reps
    del graph
    return sec


def time_leaf_kernels(model, dev) -> dict:
    """Replay every K-a launch of one step — forward and dX GEMMs (`lora_linear_kernel_v3`, `_v2`), then the dA/dB kernel
    (`lora_grad_batched_kernel`) — grouped exactly as the step launches them (LoRAModuleWrapper.launch_units: context
    launches, sibling groups, single leaves; dA/dB 32 leaves per launch), each leaf with its own frozen weight /
    shadows, so the 4.4 GB of weights stream from HBM as in the real step.  Returns device seconds and algorithmic work."""
    import torch
    from onetrainer_b200 import ops
    from onetrainer_b200.modules.module.LoRAModule import _gs_only
    g = torch.Generator(device=dev).manual_seed(0)
    acts = {}

    def act(M, C, i=0):
        if (M, C, i) not in acts:
            acts[(M, C, i)] = (torch.randn(M, C, device=dev, generator=g) * 0.05).bfloat16()
        return acts[(M, C, i)]

    POOL = 12      # dA / dB run 32 leaves after the GEMMs that produced their operands: by then x / dY have left the 126 MB L2,
    #                so their replay rotates through POOL distinct operand sets per shape (>= 250 MB) instead of re-reading one

    plan = []
    grad_leaves = []
    gemm_flops, grad_flops, grad_bytes = 0.0, 0.0, 0.0
    for unit in model.unet_lora.launch_units():
        m0 = unit[0]
        w = m0.orig_module.weight
        M, N, K, r = getattr(m0, "_last_rows", 0), w.shape[0], w[0].numel(), m0.rank
        if M == 0:
            continue
        # a leaf whose input took no gradient in the real step (context K/V): no dX GEMM is issued, only Gs (see _gs_only)
        no_dx = not getattr(m0, "_last_input_grad", True)
        entry = []
        for i, m in enumerate(unit):
            sh = m._operands()
            entry.append((act(M, K), act(M, N, i % 3), sh, m._scale()))
            slot = 100 + len(grad_leaves) % POOL
            t = act(M, r, slot)
            grad_leaves.append((act(M, K, slot), act(M, N, slot + 1000), t, t, torch.zeros(r, K, device=dev), torch.zeros(N, r, device=dev)))
            gemm_flops += 2.0 * M * K * N + 2.0 * M * r * (K + N)      # forward
            gemm_flops += 2.0 * M * r * N if no_dx else 2.0 * M * K * N + 2.0 * M * r * (K + N)   # Gs only, or dX (+ Gs)
            grad_flops += 2.0 * M * r * (K + N)                        # dA + dB
            grad_bytes += 2.0 * M * (K + N) + 4.0 * M * r + 8.0 * r * (K + N)   # stream x, dY, Gs, Ts once; fp32 RMW of dA, dB
        plan.append((entry, no_dx))

    def gemms():
        for e, no_dx in plan:
            ops.lora_linear_grouped([m[0] for m in e], [m[2]["w"] for m in e], [m[2]["a"] for m in e], [m[2]["b"] for m in e],
                                    [m[2]["bias"] for m in e], [m[3] for m in e])
            if no_dx:
                _gs_only([m[1] for m in e], [m[2] for m in e], [m[3] for m in e], 0.0, [(0, 0)] * len(e))
            else:
                ops.lora_linear_grouped([m[1] for m in e], [m[2]["wt"] for m in e], [m[2]["bt"] for m in e], [m[2]["at"] for m in e],
                                        [None] * len(e), [m[3] for m in e])

    batch = ops.grad_queue.FLUSH_LEAVES

    def grads():
        for lo in range(0, len(grad_leaves), batch):
            ops.lora_grad_batched(grad_leaves[lo:lo + batch])

    return {"gemm_sec": _graph_time(gemms, dev), "gemm_flops": gemm_flops, "gemm_launches": 2 * len(plan),
            "grad_sec": _graph_time(grads, dev), "grad_flops": grad_flops, "grad_bytes": grad_bytes,
            "grad_launches": (len(grad_leaves) + batch - 1) // batch, "grad_leaves": len(grad_leaves)}


def time_hbm_kernels(dev, hbm_peak: float) -> list:
    """The memory-bound kernels at their roofline size (2^28 elements; SURVEY.md §8d: at the training configuration they
    move 1-10 MB and are latency-bound), CUDA events around 5 launches after 3 warm-ups.  Algorithmic bytes per element:
    K-b1 10 (read fp32 latent, write bf16 noisy + fp32 target), K-b2 8 (read bf16 prediction + fp32 target, write bf16
    gradient), K-c bf16+SR+EMA 18."""
    import torch
    from onetrainer_b200 import _native as N, ops
    out = []

    def timed(fn, iters=5):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 1e3 / iters

    T = 1000
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, T, dtype=torch.float32) ** 2
    ac = torch.cumprod(1 - betas, 0)
    sa, s1 = ac.sqrt().to(dev), (1 - ac).sqrt().to(dev)
    Bb = 4096
    lat = torch.randn(Bb, 4, 128, 128, device=dev)
    sec = timed(lambda: ops.noise_addnoise(latent=lat, coef_x0=sa, coef_noise=s1, B=Bb, C_=4, hw=128 * 128, latent_dtype=torch.float32,
                                           device=dev, cast_dtype=torch.bfloat16, target=N.TARGET_EPSILON, ts_mode=N.TS_UNIFORM, seed=1,
                                           shape=lat.shape))
    out.append(("noise_addnoise_kernel (K-b1)", lat.numel() * 10 / sec / 1e9))
    pred = torch.randn(Bb, 4, 128, 128, device=dev, dtype=torch.bfloat16)
    ts = torch.randint(0, 1000, (Bb,), device=dev, dtype=torch.int32)
    sec = timed(lambda: ops.loss_fwd_bwd(pred, lat, timestep=ts, sqrt_ac=sa, sqrt_1mac=s1, loss_weight_fn=N.LW_MIN_SNR, upstream=1 / Bb))
    out.append(("loss_fwd_bwd_kernel (K-b2)", pred.numel() * 8 / sec / 1e9))
    del lat, pred
    n = 1 << 28
    p = torch.zeros(n, device=dev, dtype=torch.bfloat16); g = torch.full_like(p, 1e-3); m = torch.zeros_like(p); v = torch.ones_like(p); e = torch.zeros_like(p)
    tab = ops.OptimTable([dict(p=p, g=g, exp_avg=m, exp_avg_sq=v, ema=e)], dev)
    tab.host["flags"][:] |= N.OPT_FLAG_SR
    for k, val in ops.adam_scalars(3e-4, 0.9, 0.999, 1e-8, 1e-2, 1.0, True).items():
        tab.host[k][:] = val
    tab.upload()
    sec = timed(lambda: ops.adamw_sr_ema_multi(tab, do_ema=True, ema_one_minus_decay=0.001, seed=1, offset=2))
    out.append(("adamw_sr_ema_kernel (K-c, bf16 + SR + EMA)", n * 18 / sec / 1e9))
    return [{"kernel": k, "bound": "hbm", "achieved": a, "peak": hbm_peak, "unit": "GB/s", "frac": a / hbm_peak, "elements": 1 << 28}
            for k, a in out]


def profiled_traffic() -> tuple[float | None, str | None]:
    """DRAM bytes (read + write) of ONE launch of the dominant kernel from the committed `ncu --set full` summary
    (newest profiles/r0*_ncu_ka*.md: M=4096, N=10240, K=1280, r=32) — per launch, not per step."""
    import re
    for name in ("r02final_ncu_ka.md", "r02_ncu_ka.md", "r01b_ncu_ka_v3.md"):
        f = REPO / "profiles" / name
        if not f.exists():
            continue
        m = re.search(r"traffic = read \+ write = ([0-9.]+) (\w+) \+ ([0-9.]+) (\w+)", f.read_text())
        if not m:
            continue
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        return float(m.group(1)) * unit[m.group(2)] + float(m.group(3)) * unit[m.group(4)], \
            f"one launch at M=4096 N=10240 K=1280 r=32 (profiles/{name}); algorithmic 120.7 MB incl. the output write"
    return None, None


class Timer:
    """CUDA-event timing of K calls bracketed by barrier + synchronize on both sides, max over ranks."""

    def __init__(self, dev, world):
        self.dev, self.world = dev, world

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def __call__(self, fn, k):
        import torch
        import torch.distributed as dist
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(k):
            out = fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / 1e3, out


def pinned(batch: dict) -> dict:
    return {k: ([t.pin_memory() for t in v] if isinstance(v, list) else v.pin_memory()) for k, v in batch.items()}


def batch_bytes(batch: dict) -> int:
    return sum(sum(t.numel() * t.element_size() for t in v) if isinstance(v, list) else v.numel() * v.element_size() for v in batch.values())


# ---------------------------------------------------------------------------------------------------
# own arm: data-parallel self-check (N > 1)
# ---------------------------------------------------------------------------------------------------
def exchange_check(model, setup, cfg, dev, rank: int, world: int) -> dict:
    """One-shot parity of the data-parallel path before anything is timed (every rank takes part):
    (1) K-r (otb_grad_exchange: reduce-scatter + squared norm + all-gather over NVLink peer memory) against ncclAllReduce on a
        clone of the same arena; (2) every replica identical afterwards; (3) the norm K-r accumulated on the fly against a plain
        norm of the averaged gradient; (4) the step-0 loss of N ranks at per-rank batch b against ONE GPU computing the global
        batch N*b (DP-invariant Philox keying: global sample indices)."""
    import torch
    import torch.distributed as dist
    from onetrainer_b200 import dp, ops
    from onetrainer_b200.harness import trainer
    out = {}
    # first-step loss: N ranks x b  vs  rank 0 alone on the global batch
    with torch.no_grad():
        local = trainer.batch_to_device(trainer.synthetic_batch("sdxl", BATCH, dev), dev)
        data = setup.predict(model, local, cfg, model.train_progress)
        mine = setup.calculate_loss(model, local, data, cfg).float()
        mean = dp.average_scalar(mine.clone())
        if rank == 0:
            parts = [trainer.synthetic_batch("sdxl", BATCH, dev, rank=r) for r in range(world)]
            glob = {k: ([torch.cat([p[k][i] for p in parts]) for i in range(len(v))] if isinstance(v, list) else torch.cat([p[k] for p in parts]))
                    for k, v in parts[0].items()}
            glob = trainer.batch_to_device(glob, dev)
            gcfg = type(cfg)(**{**cfg.__dict__, "batch_size": BATCH * world})
            ws, dist_rank = dp.world_size, dp.rank
            dp.world_size, dp.rank = (lambda: 1), (lambda: 0)          # the single-process view of the same step
            try:
                gdata = setup.predict(model, glob, gcfg, model.train_progress)
                one = setup.calculate_loss(model, glob, gdata, gcfg).float().item()
            finally:
                dp.world_size, dp.rank = ws, dist_rank
            out["loss_n_ranks"] = mean.item()
            out["loss_one_gpu_global_batch"] = one
            out["loss_rel_diff"] = abs(mean.item() - one) / max(abs(one), 1e-12)
    # K-r vs NCCL on the gradient arena
    arena = model.unet_lora.grad_arena
    if arena is None or not dp.is_symmetric(arena):
        out["exchange"] = "ncclAllReduce (K-r not active: OTB_PEER_EXCHANGE=0 or no symmetric arena)"
        out["ok"] = bool(out.get("loss_rel_diff", 0.0) < 1e-3)
        return out
    saved = arena.clone()
    g = torch.Generator(device=dev).manual_seed(4242 + rank)
    arena.copy_(torch.randn(arena.shape, device=dev, generator=g) * 1e-3)
    expect = arena.clone()
    dist.all_reduce(expect, op=dist.ReduceOp.SUM)
    bucket = dp.GradientBucket(arena)
    prescale = bucket.allreduce([arena])
    torch.cuda.synchronize()
    diff = (arena - expect).abs().max().item()
    out["exchange"] = f"K-r {bucket._exchange.mode}"
    out["kr_vs_nccl_max_abs"] = diff
    out["kr_vs_nccl_bit_equal"] = bool(torch.equal(arena, expect))
    out["kr_vs_nccl_rel"] = diff / max(expect.abs().max().item(), 1e-30)
    digest = torch.stack([arena.double().sum(), arena.view(torch.int32).long().sum().double()])
    all_d = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    out["replicas_identical"] = bool(all(torch.equal(d, all_d[0]) for d in all_d))

    class _T:
        norm_coef = torch.zeros((2,), dtype=torch.float32, device=dev)
    nc = ops.grad_norm_from_partials(_T, bucket.norm_partials, 1.0, prescale).clone()
    plain = (expect.double() * prescale).norm().item()
    out["fused_norm"], out["plain_norm"] = nc[0].item(), plain
    out["fused_norm_rel_diff"] = abs(nc[0].item() - plain) / max(plain, 1e-30)
    arena.copy_(saved)
    torch.cuda.synchronize()
    dist.barrier()
    out["ok"] = bool(out["kr_vs_nccl_rel"] < 1e-6 and out["replicas_identical"] and out["fused_norm_rel_diff"] < 1e-5
                     and out.get("loss_rel_diff", 0.0) < 1e-3)
    return out


# ---------------------------------------------------------------------------------------------------
# own arm: sub-records for the other BASELINE.json configurations
# ---------------------------------------------------------------------------------------------------
def run_config(name: str, build, batch_fn, cfg, dev, world, steps: int, warmup: int, graph: bool = True) -> dict:
    """Build a model, capture its step, time `steps` replays with resident inputs.  Returns a compact record."""
    import gc
    import torch
    from onetrainer_b200 import ops
    from onetrainer_b200.harness import trainer
    timer = Timer(dev, world)
    t0 = time.perf_counter()
    model, setup = build()
    host = pinned(batch_fn())
    rec = {"config": name}
    try:
        if graph:
            stepper = trainer.GraphedTrainStep(setup, model, cfg, host)
            fn = stepper.step
        else:
            devb = trainer.batch_to_device(host, dev)
            fn = lambda: trainer.train_step(setup, model, devb, cfg)   # noqa: E731
            stepper = None
        for _ in range(max(warmup, 3)):
            fn()
        sec, loss = timer(fn, steps)
        rec.update(steps_per_sec=world * steps / sec, ms_per_step=sec / steps * 1e3, samples_per_sec=world * steps * cfg.batch_size / sec,
                   loss=float(loss), cuda_graph=graph, trainable_params=sum(p.numel() for p in model.parameters),
                   peak_mem_gb=torch.cuda.max_memory_allocated(dev) / 2 ** 30, setup_seconds=time.perf_counter() - t0 - sec)
        if stepper is not None:
            rec["gpu_launches_per_step"] = stepper.launches_per_step
            stepper.close()
        return rec, model, setup
    finally:
        gc.collect()


def free_model(*objs) -> None:
    import gc
    import torch
    for o in objs:
        del o
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


def optimizer_in_step_gbs(model, dev, steps: int = 3) -> dict:
    """K-c at the configuration's real size (all tensors of the model, their own state), CUDA events around the optimizer's
    device step: bytes = 14 B per bf16 element (+4 on EMA steps), 28 B per fp32 element."""
    import torch
    drv = model.optimizer.otb_driver
    n16 = sum(p.numel() for p in model.parameters if p.dtype == torch.bfloat16)
    n32 = sum(p.numel() for p in model.parameters if p.dtype == torch.float32)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    secs, do_ema = [], False
    for _ in range(steps):
        if not drv.host_prepare() or drv.table is None:
            return {"seconds": float("nan"), "gbs": float("nan"), "bytes": 0, "includes": "no gradients to step"}
        do_ema = bool(drv._dyn[3])
        drv.table.upload()
        e0.record()
        drv.device_step(upload=False)
        e1.record()
        torch.cuda.synchronize()
        secs.append(e0.elapsed_time(e1) / 1e3)
    sec = min(secs)
    byts = n16 * (18 if do_ema else 14) + n32 * (36 if do_ema else 28) + (n16 * 2 + n32 * 4 if drv.clip_max_norm is not None else 0)
    return {"seconds": sec, "gbs": byts / sec / 1e9, "bytes": byts,
            "includes": "grad-norm pass (one more read of every gradient) + clip + AdamW" + (" + EMA" if do_ema else " (not an EMA step)")}


def gpu_reference_record(own_model, cfg, dev, own_segments: dict, steps: int = 5) -> dict:
    """The UNMODIFIED reference (baseline/_ref) on the SAME B200, same stand-in weights, bf16 autocast, eager — SURVEY.md §2.2's
    bar ("same Python ops under torch").  Whole step and per segment, beside ours."""
    import torch
    from baseline import ref_arm
    from onetrainer_b200.harness import trainer
    R = ref_arm.load()
    rcfg = ref_arm.reference_config(train_device=str(dev), batch_size=BATCH, lora_rank=RANK, lora_alpha=1.0,
                                    loss_weight_fn=R.LossWeight.MIN_SNR_GAMMA, clip_grad_norm=1.0)
    own_model.unet_lora.remove_hook_from_module()
    try:
        # fp32 LoRA weights as in the own arm: the reference's stochastic rounding only acts on bf16 parameters
        model, setup = ref_arm.build_reference_model("sdxl", rcfg, dev, train_dtype=torch.bfloat16, stochastic_rounding=True,
                                                     lora_up_std=0.02, network=own_model.unet)
        batch = trainer.batch_to_device(trainer.synthetic_batch("sdxl", BATCH, dev), dev)
        for _ in range(2):
            ref_arm.reference_train_step(setup, model, batch, rcfg)
        torch.cuda.synchronize()
        tm = {}
        t0 = time.perf_counter()
        for _ in range(steps):
            loss = ref_arm.reference_train_step(setup, model, batch, rcfg, timings=tm)
        torch.cuda.synchronize()
        step_s = (time.perf_counter() - t0) / steps
        # segments: noise + timestep + add-noise + loss forward/backward on the step's tensors
        lat = batch["latent_image"]
        pred = torch.randn(lat.shape, device=dev, dtype=torch.bfloat16, requires_grad=True)

        def noise_loss():
            gen = torch.Generator(device=dev).manual_seed(1)
            x0 = lat * 0.13025
            noise = setup._create_noise(x0, rcfg, gen)
            t = setup._get_timestep_discrete(1000, False, gen, BATCH, rcfg)
            setup._add_noise_discrete(x0, noise, t, model.noise_scheduler.betas).to(torch.bfloat16)
            data = {"loss_type": "target", "timestep": t, "predicted": pred, "target": noise, "prediction_type": "epsilon"}
            setup._diffusion_losses(batch=batch, data=data, config=rcfg, train_device=dev, betas=model.noise_scheduler.betas).mean().backward()

        # the reference's 722 leaves alone (LoRAModule.forward + autograd backward per leaf, bf16 autocast), the launches of
        # one step captured in a CUDA graph so that DEVICE time is compared with device time (no python dispatch on either side)
        shapes = {k: (getattr(m, "_last_rows", 0), getattr(m, "_last_input_grad", True)) for k, m in own_model.unet_lora.lora_modules.items()}
        gl = torch.Generator(device=dev).manual_seed(3)
        pool = {}

        def buf(M, C, grad):
            key = (M, C, grad)
            if key not in pool:
                pool[key] = (torch.randn(M, C, device=dev, generator=gl) * 0.05).bfloat16().requires_grad_(grad)
            return pool[key]

        def ref_leaves():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                for k, m in model.unet_lora.lora_modules.items():
                    M, dx = shapes[k]
                    om = m.orig_module
                    x = buf(M, om.in_features, dx)
                    y = om(x)
                    torch.autograd.grad(y, ([x] if dx else []) + [m.lora_down.weight, m.lora_up.weight], buf(M, om.out_features, False))
        try:
            leaves_dev = _graph_time(ref_leaves, dev, 2) * 1e3
        except Exception as e:  # noqa: BLE001
            leaves_dev = f"{type(e).__name__}: {e}"[:200]

        def wall(fn, n=5):
            fn()
            torch.cuda.synchronize()
            t = time.perf_counter()
            for _ in range(n):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t) / n

        rec = {"what": "the unmodified reference modules (baseline/_ref) hooked on the same stand-in on this GPU: bf16 autocast, eager, "
                       "its own LoRAModule.forward / mixins / clip_grad_norm_ / patched AdamW (python loop over 1444 tensors)",
               "ms_per_step": step_s * 1e3, "steps_per_sec": 1.0 / step_s, "loss": loss,
               "segments_ms": {"reference": {"whole_network_forward_backward_wall": tm["forward_backward"] / steps * 1e3,
                                             "leaves_forward_backward_device": leaves_dev,
                                             "clip_optimizer_wall": tm["clip_optimizer_ema"] / steps * 1e3,
                                             "noise_addnoise_loss_wall": wall(noise_loss) * 1e3},
                               "ours": own_segments},
               "segments_note": "leaves: the 722 hooked leaves alone, forward + backward, one step's launches replayed as a CUDA graph on both "
                                "sides (reference: cuBLAS GEMMs + elementwise per leaf under autocast; ours: K-a + K-a'); the other reference "
                                "segments are host-bound eager code, timed by wall clock around synchronised regions",
               "timing": "wall clock around synchronised regions (the reference's eager step is host-bound; its own loss.item() syncs every step)"}
        model.unet_lora.remove_hook_from_module()
        del model, setup
        return rec
    finally:
        own_model.unet_lora.hook_to_module()


def own_segments_ms(model, setup, cfg, dev, ka: dict) -> dict:
    """Our side of the segment table: device time of the K-a + K-a' launches of one step (graph replay), of the fused
    noise/add-noise + loss fwd+bwd launches, and of grad-norm + clip + AdamW over the step's 1444 tensors."""
    import torch
    from onetrainer_b200 import _native as N, ops
    lat = torch.randn(BATCH, 4, LATENT, LATENT, device=dev)
    pred = torch.randn(BATCH, 4, LATENT, LATENT, device=dev, dtype=torch.bfloat16)
    betas = model.noise_scheduler.betas
    ac = torch.cumprod(1 - betas, 0)
    sa, s1 = ac.sqrt().contiguous(), (1 - ac).sqrt().contiguous()

    def noise_loss():
        o = ops.noise_addnoise(latent=lat, coef_x0=sa, coef_noise=s1, B=BATCH, C_=4, hw=LATENT * LATENT, latent_dtype=torch.float32, device=dev,
                               cast_dtype=torch.bfloat16, target=N.TARGET_EPSILON, ts_mode=N.TS_UNIFORM, latent_scale=0.13025, seed=1, shape=lat.shape)
        ops.loss_fwd_bwd(pred, o["target"], timestep=o["timestep"], sqrt_ac=sa, sqrt_1mac=s1, loss_weight_fn=N.LW_MIN_SNR, upstream=1 / BATCH)
    opt = optimizer_in_step_gbs(model, dev)
    return {"leaves_forward_backward_device": (ka["gemm_sec"] + ka["grad_sec"]) * 1e3, "noise_addnoise_loss_device": _graph_time(noise_loss, dev, 10) * 1e3,
            "clip_optimizer_device": opt["seconds"] * 1e3}


# ---------------------------------------------------------------------------------------------------
def own_arm(args) -> None:
    import torch
    import torch.distributed as dist
    from onetrainer_b200 import dp, ops
    from onetrainer_b200.harness import standin, trainer
    from onetrainer_b200.modules.util.config.TrainConfig import TrainConfig
    from onetrainer_b200.modules.util.enum import LossWeight, TimestepDistribution

    rank, local_rank, world = dp.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (own arm) needs a B200: there is no CPU fallback for this path")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    timer = Timer(dev, world)
    B = BATCH
    cfg = TrainConfig(train_device=str(dev), batch_size=B, lora_rank=RANK, lora_alpha=1.0, loss_weight_fn=LossWeight.MIN_SNR_GAMMA,
                      clip_grad_norm=1.0)
    model, setup = trainer.build_model("sdxl", cfg, dev, ema=False, stochastic_rounding=True, lora_up_std=0.02)
    assert sum(p.numel() for p in model.parameters) == LORA_PARAMS
    host_batch = pinned(trainer.synthetic_batch("sdxl", B, dev))
    dev_batch = trainer.batch_to_device(host_batch, dev)
    bucket = dp.GradientBucket(None)
    h2d = batch_bytes(host_batch)

    xcheck = None
    if world > 1:
        trainer.train_step(setup, model, dev_batch, cfg, bucket)      # creates the arena and the exchange (collective)
        xcheck = exchange_check(model, setup, cfg, dev, rank, world)

    graphed = None
    if not args.no_graph:
        graphed = trainer.GraphedTrainStep(setup, model, cfg, host_batch)

    def step_resident():
        if graphed is not None:
            return graphed.step()                      # batch already staged; device-side work only
        return trainer.train_step(setup, model, dev_batch, cfg, bucket)

    def step_e2e():
        if graphed is not None:
            return graphed.step(host_batch).item()     # host batch -> pinned staging -> H2D inside the graph -> D2H loss
        b = trainer.batch_to_device(host_batch, dev, non_blocking=True)
        return trainer.train_step(setup, model, b, cfg, bucket).item()

    l0 = ops.LAUNCHES
    if graphed is None:
        trainer.train_step(setup, model, dev_batch, cfg, bucket)
    launches_eager_step = ops.LAUNCHES - l0
    for _ in range(max(args.warmup, 3)):
        step_resident()
    with ClockSampler(local_rank) as clk:
        sec, last_loss = timer(step_resident, args.steps)
    launches = (graphed.launches_per_step if graphed is not None else launches_eager_step) * args.steps
    clocks = clk.summary()
    for _ in range(2):
        step_e2e()
    sec_e2e, loss_e2e = timer(step_e2e, args.steps)

    peaks, peak_kind = measured_peaks()
    ka = time_leaf_kernels(model, dev)
    achieved = ka["gemm_flops"] / ka["gemm_sec"] / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    hbm_peak = float(peaks["hbm_gbs"])
    hbm_kernels = time_hbm_kernels(dev, hbm_peak) if rank == 0 else []
    grad_gbs = ka["grad_bytes"] / ka["grad_sec"] / 1e9
    hbm_kernels.insert(0, {"kernel": f"lora_grad_batched_kernel (K-a', dA/dB; all {ka['grad_leaves']} leaves of one step, 32 per launch)", "bound": "hbm",
                           "achieved": grad_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": grad_gbs / hbm_peak,
                           "seconds_per_step_in_kernel": ka["grad_sec"], "launches_per_step": ka["grad_launches"]})
    traffic, traffic_of = profiled_traffic()
    value = world * args.steps / sec
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": config_dict(world),
        "samples_per_sec": value * B, "optimizer_steps_per_sec": args.steps / sec,
        "e2e": {"value": world * args.steps / sec_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": sec_e2e / args.steps * 1e3},
        "gpu_launches": launches, "cuda_graph": graphed is not None, "loss": float(last_loss), "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "lora_linear_kernel_v3 / _v2 (K-a: the forward + dX launches of all 722 leaves; context K/V of "
                                                  "the whole network, q/k/v siblings grouped)",
                     "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_of": traffic_of,
                     "peak_source": f"{peak_kind} bf16_tflops_sustained",
                     "frac_of_burst": achieved / float(peaks.get("bf16_tflops", peak)), "frac_of_nominal_2250": achieved / 2250.0,
                     "algorithmic_flops_per_step": ka["gemm_flops"], "launches_per_step": ka["gemm_launches"],
                     "seconds_per_step_in_kernel": ka["gemm_sec"], "share_of_step": ka["gemm_sec"] / (sec / args.steps)},
        "hbm_kernels": hbm_kernels, "hbm_peak_source": f"{peak_kind} hbm_gbs",
    }
    if xcheck is not None:
        line["exchange_check"] = xcheck

    extras_ok = not args.no_extras
    if extras_ok:
        own_seg = own_segments_ms(model, setup, cfg, dev, ka) if rank == 0 else {}
        # ---- gradient checkpointing ON (the reference's default, TrainConfig.py:786): one more forward of every leaf ----
        try:
            if graphed is not None:
                graphed.close()
            del graphed
            graphed = None
            n_blocks = standin.enable_checkpointing(model.unet, preserve_rng_state=False)
            ck = trainer.GraphedTrainStep(setup, model, cfg, host_batch)
            for _ in range(3):
                ck.step()
            sec_ck, loss_ck = timer(ck.step, args.steps)
            fl = leaf_flops(model.unet_lora, recompute=True)
            line["checkpointing"] = {"blocks_checkpointed": n_blocks, "ms_per_step": sec_ck / args.steps * 1e3,
                                     "steps_per_sec": world * args.steps / sec_ck, "cuda_graph": True, "loss": float(loss_ck),
                                     "algorithmic_leaf_tflop_per_step": fl / 1e12,
                                     "algorithmic_leaf_tflop_per_step_without": leaf_flops(model.unet_lora) / 1e12,
                                     "gpu_launches_per_step": ck.launches_per_step,
                                     "note": "non-reentrant torch.utils.checkpoint around every transformer block (checkpointing_util.py:150-166)"}
            ck.close()
            del ck
        except Exception as e:  # noqa: BLE001
            line["checkpointing"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        standin.disable_checkpointing(model.unet)
        # ---- the unmodified reference on this GPU (rank 0, one GPU) ----
        if rank == 0 and world == 1:
            try:
                from baseline import ref_arm
                if ref_arm.available():
                    line["gpu_reference"] = gpu_reference_record(model, cfg, dev, own_seg)
                else:
                    line["gpu_reference"] = {"unavailable": "baseline/_ref is missing"}
            except Exception as e:  # noqa: BLE001
                line["gpu_reference"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        free_model(model, setup)
        model = setup = None
        # ---- the other BASELINE.json configurations ----
        subs = []
        k = max(3, min(args.steps, 10))

        def c2():
            c = TrainConfig(train_device=str(dev), batch_size=8, lora_rank=16, lora_alpha=1.0, masked_training=True, normalize_masked_area_loss=True,
                            unmasked_weight=0.1, clip_grad_norm=1.0)
            return run_config("C2 sd15_lora_r16_bf16_512px_b8_masked", lambda: trainer.build_model("sd15", c, dev, lora_up_std=0.02),
                              lambda: trainer.synthetic_batch("sd15", 8, dev, masked=True), c, dev, world, k, 3)

        def c4():
            c = TrainConfig(train_device=str(dev), batch_size=B, clip_grad_norm=1.0, ema_decay=0.999, ema_update_step_interval=5)
            return run_config("C4 sdxl_full_finetune_bf16_sr_adamw_ema_b4", lambda: trainer.build_model("sdxl", c, dev, ema=True, stochastic_rounding=True,
                                                                                                       full_finetune=True, lr=1e-5),
                              lambda: trainer.synthetic_batch("sdxl", B, dev), c, dev, world, k, 3)

        def c5():
            c = TrainConfig(train_device=str(dev), batch_size=1, lora_rank=16, lora_alpha=1.0, timestep_distribution=TimestepDistribution.LOGIT_NORMAL,
                            clip_grad_norm=1.0)
            return run_config("C5 flux_dev_lora_r16_bf16_1024px_b1_flow", lambda: trainer.build_flux_model("flux", c, dev, lora_up_std=0.02),
                              lambda: trainer.synthetic_batch("flux", 1, dev), c, dev, world, k, 3)

        for fn in (c2, c4, c5):
            try:
                rec, m, s = fn()
                if rec["config"].startswith("C4"):
                    red = getattr(m, "grad_reducer", None)
                    if world > 1 and isinstance(red, dp.ShardedExchange):
                        # K-z: reduce-scatter -> AdamW + EMA on the owner's tensors only -> all-gather (every rank takes part)
                        rec["exchange"] = {"kind": "K-z (otb_shard_reduce / otb_shard_gather over NVLink peer memory; optimizer state and EMA sharded "
                                                   f"1/{world}: {len(m.optimizer_parameters)} of {len(m.parameters)} tensors on this rank)", **red.measure()}
                    elif world > 1 and red is not None:
                        rec["exchange"] = {"kind": f"bucketed ncclAllReduce of the {m.full_grad_arena.flat.numel() * 2 / 2 ** 30:.2f} GiB bf16 gradient "
                                                   "arena, overlapped with backward; replicated optimizer"}
                    if rank == 0 and not (world > 1 and isinstance(red, dp.ShardedExchange)):
                        kc = optimizer_in_step_gbs(m, dev)
                        rec["kc_in_step"] = {"gbs": kc["gbs"], "frac_of_hbm": kc["gbs"] / hbm_peak, "ms": kc["seconds"] * 1e3, "bytes": kc["bytes"],
                                             "includes": kc["includes"]}
                subs.append(rec)
                free_model(m, s)
            except Exception as e:  # noqa: BLE001
                subs.append({"config": fn.__name__, "error": f"{type(e).__name__}: {e}"[:300]})
        line["configs"] = subs

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from baseline import ref_arm
            if ref_arm.available():
                ref = CpuReference()
                ref.step()
                line["cpu_baseline"] = ref.summarize([ref.step() for _ in range(3)])
            else:
                line["cpu_baseline"] = {"unavailable": "baseline/_ref is missing (tools/make_baseline_ref.py needs /root/reference)"}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        shutdown(dist)


def shutdown(dist) -> None:
    """Tear the process group down — after every captured graph that holds a collective has been destroyed (round 1 left
    with os._exit because destroy_process_group() blocked while a live CUDA graph still referenced the communicator).
    Guarded by a watchdog: if NCCL still does not come back within 30 s the rank exits anyway (all work is done and
    synchronised; the JSON line has been printed)."""
    import gc
    import torch
    gc.collect()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    done = threading.Event()

    def watchdog():
        if not done.wait(30.0):
            os._exit(0)
    threading.Thread(target=watchdog, daemon=True).start()
    try:
        dist.barrier()
        dist.destroy_process_group()
    finally:
        done.set()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        own_arm(a)
