#!/usr/bin/env python
"""bench.py — the hot path's headline benchmark (contract in the task brief).

    python bench.py --gpus N --steps K --warmup W            # own arm (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N ...             # the reference's CPU path (oracle port), rank 0 only

Workload (BASELINE.json configs[2], the configuration the metric is quoted on; it fits one GPU):
SDXL LoRA rank 32, bf16, 1024 px (4x128x128 latents), batch 4 per GPU, min-SNR-5 weighting, AdamW,
global-norm clip 1.0, data parallel (weak scaling).  A step is ONE pass of the hot path over one synthetic
batch: fused noise/timestep/add-noise -> the 722 hooked LoRA leaves of an SDXL-topology stand-in forward
-> fused loss fwd+bwd -> leaves backward -> [NCCL allreduce of the LoRA gradient arena] -> grad-norm ->
fused clip+AdamW -> bf16 shadow refresh.  Synthetic data, random-init weights (no network access).

One JSON line on stdout (rank 0).  `value`: device-timed throughput with inputs resident in HBM;
`e2e`: the same loop through the public API with HOST batches (pinned H2D copy + D2H loss read per step).
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["own", "reference"], default="own")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured CUDA graph")
    return ap.parse_args()


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
def reference_arm(args) -> None:
    """The reference's CPU implementation of the path on the host cores (oracle port), rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import cpu_baseline
    for _ in range(args.warmup):
        cpu_baseline.run_sample()
    t0 = time.perf_counter()
    samples = [cpu_baseline.run_sample() for _ in range(max(args.steps, 1))]
    wall = time.perf_counter() - t0
    step_s = statistics.median(s["step_seconds"] for s in samples)
    value = 1.0 / step_s
    s0 = samples[0]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "note": "bounded CPU sample extrapolated to one full step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": s0["threads"], "kind": "port", "sample": s0["sample"],
                         "sample_seconds": s0["sample_seconds"], "parts_seconds": s0["parts"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_seconds": wall,
    }))


# ---------------------------------------------------------------------------------------------------
def leaf_flops(model, recompute: bool = False) -> tuple[float, int, list]:
    """Algorithmic FLOPs of the fused LoRA GEMMs in one step: per leaf 4MKN + 6Mr(K+N); a leaf whose
    input takes no gradient (context K/V) has no dX GEMM, here as in the reference's autograd: 2MKN + 4Mr(K+N) + 2MrN."""
    total, launches, shapes = 0.0, 0, []
    for m in model.unet_lora.lora_modules.values():
        M = getattr(m, "_last_rows", 0)
        w = m.orig_module.weight
        N, K, r = w.shape[0], w[0].numel(), m.rank
        total += 2.0 * M * K * N + 4.0 * M * r * (K + N)              # forward, dA, dB
        total += 2.0 * M * K * N + 2.0 * M * r * (K + N) if getattr(m, "_last_input_grad", True) else 2.0 * M * r * N
        launches += 3
        shapes.append((M, N, K, r))
    return total, launches, shapes


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
    """Replay every K-a launch of one step — forward and dX GEMMs (`lora_linear_kernel*`), then the dA/dB kernel
    (`lora_grad_kernel`) — each leaf with its own frozen weight / shadows, so the 4.4 GB of weights stream from HBM as
    in the real step.  Returns device seconds and algorithmic work of both families."""
    import torch
    from onetrainer_b200 import ops
    from onetrainer_b200.modules.module.LoRAModule import _gs_only
    _, _, shapes = leaf_flops(model)
    g = torch.Generator(device=dev).manual_seed(0)
    acts = {}

    def act(M, C, i=0):
        if (M, C, i) not in acts:
            acts[(M, C, i)] = (torch.randn(M, C, device=dev, generator=g) * 0.05).bfloat16()
        return acts[(M, C, i)]

    mods = list(model.unet_lora.lora_modules.values())
    shape_of = {id(m): s for m, s in zip(mods, shapes)}
    # launch units exactly as the step issues them: sibling groups (q/k/v of a self-attention, k/v of a cross-attention)
    # are ONE grouped launch forward and one for their dX GEMMs; every other leaf is launched singly
    grouped = {}
    for grp in getattr(model.unet_lora, "sibling_groups", []):
        if grp.enabled:
            for m in grp.members:
                grouped[id(m)] = grp
    units, seen = [], set()
    for m in mods:
        grp = grouped.get(id(m))
        if grp is None:
            units.append([m])
        elif id(grp) not in seen:
            seen.add(id(grp))
            units.append(list(grp.members))
    plan = []
    gemm_flops, grad_flops, grad_bytes = 0.0, 0.0, 0.0
    for unit in units:
        M, N, K, r = shape_of[id(unit[0])]
        # a leaf whose input took no gradient in the real step (context K/V): no dX GEMM is issued, only Gs (see _gs_only)
        no_dx = not getattr(unit[0], "_last_input_grad", True)
        entry = []
        for i, m in enumerate(unit):
            sh = m._operands()
            entry.append((act(M, K), act(M, N, i), sh, m._scale(), torch.zeros(r, K, device=dev), torch.zeros(N, r, device=dev),
                          torch.zeros(M, r, device=dev, dtype=torch.bfloat16)))
            gemm_flops += 2.0 * M * K * N + 2.0 * M * r * (K + N)      # forward
            gemm_flops += 2.0 * M * r * N if no_dx else 2.0 * M * K * N + 2.0 * M * r * (K + N)   # Gs only, or dX (+ Gs)
            grad_flops += 2.0 * M * r * (K + N)                        # dA + dB
            grad_bytes += 2.0 * M * (K + N) + 4.0 * M * r + 8.0 * r * (K + N)   # stream x, dY, Gs, Ts once; fp32 RMW of dA, dB
        plan.append((entry, no_dx))
    n_leaves = sum(len(e) for e, _ in plan)

    def gemms():
        for e, no_dx in plan:
            ops.lora_linear_grouped([m[0] for m in e], [m[2]["w"] for m in e], [m[2]["a"] for m in e], [m[2]["b"] for m in e],
                                    [m[2]["bias"] for m in e], [m[3] for m in e])
            if no_dx:
                _gs_only([m[1] for m in e], [m[2] for m in e], [m[3] for m in e], 0.0, [(0, 0)] * len(e))
            else:
                ops.lora_linear_grouped([m[1] for m in e], [m[2]["wt"] for m in e], [m[2]["bt"] for m in e], [m[2]["at"] for m in e],
                                        [None] * len(e), [m[3] for m in e])

    def grads():
        for e, _ in plan:
            for x, dy, sh, sc, dA, dB, t in e:
                ops.lora_grad_ab(x, dy, t, t, dA, dB)

    return {"gemm_sec": _graph_time(gemms, dev), "gemm_flops": gemm_flops, "gemm_launches": 2 * len(plan),
            "grad_sec": _graph_time(grads, dev), "grad_flops": grad_flops, "grad_bytes": grad_bytes, "grad_launches": n_leaves}


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
    (profiles/r01b_ncu_ka_v3.md: M=4096, N=10240, K=1280, r=32) — per launch, not per step."""
    import re
    f = REPO / "profiles" / "r01b_ncu_ka_v3.md"
    if not f.exists():
        return None, None
    m = re.search(r"traffic = read \+ write = ([0-9.]+) (\w+) \+ ([0-9.]+) (\w+)", f.read_text())
    if not m:
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(m.group(1)) * unit[m.group(2)] + float(m.group(3)) * unit[m.group(4)], \
        "one launch at M=4096 N=10240 K=1280 r=32 (profiles/r01b_ncu_ka_v3.md); algorithmic 120.7 MB incl. the output write"


def own_arm(args) -> None:
    import torch
    import torch.distributed as dist
    from onetrainer_b200 import dp, ops
    from onetrainer_b200.harness import trainer
    from onetrainer_b200.modules.util.config.TrainConfig import TrainConfig
    from onetrainer_b200.modules.util.enum import LossWeight

    rank, local_rank, world = dp.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (own arm) needs a B200: there is no CPU fallback for this path")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B = 4
    cfg = TrainConfig(train_device=str(dev), batch_size=B, lora_rank=32, lora_alpha=1.0, loss_weight_fn=LossWeight.MIN_SNR_GAMMA,
                      clip_grad_norm=1.0)
    model, setup = trainer.build_model("sdxl", cfg, dev, ema=False, stochastic_rounding=True, lora_up_std=0.02)
    host_batch = {k: ([t.pin_memory() for t in v] if isinstance(v, list) else v.pin_memory())
                  for k, v in trainer.synthetic_batch("sdxl", B, dev).items()}
    dev_batch = trainer.batch_to_device(host_batch, dev)
    bucket = dp.GradientBucket(None)
    h2d = sum(sum(t.numel() * t.element_size() for t in v) if isinstance(v, list) else v.numel() * v.element_size()
              for v in host_batch.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graphed = None
    if not args.no_graph:
        graphed = trainer.GraphedTrainStep(setup, model, cfg, host_batch)
        launches_per_step = None

    def step_resident():
        if graphed is not None:
            return graphed.step()                      # batch already staged; device-side work only
        return trainer.train_step(setup, model, dev_batch, cfg, bucket)

    def step_e2e():
        if graphed is not None:
            return graphed.step(host_batch).item()     # host batch -> pinned staging -> H2D inside the graph -> D2H loss
        b = trainer.batch_to_device(host_batch, dev, non_blocking=True)
        return trainer.train_step(setup, model, b, cfg, bucket).item()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.LAUNCHES
        e0.record()
        for _ in range(k):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / 1e3, ops.LAUNCHES - l0, out

    l_before = ops.LAUNCHES
    if graphed is None:
        trainer.train_step(setup, model, dev_batch, cfg, bucket)
    launches_eager_step = ops.LAUNCHES - l_before     # kernels of ours per step (a graph replays exactly these)
    for _ in range(max(args.warmup, 3)):
        step_resident()
    with ClockSampler(local_rank) as clk:
        sec, launches, last_loss = timed(step_resident, args.steps)
    if graphed is not None:
        launches = graphed.launches_per_step * args.steps
    clocks = clk.summary()
    for _ in range(2):
        step_e2e()
    sec_e2e, _, loss_e2e = timed(step_e2e, args.steps)

    peaks, peak_kind = measured_peaks()
    ka = time_leaf_kernels(model, dev)
    achieved = ka["gemm_flops"] / ka["gemm_sec"] / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    hbm_peak = float(peaks["hbm_gbs"])
    hbm_kernels = time_hbm_kernels(dev, hbm_peak) if rank == 0 else []
    grad_gbs = ka["grad_bytes"] / ka["grad_sec"] / 1e9
    hbm_kernels.insert(0, {"kernel": "lora_grad_kernel (K-a', dA/dB; all 722 leaves of one step)", "bound": "hbm", "achieved": grad_gbs,
                           "peak": hbm_peak, "unit": "GB/s", "frac": grad_gbs / hbm_peak, "seconds_per_step_in_kernel": ka["grad_sec"],
                           "launches_per_step": ka["grad_launches"]})
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline
        s = cpu_baseline.run_sample()
        cpu = {"value": 1.0 / s["step_seconds"], "unit": UNIT, "cores": s["threads"], "kind": "port", "sample": s["sample"],
               "sample_seconds": s["sample_seconds"], "parts_seconds": s["parts"]}
    traffic, traffic_of = profiled_traffic()
    if rank == 0:
        value = world * args.steps / sec
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": B, "global_batch": B * world, "lora_rank": 32, "hooked_leaves": 722,
                       "lora_params": sum(p.numel() for p in model.parameters), "parallelism": f"dp{world}",
                       "l2": "inputs larger than L2 (4.4 GB of frozen weights + 0.68 GB LoRA/optimizer state stream every step)"},
            "samples_per_sec": value * B, "optimizer_steps_per_sec": args.steps / sec,
            "e2e": {"value": world * args.steps / sec_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": sec_e2e / args.steps * 1e3},
            "gpu_launches": launches, "cuda_graph": graphed is not None, "loss": float(last_loss), "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "lora_linear_kernel_v3 / _v2 (K-a: the forward + dX launches of all 722 leaves; q/k/v siblings grouped)",
                         "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_of": traffic_of,
                         "peak_source": f"{peak_kind} bf16_tflops_sustained",
                         "algorithmic_flops_per_step": ka["gemm_flops"], "launches_per_step": ka["gemm_launches"],
                         "seconds_per_step_in_kernel": ka["gemm_sec"], "share_of_step": ka["gemm_sec"] / (sec / args.steps)},
            "hbm_kernels": hbm_kernels, "hbm_peak_source": f"{peak_kind} hbm_gbs",
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        # Leave without tearing the communicator down: with the NCCL allreduce captured inside a live CUDA graph,
        # destroy_process_group() blocked for minutes on the B200 box (gpurun_out/bench_n2.err, round 1).  Every
        # collective this process will ever issue has completed (the timed regions end on an allreduce + sync),
        # so the ranks can exit independently.
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        own_arm(a)
