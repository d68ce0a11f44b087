"""K-r (otb_grad_exchange, both modes) against ncclAllReduce on the LoRA arena size: correctness, then CUDA-event timing.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 tools/exchange_check.py [n_elems]"""
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from onetrainer_b200 import dp  # noqa: E402

rank, local, world = dp.init_from_env("nccl")
dev = torch.device("cuda", local)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 85114880          # SDXL rank-32 LoRA: 85.1 M fp32 gradients
g = torch.Generator(device=dev).manual_seed(100 + rank)
src = torch.randn(n, device=dev, generator=g) * 1e-3
ref = src.clone()
dist.all_reduce(ref)
ok = True
exchanges = {}
for mode in ("p2p", "nvls"):
    arena = dp.symmetric_zeros(n, torch.float32, dev)
    arena.copy_(src)
    ex = exchanges[mode] = dp.GradExchange(arena, mode=mode)
    partials = ex.exchange(1.0 / world)
    torch.cuda.synchronize()
    err = (arena - ref).abs().max().item()
    norm = partials.double().sum().sqrt().item()
    ref_norm = (ref.double() / world).norm().item()
    r0 = arena.clone()
    dist.broadcast(r0, 0)
    print(f"rank {rank}/{world} {mode}: n={n} blocks={ex.blocks} equal_to_nccl={torch.equal(arena, ref)} max_err={err:.3e} "
          f"replicas_identical={torch.equal(r0, arena)} norm={norm:.9g} ref_norm={ref_norm:.9g}", flush=True)
    ok = ok and err <= 1e-6 * ref.abs().max().item() + 1e-12 and torch.equal(r0, arena) and abs(norm - ref_norm) <= 1e-5 * ref_norm


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


plain = torch.zeros(n, device=dev)
t_nccl = timed(lambda: dist.all_reduce(plain))
t = {mode: timed(lambda ex=ex: ex.exchange(1.0 / world)) for mode, ex in exchanges.items()}
if rank == 0:
    mb = n * 4 / 1e6
    print(f"{world} GPUs, {mb:.0f} MB fp32: ncclAllReduce {t_nccl * 1e3:.0f} us | K-r (exchange + squared norm) "
          + " | ".join(f"{mode} {v * 1e3:.0f} us = {2 * (world - 1) / world * mb / 1e3 / (v / 1e3):.0f} GB/s bus" for mode, v in t.items()), flush=True)
    print("exchange_check:", "ok" if ok else "MISMATCH", flush=True)
torch.cuda.synchronize()
os._exit(0 if ok else 1)
