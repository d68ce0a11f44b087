"""Config C4 (SDXL full fine-tune, bf16 weights / AdamW state / EMA, stochastic rounding) under data parallelism:
one process per GPU (torchrun), gradients in a dp.FlatGradArena allreduced in place, whole step as one CUDA graph.
usage: torchrun --nproc-per-node N tools/c4_probe.py [steps]   (or plain python for N = 1)"""
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from onetrainer_b200 import dp
from onetrainer_b200.harness import trainer
from onetrainer_b200.modules.util.config.TrainConfig import TrainConfig
from onetrainer_b200.modules.util.enum import LossWeight

rank, local_rank, world = dp.init_from_env()
dev = torch.device("cuda", local_rank)
torch.cuda.set_device(dev)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = 4
cfg = TrainConfig(train_device=str(dev), batch_size=B, loss_weight_fn=LossWeight.MIN_SNR_GAMMA, clip_grad_norm=1.0)
model, setup = trainer.build_model("sdxl", cfg, dev, ema=True, stochastic_rounding=True, full_finetune=True, lr=1e-6)
host_batch = trainer.synthetic_batch("sdxl", B, dev)
graphed = trainer.GraphedTrainStep(setup, model, cfg, host_batch)
for _ in range(2):
    graphed.step()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = graphed.step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    n = sum(p.numel() for p in model.parameters)
    print(f"C4 full fine-tune, {n / 1e9:.2f} B bf16 parameters, {world} GPU(s), batch {B}/GPU: {ms.item():.1f} ms/step, "
          f"{world * 1e3 / ms.item():.2f} steps/s, loss {loss.item():.4f}, gradient arena {model.full_grad_arena.flat.numel() * 2 / 2**30:.2f} GiB, "
          f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
sys.stdout.flush()
import os
os._exit(0)
