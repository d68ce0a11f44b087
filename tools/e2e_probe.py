"""End-to-end probe on the GPU: builds a stand-in model, runs a few eager training steps on the fused
path and prints loss + timing split.  usage: python tools/e2e_probe.py {tiny|sd15|sdxl|flux|tiny_flux} [batch] [steps] [graph] [full]"""
import sys
import time
import torch
sys.path.insert(0, ".")
from onetrainer_b200.harness import trainer
from onetrainer_b200.modules.util.config.TrainConfig import TrainConfig
from onetrainer_b200.modules.util.enum import LossWeight, TimestepDistribution

kind = sys.argv[1] if len(sys.argv) > 1 else "tiny"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda:0")
flux = kind in ("flux", "tiny_flux")
use_graph = "graph" in sys.argv[4:]
full_ft = "full" in sys.argv[4:]          # config C4: every tensor trainable, SR-AdamW + EMA
if flux:   # training_presets/#flux LoRA.json: LOGIT_NORMAL timesteps, rank 16
    cfg = TrainConfig(train_device=str(dev), batch_size=B, lora_rank=16, timestep_distribution=TimestepDistribution.LOGIT_NORMAL)
else:
    cfg = TrainConfig(train_device=str(dev), batch_size=B, lora_rank=32 if kind == "sdxl" else 16, loss_weight_fn=LossWeight.MIN_SNR_GAMMA,
                      masked_training=(kind == "sd15"))
t0 = time.time()
if flux:
    model, setup = trainer.build_flux_model(kind, cfg, dev, lora_up_std=0.02)
else:
    model, setup = trainer.build_model(kind, cfg, dev, ema=True, lora_up_std=0.02, full_finetune=full_ft)
torch.cuda.synchronize()
print(f"built {kind}{' (full fine-tune)' if full_ft else ''}: {len(model.parameters)} trainable tensors, {sum(p.numel() for p in model.parameters)/1e6:.2f} M params, "
      f"{torch.cuda.memory_allocated()/2**30:.2f} GiB, {time.time()-t0:.1f}s")
batch = trainer.batch_to_device(trainer.synthetic_batch(kind, B, dev, masked=(kind == "sd15")), dev)
graphed = trainer.GraphedTrainStep(setup, model, cfg, {k: ([t.cpu() for t in v] if isinstance(v, list) else v.cpu()) for k, v in batch.items()}) if use_graph else None
for i in range(steps):
    torch.cuda.synchronize(); t = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = graphed.step() if graphed is not None else trainer.train_step(setup, model, batch, cfg)
    e1.record()
    host = time.time() - t
    torch.cuda.synchronize()
    print(f"step {i}: loss {loss.item():.6f} host {host*1e3:.1f} ms gpu-span {e0.elapsed_time(e1):.1f} ms wall {(time.time()-t)*1e3:.1f} ms", flush=True)
print("max mem GiB", torch.cuda.max_memory_allocated() / 2**30)
p0 = model.parameters[0]
print("param0 grad None after zero_grad:", p0.grad is None, "state keys", list(model.optimizer.state[p0].keys()),
      "step", model.optimizer.state[p0]["step"].item())
