"""One eager training step between cudaProfilerStart/Stop — the target of the ncu launch list:
ncu --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file out.csv python tools/step_profile.py"""
import sys
import torch
sys.path.insert(0, ".")
from onetrainer_b200.harness import trainer
from onetrainer_b200.modules.util.config.TrainConfig import TrainConfig
from onetrainer_b200.modules.util.enum import LossWeight

dev = torch.device("cuda:0")
kind = sys.argv[1] if len(sys.argv) > 1 else "sdxl"
B = 4
cfg = TrainConfig(train_device=str(dev), batch_size=B, lora_rank=32 if kind == "sdxl" else 16, loss_weight_fn=LossWeight.MIN_SNR_GAMMA)
model, setup = trainer.build_model(kind, cfg, dev, lora_up_std=0.02)
batch = trainer.batch_to_device(trainer.synthetic_batch(kind, B, dev), dev)
for _ in range(2):
    trainer.train_step(setup, model, batch, cfg)
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = trainer.train_step(setup, model, batch, cfg)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", loss.item())
