"""Device time of K-a' (dA/dB) per SDXL leaf shape: CUDA-graph replay of 8 launches on distinct operands."""
import sys
import torch
sys.path.insert(0, ".")
from onetrainer_b200 import ops
dev = torch.device("cuda:0")
R = 32
shapes = [("c1280 qkvo", 4096, 1280, 1280, 372), ("c1280 kv", 308, 1280, 2048, 120), ("c1280 ff0", 4096, 10240, 1280, 60), ("c1280 ff2", 4096, 1280, 5120, 60),
          ("c640 qkvo", 16384, 640, 640, 70), ("c640 kv", 308, 640, 2048, 20), ("c640 ff0", 16384, 5120, 640, 10), ("c640 ff2", 16384, 640, 2560, 10)]
tot = 0.0
for name, M, N, K, cnt in shapes:
    nrep = 8
    xs = [torch.randn(M, K, device=dev).bfloat16() for _ in range(nrep)]
    dys = [torch.randn(M, N, device=dev).bfloat16() for _ in range(nrep)]
    gs = torch.randn(M, R, device=dev).bfloat16(); ts = torch.randn(M, R, device=dev).bfloat16()
    dA = torch.zeros(R, K, device=dev); dB = torch.zeros(N, R, device=dev)

    def run():
        for i in range(nrep):
            ops.lora_grad_ab(xs[i], dys[i], gs, ts, dA, dB)
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        run()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 5 / nrep
    tot += us * cnt
    print(f"{name:11s} M={M:6d} N={N:6d} K={K:5d}: {us:7.1f} us  {(M * (N + K) * 2) / us / 1e3:7.0f} GB/s streamed  x{cnt}", flush=True)
print(f"weighted over one step: {tot / 1e3:.2f} ms")
