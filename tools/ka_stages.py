"""K-a v3 ring-depth sensitivity: time vs number of smem stages (graph-timed)."""
import os, sys
import torch
sys.path.insert(0, ".")
from onetrainer_b200 import ops
dev = torch.device("cuda:0")
os.environ["OTB_KA_VER"] = "3"
for (M, N, K, tn) in [(4096, 1280, 5120, 224), (4096, 1280, 5120, 128), (4096, 1280, 5120, 320), (4096, 10240, 1280, 224)]:
    for st in (2, 3, 4, 5, 6):
        nrep = 4
        xs = [torch.randn(M, K, device=dev).bfloat16() for _ in range(2)]
        ws = [(torch.randn(N, K, device=dev) / K ** 0.5).bfloat16() for _ in range(nrep)]
        a = (torch.randn(32, K, device=dev) / K ** 0.5).bfloat16(); b = torch.randn(N, 32, device=dev).bfloat16()
        def run():
            for i in range(nrep):
                ops.lora_linear(xs[i & 1], ws[i], a, b, None, 1 / 32, tile_n=tn, stages=st)
        run(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            run()
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        with torch.cuda.graph(g):
            run()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 5 / nrep
        print(f"M={M} N={N} K={K} tile_n={tn} stages(requested)={st}: {us:.1f} us  {2.0*M*N*K/us/1e6:.0f} TFLOP/s", flush=True)
