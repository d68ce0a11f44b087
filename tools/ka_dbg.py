"""Bring-up experiments on K-a v3: which part of a launch costs what (OTB_KA_DEBUG bits: 1 no TMA, 2 no MMA, 4 no epilogue).
Graph-timed (8 launches per replay).  usage: python tools/ka_dbg.py M N K"""
import os
import sys
import torch
sys.path.insert(0, ".")
from onetrainer_b200 import ops
dev = torch.device("cuda:0")
M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (4096, 1280, 1280)


def bench(R, tile_n, stages):
    nrep = 8
    xs = [torch.randn(M, K, device=dev).bfloat16() for _ in range(2)]
    ws = [(torch.randn(N, K, device=dev) / K ** 0.5).bfloat16() for _ in range(nrep)]
    a = (torch.randn(R, K, device=dev) / K ** 0.5).bfloat16() if R else None
    b = torch.randn(N, R, device=dev).bfloat16() if R else None
    bias = torch.randn(N, device=dev).bfloat16()

    def run():
        for i in range(nrep):
            ops.lora_linear(xs[i & 1], ws[i], a, b, bias, 1 / 32, tile_n=tile_n, stages=stages)
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
    return e0.elapsed_time(e1) * 1e3 / 5 / nrep


for R, tn in [(32, 0), (32, 224), (0, 0)]:
    for dbg in (0, 1, 2, 4, 3, 5, 6, 7):
        os.environ["OTB_KA_VER"] = "3"; os.environ["OTB_KA_DEBUG"] = str(dbg)
        us = bench(R, tn, 0)
        print(f"M={M} N={N} K={K} R={R} tile_n={tn} debug={dbg} ({'-TMA ' if dbg & 1 else ''}{'-MMA ' if dbg & 2 else ''}{'-EPI' if dbg & 4 else ''}): {us:.1f} us", flush=True)
