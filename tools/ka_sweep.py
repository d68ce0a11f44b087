"""Tile-width / variant sweep of K-a on a few SDXL leaf shapes (device time via CUDA-graph replay).
usage: python tools/ka_sweep.py  -> one line per (shape, R, version, tile_n)"""
import os
import sys
import torch
sys.path.insert(0, ".")
from onetrainer_b200 import ops

dev = torch.device("cuda:0")
shapes = [("ctx kv", 308, 1280, 2048), ("ctx kv dX", 308, 2048, 1280), ("c640 ctx kv", 308, 640, 2048), ("c640 ctx dX", 308, 2048, 640),
          ("c1280 qkvo", 4096, 1280, 1280), ("c1280 ff0", 4096, 10240, 1280), ("c1280 ff2", 4096, 1280, 5120), ("c640 qkvo", 16384, 640, 640)]
cfgs = [(2, 0), (2, 128), (2, 224), (3, 0), (3, 96), (3, 128), (3, 160), (3, 192), (3, 224), (3, 256), (3, 320), (3, 448)]
if len(sys.argv) > 1 and sys.argv[1] == "ctx":
    cfgs = [(2, 0), (2, 16), (2, 32), (2, 48), (2, 64), (2, 96), (3, 0), (3, 32), (3, 48), (3, 64), (3, 96), (3, 128), (3, 160)]
    shapes = [s for s in shapes if "ctx" in s[0]]
elif len(sys.argv) > 1:
    shapes = [s for s in shapes if s[0] in sys.argv[1:]]


def bench(m, n, k, R, tile_n):
    nrep = 8
    xs = [torch.randn(m, k, device=dev).bfloat16() for _ in range(2)]
    ws = [(torch.randn(n, k, device=dev) / k ** 0.5).bfloat16() for _ in range(nrep)]
    a = (torch.randn(R, k, device=dev) / k ** 0.5).bfloat16() if R else None
    b = torch.randn(n, R, device=dev).bfloat16() if R else None
    bias = torch.randn(n, device=dev).bfloat16()

    def run():
        for i in range(nrep):
            ops.lora_linear(xs[i & 1], ws[i], a, b, bias, 1 / 32, tile_n=tile_n)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        run()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
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


for name, M, N, K in shapes:
    for R in ((32,) if "ctx" in name else (32, 0)):
        for ver, tn in cfgs:
            if ver == 2 and tn + R > 256:
                continue
            if tn + R > 512 or (R == 0 and tn == 256 and ver == 2 and False):
                continue
            os.environ["OTB_KA_VER"] = str(ver)
            try:
                us = bench(M, N, K, R, tn)
            except Exception as e:  # noqa: BLE001
                print(f"{name} R={R} v{ver} tile_n={tn}: FAILED {str(e)[:100]}", flush=True)
                continue
            fl = 2.0 * M * N * K + 2.0 * M * R * (K + N)
            print(f"{name:11s} M={M} N={N} K={K} R={R:2d} v{ver} tile_n={tn:3d}: {us:7.1f} us {fl / us / 1e6:7.1f} TFLOP/s", flush=True)
