"""Bring-up experiments on K-a v3: which side of the pipeline bounds a k-block (OTB_KA_DEBUG bits)."""
import os
import sys
import torch
sys.path.insert(0, ".")
from onetrainer_b200 import ops
dev = torch.device("cuda:0")
M, N, K = 4096, 10240, 1280


def bench(R, tile_n, stages):
    nrep = 8
    xs = [torch.randn(M, K, device=dev).bfloat16() for _ in range(2)]
    ws = [(torch.randn(N, K, device=dev) / K ** 0.5).bfloat16() for _ in range(nrep)]
    a = (torch.randn(R, K, device=dev) / K ** 0.5).bfloat16() if R else None
    b = torch.randn(N, R, device=dev).bfloat16() if R else None

    def run():
        for i in range(nrep):
            ops.lora_linear(xs[i & 1], ws[i], a, b, None, 1 / 32, tile_n=tile_n, stages=stages)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / 3 / nrep


for ver, dbg, R, tn, st in [(3, 0, 0, 256, 0), (3, 1, 0, 256, 0), (3, 5, 0, 256, 0), (3, 9, 0, 256, 0), (3, 4, 0, 256, 0), (3, 8, 0, 256, 0), (3, 6, 0, 256, 0),
                            (3, 7, 0, 256, 0), (2, 0, 0, 224, 0)]:
    os.environ["OTB_KA_VER"] = str(ver); os.environ["OTB_KA_DEBUG"] = str(dbg)
    us = bench(R, tn, st)
    print(f"v{ver} debug={dbg} R={R} tile_n={tn} stages={st}: {us:.1f} us  {2.0 * M * N * K / us / 1e6:.0f} TFLOP/s-equivalent", flush=True)
