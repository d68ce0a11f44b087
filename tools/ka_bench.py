"""Device-time microbenchmark of K-a per SDXL leaf shape (CUDA-graph replay, so no host launch gaps).
usage: python tools/ka_bench.py [flux] [tile_n] [stages] [ref] ; env OTB_KA_VER=2|3 selects the variant."""
import sys
import torch
sys.path.insert(0, ".")
from onetrainer_b200 import ops

dev = torch.device("cuda:0")
import os
STATIC = bool(int(os.environ.get("KA_STATIC", "0")))      # tell K-a the weights are not written by the preceding launch
FLUX = "flux" in sys.argv[1:]
sys.argv = [a for a in sys.argv if a not in ("flux",)]
tile_n = int(sys.argv[1]) if len(sys.argv) > 1 else 0
stages = int(sys.argv[2]) if len(sys.argv) > 2 else 0
# (name, M, N, K, count per step) for SDXL B=4: forward shapes; dX swaps N and K
shapes = [("c1280 qkvo/proj", 4096, 1280, 1280, 6 * 60 + 12), ("c1280 kv(ctx)", 308, 1280, 2048, 2 * 60), ("c1280 ff0", 4096, 10240, 1280, 60),
          ("c1280 ff2", 4096, 1280, 5120, 60), ("c640 qkvo/proj", 16384, 640, 640, 6 * 10 + 10), ("c640 kv(ctx)", 308, 640, 2048, 2 * 10),
          ("c640 ff0", 16384, 5120, 640, 10), ("c640 ff2", 16384, 640, 2560, 10)]
R = 32
if FLUX:      # FLUX.1-dev at B=1 (config C5): 19 double + 38 single blocks, rank 16
    R = 16
    shapes = [("flux img qkvo", 4096, 3072, 3072, 4 * 19), ("flux txt qkvo", 512, 3072, 3072, 4 * 19), ("flux ff0", 4096, 12288, 3072, 19),
              ("flux ff2", 4096, 3072, 12288, 19), ("flux single qkv", 4608, 3072, 3072, 3 * 38)]
    tile_n = int(sys.argv[1]) if len(sys.argv) > 1 else 0
tot_t, tot_f = 0.0, 0.0
for name, M, N, K, cnt in shapes:
    for role, (m, n, k) in (("fwd", (M, N, K)), ("dX ", (M, K, N))):
        nrep = 8
        xs = [torch.randn(m, k, device=dev).bfloat16() for _ in range(2)]
        ws = [(torch.randn(n, k, device=dev) / k ** 0.5).bfloat16() for _ in range(nrep)]      # distinct weights: stream from HBM
        a = (torch.randn(R, k, device=dev) / k ** 0.5).bfloat16(); b = torch.randn(n, R, device=dev).bfloat16()
        bias = torch.randn(n, device=dev).bfloat16()

        def run():
            for i in range(nrep):
                ops.lora_linear(xs[i & 1], ws[i], a, b, bias, 1 / 32, tile_n=tile_n, stages=stages, static_weights=STATIC)
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
        us = e0.elapsed_time(e1) * 1e3 / 5 / nrep
        fl = 2.0 * m * n * k + 2.0 * m * R * (k + n)
        ref_us = 0.0
        if len(sys.argv) > 3 and sys.argv[3] == "ref":
            # the reference's op sequence (LoRAModule.py:322-323) through torch/cuBLAS on the same operands
            import torch.nn.functional as F

            def run_ref():
                for i in range(nrep):
                    F.linear(xs[i & 1], ws[i], bias) + F.linear(F.linear(xs[i & 1], a), b) * (1 / 32)
            with torch.cuda.stream(s):
                run_ref()
            torch.cuda.synchronize()
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2):
                run_ref()
            g2.replay(); torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                g2.replay()
            e1.record(); torch.cuda.synchronize()
            ref_us = e0.elapsed_time(e1) * 1e3 / 5 / nrep
            del g2
        tot_t += us * cnt; tot_f += fl * cnt
        print(f"{name:16s} {role} M={m:6d} N={n:6d} K={k:5d}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  x{cnt}"
              + (f"   torch 3-GEMM path {ref_us:8.1f} us ({ref_us / us:.2f}x)" if ref_us else ""), flush=True)
        tot_ref = globals().get("tot_ref", 0.0) + ref_us * cnt
        globals()["tot_ref"] = tot_ref
        del g
print(f"weighted over one step (fwd+dX): {tot_t / 1e3:.2f} ms, {tot_f / tot_t / 1e6:.1f} TFLOP/s"
      + (f"; torch 3-GEMM path {globals().get('tot_ref', 0.0) / 1e3:.2f} ms" if globals().get("tot_ref") else ""))
