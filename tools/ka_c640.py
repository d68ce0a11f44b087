"""K-a on the short-K SDXL shapes (C=640 level: K = 640, 10 k-blocks per tile): tile-width sweep and the
TMA / MMA / epilogue decomposition (OTB_KA_DEBUG) per shape.  Graph-timed.  usage: python tools/ka_c640.py"""
import os
import sys
import torch
sys.path.insert(0, ".")
from onetrainer_b200 import ops
dev = torch.device("cuda:0")


def bench(M, N, K, R, tile_n, stages=0):
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


os.environ["OTB_KA_VER"] = "3"
if len(sys.argv) > 1 and sys.argv[1] == "tiles":
    # tile-width sweep of every SDXL leaf shape class (after the epilogue fixes): does the planner still pick the best width?
    shapes = [("c1280 qkvo", 4096, 1280, 1280), ("c1280 ff0 fwd", 4096, 10240, 1280), ("c1280 ff0 dX", 4096, 1280, 10240),
              ("c1280 ff2 fwd", 4096, 1280, 5120), ("c1280 ff2 dX", 4096, 5120, 1280), ("c640 qkvo", 16384, 640, 640),
              ("c640 ff0 fwd", 16384, 5120, 640), ("c640 ff0 dX", 16384, 640, 5120), ("c640 ff2 fwd", 16384, 640, 2560),
              ("c640 ff2 dX", 16384, 2560, 640)]
    for name, M, N, K in shapes:
        fl = 2.0 * M * N * K + 2.0 * M * 32 * (K + N)
        res = []
        for tn in (0, 128, 160, 192, 208, 224, 288, 320, 352, 384, 416, 448, 480):
            try:
                res.append((tn, bench(M, N, K, 32, tn)))
            except Exception as e:  # noqa: BLE001
                pass
        best = min(res[1:], key=lambda r: r[1])
        print(f"{name} M={M} N={N} K={K}: auto {res[0][1]:.1f} us ({fl / res[0][1] / 1e6:.0f} TFLOP/s); best tile_n={best[0]} {best[1]:.1f} us; "
              + " ".join(f"{tn}:{us:.1f}" for tn, us in res[1:]), flush=True)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "epi":
    # where the output epilogue's time goes: TMA loads and MMAs off (debug 3), then pieces of the epilogue switched off
    for name, M, N, K, tn in [("c640 ff0 fwd", 16384, 5120, 640, 224), ("c640 ff0 fwd", 16384, 5120, 640, 192), ("c1280 ff0 fwd", 4096, 10240, 1280, 224)]:
        for dbg, what in [(0, "complete"), (3, "epilogue only"), (3 | 8, "epilogue, no TMA store"), (3 | 8 | 32, "epilogue, no TMA store, no proxy fence"),
                          (3 | 16, "epilogue, no accumulator loads"), (3 | 8 | 16 | 32, "epilogue: bias + pack + st.shared only"), (7, "all off"),
                          (8, "complete but no TMA store"), (16, "complete but no accumulator loads")]:
            os.environ["OTB_KA_DEBUG"] = str(dbg)
            for bias in (True, False):
                us = bench(M, N, K, 32, tn) if bias else None
                if us is not None:
                    print(f"{name} tile_n={tn} debug={dbg} ({what}): {us:.1f} us", flush=True)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "lookahead":
    # rank-R hand-off placement: k-blocks of the next tile issued before the up-projection MMAs of the finished one
    shapes = [("c640 qkvo", 16384, 640, 640, 224), ("c640 ff0 fwd", 16384, 5120, 640, 224), ("c640 ff2 dX", 16384, 2560, 640, 224),
              ("c1280 ff2 dX", 4096, 5120, 1280, 224), ("c1280 ff0 fwd", 4096, 10240, 1280, 224), ("c1280 ff0 fwd", 4096, 10240, 1280, 0),
              ("c1280 qkv grouped-like", 12288, 1280, 1280, 160)]
    for name, M, N, K, tn in shapes:
        fl = 2.0 * M * N * K + 2.0 * M * 32 * (K + N)
        for la in (0, 1, 2, 3, 4, 6, 8):
            os.environ["OTB_KA_LOOKAHEAD"] = str(la)
            us = bench(M, N, K, 32, tn)
            print(f"{name} M={M} N={N} K={K} tile_n={tn} lookahead={la}: {us:.1f} us  {fl / us / 1e6:.0f} TFLOP/s", flush=True)
    sys.exit(0)

shapes = [("c640 qkvo", 16384, 640, 640), ("c640 ff0 fwd", 16384, 5120, 640), ("c640 ff2 dX", 16384, 2560, 640), ("c1280 ff2 dX", 4096, 5120, 1280)]
for name, M, N, K in shapes:
    fl = 2.0 * M * N * K + 2.0 * M * 32 * (K + N)
    for R in (32, 0):
        for tn in (0, 96, 128, 160, 192, 224, 288, 320, 384, 448, 480):
            if R == 0 and tn not in (0, 128, 256 - 32, 256):
                continue
            try:
                us = bench(M, N, K, R, tn)
                print(f"{name} M={M} N={N} K={K} R={R} tile_n={tn}: {us:.1f} us  {fl / us / 1e6:.0f} TFLOP/s", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"{name} R={R} tile_n={tn}: {str(e)[:80]}", flush=True)
    for dbg in (0, 1, 2, 4, 3, 5, 6, 7):
        os.environ["OTB_KA_DEBUG"] = str(dbg)
        us = bench(M, N, K, 32, 224)
        print(f"{name} tile_n=224 debug={dbg} ({'-TMA ' if dbg & 1 else ''}{'-MMA ' if dbg & 2 else ''}{'-EPI' if dbg & 4 else ''}): {us:.1f} us", flush=True)
    os.environ["OTB_KA_DEBUG"] = "0"
