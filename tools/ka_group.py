"""Grouped K-a launches (otb_lora_linear_grouped) vs the same problems launched one by one: bit-equality and device time
(CUDA-graph replay).  usage: python tools/ka_group.py"""
import sys
import torch
sys.path.insert(0, ".")
from onetrainer_b200 import ops
dev = torch.device("cuda:0")
R = 32


def gtime(fn, nrep=4):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(nrep):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / 5 / nrep


for name, M, N, K, g, shared_x in [("self-attn qkv fwd", 4096, 1280, 1280, 3, True), ("self-attn qkv dX", 4096, 1280, 1280, 3, False),
                                   ("cross-attn kv fwd", 308, 1280, 2048, 2, True), ("cross-attn kv dX", 308, 2048, 1280, 2, False),
                                   ("c640 qkv fwd", 16384, 640, 640, 3, True), ("c640 qkv dX", 16384, 640, 640, 3, False),
                                   ("c640 kv fwd", 308, 640, 2048, 2, True), ("flux qkv fwd", 4096, 3072, 3072, 3, True)]:
    xs = [torch.randn(M, K, device=dev).bfloat16()] * g if shared_x else [torch.randn(M, K, device=dev).bfloat16() for _ in range(g)]
    ws = [(torch.randn(N, K, device=dev) / K ** 0.5).bfloat16() for _ in range(g)]
    downs = [(torch.randn(R, K, device=dev) / K ** 0.5).bfloat16() for _ in range(g)]
    ups = [torch.randn(N, R, device=dev).bfloat16() for _ in range(g)]
    biases = [torch.randn(N, device=dev).bfloat16() if shared_x else None for _ in range(g)]
    ys, ts = ops.lora_linear_grouped(xs, ws, downs, ups, biases, [1 / 32] * g)
    ok = True
    for i in range(g):
        y1, t1 = ops.lora_linear(xs[i], ws[i], downs[i], ups[i], biases[i], 1 / 32)
        ok &= torch.equal(y1, ys[i]) and torch.equal(t1, ts[i])
    tg = gtime(lambda: ops.lora_linear_grouped(xs, ws, downs, ups, biases, [1 / 32] * g))
    t1 = gtime(lambda: [ops.lora_linear(xs[i], ws[i], downs[i], ups[i], biases[i], 1 / 32) for i in range(g)])
    fl = g * (2.0 * M * N * K + 2.0 * M * R * (K + N))
    print(f"{name:18s} M={M:6d} N={N:5d} K={K:5d} x{g}: grouped {tg:7.1f} us ({fl / tg / 1e6:6.1f} TFLOP/s)  singles {t1:7.1f} us  "
          f"{'bit-identical' if ok else 'MISMATCH'}", flush=True)
