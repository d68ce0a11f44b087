"""Bring-up check of a K-a variant against an fp32 torch restatement on assorted shapes.
usage: OTB_KA_VER=3 python tools/ka_check.py [quick]
Prints max abs / rel error per shape for y and t_out and exits non-zero on a mismatch."""
import os
import sys
import torch
sys.path.insert(0, ".")
from onetrainer_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
shapes = [  # M, N, K, R, tile_n, bias
    (256, 64, 64, 16, 0, True), (256, 256, 128, 32, 0, True), (128, 64, 64, 16, 0, False), (300, 328, 136, 8, 0, True),
    (4096, 1280, 1280, 32, 0, True), (4096, 1280, 1280, 32, 224, True), (4096, 1280, 1280, 32, 320, True), (4096, 1280, 1280, 32, 160, True),
    (4096, 1280, 1280, 32, 192, False), (4000, 1272, 1280, 32, 224, True), (4096, 1280, 1280, 16, 240, True),
    (4096, 1280, 1280, 32, 64, False), (308, 1280, 2048, 32, 0, False), (16384, 640, 640, 32, 0, True),
    (4096, 10240, 1280, 32, 0, True), (4096, 1280, 5120, 32, 0, True), (1000, 1288, 72, 64, 0, True), (4096, 1280, 1280, 0, 0, True),
    (777, 200, 264, 24, 0, True), (4096, 1280, 1280, 32, 448, True),
]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    shapes = shapes[:5]
bad = 0
for (M, N, K, R, tn, has_bias) in shapes:
    x = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    a = (torch.randn(R, K, device=dev) / K ** 0.5).bfloat16() if R else None
    b = torch.randn(N, R, device=dev).bfloat16() if R else None
    bias = torch.randn(N, device=dev).bfloat16() if has_bias else None
    scale = 0.5
    try:
        y, t = ops.lora_linear(x, w, a, b, bias, scale, tile_n=tn)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"M={M} N={N} K={K} R={R} tile_n={tn}: FAILED {e}", flush=True)
        bad += 1
        continue
    ref = x.float() @ w.float().T
    if has_bias:
        ref = ref + bias.float()
    if R:
        t_ref = ((x.float() @ a.float().T) * scale).bfloat16()
        ref = ref + t_ref.float() @ b.float().T
        t_err = (t.float() - t_ref.float()).abs().max().item() / (t_ref.float().abs().max().item() + 1e-9)
    else:
        t_err = 0.0
    err = (y.float() - ref).abs().max().item()
    tol = ref.abs().max().item() * 2 ** -7 + 1e-2
    ok = err <= tol and t_err <= 2 ** -6
    bad += 0 if ok else 1
    print(f"M={M} N={N} K={K} R={R} tile_n={tn}: y max abs err {err:.4g} (tol {tol:.3g}), t rel err {t_err:.3g}  {'ok' if ok else 'MISMATCH'}", flush=True)
print("ka_check:", "all ok" if bad == 0 else f"{bad} FAILED")
sys.exit(1 if bad else 0)
