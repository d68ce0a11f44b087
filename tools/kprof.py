"""Launch one kernel a few times at its roofline size — the target of `ncu --set full`.
usage: python tools/kprof.py {kc_bf16|kc_fp32|kb1|kb2|ka|kgrad} [iters]"""
import sys
import torch
sys.path.insert(0, ".")
from onetrainer_b200 import _native as N, ops

dev = torch.device("cuda:0")
which = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
T = 1000
betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, T, dtype=torch.float32) ** 2
ac = torch.cumprod(1 - betas, 0)
sa, s1 = ac.sqrt().to(dev), (1 - ac).sqrt().to(dev)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if which in ("kc_bf16", "kc_fp32"):
    dt = torch.bfloat16 if which == "kc_bf16" else torch.float32
    n = 1 << 28 if dt == torch.bfloat16 else 1 << 27
    p = torch.zeros(n, device=dev, dtype=dt); g = torch.full_like(p, 1e-3); m = torch.zeros_like(p); v = torch.ones_like(p); e = torch.zeros_like(p)
    tab = ops.OptimTable([dict(p=p, g=g, exp_avg=m, exp_avg_sq=v, ema=e)], dev)
    if dt == torch.bfloat16:
        tab.host["flags"][:] |= N.OPT_FLAG_SR
    for k, val in ops.adam_scalars(3e-4, 0.9, 0.999, 1e-8, 1e-2, 1.0, True).items():
        tab.host[k][:] = val
    tab.upload()
    ms = timed(lambda: ops.adamw_sr_ema_multi(tab, do_ema=True, ema_one_minus_decay=0.001, seed=1, offset=2))
    bpe = 18 if dt == torch.bfloat16 else 36
    print(f"{which}: {ms:.3f} ms, {n * bpe / ms / 1e6:.0f} GB/s")
elif which == "kb1":
    Bb = 4096
    lat = torch.randn(Bb, 4, 128, 128, device=dev)
    ms = timed(lambda: ops.noise_addnoise(latent=lat, coef_x0=sa, coef_noise=s1, B=Bb, C_=4, hw=128 * 128, latent_dtype=torch.float32,
                                          device=dev, cast_dtype=torch.bfloat16, target=N.TARGET_EPSILON, ts_mode=N.TS_UNIFORM, seed=1, shape=lat.shape))
    print(f"kb1: {ms:.3f} ms, {lat.numel() * 10 / ms / 1e6:.0f} GB/s")
elif which == "kb2":
    Bb = 4096
    pred = torch.randn(Bb, 4, 128, 128, device=dev, dtype=torch.bfloat16)
    tgt = torch.randn(Bb, 4, 128, 128, device=dev)
    ts = torch.randint(0, 1000, (Bb,), device=dev, dtype=torch.int32)
    ms = timed(lambda: ops.loss_fwd_bwd(pred, tgt, timestep=ts, sqrt_ac=sa, sqrt_1mac=s1, loss_weight_fn=N.LW_MIN_SNR, upstream=1 / Bb))
    print(f"kb2: {ms:.3f} ms, {pred.numel() * 8 / ms / 1e6:.0f} GB/s")
elif which == "ka":
    M, Nn, K, R = 4096, 10240, 1280, 32
    x = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(Nn, K, device=dev) / 36).bfloat16()
    down = (torch.randn(R, K, device=dev) / 36).bfloat16(); up = torch.randn(Nn, R, device=dev).bfloat16()
    bias = torch.randn(Nn, device=dev).bfloat16()
    ms = timed(lambda: ops.lora_linear(x, w, down, up, bias, 1 / 32))
    print(f"ka: {ms * 1e3:.1f} us, {(2.0 * M * Nn * K + 2.0 * M * R * (K + Nn)) / ms / 1e9:.1f} TFLOP/s")
elif which == "kgrad":
    M, Nn, K, R = 4096, 1280, 1280, 32
    x = torch.randn(M, K, device=dev).bfloat16(); dy = torch.randn(M, Nn, device=dev).bfloat16()
    gs = torch.randn(M, R, device=dev).bfloat16(); ts_ = torch.randn(M, R, device=dev).bfloat16()
    dA = torch.zeros(R, K, device=dev); dB = torch.zeros(Nn, R, device=dev)
    ms = timed(lambda: ops.lora_grad_ab(x, dy, gs, ts_, dA, dB))
    print(f"kgrad: {ms * 1e3:.1f} us")
