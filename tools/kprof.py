"""Launch one kernel a few times at its roofline size — the target of `ncu --set full`.
usage: python tools/kprof.py {kc_bf16|kc_fp32|kb1|kb2|ka|ka1280|kaqkv|kgrad|kgradb|adafactor|came|dora|sr} [iters]"""
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
    import os
    tab = ops.OptimTable([dict(p=p, g=g, exp_avg=m, exp_avg_sq=v, ema=e)], dev, chunk_elems=int(os.environ.get("OTB_KC_CHUNK", ops.CHUNK_ELEMS)))
    if dt == torch.bfloat16:
        tab.host["flags"][:] |= N.OPT_FLAG_SR
    for k, val in ops.adam_scalars(3e-4, 0.9, 0.999, 1e-8, 1e-2, 1.0, True).items():
        tab.host[k][:] = val
    tab.upload()
    ms = timed(lambda: ops.adamw_sr_ema_multi(tab, do_ema=True, ema_one_minus_decay=0.001, seed=1, offset=2))
    bpe = 18 if dt == torch.bfloat16 else 36
    print(f"{which}: {ms:.3f} ms, {n * bpe / ms / 1e6:.0f} GB/s")
elif which == "sr":
    # the three helpers of bf16_stochastic_rounding.py at the roofline size, ours (one launch) and the reference's op sequence
    # (its own source lines, restated here so the tool needs no baseline/_ref) on the same GPU
    n = 1 << 28
    tgt = (torch.randn(n, device=dev) * 0.02).bfloat16()
    src = torch.randn(n, device=dev) * 1e-4
    t1 = (torch.randn(n, device=dev) * 1e-3).bfloat16(); t2 = (torch.rand(n, device=dev) * 1e-3 + 1e-8).bfloat16()

    def ref_copy(target, source):
        result = torch.randint_like(source, dtype=torch.int32, low=0, high=(1 << 16))
        result.add_(source.view(dtype=torch.int32))
        result.bitwise_and_(-65536)
        target.copy_(result.view(dtype=torch.float32))

    def ref_add(inp, other, alpha):
        result = other.clone() if other.dtype == torch.float32 else other.to(dtype=torch.float32)
        result.add_(inp, alpha=alpha)
        ref_copy(inp, result)

    def ref_addcdiv(inp, a, b, value):
        result = inp.to(dtype=torch.float32)
        result.addcdiv_(a, b, value=value)
        ref_copy(inp, result)

    for name, bpe, mine, ref in (
            ("copy_stochastic_", 6, lambda: ops.stochastic_round(N.SR_COPY, tgt, src, seed=1, offset=2), lambda: ref_copy(tgt, src)),
            ("add_stochastic_", 8, lambda: ops.stochastic_round(N.SR_ADD, tgt, src, alpha=1.0, seed=1, offset=2), lambda: ref_add(tgt, src, 1.0)),
            ("addcdiv_stochastic_", 8, lambda: ops.stochastic_round(N.SR_ADDCDIV, tgt, t1, t2, alpha=-3e-4, seed=1, offset=2),
             lambda: ref_addcdiv(tgt, t1, t2, -3e-4))):
        ms, ms_ref = timed(mine), timed(ref)
        print(f"sr {name}: {ms:.3f} ms, {n * bpe / ms / 1e6:.0f} GB/s of {bpe} B/elem; reference op sequence on the same GPU {ms_ref:.3f} ms ({ms_ref / ms:.1f}x)")
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
elif which == "ka1280":
    # the most frequent SDXL leaf: one round of 64 pair tiles
    M, Nn, K, R = 4096, 1280, 1280, 32
    x = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(Nn, K, device=dev) / 36).bfloat16()
    down = (torch.randn(R, K, device=dev) / 36).bfloat16(); up = torch.randn(Nn, R, device=dev).bfloat16()
    bias = torch.randn(Nn, device=dev).bfloat16()
    ms = timed(lambda: ops.lora_linear(x, w, down, up, bias, 1 / 32))
    print(f"ka1280: {ms * 1e3:.1f} us, {(2.0 * M * Nn * K + 2.0 * M * R * (K + Nn)) / ms / 1e9:.1f} TFLOP/s")
elif which == "kac640":
    # short-K shape of the C=640 level (10 k-blocks per tile): epilogue-bound
    M, Nn, K, R = 16384, 5120, 640, 32
    x = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(Nn, K, device=dev) / 25).bfloat16()
    down = (torch.randn(R, K, device=dev) / 25).bfloat16(); up = torch.randn(Nn, R, device=dev).bfloat16()
    bias = torch.randn(Nn, device=dev).bfloat16()
    ms = timed(lambda: ops.lora_linear(x, w, down, up, bias, 1 / 32))
    print(f"kac640: {ms * 1e3:.1f} us, {(2.0 * M * Nn * K + 2.0 * M * R * (K + Nn)) / ms / 1e9:.1f} TFLOP/s")
elif which == "kaqkv":
    M, Nn, K, R = 4096, 1280, 1280, 32
    x = torch.randn(M, K, device=dev).bfloat16()
    ws = [(torch.randn(Nn, K, device=dev) / 36).bfloat16() for _ in range(3)]
    dn = [(torch.randn(R, K, device=dev) / 36).bfloat16() for _ in range(3)]
    up = [torch.randn(Nn, R, device=dev).bfloat16() for _ in range(3)]
    ms = timed(lambda: ops.lora_linear_grouped([x] * 3, ws, dn, up, [None] * 3, [1 / 32] * 3))
    print(f"kaqkv: {ms * 1e3:.1f} us, {3 * (2.0 * M * Nn * K + 2.0 * M * R * (K + Nn)) / ms / 1e9:.1f} TFLOP/s")
elif which == "kgradb":
    # one batched launch: 32 leaves of the c1280 shape, each with its own operands (690 MB streamed)
    M, Nn, K, R = 4096, 1280, 1280, 32
    leaves = []
    for i in range(32):
        leaves.append((torch.randn(M, K, device=dev).bfloat16(), torch.randn(M, Nn, device=dev).bfloat16(), torch.randn(M, R, device=dev).bfloat16(),
                       torch.randn(M, R, device=dev).bfloat16(), torch.zeros(R, K, device=dev), torch.zeros(Nn, R, device=dev)))
    ms = timed(lambda: ops.lora_grad_batched(leaves))
    byts = 32 * (2.0 * M * (K + Nn) + 4.0 * M * R + 8.0 * R * (K + Nn))
    print(f"kgradb: {ms * 1e3:.1f} us for 32 leaves, {byts / ms / 1e6:.0f} GB/s")
elif which in ("adafactor", "came"):
    # 64 matrices of 2048 x 2048 fp32 (268 M elements) through the reference-facing optimizer classes
    ps = [torch.nn.Parameter(torch.randn(2048, 2048, device=dev) * 0.02) for _ in range(64)]
    for p in ps:
        p.grad = torch.randn_like(p) * 1e-3
    if which == "adafactor":
        from transformers.optimization import Adafactor
        from onetrainer_b200.modules.util.optimizer.adafactor_extensions import patch_adafactor
        opt = Adafactor(ps, lr=1e-4, relative_step=False, scale_parameter=False, warmup_init=False)
        patch_adafactor(opt, False)
        bpe = 4 + 4 + 4 + 4           # two passes over g (4+4), p read + write (4+4); factors are O(rows + cols)
    else:
        from onetrainer_b200.modules.util.optimizer.CAME import CAME
        opt = CAME(ps, lr=1e-4)
        bpe = 4 + 4 + 4 + 4 + 8       # + exp_avg read + write
    ms = timed(lambda: opt.step())
    n = sum(p.numel() for p in ps)
    print(f"{which}: {ms:.3f} ms for {n} fp32 elements, {n * bpe / ms / 1e6:.0f} GB/s at {bpe} B/element (host launch included)")
elif which == "dora":
    Nn, J = 10240, 1280
    w2 = (torch.randn(Nn, J, device=dev) / 36).bfloat16(); ba = torch.randn(Nn, J, device=dev) * 1e-3
    m = torch.ones(J, device=dev)
    ms = timed(lambda: ops.dora_compose(w2, ba, m, 1 / 32))
    print(f"dora_compose: {ms * 1e3:.1f} us, {Nn * J * (2 + 4 + 2 + 2 + 6) / ms / 1e6:.0f} GB/s (two passes over w/ba, two bf16 outputs)")
