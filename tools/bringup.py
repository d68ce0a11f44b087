"""GPU bring-up probe: runs each kernel against plain torch ops on the device and prints error
statistics.  Not a test (tests/ holds those) — a fast way to bisect a new kernel on a gpurun box:

    for g in info philox gemm grad optim noise loss; do timeout 300 python tools/bringup.py $g; done
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from onetrainer_b200 import _native as N  # noqa: E402
from onetrainer_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def philox_np(c0, c1, c2, c3, k0, k1):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return np.stack([c0, c1, c2, c3], -1).astype(np.uint32)


def g_info():
    import ctypes as C
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    rc = N.lib().otb_device_info(C.byref(a), C.byref(b), C.byref(c))
    print("device_info rc", rc, "sms", a.value, "cc", b.value, c.value, torch.cuda.get_device_name(0))


def g_philox():
    kat = philox_np([0], [0], [0], [0], 0, 0)[0]
    print("KAT zero:", [hex(v) for v in kat], "expect 6627e8d5 e169c58d bc57ac4c 9b00dbd8")
    seed, off, stream = 0x1234567890ABCDEF, 0x1_0000_0007, 4
    out = ops.philox_dump(seed, off, stream, 5, 1000, dev).cpu().numpy().view(np.uint32)
    grp = np.arange(5, 1005, dtype=np.uint64)
    ref = philox_np(grp & 0xFFFFFFFF, grp >> 32, np.full_like(grp, off & 0xFFFFFFFF),
                    np.full_like(grp, (stream << 24) | ((off >> 32) & 0xFFFFFF)), seed & 0xFFFFFFFF, seed >> 32)
    print("philox device == numpy:", bool((out == ref).all()))


def ref_lora(x, w, down, up, bias, s):
    y = x.float() @ w.float().T
    if bias is not None:
        y = y + bias.float()
    t = None
    if down is not None:
        t = (s * (x.float() @ down.float().T)).bfloat16()
        y = y + t.float() @ up.float().T
    return y, t


def g_gemm():
    torch.manual_seed(0)
    cases = [
        (128, 64, 64, 16, 0), (128, 64, 64, 0, 0), (256, 256, 128, 32, 0), (300, 328, 136, 16, 0),
        (4096, 1280, 1280, 32, 0), (308, 640, 2048, 32, 0), (4096, 1280, 1280, 32, 144), (4096, 1280, 1280, 32, 224),
        (16384, 640, 640, 32, 0), (4096, 10240, 1280, 32, 0), (4096, 1280, 5120, 32, 0), (1000, 320, 320, 16, 0),
        (128, 64, 64, 8, 0), (512, 512, 512, 64, 0),
    ]
    for (M, Nn, K, R, tn) in cases:
        x = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(Nn, K, device=dev) / K ** 0.5).bfloat16()
        bias = torch.randn(Nn, device=dev).bfloat16()
        down = (torch.randn(R, K, device=dev) / K ** 0.5).bfloat16() if R else None
        up = (torch.randn(Nn, R, device=dev) * 0.5).bfloat16() if R else None
        s = 0.7
        y, t = ops.lora_linear(x, w, down, up, bias, s, tile_n=tn)
        torch.cuda.synchronize()
        yr, tr = ref_lora(x, w, down, up, bias, s)
        err = (y.float() - yr).abs().max().item()
        scale = yr.abs().max().item()
        terr = (t.float() - tr.float()).abs().max().item() if R else 0.0
        # timing
        for _ in range(3):
            ops.lora_linear(x, w, down, up, bias, s, tile_n=tn)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 20
        for _ in range(iters):
            ops.lora_linear(x, w, down, up, bias, s, tile_n=tn)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 2.0 * M * Nn * K + 2.0 * M * R * (K + Nn)
        print(f"gemm M={M} N={Nn} K={K} R={R} tn={tn}: max|err|={err:.4g} (|y|max={scale:.3g}) t_err={terr:.3g} "
              f"{ms * 1e3:.1f} us {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
    # dropout determinism + rate
    M, Nn, K, R = 1024, 256, 256, 32
    x = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(Nn, K, device=dev).bfloat16() / 16
    down = torch.randn(R, K, device=dev).bfloat16() / 16; up = torch.randn(Nn, R, device=dev).bfloat16()
    _, t1 = ops.lora_linear(x, w, down, up, None, 1.0, dropout_p=0.25, seed=7, offset=3)
    _, t2 = ops.lora_linear(x, w, down, up, None, 1.0, dropout_p=0.25, seed=7, offset=3)
    _, t0 = ops.lora_linear(x, w, down, up, None, 1.0)
    print("dropout: same key equal:", bool((t1 == t2).all().item()), "zero frac:", (t1 == 0).float().mean().item(),
          "kept scaled ok:", torch.allclose(t1[t1 != 0].float(), (t0.float() / 0.75)[t1 != 0], rtol=2e-2, atol=1e-3))


def g_grad():
    torch.manual_seed(1)
    for (M, Nn, K, R) in [(128, 64, 64, 16), (4096, 1280, 1280, 32), (308, 640, 2048, 32), (16384, 640, 640, 32),
                          (1000, 328, 136, 16), (4096, 10240, 1280, 32), (777, 64, 72, 8)]:
        x = torch.randn(M, K, device=dev).bfloat16(); dy = torch.randn(M, Nn, device=dev).bfloat16()
        gs = torch.randn(M, R, device=dev).bfloat16(); ts = torch.randn(M, R, device=dev).bfloat16()
        dA = torch.zeros(R, K, device=dev); dB = torch.zeros(Nn, R, device=dev)
        ops.lora_grad_ab(x, dy, gs, ts, dA, dB)
        torch.cuda.synchronize()
        dAr = gs.float().T @ x.float(); dBr = dy.float().T @ ts.float()
        ea = (dA - dAr).abs().max().item() / dAr.abs().max().item()
        eb = (dB - dBr).abs().max().item() / dBr.abs().max().item()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.lora_grad_ab(x, dy, gs, ts, dA, dB)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        gb = (M * (K + Nn) * 2 + 2 * M * R * 2) / 1e9
        print(f"grad M={M} N={Nn} K={K} R={R}: rel err dA={ea:.3g} dB={eb:.3g} {ms * 1e3:.1f} us {gb / ms * 1e3:.0f} GB/s", flush=True)


def torch_adamw_ref(p, g, m, v, lr, b1, b2, eps, wd, step):
    """The reference's op sequence (adamw_extensions.py:79-146, non-SR branch) as torch eager ops."""
    import math
    p.mul_(1 - lr * wd)
    m.lerp_(g, 1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step; bc2 = 1 - b2 ** step
    step_size = lr / bc1
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-step_size)


def g_optim():
    torch.manual_seed(2)
    lr, b1, b2, eps, wd = 3e-4, 0.9, 0.999, 1e-8, 1e-2
    for dt in (torch.bfloat16, torch.float32):
        for flavor_name, flavor in (("cuda", N.FLAVOR_TORCH_CUDA), ("cpu", N.FLAVOR_TORCH_CPU)):
            sizes = [40000, 5120, 131072 + 8, 77]
            ps = [(torch.randn(n, device=dev) * 0.02).to(dt) for n in sizes]
            ms = [(torch.randn(n, device=dev) * 1e-3).to(dt) for n in sizes]
            vs = [(torch.rand(n, device=dev) * 1e-6).to(dt) for n in sizes]
            es = [p.clone() + 0.01 for p in ps]
            ref = [[t.clone() for t in ts] for ts in (ps, ms, vs, es)]
            mism = {"p": 0, "m": 0, "v": 0, "e": 0}
            total = 0
            for step in range(1, 4):
                gs = [(torch.randn(n, device=dev) * 1e-3).to(dt) for n in sizes]
                entries = [dict(p=p, g=g, exp_avg=m, exp_avg_sq=v, ema=e) for p, g, m, v, e in zip(ps, gs, ms, vs, es)]
                tab = ops.OptimTable(entries, dev)
                sc = ops.adam_scalars(lr, b1, b2, eps, wd, float(step), True)
                for k, val in sc.items():
                    tab.host[k][:] = val
                tab.upload()
                omd = 1 - min((1 + step) / (10 + step), 0.999)
                ops.adamw_sr_ema_multi(tab, flavor=flavor, do_ema=True, ema_one_minus_decay=omd)
                for i in range(len(sizes)):
                    torch_adamw_ref(ref[0][i], gs[i], ref[1][i], ref[2][i], lr, b1, b2, eps, wd, float(step))
                    ref[3][i].add_(omd * (ref[0][i] - ref[3][i]))
                torch.cuda.synchronize()
                for key, mine, r in (("p", ps, ref[0]), ("m", ms, ref[1]), ("v", vs, ref[2]), ("e", es, ref[3])):
                    for a, b in zip(mine, r):
                        mism[key] += int((a.view(torch.int16 if dt == torch.bfloat16 else torch.int32)
                                          != b.view(torch.int16 if dt == torch.bfloat16 else torch.int32)).sum().item())
                total += sum(sizes)
                # resync so mismatches do not cascade
                for mine, r in ((ps, ref[0]), (ms, ref[1]), (vs, ref[2]), (es, ref[3])):
                    for a, b in zip(mine, r):
                        b.copy_(a)
            print(f"optim {dt} flavor={flavor_name}: bit mismatches vs torch CUDA eager over {total} elems: {mism}", flush=True)
    # SR unbiasedness: p = 1 + 2^-10 (between bf16 neighbours 1 and 1.0078125), zero update
    n = 1 << 20
    p = torch.full((n,), 1.0, device=dev, dtype=torch.bfloat16)
    g = torch.zeros(n, device=dev, dtype=torch.bfloat16)
    m = torch.full((n,), 1.0, device=dev, dtype=torch.bfloat16)      # exp_avg=1 -> update = -step_size * m/(sqrt(v)+eps)
    v = torch.full((n,), 1.0, device=dev, dtype=torch.bfloat16)
    tab = ops.OptimTable([dict(p=p, g=g, exp_avg=m, exp_avg_sq=v)], dev)
    tab.host["flags"][:] |= N.OPT_FLAG_SR
    for k, val in dict(decay_mul=1.0, l2=0.0, lerp_w=0.0, beta2=1.0, one_minus_beta2=0.0, bc2_sqrt=1.0, inv_bc2_sqrt=1.0,
                       eps=0.0, neg_step_size=2.0 ** -10).items():
        tab.host[k][:] = val
    tab.upload()
    ops.adamw_sr_ema_multi(tab, seed=123, offset=1)
    torch.cuda.synchronize()
    print("SR mean of 1+2^-10:", p.float().mean().item(), "expect ~1.0009765625; values:", torch.unique(p.float()).tolist())
    # bandwidth
    for dt, bpe in ((torch.bfloat16, 18), (torch.float32, 36)):
        n = 1 << 28 if dt == torch.bfloat16 else 1 << 27
        p = torch.zeros(n, device=dev, dtype=dt); g = torch.zeros_like(p); m = torch.zeros_like(p); v = torch.ones_like(p); e = torch.zeros_like(p)
        tab = ops.OptimTable([dict(p=p, g=g, exp_avg=m, exp_avg_sq=v, ema=e)], dev)
        if dt == torch.bfloat16:
            tab.host["flags"][:] |= N.OPT_FLAG_SR
        for k, val in ops.adam_scalars(lr, b1, b2, eps, wd, 1.0, True).items():
            tab.host[k][:] = val
        tab.upload()
        for do_ema in (True, False):
            for _ in range(2):
                ops.adamw_sr_ema_multi(tab, do_ema=do_ema, ema_one_minus_decay=0.001, seed=1, offset=2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.adamw_sr_ema_multi(tab, do_ema=do_ema, ema_one_minus_decay=0.001, seed=1, offset=2)
            e1.record(); torch.cuda.synchronize()
            ms_ = e0.elapsed_time(e1) / 5
            b = bpe if do_ema else (14 if dt == torch.bfloat16 else 28)
            print(f"K-c {dt} ema={do_ema}: {ms_:.3f} ms for {n} elems -> {n * b / ms_ / 1e6:.0f} GB/s", flush=True)
        del p, g, m, v, e, tab
    # grad norm
    gs = [torch.randn(n, device=dev) for n in (1000, 123457, 1 << 22)]
    tab = ops.OptimTable([dict(p=torch.empty_like(g), g=g) for g in gs], dev)
    tab.upload()
    nc = ops.grad_norm_multi(tab, 1.0).cpu()
    ref = torch.nn.utils.get_total_norm(gs) if hasattr(torch.nn.utils, "get_total_norm") else torch.linalg.vector_norm(torch.stack([g.norm() for g in gs]))
    print("grad norm:", nc.tolist(), "torch:", ref.item())


def g_noise():
    B, Cc, H, W = 4, 4, 128, 128
    T = 1000
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, T, dtype=torch.float32) ** 2
    ac = torch.cumprod(1 - betas, 0)
    sa, s1 = ac.sqrt().to(dev), (1 - ac).sqrt().to(dev)
    lat = torch.randn(B, Cc, H, W, device=dev)
    out = ops.noise_addnoise(latent=lat, coef_x0=sa, coef_noise=s1, B=B, C_=Cc, hw=H * W, latent_dtype=torch.float32,
                             device=dev, want_x0=True, want_noise=True, want_noisy=True, cast_dtype=torch.bfloat16,
                             target=N.TARGET_V, ts_mode=N.TS_UNIFORM, latent_scale=0.13025, seed=5, offset=0, shape=lat.shape)
    torch.cuda.synchronize()
    n = out["noise"]
    print("noise mean/std/kurt:", n.mean().item(), n.std().item(), ((n - n.mean()) ** 4).mean().item() / n.var().item() ** 2,
          "timesteps", out["timestep"].tolist())
    t = out["timestep"].long()
    x0 = lat * 0.13025
    a, b = sa[t].view(-1, 1, 1, 1), s1[t].view(-1, 1, 1, 1)
    noisy = x0 * a + n * b
    v = a * n - b * x0
    print("x0 exact:", bool((out["x0"] == x0).all()), "noisy exact:", bool((out["noisy"] == noisy).all()),
          "v exact:", bool((out["target"] == v).all()), "cast exact:", bool((out["noisy_cast"] == noisy.bfloat16()).all()))
    # parity mode (given noise + timesteps)
    out2 = ops.noise_addnoise(latent=lat, noise_in=n, coef_x0=sa, coef_noise=s1, timestep=out["timestep"], B=B, C_=Cc, hw=H * W,
                              latent_dtype=torch.float32, device=dev, want_noisy=True, target=N.TARGET_EPSILON,
                              ts_mode=N.TS_GIVEN, generate_noise=False, latent_scale=0.13025, shape=lat.shape)
    print("parity mode noisy exact:", bool((out2["noisy"] == noisy).all()), "eps target exact:", bool((out2["target"] == n).all()))
    # DP invariance: rank 1 of 2 draws samples 2..3
    out3 = ops.noise_addnoise(latent=lat[2:], coef_x0=sa, coef_noise=s1, B=2, C_=Cc, hw=H * W, latent_dtype=torch.float32,
                              device=dev, want_noise=True, ts_mode=N.TS_UNIFORM, latent_scale=0.13025, seed=5, offset=0,
                              sample_base=2, shape=lat[2:].shape)
    print("DP invariance noise:", bool((out3["noise"] == n[2:]).all()), "timesteps:", out3["timestep"].tolist())
    # bandwidth at N = 2^28
    Bb = 4096
    lat = torch.randn(Bb, 4, 128, 128, device=dev)
    for _ in range(2):
        o = ops.noise_addnoise(latent=lat, coef_x0=sa, coef_noise=s1, B=Bb, C_=4, hw=128 * 128, latent_dtype=torch.float32,
                               device=dev, cast_dtype=torch.bfloat16, target=N.TARGET_EPSILON, ts_mode=N.TS_UNIFORM, seed=1, shape=lat.shape)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        o = ops.noise_addnoise(latent=lat, coef_x0=sa, coef_noise=s1, B=Bb, C_=4, hw=128 * 128, latent_dtype=torch.float32,
                               device=dev, cast_dtype=torch.bfloat16, target=N.TARGET_EPSILON, ts_mode=N.TS_UNIFORM, seed=1, shape=lat.shape)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"K-b1 N=2^28: {ms:.3f} ms (incl. allocs) -> {lat.numel() * 10 / ms / 1e6:.0f} GB/s")


def g_loss():
    torch.manual_seed(3)
    B, Cc, H, W = 4, 4, 128, 128
    T = 1000
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, T, dtype=torch.float32) ** 2
    ac = torch.cumprod(1 - betas, 0)
    sa, s1 = ac.sqrt().to(dev), (1 - ac).sqrt().to(dev)
    pred = torch.randn(B, Cc, H, W, device=dev).bfloat16().requires_grad_(True)
    tgt = torch.randn(B, Cc, H, W, device=dev)
    prior = torch.randn(B, Cc, H, W, device=dev)
    mask = (torch.rand(B, 1, H, W, device=dev) > 0.5).float()
    lw = torch.tensor([1.0, 0.5, 2.0, 1.0], device=dev)
    ts = torch.tensor([10, 499, 800, 999], device=dev, dtype=torch.int32)
    for name, kw in (("mse unmasked min-snr", dict(loss_weight_fn=N.LW_MIN_SNR)),
                     ("masked norm prior mae debiased v", dict(masked=True, normalize_masked_area_loss=True, mae_strength=0.3,
                                                                masked_prior_preservation_weight=0.4, loss_weight_fn=N.LW_DEBIASED, v_prediction=True))):
        up = 1.0 / B
        losses, dpred = ops.loss_fwd_bwd(pred.detach(), tgt, prior_target=prior if kw.get("masked") else None,
                                         mask=mask if kw.get("masked") else None, loss_weight=lw, timestep=ts, sqrt_ac=sa,
                                         sqrt_1mac=s1, upstream=up, **kw)
        # torch reference
        p32 = pred.float()
        if kw.get("masked"):
            cm = mask.clamp(0.1, 1)
            l = (p32 - tgt) ** 2 * cm / cm.mean(dim=(1, 2, 3), keepdim=True)
            l = l + (p32 - prior) ** 2 * ((1 - cm) * 0.4) / (1 - cm).mean(dim=(1, 2, 3), keepdim=True)
            l1 = (p32 - tgt).abs() * cm / cm.mean(dim=(1, 2, 3), keepdim=True)
            l1 = l1 + (p32 - prior).abs() * ((1 - cm) * 0.4) / (1 - cm).mean(dim=(1, 2, 3), keepdim=True)
            ref = l.mean((1, 2, 3)) + 0.3 * l1.mean((1, 2, 3))
        else:
            ref = ((p32 - tgt) ** 2).mean((1, 2, 3))
        ref = ref * lw
        snr = (sa / s1) ** 2
        s = snr[ts.long()]
        if kw["loss_weight_fn"] == N.LW_MIN_SNR:
            ref = ref * (torch.minimum(s, torch.full_like(s, 5.0)) / s)
        else:
            ref = ref * torch.rsqrt(s.clamp(max=1e3) + 1.0)
        (g,) = torch.autograd.grad(ref.mean(), pred)
        torch.cuda.synchronize()
        rel_l = ((losses - ref).abs() / ref.abs()).max().item()
        gd = (dpred.float() - g.float()).abs().max().item() / g.float().abs().max().item()
        print(f"loss[{name}]: rel err losses {rel_l:.3g}; dpred max rel-to-max err {gd:.3g}; bit-equal frac "
              f"{(dpred.view(torch.int16) == g.view(torch.int16)).float().mean().item():.5f}", flush=True)
    # bandwidth at 2^28
    Bb = 4096
    pred = torch.randn(Bb, 4, 128, 128, device=dev, dtype=torch.bfloat16)
    tgt = torch.randn(Bb, 4, 128, 128, device=dev)
    ts = torch.randint(0, 1000, (Bb,), device=dev, dtype=torch.int32)
    for _ in range(2):
        ops.loss_fwd_bwd(pred, tgt, timestep=ts, sqrt_ac=sa, sqrt_1mac=s1, loss_weight_fn=N.LW_MIN_SNR, upstream=1 / Bb)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.loss_fwd_bwd(pred, tgt, timestep=ts, sqrt_ac=sa, sqrt_1mac=s1, loss_weight_fn=N.LW_MIN_SNR, upstream=1 / Bb)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"K-b2 N=2^28: {ms:.3f} ms (incl. allocs) -> {pred.numel() * 8 / ms / 1e6:.0f} GB/s")


if __name__ == "__main__":
    t0 = time.time()
    for g in sys.argv[1:]:
        print(f"==== {g} ====", flush=True)
        globals()["g_" + g]()
    print(f"done in {time.time() - t0:.1f}s")
