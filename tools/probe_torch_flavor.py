"""Which arithmetic does torch eager use for the AdamW ops?  Compares torch's result of each op
(on the given device) with candidate formulas evaluated through fp64 (exact products, one final
rounding), and prints bit-mismatch counts.  Used to set OTB_FLAVOR_TORCH_CPU / _CUDA."""
import sys
import torch

dev = torch.device(sys.argv[1] if len(sys.argv) > 1 else "cpu")
torch.manual_seed(0)
n = 1 << 20
f32, f64 = torch.float32, torch.float64


def mism(a, b):
    return int((a.view(torch.int32) != b.view(torch.int32)).sum())


def fma(a, b, c):   # round(a*b + c) with a single rounding (fp64 product of two fp32 is exact; sum rounded twice: rare)
    return (a.to(f64) * b.to(f64) + c.to(f64)).to(f32)


p = (torch.randn(n, device=dev) * 0.02)
g = (torch.randn(n, device=dev) * 1e-3)
m = (torch.randn(n, device=dev) * 1e-3)
v = (torch.rand(n, device=dev) * 1e-6)
d = v.sqrt() + 1e-8
alpha = -3.3e-4
af = torch.tensor(alpha, dtype=f32, device=dev)

# lerp
w = 1 - 0.9
wf = torch.tensor(w, dtype=f32, device=dev)
ref = m.clone().lerp_(g, w)
print("lerp   fma(w,g-m,m):", mism(ref, fma(wf, g - m, m)), " m+w*(g-m):", mism(ref, m + wf * (g - m)))

# addcmul
b2 = 1 - 0.999
bf = torch.tensor(b2, dtype=f32, device=dev)
ref = v.clone().addcmul_(g, g, value=b2)
print("addcmul fma(b*g,g,v):", mism(ref, fma(bf * g, g, v)), " fma(b,g*g,v):", mism(ref, fma(bf, g * g, v)),
      " v+(b*g)*g:", mism(ref, v + (bf * g) * g), " v+b*(g*g):", mism(ref, v + bf * (g * g)))

# div by scalar
c = 0.0316
cf = torch.tensor(c, dtype=f32, device=dev)
s = v.sqrt()
ref = s / c
print("div scalar  s/c:", mism(ref, s / cf), " s*(1/c):", mism(ref, s * (1.0 / cf)))

# addcdiv
ref = p.clone().addcdiv_(m, d, value=alpha)
print("addcdiv fma(a,m/d,p):", mism(ref, fma(af, m / d, p)), " p+a*(m/d):", mism(ref, p + af * (m / d)),
      " p+(a*m)/d:", mism(ref, p + (af * m) / d), " fma(a*m,1/d,p):", mism(ref, fma(af * m, 1.0 / d, p)))

# add with alpha (Adam L2)
wd = 0.01
wdf = torch.tensor(wd, dtype=f32, device=dev)
ref = g.add(p, alpha=wd)
print("add alpha fma(wd,p,g):", mism(ref, fma(wdf, p, g)), " g+wd*p:", mism(ref, g + wdf * p))

# mul scalar
ref = p * (1 - 3e-4 * 0.01)
print("mul scalar p*float(s):", mism(ref, p * torch.tensor(1 - 3e-4 * 0.01, dtype=f32, device=dev)))

# rsqrt / sigmoid for the loss weights
x = torch.rand(n, device=dev) * 1000 + 1e-3
print("rsqrt vs 1/sqrt:", mism(torch.rsqrt(x), 1.0 / torch.sqrt(x)), " vs fp64:", mism(torch.rsqrt(x), (1.0 / x.to(f64).sqrt()).to(f32)))
z = torch.randn(n, device=dev) * 2
print("sigmoid vs fp64:", mism(torch.sigmoid(z), torch.sigmoid(z.to(f64)).to(f32)))

# bf16 variants of the same ops (opmath float, single rounding to bf16)
pb, gb, mb, vb = (t.bfloat16() for t in (p, g, m, v))
db = (vb.sqrt() + 1e-8)
def mism16(a, b):
    return int((a.view(torch.int16) != b.view(torch.int16)).sum())
ref = vb.clone().addcmul_(gb, gb, value=b2)
print("bf16 addcmul fma(b*g,g,v):", mism16(ref, fma(bf * gb.float(), gb.float(), vb.float()).bfloat16()),
      " fma(b,g*g,v):", mism16(ref, fma(bf, gb.float() * gb.float(), vb.float()).bfloat16()))
ref = pb.clone().addcdiv_(mb, db, value=alpha)
print("bf16 addcdiv fma(a,m/d,p):", mism16(ref, fma(af, mb.float() / db.float(), pb.float()).bfloat16()),
      " p+(a*m)/d:", mism16(ref, (pb.float() + (af * mb.float()) / db.float()).bfloat16()))
ref = mb.clone().lerp_(gb, w)
print("bf16 lerp fma:", mism16(ref, fma(wf, gb.float() - mb.float(), mb.float()).bfloat16()))
ref = gb.add(pb, alpha=wd)
wb = torch.tensor(wd, dtype=torch.bfloat16, device=dev).float()
print("bf16 add alpha: fma(wd_f32):", mism16(ref, fma(wdf, pb.float(), gb.float()).bfloat16()),
      " fma(wd_bf16):", mism16(ref, fma(wb, pb.float(), gb.float()).bfloat16()))
# candidates for div by python scalar
import numpy as np
inv_a = torch.tensor(float(np.float32(1.0) / np.float32(c)), dtype=f32, device=dev)
inv_b = torch.tensor(float(np.float32(1.0 / c)), dtype=f32, device=dev)
print("div scalar: s*(1f/float(c)):", mism(s / c, s * inv_a), " s*float(1/c):", mism(s / c, s * inv_b), " inv equal:", bool(inv_a == inv_b))
# the SR path: result = p.float(); result.addcdiv_(m_bf16, d_bf16, value)
ref = pb.float().addcdiv_(mb, db, value=alpha)
print("sr-path addcdiv fma(a,m/d,p):", mism(ref, fma(af, mb.float() / db.float(), pb.float())),
      " p+(a*m)/d:", mism(ref, pb.float() + (af * mb.float()) / db.float()))
