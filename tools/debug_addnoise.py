import sys
import numpy as np, torch
sys.path.insert(0, ".")
from onetrainer_b200 import _native as N, ops
from oracle import otb_oracle as O
dev = torch.device("cuda:0")
g = np.load("tests/golden/addnoise.npz"); s = np.load("tests/golden/schedule.npz")
sa, s1 = torch.from_numpy(s["sqrt_ac"]).to(dev), torch.from_numpy(s["sqrt_1mac"]).to(dev)
lat = torch.from_numpy(g["latent"]).to(dev); noise = torch.from_numpy(g["noise"]).to(dev); t = torch.from_numpy(g["t"]).to(dev)
B, C, H, W = lat.shape
out = ops.noise_addnoise(latent=lat, noise_in=noise, coef_x0=sa, coef_noise=s1, timestep=t, B=B, C_=C, hw=H * W,
                         latent_dtype=torch.float32, device=dev, want_x0=True, want_noisy=True, cast_dtype=torch.bfloat16,
                         target=N.TARGET_V, ts_mode=N.TS_GIVEN, generate_noise=False, latent_scale=float(g["scale"][0]), shape=lat.shape)
for k, ref in (("x0", g["x0"]), ("noisy", g["noisy"]), ("target", g["velocity"])):
    mine = out[k].cpu().numpy()
    bad = np.argwhere(mine.view(np.uint32) != ref.view(np.uint32))
    print(k, "mismatches", len(bad), "of", mine.size, "first", bad[:5].tolist())
    for idx in bad[:3]:
        i = tuple(idx)
        print("   ", i, mine[i], ref[i], int(mine[i].view(np.uint32)) - int(ref[i].view(np.uint32)))
print("timesteps", out["timestep"].tolist(), g["t"].tolist())
x0o, no, vo = O.add_noise(g["latent"], g["noise"], g["t"], s["sqrt_ac"], s["sqrt_1mac"], latent_scale=float(np.float32(g["scale"][0])), target=O.TARGET_V)
print("oracle vs golden noisy", (no != g["noisy"].reshape(no.shape)).sum(), "kernel vs oracle noisy", (out["noisy"].cpu().numpy().reshape(no.shape) != no).sum())
tl = t.long()
a, b = sa[tl].view(-1, 1, 1, 1), s1[tl].view(-1, 1, 1, 1)
x0 = lat * float(g["scale"][0])
print("torch-gpu vs golden noisy", int((x0 * a + noise * b != torch.from_numpy(g["noisy"]).to(dev)).sum()), "x0", int((x0 != torch.from_numpy(g["x0"]).to(dev)).sum()))
