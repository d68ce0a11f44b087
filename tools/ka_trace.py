"""Timeline of one K-a launch (bring-up): %globaltimer stamps of the first tile of clusters 0..7, relative to the end of
the PREVIOUS launch's last stamp, inside a chain of identical launches.  usage: python tools/ka_trace.py M N K [tile_n]"""
import os
import sys
import torch
sys.path.insert(0, ".")
from onetrainer_b200 import ops
dev = torch.device("cuda:0")
M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (4096, 1280, 1280)
tn = int(sys.argv[4]) if len(sys.argv) > 4 else 0
R = 32
os.environ["OTB_KA_VER"] = "3"
nrep = 6
xs = [torch.randn(M, K, device=dev).bfloat16() for _ in range(2)]
ws = [(torch.randn(N, K, device=dev) / K ** 0.5).bfloat16() for _ in range(nrep)]
a = (torch.randn(R, K, device=dev) / K ** 0.5).bfloat16(); b = torch.randn(N, R, device=dev).bfloat16()
bias = torch.randn(N, device=dev).bfloat16()
traces = [torch.zeros(8 * 16, dtype=torch.int64, device=dev) for _ in range(nrep)]


def run():
    for i in range(nrep):
        os.environ["OTB_KA_TRACE"] = hex(traces[i].data_ptr())
        ops.lora_linear(xs[i & 1], ws[i], a, b, bias, 1 / 32, tile_n=tn, static_weights=bool(int(os.environ.get('KA_STATIC', '0'))))
    os.environ.pop("OTB_KA_TRACE")


run(); torch.cuda.synchronize()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    run()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    run()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
names = ["entry", "prologue done", "prev launch complete (pdl wait)", "first operands landed", "last k-block issued", "mainloop complete (t_full)",
         "T parked (t_ready)", "accumulator complete (d_full)", "blocks handed to TMA", "stores written", "pair done"]
t = torch.stack(traces).cpu().view(nrep, 8, 16)
print(f"M={M} N={N} K={K} tile_n={tn}: launches 2..{nrep - 1} of a graph-replayed chain; microseconds after the previous launch's last 'pair done' (clusters 0..7)")
for i in range(2, nrep):
    prev_end = int(t[i - 1, :, 10].max())
    print(f"launch {i}: previous launch's slowest traced pair ended at 0; this launch:")
    for e, name in enumerate(names):
        vals = [(int(t[i, c, e]) - prev_end) / 1e3 for c in range(8) if int(t[i, c, e]) > 0]
        if vals:
            print(f"   {e:2d} {name:36s} min {min(vals):7.2f}  median {sorted(vals)[len(vals) // 2]:7.2f}  max {max(vals):7.2f}")
