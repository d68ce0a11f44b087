"""Summarise ncu captures for profiles/ (run here, no GPU needed).

  python tools/ncu_summary.py rep  gpurun_out/prof_x.ncu-rep  > profiles/rNN_x.md      # --set full capture
  python tools/ncu_summary.py list gpurun_out/launches.csv    > profiles/rNN_launches.md # launch list

`rep` prints, per profiled launch, the handful of raw metrics the roofline argument needs; `list` groups the
`gpu__time_duration.sum` launch list by kernel name (count, total, share) — shares, not absolutes, are what the
bench line is compared with (cold cache, serialised)."""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

WANT = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_tc.sum", "sm__inst_executed_pipe_tmem.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of `{path.split('/')[-1]}`\n")
    tensor_cols = [h for h in hdr if re.fullmatch(r"sm__(mem_tensor_cycles_active|inst_executed_pipe_tmem|inst_executed_pipe_uniform)"
                                                  r"\.avg\.pct_of_peak_sustained_active", h)]
    for r in rows[2:]:
        d = OrderedDict(zip(hdr, r))
        print(f"## launch {d.get('ID')}: `{d.get('Kernel Name')}`\n")
        print("| metric | value | unit |\n|---|---|---|")
        for w in WANT + [c for c in tensor_cols if c not in WANT]:
            if w in d and d[w] not in ("", "n/a"):
                print(f"| {w} | {d[w]} | {units[hdr.index(w)]} |")
        try:
            tr = float(d["dram__bytes_read.sum"].replace(",", "")), float(d["dram__bytes_write.sum"].replace(",", ""))
            print(f"\ntraffic = read + write = {tr[0]:.4g} {units[hdr.index('dram__bytes_read.sum')]} + {tr[1]:.4g} "
                  f"{units[hdr.index('dram__bytes_write.sum')]}\n")
        except Exception:
            print()


def launches(path):
    txt = open(path).read()
    start = txt.find('"ID"')
    rows = list(csv.reader(io.StringIO(txt[start:])))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("<unnamed>::", "").replace("void ", "")
        name = re.sub(r"<.*", "", name)
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(r[ui], 1.0)
        v *= scale
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    print(f"# ncu launch list `{path.split('/')[-1]}` grouped by kernel ({sum(a[0] for a in agg.values())} launches, "
          f"{total / 1e3:.2f} ms summed, cold-cache + serialised: compare shares)\n")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t:.1f} | {t / total:.3f} |")


def lines(path, top=30):
    """Warp-stall samples per SOURCE line of a capture taken with --import-source on (kernels compiled with -lineinfo)."""
    import collections
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    agg, stall, total = collections.Counter(), collections.defaultdict(collections.Counter), 0
    cur, h, ci, kernel = None, None, {}, ""

    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0
    for r in rows:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0] == "Function Name":
            kernel = r[1]
            continue
        if r and r[0] == "Line No":
            h, ci = r, {}
            for k, n in enumerate(h):
                ci.setdefault(n, k)
            continue
        if h is None or len(r) < len(h) or r[2] != "-":
            continue
        n = num(r[ci["# Samples"]])
        if n == 0:
            continue
        key = (cur, num(r[0]), r[1].strip())
        agg[key] += n
        total += n
        for name, k in ci.items():
            if name.startswith("stall_") and "Not Issued" not in name and num(r[k]):
                stall[key][name] += num(r[k])
    print(f"# warp-stall samples per source line: `{path.split('/')[-1]}`\n\n`{kernel}` — {total} samples\n")
    print("| samples | share | line | source | top stall reasons |\n|---|---|---|---|---|")
    for key, n in agg.most_common(top):
        f, ln, src = key
        why = ", ".join(f"{a} {b}" for a, b in stall[key].most_common(2))
        print(f"| {n} | {n / total:.3f} | `{f}:{ln}` | `{src[:110].replace('|', '¦')}` | {why} |")


if __name__ == "__main__":
    {"rep": rep, "list": launches, "lines": lines}[sys.argv[1]](sys.argv[2])
