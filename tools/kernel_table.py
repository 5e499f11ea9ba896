"""Per-kernel markdown table from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --csv` log (long format: one row per launch and metric).
usage: python tools/kernel_table.py <csv> [hbm_peak_GBps]"""
import csv
import re
import sys
from collections import OrderedDict, defaultdict

path = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6540.0
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ix = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Grid Size", "Metric Name", "Metric Unit", "Metric Value")}
launch = OrderedDict()
for r in rows[1:]:
    d = launch.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0}.get(u, 1.0)
    d[r[ix["Metric Name"]]] = v * scale


def label(d):
    n = re.sub(r"\(.*", "", d["name"])
    n = re.sub(r"^(void )?(pcrl::)?(tc2?::|tcg::)?", "", n)
    if n.startswith("tc_gemm_kernel"):
        g = int(re.findall(r"\d+", d["grid"])[0])
        big = d.get("dram__bytes_read.sum", 0) > 20e6
        return "tc_gemm_kernel (PointNet backward, ~100 k rows)" if big else "tc_gemm_kernel (MLP heads, M <= 512)"
    if n.startswith("pointnet_fwd_tc_kernel"):
        return "pointnet_fwd_tc_kernel (dump mode: backward recompute)"
    if n.startswith("pointnet_fwd_tc2_kernel"):
        half = d.get("dram__bytes_read.sum", 0) < 15e6
        arg = "arg-max" if re.search(r", *(1|true)>", n) else "values"
        return f"pointnet_fwd_tc2_kernel ({arg}{', 256 clouds' if half else ', 512 clouds'})"
    if "vectorized_elementwise" in n or "elementwise" in n:
        return "torch fill / add (elementwise)"
    return n


groups = defaultdict(list)
for d in launch.values():
    groups[label(d)].append(d)
total = sum(d["gpu__time_duration.sum"] for d in launch.values())
tot_rd = sum(d.get("dram__bytes_read.sum", 0) for d in launch.values())
tot_wr = sum(d.get("dram__bytes_write.sum", 0) for d in launch.values())
print("| kernel | launches | µs / launch | share | DRAM MB / launch (rd + wr) | DRAM GB/s | % of HBM peak | tensor pipe active |")
print("|---|---|---|---|---|---|---|---|")
for k, ds in sorted(groups.items(), key=lambda kv: -sum(d["gpu__time_duration.sum"] for d in kv[1])):
    t = sum(d["gpu__time_duration.sum"] for d in ds)
    rd = sum(d.get("dram__bytes_read.sum", 0) for d in ds)
    wr = sum(d.get("dram__bytes_write.sum", 0) for d in ds)
    tp = sum(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0) * d["gpu__time_duration.sum"] for d in ds) / t
    gbs = (rd + wr) / t / 1e3
    print(f"| `{k}` | {len(ds)} | {t / len(ds):.1f} | {100 * t / total:.1f} % | {rd / len(ds) / 1e6:.1f} + {wr / len(ds) / 1e6:.1f} | "
          f"{gbs:.0f} | {100 * gbs / peak:.0f} % | {tp:.1f} % |")
print(f"\n{len(launch)} launches, {total:.0f} µs serialised; DRAM read {tot_rd / 1e6:.0f} MB + write {tot_wr / 1e6:.0f} MB per update pair")
