"""Summarise an ncu report: headline metrics + warp-stall samples attributed to CUDA source lines.
usage: python profiles/analyze_ncu.py <report.ncu-rep> <kernel-substring> <out.txt> [title]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern, out = sys.argv[1], sys.argv[2], sys.argv[3]
title = sys.argv[4] if len(sys.argv) > 4 else rep


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
hdr, units, vals = raw[0], raw[1], raw[2]
keep = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
lines = [title]
for i, h in enumerate(hdr):
    if h in keep:
        lines.append("%-70s %s %s" % (h, vals[i], units[i]))

# SASS offset -> CUDA line, from the cubin inside the .so
so = os.path.join(ROOT, "muzero.jl_b200", "libmuzero_b200.so")
tmp = "/tmp/mz_elf"
subprocess.run("rm -rf %s && mkdir -p %s && cd %s && cuobjdump -xelf all %s >/dev/null 2>&1" % (tmp, tmp, tmp, so), shell=True)
m, cur, infn = {}, None, False
for cub in os.listdir(tmp):
    for ln in run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)]).split("\n"):
        if ln.startswith(".text."):
            infn = kern in ln
            continue
        if not infn:
            continue
        g = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if g:
            cur = (g.group(1).split("/")[-1], int(g.group(2)))
            continue
        g = re.match(r"\s*/\*([0-9a-f]+)\*/", ln)
        if g:
            m[int(g.group(1), 16)] = cur
rows = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv"]))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h2 = rows[hi]; col = {h: i for i, h in enumerate(h2)}
first, agg, stall, tot, st_tot = None, collections.Counter(), collections.defaultdict(collections.Counter), 0, collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < len(h2):
        continue
    try:
        a = int(r[col["Address"]], 16); n = int(r[col["# Samples"]])
    except ValueError:
        continue
    if first is None:
        first = a
    key = m.get(a - first, ("?", 0)); agg[key] += n; tot += n
    for s in h2:
        if s.startswith("stall_") and "Not Issued" not in s and r[col[s]] not in ("", "0"):
            stall[key][s[6:]] += int(r[col[s]]); st_tot[s[6:]] += int(r[col[s]])
src = {}
for f in [x for x in os.listdir(os.path.join(ROOT, "muzero.jl_b200", "csrc")) if os.path.isfile(os.path.join(ROOT, "muzero.jl_b200", "csrc", x))]:
    src[f] = open(os.path.join(ROOT, "muzero.jl_b200", "csrc", f), errors="replace").read().split("\n")
lines.append("")
lines.append("warp-stall samples (total %d): %s" % (tot, ", ".join("%s %.1f%%" % (k, 100.0 * v / max(tot, 1)) for k, v in st_tot.most_common(8))))
lines.append("hottest CUDA source lines:")
for k, v in agg.most_common(30):
    text = src.get(k[0], [""] * (k[1] + 1))[k[1] - 1].strip()[:100] if k[0] in src and k[1] > 0 else ""
    lines.append("%6d %5.1f%% %s:%d [%s] %s" % (v, 100.0 * v / max(tot, 1), k[0], k[1], ",".join("%s=%d" % x for x in stall[k].most_common(3)), text))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
