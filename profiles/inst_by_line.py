"""Instructions executed per CUDA source line / function from an ncu report (needs the matching .so for the line table).
usage: python profiles/inst_by_line.py <report.ncu-rep> <kernel-substring> [top]"""
import collections, csv, io, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
run = lambda cmd: subprocess.run(cmd, capture_output=True, text=True).stdout
so = os.path.join(ROOT, "muzero.jl_b200", "libmuzero_b200.so")
tmp = "/tmp/mz_elf_i"
subprocess.run("rm -rf %s && mkdir -p %s && cd %s && cuobjdump -xelf all %s >/dev/null 2>&1" % (tmp, tmp, tmp, so), shell=True)
m, cur, infn = {}, None, False
for cub in os.listdir(tmp):
    for ln in run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)]).split("\n"):
        if ln.startswith(".text."):
            infn = kern in ln; continue
        if not infn: continue
        g = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if g: cur = (g.group(1).split("/")[-1], int(g.group(2))); continue
        g = re.match(r"\s*/\*([0-9a-f]+)\*/", ln)
        if g: m[int(g.group(1), 16)] = cur
rows = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv"]))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
col = {h: i for i, h in enumerate(rows[hi])}
first, agg, tot = None, collections.Counter(), 0
for r in rows[hi + 1:]:
    try: a = int(r[col["Address"]], 16); n = int(r[col["Instructions Executed"]])
    except (ValueError, IndexError): continue
    if first is None: first = a
    agg[m.get(a - first, ("?", 0))] += n; tot += n
src = {}
d = os.path.join(ROOT, "muzero.jl_b200", "csrc")
for f in os.listdir(d):
    if os.path.isfile(os.path.join(d, f)): src[f] = open(os.path.join(d, f), errors="replace").read().split("\n")
print("warp instructions executed: %d" % tot)
for k, v in agg.most_common(top):
    text = src[k[0]][k[1] - 1].strip()[:110] if k[0] in src and 0 < k[1] <= len(src[k[0]]) else ""
    print("%10d %5.1f%% %s:%d %s" % (v, 100.0 * v / max(tot, 1), k[0], k[1], text))
