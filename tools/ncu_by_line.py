"""Attribute the per-SASS-instruction counters of an .ncu-rep to source lines, using nvdisasm line info of the
same build.  Usage: ncu_by_line.py file.ncu-rep lib.so kernel-mangled-name"""
import csv, subprocess, sys, io, re, os, tempfile, collections, glob
rep, lib, kname = sys.argv[1], sys.argv[2], sys.argv[3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = []
for cub in glob.glob(os.path.join(tmp, "*icp*.cubin")):
    out = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout.split("\n")
    starts = [i for i, l in enumerate(out) if l.startswith(".text.")]
    for i in starts:
        if out[i] == f".text.{kname}:":
            lines = out[i:min([j for j in starts if j > i] + [len(out)])]
cur = None; attrib = []
for l in lines:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,5}\*/', l): attrib.append((cur, l.strip()))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = [r for r in rows[2:] if len(r) == len(h)]
print("sass rows", len(data), "disasm instr", len(attrib))
iex = h.index('Instructions Executed'); isamp = h.index('# Samples')
by = collections.defaultdict(lambda: [0, 0])
n = min(len(data), len(attrib))
for k in range(n):
    key = attrib[k][0]
    by[key][0] += int(data[k][iex] or 0); by[key][1] += int(data[k][isamp] or 0)
tot = sum(v[0] for v in by.values()); ts = sum(v[1] for v in by.values())
print("by file:")
bf = collections.defaultdict(lambda: [0, 0])
for (f, l), v in [(k, v) for k, v in by.items() if k]: bf[f][0] += v[0]; bf[f][1] += v[1]
for f, v in sorted(bf.items(), key=lambda x: -x[1][0]): print(f"  {f:28s} inst {100*v[0]/tot:5.1f}%  samples {100*v[1]/ts:5.1f}%")
print("top lines by instructions:")
for k, v in sorted(by.items(), key=lambda x: -x[1][0])[:28]: print(f"  {str(k):40s} inst {100*v[0]/tot:5.1f}%  samples {100*v[1]/ts:5.1f}%")
print("top lines by samples:")
for k, v in sorted(by.items(), key=lambda x: -x[1][1])[:22]: print(f"  {str(k):40s} inst {100*v[0]/tot:5.1f}%  samples {100*v[1]/ts:5.1f}%")
