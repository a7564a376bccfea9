"""Turns the ncu artefacts of a gpurun call into the committed summaries under profiles/.
usage: make_profile_summary.py <tag> <launches.csv> <full.ncu-rep>"""
import csv, io, json, os, subprocess, sys, collections
tag, launches, rep = sys.argv[1:4]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
CMD = "python bench.py --steps 2 --warmup 3 --pool 2 --no-cpu-baseline --no-config4 --no-batch-regime"

rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = collections.OrderedDict()
seq = []
for r in rows[1:]:
    name = r[ki].split("(PersistArgs")[0].split("(const")[0].replace("void ", "")
    if not name.startswith("icp_persist_kernel"): name = name.split("(")[0]
    t = float(r[vi].replace(",", "")) / 1000.0
    a = tot.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
    seq.append((name, t))
total = sum(v[1] for v in tot.values())
with open(os.path.join(out_dir, f"{tag}_launches_summary.md"), "w") as f:
    f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none {CMD}`\n\n")
    f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes. The whole run is listed "
            "(priming + warm-up + timed steps + e2e leg, i.e. also the plane extraction and upload kernels of the e2e leg).\n\n")
    f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {v[0]} | {v[1]:.0f} | {100*v[1]/total:.1f}% | {v[1]/v[0]:.1f} |\n")
    its = [(n, t) for n, t in seq if n.startswith("icp_persist_kernel")]
    if its:
        # a registration = the search-only instance (template argument true) followed by the general instance
        regs = []
        for n, t in its:
            if n.rstrip().endswith(", 1>") or "(bool)1" in n or "true" in n or not regs: regs.append([t])
            else: regs[-1].append(t)
        f.write("\n`icp_persist_kernel` launches: one registration = 30 ICP iterations = the search-only instance (leading full-search "
                "iterations) + the general instance; us per registration (search-only + general):\n\n`"
                + " ".join(f"{sum(r):.0f}({'+'.join(f'{x:.0f}' for x in r)})" for r in regs) + "`\n")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h = rr[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__cycles_active.avg"]
def tobytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
dram = []
with open(os.path.join(out_dir, f"{tag}_icp_persist_full_summary.txt"), "w") as f:
    f.write(f"ncu --set full --clock-control none --import-source on -k regex:icp_persist -s 8 -c 4  ({CMD})\n")
    f.write("one row per captured launch; a registration of 30 iterations = a 768-thread launch (search-only instance: the leading full-search "
            "iterations) followed by a 512-thread launch (general instance)\n")
    f.write("columns: " + " | ".join(want) + "\n")
    for r in rr[2:]:
        f.write(" | ".join(r[h.index(w)] if w in h else "?" for w in want) + "\n")
        i1, i2 = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        b = tobytes(r[i1], rr[1][i1]) + tobytes(r[i2], rr[1][i2])
        if "launch__block_size" in h and r[h.index("launch__block_size")].strip() == "768" or not dram: dram.append(b)     # a new registration
        else: dram[-1] += b
    f.write("units: " + " | ".join(rr[1][h.index(w)] if w in h else "?" for w in want) + "\n")
    f.write(f"\nDRAM bytes (read+write) per registration (both launches): mean {sum(dram)/len(dram):.0f}, max {max(dram):.0f}; algorithmic bytes per registration "
            f"30 x 14745600 = 442368000.\n")
    f.write("ncu flushes the caches before every replay: the DRAM bytes are the cold first touch of the ~40 MB working set (source, cell-sorted "
            "target + normals, touched part of the cell array, 49 B of per-query state), paid again by the second launch because ncu flushes between launches; in a live run the general instance finds the working set in L2.\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows2 = list(csv.reader(io.StringIO(src)))
    # the source page repeats the header per kernel; take the first kernel
    hh = rows2[1]; data = []
    for r in rows2[2:]:
        if len(r) != len(hh): break
        data.append(r)
    isamp = hh.index('# Samples')
    ts = sum(int(r[isamp] or 0) for r in data)
    f.write(f"\nwarp-state samples of the first captured launch ({ts} samples):\n")
    for name in ['stall_barrier', 'stall_wait', 'stall_long_sb', 'stall_short_sb', 'stall_selected', 'stall_not_selected', 'stall_branch_resolving',
                 'stall_math', 'stall_no_inst', 'stall_lg', 'stall_mio', 'stall_dispatch', 'stall_membar']:
        if name in hh:
            v = sum(int(r[hh.index(name)] or 0) for r in data)
            f.write(f"   {name:26s} {100.0*v/max(1,ts):5.1f}%\n")
json.dump({"kernel": "icp_persist_kernel", "dram_bytes_per_launch": sum(dram) / len(dram), "launches": len(dram),
           "note": "ncu cold-cache replay; per registration of 30 iterations = both cooperative launches (search-only + general instance)"}, open(os.path.join(out_dir, "traffic.json"), "w"))
print("written", tag)
