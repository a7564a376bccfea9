"""Turns the ncu artefacts of a gpurun call into the committed summaries under profiles/.
usage: make_profile_summary.py <tag> <launches.csv> <full.ncu-rep>"""
import csv, io, json, os, subprocess, sys, collections
tag, launches, rep = sys.argv[1:4]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = collections.OrderedDict()
seq = []
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("void ", "")
    t = float(r[vi]) / 1000.0
    a = tot.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
    seq.append((name, t))
total = sum(v[1] for v in tot.values())
with open(os.path.join(out_dir, f"{tag}_launches_summary.md"), "w") as f:
    f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 3 --pool 2`\n\n")
    f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes. The whole run is listed (priming + warm-up + timed steps + e2e leg).\n\n")
    f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {v[0]} | {v[1]:.0f} | {100*v[1]/total:.1f}% | {v[1]/v[0]:.1f} |\n")
    # one registration step: the launches between two bbox_init of the fine grid (first 30 icp_iter after a build)
    idx = [i for i, (n, _) in enumerate(seq) if n.startswith("icp_iter_kernel")]
    if idx:
        # find a run of 30 consecutive iteration kernels
        for s in range(len(idx) - 29):
            if idx[s + 29] - idx[s] == 29:
                its = [seq[i][1] for i in idx[s:s + 30]]
                f.write("\nOne registration (30 consecutive `icp_iter_kernel` launches), us per launch:\n\n`" + " ".join(f"{t:.0f}" for t in its) + "`\n")
                f.write(f"\nsum {sum(its):.0f} us; first two iterations {its[0]+its[1]:.0f} us ({100*(its[0]+its[1])/sum(its):.0f}%), last twenty {sum(its[10:]):.0f} us.\n")
                break

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h = rr[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__cycles_active.avg"]
def tobytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
dram = []
with open(os.path.join(out_dir, f"{tag}_icp_iter_full_summary.txt"), "w") as f:
    f.write(f"ncu --set full --clock-control none --import-source on -k regex:icp_iter (30 consecutive launches inside `python bench.py --steps 2 --warmup 3 --pool 2`)\n")
    f.write("columns: " + " | ".join(want) + "\n")
    for r in rr[2:]:
        vals = []
        for w in want:
            vals.append(r[h.index(w)] if w in h else "?")
        f.write(" | ".join(vals) + "\n")
        i1, i2 = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        dram.append(tobytes(r[i1], rr[1][i1]) + tobytes(r[i2], rr[1][i2]))
    f.write("units: " + " | ".join(rr[1][h.index(w)] if w in h else "?" for w in want) + "\n")
    f.write(f"\nDRAM bytes (read+write) per launch: mean {sum(dram)/len(dram):.0f}, max {max(dram):.0f}; algorithmic bytes per iteration 14745600.\n")
    f.write("ncu flushes the caches before every replay (cold cache): with warm caches the working set (~45 MB) is L2 resident and DRAM traffic of an iteration is ~0.\n")
json.dump({"dram_bytes_per_launch": sum(dram) / len(dram), "source": f"profiles/{tag}_icp_iter_full_summary.txt (ncu --set full, cold cache, mean of {len(dram)} launches)"},
          open(os.path.join(out_dir, "traffic.json"), "w"), indent=1)
print("ok", len(seq), "launches;", len(dram), "full captures")
