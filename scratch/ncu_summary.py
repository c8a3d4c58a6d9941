"""scratch: summarise an .ncu-rep (raw metrics + top stalled source lines) into a markdown file.
usage: python scratch/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.md [title]"""
import csv, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sectors_srcunit_tex_op_atom.sum",
        "lts__t_sectors_srcunit_tex_op_red.sum", "smsp__cycles_active.avg"]
L = [f"# {title}", "", "`ncu --set full --clock-control none --import-source on` (cold-cache, serialised replays: shares, not absolutes).", ""]
L.append("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
L.append("|---|---|" + "---|" * len(data))
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        L.append(f"| `{k}` | {units[i]} | " + " | ".join(r[i][:90] for r in data) + " |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur, names, lines, tot = None, None, {}, 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        names = r
    elif r[0].isdigit() and names:
        s = int(r[4]) if r[4].isdigit() else 0
        key = (cur, int(r[0]), r[1].strip()[:100])
        st = lines.setdefault(key, [0, {}])
        st[0] += s
        tot += s
        for i, n in enumerate(names):
            if n.startswith("stall_") and "Not Issued" not in n and r[i].isdigit():
                st[1][n[6:]] = st[1].get(n[6:], 0) + int(r[i])
L += ["", f"## Top source lines by warp-stall samples (all captured launches, {tot} samples)", "",
      "| % | file:line | source | top stall reasons |", "|---|---|---|---|"]
for (f, ln, s), (n, st) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:30]:
    top = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    L.append(f"| {100 * n / max(tot, 1):.1f} | {f}:{ln} | `{s.replace('|', '/')}` | {top} |")
open(out, "w").write("\n".join(L) + "\n")
print("\n".join(L[:30]))
