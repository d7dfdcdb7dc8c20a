"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-kernel lines kept under profiles/, and derive
profiles/traffic.json (DRAM bytes per launch / per step) from the same capture.
    python profiles/summarize_ncu.py raw.csv "header line" out.txt [traffic.json]"""
import csv
import json
import sys

raw, header, out_path = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
idx = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("_not_issued")]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = [header, ""]
traffic = {}
for r in rows[2:]:
    for w in want:
        if w in idx:
            out.append(f"{w} [{units[idx[w]]}] = {r[idx[w]]}")
    tot = sum(float(r[idx[h]] or 0) for h in stall) or 1.0
    top = sorted(((float(r[idx[h]] or 0) / tot * 100, h.replace("smsp__pcsamp_warps_issue_stalled_", ""))
                  for h in stall), reverse=True)[:10]
    out.append("stall samples: " + ", ".join(f"{n} {p:.1f}%" for p, n in top))
    out.append("")
    b = sum(float(r[idx[m]]) * SCALE[units[idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    traffic[name] = traffic.get(name, 0.0) + b
open(out_path, "w").write("\n".join(out))
if len(sys.argv) > 4:
    k2 = max((k for k in traffic if k.startswith("k2_chamfer<")), key=lambda k: traffic[k])
    json.dump({"kernel": k2 + " (narrow tiles, the dominant kernel)", "dram_bytes_per_launch": traffic[k2],
               "source": out_path + " (dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full capture of "
                                    "one step of the pipelined configuration, k3_sky on)",
               "note": "traffic above the algorithmic bytes = forward-state scratch of the scan (written once, read "
                       "back once) and the bit rows / prefixes / depth_list the first stage leaves for it",
               "dram_bytes_per_step": sum(traffic.values()), "per_kernel_dram_bytes": traffic,
               "algorithmic_bytes_per_step": 13 * 256 * 352 * 1216}, open(sys.argv[4], "w"), indent=1)
print("\n".join(out[:3]), "...", json.dumps(traffic))
