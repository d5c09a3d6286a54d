"""Export what bench.py and the profile summaries need from an `ncu --set full` report (run here, no GPU needed):

  python tools/ncu_traffic.py gpurun_out/r2_pmpc_full.ncu-rep profiles/r2_pmpc

writes <prefix>_ncu_raw.csv (the raw page, one row per captured launch), <prefix>_ncu_traffic.json (dram bytes per launch:
bench.py's roofline.traffic), <prefix>_lines.txt (stall samples and instructions by source line) and
<prefix>_sass_tensor.txt (the tcgen05 / TMA mnemonics of the shipped library, if any, for the named kernel)."""
import csv, json, os, subprocess, sys

rep, prefix = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(prefix + "_ncu_raw.csv", "w").write(raw)
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v = float(r[idx[name]].replace(",", ""))
    u = units[idx[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


launches = rows[2:]
rd = sum(val(r, "dram__bytes_read.sum") for r in launches) / len(launches)
wr = sum(val(r, "dram__bytes_write.sum") for r in launches) / len(launches)
rec = {"kernel": launches[0][idx["Kernel Name"]], "launches_captured": len(launches), "dram_bytes_read": rd, "dram_bytes_write": wr,
       "duration_us": [float(r[idx["gpu__time_duration.sum"]]) for r in launches],
       "registers_per_thread": int(float(launches[0][idx["launch__registers_per_thread"]])),
       "source": f"ncu --set full --clock-control none ({os.path.basename(rep)}), mean over {len(launches)} launches; tools/ncu_traffic.py"}
for k in ("sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"):
    if k in idx:
        rec[k] = float(launches[0][idx[k]].replace(",", ""))
st = {h: float(launches[0][idx[h]].replace(",", "")) for h in hdr
      if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")}
tot = sum(st.values()) or 1.0
rec["stall_share_pct"] = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(100 * v / tot, 1)
                          for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]}
json.dump(rec, open(prefix + "_ncu_traffic.json", "w"), indent=1)
here = os.path.dirname(os.path.abspath(__file__))
lines = subprocess.run([sys.executable, os.path.join(here, "ncu_lines.py"), rep, "40"], capture_output=True, text=True).stdout
open(prefix + "_lines.txt", "w").write(lines)
print(json.dumps(rec, indent=1))
