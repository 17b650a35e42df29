"""ncu -i <rep> --page raw --csv -> profiles/<name>_summary.json (+ _details.csv); one entry per captured kernel launch"""
import csv, json, subprocess, sys, re
rep, name = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(f"profiles/{name}_details.csv", "w").write(raw)
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
outs = []
for vals in rows[2:]:
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    def f(k):
        try: return float(d[k].replace(",", ""))
        except Exception: return None
    out = {"kernel": d.get("Kernel Name"), "report": rep.split("/")[-1]}
    for k in keys:
        if k in d: out[k] = {"value": f(k), "unit": u.get(k)}
    for k in hdr:
        if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", k):
            v = f(k)
            if v and v > 0.02: out.setdefault("stalls_per_issue", {})[k.split("stalled_")[1].split("_per_issue")[0]] = round(v, 3)
    outs.append(out)
json.dump(outs if len(outs) > 1 else outs[0], open(f"profiles/{name}_summary.json", "w"), indent=1)
for o in outs:
    print(o["kernel"], {k.split(".")[0]: v["value"] for k, v in o.items() if isinstance(v, dict) and "value" in v and k in keys[:10]}, o.get("stalls_per_issue"))
