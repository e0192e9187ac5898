"""profiles/ from an ncu report: python scripts/make_profile_summary.py gpurun_out/prof_X.ncu-rep TAG [paths time_steps]
writes profiles/TAG_ncu_full.csv (selected raw metrics per kernel), profiles/TAG_ncu_summary.txt (scripts/ncu_summary.py
output incl. the top stalled SASS instructions) and profiles/ncu_traffic.json (DRAM bytes per launch, read by bench.py)."""
import csv, io, json, os, subprocess, sys, contextlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import ncu_summary  # noqa: E402

rep, tag = sys.argv[1], sys.argv[2]
paths = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 100
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
open("/tmp/_raw.csv", "w").write(raw)
open("/tmp/_src.csv", "w").write(src)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keep = ["Kernel Name", "Block Size", "Grid Size"] + [k for k in ncu_summary.WANT if k in hdr] + \
       [k for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")] + \
       [k for k in hdr if k in ("launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_warps", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                                "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active")]
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", tag + "_ncu_full.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(keep)
    w.writerow([units[hdr.index(k)] for k in keep])
    for r in rows[2:]:
        w.writerow([r[hdr.index(k)] for k in keep])
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    ncu_summary.raw("/tmp/_raw.csv")
    ncu_summary.source("/tmp/_src.csv", 25)
open(os.path.join(ROOT, "profiles", tag + "_ncu_summary.txt"), "w").write(buf.getvalue())
traffic = {"paths": paths, "time_steps": steps, "source": tag + "_ncu_full.csv (ncu --set full --clock-control none, one launch each)", "kernels": {}}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"]
    short = "sim_merton_kernel" if "sim_merton" in name else "reg_forward_tc" if "reg_forward" in name else \
            "reg_backward_tc" if "reg_backward" in name else name.split("(")[0].split("<")[0].split("::")[-1]
    rd = float(d["dram__bytes_read.sum"]) * scale[units[hdr.index("dram__bytes_read.sum")]]
    wr = float(d["dram__bytes_write.sum"]) * scale[units[hdr.index("dram__bytes_write.sum")]]
    traffic["kernels"][short] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "duration_us_under_ncu": float(d["gpu__time_duration.sum"])}
json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
