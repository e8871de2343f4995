"""profiles/traffic.json from an `ncu --set full` capture of the three bench launches (1 %, 10 %, 50 % density).

usage: python scripts/ncu_traffic.py gpurun_out/prof_bench_r1.ncu-rep FRAMES_PER_LAUNCH "capture command" > profiles/traffic.json
"""
import csv, io, json, subprocess, sys

rep, frames, source = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]


def col(row, name, scale_unit=True):
    i = hdr.index(name)
    v = float(row[i].replace(",", ""))
    u = units[i]
    if scale_unit:
        v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
    return v


per = []
for row, d in zip(rows[2:5], (10000, 100000, 500000)):
    rd, wr = col(row, "dram__bytes_read.sum") / frames, col(row, "dram__bytes_write.sum") / frames
    per.append({"density_ppm": d, "dram_read_bytes_per_frame": rd, "dram_write_bytes_per_frame": wr,
                "dram_bytes_per_frame": rd + wr,
                "duration_under_ncu": row[hdr.index("gpu__time_duration.sum")] + " " + units[hdr.index("gpu__time_duration.sum")],
                "inst_executed": col(row, "smsp__inst_executed.sum", False),
                "issue_active_pct": col(row, "smsp__issue_active.avg.pct_of_peak_sustained_active", False),
                "registers_per_thread": int(col(row, "launch__registers_per_thread", False)),
                "kernel": row[hdr.index("Kernel Name")]})
print(json.dumps({"source": source, "frames_per_captured_launch": frames, "per_density": per,
                  "dram_bytes_per_frame_mean": sum(p["dram_bytes_per_frame"] for p in per) / len(per)}, indent=1))
