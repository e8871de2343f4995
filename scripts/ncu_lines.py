"""Summarises an ncu report: headline metrics + instructions / stall samples per source line."""
import csv, subprocess, sys, io
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__cycles_elapsed.avg",
        "launch__grid_size", "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for v in rows[2:]:
    print("kernel:", v[hdr.index("Kernel Name")][:80])
    for w in want:
        if w in hdr:
            print("  ", w, v[hdr.index(w)], units[hdr.index(w)])
mix = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur, h, out = None, None, []
for r in csv.reader(io.StringIO(mix)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        h = r
        continue
    if h and len(r) > 8 and r[2] == "-":
        ix = {n: i for i, n in enumerate(h)}
        try:
            inst, samp = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        except Exception:
            continue
        stalls = sorted(((int(r[i] or 0), n) for n, i in ix.items() if n.startswith("stall_") and "Not Issued" not in n), reverse=True)[:2]
        out.append((cur, int(r[0]), inst, samp, r[1][:70], stalls))
ti, ts = sum(o[2] for o in out), sum(o[3] for o in out)
print("total inst", ti, "samples", ts)
print("--- top by samples")
for o in sorted(out, key=lambda o: -o[3])[:topn]:
    print(o[0][:20].ljust(20), str(o[1]).rjust(4), f"{100*o[2]/ti:5.1f}%i {100*o[3]/ts:5.1f}%s", o[4].ljust(70), [(a, b[6:]) for a, b in o[5] if a])
print("--- top by instructions")
for o in sorted(out, key=lambda o: -o[2])[:topn]:
    print(o[0][:20].ljust(20), str(o[1]).rjust(4), f"{100*o[2]/ti:5.1f}%i {100*o[3]/ts:5.1f}%s", o[4])
