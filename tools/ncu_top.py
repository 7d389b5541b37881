"""Summarise an .ncu-rep: duration / DRAM bytes, stall reasons, hottest SASS lines with their source line."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
for i, h in enumerate(hdr):
    if h in want:
        print(f"  {h} = {vals[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hs = [i for i, r in enumerate(rows) if r and r[0] == "Address"]   # one header row per profiled launch: the first one
h = hs[0]
hdr = rows[h]; data = [r for r in rows[h + 1:(hs[1] if len(hs) > 1 else len(rows))] if len(r) == len(hdr)]
ci = {x: i for i, x in enumerate(hdr)}
stalls = [x for x in hdr if x.startswith("stall_") and "Not Issued" not in x]
tot = sum(int(r[ci["# Samples"]]) for r in data)
agg = sorted(((s, sum(int(r[ci[s]]) for r in data)) for s in stalls), key=lambda x: -x[1])[:8]
print("samples", tot, [(s, f"{100*v/tot:.1f}%") for s, v in agg])
for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:topn]:
    st = sorted(((s, int(r[ci[s]])) for s in stalls), key=lambda x: -x[1])[:2]
    print(f"{100*int(r[ci['# Samples']])/tot:6.2f}%  {r[ci['Source']].strip()[:64]:64s} {st}")
