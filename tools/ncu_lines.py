"""Per CUDA source line: share of executed warp instructions and of stall samples (ncu cuda,sass view)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines = []; fname = ""; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        ci = {x: i for i, x in enumerate(hdr)}
        lines.append((fname, int(r[0]), r[1].strip(), int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]])))
ti = sum(l[3] for l in lines); ts = sum(l[4] for l in lines)
print("warp instr", ti, "samples", ts)
for l in sorted(lines, key=lambda l: -l[3])[:topn]:
    print(f"{100*l[3]/ti:5.1f}% instr {100*l[4]/ts:5.1f}% smpl  {l[0]}:{l[1]:<4} {l[2][:92]}")
