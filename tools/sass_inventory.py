"""SASS inventory of libhtlr_cuda.so: per kernel, registers / stack and the counts of the mnemonics that prove the design points
DESIGN.md names. python tools/sass_inventory.py > profiles/rNN_sass_k_traj.txt (needs cuobjdump; no GPU)."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "htlr_b200", "libhtlr_cuda.so")
KEYS = ["UBLKCP", "UBLKPF", "SYNCS", "UCGABAR", "CGAERRBAR", "STAS", "LDS.128", "LDS.64", "LDS", "DFMA", "DADD", "DMUL", "SHFL", "BAR.SYNC",
        "MEMBAR", "RED", "ATOM", "LDGSTS", "LDG", "STG", "DMMA", "HMMA", "UTCMMA", "STL", "LDL", "MUFU"]
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", res):
    usage[m.group(1)] = (int(m.group(2)), int(m.group(3)))
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, cur = {}, None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k in ("LDS", "LDG", "STG", "RED", "ATOM", "MEMBAR", "SYNCS", "UCGABAR") and op.startswith(k)):
                counts[cur][k] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("SASS inventory of the kernels in htlr_b200/libhtlr_cuda.so (sm_100a), tools/sass_inventory.py.")
print("cuobjdump -sass | per function: counts of the mnemonics that prove the design points DESIGN.md names:")
print("  UBLKCP = cp.async.bulk (1-D TMA bulk copy) | SYNCS = mbarrier ops | STAS = st.async (counted stores into another block's shared")
print("  memory) | UCGABAR = cluster barrier (CGAERRBAR + MEMBAR.ALL.GPU come with its release) | LDS.128 = 128-bit shared loads of the column")
print("  tile | DFMA | SHFL | BAR = named barriers | MEMBAR/RED/ATOM = grid barrier | LDGSTS = cp.async | DMMA = FP64 tensor-core MMA")
print("  (prediction only) | STL/LDL = spill traffic. No HMMA/UTCMMA anywhere: the path is FP64, skinny, HBM-bound.\n")
for mang, name in sorted(zip(counts, dem), key=lambda t: t[1]):
    short = re.sub(r"htlr::|\(anonymous namespace\)::|\(.*$", "", name)
    r, st = usage.get(mang, (0, 0))
    body = " ".join(f"{k}={counts[mang][k]}" for k in KEYS if counts[mang][k])
    print(f"{short:42s} regs={r:3d} stack={st:3d}  {body}")
