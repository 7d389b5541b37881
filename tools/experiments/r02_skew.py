import ctypes, json, os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from htlr_b200 import _lib
out = {}
for wl in sys.argv[1:]:
    r = bench.measure_workload(torch, wl, 16, 5, 0, 0, None, want_profile=False)
    lib = ctypes.CDLL(os.environ["HTLR_CUDA_LIB"])
    buf = (ctypes.c_ulonglong * 4096)()
    lib.htlr_dbg_read(buf)
    a = np.array(buf[:4000], dtype=np.int64).reshape(-1, 160)[:9, :148].astype(np.float64)
    t0 = a[0].min()
    smid = a[6].copy(); d10 = (a[8] - a[7]) / 1e3
    a = (a - t0) / 1e3   # us
    a[6] = smid; a[7] = d10
    d = {"it_s": round(16 / (r["ms"] * 1e-3), 1)}
    names = ["step_start", "sweep_end", "published", "after_barrier1", "rowphase_end", "after_barrier2"]
    for k, nm in enumerate(names):
        d[nm] = {"min": round(a[k].min(), 2), "p50": round(float(np.median(a[k])), 2), "max": round(a[k].max(), 2)}
    d["sweep_dur"] = {"min": round((a[1] - a[0]).min(), 2), "p50": round(float(np.median(a[1] - a[0])), 2), "max": round((a[1] - a[0]).max(), 2)}
    order = np.argsort(a[1])
    d["slowest_blocks"] = [(int(i), round(a[1][i], 2)) for i in order[-8:]]
    d["fastest_blocks"] = [(int(i), round(a[1][i], 2)) for i in order[:8]]
    d["sweep_end_by_block_decile"] = [round(float(x), 2) for x in np.percentile(a[1], [0, 10, 25, 50, 75, 90, 100])]
    out[wl] = d
    np.save(f"/root/repo/gpurun_out/skew_{wl}.npy", a)
print(json.dumps(out))
