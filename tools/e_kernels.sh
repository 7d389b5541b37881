#!/bin/bash
python bench.py --workload E --n-total ${NTOT:-25000} --steps 2 --warmup 3 > gpurun_out/plain_E25k.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -s 400 -c 16 --csv --log-file gpurun_out/launches_E25k.csv python bench.py --workload E --n-total ${NTOT:-25000} --steps 2 --warmup 3 > gpurun_out/ncu_E25k.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_E25k.csv")) if len(r)>10 and r[0].isdigit()]
seen={}
for r in rows:
    seen.setdefault(r[0],{"name":r[4][:50],"grid":r[8]})[r[12]]=r[-1]
for k,v in list(seen.items())[:8]:
    t=float(v.get("gpu__time_duration.sum",0))/1e3; by=float(v.get("dram__bytes_read.sum",0))
    print("%-50s grid=%-16s %9.1f us  dram_read=%8.1f MB  %7.1f GB/s"%(v["name"],v["grid"],t,by/1e6, by/1e3/max(t,1e-9)))
PY
