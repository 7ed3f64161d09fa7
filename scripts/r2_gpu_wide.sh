#!/bin/bash
# round-2 session script: layered-path tests + timings (+ ncu launch list of one layered step)
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_wide.py -x -q -m gpu 2>&1 | tail -3
for cfg in "dgru 64 64 2048 1" "dgru 40 64 2048 2" "gru 48 64 2048 1" "lstm 64 64 2048 1" "dgru 13 64 2048 2"; do
  set -- $cfg
  KT_LAYERS=$5 timeout 300 python scripts/ktime.py $1 $2 $3 $4 2>&1 | tail -1 | cut -c1-330
done
KT_LAYERS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_wide_launches.csv python scripts/ktime.py dgru 64 64 2048 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('/root/repo/gpurun_out/r2_wide_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); 
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    k=r[ki][:70]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in agg.items(): print(f"{n:3d} x {t/n/1000:9.1f} us  {k}")
PY
