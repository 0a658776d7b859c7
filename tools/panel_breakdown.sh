#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python tools/one_step.py cyclic9-15 15000000 9 > gpurun_out/pb_plain.log 2>&1; tail -1 gpurun_out/pb_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/pb_c9.csv python tools/one_step.py cyclic9-15 15000000 9 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/pb_k14.csv python tools/one_step.py kat14-15 100000000 2 panel_max_top=1000000 > /dev/null 2>&1
python - <<'PY'
import csv, collections
for f in ['gpurun_out/pb_c9.csv','gpurun_out/pb_k14.csv']:
    rows=[r for r in csv.reader(open(f)) if len(r)>5]
    hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
    tot=collections.Counter(); cnt=collections.Counter()
    for r in rows[1:]:
        try: v=float(r[vi].replace(',',''))
        except: continue
        n=r[ki].split('(')[0][:40]; tot[n]+=v; cnt[n]+=1
    print(f)
    for n,v in tot.most_common(8): print(f'  {n:42s} {cnt[n]:5d} launches {v/1e6:9.2f} ms')
PY
