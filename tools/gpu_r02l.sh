#!/bin/bash
mkdir -p gpurun_out
echo "== imputed timing 512 genes, flush trace"; IMP_GENES=512 RVT_FLUSH_TRACE=1 timeout 900 python tools/imputed_time.py > gpurun_out/r02l_imputed.log 2>&1; echo "rc=$?"; grep -v "^\[flush\] enter" gpurun_out/r02l_imputed.log | tail -60
echo "== launch list of an augmented flush"; IMP_GENES=256 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_sweep_aug|k_aug_stats|k_sweep_tc|k_finalize|k_tile' --csv --log-file gpurun_out/r02l_launches_aug.csv python tools/imputed_time.py > /dev/null 2>&1; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02l_launches_aug.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:]: print(r[ki].split('(')[0][:60], float(r[vi].replace(',',''))/1e6,'ms')
PY
