#!/bin/bash
mkdir -p gpurun_out
echo "== aug + bed + parity tests"; timeout 900 python -m pytest tests/test_gpu_aug.py tests/test_bed_format.py tests/test_gpu_parity.py tests/test_gpu_zz_fp64_skato.py -m gpu -q > gpurun_out/r02m_tests.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r02m_tests.log
echo "== headline sweep (skip check added)"; timeout 300 python tools/overlap_time.py 2500 quick 2>&1 | grep "skato"
echo "== imputed timing 512 genes"; IMP_GENES=512 timeout 900 python tools/imputed_time.py > gpurun_out/r02m_imputed.log 2>&1; echo "rc=$?"; cat gpurun_out/r02m_imputed.log
echo "== launch list"; IMP_GENES=256 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_sweep_aug|k_aug_stats|k_sweep_tc' -c 12 --csv --log-file gpurun_out/r02m_launches_aug.csv python tools/imputed_time.py > /dev/null 2>&1; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02m_launches_aug.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:]: print(r[ki].split('(')[0][:60], float(r[vi].replace(',',''))/1e6,'ms')
PY
