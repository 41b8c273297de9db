#!/bin/bash
# round 2, run C: warp kernels after the instruction diet; DRAM bytes of the two lookup access methods
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "warp or homo or adapter or config1 or hot_path or step or composite or smoke" 2>&1 | tail -8 > gpurun_out/r2c_tests.log
python tools/kernel_bench.py 2>&1 | grep -E "flow_warp|homo_warp|range_map|lookup|morph" > gpurun_out/r2c_kernel_bench.txt
python tools/lookup_generic_prof.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum --clock-control none -k regex:"corr_lookup" --csv --log-file gpurun_out/r2c_lookup_ncu.csv python tools/lookup_generic_prof.py > gpurun_out/r2c_ncu.log 2>&1
cat gpurun_out/r2c_tests.log gpurun_out/r2c_kernel_bench.txt
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c_lookup_ncu.csv')) if len(r)>10]
hdr=rows[0]; 
ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
import collections
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:40]),{})[r[mi]]=r[vi]
for k,v in d.items(): print(k, v)
PY
