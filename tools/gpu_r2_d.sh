#!/bin/bash
# round 2, run D: first run of the PatchEmbed kernel (N4) + lookup DRAM comparison
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "patch_embed" 2>&1 | tail -30 > gpurun_out/r2d_tests.log
cat gpurun_out/r2d_tests.log
python -c "
import sys; sys.path.insert(0,'.')
import stitch_b200
print('debug word', hex(stitch_b200._lib.load().sb_debug_word()))"
timeout 300 python tools/kernel_bench.py 2>&1 | grep -E "patch_embed|tensor|torch/cuDNN" | tail -4
python tools/lookup_generic_prof.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum --clock-control none -k regex:"corr_lookup" --csv --log-file gpurun_out/r2d_lookup_ncu.csv python tools/lookup_generic_prof.py > gpurun_out/r2d_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2d_lookup_ncu.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:40]),{})[r[mi]]=r[vi]
for k,v in d.items(): print(k, v)
PY
