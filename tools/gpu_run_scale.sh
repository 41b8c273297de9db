# 1 -> N GPU scaling of bench.py on one box (torchrun, one rank per GPU)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
NG=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  [ $n -le $NG ] || continue
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n$n.log 2>&1
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/scale_n$n.log 2>&1
  fi
  echo "n=$n exit $?"
  grep '^{' gpurun_out/scale_n$n.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  value', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
