# builds libstitchb200 variants with different corr_umma_kernel pipeline shapes into tools/probes/
set -e
P=$(ls -d seamless*_b200)
python $P/build.py > /dev/null
mkdir -p tools/probes
for v in "2 2" "1 2" "1 4" "1 6"; do
  set -- $v
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC \
    -DSB_CORR_BSTAGES=$1 -DSB_CORR_SBUFS=$2 -c $P/csrc/corr_tcgen05.cu -o /tmp/corr_v_$1_$2.o
  objs=$(ls $P/build/*.o | grep -v corr_tcgen05.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/probes/libstitch_b$1_s$2.so $objs /tmp/corr_v_$1_$2.o
  echo built tools/probes/libstitch_b$1_s$2.so
done
