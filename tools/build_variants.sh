# Builds variants of ONE kernel file of libstitchb200 into tools/probes/ (all other objects from the regular build):
#   usage: bash tools/build_variants.sh corr_tcgen05 "name1:-DFLAG=.. -DFLAG2=.." "name2:..."
#   -> tools/probes/libstitch_<file>_<name>.so, loaded with STITCH_B200_LIB=...
set -e
P=$(ls -d seamless*_b200)
F=$1; shift
python $P/build.py > /dev/null
mkdir -p tools/probes
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -I include \
    $flags -c $P/csrc/$F.cu -o /tmp/var_${F}_$name.o
  objs=$(ls $P/build/*.o | grep -v "/$F.o")
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/probes/libstitch_${F}_$name.so $objs /tmp/var_${F}_$name.o
  echo built tools/probes/libstitch_${F}_$name.so
done
