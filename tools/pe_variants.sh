# Builds libstitchb200 variants of the PatchEmbed kernel (csrc/patch_embed.cu) into tools/probes/:
#   usage: bash tools/pe_variants.sh "name1:-DFLAG=.. -DFLAG2=.." "name2:..."
set -e
P=$(ls -d seamless*_b200)
python $P/build.py > /dev/null
mkdir -p tools/probes
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC \
    $flags -c $P/csrc/patch_embed.cu -o /tmp/pe_v_$name.o
  objs=$(ls $P/build/*.o | grep -v patch_embed.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/probes/libstitch_pe_$name.so $objs /tmp/pe_v_$name.o
  echo built tools/probes/libstitch_pe_$name.so
done
