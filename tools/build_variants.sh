#!/bin/bash
# Development aid: builds variants of libstb_b200.so that differ in -D flags of the fill kernel,
# into libstb_b200/lib/exp/libstb_b200_<name>.so (select one with STB_B200_LIB=...).
# usage: tools/build_variants.sh name1="-DFLAG1 -DFLAG2" name2="..."
set -e
cd "$(dirname "$0")/../libstb_b200/csrc"
make -j8 >/dev/null
mkdir -p ../lib/exp /tmp/stb_variants
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  (
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xcompiler -fPIC -I../../include -I. $flags \
      -c stb_cuda.cu -o /tmp/stb_variants/stb_cuda_$name.o
    objs=$(ls ../lib/obj/*.o | grep -v "stb_cuda.cu.o")
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/exp/libstb_b200_$name.so $objs /tmp/stb_variants/stb_cuda_$name.o -lpthread -lm
    echo "built $name ($flags)"
  ) &
done
wait
