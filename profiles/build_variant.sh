#!/bin/bash
# build an experimental variant of the library next to the product one:  build_variant.sh <name> <nvcc flags...>
# (selected at run time with PP_B200_LIB=powerpaint_b200/_variants/lib<name>.so)
name=$1; shift
set -e
cd "$(dirname "$0")/.."
mkdir -p powerpaint_b200/_variants/_obj_$name
for f in powerpaint_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
       -I include -c $f -o powerpaint_b200/_variants/_obj_$name/$(basename $f .cu).o &
done
wait
nvcc -shared -o powerpaint_b200/_variants/lib$name.so powerpaint_b200/_variants/_obj_$name/*.o -lcudart_static -lpthread -ldl -lrt
rm -rf powerpaint_b200/_variants/_obj_$name
ls -la powerpaint_b200/_variants/lib$name.so
