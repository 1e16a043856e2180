#!/bin/bash
# Build an experimental variant of libhv_swin.so: recompile the named sources with extra nvcc flags, link with the
# production objects.   tools/build_variant.sh NAME "FLAGS" file1.cu [file2.cu ...]  ->  hierarchical_vision_b200/libhv_swin_NAME.so
# (select it at run time with HV_SWIN_LIB=<path>; the variants are git-ignored)
set -e
name=$1; flags=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
pkg=$root/hierarchical_vision_b200
out=/tmp/hv_variant_$name
mkdir -p $out
python -m hierarchical_vision_b200.build > /dev/null
objs=""
for o in $pkg/_build/*.o; do
  b=$(basename $o .o)
  skip=0
  for f in "$@"; do [ "$(basename $f .cu)" = "$b" ] && skip=1; done
  [ $skip = 0 ] && objs="$objs $o"
done
for f in "$@"; do
  b=$(basename $f .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
    --expt-relaxed-constexpr $flags -c $pkg/csrc/$b.cu -o $out/$b.o &
done
wait
for f in "$@"; do objs="$objs $out/$(basename $f .cu).o"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $pkg/libhv_swin_$name.so $objs -Xcompiler -fPIC
echo $pkg/libhv_swin_$name.so
