#!/bin/bash
# Builds dextractor_b200/libdexb200.so for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
OUT=../libdexb200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -I../../include -I."
mkdir -p build
objs=""
pids=""
HDRS="dx_common.cuh dx_bits.cuh dx_internal.h ../../include/dexb200.h"
newer() { for h in $HDRS $1; do [ $h -nt $2 ] && return 0; done; return 1; }
for f in dx_frame.cu dx_qv_stats.cu dx_qv_encode.cu dx_qv_decode.cu dx_qv_decode5.cu dx_qv_decode6.cu dx_qv_plan.cu dx_pack.cu dx_pack2.cu dx_pack3.cu dx_api.cpp dx_coding.cpp dx_pipe.cpp; do
  [ -f $f ] || continue
  o=build/${f%.*}.o
  if [ ! -f $o ] || newer $f $o; then
    rm -f $o                       # a failed compile must not leave a stale object for the link
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -x cu -c $f -o $o &
    pids="$pids $!"
  fi
  objs="$objs $o"
done
for p in $pids; do wait $p || { echo "compile failed" >&2; exit 1; }; done
for o in $objs; do [ -f $o ] || { echo "missing $o" >&2; exit 1; }; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o $OUT $objs -lcudart_static -lpthread -ldl -lrt
echo "built $OUT"
# libdexcompat.so: the reference's QV.h / DB.h symbols over libdexb200.so
if [ ! -f ../libdexcompat.so ] || [ dx_compat.cu -nt ../libdexcompat.so ] || [ ../../include/dexb200.h -nt ../libdexcompat.so ]; then
  $NVCC $FLAGS -shared -o ../libdexcompat.so dx_compat.cu -L.. -ldexb200 -Xlinker -rpath -Xlinker '$ORIGIN' -lcudart_static -lpthread -ldl -lrt
  echo "built ../libdexcompat.so"
fi
# libdexcompat_i.so: the same symbols under the reference's -DINTERACTIVE error convention (DB.h:28-47):
# messages into the exported Ebuffer, error values returned instead of exit()
if [ ! -f ../libdexcompat_i.so ] || [ dx_compat.cu -nt ../libdexcompat_i.so ] || [ ../../include/dexb200.h -nt ../libdexcompat_i.so ]; then
  $NVCC $FLAGS -DINTERACTIVE -shared -o ../libdexcompat_i.so dx_compat.cu -L.. -ldexb200 -Xlinker -rpath -Xlinker '$ORIGIN' -lcudart_static -lpthread -ldl -lrt
  echo "built ../libdexcompat_i.so"
fi
