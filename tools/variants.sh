#!/bin/bash
# Build the library with other tuning macros of the producer/consumer apply kernel (csrc/retrieval_pc.cu) into
# build/variants/ for A/B timing on the GPU box:   tools/variants.sh            (here, no GPU)
#                                                  tools/variants.sh time       (on the box: tools/time_apply.py per variant)
set -e
cd "$(dirname "$0")/.."
VARIANTS=("base:" "batch2:-DRANGE_PC_BATCH=2" "batch8:-DRANGE_PC_BATCH=8" "ring32:-DRANGE_PC_RING=32" "acc256:-DRANGE_PC_ACC_WINDOW=256" \
          "win128:-DRANGE_PC_WINDOW=128" "win32:-DRANGE_PC_WINDOW=32")
if [ "${1:-build}" = "build" ]; then
  mkdir -p build/variants
  FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --use_fast_math -diag-suppress 177"
  for v in "${VARIANTS[@]}"; do
    tag=${v%%:*}; def=${v#*:}
    nvcc $FLAGS $def -c range_b200/csrc/retrieval_pc.cu -o build/variants/retrieval_pc_$tag.o
    nvcc -arch=sm_100a -shared -cudart static -o build/variants/librange_b200_$tag.so build/encoder.o build/encoder_tc.o \
      build/encoder_raster.o build/retrieval.o build/variants/retrieval_pc_$tag.o build/sort.o build/merge.o build/capi.o build/host.o -Xlinker -lpthread
    echo "built $tag ($def)"
  done
else
  for rep in 1 2; do
    for v in "${VARIANTS[@]}"; do
      tag=${v%%:*}
      RANGE_B200_LIB=$PWD/build/variants/librange_b200_$tag.so REPS=8 python tools/time_apply.py 2>/dev/null | tail -1 | sed "s/^/$tag: /"
    done
  done
fi
