#!/bin/bash
# Build the library with other tuning macros of the producer/consumer apply kernel (csrc/retrieval_pc.cu) into
# build/variants/ for A/B timing on the GPU box:   tools/variants.sh            (here, no GPU)
#                                                  tools/variants.sh time       (on the box: tools/time_apply.py per variant)
set -e
cd "$(dirname "$0")/.."
SRC=${SRC:-retrieval_pc}
VARIANTS=("base:" "batch2:-DRANGE_PC_BATCH=2" "batch8:-DRANGE_PC_BATCH=8" "ring32:-DRANGE_PC_RING=32" "acc256:-DRANGE_PC_ACC_WINDOW=256" \
          "win128:-DRANGE_PC_WINDOW=128" "win32:-DRANGE_PC_WINDOW=32")
if [ "${1:-build}" = "build" ]; then
  mkdir -p build/variants
  FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --use_fast_math -diag-suppress 177"
  for v in "${VARIANTS[@]}"; do
    tag=${v%%:*}; def=${v#*:}
    nvcc $FLAGS $def -c range_b200/csrc/$SRC.cu -o build/variants/${SRC}_$tag.o
    objs=""
    for f in encoder encoder_tc encoder_raster retrieval retrieval_pc sort merge capi host; do
      if [ $f = $SRC ]; then objs="$objs build/variants/${SRC}_$tag.o"; else objs="$objs build/$f.o"; fi
    done
    nvcc -arch=sm_100a -shared -cudart static -o build/variants/librange_b200_$tag.so $objs -Xlinker -lpthread
    echo "built $tag ($def)"
  done
else
  for rep in 1 2; do
    for v in "${VARIANTS[@]}"; do
      tag=${v%%:*}
      RANGE_B200_LIB=$PWD/build/variants/librange_b200_$tag.so REPS=${REPS:-8} python tools/time_apply.py 2>/dev/null | tail -1 | sed "s/^/$tag: /"
    done
  done
fi
