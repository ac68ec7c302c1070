for cfg in "-DRANGE_PC_BATCH=2" "-DRANGE_PC_BATCH=8 -DRANGE_PC_RING=32" "-DRANGE_PC_WINDOW=32" "-DRANGE_PC_WINDOW=128" "-DRANGE_PC_RING=32" ""; do
  NVCC_EXTRA="$cfg" bash range_b200/csrc/build.sh > /dev/null 2>&1
  echo "== [$cfg]"; for i in 1 2; do timeout 200 python tools/time_apply.py 2>&1 | grep -v "^$" | tail -1; done
done
