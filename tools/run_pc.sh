for poly in 0 4 0 4 3 0; do
  NVCC_EXTRA="-DRANGE_PC_POLY=$poly" bash range_b200/csrc/build.sh > /dev/null 2>&1
  echo "== poly $poly"; timeout 200 python tools/time_apply.py 2>&1 | grep -v "^$" | tail -1
done
