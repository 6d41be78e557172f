#!/bin/bash
# config-2 timing A/B: tools/r02_ab3.sh lib1.so ...
for lib in "$@"; do
  n=$(basename $lib .so)
  MOVFEM_B200_LIB=$PWD/$lib timeout 200 python tools/run_one.py 2 None 6 2>/dev/null | python -c "
import sys,ast
d=ast.literal_eval(sys.stdin.read().strip().splitlines()[-1]); print('$n', {k: round(d[k],4) for k in ('ms_total','ms_element','ms_exact','ms_gather','nz')})"
done
