#!/bin/bash
cat > /tmp/one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from movfem_b200 import mesh, host
m = mesh.build_model("san_fused", 14, 9, 8, 1000., 1100., 900., 2, 2, 1, dirichlet=0, gpml_sch=1, freqs=(0.5, 3.0), sigma_fn=mesh._layered((1500., 1500., 500., 1500.)), topo_amp=50.0)
asm = host.Assembly(m); r = asm.global_vfem(1, m.omega(1), m.sigma_for(1)); print("ok", r[4]); asm.close()
PY
for g in 0 2; do
  if [ $g = 0 ]; then unset MOVFEM_TEST_FUSED_GRID; else export MOVFEM_TEST_FUSED_GRID=$g; fi
  echo "== grid hook $g"
  timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 200 python /tmp/one.py > gpurun_out/r02_race_g$g.log 2>&1
  grep -E "RACECHECK SUMMARY|^ok" gpurun_out/r02_race_g$g.log
  grep -E "hazard detected|Race reported" gpurun_out/r02_race_g$g.log | sed 's/at 0x[0-9a-f]* //; s/+0x[0-9a-f]*//' | cut -c1-150 | sort | uniq -c | sort -rn | head -12
done
