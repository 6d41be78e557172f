for v in ${VARIANTS:-k64 k128 base}; do
  if [ $v = base ]; then unset MOVFEM_B200_LIB; else export MOVFEM_B200_LIB=$PWD/ab/lib_$v.so; fi
  timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python -c "
import json; b=json.load(open('gpurun_out/ab_$v.json')); print('$v', round(b['ms_per_step'],4), {k: round(x,4) for k,x in b['phases_ms'].items()})"
done
