for v in base pf1 pf2 base pf1 pf2; do
  lib=build/variants/$v/libcrb200.so; [ $v = base ] && lib=cudaraster-linux_b200/libcrb200.so
  CRB200_LIBRARY=$lib python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-ref-kernels 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value']), {k:round(v*1000,1) for k,v in d['stage_ms'].items()})" >> gpurun_out/exp1.log
done
cat gpurun_out/exp1.log
