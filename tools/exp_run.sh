# developer tool: runs bench.py on experiment builds (tools/build_variant.sh) and prints one line per run
# usage: BENCH_ARGS="--workload c2" tools/exp_run.sh base variant1 variant2 ...
out=gpurun_out/exp.log; rm -f $out
for v in "$@"; do
  lib=build/variants/$v/libcrb200.so; [ $v = base ] && lib=cudaraster-linux_b200/libcrb200.so
  CRB200_LIBRARY=$lib python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-ref-kernels ${BENCH_ARGS} 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', '${BENCH_ARGS}', round(d['value']), 'chain', round(d['value_unbroken_chain']), 'two', round(d['value_two_in_flight'] or 0), {k:round(v*1000,1) for k,v in d['stage_ms'].items() if k != 'frames'}, 'e2e', round(d['e2e']['value']))" >> $out
done
cat $out
