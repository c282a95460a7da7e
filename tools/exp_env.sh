# developer tool: runs bench.py under different environment settings ("NAME=VALUE" or "-"), one line per run
out=gpurun_out/exp.log; rm -f $out
for e in "$@"; do
  ( [ "$e" != "-" ] && export $e; python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-ref-kernels ${BENCH_ARGS} 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$e', round(d['value']), {k:round(v*1000,1) for k,v in d['stage_ms'].items()}, 'e2e', round(d['e2e']['value']))" >> $out )
done
cat $out
