#!/bin/bash
# developer tool (run under gpurun): GPU test suite on the in-tree library, then A/B bench lines of experiment builds
# usage: tools/gpu_ab.sh <tag> "<variants for c2>" ["<variants for c4>"] ["<variants for c3>"]
tag=${1:-ab}
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
  tail -6 gpurun_out/${tag}_pytest.log
fi
BENCH_ARGS="--workload c2" tools/exp_run.sh $2; cp gpurun_out/exp.log gpurun_out/${tag}_exp_c2.log
[ -n "$3" ] && { BENCH_ARGS="--workload c4" tools/exp_run.sh $3; cp gpurun_out/exp.log gpurun_out/${tag}_exp_c4.log; }
[ -n "$4" ] && { BENCH_ARGS="--workload c3" tools/exp_run.sh $4; cp gpurun_out/exp.log gpurun_out/${tag}_exp_c3.log; }
true
