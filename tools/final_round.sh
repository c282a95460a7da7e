#!/bin/bash
# developer tool (run under gpurun): the round's closing measurements -- GPU tests, the three bench lines (C2 with the
# sync-patched reference kernels timed beside it), the reference arm, a launch list and one ncu --set full capture per workload.
tag=${1:-r2f}
REFK=1 tools/gpu_round.sh $tag
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err; echo "reference arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 100 --csv --log-file gpurun_out/${tag}_launches_c2.csv python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-ref-kernels > gpurun_out/${tag}_launches.log 2>&1; echo "launch list rc=$?"
for wl in c2 c3 c4; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"triangleSetup|fineRaster|directAlloc|directScatter" -s 40 -c 4 -o gpurun_out/${tag}_${wl}_full python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline --no-ref-kernels > gpurun_out/${tag}_ncu_${wl}.log 2>&1; echo "ncu $wl rc=$?"
done
ls -la gpurun_out/${tag}_*
