#!/bin/bash
# developer tool (run under gpurun): the GPU test suite, then the three bench lines; everything lands in gpurun_out/
tag=${1:-r2a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
for wl in c2 c3 c4; do
  extra="--no-ref-kernels"; [ $wl = c2 ] && [ "${REFK:-0}" = 1 ] && extra=""
  timeout 900 python bench.py --workload $wl $extra > gpurun_out/${tag}_bench_$wl.json 2> gpurun_out/${tag}_bench_$wl.err; echo "bench $wl rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench_$wl.json"))
    print("$wl", "value", round(d["value"]), "chain", round(d["value_unbroken_chain"]), "two", d["value_two_in_flight"] and round(d["value_two_in_flight"]), "enq", round(d["enqueue_ms_per_step"], 4),
          {k: round(v * 1000, 1) for k, v in d["stage_ms"].items() if k != "frames"}, "e2e", round(d["e2e"]["value"]), "roof", d.get("roofline") and (d["roofline"]["kernel"], round(d["roofline"]["frac"], 3)),
          "frame_roof", d.get("frame_roofline") and round(d["frame_roofline"]["frac"], 3), "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"], 1))
except Exception as e:
    print("$wl: no line", e)
PY
done
