#!/bin/bash
# developer tool (run under gpurun --gpus N): the default bench, the sort-first 4K frame and the 48-view batch at N GPUs
N=$1; tag=${2:-r2}
mkdir -p gpurun_out
run() { if [ $N = 1 ]; then python "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 "$@"; fi; }
run bench.py --gpus $N --steps 30 --warmup 3 --no-ref-kernels 2> gpurun_out/${tag}_bench_c2_n$N.err | tail -1 > gpurun_out/${tag}_bench_c2_n$N.json
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench_c2_n$N.json"))
print("bench N=$N value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "bracket", round(d["bracket_ms_per_step"], 4), "enq", round(d["enqueue_ms_per_step"], 4), "chain", round(d["value_unbroken_chain"]), "two", round(d["value_two_in_flight"] or 0), "e2e", round(d["e2e"]["value"]), d["notes"]["composite"] and d["notes"]["composite"][-12:])
PY
run tools/sort_first_4k.py --frames 30 2> gpurun_out/${tag}_sf_n$N.err | tail -1 > gpurun_out/${tag}_sortfirst4k_n$N.json; cat gpurun_out/${tag}_sortfirst4k_n$N.json | cut -c1-120
run tools/sort_first_4k.py --frames 30 --no-bounds 2>> gpurun_out/${tag}_sf_n$N.err | tail -1 > gpurun_out/${tag}_sortfirst4k_nobounds_n$N.json; cat gpurun_out/${tag}_sortfirst4k_nobounds_n$N.json | cut -c1-120
if [ -f tools/views_48.py ]; then run tools/views_48.py 2> gpurun_out/${tag}_views_n$N.err | tail -1 > gpurun_out/${tag}_views48_n$N.json; cat gpurun_out/${tag}_views48_n$N.json | cut -c1-160; fi
