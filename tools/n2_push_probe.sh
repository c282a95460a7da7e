for v in base r2base; do
lib=cudaraster-linux_b200/libcrb200.so; [ $v = r2base ] && lib=build/variants/r2base/libcrb200.so
CRB200_LIBRARY=$lib python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 30 --warmup 3 --no-ref-kernels --no-cpu-baseline --composite push 2>/dev/null | tail -1 > gpurun_out/n2_push_$v.json
python - <<PY
import json
d=json.load(open("gpurun_out/n2_push_$v.json"))
print("$v", round(d["value"]), round(d["ms_per_step"],4), d.get("per_rank"))
PY
done
