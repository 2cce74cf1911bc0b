#!/bin/bash
# One multi-GPU measurement set: C2 (weak), C4 (strong) and the online C3 run on N GPUs of one box.
# usage: benchmarks/scaling_run.sh N   (results in gpurun_out/scale_n<N>_*.json)
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "1" ]; then RUN="python"; fi
timeout 600 $RUN bench.py --gpus $N --no-cpu > $OUT/scale_n${N}_c2.json 2> $OUT/scale_n${N}_c2.err
echo "C2 rc=$?"; tail -c 600 $OUT/scale_n${N}_c2.json | head -c 0; python - <<PY
import json
try:
    d = json.loads(open("$OUT/scale_n${N}_c2.json").read().strip().splitlines()[-1])
    print("C2 N=%d value %.4g ms/step %.3f e2e %s" % (d["n_gpus"], d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value")))
except Exception as e:
    print("C2 parse failed", e)
PY
timeout 900 $RUN bench.py --gpus $N --config C4 --steps 10 --no-cpu --no-e2e > $OUT/scale_n${N}_c4.json 2> $OUT/scale_n${N}_c4.err
echo "C4 rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/scale_n${N}_c4.json").read().strip().splitlines()[-1])
    print("C4 N=%d value %.4g ms/step %.3f" % (d["n_gpus"], d["value"], d["ms_per_step"]))
except Exception as e:
    print("C4 parse failed", e)
PY
timeout 900 $RUN tests/tools/online_bench.py --cpu-cells 0 > $OUT/scale_n${N}_c3.json 2> $OUT/scale_n${N}_c3.err
echo "C3 rc=$?"; tail -1 $OUT/scale_n${N}_c3.json | cut -c1-700
