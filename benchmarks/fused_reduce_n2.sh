#!/bin/bash
# 2-GPU comparison of the two reduction modes (run under gpurun --gpus 2): bench.py with --fused-reduce 1 / 0, C2 + C4.
# Results: profiles/r2_fused_reduce_n2.txt.
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for f in 1 0; do for g in 1 0; do
timeout 400 $RUN bench.py --gpus 2 --no-cpu --no-e2e --extra c4 --fused-reduce $f --graph $g > gpurun_out/n2_f${f}_g${g}.json 2> gpurun_out/n2_f${f}_g${g}.err; echo "fused=$f graph=$g rc=$?"
python - <<PY
import json
try:
    L=[l for l in open("gpurun_out/n2_f${f}_g${g}.json").read().splitlines() if l.startswith("{")]
    d=json.loads(L[-1]); k=d["roofline"]["kernels"]
    print("C2 ms %.4f value %.4g" % (d["ms_per_step"], d["value"]), {p:k[p]["ms"] for p in k}, d["config"]["parallelism"][:60])
    c=d["extra"]["c4_strong"]; k=c["roofline"]["kernels"]
    print("C4 ms %.4f value %.4g fused=%s" % (c["ms_per_step"], c["value"], c.get("fused_reduce")), {p:k[p]["ms"] for p in k})
except Exception as e:
    print("parse failed", e)
PY
done; done
