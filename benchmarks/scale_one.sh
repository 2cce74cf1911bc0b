#!/bin/bash
# usage: benchmarks/scale_one.sh N   (under gpurun --gpus N): the bench line at N GPUs -> gpurun_out/r2_bench_nN.json
N=$1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
timeout 700 $RUN bench.py --gpus $N --no-cpu --extra c4,c3,parity,fused > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo rc=$?
python - <<PY
import json
L=[l for l in open("gpurun_out/r2_bench_n$N.json").read().splitlines() if l.startswith("{")]
d=json.loads(L[-1]); k=d["roofline"]["kernels"]
print("C2 ms %.4f value %.4g e2e %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), {p:k[p]["ms"] for p in k})
for n,v in d["extra"].items():
    if isinstance(v,dict): print(n, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ("ms_per_step","value","active","pass2_plus_reduction_ms","other_mode_pass2_plus_reduction_ms","train_cell_visits_per_s","final_H_cells_per_s","objective_rel","H_rel","ok")})
PY
