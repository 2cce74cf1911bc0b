#!/bin/bash
# Round-2 profile session on one B200 (run under gpurun): bench line, reference arm, ncu launch list and one
# `ncu --set full` capture of the passes and the H solve.  Outputs under gpurun_out/; benchmarks/make_profiles.py turns
# them into the tracked summaries under profiles/.
set -u
O=gpurun_out
python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_ref.json 2> $O/r2_bench_ref.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra --graph 0"
$CMD > $O/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2_launches.csv $CMD > $O/r2_ncu_l.log 2>&1
$CMD > $O/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:sell32_pass|bpp16_kernel' -s 13 -c 3 -o $O/r2_prof $CMD > $O/r2_ncu_f.log 2>&1
tail -2 $O/r2_ncu_f.log
