#!/bin/bash
# Round-2 profile session on one B200 (run under gpurun): bench line, reference arm, ncu launch list and one
# `ncu --set full` capture of the passes and the H solve.  Outputs under gpurun_out/; benchmarks/make_profiles.py turns
# them into the tracked summaries under profiles/.
set -u
O=gpurun_out
python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
if [ "${SKIP_REF:-0}" != "1" ]; then python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_ref.json 2> $O/r2_bench_ref.err; fi
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra --graph 0"
$CMD > $O/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2_launches.csv $CMD > $O/r2_ncu_l.log 2>&1
$CMD > $O/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:sell32_pass|bpp16_kernel' -s 13 -c 3 -o $O/r2_prof $CMD > $O/r2_ncu_f.log 2>&1
tail -2 $O/r2_ncu_f.log
# k = 40 passes (one lane per row + remainder tile) on a C4 slice, and the online step's kernels
C4="python benchmarks/pass_variants.py --config C4 --variants sell"
$C4 > $O/r2_c4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sell32r_pass -s 4 -c 2 -o $O/r2_prof_c4 $C4 > $O/r2_ncu_c4.log 2>&1
ON="python tests/tools/online_bench.py --cells 10000 --cpu-cells 0"
$ON > $O/r2_online_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:hals_grouped|sell32r_pass|blocked_split' -s 12 -c 3 -o $O/r2_prof_online $ON > $O/r2_ncu_on.log 2>&1
tail -n 2 $O/r2_ncu_c4.log; tail -n 2 $O/r2_ncu_on.log
# the reports themselves are too large to travel back (64 MiB limit): summarise here, keep the text
for r in r2_prof r2_prof_c4 r2_prof_online; do
  if [ -f $O/$r.ncu-rep ]; then python benchmarks/ncu_summary.py $O/$r.ncu-rep > $O/$r.summary.txt 2>&1; rm -f $O/$r.ncu-rep; fi
done
ls -la $O | tail -n 20
