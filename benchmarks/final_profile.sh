#!/bin/bash
# Round-end measurement set on one B200 (run under gpurun): tests, bench (both arms), ncu launch list + full capture.
# Every ncu command runs only after the same command has exited 0 without ncu.
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py > $OUT/bench_final_n1.json 2> $OUT/bench_final_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_final_ref.json 2> $OUT/bench_final_ref.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $OUT/plain_final.log 2>&1; rc=$?; echo "plain rc=$rc"
if [ $rc -eq 0 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_final.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_l_final.log 2>&1; echo "ncu list rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:'blocked_spmm|bpp_kernel' -s 20 -c 5 \
      -o $OUT/prof_final -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_f_final.log 2>&1; echo "ncu full rc=$?"
fi
python tests/tools/bpp_microbench.py 2>/dev/null > $OUT/c5_bpp_final.jsonl; echo "c5 rc=$?"
python tests/tools/c1_parity.py > $OUT/c1_parity_final.json 2> $OUT/c1_final.err; echo "c1 rc=$?"
tail -1 $OUT/bench_final_n1.json | cut -c1-600
