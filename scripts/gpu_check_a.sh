#!/bin/bash
# On-box: GPU tests, smoke, bench, then ncu launch list + one full capture of the dequant-fused GEMM.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -q -m gpu -s --durations=12 -p no:cacheprovider > $O/r2_t2.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2_t2.log
tail -4 $O/r2_t2.log
python __graft_entry__.py smoke > $O/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2_smoke.log
python bench.py --steps 5 --warmup 3 > $O/r2_bench_a.json 2> $O/r2_bench_a.err; echo "bench rc=$?"; tail -c 600 $O/r2_bench_a.err
python bench.py --steps 2 --warmup 1 --no-extras > $O/r2_plain.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-extras > $O/r2_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tc_skinny_q -s 1500 -c 6 -o $O/r2_skq python bench.py --steps 2 --warmup 1 --no-extras > $O/r2_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la $O | tail -12
