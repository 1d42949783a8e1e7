#!/bin/bash
# Everything the profiles/ tables of a round are made of, in one GPU call: the full GPU test suite, smoke(), the bench line with
# both breakdown tables, the reference arm, the config-4 / 1 / 5 JSON lines, the memory-bound kernel tables and the ncu launch
# list (two steps, set-up launches skipped) of the same bench command.  usage: scripts/round_end.sh OUTDIR   (OUTDIR under gpurun_out/)
out=$1; mkdir -p $out
python -m pytest tests -q -m gpu > $out/pytest_gpu.log 2>&1; echo "rc=$?" >> $out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "rc=$?" >> $out/smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
python bench.py --steps 20 --warmup 5 --breakdown $out/gemm_breakdown.csv --breakdown-all $out/launchers.csv > $out/bench.json 2> $out/bench.err
python bench.py --size 512 640 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > $out/bench_config4.json 2> $out/bench_config4.err
python scripts/bench_infer.py > $out/bench_infer.json 2> $out/bench_infer.err
python scripts/bench_elem_large.py > $out/bench_elem_large.json 2> $out/bench_elem_large.err
python scripts/bench_elem.py > $out/bench_elem.txt 2>&1
python scripts/prof_gathers.py 7 > $out/gathers.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 257 -c 444 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-eager-baseline > $out/ncu_bench.log 2>&1
tail -3 $out/pytest_gpu.log; tail -1 $out/smoke.log; head -c 600 $out/bench.json
