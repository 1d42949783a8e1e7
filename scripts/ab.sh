#!/bin/bash
# A/B timing of the captured train step under environment-variable variants, interleaved so that box-level drift
# (power cap, clocks) hits every variant alike.  usage: scripts/ab.sh OUTFILE REPS "VAR=a VAR2=b" "VAR=c" ...
out=$1; reps=$2; shift 2
: > $out
for r in $(seq $reps); do
  for v in "$@"; do
    ms=$(env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline 2>/dev/null | python -c "import json,sys; print(json.loads(sys.stdin.readline())['ms_per_step'])")
    echo "$r [$v] $ms" >> $out
  done
done
