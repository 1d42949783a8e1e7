#!/bin/bash
# N = 2 data-parallel bench under a few SM-reservation / NCCL-channel settings (needs gpurun --gpus 2)
out=$1; : > $out
run() {  # label, env...
  label=$1; shift
  ms=$(env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline 2>/dev/null | python -c "import json,sys; print(json.loads(sys.stdin.readline())['ms_per_step'])")
  echo "$label $ms" >> $out
}
one=$(python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline 2>/dev/null | python -c "import json,sys; print(json.loads(sys.stdin.readline())['ms_per_step'])")
echo "n1 $one" >> $out
for rep in 1 2; do
run "n2_default" IRC_X=0
run "n2_sm140" IRC_SM_LIMIT=140
run "n2_sm132" IRC_SM_LIMIT=132
run "n2_nch4" NCCL_MAX_NCHANNELS=4
run "n2_nch4_sm144" NCCL_MAX_NCHANNELS=4 IRC_SM_LIMIT=144
run "n2_nch8_sm140" NCCL_MAX_NCHANNELS=8 IRC_SM_LIMIT=140
done
one=$(IRC_SM_LIMIT=140 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline 2>/dev/null | python -c "import json,sys; print(json.loads(sys.stdin.readline())['ms_per_step'])")
echo "n1_sm140 $one" >> $out
