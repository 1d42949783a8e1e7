#!/bin/bash
# N-GPU data-parallel bench under NCCL channel settings (needs gpurun --gpus N)
out=$1; n=$2; : > $out
run() {
  label=$1; shift
  ms=$(env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline 2>/dev/null | python -c "import json,sys; print(json.loads(sys.stdin.readline())['ms_per_step'])")
  echo "$label $ms" >> $out
}
one=$(python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline 2>/dev/null | python -c "import json,sys; print(json.loads(sys.stdin.readline())['ms_per_step'])")
echo "n1 $one" >> $out
run "n${n}_default" IRC_X=0
run "n${n}_nch4" NCCL_MAX_NCHANNELS=4
run "n${n}_nch8" NCCL_MAX_NCHANNELS=8
run "n${n}_default" IRC_X=0
run "n${n}_nch4" NCCL_MAX_NCHANNELS=4
