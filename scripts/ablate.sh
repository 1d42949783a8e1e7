#!/bin/bash
# real cost of each launcher inside the captured step: ms/step with that launcher turned into a no-op (results are wrong;
# timing only).  Usage (GPU box): bash scripts/ablate.sh > gpurun_out/ablate.txt
run() { IRC_SKIP=$1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; }
echo "none $(run '')"
for k in in_bwd in_stats gather fold_inplace im2col col2im gather_sum tap_expand tap_reduce maxpool2,maxpool2_bwd pack_bf16,adam ssim_fwd,ssim_bwd,pixel_loss,feat_l1,hinge,colsum tn_gemm conv_gemm; do
  echo "$k $(run $k)"
done
