"""Test-mode throughput (BASELINE configs 1 and 5): generator forward + truncating quantisation + MAE/MSE/PSNR per image,
device-resident inputs, CUDA events.  Usage: python scripts/bench_infer.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import torch
import irc_b200 as R
from irc_b200.train import batch_metrics

cfg = R.Config(); cfg.device = "cuda"
model = R.IRColorizationModel(cfg).eval()
for B, H, W in ((1, 256, 256), (16, 256, 256), (8, 512, 640), (64, 512, 640)):
    g = torch.Generator().manual_seed(1)
    ir = (torch.rand(B, 1, H, W, generator=g) * 2 - 1).cuda(); gt = torch.rand(B, 3, H, W, generator=g).cuda()
    with torch.no_grad():
        for _ in range(3):
            fake = model(ir); batch_metrics(fake, gt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            fake = model(ir)
            u8, mae, mse, psnr = batch_metrics(fake, gt)      # includes the D2H of the per-image sums
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gflop = 136.94 * B * H * W / 65536
    print(f"B={B:3d} {H}x{W}: {ms:8.3f} ms/batch  {B / ms * 1e3:8.1f} img/s  {gflop / ms:7.1f} TFLOP/s (generator conv FLOPs)")
