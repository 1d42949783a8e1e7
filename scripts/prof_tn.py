import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend
from irc_b200 import layout as L
be = CudaBackend()
def run(name, B, H, W, Cout, Cin, splits_list, a_off=0, b_off=0, ldb=None):
    fr = L.Frame(B, H, W, 1, ldb or Cin, "cuda"); fr.t.normal_()
    dz = L.Frame(B, H, W, 1, Cout, "cuda"); dz.t.normal_()
    taps = L.taps_centered(3, 3, fr.wp)
    for splits in splits_list:
        part = torch.zeros(splits * Cout * 9 * Cin, device="cuda")
        f = lambda: be.tn_gemm(dz.t, 0, Cout, fr.t, b_off, Cin, fr.rows, [0] * 9, taps, part, Cin, 9 * Cin, 1, splits, Cout * 9 * Cin)
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name} splits={splits:3d}: {ms*1e3:7.1f} us  {2*B*H*W*Cout*9*Cin/ms/1e9:6.0f} TFLOP/s")
run("res 256x256 64^2", 16, 64, 64, 256, 256, [4, 8, 12, 16, 24, 32])
run("up2 64<-192 256^2", 16, 256, 256, 64, 192, [8, 16, 33, 49, 64])
run("down1 128<-64 256^2", 16, 256, 256, 128, 64, [8, 16, 33, 64])
run("up1 128<-384 128^2", 16, 128, 128, 128, 384, [4, 8, 16])
