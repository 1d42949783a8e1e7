"""Pure tcgen05.mma issue/execute rate (no TMA, no commits): cycles per MMA vs N and number of interleaved accumulators."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend
from irc_b200 import layout as L
be = CudaBackend()
B, H, W, Cin = 16, 128, 128, 128
fr = L.Frame(B, H, W, 1, Cin, "cuda"); fr.t.normal_()
taps = L.taps_centered(3, 3, fr.wp)
for Cout in (32, 64, 128, 256):
    for mt in (1, 2, 4):
        if mt * Cout > 512: continue
        out = torch.zeros(fr.rows, Cout, device="cuda", dtype=torch.bfloat16)
        w = (torch.randn(Cout, 9 * Cin, device="cuda") * 0.02).bfloat16()
        be.conv_mt = mt
        for mode in (4, 0):

            be.conv_dbg_mode = mode
            dbg = torch.zeros(148 * 8, device="cuda", dtype=torch.int64)
            be.conv_gemm(fr.t, 0, Cin, taps, w, Cout, out); torch.cuda.synchronize()
            be.conv_dbg = dbg
            be.conv_gemm(fr.t, 0, Cin, taps, w, Cout, out); torch.cuda.synchronize()
            be.conv_dbg = None
            d = dbg.view(148, 8).double()
            tiles = -(-fr.rows // (128 * mt))
            mmas_per_cta = tiles / 148 * 18 * 4 * mt
            print(f"N={Cout:3d} mt={mt} mode={mode}: cycles/CTA {d[:,4].mean().item():9.0f}  MMAs/CTA {mmas_per_cta:7.0f}  cycles/MMA {d[:,4].mean().item()/mmas_per_cta:6.1f} (nominal {128*Cout/256:.0f}) mma-waits-tmem {100*d[:,1].mean().item()/d[:,4].mean().item():.0f}%")
be.conv_dbg_mode = 0
