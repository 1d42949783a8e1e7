"""Where do the conv_gemm pipeline roles wait?  Per-CTA cycle counters for a few layer shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend
from irc_b200 import layout as L
be = CudaBackend()
def run(name, B, H, W, Cin, Cout, mt):
    fr = L.Frame(B, H, W, 1, Cin, "cuda"); fr.t.normal_()
    out = torch.zeros(fr.rows, Cout, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(Cout, 9 * Cin, device="cuda") * 0.02).bfloat16()
    taps = L.taps_centered(3, 3, fr.wp)
    be.conv_mt = mt
    for _ in range(2): be.conv_gemm(fr.t, 0, Cin, taps, w, Cout, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(5): be.conv_gemm(fr.t, 0, Cin, taps, w, Cout, out)
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 5
    dbg = torch.zeros(148 * 8, device="cuda", dtype=torch.int64)
    be.conv_dbg = dbg
    be.conv_gemm(fr.t, 0, Cin, taps, w, Cout, out); torch.cuda.synchronize()
    be.conv_dbg = None
    d = dbg.view(148, 8).double()
    tot = d[:, 4].mean().item()
    for mode in (1, 2, 3, 4):
        be.conv_dbg_mode = mode
        be.conv_gemm(fr.t, 0, Cin, taps, w, Cout, out); torch.cuda.synchronize()
        e0.record()
        for _ in range(5): be.conv_gemm(fr.t, 0, Cin, taps, w, Cout, out)
        e1.record(); torch.cuda.synchronize()
        print(f"    mode {mode} ({ {1: 'MMA only, no TMA', 2: 'TMA only, no MMA', 3: 'MMA only, no per-k-block commit', 4: 'MMA only, no commit, no fence'}[mode] }): {e0.elapsed_time(e1) / 5 * 1e3:.1f} us")
    be.conv_dbg_mode = 0
    print(f"{name} mt={mt}: {ms*1e3:.1f} us {2*B*H*W*Cin*9*Cout/ms/1e9:.0f} TF | cycles/CTA {tot:.0f}: mma waits data {100*d[:,0].mean().item()/tot:.0f}%, mma waits tmem {100*d[:,1].mean().item()/tot:.0f}%, producer waits empty {100*d[:,2].mean().item()/tot:.0f}%, epilogue waits acc {100*d[:,3].mean().item()/tot:.0f}%")
run("res 256->256 64^2", 16, 64, 64, 256, 256, 1)
run("up2 192->64 256^2", 16, 256, 256, 192, 64, 1)
run("up2 192->64 256^2", 16, 256, 256, 192, 64, 2)
