import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend
be = CudaBackend()
def run(name, rows, Cin, Cout, taps, mt, direct=0):
    a = torch.randn(rows, Cin, device="cuda").bfloat16()
    out = torch.zeros(rows, Cout, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(Cout, len(taps) * Cin, device="cuda") * 0.02).bfloat16()
    be.conv_mt = mt; be.conv_epilogue_direct = direct
    for _ in range(2): be.conv_gemm(a, 0, Cin, taps, w, Cout, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): be.conv_gemm(a, 0, Cin, taps, w, Cout, out)
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 5
    dbg = torch.zeros(148 * 8, device="cuda", dtype=torch.int64)
    be.conv_dbg = dbg
    be.conv_gemm(a, 0, Cin, taps, w, Cout, out); torch.cuda.synchronize()
    be.conv_dbg = None
    d = dbg.view(148, 8).double(); tot = d[:, 4].mean().item()
    gb = rows * (Cin + Cout) * 2 / ms / 1e6
    print(f"{name} mt={mt} direct={direct}: {ms*1e3:.1f} us  {gb:.0f} GB/s in+out | cycles/CTA {tot:.0f}: mma waits data {100*d[:,0].mean().item()/tot:.0f}%, mma waits tmem {100*d[:,1].mean().item()/tot:.0f}%, producer waits empty {100*d[:,2].mean().item()/tot:.0f}%, epilogue waits acc {100*d[:,3].mean().item()/tot:.0f}%, epilogue in tcgen05.ld {100*d[:,5].mean().item()/tot:.0f}%, epilogue busy {100*d[:,6].mean().item()/tot:.0f}% (warp2 lifetime {d[:,7].mean().item():.0f})")
R = 32 * 258 * 258
for mode in (0,):
    be.conv_dbg_mode = mode
    print("dbg_mode", mode)
    run("1tap 64->64", R, 64, 64, [0], 4, 1)
be.conv_dbg_mode = 0
run("9tap 64->128 (down1)", R // 2, 64, 128, [-259, -258, -257, -1, 0, 1, 257, 258, 259], 2, 1)
