import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200 import _native as nat
nat.arch_check()
def ref(A, taps, W, cin):
    rows = A.shape[0]; Af = A.float(); out = torch.zeros(rows, W.shape[0], device=A.device); q = torch.arange(rows, device=A.device)
    for t, sh in enumerate(taps):
        idx = q + sh; ok = (idx >= 0) & (idx < rows)
        out += (Af[idx.clamp(0, rows - 1), :cin] * ok[:, None]) @ W[:, t * cin:(t + 1) * cin].float().t()
    return out
def run(rows, cin, taps, n_out, reuse):
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(rows, cin, device="cuda", generator=g).bfloat16()
    W = (torch.randn(n_out, len(taps) * cin, device="cuda", generator=g) * 0.05).bfloat16()
    out = torch.zeros(rows, n_out, device="cuda", dtype=torch.bfloat16)
    a = nat.ConvGemmArgs()
    a.a = A.data_ptr(); a.a_rows = rows; a.a_ld = cin; a.cin = cin; a.ntaps = len(taps)
    for i, t in enumerate(taps): a.taps[i] = t
    a.w = W.data_ptr(); a.n_out = n_out; a.out = out.data_ptr(); a.out_ld = n_out; a.reuse = reuse; a.mt = 1
    nat.check(nat.lib().irc_conv_gemm(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    r = ref(A, taps, W, cin)
    return ((out.float() - r).norm() / r.norm()).item()
for reuse in (0, 1, 3):
    for (rows, cin, taps, n_out) in [(3000, 64, [-1, 0, 1], 64), (5000, 128, [-67, -66, -65, -1, 0, 1, 65, 66, 67], 256), (4000, 256, [0, 1, 2, 3, 34, 35, 36, 37], 128)]:
        try:
            print("reuse", reuse, rows, cin, len(taps), n_out, "rel err %.5f" % run(rows, cin, taps, n_out, reuse))
        except Exception as e:
            print("reuse", reuse, "FAILED", e); break
