import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irc_b200  # noqa
from irc_b200._native import CudaBackend
be = CudaBackend()
rows = 16 * 258 * 258
a = torch.randn(rows, 64, device="cuda").bfloat16()
out = torch.zeros(rows, 64, device="cuda", dtype=torch.bfloat16)
w = (torch.randn(64, 64, device="cuda") * 0.02).bfloat16()
be.conv_mt = 4; be.conv_epilogue_direct = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(3): be.conv_gemm(a, 0, 64, [0], w, 64, out)
torch.cuda.synchronize()
print("ok")
