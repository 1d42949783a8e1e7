import sys, os, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import torch
import irc_b200 as R
cfg = R.Config(); cfg.device = "cuda"
model = R.IRColorizationModel(cfg).eval()
ir = torch.rand(1, 1, 256, 256).cuda()
with torch.no_grad():
    for _ in range(3):
        model(ir)
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(3):
        model(ir)
    torch.cuda.synchronize()
    pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
