import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import torch
import irc_oracle as O
import irc_b200
from irc_b200._native import CudaBackend
from irc_b200.train_step import TrainStep
B,H,W=1,32,32
pG = O.seeded_params(O.generator_shapes(), 1, bias_std=0.02)
pD = O.seeded_params(O.discriminator_shapes(), 2, bias_std=0.02)
pV = O.seeded_params(O.vgg_shapes(), 3, kaiming=True, bias_std=0.05)
ir, rgb = O.synthetic_pair(B,H,W)
for graph in (False, False, True, True):
    ts = TrainStep(CudaBackend(), B,H,W,'cuda', use_graph=graph)
    ts.load(pG,pD,pV)
    for i in range(3):
        ts.step(ir.cuda(), rgb.cuda()); torch.cuda.synchronize()
        l = ts.losses()
        print(graph, i, {k: round(v,5) for k,v in l.items()}, ts.D2.arena.flat.abs().sum().item(), ts.G.arena.flat.abs().sum().item(), ts.optD.dev.tolist()[:7])
