import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn, torch.nn.functional as F
from pika_b200 import engine as E, kernels as K
torch.manual_seed(0)
def g(*shape, seed=0):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=gen)
for prec in ("bf16", "fp32"):
    E.set_precision(prec)
    for (dil, stride, T, C) in [(3, 1, 90, 256), (3, 1, 140, 256), (1, 1, 90, 256), (3, 1, 90, 128), (3, 1, 200, 1024)]:
        conv = nn.Conv2d(1, C, (3, C), dilation=(dil, 1), stride=(stride, 1)).cuda()
        for rep in range(2):
            x = g(3, T, C, seed=8).requires_grad_(True)
            xin = x if prec == "fp32" else x.detach().to(torch.bfloat16).requires_grad_(True)
            y = E.TdnnFn.apply(xin, conv.weight, conv.bias, dil, stride)
            ref = F.relu(conv(x.unsqueeze(1))).squeeze(-1).transpose(1, 2)
            dy = g(*ref.shape, seed=9)
            (gx,) = torch.autograd.grad(ref, [x], dy)
            conv.weight.grad = conv.bias.grad = None
            y.backward(dy.to(y.dtype))
            err = (xin.grad.float() - gx)
            rowerr = err.norm(dim=2) / gx.norm(dim=2).clamp_min(1e-9)     # [B,T]
            bad = (rowerr > 2e-2).nonzero()
            print(prec, (dil, stride, T, C), "rep", rep, "rel", (err.norm() / gx.norm()).item(), "bad rows", bad[:12].tolist(), len(bad), flush=True)
