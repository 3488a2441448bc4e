"""Debug helper: fused step through the FFMA kernels vs the tcgen05 kernels, several shapes, repeated; prints which
gradient tensors / rows disagree."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import pcvae_oracle as O
from vae_posterior_consistency_b200 import kernels as KR, lib as L

lib = L.load()
torch.manual_seed(0)
for (B, D) in [(300, 100), (129, 20), (1000, 100), (4100, 100), (513, 96), (65536, 100)]:
    p = O.init_params("mlp", D, 0, seed=B + D)
    g = torch.Generator().manual_seed(B)
    x = torch.rand(B, D, generator=g).cuda()
    mask = (torch.rand(B, D, generator=g) < 0.7)
    mask_p = (mask & (torch.rand(B, D, generator=g) < 0.7)).cuda()
    mask = mask.cuda()
    eq, ep = torch.randn(B, 10, generator=g).cuda(), torch.randn(B, 10, generator=g).cuda()
    theta = KR.flatten_params(p, L.FAMILY_MLP, "cuda")
    res = {}
    for tc in (0, 1):
        prev = lib.pcvae_set_train_tensor_cores(tc)
        outs = []
        for rep in range(1 if tc == 0 else 6):
            tr = KR.FusedTrainer(L.FAMILY_MLP, D, 0, theta, regularised=True, alpha=0.6, beta_w=0.9)
            tr.forward_backward(x, mask, mask_p, eq, ep)
            torch.cuda.synchronize()
            outs.append(KR.unflatten_params(tr.grad.clone(), L.FAMILY_MLP, D, 0))
        lib.pcvae_set_train_tensor_cores(prev)
        res[tc] = outs
    ref = res[0][0]
    for rep, got in enumerate(res[1]):
        bad = []
        for k in ref:
            a, b = got[k].cpu(), ref[k].cpu()
            tol = 2e-4 * b.abs() + 2e-6 * float(b.abs().max())
            m = (a - b).abs() > tol
            if m.any():
                rows = sorted(set(m.nonzero()[:, 0].tolist()))
                bad.append((k, int(m.sum()), rows[:12], float((a - b).abs().max())))
        print(f"B={B} D={D} rep={rep}: {'OK' if not bad else bad}", flush=True)
