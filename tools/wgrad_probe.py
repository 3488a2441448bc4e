"""Repeats forward + backward on the same inputs and compares the reduced gradients bit for bit (development probe for
the weight-gradient kernel; PCVAE_LIB selects a variant library)."""
import os, sys
import torch
sys.path.insert(0, ".")
from vae_posterior_consistency_b200 import lib as L
if os.environ.get("PCVAE_LIB"):
    L.LIB_PATH = os.path.abspath(os.environ["PCVAE_LIB"])
from vae_posterior_consistency_b200 import VAE, kernels as KR

D = 100
torch.manual_seed(0)
model = VAE.Reg_VAE(D, 500, 0, 10, {"batch_size": 64, "patience": 100}, "probe", "kl_reg")
theta = model.flat_theta().detach().clone().cuda()
offs = L.param_offsets(L.model(L.FAMILY_MLP, D, 0))
print("offsets", offs)
for B in [int(a) for a in sys.argv[1:]] or [4096, 65536]:
    tr = KR.FusedTrainer(L.FAMILY_MLP, D, 0, theta.clone(), regularised=True)
    x = torch.rand(B, D, device="cuda")
    mask = torch.rand(B, D, device="cuda") < 0.7
    mask_p = mask & (torch.rand(B, D, device="cuda") < 0.7)
    eq, ep = torch.randn(B, 10, device="cuda"), torch.randn(B, 10, device="cuda")
    tr.forward_backward(x, mask, mask_p, eq, ep)
    g0 = tr.grad.clone()
    bad = 0
    for it in range(30):
        tr.forward_backward(x, mask, mask_p, eq, ep)
        ne = (tr.grad != g0).nonzero().flatten()
        if ne.numel():
            bad += 1
            if bad <= 3:
                i = ne[:8].tolist()
                print(f"B={B} run {it}: {ne.numel()} of {g0.numel()} differ, first {i}, last {int(ne[-1])}, "
                      f"max |d| {float((tr.grad - g0).abs().max()):.3e} at |g| {float(g0.abs().max()):.3e}")
    print(f"B={B}: {bad} of 30 repeats differ")
