"""Eager Reg_VAE training steps at cfg4 size (batch 65536 x 100), for `ncu --metrics gpu__time_duration.sum` and variant
libraries (PCVAE_LIB=...): development probe, results of variant builds may be wrong on purpose."""
import os, sys
import torch
sys.path.insert(0, ".")
from vae_posterior_consistency_b200 import lib as L
if os.environ.get("PCVAE_LIB"):
    L.LIB_PATH = os.path.abspath(os.environ["PCVAE_LIB"])
from vae_posterior_consistency_b200 import VAE, kernels as KR

B, D = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 100
torch.manual_seed(0)
model = VAE.Reg_VAE(D, 500, 0, 10, {"batch_size": 64, "patience": 100}, "probe", "kl_reg")
theta = model.flat_theta().detach().clone().cuda()
tr = KR.FusedTrainer(L.FAMILY_MLP, D, 0, theta, regularised=True)
x = torch.rand(B, D, device="cuda")
mask = torch.rand(B, D, device="cuda") < 0.7
mask_p = mask & (torch.rand(B, D, device="cuda") < 0.7)
eq, ep = torch.randn(B, 10, device="cuda"), torch.randn(B, 10, device="cuda")
for _ in range(3):
    tr.step(x, mask, mask_p, eq, ep)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    loss = tr.step(x, mask, mask_p, eq, ep)
e1.record()
torch.cuda.synchronize()
print(f"eager step {e0.elapsed_time(e1) / 20:.4f} ms, loss {float(loss):.5f}")
