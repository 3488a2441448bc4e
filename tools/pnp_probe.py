"""Eager Reg_EDDI (PNP, K = 20) training steps at cfg4 size, for `ncu --metrics gpu__time_duration.sum` (development probe)."""
import sys
import torch
sys.path.insert(0, ".")
from vae_posterior_consistency_b200 import VAE, kernels as KR, lib as L

B, D, K = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 100, 20
torch.manual_seed(0)
model = VAE.Reg_EDDI(D, 500, K, 10, {"batch_size": 64, "patience": 100}, "probe", "kl_reg")
theta = model.flat_theta().detach().clone().cuda()
tr = KR.FusedTrainer(L.FAMILY_PNP, D, K, theta, regularised=True)
x = torch.rand(B, D, device="cuda")
mask = torch.rand(B, D, device="cuda") < 0.7
mask_p = mask & (torch.rand(B, D, device="cuda") < 0.7)
eq, ep = torch.randn(B, 10, device="cuda"), torch.randn(B, 10, device="cuda")
for _ in range(3):
    tr.step(x, mask, mask_p, eq, ep)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    loss = tr.step(x, mask, mask_p, eq, ep)
e1.record()
torch.cuda.synchronize()
print(f"PNP eager step {e0.elapsed_time(e1) / 10:.3f} ms, loss {float(loss):.5f}")
