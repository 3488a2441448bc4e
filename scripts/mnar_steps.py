"""A few eager REG_notMIWAE_v2 training steps at cfg2 sizes (for an ncu launch list)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_posterior_consistency_b200 import VAE

dev = torch.device("cuda")
D, B, S = 50, 128, 20
torch.manual_seed(0)
model = VAE.REG_notMIWAE_v2(D, 500, 20, 10, {"batch_size": B, "patience": 100}, S, 10).to(dev)
model.noise = "device"
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    x = torch.rand(B, D, device=dev)
    mask = (torch.rand(B, D, device=dev) < 0.7).float()
    mask_p = mask * (torch.rand(B, D, device=dev) < 0.5).float()
    mean_p, logvar_p, xm_p, xlv_p, mean_q, logvar_q, xm_q, xlv_q = model.forward(x, mask, mask_p, stage="train")
    loss = model.loss(x, xm_p, xlv_p, mean_p, logvar_p, xm_q, xlv_q, mean_q, logvar_q, mask, mask_p, 1, alpha=1.0)[1]
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("ok", float(loss))
