"""First-item check of k_enc_fwd_tc: the first work item of every CTA against the oracle, three launches in a fresh
process (B = 300, D = 100: every CTA has at most two items).  PCVAE_ENC_FLAT=0 selects the nested item order."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pcvae_oracle as O
from vae_posterior_consistency_b200 import kernels as KR, lib as L
B, D = 300, 100
p = O.init_params("mlp", D, 0, seed=B + D)
g = torch.Generator().manual_seed(B + D + 100)
x = torch.rand(B, D, generator=g)
mask = torch.rand(B, D, generator=g) < 0.7
mask_p = mask & (torch.rand(B, D, generator=g) < 0.7)
eq, ep = torch.randn(B, 10, generator=g), torch.randn(B, 10, generator=g)
mu_q, lv_q = O.encoder_stats(p, x, mask); mu_p, lv_p = O.encoder_stats(p, x, mask_p)
theta = KR.flatten_params(p, L.FAMILY_MLP, "cuda")
if os.environ.get("POISON", "1") == "1":
    junk = torch.full((64 << 20,), float("nan"), device="cuda"); del junk
eng = KR.Engine(L.FAMILY_MLP, D, 0, "cuda")
xc, mc, mpc, eqc, epc = x.cuda(), mask.cuda(), mask_p.cuda(), eq.cuda(), ep.cuda()
for call in range(3):
    mean, logvar, z, ws = eng.enc_fwd(theta, xc, [mc, mpc], [eqc, epc], save=True)
    torch.cuda.synchronize()
    e0 = float((mean[0].cpu() - mu_q).abs().max()); e1 = float((mean[1].cpu() - mu_p).abs().max())
    print(f"flat={os.environ.get('PCVAE_ENC_FLAT','0')} poison={os.environ.get('POISON','1')} call {call}: err q {e0:.2e} p {e1:.2e}")
