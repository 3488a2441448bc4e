"""Compare the tcgen05 decoder path with the FFMA one piece by piece (debug aid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pcvae_oracle as O
from vae_posterior_consistency_b200 import kernels as KR, lib as L

B, D = int(sys.argv[1]), int(sys.argv[2])
p = O.init_params("mlp", D, 0, seed=5)
g = torch.Generator().manual_seed(7)
x = torch.rand(B, D, generator=g); mask = torch.rand(B, D, generator=g) < 0.7
mask_p = mask & (torch.rand(B, D, generator=g) < 0.7)
eq, ep = torch.randn(B, 10, generator=g), torch.randn(B, 10, generator=g)
lib = L.load()
res = []
for tc in (0, 1):
    lib.pcvae_set_train_tensor_cores(tc)
    eng = KR.Engine(L.FAMILY_MLP, D, 0, "cuda"); theta = KR.flatten_params(p, L.FAMILY_MLP, "cuda")
    masks = [mask.cuda(), mask_p.cuda()]; eps = [eq.cuda(), ep.cuda()]
    mean, logvar, z, ws = eng.enc_fwd(theta, x.cuda(), masks, eps, save=True)
    eng.grad_partials().zero_()
    out = eng.dec(L.DEC_TRAIN, theta, z, x=x.cuda(), masks=masks, mean=mean, logvar=logvar, eps=eps, alpha=0.6, beta_w=0.9, loss_scale=1.0 / B)
    if tc:
        w = eng._last_tcw.cpu(); R2 = 2 * B; R2P = (R2 + 31) // 32 * 32; o = 0
        for name, feats in (("zT", 16), ("h4T", 56), ("h5T", 104), ("dp6T", 104), ("dp5T", 104), ("dp4T", 56)):
            buf = w[o:o + R2P * feats].view(feats, R2P); o += R2P * feats
            print(name, "absmax", float(buf[:, :R2].abs().max()), "pad absmax", float(buf[:, R2:].abs().max()) if R2P > R2 else 0.0,
                  "row0 feats[:6]", [round(v, 5) for v in buf[:6, 0].tolist()])
    grad = torch.zeros_like(theta)
    eng.reduce_grads(grad)
    torch.cuda.synchronize()
    res.append(dict(dm=[t.cpu() for t in out["d_mean"]], dv=[t.cpu() for t in out["d_logvar"]],
                    grads={k: v.cpu().clone() for k, v in KR.unflatten_params(grad, L.FAMILY_MLP, D, 0).items()}))
a, b = res
for br in range(2):
    print("d_mean", br, float((a["dm"][br] - b["dm"][br]).abs().max()), float(a["dm"][br].abs().max()))
    print("d_logvar", br, float((a["dv"][br] - b["dv"][br]).abs().max()), float(a["dv"][br].abs().max()))
for k in a["grads"]:
    if "decoder" in k:
        d = (a["grads"][k] - b["grads"][k]).abs()
        print(k, tuple(d.shape), "maxerr", float(d.max()), "ref max", float(a["grads"][k].abs().max()), "tc max", float(b["grads"][k].abs().max()))
        if d.max() > 1e-4 * a["grads"][k].abs().max() and d.dim() == 2:
            bad = (d > 1e-4 * a["grads"][k].abs().max())
            print("   bad rows", bad.any(1).nonzero().flatten()[:12].tolist(), "bad cols", bad.any(0).nonzero().flatten()[:12].tolist())
            print("   ref[0,:6]", a["grads"][k][0, :6].tolist()); print("   tc [0,:6]", b["grads"][k][0, :6].tolist())
