import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import pcvae_oracle as O
from vae_posterior_consistency_b200 import kernels as KR, lib as L
lib = L.load()
B, D = int(sys.argv[1]), 100
p = O.init_params("mlp", D, 0, seed=B + D)
g = torch.Generator().manual_seed(B)
x = torch.rand(B, D, generator=g).cuda()
mask = (torch.rand(B, D, generator=g) < 0.7)
mask_p = (mask & (torch.rand(B, D, generator=g) < 0.7)).cuda(); mask = mask.cuda()
eq, ep = torch.randn(B, 10, generator=g).cuda(), torch.randn(B, 10, generator=g).cuda()
theta = KR.flatten_params(p, L.FAMILY_MLP, "cuda")
eng = KR.Engine(L.FAMILY_MLP, D, 0, "cuda")
res = {}
for tc in (0, 1):
    lib.pcvae_set_train_tensor_cores(tc)
    mean, logvar, z, ws = eng.enc_fwd(theta, x, [mask, mask_p], [eq, ep], save=True)
    torch.cuda.synchronize()
    res[tc] = (mean, logvar, z, ws)
for i, nm in enumerate(["mean", "logvar", "z"]):
    for br in range(2):
        d = (res[0][i][br] - res[1][i][br]).abs()
        print(nm, br, "max diff", float(d.max()), "rows bad", (d.max(dim=1).values > 1e-4).nonzero().flatten().tolist()[:10])
ws = res[1][3]
ntiles = (B + 127) // 128
nvt = 2 * ntiles
inT = ws[:nvt * 128 * 104].view(nvt, 4, 104, 32)      # [vt][slab][f][32]
for br, m in enumerate([mask, mask_p]):
    xm = (x * m.float())
    for t in range(ntiles):
        blk = inT[br * ntiles + t]                      # [4][104][32]
        rows = blk.permute(0, 2, 1).reshape(128, 104)   # [row][f]
        n = min(128, B - t * 128)
        d = (rows[:n, :D] - xm[t * 128:t * 128 + n]).abs().max()
        one = rows[:, D]
        print("inT br", br, "tile", t, "max diff", float(d), "bias col min/max", float(one.min()), float(one.max()), "pad rows absmax", float(rows[n:, :D].abs().max()) if n < 128 else 0.0)
m0, m1 = res[0][0], res[1][0]
for br in range(2):
    d = (m0[br] - m1[br]).abs().max(dim=1).values
    print("br", br, "rows with diff>1e-4:", int((d > 1e-4).sum()), "of", B, "first ok rows", (d <= 1e-4).nonzero().flatten().tolist()[:10])
    print("  tc row0", m1[br][0, :5].tolist(), "ffma row0", m0[br][0, :5].tolist())
    # does the TC result match the FFMA result of the other branch?
    print("  match other branch:", float((m0[1 - br] - m1[br]).abs().max()))
# h1 scratch vs FFMA-equivalent recompute
W1 = p["seq_encoder.0.weight"].cuda(); b1 = p["seq_encoder.0.bias"].cuda()
h1T = ws[nvt * 128 * 104: 2 * nvt * 128 * 104].view(nvt, 4, 104, 32)
for br, m in enumerate([mask, mask_p]):
    h1 = torch.relu((x * m.float()) @ W1.T + b1)
    rows = h1T[br * ntiles].permute(0, 2, 1).reshape(128, 104)
    print("h1 br", br, "tile 0 max diff", float((rows[:, :100] - h1[:128]).abs().max()), "col100", float(rows[:, 100].min()), float(rows[:, 100].max()))
