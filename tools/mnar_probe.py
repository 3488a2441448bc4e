"""cfg2 REG_notMIWAE_v2 steps, eager and from the CUDA graph (development probe; bench.py's `mnar` section is the record).
Under `ncu --metrics gpu__time_duration.sum` the eager steps give the launch list of one step."""
import sys
import torch
sys.path.insert(0, ".")
from vae_posterior_consistency_b200 import VAE
from vae_posterior_consistency_b200.graphed import GraphedTrainer

dev = torch.device("cuda")
D, B, S, N = 50, 128, 20, 10_000
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
torch.manual_seed(0)
model = VAE.REG_notMIWAE_v2(D, 500, 20, 10, {"batch_size": B, "patience": 100}, S, 10).to(dev)
model.noise = "device"
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
g = torch.Generator(device=dev).manual_seed(9)
table = torch.rand(N, D, device=dev, generator=g)
mtable = torch.ones(N, D, device=dev)
mtable[:, :D // 2] = (table[:, :D // 2] <= table[:, :D // 2].mean(0)).float()


def batch():
    idx = torch.randint(0, N, (B,), device=dev, generator=g)
    x, mask = table[idx], mtable[idx]
    return x, mask, mask * (torch.rand(B, D, device=dev, generator=g) < 0.5).float()


def fwd_loss(x, mask, mask_p):
    mean_p, logvar_p, xm_p, xlv_p, mean_q, logvar_q, xm_q, xlv_q = model.forward(x, mask, mask_p, stage="train")
    return model.loss(x, xm_p, xlv_p, mean_p, logvar_p, xm_q, xlv_q, mean_q, logvar_q, mask, mask_p, 1, alpha=1.0,
                      stage="train")[1]


def eager_step():
    loss = fwd_loss(*batch())
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss


def timed(fn, n):
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n):
        out = fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n, out


for _ in range(3):
    eager_step()
torch.cuda.synchronize()
print("MARK eager")
ms, _ = timed(eager_step, 3)
print(f"eager {ms:.3f} ms/step")
if "--eager-only" not in sys.argv:
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True, fused=True)
    gt = GraphedTrainer(model, lambda: fwd_loss, opt, VAE.fill_normal_)
    for _ in range(5):
        gt.step(*batch())
    ms, loss = timed(lambda: gt.step(*batch()), steps)
    bs = [batch() for _ in range(8)]
    ms2, loss = timed(lambda: gt.step(*bs[0]), steps)
    print(f"graph {ms:.3f} ms/step (with batch assembly), {ms2:.3f} ms/step (graph replay only), loss {float(loss):.4f}")
