"""Time of the batch preparation alone (cfg4: 65 536 rows gathered from a 1 M-row table) and of the graphed step."""
import os, sys
import torch
sys.path.insert(0, ".")
from vae_posterior_consistency_b200 import lib as L
if os.environ.get("PCVAE_LIB"):
    L.LIB_PATH = os.path.abspath(os.environ["PCVAE_LIB"])
from vae_posterior_consistency_b200 import VAE, kernels as KR
B, D, T = 65536, 100, 1_000_000
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = VAE.Reg_VAE(D, 500, 0, 10, {"batch_size": 64, "patience": 100}, "probe", "kl_reg")
theta = model.flat_theta().detach().clone().cuda()
table = torch.rand(T, D, device=dev)
mtable = (torch.rand(T, D, device=dev) < 0.7)
nb = 15
tr = KR.GraphedFusedTrainer(L.FAMILY_MLP, D, 0, theta, table, mtable, B, nb, keep=0.7, seed=99, regularised=True)
tr.set_batches(torch.randperm(T, device=dev)[:nb * B].view(nb, B))
tr.capture(warmup=3)
def timed(fn, n):
    for _ in range(6): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def prep():
    tr._prep(tr.prep_count); tr.prep_count += 1
print(f"{os.environ.get('PCVAE_LIB', 'default')}: prep alone {1e3 * timed(prep, 300):.1f} us, step {timed(tr.step_graph, 300):.4f} ms, checksum {float(tr.x.double().sum()) + float(tr.eps.double().sum()):.6f}")
