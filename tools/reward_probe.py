"""Times pcvae_reward_chain at cfg5 size with each main kernel (development probe; bench.py is the record)."""
import sys, time
import torch
sys.path.insert(0, ".")
import os
from vae_posterior_consistency_b200 import lib as L
if os.environ.get("PCVAE_LIB"):
    L.LIB_PATH = os.path.abspath(os.environ["PCVAE_LIB"])
from vae_posterior_consistency_b200 import kernels as KR

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
modes = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2, 1]
D, M = 101, 50
torch.manual_seed(0)
from vae_posterior_consistency_b200 import VAE
theta = VAE.Reg_VAE(D, 500, 0, 10, {"batch_size": 64, "patience": 100}, "probe", "kl_reg").flat_theta().detach().clone().cuda()
eng = KR.Engine(L.FAMILY_MLP, D, 0)
x = torch.rand(rows, D, device="cuda")
mask = torch.zeros(rows, D, device="cuda")
im = torch.rand(M, rows, D, device="cuda")
lib = L.load()
out = {}
for mode in modes:
    lib.pcvae_set_reward_tensor_cores(mode)
    for _ in range(2):
        R, _ = eng.reward(theta, x, mask, im)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        R, _ = eng.reward(theta, x, mask, im)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    out[mode] = R.clone()
    print(f"mode {mode}: {ms:.2f} ms  {rows * 100 * M / ms / 1e6:.2f} G triples/s", flush=True)
if len(out) == 2:
    a, b = list(out.values())
    print("bit-identical:", torch.equal(a, b), "max |diff|", float((a - b).abs().max()), "scale", float(a.abs().max()))
