"""cfg4 step from CUDA graphs holding 1, 2, 4, 8 steps each (development probe: how much of the step is the gap between
two graph launches).  Everything the step needs is counted on the device, so a graph of U steps is U recorded steps."""
import os, sys, time
import torch
sys.path.insert(0, ".")
from vae_posterior_consistency_b200 import lib as L, VAE, kernels as KR

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
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(f"1 step per graph: {timed(tr.step_graph, 400):.4f} ms/step")
for U in (2, 4, 8):
    side = torch.cuda.Stream(device=dev, priority=-1)
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(U):
            tr._launch_step()
    print(f"{U} steps per graph: {timed(g.replay, 400 // U) / U:.4f} ms/step")

# how much of the step is the (forked) gather of the next batch: the same graph without it (not a valid training step)
tr._prep = lambda counter: None
side = torch.cuda.Stream(device=dev, priority=-1)
side.wait_stream(torch.cuda.current_stream())
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=side):
    tr._launch_step()
print(f"1 step per graph, no batch gather: {timed(g.replay, 400):.4f} ms/step")
tr.eng.build_weight_images = lambda theta, wimg: None
g2 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g2, stream=side):
    tr._launch_step()
print(f"1 step per graph, no batch gather, no image build: {timed(g2.replay, 400):.4f} ms/step")
