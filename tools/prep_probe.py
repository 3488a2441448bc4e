"""cfg4 step with the gather of the NEXT batch forked at the START of the step into a second set of batch buffers (two
graphs replayed alternately), against the kept order (gather beside reduce + Adam).  Development probe."""
import os, sys
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
    for _ in range(6):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(f"kept order: {timed(tr.step_graph, 400):.4f} ms/step")
where = os.environ.get("FORK_AT", "start")
bufs = [(tr.x, tr.mask, tr.mask_p, tr.eps),
        (torch.empty_like(tr.x), torch.empty_like(tr.mask), torch.empty_like(tr.mask_p), torch.empty_like(tr.eps))]


def launch(cur, nxt):
    main = torch.cuda.current_stream()
    x, m, mp, eps = bufs[cur]

    def fork_prep():
        tr._fork.wait_stream(main)
        with torch.cuda.stream(tr._fork):
            tr.x, tr.mask, tr.mask_p, tr.eps = bufs[nxt]
            tr._prep(tr.prep_count)
            tr.prep_count += 1
            tr.x, tr.mask, tr.mask_p, tr.eps = bufs[0]
    if where == "start":
        fork_prep()
    e = tr.eng
    masks, epsl = [m, mp], [eps[0], eps[1]]
    mean, logvar, z, ws = e.enc_fwd(tr.theta, x, masks, epsl, save=True, wimg=tr.wimg)
    if where == "enc":
        fork_prep()
    out = e.dec(L.DEC_TRAIN, tr.theta, z, x=x, masks=masks, mean=mean, logvar=logvar, eps=epsl, alpha=tr.alpha,
                beta_w=tr.beta_w, loss_scale=1.0 / B, wimg=tr.wimg)
    e.enc_bwd(tr.theta, x, masks, ws, out["d_mean"], out["d_logvar"], wimg=tr.wimg)
    tr._launch_tail()
    e.build_weight_images(tr.theta, tr.wimg)
    main.wait_stream(tr._fork)


side = torch.cuda.Stream(device=dev, priority=-1)
side.wait_stream(torch.cuda.current_stream())
graphs = []
with torch.cuda.stream(side):
    launch(0, 1); launch(1, 0)               # warm, leaves buffers 0 prepared
side.synchronize()
for cur in (0, 1):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        launch(cur, 1 - cur)
    graphs.append(g)
k = [0]


def step():
    graphs[k[0] & 1].replay()
    k[0] += 1


print(f"gather forked at {where}: {timed(step, 400):.4f} ms/step")
