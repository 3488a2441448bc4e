"""Timing probe: one fused training step (8 launches) replayed from a CUDA graph with frozen per-step scalars vs launched
eagerly, at the bench shape and at a cfg1-like small batch.  The frozen scalars make the replay numerically
meaningless; only the launch-gap difference is of interest."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pcvae_oracle as O
from vae_posterior_consistency_b200 import kernels as KR, lib as L
lib = L.load()
dev = torch.device("cuda")
for B, D, T in ((65536, 100, 1_000_000), (64, 20, 455), (64, 13, 455)):
    g = torch.Generator(device=dev).manual_seed(0)
    table = torch.rand(T, D, device=dev, generator=g); mtable = torch.rand(T, D, device=dev, generator=g) < 0.7
    p = O.init_params("mlp", D, seed=0)
    theta = KR.flatten_params(p, L.FAMILY_MLP, dev)
    tr = KR.FusedTrainer(L.FAMILY_MLP, D, 0, theta, regularised=True)
    idx = torch.randint(0, T, (B,), device=dev)
    x = torch.empty(B, D, device=dev); m = torch.empty(B, D, device=dev, dtype=torch.bool); mp = torch.empty_like(m)
    eps = torch.empty(2, B, 10, device=dev)
    def step():
        if D % 4 == 0:
            L.check(lib.pcvae_prep_batch(table.data_ptr(), mtable.data_ptr(), idx.data_ptr(), x.data_ptr(), m.data_ptr(), mp.data_ptr(),
                                         eps.data_ptr(), B, D, 2, 0.7, 99, 8, torch.cuda.current_stream().cuda_stream), "prep")
        else:
            L.check(lib.pcvae_gather_rows(table.data_ptr(), mtable.data_ptr(), idx.data_ptr(), x.data_ptr(), m.data_ptr(), B, D, L.MASK_U8,
                                          torch.cuda.current_stream().cuda_stream), "gather")
            L.check(lib.pcvae_draw_submask(m.data_ptr(), mp.data_ptr(), B * D, 0.7, 99, 8, torch.cuda.current_stream().cuda_stream), "sub")
            L.check(lib.pcvae_draw_normal(eps.data_ptr(), 2 * B * 10, 77, 8, torch.cuda.current_stream().cuda_stream), "norm")
        return tr.step(x, m, mp, eps[0], eps[1])
    for _ in range(5): step()
    torch.cuda.synchronize()
    n = 200
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    w0 = time.time(); t0.record()
    for _ in range(n): step()
    t1.record(); torch.cuda.synchronize(); w1 = time.time()
    eager = t0.elapsed_time(t1) / n
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): step()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s):
        loss = step()
    for _ in range(5): gr.replay()
    torch.cuda.synchronize()
    t0.record()
    for _ in range(n): gr.replay()
    t1.record(); torch.cuda.synchronize()
    print(f"B={B} D={D}: eager {eager*1e3:.1f} us/step (wall {(w1-w0)/n*1e6:.1f}), graph {t0.elapsed_time(t1)/n*1e3:.1f} us/step")
