#!/usr/bin/env python
"""Benchmark of the partial-VAE posterior-consistency hot path (BASELINE.json metric:
train rows/s (fwd+bwd) and active-selection rewards/s).

    python bench.py --gpus N --steps K --warmup W              # this repo's sm_100a path
    python bench.py --impl reference --steps K --warmup W      # the UNMODIFIED reference (baseline/_ref) on the host cores

One JSON line on rank 0.  `value` = consistency-regularised training rows/s (Reg_VAE, cfg4: 1M x 100 synthetic table,
batch 65 536 rows per GPU, fwd + bwd + Adam, inputs resident in HBM); `secondary` = active-selection reward triples/s
(cfg5: 100k rows x 100 candidates x 50 samples, rows sharded over the GPUs); further sections: `al_loop` (cfg3 through
active_learning_func), `mnar` (cfg2), `pnp` (Reg_EDDI), `strong` (cfg4's global batch split over the GPUs), `long_run`
(the same step timed over 200 steps).  See DESIGN.md section "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D_TRAIN = 100
D_REWARD = 101
# algorithmic FLOP per unit (FLOP = 2*MAC, transcendental work excluded)
FLOP_TRAIN_ROW = 338_000          # Reg_VAE D=100 fwd+bwd, both branches
FLOP_PNP_ROW = 306_000            # Reg_EDDI K=20 D=100, collapsed form
FLOP_MNAR_ROW = 7_590_000         # REG_notMIWAE_v2, D=50, S=20, fwd+bwd both branches
# per-kernel algorithmic FLOP per branch-row at D=100 (they add up to FLOP_TRAIN_ROW / 2 = 169 000)
KERNEL_FLOP_BRANCH_ROW = {
    "k_enc_fwd_tc": 32_000,        # 2 x (100x100 + 100x50 + 50x20)
    "k_dec_fwd_tc": 31_000,        # 2 x (10x50 + 50x100 + 100x100)
    "k_dec_bwd_tc": 31_000,        # data gradients of the three decoder layers
    "k_wgrad_tc[dec]": 31_000,     # weight gradients of the three decoder layers
    "k_enc_bwd_tc": 12_000,        # data gradients of encoder layers 3 and 2
    "k_wgrad_tc[enc]": 32_000,     # weight gradients of the three encoder layers
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch at batch 65 536 from the ncu --set full captures
# (profiles/r01_ncu_summary.md capture 6; unchanged kernels)
KERNEL_DRAM_BYTES = {"k_enc_fwd_tc": 142.1e6, "k_dec_fwd_tc": 136.6e6, "k_dec_bwd_tc": 118.7e6, "k_wgrad_tc[dec]": 235.4e6,
                     "k_enc_bwd_tc": 49.5e6, "k_wgrad_tc[enc]": 238.9e6}
FLOP_REWARD_TRIPLE = 24_460       # Reg_VAE incremental form: 2 tail evaluations + amortised bases
# dram bytes of the reward main kernel at cfg5 (100k rows x 100 candidates x 50 samples), profiles/r02_ncu_summary.md
# (same reads in the lock-step and the warp-specialised kernel: imputations and base posteriors, once each)
REWARD_DRAM_BYTES = 2.993e9       # read 2.945 GB (imputations 2.02 GB + per-(row, sample) base posteriors) + write 0.048 GB

#: the workload both arms are run on (identical `config` in the two JSON lines)
def config_dict(args):
    return {"workload": f"cfg4: Reg_VAE (consistency-regularised partial VAE, kl_reg, alpha 1), obs_dim {D_TRAIN}, synthetic "
                        f"{args.table_rows} x {D_TRAIN} table U(0,1), MCAR 30 % mask, sub-mask 30 %, batch {args.batch} rows "
                        "per step (per GPU), one step = forward + loss + backward + Adam",
            "batch": args.batch, "table_rows": args.table_rows, "obs_dim": D_TRAIN,
            "l2": "inputs larger than L2: a fresh batch (33 MB) gathered from the 500 MB table every step, ~0.9 GB of "
                  "scratch traffic per step"}


def tensor_peak_3xtf32():
    """Tensor-core roofline of an fp32-accurate product issued as three kind::tf32 MMAs: the measured dense bf16
    rate / 2 (tf32 runs at half the bf16 rate) / 3."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]) / 6.0, "MEASURED_PEAKS.json bf16_tflops (burst) / 2 (tf32) / 3 (3xTF32)"
    except Exception:
        return 1590.0 / 6.0, "fallback 1.59 PFLOP/s bf16 / 2 (tf32) / 3 (3xTF32)"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi sampling in the background for the whole run; only the samples whose timestamp falls
    inside a timed window (train / e2e / reward loops) are summarised."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.windows = [], None, index, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        if self.proc is not None:
            time.sleep(0.12)
            self.proc.terminate()
        sm, mx, reasons, power = [], 0, set(), 0.0
        for ts, r in self.rows:
            if not any(a - 0.05 <= ts <= b + 0.1 for a, b in self.windows):
                continue
            f = [s.strip() for s in r.split(",")]
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2])); power = max(power, float(f[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": power or None}


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return world, rank, local


def max_over_ranks(ms, world):
    if world == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


# ----------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference (baseline/_ref, installed by baseline/install_ref.py) on the
# host cores; the oracle port (oracle/pcvae_oracle.py) when the install is absent, and as a second number beside it
# ----------------------------------------------------------------------------------------

def load_reference():
    """(src.models.VAE, src.experiment_main.evaluate, src.utils.utils) of the unmodified reference, or None."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "src", "models")):
        return None
    import types
    try:
        import matplotlib.pyplot  # noqa: F401  (imported, unused, at the reference's evaluate.py:10)
    except Exception:
        mpl = types.ModuleType("matplotlib")
        mpl.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules.setdefault("matplotlib", mpl)
        sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import src.models.VAE as V
    import src.experiment_main.evaluate as E
    import src.utils.utils as U
    return V, E, U


def ref_train_rows_per_s(ref, batch, steps, warmup):
    """One training step as the reference's train.py:49-116 runs it for 'reg_vae': sub-mask (utils.py:36-39), forward,
    loss, zero_grad, backward, Adam.step -- its own classes, torch CPU eager, all host threads."""
    V, _, U = ref
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model = V.Reg_VAE(D_TRAIN, 500, 20, 10, {"batch_size": batch, "patience": 100}, "bench", "kl_reg")
    opt = torch.optim.Adam(model.parameters(), lr=0.001)                    # train.py:21
    g = torch.Generator().manual_seed(1)
    x = torch.rand(batch, D_TRAIN, generator=g)
    mask = torch.rand(batch, D_TRAIN, generator=g) < 0.7
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        mask_p = U.create_missing_uci(x.shape, 30) * mask                  # train.py:53-55
        mean_p, logvar_p, xm_p, xlv_p, mean_q, logvar_q, xm_q, xlv_q = model.forward(x, mask, mask_p, stage="train")
        _, train_loss = model.loss(x, xm_p, xlv_p, mean_p, logvar_p, xm_q, xlv_q, mean_q, logvar_q, mask, mask_p, s + 1,
                                   beta_annealing=False, beta=1.0, alpha=1.0, alpha_annealing=True, stage="train")
        opt.zero_grad()
        train_loss.backward()
        opt.step()
        times.append(time.perf_counter() - t0)
    t = sum(times[warmup:]) / steps
    return batch / t, t * 1e3


def ref_reward_triples_per_s(ref, n_rows, n_cand, samples):
    """R_lindley_chain of the reference (evaluate.py:514-634) for `n_cand` of the candidates on `n_rows` rows."""
    V, E, _ = ref
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model = V.Reg_VAE(D_REWARD, 500, 20, 10, {"batch_size": 64, "patience": 100}, "bench", "kl_reg")
    g = torch.Generator().manual_seed(2)
    x = torch.rand(n_rows, D_REWARD, generator=g)
    mask = torch.zeros(n_rows, D_REWARD)
    im = torch.rand(samples, n_rows, D_REWARD, generator=g)
    loc = list(range(n_rows))
    t0 = time.perf_counter()
    with torch.no_grad():
        for u in range(n_cand):
            E.R_lindley_chain(u, x, mask, samples, model, im, loc)
    dt = time.perf_counter() - t0
    return n_rows * n_cand * samples / dt, dt * 1e3


def port_train_rows_per_s(batch, steps, warmup):
    from oracle import pcvae_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    p = O.init_params("mlp", D_TRAIN, seed=0)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(batch, D_TRAIN, generator=g)
    mask = torch.rand(batch, D_TRAIN, generator=g) < 0.7
    names = O.trainable_names(p)
    m = {k: torch.zeros_like(p[k]) for k in names}
    v = {k: torch.zeros_like(p[k]) for k in names}
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        mask_p = mask & (torch.rand(batch, D_TRAIN) < 0.7)             # train.py:53-55
        eq, ep = torch.randn(batch, 10), torch.randn(batch, 10)        # the two rsample() draws
        _, grads, _ = O.train_step(p, x, mask, mask_p, eq, ep)
        for k in names:
            p[k], m[k], v[k] = O.adam_step(p[k], grads[k], m[k], v[k], s + 1)
        times.append(time.perf_counter() - t0)
    t = sum(times[warmup:]) / steps
    return batch / t, t * 1e3


def port_reward_triples_per_s(n_rows, n_cand, samples):
    from oracle import pcvae_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    p = O.init_params("mlp", D_REWARD, seed=0)
    g = torch.Generator().manual_seed(2)
    x = torch.rand(n_rows, D_REWARD, generator=g)
    mask = torch.zeros(n_rows, D_REWARD)
    im = torch.rand(samples, n_rows, D_REWARD, generator=g)
    loc = torch.arange(n_rows)
    t0 = time.perf_counter()
    with torch.no_grad():
        for u in range(n_cand):
            O.reward_chain_as_written(p, u, x, mask, im, loc)       # evaluate.py:514-634, as written
    dt = time.perf_counter() - t0
    return n_rows * n_cand * samples / dt, dt * 1e3


def port_mnar_rows_per_s(batch, samples, steps, warmup):
    from oracle import pcvae_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    D = 50
    p = O.init_mnar_params(D, seed=0)
    g = torch.Generator().manual_seed(4)
    x = torch.rand(batch, D, generator=g)
    mask = (torch.rand(batch, D, generator=g) < 0.7).float()
    names = O.mnar_trainable_names(p)
    m = {k: torch.zeros_like(p[k]) for k in names}
    v = {k: torch.zeros_like(p[k]) for k in names}
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        mask_p = mask * (torch.rand(batch, D) < 0.5).float()
        eq, ep = torch.randn(batch, samples, 10), torch.randn(batch, samples, 10)
        _, grads, _ = O.mnar_train_step(p, x, mask, mask_p, eq, ep)
        for k in names:
            p[k], m[k], v[k] = O.adam_step(p[k], grads[k], m[k], v[k], s + 1)
        times.append(time.perf_counter() - t0)
    t = sum(times[warmup:]) / steps
    return batch / t, t * 1e3


def cpu_train_baseline(batch, steps, warmup):
    """cpu_baseline object of the training metric: the unmodified reference when installed, else the oracle port."""
    cores = os.cpu_count() or 1
    ref = load_reference()
    if ref is not None:
        v, ms = ref_train_rows_per_s(ref, batch, steps, warmup)
        return {"value": v, "unit": "rows/s", "cores": cores, "kind": "reference", "ms_per_step": ms,
                "sample": f"{steps} steps of one {batch}-row batch through the unmodified reference classes (baseline/_ref: "
                          "Reg_VAE.forward + loss + backward + Adam as train.py:49-116), torch CPU eager, "
                          f"{torch.get_num_threads()} threads"}
    v, ms = port_train_rows_per_s(batch, steps, warmup)
    return {"value": v, "unit": "rows/s", "cores": cores, "kind": "port", "ms_per_step": ms,
            "sample": f"{steps} steps of one {batch}-row batch, oracle port of train.py:87-116 (baseline/_ref not installed)"}


def cpu_reward_baseline(n_rows, n_cand, samples):
    cores = os.cpu_count() or 1
    ref = load_reference()
    if ref is not None:
        v, ms = ref_reward_triples_per_s(ref, n_rows, n_cand, samples)
        kind = "reference"
    else:
        v, ms = port_reward_triples_per_s(n_rows, n_cand, samples)
        kind = "port"
    return {"value": v, "unit": "triples/s", "cores": cores, "kind": kind, "ms": ms,
            "sample": f"R_lindley_chain as written (evaluate.py:514-634): {n_cand} of {D_REWARD - 1} candidates on {n_rows} rows x "
                      f"{samples} samples, D={D_REWARD}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_train_baseline(args.batch, args.steps, args.warmup)
    rw = cpu_reward_baseline(args.ref_reward_rows, 2, args.reward_samples)
    line = {
        "impl": "reference", "metric": "train rows/s (fwd+bwd+Adam), consistency-regularised partial VAE",
        "value": cb["value"], "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(args), "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "secondary": {"metric": "active-selection rewards/s", "value": rw["value"], "unit": "triples/s", "cpu_baseline": rw},
    }
    if cb["kind"] == "reference" and not args.no_port:
        pv, pms = port_train_rows_per_s(args.batch, min(args.steps, 5), 1)
        line["port"] = {"value": pv, "unit": "rows/s", "ms_per_step": pms,
                        "note": "the oracle port (oracle/pcvae_oracle.py, closed-form restatement: fewer ATen ops than the "
                                "reference) on the same batch, for comparison with round 1's reference arm"}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------

def fresh_theta(family, D, K, dev, seed=0):
    """Flat parameter vector of a freshly initialised mirror model (nn.Linear default init, as the reference's)."""
    from vae_posterior_consistency_b200 import VAE, lib as L
    torch.manual_seed(seed)
    tp = {"batch_size": 64, "patience": 100}
    if family == L.FAMILY_PNP:
        model = VAE.Reg_EDDI(D, 500, K, 10, tp, "bench", "kl_reg")
    else:
        model = VAE.Reg_VAE(D, 500, K, 10, tp, "bench", "kl_reg")
    return model.flat_theta().detach().clone().to(dev), model


def timed_graph_steps(tr, n, world, clocks=None):
    barrier(world)
    w0 = time.time()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n):
        sums = tr.step_graph()
    t1.record()
    barrier(world)
    if clocks is not None:
        clocks.window(w0, time.time())
    return max_over_ranks(t0.elapsed_time(t1), world) / n, sums


def gpu_mnar(args, dev, ffma_tflops, with_cpu):
    """cfg2: REG_notMIWAE_v2 on a synthetic 10 000 x 50 MNAR table, batch 128, S = train_k = 20; a step is
    model.forward + model.loss + backward (pcvae:: custom ops) + torch.optim.Adam, as train.py:87-116 runs it."""
    from vae_posterior_consistency_b200 import VAE
    D, B, S, N = 50, 128, 20, 10_000
    torch.manual_seed(0)
    model = VAE.REG_notMIWAE_v2(D, 500, 20, 10, {"batch_size": B, "patience": 100}, S, 10).to(dev)
    model.noise = "device"
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    g = torch.Generator(device=dev).manual_seed(9)
    table = torch.rand(N, D, device=dev, generator=g)
    mtable = torch.ones(N, D, device=dev)
    mtable[:, :D // 2] = (table[:, :D // 2] <= table[:, :D // 2].mean(0)).float()     # self-masking MNAR
    steps, warm = args.mnar_steps, 5

    def batch():
        idx = torch.randint(0, N, (B,), device=dev, generator=g)
        x, mask = table[idx], mtable[idx]
        return x, mask, mask * (torch.rand(B, D, device=dev, generator=g) < 0.5).float()

    def fwd_loss(x, mask, mask_p):
        mean_p, logvar_p, xm_p, xlv_p, mean_q, logvar_q, xm_q, xlv_q = model.forward(x, mask, mask_p, stage="train")
        return model.loss(x, xm_p, xlv_p, mean_p, logvar_p, xm_q, xlv_q, mean_q, logvar_q, mask, mask_p, 1,
                          alpha=1.0, stage="train")[1]

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

    for _ in range(warm):
        eager_step()
    eager_ms, _ = timed(eager_step, max(steps // 4, 10))
    # the same step replayed from a CUDA graph (train.py does this under PCVAE_MODE=throughput): batch assembly,
    # sub-mask and noise draws stay outside the graph, forward + loss + backward + Adam are one graph launch
    from vae_posterior_consistency_b200.graphed import GraphedTrainer
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True, fused=True)      # as train() builds it for the graph
    gt = GraphedTrainer(model, lambda: fwd_loss, opt, VAE.fill_normal_)
    for _ in range(warm):
        gt.step(*batch())
    ms, loss = timed(lambda: gt.step(*batch()), steps)
    assert gt.replays >= steps
    # e2e: the batch (x, mask, sub-mask) arrives from pinned host memory every step, the loss goes back
    hb = [tuple(t.cpu().pin_memory() for t in batch()) for _ in range(4)]
    hloss = torch.empty(steps + warm).pin_memory()
    it = [0]

    def e2e_step():
        i = it[0]
        it[0] += 1
        x, m, mp = (t.to(dev, non_blocking=True) for t in hb[i % 4])
        out = gt.step(x, m, mp)
        hloss[i].copy_(out, non_blocking=True)
        return out
    for _ in range(warm):
        e2e_step()
    e2e_ms, _ = timed(e2e_step, steps)
    res = {"metric": "MNAR train rows/s (REG_notMIWAE_v2, fwd+bwd+Adam)", "value": B / (ms * 1e-3), "unit": "rows/s",
           "ms_per_step": ms, "steps": steps, "eager_ms_per_step": eager_ms,
           "config": {"workload": f"cfg2: synthetic {N} x {D} self-masking MNAR table, batch {B}, train_k {S}, "
                                  "module API + autograd over pcvae:: dense / mnar ops + Adam replayed from a CUDA graph, "
                                  "device noise (eager_ms_per_step = the same step launched op by op)"},
           "e2e": {"value": B / (e2e_ms * 1e-3), "unit": "rows/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": 3 * B * D * 4, "d2h_bytes_per_step": 4},
           "roofline": {"bound": "FP32 FFMA (2 x 2 560 virtual rows per step, ~60 kernels of a few microseconds each: launch- "
                                 "and latency-bound)",
                        "achieved": FLOP_MNAR_ROW * B / (ms * 1e-3) / 1e12, "peak": ffma_tflops, "unit": "TFLOP/s",
                        "frac": FLOP_MNAR_ROW * B / (ms * 1e-3) / 1e12 / ffma_tflops},
           "final_loss": float(loss.detach())}
    if with_cpu:
        cv, cms = port_mnar_rows_per_s(B, S, 5, 1)
        res["cpu_baseline"] = {"value": cv, "unit": "rows/s", "cores": os.cpu_count() or 1, "kind": "port",
                               "sample": f"5 steps, batch {B}, S={S}, oracle port of REG_notMIWAE_v2 step + Adam",
                               "ms_per_step": cms}
    return res


def gpu_pnp_train(dev, ffma_tflops, batch, table, mtable, steps=40):
    """The same step for the PNP/EDDI set-encoder family (Reg_EDDI, K=20, D=100) through the gather + graph path."""
    from vae_posterior_consistency_b200 import kernels as KR, lib as L
    D, K = D_TRAIN, 20
    theta, _ = fresh_theta(L.FAMILY_PNP, D, K, dev)
    n_total = steps + 8
    T = table.shape[0]
    g = torch.Generator(device=dev).manual_seed(77)
    perm = torch.cat([torch.randperm(T, device=dev, generator=g) for _ in range((n_total * batch + T - 1) // T + 1)])
    tr = KR.GraphedFusedTrainer(L.FAMILY_PNP, D, K, theta, table, mtable, batch, n_total, keep=0.7, seed=5, regularised=True,
                                alpha=1.0)
    tr.set_batches(perm[:n_total * batch].view(n_total, batch))
    tr.capture(warmup=3)
    for _ in range(5):
        tr.step_graph()
    ms, sums = timed_graph_steps(tr, steps, 1)
    loss = float(KR.loss_from_sums(sums, batch, 1.0, 1.0, True))
    return {"metric": "train rows/s (Reg_EDDI K=20, gather + fused step replayed from a CUDA graph)", "value": batch / (ms * 1e-3),
            "unit": "rows/s", "ms_per_step": ms, "steps": steps,
            "roofline": {"bound": "fp32_ffma (set-encoder embedding and its backward) + tensor (decoder, weight gradients)",
                         "achieved": FLOP_PNP_ROW * batch / (ms * 1e-3) / 1e12, "peak": ffma_tflops,
                         "unit": "TFLOP/s", "frac": FLOP_PNP_ROW * batch / (ms * 1e-3) / 1e12 / ffma_tflops},
            "final_loss": loss}


def gpu_al_loop(args, dev, world, rank):
    """cfg3: active_learning_func (the acquisition loop of active_learning.py) on a synthetic 2 000 x 20 test set, M = 50,
    all 19 steps, rows sharded over the ranks; device noise (PCVAE_MODE=throughput).  rewards/s counts the (row,
    candidate, sample) triples of the reward calls."""
    from vae_posterior_consistency_b200 import VAE, evaluate, loaders
    N, D, M = args.al_rows, 20, 50
    exp, data_type, vae_type = "bench_al", "synth", "reg_vae1"
    tp = {"batch_size": 64, "patience": 100}
    root = tempfile.mkdtemp(prefix="pcvae_bench_al_") if rank == 0 else None
    if world > 1:
        import torch.distributed as dist
        box = [root]
        dist.broadcast_object_list(box, src=0)
        root = box[0]
    cwd = os.getcwd()
    prev = os.environ.get("PCVAE_MODE")
    try:
        if rank == 0:
            torch.manual_seed(3)
            model = VAE.Reg_VAE(D, 500, 10, 10, tp, exp, "kl_reg")
            path = os.path.join(root, loaders.checkpoint_path(exp, data_type, vae_type, 30, 1.0, 30, "kl_reg", "reg_vae"))
            os.makedirs(os.path.dirname(path), exist_ok=True)
            torch.save(model.state_dict(), path)
        barrier(world)
        os.chdir(root)
        os.environ["PCVAE_MODE"] = "throughput"
        g = torch.Generator().manual_seed(4)
        test = torch.rand(N, D, generator=g)
        tmask = torch.rand(N, D, generator=g) < 0.7
        out = {}
        import contextlib
        for rep in range(2):                                              # first pass warms up (lazy module load, allocator)
            barrier(world)
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(sys.stderr):                  # the loop prints its progress like the reference's
                evaluate.active_learning_func(None, test, tmask, 30, D, 500, 10, M, 10, data_type, tp, exp, vae_type, 1, 5000,
                                              10, device=dev, alpha=1.0, p_missingness=30, reg_type="kl_reg", Repeat=1)
            barrier(world)
            wall = time.perf_counter() - t0
            tm = dict(evaluate.LAST_TIMING)
            out = {"wall_ms": max_over_ranks(wall * 1e3, world), "reward_ms": max_over_ranks(tm["reward_ms"], world),
                   "calls": tm["reward_calls"]}
        triples = N * M * (D - 1) * D // 2                                  # N M (C + (C-1) + ... + 1), C = D - 1
        return {"metric": "active-selection rewards/s through active_learning_func (cfg3)", "unit": "triples/s",
                "value": triples / (out["reward_ms"] * 1e-3), "reward_ms": out["reward_ms"], "reward_calls": out["calls"],
                "whole_loop": {"value": triples / (out["wall_ms"] * 1e-3), "unit": "triples/s", "wall_ms": out["wall_ms"],
                               "note": "wall clock of the whole call: 2 x 20 x M decoder passes for the imputations, 19 reward "
                                       "calls, argmax / mask updates, histories copied to the host and saved (im_CHAI: "
                                       f"{19 * M * N * D * 4 / 1e6:.0f} MB)"},
                "config": {"workload": f"cfg3: active_learning_func, synthetic {N} x {D} test set, M = {M}, {D - 1} acquisition "
                                       f"steps, rows sharded over {world} GPU(s), device noise"},
                "scaling": "strong", "triples": triples}
    finally:
        os.chdir(cwd)
        if prev is None:
            os.environ.pop("PCVAE_MODE", None)
        else:
            os.environ["PCVAE_MODE"] = prev


def run_ours(args):
    import ctypes as C
    world, rank, local = dist_setup(args.gpus)
    from vae_posterior_consistency_b200 import kernels as KR, lib as L

    dev = torch.device("cuda", local)
    clocks = ClockSampler(local)
    clocks.start()
    lib = L.load()
    stream = lambda: torch.cuda.current_stream().cuda_stream
    hbm_peak, peak_src = peaks()
    tensor_peak, tensor_src = tensor_peak_3xtf32()

    # ---- FP32 FFMA peak of this box (roofline denominator of the FFMA-bound kernels) ----
    scratch = torch.empty(148 * 4 * 512 * 2, device=dev)
    ffma_tflops = 0.0
    for _ in range(5):
        fl = C.c_double(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.pcvae_ffma_probe(scratch.data_ptr(), 4000, C.byref(fl), stream()), "ffma_probe")
        e1.record()
        torch.cuda.synchronize()
        ffma_tflops = max(ffma_tflops, fl.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)

    # =========================== training (cfg4) ===========================
    B, D, T = args.batch, D_TRAIN, args.table_rows
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    table = torch.rand(T, D, device=dev, generator=g)
    mtable = torch.rand(T, D, device=dev, generator=g) < 0.7                 # MCAR 30 % missing, bool
    theta, _ = fresh_theta(L.FAMILY_MLP, D, 20, dev)
    dist_group = None
    if world > 1:
        import torch.distributed as dist
        dist_group = dist.group.WORLD
        dist.broadcast(theta, src=0)                     # identical replicas (train() does the same, train.py:_sync_replicas)
    n_long = args.long_steps if world == 1 else 0
    n_total = args.warmup + args.steps + n_long
    perm = torch.cat([torch.randperm(T, device=dev, generator=g) for _ in range((n_total * B + T - 1) // T + 1)])
    graphed = os.environ.get("PCVAE_GRAPH", "1") != "0" and (world == 1 or os.environ.get("PCVAE_DP", "peer") != "nccl")
    tr = None
    if graphed:
        # the whole step (gather + sub-mask + noise, six training kernels, reduce [+ NVLink gradient exchange] + Adam) is
        # replayed from a CUDA graph; the per-step scalars live in a device counter (KR.GraphedFusedTrainer)
        try:
            tr = KR.GraphedFusedTrainer(L.FAMILY_MLP, D, 0, theta, table, mtable, B, n_total, keep=0.7, seed=99, regularised=True,
                                        alpha=1.0, dist_group=dist_group, world_size=world, global_rows=B * world)
            tr.set_batches(perm[:n_total * B].view(n_total, B))
        except L.PcvaeError:                             # no peer access between the GPUs (every rank agrees, see
            graphed = False                              # PeerExchange.create_or_none): eager launches + NCCL all-reduce
    if not graphed:
        tr = KR.FusedTrainer(L.FAMILY_MLP, D, 0, theta, regularised=True, alpha=1.0, dist_group=dist_group,
                             world_size=world)
    eng = tr.eng
    x = torch.empty(B, D, device=dev)
    mask = torch.empty(B, D, device=dev, dtype=torch.bool)
    mask_p = torch.empty(B, D, device=dev, dtype=torch.bool)
    eps = torch.empty(2, B, 10, device=dev)
    launches = [0]
    # per-kernel CUDA events recorded by the library itself (pcvae_profile_events): 9 marks per step
    # enc_fwd: [0] k [1]   dec: [2] k_fwd [3] k_bwd [4] k_wgrad [5]   enc_bwd: [6] k_bwd [7] k_wgrad [8]
    KERNELS = [("k_enc_fwd_tc", 0, 1), ("k_dec_fwd_tc", 2, 3), ("k_dec_bwd_tc", 3, 4), ("k_wgrad_tc[dec]", 4, 5),
               ("k_enc_bwd_tc", 6, 7), ("k_wgrad_tc[enc]", 7, 8)]
    ev_sets = []

    def new_events():
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
        for e in evs:
            e.record()                                   # creates the underlying cudaEvent_t
        return evs, (C.c_void_p * 9)(*[e.cuda_event for e in evs])

    def prep_step(s, xs, ms):
        """device-side batch preparation in one launch: gather by the sampler permutation, draw sub-mask and noise"""
        idx = perm[s * B:(s + 1) * B]
        L.check(lib.pcvae_prep_batch(table.data_ptr(), mtable.data_ptr(), idx.data_ptr(), xs.data_ptr(), ms.data_ptr(),
                                     mask_p.data_ptr(), eps.data_ptr(), B, D, 2, 0.7, 99, s * 4, stream()), "prep_batch")
        launches[0] += 1

    def train_step(xs, ms, timed):
        masks, e = [ms, mask_p], [eps[0], eps[1]]
        if timed:
            evs, arr = new_events()
            lib.pcvae_profile_events(arr, 9)
        wimg = getattr(tr, "wimg", None)                 # the weight images of this step's theta, as the trainers build them
        if wimg is not None:
            eng.build_weight_images(theta, wimg)
            launches[0] += 1
        mean, logvar, z, ws = eng.enc_fwd(theta, xs, masks, e, save=True, wimg=wimg)
        out = eng.dec(L.DEC_TRAIN, theta, z, x=xs, masks=masks, mean=mean, logvar=logvar, eps=e, alpha=1.0,
                      beta_w=1.0, loss_scale=1.0 / (B * world), wimg=wimg)
        eng.enc_bwd(theta, xs, masks, ws, out["d_mean"], out["d_logvar"], wimg=wimg)
        if timed:
            if lib.pcvae_profile_events(None, 0) == 9:   # all nine marks were recorded (tensor-core path taken)
                ev_sets.append(evs)
        tr.step_count += 1
        if world > 1 and tr.xch is not None:             # reduce + NVLink peer exchange + Adam in one launch
            sums = eng.dp_reduce_adam(tr.xch, tr.grad, theta, tr.exp_avg, tr.exp_avg_sq, tr.step_count, B)
            launches[0] += 6 + 1
        elif world > 1:                                  # PCVAE_DP=nccl: reduce, NCCL all-reduce, Adam
            eng.reduce_grads(tr.grad)
            sums = eng.reduce_sums(B)
            torch.distributed.all_reduce(tr.grad, group=dist_group)
            eng.adam_step(theta, tr.grad, tr.exp_avg, tr.exp_avg_sq, tr.step_count)
            launches[0] += 6 + 3
        else:
            sums = eng.reduce_adam(tr.grad, theta, tr.exp_avg, tr.exp_avg_sq, tr.step_count, B)
            launches[0] += 6 + 1
        return sums

    long_run = None
    if graphed:
        w_graph = min(3, args.warmup)
        tr.capture(warmup=w_graph)                       # `w_graph` eager warm-up steps, then the capture
        for s in range(args.warmup - w_graph):
            tr.step_graph()
        train_ms, sums = timed_graph_steps(tr, args.steps, world, clocks)
        # the kernels of the captured step, per replay: prep, [weight images], six training kernels, reduce + Adam
        train_launches = (9 if tr.wimg is not None else 8) * args.steps
        sums = sums.clone()
        if n_long > 0:                                   # the same step over a longer window (the contract's K can be small)
            long_ms, _ = timed_graph_steps(tr, n_long, world, clocks)
            long_run = {"steps": n_long, "ms_per_step": long_ms, "value": B * world / (long_ms * 1e-3), "unit": "rows/s"}
        # per-kernel times: the same step launched kernel by kernel, the library's own events between its launches
        if world == 1:
            for s in range(24):
                prep_step(s, x, mask); train_step(x, mask, s % 4 == 0)
            torch.cuda.synchronize()
    else:
        for s in range(args.warmup):
            prep_step(s, x, mask); train_step(x, mask, False)
        barrier(world)
        launches[0] = 0
        w0 = time.time()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for s in range(args.warmup, args.warmup + args.steps):
            prep_step(s, x, mask); sums = train_step(x, mask, (s - args.warmup) % 4 == 0)
        t1.record()
        barrier(world)
        clocks.window(w0, time.time())
        train_ms = max_over_ranks(t0.elapsed_time(t1), world) / args.steps
        train_launches = launches[0]
    kernel_us = {}
    for name, i0, i1 in KERNELS:
        if ev_sets:
            kernel_us[name] = 1e3 * sum(evs[i0].elapsed_time(evs[i1]) for evs in ev_sets) / len(ev_sets)
    loss = float(KR.loss_from_sums(sums, B * world, 1.0, 1.0, True))
    rows_s = B * world / (train_ms * 1e-3)

    # ---- strong scaling of cfg4: the global batch of 65 536 rows split over the GPUs ----
    strong = None
    if world > 1 and graphed and B % world == 0:
        Bs = B // world
        ns = args.warmup + args.steps
        th2, _ = fresh_theta(L.FAMILY_MLP, D, 20, dev)
        torch.distributed.broadcast(th2, src=0)
        try:
            ts = KR.GraphedFusedTrainer(L.FAMILY_MLP, D, 0, th2, table, mtable, Bs, ns, keep=0.7, seed=98, regularised=True,
                                        alpha=1.0, dist_group=dist_group, world_size=world, global_rows=B)
            ts.set_batches(perm[:ns * Bs].view(ns, Bs))
            ts.capture(warmup=min(3, args.warmup))
            for s in range(args.warmup - min(3, args.warmup)):
                ts.step_graph()
            s_ms, _ = timed_graph_steps(ts, args.steps, world, clocks)
            strong = {"metric": "train rows/s, global batch fixed", "value": B / (s_ms * 1e-3), "unit": "rows/s", "ms_per_step": s_ms,
                      "scaling": "strong", "global_batch": B, "batch_per_gpu": Bs, "n_gpus": world}
            ts.xch.close()
        except L.PcvaeError as e:
            strong = {"unavailable": str(e)}

    # ---- e2e: host batch (pinned) -> H2D -> step -> D2H loss sums, copies inside the timed region ----
    # The host batch is what the loader keeps for a table that lives in host memory: x dense fp32 and the mask
    # bit-packed (KR.pack_mask_bits, built once per table like the loader's min-max pass).  `observed`: the loader's
    # observed-entry form (KR.compact_rows), x's entries under mask == 0 are not shipped (they never reach the loss).
    nbuf = 2
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    hsum = torch.empty(args.steps + args.warmup, L.NSUMS, dtype=torch.float64).pin_memory()
    dx = [torch.empty(B, D, device=dev) for _ in range(nbuf)]
    dm = [torch.empty(B, D, device=dev, dtype=torch.bool) for _ in range(nbuf)]
    Wb = (D + 31) // 32
    dbits = [torch.empty(B, Wb, device=dev, dtype=torch.int32) for _ in range(nbuf)]
    dvals = [torch.empty(B * D, device=dev) for _ in range(nbuf)]
    doff = [torch.empty(B, device=dev, dtype=torch.int32) for _ in range(nbuf)]
    hx, hbits, hvals, hoff = [], [], [], []
    for _ in range(nbuf):
        xb_, mb_ = torch.rand(B, D), torch.rand(B, D) < 0.7
        v_, o_, b_ = KR.compact_rows(xb_, mb_)
        hx.append(xb_.pin_memory()); hbits.append(b_.pin_memory()); hvals.append(v_.pin_memory()); hoff.append(o_.pin_memory())

    def run_e2e(observed):
        ready = [torch.cuda.Event() for _ in range(nbuf)]
        freed = [torch.cuda.Event() for _ in range(nbuf)]
        h2d_bytes = [0]

        def loop(first, count):
            for i in range(first, first + count):
                b = i % nbuf
                with torch.cuda.stream(copy_stream):
                    if i >= first + nbuf:
                        copy_stream.wait_event(freed[b])
                    dbits[b].copy_(hbits[b], non_blocking=True)
                    n = hbits[b].numel() * 4
                    if observed:
                        nv = hvals[b].numel()
                        dvals[b][:nv].copy_(hvals[b], non_blocking=True)
                        doff[b].copy_(hoff[b], non_blocking=True)
                        n += nv * 4 + hoff[b].numel() * 4
                    else:
                        dx[b].copy_(hx[b], non_blocking=True)
                        n += hx[b].numel() * 4
                    h2d_bytes[0] += n
                    ready[b].record(copy_stream)
                main.wait_event(ready[b])
                if observed:
                    eng.prep_packed(dbits[b], dm[b], mask_p, eps, vals=dvals[b], row_off=doff[b], x=dx[b], keep=0.7, seed=99,
                                    offset=i * 4)
                else:
                    eng.prep_packed(dbits[b], dm[b], mask_p, eps, keep=0.7, seed=99, offset=i * 4)
                launches[0] += 1
                s_ = train_step(dx[b], dm[b], False)
                freed[b].record(main)
                hsum[i].copy_(s_, non_blocking=True)

        loop(0, args.warmup)
        barrier(world)
        h2d_bytes[0] = 0
        w0 = time.time()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        loop(args.warmup, args.steps)
        t1.record()
        barrier(world)
        clocks.window(w0, time.time())
        ms = max_over_ranks(t0.elapsed_time(t1), world) / args.steps
        return ms, h2d_bytes[0] // args.steps

    e2e_ms, h2d = run_e2e(False)
    e2e_rows_s = B * world / (e2e_ms * 1e-3)
    obs_ms, obs_h2d = run_e2e(True)
    e2e_observed = {"value": B * world / (obs_ms * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": obs_h2d,
                    "d2h_bytes_per_step": L.NSUMS * 8, "ms_per_step": obs_ms,
                    "h2d_gb_per_s_all_ranks": obs_h2d * world / (obs_ms * 1e-3) / 1e9,
                    "host_format": "observed entries of x only (fp32 stream + row offsets) + bit-packed mask; "
                                   "entries under mask == 0 never reach the training loss"}
    d2h = L.NSUMS * 8
    del dvals, dbits, doff, hvals, hbits, hoff, hx, dx, dm
    pnp = None
    if world == 1 and args.mnar_steps > 0:
        pnp = gpu_pnp_train(dev, ffma_tflops, B, table, mtable)
    del table, mtable, perm
    torch.cuda.empty_cache()

    # =========================== reward (cfg5) ===========================
    sec = None
    if args.reward_rows > 0:
        Dr, M = D_REWARD, args.reward_samples
        Nloc = args.reward_rows // world
        theta_r, _ = fresh_theta(L.FAMILY_MLP, Dr, 20, dev, seed=3)
        theta_r = theta_r * 2.0                                              # spread-out posteriors: rewards away from zero
        er = KR.Engine(L.FAMILY_MLP, Dr, 0, dev)
        gx = torch.Generator(device=dev).manual_seed(5 + rank)
        xr = torch.rand(Nloc, Dr, device=dev, generator=gx)
        mr = torch.zeros(Nloc, Dr, device=dev)                              # step t=0: nothing selected
        im = torch.rand(M, Nloc, Dr, device=dev, generator=gx)
        ws = None
        R, ws = er.reward(theta_r, xr, mr, im, ws)
        barrier(world)
        w0 = time.time()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.reward_steps):
            R, ws = er.reward(theta_r, xr, mr, im, ws)
        t1.record()
        barrier(world)
        clocks.window(w0, time.time())
        r_ms = max_over_ranks(t0.elapsed_time(t1), world) / args.reward_steps
        triples = Nloc * world * (Dr - 1) * M
        # e2e: x / mask / im in pinned host memory -> R in pinned host memory; rows go in blocks whose H2D copies run under
        # the reward kernel of the previous block (KR.Engine.reward_streamed)
        hxr, hmr, him = xr.cpu().pin_memory(), mr.cpu().pin_memory(), im.cpu().pin_memory()
        hR = torch.empty(Nloc, Dr - 1).pin_memory()
        R_dev = R.clone()
        del xr, mr, im, ws
        torch.cuda.empty_cache()
        er.reward_streamed(theta_r, hxr, hmr, him, hR, chunks=args.reward_chunks, copy_stream=copy_stream)   # warm-up
        barrier(world)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        h2d_r, d2h_r = er.reward_streamed(theta_r, hxr, hmr, him, hR, chunks=args.reward_chunks, copy_stream=copy_stream)
        t1.record()
        barrier(world)
        re2e_ms = max_over_ranks(t0.elapsed_time(t1), world)
        assert torch.equal(hR, R_dev.cpu()), "streamed reward differs from the resident one"
        r_tflops = triples / world * FLOP_REWARD_TRIPLE / (r_ms * 1e-3) / 1e12
        sec = {
            "metric": "active-selection rewards/s", "value": triples / (r_ms * 1e-3), "unit": "triples/s",
            "ms_per_step": r_ms, "steps": args.reward_steps, "scaling": "strong",
            "config": {"workload": f"cfg5: Reg_VAE reward sweep {Nloc * world} rows x {Dr - 1} candidates x {M} samples, "
                                   f"rows sharded over {world} GPU(s), no collective"},
            "e2e": {"value": triples / (re2e_ms * 1e-3), "unit": "triples/s", "ms": re2e_ms,
                    "h2d_bytes_per_step": int(h2d_r), "d2h_bytes_per_step": int(d2h_r),
                    "note": f"rows streamed in {args.reward_chunks} blocks, copies overlapped with the kernel of the previous block; "
                            "R bit-identical to the resident call"},
            "roofline": {"bound": "tensor",
                         "kernel": "k_reward_main_ws (warp-specialised: constructor / epilogue+KL / issuer warpgroups pipelined "
                                   "through mbarriers; tcgen05.mma kind::tf32, fp32-accurate 3xTF32 split, both dense layers "
                                   "of the tail MLP with the A operand in tensor memory) + k_reward_prep: whole "
                                   "pcvae_reward_chain call",
                         "achieved": r_tflops, "peak": tensor_peak, "unit": "TFLOP/s", "frac": r_tflops / tensor_peak,
                         "peak_source": tensor_src, "traffic": REWARD_DRAM_BYTES,
                         "vs_fp32_ffma": {"peak": ffma_tflops, "frac": r_tflops / ffma_tflops,
                                          "note": "the same algorithmic fp32 FLOP against the CUDA-core peak (can exceed 1: the "
                                                  "products run on the tensor pipe)"}},
            "gpu_launches": 4 * args.reward_steps,
        }
        del hxr, hmr, him

    al = gpu_al_loop(args, dev, world, rank) if args.al_rows > 0 else None

    clk = clocks.stop()
    if rank == 0:
        # dominant kernel of the step = the longest of the six tensor-core kernels, timed with the library's own events
        if kernel_us:
            dom = max(kernel_us, key=kernel_us.get)
            dom_flops = KERNEL_FLOP_BRANCH_ROW[dom] * 2 * B
            achieved = dom_flops / (kernel_us[dom] * 1e-6) / 1e12
        else:
            dom, achieved = "unavailable (per-kernel events are taken on one GPU)", 0.0
        step_tflops = FLOP_TRAIN_ROW * B / (train_ms * 1e-3) / 1e12
        cfg = config_dict(args)
        line = {
            "metric": "train rows/s (fwd+bwd+Adam), consistency-regularised partial VAE",
            "value": rows_s, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": train_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "arithmetic": "fp32 storage and accumulation; dense products as three kind::tf32 MMAs (hi*hi + hi*lo + lo*hi, "
                          "~2^-21 relative per product); losses, KL terms, Adam in fp32 / fp64 sums",
            "run": {"global_batch": B * world, "final_loss": loss, "graph": bool(graphed),
                    "tail": ("reduce + NVLink peer exchange + Adam in one launch" if (world > 1 and getattr(tr, "xch", None) is not None)
                             else ("NCCL grad all-reduce" if world > 1 else "reduce + Adam in one launch")),
                    "input": "device-side gather by the sampler permutation + Philox sub-mask / noise"},
            "e2e": {"value": e2e_rows_s, "unit": "rows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "h2d_gb_per_s_all_ranks": h2d * world / (e2e_ms * 1e-3) / 1e9,
                    "host_format": "x dense fp32 (every entry) + bit-packed mask; pcvae_prep_packed unpacks the mask and "
                                   "draws sub-mask / noise, then the same 7 launches as `value`"},
            "e2e_observed_only": e2e_observed,
            "gpu_launches": train_launches,
            "clocks": clk,
            "roofline": {"bound": "tensor", "kernel": dom + " (tcgen05.mma kind::tf32, fp32-accurate 3xTF32 split; longest of the "
                                                           "six kernels of the step)",
                         "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                         "traffic": KERNEL_DRAM_BYTES.get(dom) if B == 65536 else None,
                         "peak_source": tensor_src,
                         "kernel_us": kernel_us,
                         "step_achieved": step_tflops, "step_frac": step_tflops / tensor_peak,
                         # the same algorithmic fp32 FLOP against the FP32 FFMA peak measured on this GPU (the roofline
                         # of an implementation that keeps the products on the CUDA cores, as rounds before did)
                         "vs_fp32_ffma": {"peak": ffma_tflops, "kernel_frac": achieved / ffma_tflops,
                                          "step_frac": step_tflops / ffma_tflops,
                                          "peak_source": "pcvae_ffma_probe on this GPU"},
                         "hbm": {"achieved": sum(KERNEL_DRAM_BYTES.values()) / (train_ms * 1e-3) / 1e9 if B == 65536 else None,
                                 "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src,
                                 "note": "ncu DRAM bytes of the six kernels / step time"}},
            "secondary": sec,
        }
        if long_run is not None:
            line["long_run"] = long_run
        if strong is not None:
            line["strong"] = strong
        if al is not None:
            line["al_loop"] = al
        if world == 1 and args.mnar_steps > 0:
            line["mnar"] = gpu_mnar(args, dev, ffma_tflops, not args.no_cpu)
            line["pnp"] = pnp
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_train_baseline(B, 3, 1)
            if sec is not None:
                sec["cpu_baseline"] = cpu_reward_baseline(args.ref_reward_rows, 2, args.reward_samples)
        print(json.dumps(line))
    if getattr(tr, "xch", None) is not None:
        tr.xch.close()
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--table-rows", type=int, default=1_000_000)
    ap.add_argument("--long-steps", type=int, default=200, help="extra timed window of this many steps (1 GPU)")
    ap.add_argument("--reward-rows", type=int, default=100_000)
    ap.add_argument("--reward-samples", type=int, default=50)
    ap.add_argument("--reward-steps", type=int, default=3)
    ap.add_argument("--reward-chunks", type=int, default=8)
    ap.add_argument("--ref-reward-rows", type=int, default=100_000, help="rows of the CPU reward sample (2 candidates)")
    ap.add_argument("--al-rows", type=int, default=2000)
    ap.add_argument("--mnar-steps", type=int, default=100)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-port", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
