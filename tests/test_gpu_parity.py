"""CUDA path (through the C ABI) vs the oracle and the golden fixtures.  Needs a B200:
run with `pytest -m gpu` under gpurun.  Tolerances: north_star's 1e-4 relative on fp32
losses / ELBO / RMSE; masks bit-exact (the kernels read the mask bytes as given)."""
import math
import os
import sys

import ctypes as C

import pytest
import torch

from oracle import pcvae_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _mods():
    from vae_posterior_consistency_b200 import kernels as KR, lib as L
    return KR, L


def fam_of(p, L):
    if "type_pars1" in p:
        return L.FAMILY_PNP
    w1 = p["seq_encoder.0.weight"]                       # mask-augmented family: fan-in 2D (reference VAE.py:526)
    return L.FAMILY_MLP_MASK if w1.shape[1] == 2 * p["seq_decoder.4.bias"].numel() else L.FAMILY_MLP


def dims(p):
    D = p["seq_decoder.4.bias"].numel()
    K = p["type_pars1"].shape[1] if "type_pars1" in p else 0
    return D, K


def close(a, b, rtol=RTOL, atol=1e-6, msg=""):
    a_ = a.detach().cpu().to(b.dtype)
    if b.numel():
        from conftest import record_error
        rel = ((a_ - b).abs() / (b.abs() + atol / rtol)).max()
        record_error(f"losses / statistics, relative error (limit {rtol:g})", rel)
    torch.testing.assert_close(a_, b, rtol=rtol, atol=atol, msg=lambda m: f"{msg}: {m}")


def grad_close(g, ref, name):
    g = g.detach().cpu()
    atol = 2e-5 * float(ref.abs().max() + 1e-3)
    if ref.numel():
        from conftest import record_error
        # error of a gradient tensor relative to its largest entry (the unit the absolute floor is stated in)
        record_error("gradient entry, |err| / max|ref| (limit 2e-5 + 2e-3 rel)", (g - ref).abs().max() / (ref.abs().max() + 1e-3),
                     ref.abs().max())
    torch.testing.assert_close(g, ref, rtol=2e-3, atol=atol, msg=lambda m: f"grad {name}: {m}")


def engine_for(p):
    KR, L = _mods()
    D, K = dims(p)
    fam = fam_of(p, L)
    eng = KR.Engine(fam, D, K, "cuda")
    theta = KR.flatten_params(p, fam, "cuda")
    return eng, theta, fam


@pytest.fixture(params=[0, 1], ids=["ffma", "tcgen05"])
def train_tc(request):
    """Run the fused-step tests with both sets of training kernels: FP32 FFMA and tcgen05 3xTF32 (decoder: obs_dim % 4 == 0
    and obs_dim <= 104; encoder: MLP family with obs_dim % 4 == 0 and obs_dim <= 100, PNP family with emb_dim % 4 == 0 --
    its pooled embedding stays on the CUDA cores, the MLP tail runs on the tensor cores; other shapes take the FFMA
    kernels under either setting)."""
    KR, L = _mods()
    prev = L.load().pcvae_set_train_tensor_cores(request.param)
    yield request.param
    L.load().pcvae_set_train_tensor_cores(prev)


REG = ["reg_vae_b64_d13", "reg_vae_b37_d20_a05", "reg_eddi_b64_d13_k20", "reg_eddi_b33_d7_k10_a07",
       "reg_vae_mask_b64_d13", "reg_vae_mask_b37_d20_a05"]
VAN = ["vanilla_vae_b64_d13", "vanilla_eddi_b64_d13_k20", "vanilla_vae_mask_b64_d13"]


@pytest.mark.parametrize("name", REG)
def test_golden_reg_forward_and_fused_step(golden, name, train_tc):
    KR, L = _mods()
    g = golden(name)
    p = g["state_dict"]
    eng, theta, fam = engine_for(p)
    cu = lambda t: t.cuda()
    x, mask, mask_p = cu(g["x"]), cu(g["mask"]), cu(g["mask_p"])
    eq, ep = cu(g["eps_q"]), cu(g["eps_p"])
    mean, logvar, z, ws = eng.enc_fwd(theta, x, [mask, mask_p], [eq, ep], save=True)
    close(mean[0], g["mean_q"], msg="mean_q"); close(logvar[0], g["logvar_q"], msg="logvar_q")
    close(mean[1], g["mean_p"], msg="mean_p"); close(logvar[1], g["logvar_p"], msg="logvar_p")
    xh = eng.dec(L.DEC_FWD, theta, z)["xhat"]
    close(xh[0], g["xh_q"], msg="xh_q"); close(xh[1], g["xh_p"], msg="xh_p")
    # fused step: loss + all parameter gradients vs the reference's autograd
    B = x.shape[0]
    tr = KR.FusedTrainer(fam, eng.D, eng.K, theta.clone(), regularised=True, alpha=g["alpha"])
    sums = tr.forward_backward(x, mask, mask_p, eq, ep)
    loss = KR.loss_from_sums(sums, B, g["alpha"], 1.0, True)
    close(loss.float(), g["train_loss"], msg="train_loss")
    grads = KR.unflatten_params(tr.grad, fam, eng.D, eng.K)
    for k, ref in g["grads"].items():
        grad_close(grads[k], ref, k)
    # evaluate-stage sums
    eng.dec(L.DEC_EVAL, theta, [z[0]], x=x, masks=[mask], mean=[mean[0]], logvar=[logvar[0]])
    s = eng.reduce_sums(B).cpu()
    close(((s[L.S_RE_Q] + s[L.S_KL_Q]) / B).float(), g["eval_loss"], msg="eval_loss")
    close((s[L.S_RE_Q] / B).float(), g["negl"], msg="negl")
    close((s[L.S_RE_IMP] / B).float(), g["negl_imp"], msg="negl_imp")
    n_unobs = (~g["mask"]).sum()
    close(torch.sqrt(s[L.S_SSE_UNOBS] / n_unobs).float(), g["rmse"], msg="rmse")


@pytest.mark.parametrize("name", VAN)
def test_golden_vanilla_fused_step(golden, name):
    KR, L = _mods()
    g = golden(name)
    p = g["state_dict"]
    eng, theta, fam = engine_for(p)
    x, eq = g["x"].cuda(), g["eps_q"].cuda()
    maskf = (g["mask"] * torch.ones(g["x"].shape)).cuda()          # float32 mask, train.py:58,97
    tr = KR.FusedTrainer(fam, eng.D, eng.K, theta.clone(), regularised=False)
    sums = tr.forward_backward(x, maskf, None, eq, None)
    loss = KR.loss_from_sums(sums, x.shape[0], 0.0, 1.0, False)
    close(loss.float(), g["train_loss"], msg="train_loss")
    grads = KR.unflatten_params(tr.grad, fam, eng.D, eng.K)
    for k, ref in g["grads"].items():
        grad_close(grads[k], ref, k)


@pytest.mark.parametrize("name", ["traj_reg_vae_b32_d13", "traj_reg_eddi_b32_d13_k10", "traj_reg_vae_mask_b32_d13"])
def test_golden_training_trajectory_with_adam(golden, name, train_tc):
    KR, L = _mods()
    g = golden(name)
    p = g["state_dict0"]
    eng, theta, fam = engine_for(p)
    tr = KR.FusedTrainer(fam, eng.D, eng.K, theta, regularised=True, alpha=1.0)
    for s in range(g["x"].shape[0]):
        loss = tr.step(g["x"][s].cuda(), g["mask"][s].cuda(), g["mask_p"][s].cuda(), g["eps_q"][s].cuda(),
                       g["eps_p"][s].cuda())
        close(loss.float(), g["losses"][s], msg=f"loss step {s}")
    end = KR.unflatten_params(tr.theta, fam, eng.D, eng.K)
    for k, v in end.items():
        torch.testing.assert_close(v.cpu(), g["state_dict_end"][k], rtol=1e-3, atol=1e-5, msg=lambda m: f"{k}: {m}")


def rand_case(family, B, D, K, seed, mask_float=False):
    p = O.init_params(family, D, K, seed=seed)
    g = torch.Generator().manual_seed(seed + 100)
    x = torch.rand(B, D, generator=g)
    mask = torch.rand(B, D, generator=g) < 0.7
    mask_p = mask & (torch.rand(B, D, generator=g) < 0.7)
    eq, ep = torch.randn(B, 10, generator=g), torch.randn(B, 10, generator=g)
    if mask_float:
        mask, mask_p = mask.float(), mask_p.float()
    return p, x, mask, mask_p, eq, ep


SHAPES = [("mlp", 1, 4, 0), ("mlp", 129, 8, 0), ("mlp", 300, 100, 0), ("mlp", 257, 104, 0), ("pnp", 131, 12, 10),
          ("mlp", 1, 2, 0), ("mlp", 63, 13, 0), ("mlp", 65, 50, 0), ("mlp", 200, 100, 0), ("mlp", 130, 101, 0),
          ("mlp", 70, 128, 0), ("pnp", 1, 2, 10), ("pnp", 65, 13, 20), ("pnp", 200, 100, 20), ("pnp", 97, 50, 7),
          ("pnp", 64, 128, 20), ("pnp", 300, 40, 4), ("pnp", 129, 60, 32), ("pnp", 1000, 100, 8),
          ("mlp_mask", 1, 2, 0), ("mlp_mask", 65, 13, 0), ("mlp_mask", 130, 50, 0), ("mlp_mask", 200, 100, 0),
          ("mlp_mask", 97, 21, 0)]


@pytest.mark.parametrize("family,B,D,K", SHAPES)
@pytest.mark.parametrize("alpha", [1.0, 0.3])
def test_fused_step_vs_oracle_random_shapes(family, B, D, K, alpha, train_tc):
    KR, L = _mods()
    p, x, mask, mask_p, eq, ep = rand_case(family, B, D, K, seed=B + D)
    eng, theta, fam = engine_for(p)
    tr = KR.FusedTrainer(fam, D, K, theta, regularised=True, alpha=alpha, beta_w=0.7)
    sums = tr.forward_backward(x.cuda(), mask.cuda(), mask_p.cuda(), eq.cuda(), ep.cuda())
    loss = KR.loss_from_sums(sums, B, alpha, 0.7, True)
    ref_loss, ref_grads, aux = O.train_step(p, x, mask, mask_p, eq, ep, alpha=alpha, beta=0.7)
    close(loss.float(), ref_loss, msg="loss")
    grads = KR.unflatten_params(tr.grad, fam, D, K)
    for k in O.trainable_names(p):
        grad_close(grads[k], ref_grads[k], k)


@pytest.mark.parametrize("family,B,D,K", [("mlp", 100, 20, 0), ("pnp", 100, 20, 10), ("mlp_mask", 100, 20, 0)])
def test_modular_ops_vs_oracle(family, B, D, K):
    """decoder backward from an arbitrary d_xhat, encoder backward from arbitrary d_mean/d_logvar,
    and the stand-alone loss kernel (what the nn.Module API composes)."""
    KR, L = _mods()
    p, x, mask, mask_p, eq, ep = rand_case(family, B, D, K, seed=5, mask_float=(family == "pnp"))
    eng, theta, fam = engine_for(p)
    g = torch.Generator().manual_seed(9)
    names = O.trainable_names(p)
    q = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in p.items()}
    mu, lv = O.encoder_stats(q, x, mask, collapsed=True)
    z = O.reparam(mu, lv, eq)
    xh = O.decoder(q, z)
    w_xh, w_mu, w_lv = torch.randn(B, D, generator=g), torch.randn(B, 10, generator=g), torch.randn(B, 10, generator=g)
    obj = (xh * w_xh).sum() + (mu * w_mu).sum() + (lv * w_lv).sum()
    ref = dict(zip(names, torch.autograd.grad(obj, [q[k] for k in names], allow_unused=True)))
    # CUDA: enc_fwd -> dec BWD with d_xhat = w_xh -> d_z; chain to d_mean/d_logvar; enc_bwd
    xc, mc = x.cuda(), mask.cuda()
    mean, logvar, zc, ws = eng.enc_fwd(theta, xc, [mc], [eq.cuda()], save=True)
    close(zc[0], z.detach(), msg="z")
    out = eng.dec(L.DEC_BWD, theta, zc, d_xhat=[w_xh.cuda()])
    dz = out["d_z"][0]
    d_mean = dz + w_mu.cuda()
    d_logvar = dz * 0.5 * torch.exp(0.5 * logvar[0]) * eq.cuda() + w_lv.cuda()
    eng.enc_bwd(theta, xc, [mc], ws, [d_mean], [d_logvar])
    grad = torch.zeros_like(theta)
    eng.reduce_grads(grad)
    grads = KR.unflatten_params(grad, fam, D, K)
    for k in names:
        grad_close(grads[k], ref[k] if ref[k] is not None else torch.zeros_like(p[k]), k)
    # stand-alone loss terms + gradients
    mu_p, lv_p = O.encoder_stats(p, x, mask_p, collapsed=True)
    xh_p = O.decoder(p, O.reparam(mu_p, lv_p, ep))
    leaves = [t.detach().clone().requires_grad_(True) for t in (xh_p, mu_p, lv_p, xh.detach(), mu.detach(), lv.detach())]
    loss, _, _ = O.reg_loss(x, leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], leaves[5], mask, mask_p,
                            beta=0.9, alpha=0.6)
    rg = torch.autograd.grad(loss, leaves)
    sums, d_xhat, d_mean2, d_logvar2 = eng.loss_terms(
        xc, [mc, mask_p.cuda()], [leaves[3].detach().cuda(), leaves[0].detach().cuda()],
        [leaves[4].detach().cuda(), leaves[1].detach().cuda()], [leaves[5].detach().cuda(), leaves[2].detach().cuda()],
        alpha=0.6, beta_w=0.9, loss_scale=1.0 / B, want_grads=True)
    close(KR.loss_from_sums(sums, B, 0.6, 0.9, True).float(), loss.detach(), msg="loss_terms loss")
    for got, want, nm in [(d_xhat[1], rg[0], "d_xh_p"), (d_mean2[1], rg[1], "d_mu_p"), (d_logvar2[1], rg[2], "d_lv_p"),
                          (d_xhat[0], rg[3], "d_xh_q"), (d_mean2[0], rg[4], "d_mu_q"), (d_logvar2[0], rg[5], "d_lv_q")]:
        torch.testing.assert_close(got.cpu(), want, rtol=1e-4, atol=1e-7, msg=lambda m: f"{nm}: {m}")


@pytest.mark.parametrize("D", [20, 13])
def test_enc_bwd_folds_reparameterisation(D, train_tc):
    """pcvae_enc_bwd with d_z / eps / logvar given == the caller chaining z = mean + eps * exp(logvar / 2) by hand
    (VAE.py:390-392); D = 20 takes the tcgen05 encoder when enabled, D = 13 the FFMA one."""
    KR, L = _mods()
    B = 150
    p, x, mask, mask_p, eq, ep = rand_case("mlp", B, D, 0, seed=77)
    eng, theta, fam = engine_for(p)
    g = torch.Generator().manual_seed(4)
    d_z, w_mu, w_lv = (torch.randn(B, 10, generator=g).cuda() for _ in range(3))
    xc, mc, eqc = x.cuda(), mask.cuda(), eq.cuda()
    grads = []
    for folded in (True, False):
        mean, logvar, z, ws = eng.enc_fwd(theta, xc, [mc], [eqc], save=True)
        if folded:
            eng.enc_bwd(theta, xc, [mc], ws, [w_mu], [w_lv], d_z=[d_z], eps=[eqc], logvar=logvar)
        else:
            eng.enc_bwd(theta, xc, [mc], ws, [w_mu + d_z], [w_lv + d_z * 0.5 * torch.exp(0.5 * logvar[0]) * eqc])
        grad = torch.zeros_like(theta)
        eng.reduce_grads(grad, 0, eng.dec_off)
        grads.append(grad.cpu())
    torch.testing.assert_close(grads[0], grads[1], rtol=1e-5, atol=1e-6 * float(grads[1].abs().max()))
    assert float(grads[1].abs().max()) > 0


@pytest.mark.parametrize("B,D", [(1, 4), (127, 12), (129, 20), (300, 8), (65, 96), (1000, 100), (513, 104), (4100, 100)])
@pytest.mark.parametrize("regularised,mask_float", [(True, False), (False, True), (True, True)])
def test_tcgen05_decoder_matches_ffma_decoder(B, D, regularised, mask_float):
    """Same fused step through the FFMA kernels and through k_enc_*_tc + k_dec_*_tc + k_wgrad_tc (3xTF32): loss sums and
    every gradient must agree to fp32 rounding; covers one-branch (vanilla) training, float masks, ragged tiles and
    more rows than one wave of CTAs."""
    KR, L = _mods()
    p, x, mask, mask_p, eq, ep = rand_case("mlp", B, D, 0, seed=3 * B + D, mask_float=mask_float)
    lib = L.load()
    out = []
    for tc in (0, 1):
        prev = lib.pcvae_set_train_tensor_cores(tc)
        try:
            eng, theta, fam = engine_for(p)
            tr = KR.FusedTrainer(fam, D, 0, theta, regularised=regularised, alpha=0.6, beta_w=0.9)
            sums = tr.forward_backward(x.cuda(), mask.cuda(), mask_p.cuda() if regularised else None, eq.cuda(),
                                       ep.cuda() if regularised else None).cpu()
            out.append((sums, tr.grad.cpu().clone()))
        finally:
            lib.pcvae_set_train_tensor_cores(prev)
    (s0, g0), (s1, g1) = out
    torch.testing.assert_close(s1, s0, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(g1, g0, rtol=2e-4, atol=2e-6 * float(g0.abs().max()))


def test_throughput_prep_then_step_tcgen05_matches_ffma():
    """Throughput-mode sequence at the bench shape: pcvae_prep_batch (gather + Philox sub-mask + noise in one launch)
    followed by the fused step, through the tcgen05 kernels and through the FFMA kernels on the same prepared batch.
    (Running other kernels right before the tensor-core ones also guards against state they might depend on.)"""
    KR, L = _mods()
    lib = L.load()
    B, D, T = 65536, 100, 200_000
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(11)
    table = torch.rand(T, D, device=dev, generator=g)
    mtable = torch.rand(T, D, device=dev, generator=g) < 0.7
    idx = torch.randint(0, T, (B,), device=dev, generator=g)
    p = O.init_params("mlp", D, 0, seed=5)
    out = []
    for tc in (0, 1):
        prev = lib.pcvae_set_train_tensor_cores(tc)
        try:
            x = torch.empty(B, D, device=dev); mask = torch.empty(B, D, device=dev, dtype=torch.bool)
            mask_p = torch.empty(B, D, device=dev, dtype=torch.bool); eps = torch.empty(2, B, 10, device=dev)
            L.check(lib.pcvae_prep_batch(table.data_ptr(), mtable.data_ptr(), idx.data_ptr(), x.data_ptr(), mask.data_ptr(),
                                         mask_p.data_ptr(), eps.data_ptr(), B, D, 2, 0.7, 99, 8,
                                         torch.cuda.current_stream().cuda_stream), "pcvae_prep_batch")
            theta = KR.flatten_params(p, L.FAMILY_MLP, "cuda")
            tr = KR.FusedTrainer(L.FAMILY_MLP, D, 0, theta, regularised=True, alpha=1.0)
            loss = tr.step(x, mask, mask_p, eps[0], eps[1])
            torch.cuda.synchronize()
            out.append((float(loss), tr.grad.cpu().clone(), tr.theta.cpu().clone(), x.cpu(), mask.cpu(), mask_p.cpu(), eps.cpu()))
        finally:
            lib.pcvae_set_train_tensor_cores(prev)
    (l0, g0, t0, x0, m0, mp0, e0), (l1, g1, t1, x1, m1, mp1, e1) = out
    assert torch.equal(x0, x1) and torch.equal(m0, m1) and torch.equal(mp0, mp1) and torch.equal(e0, e1)   # Philox: same counters
    assert torch.equal(x0, table[idx].cpu()) and torch.equal(m0, mtable[idx].cpu())                          # the gather itself
    assert bool((mp0 <= m0).all()) and 0.66 < float(mp0.float().sum() / m0.float().sum()) < 0.74
    assert abs(float(e0.mean())) < 5e-3 and abs(float(e0.std()) - 1.0) < 5e-3
    assert abs(l1 - l0) <= 1e-5 * abs(l0)
    torch.testing.assert_close(g1, g0, rtol=2e-3, atol=2e-5 * float(g0.abs().max()))
    torch.testing.assert_close(t1, t0, rtol=0, atol=2.1e-3)      # one Adam step moves every weight by at most lr = 1e-3


@pytest.mark.parametrize("B,D", [(1, 4), (777, 100), (4096, 128), (300, 36)])
def test_prep_packed_equals_prep_batch_on_the_same_rows(B, D):
    """Host-streamed batch format (bit-packed mask + observed-entry stream of x): pcvae_prep_packed must rebuild
    the dense mask bit for bit, x * mask exactly, and draw the SAME sub-mask and noise as pcvae_prep_batch does
    for those rows (same Philox counters); a fused step on either batch then gives identical results."""
    KR, L = _mods()
    lib = L.load()
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(B + D)
    xh = torch.rand(B, D, generator=g)
    mh = torch.rand(B, D, generator=g) < 0.7
    if B > 2:
        mh[1] = False                                     # a row with nothing observed
        mh[2] = True                                      # and a fully observed one
    vals, row_off, bits = KR.compact_rows(xh, mh)
    assert vals.numel() == int(mh.sum()) and bits.shape == (B, (D + 31) // 32)
    eng = KR.Engine(L.FAMILY_MLP, D, 0, dev)
    st = torch.cuda.current_stream().cuda_stream
    # reference: the resident-table path with the identity permutation
    table, mtable = xh.to(dev), mh.to(dev)
    idx = torch.arange(B, device=dev)
    x0 = torch.empty(B, D, device=dev); m0 = torch.empty(B, D, device=dev, dtype=torch.bool)
    mp0 = torch.empty_like(m0); e0 = torch.empty(2, B, 10, device=dev)
    L.check(lib.pcvae_prep_batch(table.data_ptr(), mtable.data_ptr(), idx.data_ptr(), x0.data_ptr(), m0.data_ptr(),
                                 mp0.data_ptr(), e0.data_ptr(), B, D, 2, 0.7, 99, 24, st), "pcvae_prep_batch")
    # compact form: observed entries only
    x1 = torch.full((B, D), float("nan"), device=dev); m1 = torch.empty(B, D, device=dev, dtype=torch.bool)
    mp1 = torch.empty_like(m1); e1 = torch.empty(2, B, 10, device=dev)
    eng.prep_packed(bits.to(dev), m1, mp1, e1, vals=vals.to(dev) if vals.numel() else torch.zeros(1, device=dev),
                    row_off=row_off.to(dev), x=x1, keep=0.7, seed=99, offset=24)
    assert torch.equal(m1, m0) and torch.equal(mp1, mp0) and torch.equal(e1, e0)
    assert torch.equal(x1, x0 * m0)                       # unobserved entries are zero, observed ones exact
    # dense x copied separately, mask bit-packed only
    m2 = torch.empty(B, D, device=dev, dtype=torch.bool); mp2 = torch.empty_like(m2); e2 = torch.empty(2, B, 10, device=dev)
    eng.prep_packed(bits.to(dev), m2, mp2, e2, keep=0.7, seed=99, offset=24)
    assert torch.equal(m2, m0) and torch.equal(mp2, mp0) and torch.equal(e2, e0)
    if D % 4 == 0 and D <= 100:
        p = O.init_params("mlp", D, 0, seed=3)
        res = []
        for xs, ms, mps, es in ((x0, m0, mp0, e0), (x1, m1, mp1, e1)):
            tr = KR.FusedTrainer(L.FAMILY_MLP, D, 0, KR.flatten_params(p, L.FAMILY_MLP, "cuda"), regularised=True)
            res.append((tr.forward_backward(xs, ms, mps, es[0], es[1]).clone(), tr.grad.clone()))
        assert torch.equal(res[0][0][:6], res[1][0][:6])  # every training sum (the imputed / SSE slots read unobserved x)
        assert torch.equal(res[0][1], res[1][1])          # and every gradient, bit for bit


@pytest.mark.parametrize("B,D,T", [(64, 20, 455), (4096, 100, 20000)])
def test_graph_replay_of_the_fused_step_is_bit_identical_to_eager_launches(B, D, T, monkeypatch):
    """GraphedFusedTrainer: the whole step (prep_batch_dev + six kernels + reduce_adam_dev, per-step scalars from a device
    counter) replayed from a CUDA graph == the same launches issued eagerly, bit for bit; and == the host-counted
    sequence pcvae_prep_batch(offset = 8 step) -> FusedTrainer.step up to the rounding of Adam's bias correction."""
    KR, L = _mods()
    lib = L.load()
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(3)
    table = torch.rand(T, D, device=dev, generator=g)
    mtable = torch.rand(T, D, device=dev, generator=g) < 0.7
    nb, steps = T // B, 9
    idx = torch.stack([torch.randperm(T, device=dev, generator=g)[:B] for _ in range(nb)])
    p = O.init_params("mlp", D, 0, seed=7)
    mk = lambda: KR.GraphedFusedTrainer(L.FAMILY_MLP, D, 0, KR.flatten_params(p, L.FAMILY_MLP, dev), table, mtable, B, nb,
                                        keep=0.7, seed=99)
    a, b = mk(), mk()
    a.set_batches(idx); b.set_batches(idx)
    a.capture()                                            # three eager warm-up steps, then the capture
    for _ in range(steps):
        a.step_graph()
    for _ in range(3 + steps):
        b.step_eager_dev()
    torch.cuda.synchronize()
    assert a.step_count == b.step_count == 3 + steps and int(a.state[0]) == int(b.state[0]) == 3 + steps
    assert torch.equal(a.theta, b.theta) and torch.equal(a.exp_avg_sq, b.exp_avg_sq) and torch.equal(a.total, b.total)
    # a and b prepare the batch of step n + 1 on a forked stream beside the last launch of step n (the default); c runs the
    # plain prep -> step order (PCVAE_PREP_AHEAD=0); new index lists in the middle of the run
    assert a.ahead and b.ahead
    monkeypatch.setenv("PCVAE_PREP_AHEAD", "0")
    c = mk()
    monkeypatch.delenv("PCVAE_PREP_AHEAD")
    assert not c.ahead
    c.set_batches(idx)
    for _ in range(3 + steps):
        c.step_eager_dev()
    assert torch.equal(a.theta, c.theta) and torch.equal(a.total, c.total)
    idx2 = torch.stack([torch.randperm(T, device=dev, generator=g)[:B] for _ in range(nb)])
    a.set_batches(idx2); c.set_batches(idx2)
    for _ in range(nb + 2):
        a.step_graph(); c.step_eager_dev()
    torch.cuda.synchronize()
    assert torch.equal(a.theta, c.theta) and torch.equal(a.total, c.total)
    # host-counted reference sequence (against b, which stopped after 3 + steps steps and equalled a there bit for bit)
    tr = KR.FusedTrainer(L.FAMILY_MLP, D, 0, KR.flatten_params(p, L.FAMILY_MLP, dev), regularised=True)
    x = torch.empty(B, D, device=dev); m = torch.empty(B, D, device=dev, dtype=torch.bool); mp = torch.empty_like(m)
    eps = torch.empty(2, B, 10, device=dev)
    total = 0.0
    for s_ in range(3 + steps):
        L.check(lib.pcvae_prep_batch(table.data_ptr(), mtable.data_ptr(), idx[s_ % nb].data_ptr(), x.data_ptr(), m.data_ptr(),
                                     mp.data_ptr(), eps.data_ptr(), B, D, 2, 0.7, 99, 8 * s_,
                                     torch.cuda.current_stream().cuda_stream), "pcvae_prep_batch")
        total += float(tr.step(x, m, mp, eps[0], eps[1]))
    torch.testing.assert_close(b.theta, tr.theta, rtol=0, atol=1e-6)
    assert abs(float(b.total) - total) <= 1e-6 * abs(total)


def test_empty_batch_is_a_no_op():
    KR, L = _mods()
    p = O.init_params("mlp", 13)
    eng, theta, fam = engine_for(p)
    x = torch.zeros(0, 13, device="cuda")
    mean, logvar, z, _ = eng.enc_fwd(theta, x, [torch.zeros(0, 13, dtype=torch.bool, device="cuda")])
    assert mean[0].shape == (0, 10) and z[0].shape == (0, 10)        # VAE.py:723-724 empty guard


@pytest.fixture(params=[0, 1, 2], ids=["ffma", "tcgen05", "tcgen05-lockstep"])
def reward_tc(request):
    """Run the reward tests with all three main kernels: FP32 FFMA, tcgen05 3xTF32 warp-specialised (the default) and
    tcgen05 3xTF32 in lock-step phases (MLP family)."""
    KR, L = _mods()
    prev = L.load().pcvae_set_reward_tensor_cores(request.param)
    yield request.param
    L.load().pcvae_set_reward_tensor_cores(prev)


@pytest.mark.parametrize("name", ["reward_reg_vae_n24_d8_m5", "reward_reg_eddi_n24_d8_k10_m5"])
def test_golden_reward(golden, name, reward_tc):
    KR, L = _mods()
    g = golden(name)
    eng, theta, fam = engine_for(g["state_dict"])
    R, _ = eng.reward(theta, g["x"].cuda(), g["mask"].cuda(), g["im"].cuda())
    R = R.cpu()
    ref = g["R"]
    sel = g["mask"][:, :-1] != 0
    assert torch.all(R[sel] == -1e4)
    from conftest import record_error
    record_error("reward vs reference, |err| (limit 1e-4 |R| + 2e-6)", (R[~sel] - ref[~sel]).abs().max(), ref[~sel].abs().max())
    torch.testing.assert_close(R[~sel], ref[~sel], rtol=1e-4, atol=2e-6)
    # selection order equal except where the reward gap is below tolerance
    top2 = ref.topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 1e-5
    assert torch.equal(R.argmax(1)[decided], ref.argmax(1)[decided])


@pytest.mark.parametrize("family,N,D,K,M", [("mlp", 150, 20, 0, 7), ("pnp", 150, 20, 10, 7), ("mlp", 70, 101, 0, 3),
                                            ("pnp", 40, 101, 20, 2), ("mlp", 5, 2, 0, 1)])
def test_reward_vs_oracle_random(family, N, D, K, M, reward_tc):
    KR, L = _mods()
    p = O.init_params(family, D, K, seed=N + D)
    p = {k: (v * 2.0 if v.dtype == torch.float32 and not k.startswith("prior") else v) for k, v in p.items()}
    g = torch.Generator().manual_seed(N)
    x = torch.rand(N, D, generator=g)
    mask = (torch.rand(N, D, generator=g) < 0.4).float()
    mask[:, -1] = 0
    im = torch.rand(M, N, D, generator=g)
    eng, theta, fam = engine_for(p)
    R, _ = eng.reward(theta, x.cuda(), mask.cuda(), im.cuda())
    ref = O.reward_all(p, x, mask, im, incremental=True)
    ref64 = O.reward_all({k: v.double() for k, v in p.items()}, x.double(), mask.double(), im.double())
    sel = mask[:, :-1] != 0
    assert torch.all(R.cpu()[sel] == -1e4)
    # both fp32 evaluations must sit within the fp32 noise floor of the fp64 value
    err_cuda = (R.cpu()[~sel].double() - ref64[~sel]).abs().max()
    err_ref = (ref[~sel].double() - ref64[~sel]).abs().max()
    scale = ref64[~sel].abs().max()
    from conftest import record_error
    record_error("reward vs fp64, |err| / (fp32 CPU reference's own |err|)", err_cuda / max(float(err_ref), 1e-12), scale)
    assert err_cuda <= max(4 * err_ref, 1e-4 * scale + 2e-6), (err_cuda, err_ref, scale)


def test_reward_is_row_shardable_bit_exact(reward_tc):
    """Rows are independent: evaluating two row blocks separately must reproduce the single
    call bit for bit (what the multi-GPU sharding relies on)."""
    KR, L = _mods()
    p = O.init_params("mlp", 20, seed=3)
    g = torch.Generator().manual_seed(1)
    N, D, M = 333, 20, 6
    x = torch.rand(N, D, generator=g).cuda()
    mask = (torch.rand(N, D, generator=g) < 0.3).float()
    mask[:, -1] = 0
    mask = mask.cuda()
    im = torch.rand(M, N, D, generator=g).cuda()
    eng, theta, fam = engine_for(p)
    R, _ = eng.reward(theta, x, mask, im)
    cut = 140
    Ra, _ = eng.reward(theta, x[:cut], mask[:cut], im[:, :cut].contiguous())
    Rb, _ = eng.reward(theta, x[cut:], mask[cut:], im[:, cut:].contiguous())
    assert torch.equal(R, torch.cat([Ra, Rb]))


@pytest.mark.parametrize("N,D,M", [(333, 20, 6), (70, 101, 3), (5, 2, 1), (1500, 100, 50), (9, 8, 2)])
def test_warp_specialised_reward_kernel_against_lockstep_kernel(N, D, M):
    """pcvae_reward_ws.cu computes what pcvae_reward_tc.cu computes with the same operand split and scalar arithmetic;
    the scheduling differs (warpgroup roles pipelined through mbarriers, sample counter running across tiles) and layer 3
    keeps hi*lo in a second accumulator, so the two agree to rounding of the layer-3 outputs (observed difference
    printed), and the warp-specialised kernel reproduces itself bit for bit."""
    KR, L = _mods()
    p = O.init_params("mlp", D, seed=N)
    g = torch.Generator().manual_seed(N + M)
    x = torch.rand(N, D, generator=g).cuda()
    mask = (torch.rand(N, D, generator=g) < 0.3).float()
    mask[:, -1] = 0
    mask = mask.cuda()
    im = torch.rand(M, N, D, generator=g).cuda()
    eng, theta, fam = engine_for(p)
    lib = L.load()
    prev = lib.pcvae_set_reward_tensor_cores(2)
    try:
        R_lock, _ = eng.reward(theta, x, mask, im)
        R_lock = R_lock.clone()
        lib.pcvae_set_reward_tensor_cores(1)
        R_ws, _ = eng.reward(theta, x, mask, im)
        torch.cuda.synchronize()
    finally:
        lib.pcvae_set_reward_tensor_cores(prev)
    sel = mask[:, :-1] != 0
    assert torch.equal(R_ws[sel], R_lock[sel]) and torch.all(R_ws[sel] == -1e4)
    diff = float((R_ws - R_lock)[~sel].abs().max()) if (~sel).any() else 0.0
    scale = float(R_lock[~sel].abs().max()) if (~sel).any() else 0.0
    from conftest import record_error
    record_error("reward, warp-specialised vs lock-step kernel, |diff|", diff, scale)
    assert diff <= 2e-5 * scale + 1e-6
    # and a second call reproduces it (no state left in the barriers' phases, no race between the roles)
    R_again, _ = eng.reward(theta, x, mask, im)
    assert torch.equal(R_again, R_ws)


@pytest.mark.parametrize("family,B,D,K", [("mlp", 300, 100, 0), ("mlp", 64, 20, 0), ("pnp", 200, 100, 20), ("mlp_mask", 130, 52, 0)])
def test_prebuilt_weight_images_give_bit_identical_steps(family, B, D, K, monkeypatch):
    """pcvae_build_weight_images: the operand images of the weights built once per theta and fetched by the tensor-core
    kernels with bulk copies == the images every CTA builds for itself (the builder runs the kernels' own build code)."""
    KR, L = _mods()
    p, x, mask, mask_p, eq, ep = rand_case(family, B, D, K, seed=B + D)
    eng, theta, fam = engine_for(p)
    args = (x.cuda(), mask.cuda(), mask_p.cuda(), eq.cuda(), ep.cuda())
    tr = KR.FusedTrainer(fam, D, K, theta.clone(), regularised=True)
    assert (tr.wimg is not None) == (L.load().pcvae_weight_images_floats(C.byref(tr.eng.model)) > 0)
    monkeypatch.setenv("PCVAE_WEIGHT_IMAGES", "0")
    tr0 = KR.FusedTrainer(fam, D, K, theta.clone(), regularised=True)
    monkeypatch.delenv("PCVAE_WEIGHT_IMAGES")
    assert tr0.wimg is None
    for _ in range(3):                                    # theta changes every step: the images must follow it
        l1, l0 = tr.step(*args), tr0.step(*args)
        assert torch.equal(l1, l0)
    assert torch.equal(tr.theta, tr0.theta) and torch.equal(tr.grad, tr0.grad)


_PDL_SNIPPET = r"""
import hashlib, sys, torch
sys.path.insert(0, %r)
from vae_posterior_consistency_b200 import lib as L, VAE, kernels as KR
B, D = 16384, 100
torch.manual_seed(3)
model = VAE.Reg_VAE(D, 500, 0, 10, {"batch_size": 64, "patience": 100}, "pdl", "kl_reg")
tr = KR.FusedTrainer(L.FAMILY_MLP, D, 0, model.flat_theta().detach().clone().cuda(), regularised=True)
g = torch.Generator().manual_seed(5)
x = torch.rand(B, D, generator=g).cuda(); mask = (torch.rand(B, D, generator=g) < 0.7).cuda()
mask_p = mask & (torch.rand(B, D, generator=g) < 0.7).cuda()
eq, ep = torch.randn(B, 10, generator=g).cuda(), torch.randn(B, 10, generator=g).cuda()
for _ in range(4):
    loss = tr.step(x, mask, mask_p, eq, ep)
torch.cuda.synchronize()
print("HASH", hashlib.sha256(tr.theta.cpu().numpy().tobytes() + tr.exp_avg_sq.cpu().numpy().tobytes()).hexdigest(), float(loss))
"""


def test_dependent_launches_do_not_change_a_bit():
    """PCVAE_PDL (programmatic dependent launch along the training chain, read once per process): four eager steps at
    16 384 x 100 in a process with it and in one without must leave bit-identical theta and Adam state."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {}
    for v in ("0", "1"):
        env = dict(os.environ, PCVAE_PDL=v)
        r = subprocess.run([sys.executable, "-c", _PDL_SNIPPET % root], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        line = [l for l in r.stdout.splitlines() if l.startswith("HASH")]
        assert line, r.stdout[-500:]
        out[v] = line[0]
    assert out["0"] == out["1"], out


def test_large_batch_properties(train_tc):
    """BASELINE cfg4 shape (batch 65536 x 100): the fused step must be deterministic (two runs
    bit-identical) and linear in loss_scale; loss must match the oracle on a 4096-row slice."""
    KR, L = _mods()
    B, D = 65536, 100
    p, x, mask, mask_p, eq, ep = rand_case("mlp", B, D, 0, seed=11)
    eng, theta, fam = engine_for(p)
    xc, mc, mpc, eqc, epc = x.cuda(), mask.cuda(), mask_p.cuda(), eq.cuda(), ep.cuda()
    tr = KR.FusedTrainer(fam, D, 0, theta, regularised=True)
    s1 = tr.forward_backward(xc, mc, mpc, eqc, epc).clone()
    g1 = tr.grad.clone()
    s2 = tr.forward_backward(xc, mc, mpc, eqc, epc)
    assert torch.equal(s1, s2) and torch.equal(g1, tr.grad)
    n = 4096
    s3 = tr.forward_backward(xc[:n], mc[:n], mpc[:n], eqc[:n], epc[:n])
    ref_loss, ref_grads, _ = O.train_step(p, x[:n], mask[:n], mask_p[:n], eq[:n], ep[:n])
    close(KR.loss_from_sums(s3, n, 1.0, 1.0, True).float(), ref_loss, msg="loss 4096")
    grads = KR.unflatten_params(tr.grad, fam, D, 0)
    for k in O.trainable_names(p):
        grad_close(grads[k], ref_grads[k], k)
