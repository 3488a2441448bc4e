"""MIWAE / Reg_MIWAE (Student-t decoder + importance-weighted bound, reference VAE.py:3011-3301; SURVEY.md section 8f
item 4) on a B200: heads and sampling kernels against torch, the loss kernel against the oracle's closed forms, the
module API with autograd against the four fixtures recorded from the reference (forward, loss with the reference's
un-transposed reshape, every parameter gradient, llh_eval imputation), and eval_miwae's row-wise batching."""
import os
import sys

import pytest
import torch

from oracle import pcvae_oracle as O

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

MIWAE = ["miwae_b12_d6_s4", "miwae_b7_d9_s5", "reg_miwae_b12_d6_s4", "reg_miwae_b7_d9_s5_a06"]


class FeedBSL:
    """Replays recorded [B, S, L] draws in call order (the models draw through VAE.draw_noise_bsl)."""

    def __init__(self, VAE, draws):
        self.VAE, self.draws, self.i = VAE, list(draws), 0

    def __enter__(self):
        self.orig = self.VAE.draw_noise_bsl

        def feed(rows, samples, latent, device, mode):
            e = self.draws[self.i]
            self.i += 1
            assert e.shape == (rows, samples, latent)
            return e.to(device)
        self.VAE.draw_noise_bsl = feed
        return self

    def __exit__(self, *a):
        self.VAE.draw_noise_bsl = self.orig


@pytest.mark.parametrize("R,W", [(1, 3), (37, 10), (700, 13), (65, 50)])
def test_heads_forward_backward(R, W):
    from vae_posterior_consistency_b200 import kernels as KR, lib as L
    g = torch.Generator().manual_seed(R + W)
    sp = torch.nn.functional.softplus
    for mode, C in ((L.MIWAE_HEADS_ENC, 2), (L.MIWAE_HEADS_DEC, 3)):
        raw = (torch.randn(R, C * W, generator=g) * 6).requires_grad_(True)        # covers both softplus regimes
        raw.data[0, :C] = torch.tensor([25.0, -30.0, 21.0][:C])
        if mode == L.MIWAE_HEADS_ENC:
            ref = [raw[:, :W], sp(raw[:, W:])]
        else:
            ref = [torch.sigmoid(raw[:, :W]), sp(raw[:, W:2 * W]) + 0.001, sp(raw[:, 2 * W:]) + 3.0]
        outs = KR.miwae_heads(raw.detach().cuda(), mode)
        for o, r in zip(outs, ref):
            torch.testing.assert_close(o.cpu(), r.detach(), rtol=2e-6, atol=1e-7)
        gs = [torch.randn(R, W, generator=g) for _ in range(C)]
        torch.autograd.backward(ref, gs)
        d_raw = KR.miwae_heads_bwd(raw.detach().cuda(), mode, [t.cuda() for t in gs])
        torch.testing.assert_close(d_raw.cpu(), raw.grad, rtol=1e-5, atol=1e-7)


def test_sample_z_forward_backward():
    from vae_posterior_consistency_b200 import kernels as KR
    g = torch.Generator().manual_seed(3)
    B, S, Lt = 9, 7, 10
    mean = torch.randn(B, Lt, generator=g).requires_grad_(True)
    scale = (torch.rand(B, Lt, generator=g) + 0.1).requires_grad_(True)
    eps = torch.randn(B, S, Lt, generator=g)
    z_ref = mean.unsqueeze(1) + scale.unsqueeze(1) * eps
    dz = torch.randn(B, S, Lt, generator=g)
    z_ref.backward(dz)
    z = KR.miwae_sample_z(mean.detach().cuda(), scale.detach().cuda(), eps.cuda(), S)
    torch.testing.assert_close(z.cpu(), z_ref.detach(), rtol=1e-6, atol=1e-7)
    dm, ds = KR.miwae_sample_z_bwd(dz.cuda(), eps.cuda())
    torch.testing.assert_close(dm.cpu(), mean.grad, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ds.cpu(), scale.grad, rtol=1e-5, atol=1e-6)
    z0 = KR.miwae_sample_z(mean.detach().cuda(), scale.detach().cuda(), None, S)
    assert torch.equal(z0.cpu(), mean.detach().unsqueeze(1).expand(B, S, Lt))


@pytest.mark.parametrize("B,S,D,reg,alpha", [(12, 4, 6, False, 1.0), (7, 5, 9, True, 0.6), (33, 20, 13, True, 1.0),
                                             (5, 300, 50, False, 1.0), (64, 20, 13, True, 0.3)])
def test_loss_kernel_against_the_closed_forms_of_the_oracle(B, S, D, reg, alpha):
    """pcvae_miwae_loss against oracle.miwae_loss / reg_miwae_loss (values) and the closed-form gradients, fp32 and
    fp64 evaluation of the same formulas (the tolerance is the oracle's own fp32-vs-fp64 distance, with a floor)."""
    from vae_posterior_consistency_b200 import kernels as KR
    g = torch.Generator().manual_seed(B * 1000 + S * 10 + D)
    Lt = 10
    x = torch.rand(B, D, generator=g)
    mask = torch.rand(B, D, generator=g) < 0.7
    mask_p = mask & (torch.rand(B, D, generator=g) < 0.6)

    def branch():
        raw = torch.randn(B, S, 3 * D, generator=g)
        mean = torch.randn(B, Lt, generator=g) * 0.5
        scale = torch.rand(B, Lt, generator=g) + 0.2
        return raw, mean, scale, torch.randn(B, S, Lt, generator=g)
    raw_q, mean_q, scale_q, e2_q = branch()
    raw_p, mean_p, scale_p, e2_p = branch()

    def heads(raw):
        sp = torch.nn.functional.softplus
        return torch.sigmoid(raw[..., :D]), sp(raw[..., D:2 * D]) + 0.001, sp(raw[..., 2 * D:]) + 3.0

    def oracle(dt):
        c = lambda t: t.to(dt)
        if reg:
            out = O.reg_miwae_loss_closed_form_grads(c(x), mask, mask_p, c(raw_q), c(mean_q), c(scale_q), c(e2_q), c(raw_p),
                                                     c(mean_p), c(scale_p), c(e2_p), alpha)
            _, xm_imp = O.reg_miwae_loss(c(x), mask, mask_p, (*heads(c(raw_q)), c(mean_q), c(scale_q)),
                                         (*heads(c(raw_p)), c(mean_p), c(scale_p)), c(e2_q), c(e2_p), alpha)
            return out, xm_imp
        out = O.miwae_loss_closed_form_grads(c(x), mask, c(raw_q), c(mean_q), c(scale_q), c(e2_q))
        _, xm_imp, _ = O.miwae_loss(c(x), mask, *heads(c(raw_q)), c(mean_q), c(scale_q), c(e2_q))
        return out, xm_imp
    (o32, imp32), (o64, imp64) = oracle(torch.float32), oracle(torch.float64)
    cu = lambda t: t.cuda()
    hq, hp = [cu(t) for t in heads(raw_q)], [cu(t) for t in heads(raw_p)]
    br = lambda q, p: [q, p] if reg else [q]
    r = KR.miwae_loss(cu(x), cu(mask), cu(mask_p) if reg else None, br(hq[0], hp[0]), br(hq[1], hp[1]), br(hq[2], hp[2]),
                      br(cu(mean_q), cu(mean_p)), br(cu(scale_q), cu(scale_p)), br(cu(e2_q), cu(e2_p)), alpha, reg,
                      want_grads=True, want_imputed=True)
    assert abs(float(r["out"][0]) - float(o64[0])) <= 1e-5 * abs(float(o64[0])) + 1e-6
    torch.testing.assert_close(r["xm_imputed"].cpu().double(), imp64, rtol=1e-4, atol=1e-6)
    # gradients w.r.t. the raw decoder output through the heads' backward, against the oracle's d_raw
    from vae_posterior_consistency_b200 import lib as L
    names = ["d_raw_q", "d_mean_q", "d_scale_q"] + (["d_raw_p", "d_mean_p", "d_scale_p"] if reg else [])
    for bi in range(2 if reg else 1):
        raw = raw_q if bi == 0 else raw_p
        d_raw = KR.miwae_heads_bwd(cu(raw.reshape(-1, 3 * D)), L.MIWAE_HEADS_DEC,
                                   [r["d_xm"][bi].reshape(-1, D), r["d_xs"][bi].reshape(-1, D), r["d_df"][bi].reshape(-1, D)])
        got = [d_raw.view(B, S, 3 * D).cpu(), r["d_mean"][bi].cpu(), r["d_scale"][bi].cpu()]
        for k, gt in enumerate(got):
            ref64, ref32 = o64[1 + 3 * bi + k], o32[1 + 3 * bi + k]
            floor = 2e-5 * float(ref64.abs().max()) + 1e-9
            err = float((gt.double() - ref64).abs().max())
            own = float((ref32.double() - ref64).abs().max())
            print(f"{names[3 * bi + k]}: max |err| {err:.3e} (oracle fp32 vs fp64 {own:.3e}, max |ref| {float(ref64.abs().max()):.3e})")
            assert err <= 4 * own + floor


def test_rowwise_equals_one_call_per_row():
    """eval_miwae stacks the reference's per-row calls (evaluate.py:96-112): rowwise = 1 on a batch must equal B separate
    single-row calls with the reference indexing."""
    from vae_posterior_consistency_b200 import kernels as KR
    g = torch.Generator().manual_seed(5)
    B, S, D, Lt = 6, 17, 8, 10
    x = torch.rand(B, D, generator=g).cuda()
    mask = (torch.rand(B, D, generator=g) < 0.7).cuda()
    sp = torch.nn.functional.softplus
    raw = torch.randn(B, S, 3 * D, generator=g).cuda()
    xm, xs, df = torch.sigmoid(raw[..., :D]).contiguous(), (sp(raw[..., D:2 * D]) + 0.001).contiguous(), (sp(raw[..., 2 * D:]) + 3).contiguous()
    mean, scale = torch.randn(B, Lt, generator=g).cuda(), (torch.rand(B, Lt, generator=g) + 0.2).cuda()
    e2 = torch.randn(B, S, Lt, generator=g).cuda()
    whole = KR.miwae_loss(x, mask, None, [xm], [xs], [df], [mean], [scale], [e2], 1.0, False, rowwise=True, want_imputed=True)
    for b in range(B):
        s = slice(b, b + 1)
        one = KR.miwae_loss(x[s], mask[s], None, [xm[s].contiguous()], [xs[s].contiguous()], [df[s].contiguous()],
                            [mean[s].contiguous()], [scale[s].contiguous()], [e2[s].contiguous()], 1.0, False,
                            rowwise=False, want_imputed=True)
        assert torch.equal(one["xm_imputed"][0], whole["xm_imputed"][b])


@pytest.mark.parametrize("name", MIWAE)
def test_module_api_against_reference_fixture(golden, name):
    """forward / loss / backward / llh_eval of the mirror classes with the reference's recorded noise."""
    from vae_posterior_consistency_b200 import VAE
    g = golden(name)
    reg = g["cls"] == "Reg_MIWAE"
    D, S = g["D"], g["S"]
    B = g["x"].shape[0]
    cls = getattr(VAE, g["cls"])
    model = cls(D, 500, 20, 10, {"batch_size": B, "patience": 100}, S, 10)
    assert list(model.state_dict().keys()) == list(g["state_dict"].keys())
    model.load_state_dict(g["state_dict"])
    model = model.cuda()
    x, mask, mask_p = g["x"].cuda(), g["mask"].cuda(), g["mask_p"].cuda()
    with FeedBSL(VAE, g["draws"]):
        if reg:
            mean_p, scale_p, xm_p, xs_p, df_p, mean_q, scale_q, xm_q, xs_q, df_q = model.forward(x, mask, mask_p)
            pl, loss = model.loss(x, xm_p, xs_p, df_p, mean_p, scale_p, xm_q, xs_q, df_q, mean_q, scale_q, mask, mask_p, 1,
                                  beta_annealing=False, beta=1.0, alpha=g["alpha"])
        else:
            mean_q, scale_q, xm_q, xs_q, df_q = model.forward(x, mask)
            pl, loss = model.loss(x, xm_q, xs_q, df_q, mean_q, scale_q, mask, 1)
    for nm, t in (("mean_q", mean_q), ("scale_q", scale_q), ("xm_q", xm_q), ("xs_q", xs_q), ("df_q", df_q)):
        assert t.shape == g[nm].shape, nm
        torch.testing.assert_close(t.detach().cpu(), g[nm], rtol=1e-4, atol=1e-5, msg=lambda m: f"{nm}: {m}")
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * abs(float(g["loss"])), (float(loss), float(g["loss"]))
    model.zero_grad()
    loss.backward()
    worst = 0.0
    for k, ref in g["grads"].items():
        got = dict(model.named_parameters())[k].grad
        assert got is not None, k
        tol = 2e-5 * float(ref.abs().max()) + 1e-8
        err = float((got.cpu() - ref).abs().max())
        worst = max(worst, err / (float(ref.abs().max()) + 1e-12))
        torch.testing.assert_close(got.cpu(), ref, rtol=2e-3, atol=tol, msg=lambda m: f"grad {k}: {m}")
    print(f"{name}: loss {float(loss):.6f} (reference {float(g['loss']):.6f}), worst gradient error / max |grad| = {worst:.2e}")
    with torch.no_grad(), FeedBSL(VAE, g["eval_draws"]):
        if reg:
            xm_imp, ev_loss, _ = model.loss(x, xm_p, xs_p, df_p, mean_p, scale_p, xm_q, xs_q, df_q, mean_q, scale_q, mask,
                                            mask_p, 1, llh_eval=True, alpha=g["alpha"])
        else:
            xm_imp, ev_loss, imp = model.loss(x, xm_q, xs_q, df_q, mean_q, scale_q, mask, 1, llh_eval=True)
            assert abs(float(imp) - float(g["imp"])) <= 1e-4 * abs(float(g["imp"])) + 1e-7
    torch.testing.assert_close(xm_imp.cpu(), g["xm_imp"], rtol=1e-4, atol=1e-5)
    assert abs(float(ev_loss) - float(g["eval_loss"])) <= 1e-4 * abs(float(g["eval_loss"]))


def test_wide_decoder_head_is_sliced():
    """3 * obs_dim > 128 outputs: the decoder's last layer runs as row slices of its weight matrix."""
    from vae_posterior_consistency_b200 import VAE
    torch.manual_seed(0)
    D, B, S = 50, 5, 3
    model = VAE.MIWAE(D, 500, 20, 10, {"batch_size": B, "patience": 100}, S, 10).cuda()
    z = torch.randn(B, S, 10, device="cuda")
    xm, xs, df = model.decoder(z)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    rm, rs, rd = O.miwae_decoder(sd, z.cpu())
    torch.testing.assert_close(xm.cpu(), rm, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(xs.cpu(), rs, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(df.cpu(), rd, rtol=1e-4, atol=1e-5)
