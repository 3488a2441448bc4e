"""not-MIWAE MNAR family (REG_notMIWAE_v2 / notMIWAE_myversion, reference VAE.py:2327-2505, 2691-2847) on a
B200: generic dense kernels, the (row, sample) loss kernels, the module API with autograd against fixtures
recorded from the reference, and the imputation_mnar.py call sequence against the reference's artefacts."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import pcvae_oracle as O

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from synth import MNAR_CASES, MNAR_CFG, make_tree_mnar  # noqa: E402


def _act_ref(v, act):
    return [lambda t: t, torch.relu, torch.sigmoid, O.elu, lambda t: torch.clamp(t, -10.0, 0.0)][act](v)


@pytest.mark.parametrize("R,K,N", [(70, 50, 128), (1, 10, 128), (130, 128, 128), (65, 128, 50), (64, 128, 10), (33, 7, 5)])
@pytest.mark.parametrize("act", [0, 1, 2, 3, 4])
def test_dense_layer_forward_backward(R, K, N, act):
    from vae_posterior_consistency_b200 import kernels as KR
    g = torch.Generator().manual_seed(R + K + N + act)
    x = torch.randn(R, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) * (3.0 if act == 4 else 0.3) - (3.0 if act == 4 else 0.0)
    mask = (torch.rand(R, K, generator=g) < 0.7).float() if K == 50 else None
    dy = torch.randn(R, N, generator=g)
    xr, Wr, br = x.clone().requires_grad_(True), W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y_ref = _act_ref((xr * mask if mask is not None else xr) @ Wr.t() + br, act)
    y_ref.backward(dy)
    y = KR.dense_fwd(x.cuda(), W.cuda(), b.cuda(), act, None if mask is None else mask.cuda())
    torch.testing.assert_close(y.cpu(), y_ref.detach(), rtol=1e-4, atol=1e-5)
    dx, dW, db = KR.dense_bwd(x.cuda(), W.cuda(), y, dy.cuda(), act, None if mask is None else mask.cuda())
    torch.testing.assert_close(dx.cpu(), xr.grad, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(dW.cpu(), Wr.grad, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(db.cpu(), br.grad, rtol=1e-3, atol=1e-4)


class FeedBSL:
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


@pytest.mark.parametrize("name", ["mnar_reg_v2_b16_d8_s5", "mnar_reg_v2_b9_d50_s20_a06", "mnar_vanilla_b16_d8_s5"])
def test_mnar_module_api_matches_reference(golden, name):
    from vae_posterior_consistency_b200 import VAE
    g = golden(name)
    reg = g["cls"] == "REG_notMIWAE_v2"
    model = getattr(VAE, g["cls"])(g["D"], 500, 20, 10, {"batch_size": 16, "patience": 1}, g["S"], 10)
    assert list(model.state_dict().keys()) == list(g["state_dict"].keys())
    model.load_state_dict(g["state_dict"])
    model.cuda()
    x, mask, mask_p = g["x"].cuda(), g["mask"].cuda(), g["mask_p"].cuda()
    with FeedBSL(VAE, g["draws"]):
        if reg:
            mean_p, logvar_p, xm_p, xlv_p, mean_q, logvar_q, xm_q, xlv_q = model.forward(x, mask, mask_p, stage="train")
            _, loss = model.loss(x, xm_p, xlv_p, mean_p, logvar_p, xm_q, xlv_q, mean_q, logvar_q, mask, mask_p, 1,
                                 beta_annealing=False, beta=1.0, alpha=g["alpha"], alpha_annealing=True, stage="train")
        else:
            mean_q, logvar_q, xm_q, xlv_q = model.forward(x, mask)
            _, loss = model.loss(x, xm_q, xlv_q, mean_q, logvar_q, 1, mask, beta_annealing=False, beta=1.0, stage="train")
        assert mean_q.shape == g["mean_q"].shape
        torch.testing.assert_close(mean_q.cpu(), g["mean_q"], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(xm_q.detach().cpu(), g["xm_q"], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(xlv_q.detach().cpu(), g["xlv_q"], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(loss.detach().cpu(), g["loss"], rtol=1e-4, atol=1e-6)
        model.zero_grad()
        loss.backward()
        for k, prm in model.named_parameters():
            if k in g["grads"]:
                ref = g["grads"][k]
                torch.testing.assert_close(prm.grad.cpu(), ref, rtol=2e-3, atol=2e-5 * float(ref.abs().max() + 1e-3),
                                           msg=lambda m: f"{k}: {m}")
            else:
                assert k.startswith("logits.") and prm.grad is None
        with torch.no_grad():
            if reg:
                xm_imp, _, re = model.loss(x, xm_p, xlv_p, mean_p, logvar_p, xm_q, xlv_q, mean_q, logvar_q, mask,
                                           mask_p, 1, llh_eval=True, alpha=g["alpha"], stage="evaluate")
            else:
                xm_imp, _, re = model.loss(x, xm_q, xlv_q, mean_q, logvar_q, 1, mask, llh_eval=True, stage="evaluate")
        torch.testing.assert_close(xm_imp.cpu(), g["xm_imp"], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(re.cpu(), g["re"], rtol=1e-4, atol=1e-6)


def test_mnar_loss_kernel_vs_oracle_large_sample_count():
    """cfg2-like shape (B=128, D=50, S=20) and the evaluation shape (B=3, S=2000): loss + gradients vs the oracle."""
    from vae_posterior_consistency_b200 import kernels as KR
    for B, D, S in ((128, 50, 20), (3, 50, 2000)):
        p = O.init_mnar_params(D, seed=B)
        g = torch.Generator().manual_seed(S)
        x = torch.rand(B, D, generator=g)
        mask = (torch.rand(B, D, generator=g) < 0.6).float()
        mask_p = mask * (torch.rand(B, D, generator=g) < 0.5).float()
        eq, ep = torch.randn(B, S, 10, generator=g), torch.randn(B, S, 10, generator=g)
        mu_q, lv_q = O.mnar_encoder_stats(p, x, mask)
        mu_p, lv_p = O.mnar_encoder_stats(p, x, mask_p)
        xm_q, xlv_q = O.mnar_decoder(p, mu_q.unsqueeze(1) + eq * torch.exp(lv_q / 2).unsqueeze(1))
        xm_p, xlv_p = O.mnar_decoder(p, mu_p.unsqueeze(1) + ep * torch.exp(lv_p / 2).unsqueeze(1))
        leaves = [t.detach().clone().requires_grad_(True) for t in (xm_q, xlv_q, mu_q, lv_q, xm_p, xlv_p, mu_p, lv_p)]
        pp = dict(p); pp["W"] = p["W"].clone().requires_grad_(True); pp["b"] = p["b"].clone().requires_grad_(True)
        loss, xm_imp, re = O.mnar_reg_loss(pp, x, mask, mask_p, leaves[2], leaves[3], leaves[6], leaves[7], leaves[0],
                                           leaves[1], leaves[4], leaves[5], alpha=0.7)
        ref = torch.autograd.grad(loss, leaves + [pp["W"], pp["b"]])
        cu = lambda t: t.detach().cuda()
        r = KR.mnar_loss(cu(x), cu(mask), cu(mask_p), [cu(xm_q), cu(xm_p)], [cu(xlv_q), cu(xlv_p)], [cu(mu_q), cu(mu_p)],
                         [cu(lv_q), cu(lv_p)], cu(p["W"]), cu(p["b"]), 0.7, True, want_grads=True, want_imputed=True)
        torch.testing.assert_close(r["out"][0].float().cpu(), loss.detach(), rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(r["out"][1].float().cpu(), re.detach(), rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(r["xm_imputed"].cpu(), xm_imp.detach(), rtol=1e-4, atol=1e-6)
        got = [r["d_xm"][0], r["d_xlv"][0], r["d_mean"][0], r["d_logvar"][0], r["d_xm"][1], r["d_xlv"][1], r["d_mean"][1],
               r["d_logvar"][1], r["d_W"].view(1, 1, -1), r["d_b"].view(1, 1, -1)]
        for i, (a, b) in enumerate(zip(got, ref)):
            torch.testing.assert_close(a.cpu(), b, rtol=2e-3, atol=2e-5 * float(b.abs().max() + 1e-6),
                                       msg=lambda m: f"grad {i} (B={B},S={S}): {m}")


@pytest.mark.parametrize("name,vae_type", MNAR_CASES)
def test_mnar_driver_sequence_matches_reference_artifacts(golden, tmp_path, name, vae_type):
    from vae_posterior_consistency_b200 import evaluate, loaders, train as train_mod
    c = MNAR_CFG
    g = golden("drivers_mnar_40x6")[name]
    make_tree_mnar(str(tmp_path), c["data_type"], c["n_rows"], c["obs_dim"], seed=3, experiment_type=c["experiment_type"])
    cwd = os.getcwd()
    os.chdir(tmp_path)
    dev = torch.device("cuda:0")
    tp = {"batch_size": c["batch_size"], "patience": 100}
    try:
        torch.manual_seed(0); np.random.seed(0)
        loader, obs_dim = loaders.data_loader_mnar("Data", vae_type, c["missing_rate"], c["batch_size"], c["data_type"],
                                                   device=dev)
        data = torch.load(os.path.join("Data", c["data_type"], "data.pt"))[:, :-1]
        perm = torch.load(os.path.join("Data", c["data_type"], "rand_perm1.pt")).numpy()
        data = data[perm, :]
        mask = torch.load(os.path.join("Data", c["data_type"], "mnar_mask_missing1.pt"))[:, :-1]
        data = (data - data.min(axis=0).values) / (data.max(axis=0).values - data.min(axis=0).values)
        import tqdm as tqdm_mod
        losses, orig = [], tqdm_mod.tqdm.write
        tqdm_mod.tqdm.write = staticmethod(lambda s, *a, **k: losses.append(float(s.split("Total Loss:")[1])))
        try:
            train_mod.train(loader, c["missing_rate"], obs_dim, 500, 20, c["M"], 10, c["data_type"], tp,
                            c["experiment_type"], vae_type, c["train_k"], 10, c["epochs"], device=dev, alpha=c["alpha"],
                            p_missingness=c["p_missingness"], reg_type=c["reg_type"], not_miwae_type="changed")
        finally:
            tqdm_mod.tqdm.write = orig
        torch.testing.assert_close(torch.tensor(losses), g["epoch_losses"], rtol=2e-4, atol=1e-5)
        evaluate.eval_vae_mnar(data, mask, c["missing_rate"], obs_dim, 500, 20, c["M"], 10, c["data_type"], tp,
                               c["experiment_type"], vae_type, c["epochs"], c["valid_k"], 10, device=dev,
                               alpha=c["alpha"], p_missingness=c["p_missingness"], reg_type=c["reg_type"],
                               not_miwae_type="changed")
        assert len(g["files"]) == 2
        for rel, ref in g["files"].items():
            got = torch.load(os.path.join("experiments", rel))
            if isinstance(ref, dict):
                assert list(got.keys()) == list(ref.keys())
                for k in ref:
                    assert got[k].dtype == ref[k].dtype, k
                    torch.testing.assert_close(got[k], ref[k], rtol=2e-3, atol=2e-5, msg=lambda m: f"{rel}:{k}: {m}")
            else:
                torch.testing.assert_close(got.float(), ref.float(), rtol=5e-4, atol=1e-5, msg=lambda m: f"{rel}: {m}")
    finally:
        os.chdir(cwd)


@pytest.mark.parametrize("cls", ["REG_notMIWAE_v2", "notMIWAE_myversion"])
def test_graph_replayed_step_equals_eager_step(cls):
    """graphed.py: forward + loss + backward + Adam replayed from a CUDA graph gives the same parameters as the
    same steps launched op by op (same batches, same noise, same capturable Adam)."""
    from vae_posterior_consistency_b200 import VAE
    from vae_posterior_consistency_b200.graphed import GraphedTrainer
    dev = torch.device("cuda")
    D, B, S = 12, 24, 5
    reg = cls == "REG_notMIWAE_v2"

    def run(eager_first, n_steps=7):
        torch.manual_seed(3)
        model = getattr(VAE, cls)(D, 500, 20, 10, {"batch_size": B, "patience": 100}, S, 10).to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)
        gn = torch.Generator(device=dev).manual_seed(11)
        gb = torch.Generator(device=dev).manual_seed(12)

        if reg:
            def fwd_loss(x, m, mp):
                mean_p, logvar_p, xm_p, xlv_p, mean_q, logvar_q, xm_q, xlv_q = model.forward(x, m, mp, stage="train")
                return model.loss(x, xm_p, xlv_p, mean_p, logvar_p, xm_q, xlv_q, mean_q, logvar_q, m, mp, 1, alpha=0.7)[1]
        else:
            def fwd_loss(x, m):
                mean, logvar, xm, xlv = model.forward(x, m)
                return model.loss(x, xm, xlv, mean, logvar, 1, m)[1]
        gt = GraphedTrainer(model, lambda: fwd_loss, opt, lambda t: t.normal_(generator=gn), eager_first=eager_first)
        losses = []
        for s in range(n_steps):
            rows = B if s != 4 else B - 5                      # one ragged batch in the middle (runs eagerly)
            x = torch.rand(rows, D, device=dev, generator=gb)
            m = (torch.rand(rows, D, device=dev, generator=gb) < 0.7).float()
            mp = m * (torch.rand(rows, D, device=dev, generator=gb) < 0.5).float()
            losses.append(float(gt.step(x, m, mp) if reg else gt.step(x, m)))
        return model, losses, gt.replays

    m_eager, l_eager, r0 = run(10 ** 9)
    m_graph, l_graph, r1 = run(2)
    assert r0 == 0 and r1 == 4                                  # steps 2, 3, 5, 6 were graph replays
    np.testing.assert_allclose(l_graph, l_eager, rtol=1e-6)
    for (k, a), (_, b) in zip(m_eager.state_dict().items(), m_graph.state_dict().items()):
        torch.testing.assert_close(b, a, rtol=1e-6, atol=1e-7, msg=k)


@pytest.mark.parametrize("cls,N,D,S", [("REG_notMIWAE_v2", 130, 50, 10000), ("notMIWAE_myversion", 100, 50, 10000),
                                       ("REG_notMIWAE_v2", 7, 6, 5), ("notMIWAE_myversion", 300, 13, 77),
                                       ("REG_notMIWAE_v2", 3, 64, 1000)])
def test_one_pass_imputation_at_full_valid_k(cls, N, D, S):
    """pcvae_mnar_impute (one pass over [rows, samples], online softmax) at the reference's
    valid_k = 10 000 on >= 100 rows: against the row-blocked path through the generic dense / loss kernels with the
    same noise, and against the oracle in float64 on a few rows."""
    from vae_posterior_consistency_b200 import VAE, kernels as KR, ops
    torch.manual_seed(5)
    dev = torch.device("cuda")
    model = getattr(VAE, cls)(D, 500, 20, 10, {"batch_size": 64, "patience": 100}, S, 10).to(dev)
    with torch.no_grad():                                     # spread the importance weights: sharper decoder, real slopes
        for prm in model.parameters():
            prm.mul_(1.7)
    reg = cls == "REG_notMIWAE_v2"
    g = torch.Generator().manual_seed(N + D + S)
    x = torch.rand(N, D, generator=g).to(dev)
    mask = (torch.rand(N, D, generator=g) < 0.6).float().to(dev)
    eps = torch.randn(N, S, 10, generator=g).to(dev)
    eps_kl = torch.randn(N, S, 10, generator=g).to(dev)
    with torch.no_grad():
        mean, logvar = model._stats(x, mask)
        got = KR.mnar_impute(model, x, mask, mean, logvar, S, reg, eps=eps, eps_kl=None if reg else eps_kl)
        # row-blocked cross-check (materialises [block, S, D])
        ref = torch.empty_like(got)
        block = max(1, (1 << 22) // (S * D))
        for lo in range(0, N, block):
            hi = min(N, lo + block)
            z = ops.mnar_sample_z_op(mean[lo:hi], logvar[lo:hi], eps[lo:hi].contiguous(), S)
            xm, xlv = model.decoder(z)
            if reg:
                out = ops.mnar_loss_op(x[lo:hi], mask[lo:hi], mask[lo:hi], xm, xlv, xm, xlv, mean[lo:hi], logvar[lo:hi],
                                       mean[lo:hi], logvar[lo:hi], model.W, model.b, None, 1.0, False, True)
            else:
                out = ops.mnar_loss_op(x[lo:hi], mask[lo:hi], None, xm, xlv, None, None, mean[lo:hi], logvar[lo:hi], None,
                                       None, model.W, model.b, eps_kl[lo:hi].contiguous(), 1.0, False, True)
            ref[lo:hi] = out[2]
    err = float((got - ref).abs().max())
    print(f"{cls} N={N} D={D} S={S}: max |one-pass - blocked| = {err:.3e}")
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=2e-6)
    # oracle in float64 on the first rows (the reference formulas, oracle/pcvae_oracle.py)
    k = min(N, 4)
    sd = {kk: v.detach().cpu().double() for kk, v in model.state_dict().items() if v.dtype.is_floating_point}
    xs, ms = x[:k].cpu().double(), mask[:k].cpu().double()
    mu, lv = O.mnar_encoder_stats(sd, xs, ms)
    zz = mu.unsqueeze(1) + torch.exp(lv / 2).unsqueeze(1) * eps[:k].cpu().double()
    xm64, xlv64 = O.mnar_decoder(sd, zz)
    if reg:
        _, imp64, _ = O.mnar_reg_loss(sd, xs, ms, ms, mu, lv, mu, lv, xm64, xlv64, xm64, xlv64, alpha=1.0)
    else:
        _, imp64, _ = O.mnar_vanilla_loss(sd, xs, ms, mu, lv, xm64, xlv64, eps_kl[:k].cpu().double())
    torch.testing.assert_close(got[:k].cpu().double(), imp64, rtol=2e-4, atol=1e-5)


def test_one_pass_imputation_with_device_noise_is_statistically_the_same():
    """Throughput mode: the kernel draws its own noise (Philox); with S = 10 000 samples the importance-weighted means of
    two independent noise streams agree to Monte-Carlo accuracy, and the call is deterministic."""
    from vae_posterior_consistency_b200 import VAE, kernels as KR
    torch.manual_seed(6)
    dev = torch.device("cuda")
    N, D, S = 64, 20, 10000
    model = VAE.notMIWAE_myversion(D, 500, 20, 10, {"batch_size": 64, "patience": 100}, S, 10).to(dev)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(N, D, generator=g).to(dev)
    mask = (torch.rand(N, D, generator=g) < 0.6).float().to(dev)
    with torch.no_grad():
        mean, logvar = model._stats(x, mask)
        a = KR.mnar_impute(model, x, mask, mean, logvar, S, False, seed=11)
        a2 = KR.mnar_impute(model, x, mask, mean, logvar, S, False, seed=11)
        b = KR.mnar_impute(model, x, mask, mean, logvar, S, False, seed=12)
        h = KR.mnar_impute(model, x, mask, mean, logvar, S, False, eps=torch.randn(N, S, 10, device=dev),
                           eps_kl=torch.randn(N, S, 10, device=dev))
    assert torch.equal(a, a2) and not torch.equal(a, b)
    assert float((a - b).abs().max()) < 0.02 and float((a - h).abs().max()) < 0.02
    assert float((a - h).abs().mean()) < 0.004
