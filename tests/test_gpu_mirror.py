"""The host-side mirror of the reference API (VAE.py modules, train / eval_vae /
active_learning_func, loaders, driver injection) on a B200, against artefacts recorded from
the reference's own functions (tests/golden/make_golden.py: `drivers_synth_150x6`) and the
module-level golden fixtures."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from synth import DRIVER_CASES, DRIVER_CFG, MASK_DRIVER_CASES, make_tree  # noqa: E402


def _pkg():
    from vae_posterior_consistency_b200 import VAE, evaluate, inject, loaders, train
    return VAE, train, evaluate, loaders, inject


class FeedNoise:
    """Replay recorded reference draws through VAE.draw_noise (parity mode feeds host noise)."""

    def __init__(self, VAE, draws):
        self.VAE, self.draws, self.i = VAE, list(draws), 0

    def __enter__(self):
        self.orig = self.VAE.draw_noise

        def feed(rows, latent, device, mode):
            e = self.draws[self.i]
            self.i += 1
            assert e.shape == (rows, latent)
            return e.to(device)
        self.VAE.draw_noise = feed
        return self

    def __exit__(self, *a):
        self.VAE.draw_noise = self.orig


@pytest.mark.parametrize("name", ["reg_vae_b64_d13", "reg_eddi_b64_d13_k20", "reg_eddi_b33_d7_k10_a07",
                                  "reg_vae_b37_d20_a05", "reg_vae_mask_b64_d13", "reg_vae_mask_b37_d20_a05"])
def test_module_api_autograd_matches_reference(golden, name):
    VAE, *_ = _pkg()
    g = golden(name)
    cls = getattr(VAE, g["cls"])
    model = cls(g["D"], 500, g["K"], 10, {"batch_size": 64, "patience": 100}, "exp", "kl_reg", 1, 10)
    model.load_state_dict(g["state_dict"])            # checkpoint interchange: same keys and shapes
    assert list(model.state_dict().keys()) == list(g["state_dict"].keys())
    model.to("cuda")
    x, mask, mask_p = g["x"].cuda(), g["mask"].cuda(), g["mask_p"].cuda()
    with FeedNoise(VAE, [g["eps_q"], g["eps_p"]]):
        out = model.forward(x, mask, mask_p, stage="train")
    mean_p, logvar_p, xh_p, lv_p, mean_q, logvar_q, xh_q, lv_q = out
    for got, key in ((mean_p, "mean_p"), (logvar_p, "logvar_p"), (xh_p, "xh_p"), (mean_q, "mean_q"),
                     (logvar_q, "logvar_q"), (xh_q, "xh_q")):
        torch.testing.assert_close(got.detach().cpu(), g[key], rtol=1e-4, atol=1e-6)
    assert lv_q.shape == (1,) and not lv_q.is_cuda                    # plain CPU attribute as in VAE.py:379
    print_loss, train_loss = model.loss(x, xh_p, lv_p, mean_p, logvar_p, xh_q, lv_q, mean_q, logvar_q, mask, mask_p,
                                        1, beta_annealing=False, beta=1.0, alpha=g["alpha"], alpha_annealing=True,
                                        stage="train")
    torch.testing.assert_close(train_loss.detach().cpu(), g["train_loss"], rtol=1e-4, atol=1e-6)
    model.zero_grad()
    train_loss.backward()
    for k, p in model.named_parameters():
        if k in g["grads"]:
            ref = g["grads"][k]
            torch.testing.assert_close(p.grad.cpu(), ref, rtol=2e-3, atol=2e-5 * float(ref.abs().max() + 1e-3),
                                       msg=lambda m: f"{k}: {m}")
    with torch.no_grad():
        _, ev, negl, negl_imp = model.loss(x, xh_p, lv_p, mean_p, logvar_p, xh_q, lv_q, mean_q, logvar_q, mask,
                                           mask_p, 1, llh_eval=True, beta=1.0, alpha=g["alpha"], stage="evaluate")
    for got, key in ((ev, "eval_loss"), (negl, "negl"), (negl_imp, "negl_imp")):
        torch.testing.assert_close(got.cpu(), g[key], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name", ["vanilla_vae_b64_d13", "vanilla_eddi_b64_d13_k20", "vanilla_vae_mask_b64_d13"])
def test_vanilla_module_api(golden, name):
    VAE, *_ = _pkg()
    g = golden(name)
    model = getattr(VAE, g["cls"])(g["D"], 500, g["K"], 10, {"batch_size": 64, "patience": 100}, "exp", 1, 10)
    model.load_state_dict(g["state_dict"])
    model.to("cuda")
    x = g["x"].cuda()
    maskf = (g["mask"] * torch.ones(g["x"].shape)).cuda()
    with FeedNoise(VAE, [g["eps_q"]]):
        mean_q, logvar_q, xh_q, lv = model.forward(x, maskf)
    _, train_loss = model.loss(x, xh_q, lv, mean_q, logvar_q, 1, maskf, beta_annealing=False, beta=1.0, stage="train")
    torch.testing.assert_close(train_loss.detach().cpu(), g["train_loss"], rtol=1e-4, atol=1e-6)
    train_loss.backward()
    for k, p in model.named_parameters():
        if k in g["grads"]:
            ref = g["grads"][k]
            torch.testing.assert_close(p.grad.cpu(), ref, rtol=2e-3, atol=2e-5 * float(ref.abs().max() + 1e-3))
    with torch.no_grad():
        _, ev, negl, negl_imp = model.loss(x, xh_q, lv, mean_q, logvar_q, 1, g["mask"].cuda(), llh_eval=True,
                                           beta=1.0, stage="evaluate")
    torch.testing.assert_close(ev.cpu(), g["eval_loss"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(negl_imp.cpu(), g["negl_imp"], rtol=1e-4, atol=1e-6)


def test_cpu_tensors_fail_loudly():
    VAE, *_ = _pkg()
    from vae_posterior_consistency_b200.lib import PcvaeError
    model = VAE.Reg_VAE(5, 500, 20, 10, {"batch_size": 8, "patience": 1}, "exp", "kl_reg", 1, 10)
    with pytest.raises(PcvaeError):
        model.encoder(torch.rand(4, 5), torch.ones(4, 5, dtype=torch.bool))


def test_empty_batch_guard():
    VAE, *_ = _pkg()
    model = VAE.Reg_EDDI(5, 500, 10, 10, {"batch_size": 8, "patience": 1}, "exp", "kl_reg", 1, 10).cuda()
    z, m, lv = model.encoder(torch.zeros(0, 5, device="cuda"), torch.zeros(0, 5, device="cuda"))
    assert z.shape == m.shape == lv.shape == (0, 10)                  # VAE.py:723-724


@pytest.mark.parametrize("name,vae_type,K", DRIVER_CASES + MASK_DRIVER_CASES)
def test_driver_sequence_matches_reference_artifacts(golden, tmp_path, name, vae_type, K):
    """imputation.py:28-59 and active_learning.py:58-74 call sequences through the mirror on the GPU, same
    seeds, parity (host-noise) mode; every .pt artefact must match what the reference wrote on CPU."""
    VAE, train_mod, evaluate, loaders, inject = _pkg()
    c = DRIVER_CFG
    mask_augm = "mask_augm" in vae_type              # imputation.py path only: active_learning.py never selects it
    g = golden("drivers_mask_augm_150x6" if mask_augm else "drivers_synth_150x6")[name]
    make_tree(str(tmp_path), c["data_type"], c["n_rows"], c["obs_dim"], seed=0, missing_rate=c["missing_rate"],
              experiment_type=c["experiment_type"])
    cwd = os.getcwd()
    os.chdir(tmp_path)
    dev = torch.device("cuda:0")
    tp = {"batch_size": c["batch_size"], "patience": 100}
    try:
        torch.manual_seed(0); np.random.seed(0)
        tr, te, obs_dim = loaders.data_loader("Data", vae_type, c["missing_rate"], c["batch_size"], c["data_type"],
                                              device=dev)
        import tqdm as tqdm_mod
        losses, orig = [], tqdm_mod.tqdm.write
        tqdm_mod.tqdm.write = staticmethod(lambda s, *a, **k: losses.append(float(s.split("Total Loss:")[1])))
        try:
            train_mod.train(tr, c["missing_rate"], obs_dim, 500, K, 1, 10, c["data_type"], tp, c["experiment_type"],
                            vae_type, 20, 10, c["epochs"], device=dev, alpha=c["alpha"],
                            p_missingness=c["p_missingness"], reg_type=c["reg_type"])
        finally:
            tqdm_mod.tqdm.write = orig
        torch.testing.assert_close(torch.tensor(losses), g["epoch_losses"], rtol=2e-4, atol=1e-5)
        evaluate.eval_vae([tr, te], c["missing_rate"], obs_dim, 500, K, c["M_eval"], 10, c["data_type"], tp,
                          c["experiment_type"], vae_type, c["epochs"], 5000, 10, device=dev, alpha=c["alpha"],
                          p_missingness=c["p_missingness"], reg_type=c["reg_type"])
        data = torch.load(os.path.join("Data", c["data_type"], "data.pt"))
        test_idx = np.loadtxt(os.path.join("Data", c["data_type"], "test_index1.csv"), delimiter=",")
        mask = torch.load(os.path.join("Data", c["data_type"], f"mask_{c['missing_rate']}_missing1.pt"))
        norm = (data - data.min(axis=0).values) / (data.max(axis=0).values - data.min(axis=0).values)
        if not mask_augm:
            evaluate.active_learning_func(tr[0], norm[test_idx], mask[test_idx], c["missing_rate"], obs_dim, 500, K,
                                          c["M_al"], 10, c["data_type"], tp, c["experiment_type"], vae_type,
                                          c["epochs"], 5000, 10, device=dev, alpha=c["alpha"],
                                          p_missingness=c["p_missingness"], reg_type=c["reg_type"], Repeat=1)
        checked = 0
        for rel, ref in g["files"].items():
            path = os.path.join("experiments", rel)
            assert os.path.exists(path), f"missing artefact {rel}"
            got = torch.load(path)
            if isinstance(ref, dict):                                   # checkpoint: same keys, close values
                assert list(got.keys()) == list(ref.keys())
                for k in ref:
                    torch.testing.assert_close(got[k], ref[k], rtol=2e-3, atol=2e-5, msg=lambda m: f"{rel}:{k}: {m}")
            elif "action_CHAI" in rel:
                R_ref = g["files"][rel.replace("action_CHAI", "R_hist_CHAI")]
                top2 = R_ref[0].topk(2, dim=2).values                  # [step, row, 2]
                decided = ((top2[..., 0] - top2[..., 1]) > 2e-5).t()    # [row, step]
                assert torch.equal(got[0][decided], ref[0][decided]), rel
            elif "R_hist_CHAI" in rel:
                torch.testing.assert_close(got, ref, rtol=1e-3, atol=5e-6, msg=lambda m: f"{rel}: {m}")
            else:
                torch.testing.assert_close(got.float(), ref.float(), rtol=5e-4, atol=1e-5, msg=lambda m: f"{rel}: {m}")
            checked += 1
        assert checked >= 8
    finally:
        os.chdir(cwd)


def test_throughput_mode_trains(tmp_path, monkeypatch):
    """PCVAE_MODE=throughput: device-side gather + Philox sub-mask/noise; no bitwise parity, the loss must fall."""
    VAE, train_mod, evaluate, loaders, inject = _pkg()
    c = DRIVER_CFG
    make_tree(str(tmp_path), c["data_type"], 600, 8, seed=1, experiment_type=c["experiment_type"])
    monkeypatch.setenv("PCVAE_MODE", "throughput")
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        torch.manual_seed(0); np.random.seed(0)
        dev = torch.device("cuda:0")
        tr, te, obs_dim = loaders.data_loader("Data", "reg_vae1", 30, 64, c["data_type"], device=dev)
        import tqdm as tqdm_mod
        losses, orig = [], tqdm_mod.tqdm.write
        tqdm_mod.tqdm.write = staticmethod(lambda s, *a, **k: losses.append(float(s.split("Total Loss:")[1])))
        try:
            train_mod.train(tr, 30, obs_dim, 500, 20, 1, 10, c["data_type"], {"batch_size": 64, "patience": 1},
                            c["experiment_type"], "reg_vae1", 20, 10, 25, device=dev, alpha=1.0, p_missingness=30,
                            reg_type="kl_reg")
        finally:
            tqdm_mod.tqdm.write = orig
        assert losses[-1] < losses[0] and all(np.isfinite(losses))
    finally:
        os.chdir(cwd)
