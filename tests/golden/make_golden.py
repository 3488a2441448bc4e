"""Generate golden fixtures from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # needs /root/reference

The reference has no tests, seeds or golden vectors of its own,
so parity is pinned against the reference *executed here*: this script imports
`src.models.VAE` / `src.experiment_main.evaluate` from /root/reference, fixes
seeds, records every input, every N(0,1) draw (`_standard_normal`) and every
output, and writes small `.pt` files next to itself.  The fixtures travel to the
GPU box; /root/reference does not.  Nothing under tests/ or the package reads
/root/reference at test time.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("PCVAE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    sys.path.insert(0, REF)
    # matplotlib is imported (unused) at src/experiment_main/evaluate.py:10 and is not installed
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)
    import src.models.VAE as V
    import src.experiment_main.evaluate as E
    return V, E


class NoiseTape:
    """Records every standard-normal draw Normal.rsample makes (order matters)."""

    def __init__(self):
        import torch.distributions.normal as tdn
        self.tdn = tdn
        self.orig = tdn._standard_normal
        self.draws = []

    def __enter__(self):
        def rec(shape, dtype, device):
            e = self.orig(shape, dtype=dtype, device=device)
            self.draws.append(e.clone())
            return e
        self.tdn._standard_normal = rec
        return self

    def __exit__(self, *a):
        self.tdn._standard_normal = self.orig


def sd_clone(model):
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def synth(B, D, seed, missing=0.3, sub=0.3):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, D, generator=g)
    mask = torch.rand(B, D, generator=g) < (1 - missing)
    mask_p = mask & (torch.rand(B, D, generator=g) < (1 - sub))
    return x, mask, mask_p


def reg_case(V, cls_name, B, D, K, seed, alpha):
    torch.manual_seed(seed)
    np.random.seed(seed)
    cls = getattr(V, cls_name)
    model = cls(D, 500, K, 10, {"batch_size": B, "patience": 100}, "exp", "kl_reg", 1, 10)
    x, mask, mask_p = synth(B, D, seed + 1)
    with NoiseTape() as tape:
        out = model.forward(x, mask, mask_p, stage="train")
    eps_q, eps_p = tape.draws
    mean_p, logvar_p, xh_p, lv_p, mean_q, logvar_q, xh_q, lv_q = out
    _, train_loss = model.loss(x, xh_p, lv_p, mean_p, logvar_p, xh_q, lv_q, mean_q, logvar_q,
                               mask, mask_p, 1, beta_annealing=False, beta=1.0, alpha=alpha,
                               alpha_annealing=True, stage="train")
    model.zero_grad()
    train_loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    with torch.no_grad():
        _, ev_loss, negl, negl_imp = model.loss(x, xh_p, lv_p, mean_p, logvar_p, xh_q, lv_q, mean_q,
                                                logvar_q, mask, mask_p, 1, llh_eval=True, beta=1.0,
                                                alpha=alpha, stage="evaluate")
        rmse = torch.sqrt(torch.sum(torch.square(xh_q * ~mask - x * ~mask)) / torch.sum(~mask))
    return dict(cls=cls_name, D=D, K=K, alpha=alpha, state_dict=sd_clone(model), x=x, mask=mask,
                mask_p=mask_p, eps_q=eps_q, eps_p=eps_p,
                mean_p=mean_p.detach(), logvar_p=logvar_p.detach(), xh_p=xh_p.detach(),
                mean_q=mean_q.detach(), logvar_q=logvar_q.detach(), xh_q=xh_q.detach(),
                x_logvar=lv_q.clone(), train_loss=train_loss.detach(), grads=grads,
                eval_loss=ev_loss, negl=negl, negl_imp=negl_imp, rmse=rmse)


def vanilla_case(V, cls_name, B, D, K, seed):
    torch.manual_seed(seed)
    cls = getattr(V, cls_name)
    model = cls(D, 500, K, 10, {"batch_size": B, "patience": 100}, "exp", 1, 10)
    x, mask, _ = synth(B, D, seed + 1)
    maskf = mask * torch.ones(x.shape)            # train.py:58,97 -> float32 mask
    with NoiseTape() as tape:
        mean_q, logvar_q, xh_q, lv = model.forward(x, maskf)
    (eps_q,) = tape.draws
    _, train_loss = model.loss(x, xh_q, lv, mean_q, logvar_q, 1, maskf, beta_annealing=False,
                               beta=1.0, stage="train")
    model.zero_grad()
    train_loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    with torch.no_grad():
        _, ev_loss, negl, negl_imp = model.loss(x, xh_q, lv, mean_q, logvar_q, 1, mask, llh_eval=True,
                                                beta=1.0, stage="evaluate")
    return dict(cls=cls_name, D=D, K=K, state_dict=sd_clone(model), x=x, mask=mask, eps_q=eps_q,
                mean_q=mean_q.detach(), logvar_q=logvar_q.detach(), xh_q=xh_q.detach(),
                train_loss=train_loss.detach(), grads=grads, eval_loss=ev_loss, negl=negl,
                negl_imp=negl_imp)


def train_traj_case(V, cls_name, B, D, K, seed, steps):
    """A few full reference steps (forward, loss, backward, Adam) with recorded noise."""
    torch.manual_seed(seed)
    cls = getattr(V, cls_name)
    model = cls(D, 500, K, 10, {"batch_size": B, "patience": 100}, "exp", "kl_reg", 1, 10)
    sd0 = sd_clone(model)
    opt = torch.optim.Adam(model.parameters(), lr=0.001)
    xs, ms, mps, eqs, eps_, losses = [], [], [], [], [], []
    for s in range(steps):
        x, mask, mask_p = synth(B, D, seed + 10 + s)
        with NoiseTape() as tape:
            out = model.forward(x, mask, mask_p, stage="train")
        mean_p, logvar_p, xh_p, lv_p, mean_q, logvar_q, xh_q, lv_q = out
        _, loss = model.loss(x, xh_p, lv_p, mean_p, logvar_p, xh_q, lv_q, mean_q, logvar_q, mask,
                             mask_p, s + 1, beta=1.0, alpha=1.0, stage="train")
        opt.zero_grad()
        loss.backward()
        opt.step()
        xs.append(x); ms.append(mask); mps.append(mask_p)
        eqs.append(tape.draws[0]); eps_.append(tape.draws[1]); losses.append(loss.detach())
    return dict(cls=cls_name, D=D, K=K, state_dict0=sd0, state_dict_end=sd_clone(model),
                x=torch.stack(xs), mask=torch.stack(ms), mask_p=torch.stack(mps),
                eps_q=torch.stack(eqs), eps_p=torch.stack(eps_), losses=torch.stack(losses))


def reward_case(V, E, cls_name, N, D, K, M, seed, n_selected):
    torch.manual_seed(seed)
    cls = getattr(V, cls_name)
    model = cls(D, 500, K, 10, {"batch_size": 64, "patience": 100}, "exp", "kl_reg", 1, 10)
    # make the encoder less degenerate than a fresh init so rewards are not ~0
    with torch.no_grad():
        for prm in model.parameters():
            prm.mul_(2.0)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(N, D, generator=g)
    mask = torch.zeros(N, D)
    for n in range(N):                               # rows have different selected sets
        k = int(torch.randint(0, n_selected + 1, (1,), generator=g))
        sel = torch.randperm(D - 1, generator=g)[:k]
        mask[n, sel] = 1.0
    mask_p = torch.zeros(N, D)
    with torch.no_grad():
        with NoiseTape() as tape:
            ims = [model.forward(x, mask, mask_p, stage="evaluate")[6] for _ in range(M)]
        im = torch.stack(ims, 0)
        R = -1e4 * torch.ones(N, D - 1)
        for u in range(D - 1):
            loc = np.where(mask[:, u] == 0)[0]
            R[loc, u] = E.R_lindley_chain(u, x, mask, M, model, im, loc).float()
    return dict(cls=cls_name, D=D, K=K, M=M, state_dict=sd_clone(model), x=x, mask=mask, im=im,
                im_eps_q=torch.stack(tape.draws[0::2]), R=R)


def mnar_case(V, cls_name, B, D, S, seed, alpha):
    """REG_notMIWAE_v2 / notMIWAE_myversion: forward, loss, backward, llh_eval imputation with recorded noise."""
    torch.manual_seed(seed)
    cls = getattr(V, cls_name)
    model = cls(D, 500, 20, 10, {"batch_size": B, "patience": 100}, S, 10)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(B, D, generator=g)
    mask = (torch.rand(B, D, generator=g) < 0.7).float()            # MNAR path uses float32 masks
    mask_p = mask * (torch.rand(B, D, generator=g) < 0.5).float()
    reg = cls_name == "REG_notMIWAE_v2"
    with NoiseTape() as tape:
        if reg:
            mean_p, logvar_p, xm_p, xlv_p, mean_q, logvar_q, xm_q, xlv_q = model.forward(x, mask, mask_p, stage="train")
            _, loss = model.loss(x, xm_p, xlv_p, mean_p, logvar_p, xm_q, xlv_q, mean_q, logvar_q, mask, mask_p, 1,
                                 beta_annealing=False, beta=1.0, alpha=alpha, alpha_annealing=True, stage="train")
        else:
            mean_q, logvar_q, xm_q, xlv_q = model.forward(x, mask)
            _, loss = model.loss(x, xm_q, xlv_q, mean_q, logvar_q, 1, mask, beta_annealing=False, beta=1.0, stage="train")
    draws = [d.clone() for d in tape.draws]
    model.zero_grad()
    loss.backward()
    grads = {k: prm.grad.detach().clone() for k, prm in model.named_parameters() if prm.grad is not None}
    with torch.no_grad():
        if reg:
            xm_imp, _, re = model.loss(x, xm_p, xlv_p, mean_p, logvar_p, xm_q, xlv_q, mean_q, logvar_q, mask, mask_p, 1,
                                       llh_eval=True, alpha=alpha, stage="evaluate")
        else:
            with NoiseTape() as tape2:
                xm_imp, _, re = model.loss(x, xm_q, xlv_q, mean_q, logvar_q, 1, mask, llh_eval=True, stage="evaluate")
            draws.append(tape2.draws[0].clone())
    return dict(cls=cls_name, D=D, S=S, alpha=alpha, state_dict=sd_clone(model), x=x, mask=mask, mask_p=mask_p,
                draws=draws, mean_q=mean_q.detach(), logvar_q=logvar_q.detach(), xm_q=xm_q.detach(),
                xlv_q=xlv_q.detach(), loss=loss.detach(), grads=grads, xm_imp=xm_imp, re=re)


def miwae_case(V, cls_name, B, D, S, seed, alpha):
    """MIWAE / Reg_MIWAE (VAE.py:3011-3301): forward, loss, backward and the llh_eval imputation with recorded noise."""
    torch.manual_seed(seed)
    cls = getattr(V, cls_name)
    model = cls(D, 500, 20, 10, {"batch_size": B, "patience": 100}, S, 10)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(B, D, generator=g)
    mask = torch.rand(B, D, generator=g) < 0.7                       # bool: the losses use ~mask
    mask_p = mask & (torch.rand(B, D, generator=g) < 0.6)
    reg = cls_name == "Reg_MIWAE"
    with NoiseTape() as tape:
        if reg:
            mean_p, scale_p, xm_p, xs_p, df_p, mean_q, scale_q, xm_q, xs_q, df_q = model.forward(x, mask, mask_p)
            _, loss = model.loss(x, xm_p, xs_p, df_p, mean_p, scale_p, xm_q, xs_q, df_q, mean_q, scale_q, mask, mask_p, 1,
                                 beta_annealing=False, beta=1.0, alpha=alpha)
        else:
            mean_q, scale_q, xm_q, xs_q, df_q = model.forward(x, mask)
            _, loss = model.loss(x, xm_q, xs_q, df_q, mean_q, scale_q, mask, 1)
    draws = [d.clone() for d in tape.draws]
    model.zero_grad()
    loss.backward()
    grads = {k: prm.grad.detach().clone() for k, prm in model.named_parameters() if prm.grad is not None}
    # llh_eval redraws the loss-internal noise: record it and the imputation it produces
    with torch.no_grad():
        with NoiseTape() as tape2:
            if reg:
                xm_imp, ev_loss, _ = model.loss(x, xm_p, xs_p, df_p, mean_p, scale_p, xm_q, xs_q, df_q, mean_q, scale_q, mask,
                                                mask_p, 1, llh_eval=True, alpha=alpha)
                imp = None
            else:
                xm_imp, ev_loss, imp = model.loss(x, xm_q, xs_q, df_q, mean_q, scale_q, mask, 1, llh_eval=True)
    return dict(cls=cls_name, D=D, S=S, alpha=alpha, state_dict=sd_clone(model), x=x, mask=mask, mask_p=mask_p, draws=draws,
                eval_draws=[d.clone() for d in tape2.draws], mean_q=mean_q.detach(), scale_q=scale_q.detach(),
                xm_q=xm_q.detach(), xs_q=xs_q.detach(), df_q=df_q.detach(), loss=loss.detach(), grads=grads,
                xm_imp=xm_imp, eval_loss=ev_loss, imp=imp)


def driver_cases(V, E, cases=None, with_al=True):
    """Run the reference's own train() / eval_vae() / active_learning_func() (the call sequence of
    imputation.py:28-59 and active_learning.py:58-74) on a tiny synthetic Data/ tree with fixed seeds
    and record the artefacts they write."""
    import glob
    import tempfile
    import tqdm as tqdm_mod
    sys.path.insert(0, os.path.dirname(HERE))
    from synth import DRIVER_CASES, DRIVER_CFG, make_tree
    import src.experiment_main.train as T
    import src.utils.loaders as LD
    c = DRIVER_CFG
    out = {}
    cwd = os.getcwd()
    for name, vae_type, K in (cases or DRIVER_CASES):
        with tempfile.TemporaryDirectory() as root:
            make_tree(root, c["data_type"], c["n_rows"], c["obs_dim"], seed=0, missing_rate=c["missing_rate"],
                      experiment_type=c["experiment_type"])
            os.chdir(root)
            try:
                torch.manual_seed(0); np.random.seed(0)
                tr, te, obs_dim = LD.data_loader("Data", vae_type, c["missing_rate"], c["batch_size"], c["data_type"])
                losses = []
                orig_write = tqdm_mod.tqdm.write
                tqdm_mod.tqdm.write = staticmethod(lambda s, *a, **k: losses.append(float(s.split("Total Loss:")[1])))
                try:
                    T.train(tr, c["missing_rate"], obs_dim, 500, K, 1, 10, c["data_type"],
                            {"batch_size": c["batch_size"], "patience": 100}, c["experiment_type"], vae_type, 20, 10,
                            c["epochs"], device=torch.device("cpu"), alpha=c["alpha"],
                            p_missingness=c["p_missingness"], reg_type=c["reg_type"])
                finally:
                    tqdm_mod.tqdm.write = orig_write
                E.eval_vae([tr, te], c["missing_rate"], obs_dim, 500, K, c["M_eval"], 10, c["data_type"],
                           {"batch_size": c["batch_size"], "patience": 100}, c["experiment_type"], vae_type,
                           c["epochs"], 5000, 10, device=torch.device("cpu"), alpha=c["alpha"],
                           p_missingness=c["p_missingness"], reg_type=c["reg_type"])
                # active_learning.py:24-74 call sequence (min-max, split, DataLoader, active_learning_func)
                if not with_al:
                    files = {os.path.relpath(f, "experiments"): torch.load(f)
                             for f in glob.glob(os.path.join("experiments", "**", "*.pt"), recursive=True)}
                    out[name] = dict(vae_type=vae_type, K=K, epoch_losses=torch.tensor(losses), files=files)
                    continue
                data = torch.load(os.path.join("Data", c["data_type"], "data.pt"))
                test_idx = np.loadtxt(os.path.join("Data", c["data_type"], "test_index1.csv"), delimiter=",")
                mask = torch.load(os.path.join("Data", c["data_type"], f"mask_{c['missing_rate']}_missing1.pt"))
                norm = (data - data.min(axis=0).values) / (data.max(axis=0).values - data.min(axis=0).values)
                E.active_learning_func(tr[0], norm[test_idx], mask[test_idx], c["missing_rate"], obs_dim, 500, K,
                                       c["M_al"], 10, c["data_type"], {"batch_size": c["batch_size"], "patience": 100},
                                       c["experiment_type"], vae_type, c["epochs"], 5000, 10,
                                       device=torch.device("cpu"), alpha=c["alpha"], p_missingness=c["p_missingness"],
                                       reg_type=c["reg_type"], Repeat=1)
                files = {}
                for f in glob.glob(os.path.join("experiments", "**", "*.pt"), recursive=True):
                    if "im_CHAI" in f:
                        continue
                    files[os.path.relpath(f, "experiments")] = torch.load(f)
                out[name] = dict(vae_type=vae_type, K=K, epoch_losses=torch.tensor(losses), files=files)
            finally:
                os.chdir(cwd)
    return out


def mnar_driver_cases(V, E):
    """imputation_mnar.py:41-85 call sequence (data_loader_mnar -> train -> eval_vae_mnar) with the reference's
    own functions on a tiny synthetic MNAR tree."""
    import glob
    import tempfile
    import tqdm as tqdm_mod
    sys.path.insert(0, os.path.dirname(HERE))
    from synth import MNAR_CASES, MNAR_CFG, make_tree_mnar
    import src.experiment_main.train as T
    import src.utils.loaders as LD
    c = MNAR_CFG
    out = {}
    cwd = os.getcwd()
    for name, vae_type in MNAR_CASES:
        with tempfile.TemporaryDirectory() as root:
            make_tree_mnar(root, c["data_type"], c["n_rows"], c["obs_dim"], seed=3, experiment_type=c["experiment_type"])
            os.chdir(root)
            try:
                torch.manual_seed(0); np.random.seed(0)
                loader, obs_dim = LD.data_loader_mnar("Data", vae_type, c["missing_rate"], c["batch_size"], c["data_type"])
                data = torch.load(os.path.join("Data", c["data_type"], "data.pt"))[:, :-1]
                perm = torch.load(os.path.join("Data", c["data_type"], "rand_perm1.pt")).numpy()
                data = data[perm, :]
                mask = torch.load(os.path.join("Data", c["data_type"], "mnar_mask_missing1.pt"))[:, :-1]
                data = (data - data.min(axis=0).values) / (data.max(axis=0).values - data.min(axis=0).values)
                tp = {"batch_size": c["batch_size"], "patience": 100}
                losses = []
                orig_write = tqdm_mod.tqdm.write
                tqdm_mod.tqdm.write = staticmethod(lambda s, *a, **k: losses.append(float(s.split("Total Loss:")[1])))
                try:
                    T.train(loader, c["missing_rate"], obs_dim, 500, 20, c["M"], 10, c["data_type"], tp,
                            c["experiment_type"], vae_type, c["train_k"], 10, c["epochs"], device=torch.device("cpu"),
                            alpha=c["alpha"], p_missingness=c["p_missingness"], reg_type=c["reg_type"],
                            not_miwae_type="changed")
                finally:
                    tqdm_mod.tqdm.write = orig_write
                E.eval_vae_mnar(data, mask, c["missing_rate"], obs_dim, 500, 20, c["M"], 10, c["data_type"], tp,
                                c["experiment_type"], vae_type, c["epochs"], c["valid_k"], 10,
                                device=torch.device("cpu"), alpha=c["alpha"], p_missingness=c["p_missingness"],
                                reg_type=c["reg_type"], not_miwae_type="changed")
                files = {os.path.relpath(f, "experiments"): torch.load(f)
                         for f in glob.glob(os.path.join("experiments", "**", "*.pt"), recursive=True)}
                out[name] = dict(vae_type=vae_type, epoch_losses=torch.tensor(losses), files=files)
            finally:
                os.chdir(cwd)
    return out


def _run_reference_driver(driver, seed=0):
    """runpy the UNMODIFIED driver file of the reference (CPU), seeds fixed by the harness (the reference sets none),
    epoch losses captured from its tqdm.write lines."""
    import runpy
    import tqdm as tqdm_mod
    losses = []
    orig_write, old_argv = tqdm_mod.tqdm.write, sys.argv
    tqdm_mod.tqdm.write = staticmethod(lambda s, *a, **k: losses.append(float(s.split("Total Loss:")[1])))
    sys.argv = [driver]
    try:
        torch.manual_seed(seed); np.random.seed(seed)
        runpy.run_path(os.path.join(REF, "src", "experiment_main", driver), run_name="__main__")
    finally:
        tqdm_mod.tqdm.write, sys.argv = orig_write, old_argv
    return torch.tensor(losses)


def _template_line():
    import json
    return json.loads(open(os.path.join(REF, "Data", "imputation_args.json")).readline())


def full_size_driver_cases():
    """BASELINE.json configs 1 and 3 through the reference's own, unmodified driver FILES on CPU:
    cfg1  imputation.py on a 506 x 13 table (four model lines in one Data/imputation_args.json, so the RNG stream runs
          across experiments exactly as in the reference), every scalar artefact recorded;
    cfg3  imputation.py (training) then active_learning.py on a 2 000 x 20 test set with M = 50: action_CHAI (uint8),
          the information curve, R_hist_CHAI for the first three acquisition steps on all rows and for every step on
          every 8th row.  im_CHAI (152 MB) is not stored."""
    import glob
    import tempfile
    sys.path.insert(0, os.path.dirname(HERE))
    from synth import CFG1, CFG1_MIWAE, CFG3, family_dirs, make_tree, write_args_json
    out = {}
    cwd = os.getcwd()
    fams = ("reg_vae", "vanilla_vae", "reg_EDDI", "vanilla_EDDI", "reg_MIWAE", "vanilla_MIWAE")
    with tempfile.TemporaryDirectory() as root:                              # MIWAE lines: train() + eval_miwae()
        c = CFG1_MIWAE
        make_tree(root, c["data_type"], c["n_rows"], c["obs_dim"], seed=0, missing_rate=c["missing_rate"], test_frac=c["test_frac"])
        family_dirs(root, c["data_type"], fams)
        write_args_json(root, _template_line(), c, train_k=c["train_k"], valid_k=c["valid_k"])
        os.chdir(root)
        try:
            losses = _run_reference_driver("imputation.py")
            files = {os.path.relpath(f, "experiments"): torch.load(f)
                     for f in glob.glob(os.path.join("experiments", "**", "*.pt"), recursive=True) if "checkpoint" not in f}
            out["cfg1_miwae"] = dict(epoch_losses=losses, files=files)
        finally:
            os.chdir(cwd)
    with tempfile.TemporaryDirectory() as root:
        c = CFG1
        make_tree(root, c["data_type"], c["n_rows"], c["obs_dim"], seed=0, missing_rate=c["missing_rate"], test_frac=c["test_frac"])
        family_dirs(root, c["data_type"], fams)
        write_args_json(root, _template_line(), c)
        os.chdir(root)
        try:
            losses = _run_reference_driver("imputation.py")
            files = {}
            for f in glob.glob(os.path.join("experiments", "**", "*.pt"), recursive=True):
                if "checkpoint" in f:
                    continue
                files[os.path.relpath(f, "experiments")] = torch.load(f)
            out["cfg1"] = dict(epoch_losses=losses, files=files)
        finally:
            os.chdir(cwd)
    with tempfile.TemporaryDirectory() as root:
        c = CFG3
        make_tree(root, c["data_type"], c["n_rows"], c["obs_dim"], seed=1, missing_rate=c["missing_rate"], test_frac=c["test_frac"],
                  factors=c["factors"])
        family_dirs(root, c["data_type"], fams)
        write_args_json(root, _template_line(), c, M=1)                    # imputation.py: M = eval repeats
        os.chdir(root)
        try:
            losses = _run_reference_driver("imputation.py")
            write_args_json(root, _template_line(), c)                       # active_learning.py: M = 50 reward samples
            _run_reference_driver("active_learning.py", seed=1)
            files = {}
            for f in glob.glob(os.path.join("experiments", "**", "*.pt"), recursive=True):
                rel = os.path.relpath(f, "experiments")
                if "im_CHAI" in f:
                    continue
                t = torch.load(f)
                if "checkpoint" in f:        # the weights active_learning.py loaded: the GPU run of the acquisition loop starts
                    files[rel] = t           # from the same ones (the 600-step training trajectory is compared separately)
                    continue
                if "action_CHAI" in f:
                    t = t.to(torch.uint8)
                elif "R_hist_CHAI" in f:                                     # [1, step, row, candidate]
                    top2 = t[0].topk(2, dim=2).values                        # reference's best-minus-second reward per
                    files[rel + "#gap"] = (top2[..., 0] - top2[..., 1]).clone()   # (step, row): which selections are decided
                    files[rel + "#first3"] = t[:, :3].clone()
                    t = t[:, :, ::8].clone()
                    rel = rel + "#rows8"
                elif "information_curve" in f:                               # broadcast over rows: keep row 0
                    t = t[:, :1].clone()
                files[rel] = t
            out["cfg3"] = dict(epoch_losses=losses, files=files)
        finally:
            os.chdir(cwd)
    return out


def main():
    sys.path.insert(0, os.path.dirname(HERE))
    V, E = _import_reference()
    torch.set_num_threads(1)
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]

    class Lazy(dict):                       # name -> fixture; with --only=... the other cases are not even run
        def __setitem__(self, k, thunk):
            if not only or k in only:
                dict.__setitem__(self, k, thunk())
    fx = Lazy()
    fx["reg_vae_b64_d13"] = lambda: reg_case(V, "Reg_VAE", 64, 13, 20, 0, 1.0)
    fx["reg_vae_b37_d20_a05"] = lambda: reg_case(V, "Reg_VAE", 37, 20, 20, 1, 0.5)
    fx["reg_eddi_b64_d13_k20"] = lambda: reg_case(V, "Reg_EDDI", 64, 13, 20, 2, 1.0)
    fx["reg_eddi_b33_d7_k10_a07"] = lambda: reg_case(V, "Reg_EDDI", 33, 7, 10, 3, 0.7)
    fx["vanilla_vae_b64_d13"] = lambda: vanilla_case(V, "vanilla_VAE", 64, 13, 20, 4)
    fx["vanilla_eddi_b64_d13_k20"] = lambda: vanilla_case(V, "vanilla_EDDI", 64, 13, 20, 5)
    fx["traj_reg_vae_b32_d13"] = lambda: train_traj_case(V, "Reg_VAE", 32, 13, 20, 6, 4)
    fx["traj_reg_eddi_b32_d13_k10"] = lambda: train_traj_case(V, "Reg_EDDI", 32, 13, 10, 7, 4)
    fx["reward_reg_vae_n24_d8_m5"] = lambda: reward_case(V, E, "Reg_VAE", 24, 8, 20, 5, 8, 3)
    fx["reward_reg_eddi_n24_d8_k10_m5"] = lambda: reward_case(V, E, "Reg_EDDI", 24, 8, 10, 5, 9, 3)
    fx["mnar_reg_v2_b16_d8_s5"] = lambda: mnar_case(V, "REG_notMIWAE_v2", 16, 8, 5, 20, 1.0)
    fx["mnar_reg_v2_b9_d50_s20_a06"] = lambda: mnar_case(V, "REG_notMIWAE_v2", 9, 50, 20, 21, 0.6)
    fx["mnar_vanilla_b16_d8_s5"] = lambda: mnar_case(V, "notMIWAE_myversion", 16, 8, 5, 22, 1.0)
    # mask-augmented zero-imputation family (VAE.py:510-667, 995-1116): first layer reads [x*mask, mask]
    fx["reg_vae_mask_b64_d13"] = lambda: reg_case(V, "Reg_VAE_mask", 64, 13, 20, 30, 1.0)
    fx["reg_vae_mask_b37_d20_a05"] = lambda: reg_case(V, "Reg_VAE_mask", 37, 20, 20, 31, 0.5)
    fx["vanilla_vae_mask_b64_d13"] = lambda: vanilla_case(V, "vanilla_VAE_mask", 64, 13, 20, 32)
    fx["traj_reg_vae_mask_b32_d13"] = lambda: train_traj_case(V, "Reg_VAE_mask", 32, 13, 20, 33, 4)
    # MIWAE family (Student-t decoder, importance-weighted bound), VAE.py:3011-3301: oracle pinned ahead of its kernels
    fx["miwae_b12_d6_s4"] = lambda: miwae_case(V, "MIWAE", 12, 6, 4, 40, 1.0)
    fx["miwae_b7_d9_s5"] = lambda: miwae_case(V, "MIWAE", 7, 9, 5, 41, 1.0)
    fx["reg_miwae_b12_d6_s4"] = lambda: miwae_case(V, "Reg_MIWAE", 12, 6, 4, 42, 1.0)
    fx["reg_miwae_b7_d9_s5_a06"] = lambda: miwae_case(V, "Reg_MIWAE", 7, 9, 5, 43, 0.6)
    if "--skip-drivers" not in sys.argv:
        from synth import MASK_DRIVER_CASES
        if not only or "drivers_synth_150x6" in only:
            fx["drivers_synth_150x6"] = lambda: driver_cases(V, E)
        if not only or "drivers_mnar_40x6" in only:
            fx["drivers_mnar_40x6"] = lambda: mnar_driver_cases(V, E)
        if not only or "drivers_mask_augm_150x6" in only:
            # imputation.py call sequence only: active_learning.py never selects the mask-augmented family
            fx["drivers_mask_augm_150x6"] = lambda: driver_cases(V, E, MASK_DRIVER_CASES, with_al=False)
    if "--skip-drivers" not in sys.argv and (not only or "drivers_full_size" in only):
        fx["drivers_full_size"] = lambda: full_size_driver_cases()
    for name, d in fx.items():
        if only and name not in only:
            continue
        path = os.path.join(HERE, name + ".pt")
        torch.save(d, path)
        print(f"{name}: {os.path.getsize(path) / 1024:.1f} KB")


if __name__ == "__main__":
    main()
