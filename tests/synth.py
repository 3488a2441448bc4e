"""Deterministic synthetic `Data/` trees in the layout the reference's loaders expect
(SURVEY.md section 8c): data.pt, train_index<i>.csv, test_index<i>.csv, mask_<rate>_missing<i>.pt,
plus the experiments/ output folders the reference never creates itself."""
import os

import numpy as np
import torch


def make_tree(root, data_type, n_rows, obs_dim, seed=0, missing_rate=30, index="1", test_frac=0.25,
              experiment_type="UCI_experiments_consistency_missingness", factors=0):
    """`factors` > 0: columns are noisy mixtures of that many latent factors (features carry information about the
    target column, so the acquisition rewards are well separated); 0: independent columns."""
    g = torch.Generator().manual_seed(seed)
    folder = os.path.join(root, "Data", data_type)
    os.makedirs(folder, exist_ok=True)
    data = torch.randn(n_rows, obs_dim, generator=g) * 2 + 1
    if factors:
        lat = torch.randn(n_rows, factors, generator=g)
        mix = torch.randn(factors, obs_dim, generator=g)
        data = lat @ mix + 0.25 * torch.randn(n_rows, obs_dim, generator=g)
    mask = torch.rand(n_rows, obs_dim, generator=g) < (1 - missing_rate / 100)
    perm = torch.randperm(n_rows, generator=g).numpy()
    n_test = max(1, int(round(n_rows * test_frac)))
    torch.save(data, os.path.join(folder, "data.pt"))
    torch.save(mask, os.path.join(folder, f"mask_{missing_rate}_missing{index}.pt"))
    np.savetxt(os.path.join(folder, f"test_index{index}.csv"), perm[:n_test].astype(np.float64), delimiter=",")
    np.savetxt(os.path.join(folder, f"train_index{index}.csv"), perm[n_test:].astype(np.float64), delimiter=",")
    for kind in ("checkpoints", "rest", "elbos"):
        for fam in ("reg_vae", "vanilla_vae", "reg_EDDI", "vanilla_EDDI"):
            os.makedirs(os.path.join(root, "experiments", experiment_type, data_type, kind, fam), exist_ok=True)
    return folder


DRIVER_CASES = [
    # name, vae_type, K
    ("reg_vae", "reg_vae1", 20),
    ("reg_eddi", "reg_EDDI1", 10),
    ("vanilla_vae", "vanilla_vae1", 20),
    ("vanilla_eddi", "vanilla_EDDI1", 20),
]
#: imputation_args.json lines 16-18 / 31-33 ("*_mask_augm"): Reg_VAE_mask / vanilla_VAE_mask (imputation.py path only)
MASK_DRIVER_CASES = [
    ("reg_vae_mask", "reg_vae1_mask_augm", 20),
    ("vanilla_vae_mask", "vanilla_vae1_mask_augm", 20),
]
DRIVER_CFG = dict(data_type="synth", n_rows=150, obs_dim=6, batch_size=64, epochs=3, M_eval=2, M_al=3,
                  missing_rate=30, p_missingness=30, alpha=1.0, reg_type="kl_reg",
                  experiment_type="UCI_experiments_consistency_missingness")


def make_tree_mnar(root, data_type, n_rows, obs_dim, seed=0, index="1",
                   experiment_type="UCI_experiments_consistency_missingness"):
    """MNAR tree (reference loaders.py:357-384): data.pt has a trailing target column that the loader drops,
    rand_perm<i>.pt, mnar_mask_missing<i>.pt (float32, self-masking: the first D/2 columns are hidden where the
    value exceeds the column mean -- the rule of reference utils.py:48-60)."""
    g = torch.Generator().manual_seed(seed)
    folder = os.path.join(root, "Data", data_type)
    os.makedirs(folder, exist_ok=True)
    data = torch.rand(n_rows, obs_dim + 1, generator=g)
    mask = torch.ones(n_rows, obs_dim + 1)
    half = obs_dim // 2
    mask[:, :half] = (data[:, :half] <= data[:, :half].mean(0)).float()
    torch.save(data, os.path.join(folder, "data.pt"))
    torch.save(mask, os.path.join(folder, f"mnar_mask_missing{index}.pt"))
    torch.save(torch.randperm(n_rows, generator=g), os.path.join(folder, f"rand_perm{index}.pt"))
    for kind in ("checkpoints", "rest", "elbos"):
        for fam in ("reg_notMIWAE", "vanilla_notMIWAE"):
            os.makedirs(os.path.join(root, "experiments", experiment_type, data_type, kind, fam), exist_ok=True)
    return folder


MNAR_CASES = [("reg_notmiwae", "reg_notMIWAE1"), ("vanilla_notmiwae", "vanilla_notMIWAE1")]
MNAR_CFG = dict(data_type="synthmnar", n_rows=40, obs_dim=6, batch_size=16, epochs=2, train_k=4, valid_k=6, M=1,
                missing_rate=30, p_missingness=50, alpha=1.0, reg_type="kl_reg",
                experiment_type="UCI_experiments_consistency_missingness")


# ---- BASELINE.json config sizes, run through the UNMODIFIED driver files (imputation.py, active_learning.py) ----
#: cfg1: Boston-shaped 506 x 13 table, 90/10 split, MCAR 30 %, batch 64; few epochs
CFG1 = dict(data_type="boston_synth", n_rows=506, obs_dim=13, test_frac=0.1, epochs=3, M=2, batch_size=64,
            missing_rate=30, lines=[("reg_vae1", 20), ("vanilla_vae1", 20), ("reg_EDDI1", 10), ("vanilla_EDDI1", 20)])
#: the MIWAE lines of Data/imputation_args.json (train + eval_miwae) on the cfg1 table; S reduced from 20 / 5000
CFG1_MIWAE = dict(data_type="boston_synth", n_rows=506, obs_dim=13, test_frac=0.1, epochs=2, M=1, batch_size=64,
                  missing_rate=30, train_k=5, valid_k=40, lines=[("reg_MIWAE1", 10), ("vanilla_MIWAE1", 10)])
#: cfg3: 2 000 x 20 test set (train part 600 rows), M = 50 reward samples, all 19 acquisition steps
CFG3 = dict(data_type="al_synth", n_rows=2600, obs_dim=20, test_frac=2000 / 2600, epochs=60, M=50, batch_size=64,
            missing_rate=30, factors=3, lines=[("reg_vae1", 10)])


def write_args_json(root, template_line, cfg, lines=None, fname="imputation_args.json", **overrides):
    """Data/imputation_args.json in the reference's JSON-lines format: one experiment per line, every key a
    {"type", "default", "help"} record (utils.py:177-190 builds an argparse parser from it)."""
    import copy
    import json
    out = []
    for vae_type, K in (lines or cfg["lines"]):
        d = copy.deepcopy(template_line)
        d["vae_type"]["default"] = vae_type
        d["K"]["default"] = K
        d["data_type"]["default"] = cfg["data_type"]
        d["missing_rate"]["default"] = cfg["missing_rate"]
        d["epoch"]["default"] = cfg["epochs"]
        d["M"]["default"] = cfg["M"]
        d["batch_size"]["default"] = cfg["batch_size"]
        for k, v in overrides.items():
            d[k]["default"] = v
        out.append(json.dumps(d))
    with open(os.path.join(root, "Data", fname), "w") as f:
        f.write("\n".join(out) + "\n")


def family_dirs(root, data_type, families, experiment_type="UCI_experiments_consistency_missingness"):
    for kind in ("checkpoints", "rest", "elbos"):
        for fam in families:
            os.makedirs(os.path.join(root, "experiments", experiment_type, data_type, kind, fam), exist_ok=True)
