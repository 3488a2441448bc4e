"""The row-sharded acquisition loop (active_learning_func with torch.distributed initialised, SURVEY.md section 8e;
reference loop src/experiment_main/evaluate.py:387-459) with world_size 2 and 3 against the single-process run:
action_CHAI, R_hist_CHAI, im_CHAI and information_curve_CHAI must be BIT-IDENTICAL.  On a single-GPU box the ranks
share the GPU over gloo (no kernel of this path waits for another rank); with two or more GPUs the same test runs
over NCCL, one GPU per rank."""
import glob
import json
import os
import shutil
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from synth import make_tree  # noqa: E402

EXP = "UCI_experiments_consistency_missingness"


def _artefacts(root):
    out = {}
    for f in glob.glob(os.path.join(root, "experiments", "**", "*CHAI*.pt"), recursive=True):
        out[os.path.basename(f)] = torch.load(f)
    return out


@pytest.mark.parametrize("vae_type,K,world", [("reg_vae1", 20, 2), ("reg_EDDI1", 10, 2), ("vanilla_vae1", 20, 3)])
def test_sharded_acquisition_loop_is_bit_identical_to_one_rank(tmp_path, vae_type, K, world):
    from vae_posterior_consistency_b200 import evaluate, loaders, train as train_mod
    root = str(tmp_path)
    data_type, D, M = "alshard", 8, 4
    make_tree(root, data_type, 700, D, seed=3, test_frac=301 / 700, factors=2)     # 301 test rows: uneven row blocks
    cwd = os.getcwd()
    os.chdir(root)
    try:
        dev = torch.device("cuda:0")
        tp = {"batch_size": 64, "patience": 100}
        torch.manual_seed(0); np.random.seed(0)
        tr, te, obs_dim = loaders.data_loader("Data", vae_type, 30, 64, data_type, device=dev)
        train_mod.train(tr, 30, obs_dim, 500, K, 1, 10, data_type, tp, EXP, vae_type, 20, 10, 8, device=dev, alpha=1.0,
                        p_missingness=30, reg_type="kl_reg")
        data = torch.load(os.path.join("Data", data_type, "data.pt"))
        test_idx = np.loadtxt(os.path.join("Data", data_type, "test_index1.csv"), delimiter=",")
        mask = torch.load(os.path.join("Data", data_type, "mask_30_missing1.pt"))
        norm = (data - data.min(axis=0).values) / (data.max(axis=0).values - data.min(axis=0).values)
        torch.manual_seed(7); np.random.seed(7)
        evaluate.active_learning_func(tr[0], norm[test_idx], mask[test_idx], 30, obs_dim, 500, K, M, 10, data_type, tp, EXP,
                                      vae_type, 8, 5000, 10, device=dev, alpha=1.0, p_missingness=30, reg_type="kl_reg",
                                      Repeat=2)
    finally:
        os.chdir(cwd)
    one = _artefacts(root)
    assert len(one) == 4
    for f in glob.glob(os.path.join(root, "experiments", "**", "*CHAI*.pt"), recursive=True):
        os.remove(f)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    env = dict(os.environ, PCVAE_TEST_BACKEND=backend,
               PCVAE_AL_CFG=json.dumps(dict(root=root, seed=7, vae_type=vae_type, data_type=data_type, K=K, M=M, epochs=8,
                                            experiment_type=EXP, repeat=2)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "al_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "al_worker ok" in out.stdout
    many = _artefacts(root)
    assert set(many) == set(one)
    for name, ref in one.items():
        assert many[name].shape == ref.shape
        assert torch.equal(many[name], ref), f"{name}: {world}-rank run differs from the single-rank run " \
                                             f"(max |diff| {float((many[name] - ref).abs().max()):.3e})"
    act = [v for k, v in one.items() if "action_CHAI" in k][0]
    assert act.shape == (2, 301, D - 1) and len(torch.unique(act[0, 0])) == D - 1      # every candidate picked once per row
