"""Worker of tests/test_gpu_al_sharded.py: one rank of a row-sharded active_learning_func run (launched by
torch.distributed.run).  Backend from PCVAE_TEST_BACKEND: "gloo" lets two ranks share ONE GPU (the acquisition loop has
no kernel that waits for another rank: the reward needs no collective, the per-step scalar all-reduce and the final
merge go through torch.distributed), "nccl" is the real thing on two GPUs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    import json
    cfg = json.loads(os.environ["PCVAE_AL_CFG"])
    backend = os.environ.get("PCVAE_TEST_BACKEND", "gloo")
    local = int(os.environ["LOCAL_RANK"])
    ndev = torch.cuda.device_count()
    dev = torch.device("cuda", local % ndev)
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    from vae_posterior_consistency_b200 import evaluate, loaders
    os.chdir(cfg["root"])
    torch.manual_seed(cfg["seed"]); np.random.seed(cfg["seed"])          # every rank draws the same host noise
    tr, te, obs_dim = loaders.data_loader("Data", cfg["vae_type"], 30, 64, cfg["data_type"], device=dev)
    data = torch.load(os.path.join("Data", cfg["data_type"], "data.pt"))
    test_idx = np.loadtxt(os.path.join("Data", cfg["data_type"], "test_index1.csv"), delimiter=",")
    mask = torch.load(os.path.join("Data", cfg["data_type"], "mask_30_missing1.pt"))
    norm = (data - data.min(axis=0).values) / (data.max(axis=0).values - data.min(axis=0).values)
    evaluate.active_learning_func(tr[0], norm[test_idx], mask[test_idx], 30, obs_dim, 500, cfg["K"], cfg["M"], 10,
                                  cfg["data_type"], {"batch_size": 64, "patience": 100}, cfg["experiment_type"],
                                  cfg["vae_type"], cfg["epochs"], 5000, 10, device=dev, alpha=1.0, p_missingness=30,
                                  reg_type="kl_reg", Repeat=cfg["repeat"])
    dist.barrier()
    if dist.get_rank() == 0:
        print(f"al_worker ok: world {dist.get_world_size()} backend {backend}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
