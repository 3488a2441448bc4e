"""Mirror of the reference's `src/utils/loaders.py` for the in-scope families: same function
names, positional signatures, return values, checkpoint naming."""
import os

import torch
from numpy import loadtxt
from torch.utils.data import DataLoader

from .VAE import (MIWAE, REG_notMIWAE_v2, Reg_EDDI, Reg_MIWAE, Reg_VAE, Reg_VAE_mask, notMIWAE_myversion, vanilla_EDDI,
                  vanilla_VAE, vanilla_VAE_mask)

_OUT_OF_SCOPE = ("flow",)


def _strip_digits(s):
    return ''.join(c for c in s if not c.isdigit())


def checkpoint_path(experiment_type, data_type, vae_type, missing_rate, alpha, p_missingness, reg_type, family_dir):
    """File the reference writes at train.py:120-131 and reads back at loaders.py:81-88,178-183,129-134,213-218."""
    base = os.path.join('experiments', experiment_type, data_type, 'checkpoints', family_dir)
    if 'vanilla' in vae_type:
        return os.path.join(base, f'checkpoint_{vae_type}_{missing_rate}_missing_rate_test.pt')
    return os.path.join(base, f'checkpoint_{vae_type}_{alpha}_{p_missingness}_{reg_type}_{missing_rate}'
                              '_missing_rate_full_reg_test.pt')


def model_loader(stage, obs_dim, hid_dim, K, latent_dim, missing_rate, data_type, training_parameters, max_epochs,
                 num_samples, num_estimates, experiment_type, reg_type, vae_type='vae', alpha=1.0, p_missingness=30,
                 beta=0.5, beta_annealing=True, alpha_annealing=True, not_miwae_type='changed'):
    """Substring dispatch vae_type -> class in the reference's order (loaders.py:19-245); stage 'train' gives a
    fresh model, anything else loads the checkpoint written by train()."""
    for tag in _OUT_OF_SCOPE:
        if tag in vae_type:
            raise NotImplementedError(f"vae_type {vae_type!r} selects a family outside the B200 hot path "
                                      "(SURVEY.md section 8f); use the reference's eager implementation")
    if 'reg_vae' in vae_type:
        # loaders.py:51-89: '*_mask_augm' selects the mask-augmented class and always loads from .../reg_vae/
        cls = Reg_VAE_mask if 'mask_augm' in vae_type else Reg_VAE
        model = cls(obs_dim, hid_dim, K, latent_dim, training_parameters, experiment_type, reg_type, num_samples,
                    num_estimates)
        load_dir = 'reg_vae' if 'mask_augm' in vae_type else _strip_digits(vae_type)
    elif 'reg_notMIWAE' in vae_type:
        model = REG_notMIWAE_v2(obs_dim, hid_dim, K, latent_dim, training_parameters, num_samples, num_estimates)
        load_dir = _strip_digits(vae_type)
    elif 'reg_EDDI' in vae_type:
        if data_type == 'mnist':
            raise NotImplementedError("the MNIST variants are outside the B200 hot path (SURVEY.md section 2 #16)")
        model = Reg_EDDI(obs_dim, hid_dim, K, latent_dim, training_parameters, experiment_type, reg_type, num_samples,
                         num_estimates)
        load_dir = _strip_digits(vae_type)
    elif 'reg_MIWAE' in vae_type:                                                    # loaders.py:135-147
        model = Reg_MIWAE(obs_dim, hid_dim, K, latent_dim, training_parameters, num_samples, num_estimates)
        load_dir = _strip_digits(vae_type)
    elif 'vanilla_vae' in vae_type:
        cls = vanilla_VAE_mask if 'mask_augm' in vae_type else vanilla_VAE          # loaders.py:149-184
        model = cls(obs_dim, hid_dim, K, latent_dim, training_parameters, experiment_type, num_samples,
                    num_estimates)
        load_dir = 'vanilla_vae'
    elif 'vanilla_EDDI' in vae_type:
        if data_type == 'mnist':
            raise NotImplementedError("the MNIST variants are outside the B200 hot path (SURVEY.md section 2 #16)")
        model = vanilla_EDDI(obs_dim, hid_dim, K, latent_dim, training_parameters, experiment_type, num_samples,
                             num_estimates)
        load_dir = 'vanilla_EDDI'
    elif 'vanilla_notMIWAE' in vae_type:
        model = notMIWAE_myversion(obs_dim, hid_dim, K, latent_dim, training_parameters, num_samples, num_estimates)
        load_dir = _strip_digits(vae_type)
    else:                                                                            # loaders.py:234-245: everything else is MIWAE
        model = MIWAE(obs_dim, hid_dim, K, latent_dim, training_parameters, num_samples, num_estimates)
        load_dir = _strip_digits(vae_type)
    if stage == 'train':
        print("Initializing fresh model")
    else:
        print("Loading saved model")
        path = checkpoint_path(experiment_type, data_type, vae_type, missing_rate, alpha, p_missingness, reg_type,
                               load_dir)
        model.load_state_dict(torch.load(path, map_location=torch.device('cpu')))
    return model


class ConcatDataset(torch.utils.data.Dataset):
    """Zip of equally indexed tensors (reference loaders.py:389-397)."""

    def __init__(self, *datasets):
        self.datasets = datasets

    def __getitem__(self, i):
        return tuple(d[i] for d in self.datasets)

    def __len__(self):
        return min(len(d) for d in self.datasets)


def _minmax(data, data_transform):
    if data_transform == 'minmax':
        lo, hi = data.min(axis=0).values, data.max(axis=0).values
        return (data - lo) / (hi - lo)
    data = data - data.mean(0)
    return data / data.std(0)


def data_loader(data_path, vae_type, missing_rate, batch_size, data_type, device=torch.device('cpu'), shuffle=True,
                data_transform='minmax'):
    """reference loaders.py:319-354: ([train_loader,'train'], [test_loader,'test'], obs_dim).  The whole table is
    moved to `device`; the loaders additionally expose it as `.pcvae_table` so the fused trainer can gather
    batches on the device from the same sampler permutation."""
    index = [c for c in vae_type if c.isdigit()][0]
    folder = os.path.join(data_path, data_type)
    train_idx = loadtxt(os.path.join(folder, f'train_index{index}.csv'), delimiter=',')
    test_idx = loadtxt(os.path.join(folder, f'test_index{index}.csv'), delimiter=',')
    data = torch.load(os.path.join(folder, 'data.pt'))
    mask = torch.load(os.path.join(folder, f'mask_{missing_rate}_missing{index}.pt'))
    data = _minmax(data, data_transform)
    out = []
    for idx, name in ((train_idx, 'train'), (test_idx, 'test')):
        d, m = data[idx].to(device), mask[idx].to(device)
        loader = DataLoader(ConcatDataset(d, m), batch_size=batch_size, shuffle=shuffle, drop_last=False,
                            num_workers=0)
        loader.pcvae_table = (d, m)
        out.append([loader, name])
    return out[0], out[1], data.shape[1]


def data_loader_mnar(data_path, vae_type, missing_rate, batch_size, data_type, device=torch.device('cpu'), shuffle=True,
                     data_transform='minmax'):
    """reference loaders.py:357-384: rows permuted by rand_perm<i>.pt, LAST column of data and mask dropped,
    min-max, one bare DataLoader; returns (loader, obs_dim).  The mask is float32."""
    index = [c for c in vae_type if c.isdigit()][0]
    folder = os.path.join(data_path, data_type)
    data = torch.load(os.path.join(folder, 'data.pt'))
    perm = torch.load(os.path.join('Data', data_type, f'rand_perm{index}.pt')).numpy()
    data = data[perm, :][:, :-1]
    mask = torch.load(os.path.join(folder, f'mnar_mask_missing{index}.pt'))[:, :-1]
    data = _minmax(data, data_transform)
    d, m = data.to(device), mask.to(device)
    loader = DataLoader(ConcatDataset(d, m), batch_size=batch_size, shuffle=shuffle, drop_last=False, num_workers=0)
    loader.pcvae_table = (d, m)
    return loader, data.shape[1]
