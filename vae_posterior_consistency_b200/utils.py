"""Mirror of the helpers of the reference's `src/utils/utils.py` that the hot path's callers use.

Host-side only (NumPy / argparse): `create_missing_uci` defines the bits of the per-batch
sub-mask (reference utils.py:36-39) and must consume NumPy's global RNG exactly like the
reference so that seeded runs produce bit-identical masks.
"""
import argparse

import numpy as np
import torch


def create_missing_uci(shape, missing_rate):
    """Bernoulli keep-mask with keep probability 1 - missing_rate/100 drawn from NumPy's global
    generator (reference utils.py:36-39): one `np.random.rand(*shape)` call, `<` comparison, bool tensor."""
    keep = 1 - missing_rate / 100
    return torch.from_numpy(np.random.rand(*shape) < keep)


def create_missing_uci_drop_eddi(shape):
    """EDDI drop mask (reference utils.py:42-45): per-entry keep probability 1 - min(u, 0.99)."""
    from scipy.stats import bernoulli
    u = np.minimum(np.random.rand(*shape), 0.99)
    return torch.from_numpy(bernoulli.rvs(1 - u))


def setup_parser(arguments, title):
    """JSON-line {name: {type, default, help}} -> argparse with single-dash flags (reference utils.py:177-189)."""
    parser = argparse.ArgumentParser(description=title, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    for key, spec in arguments.items():
        parser.add_argument('-%s' % key, type=type(spec["default"]), help=spec["help"], default=spec["default"])
    return parser


def completion(x, mask, mask_p, M, model):
    """M imputations conditioned on the observed entries (reference utils.py:192-208): [M, N, D]."""
    im = torch.zeros((M, x.shape[0], x.shape[1]), device=x.device)
    for m in range(M):
        im[m] = model.forward(x, mask, mask_p, 'evaluate')[6]
    return im
