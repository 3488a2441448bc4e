"""Run the reference's UNMODIFIED experiment drivers on top of this package.

The drivers (`src/experiment_main/imputation.py`, `active_learning.py`) import
`src.models.VAE`, `src.utils.loaders`, `src.utils.utils`, `src.utils.AIS`,
`src.experiment_main.train` and `src.experiment_main.evaluate` (imputation.py:5-11).
`install()` registers this package's mirrors in `sys.modules` under exactly those names
(plus a `matplotlib.pyplot` stub for the unused import at evaluate.py:10), after which

    python -m vae_posterior_consistency_b200.inject /path/to/reference/src/experiment_main/imputation.py [driver args]

executes the driver file as `__main__` with the current directory as its working directory
(it must contain `Data/`).  See INTEGRATION.md.
"""
import runpy
import sys
import types


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def _ais_unavailable(*a, **k):
    raise NotImplementedError("AIS evaluation is dead code in the reference drivers (imported, never called) "
                              "and is outside the B200 hot path")


def install():
    from . import VAE, evaluate, loaders, train, utils
    pkgs = {
        "src": _module("src"), "src.models": _module("src.models"), "src.utils": _module("src.utils"),
        "src.experiment_main": _module("src.experiment_main"),
    }
    for m in pkgs.values():
        m.__path__ = []
    mods = {
        "src.models.VAE": VAE,
        "src.utils.loaders": loaders,
        "src.utils.utils": utils,
        "src.utils.AIS": _module("src.utils.AIS", linear_schedule=_ais_unavailable, eval_ais=_ais_unavailable,
                                 sigmoidial_schedule=_ais_unavailable),
        "src.utils.pytorchtools": _module("src.utils.pytorchtools", EarlyStopping=object),
        "src.experiment_main.train": train,
        "src.experiment_main.evaluate": evaluate,
    }
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = _module("matplotlib")
            mpl.pyplot = _module("matplotlib.pyplot")
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = mpl.pyplot
    sys.modules.update(pkgs)
    sys.modules.update(mods)
    for full, mod in mods.items():
        parent, leaf = full.rsplit(".", 1)
        setattr(sys.modules[parent], leaf, mod)
    # names the drivers import from src.utils.loaders that this package does not reimplement
    for missing in ("data_loader_mnist",):
        if not hasattr(loaders, missing):
            setattr(loaders, missing, _ais_unavailable)
    return mods


def run_driver(path, argv=()):
    """Execute the driver file as __main__.  The drivers switch autograd's anomaly mode on at import time
    (imputation.py:19, active_learning.py:21) and never off; the caller's setting is restored afterwards (anomaly mode
    synchronises after every backward node, which among other things forbids CUDA-graph capture)."""
    import torch
    install()
    old = sys.argv
    anomaly = torch.is_anomaly_enabled()
    sys.argv = [path, *argv]
    try:
        return runpy.run_path(path, run_name="__main__")
    finally:
        sys.argv = old
        torch.autograd.set_detect_anomaly(anomaly)


if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    run_driver(sys.argv[1], sys.argv[2:])
