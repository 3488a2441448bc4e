"""The unmodified reference drivers import and call this package's mirrors with arguments that bind to
the mirror signatures.  Needs /root/reference (build container only; skipped on the GPU box)."""
import inspect
import json
import os
import sys

import pytest

REF = os.environ.get("PCVAE_REFERENCE", "/root/reference")
DRIVERS = os.path.join(REF, "src", "experiment_main")

pytestmark = pytest.mark.skipif(not os.path.isdir(DRIVERS), reason="reference checkout not present")


@pytest.fixture()
def injected(tmp_path, monkeypatch):
    from vae_posterior_consistency_b200 import evaluate, inject, loaders, train
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from synth import make_tree
    make_tree(str(tmp_path), "synth", 60, 5, seed=0)
    # one experiment line, reference JSON-lines format (Data/imputation_args.json)
    ref_line = json.loads(open(os.path.join(REF, "Data", "imputation_args.json")).readline())
    ref_line["vae_type"]["default"] = "reg_vae1"
    ref_line["data_type"]["default"] = "synth"
    ref_line["missing_rate"]["default"] = 30
    ref_line["epoch"]["default"] = 1
    ref_line["M"]["default"] = 1
    with open(tmp_path / "Data" / "imputation_args.json", "w") as f:
        f.write(json.dumps(ref_line) + "\n")
    calls = []

    def recorder(name, real):
        sig = inspect.signature(real)

        def rec(*a, **k):
            calls.append((name, sig.bind(*a, **k)))       # raises TypeError if the driver's call does not bind
        return rec

    monkeypatch.setattr(train, "train", recorder("train", train.train))
    monkeypatch.setattr(evaluate, "eval_vae", recorder("eval_vae", evaluate.eval_vae))
    monkeypatch.setattr(evaluate, "active_learning_func",
                        recorder("active_learning_func", evaluate.active_learning_func))
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(sys, "argv", ["driver"])
    yield inject, calls
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]


def test_imputation_driver_binds(injected):
    inject, calls = injected
    inject.run_driver(os.path.join(DRIVERS, "imputation.py"))
    names = [n for n, _ in calls]
    assert names == ["train", "eval_vae"]
    tr = calls[0][1].arguments
    assert tr["vae_type"] == "reg_vae1" and tr["obs_dim"] == 5 and tr["reg_type"] == "kl_reg"
    assert tr["num_estimates"] == 10 and tr["p_missingness"] == 30 and tr["alpha"] == 1.0
    ev = calls[1][1].arguments
    assert ev["max_epochs"] == 1 and ev["valid_k"] == 5000 and len(ev["list_loaders"]) == 2


def test_active_learning_driver_binds(injected):
    inject, calls = injected
    inject.run_driver(os.path.join(DRIVERS, "active_learning.py"))
    assert [n for n, _ in calls] == ["active_learning_func"]
    a = calls[0][1].arguments
    assert a["Repeat"] == 1 and a["obs_dim"] == 5 and a["test_data"].shape[1] == 5
    assert a["test_mask"].dtype.is_floating_point is False


def test_mnar_driver_binds(injected, tmp_path):
    inject, calls = injected
    from synth import make_tree_mnar
    from vae_posterior_consistency_b200 import evaluate
    import inspect as _inspect
    make_tree_mnar(str(tmp_path), "synthmnar", 30, 6, seed=1)
    line = json.loads(open(os.path.join(REF, "Data", "imputation_args_mnar.json")).readline())
    line["vae_type"]["default"] = "reg_notMIWAE1"
    line["data_type"]["default"] = "synthmnar"
    with open(tmp_path / "Data" / "imputation_args_mnar.json", "w") as f:
        f.write(json.dumps(line) + "\n")
    sig = _inspect.signature(evaluate.eval_vae_mnar)
    rec = []
    evaluate.eval_vae_mnar, orig = (lambda *a, **k: rec.append(sig.bind(*a, **k))), evaluate.eval_vae_mnar
    try:
        inject.run_driver(os.path.join(DRIVERS, "imputation_mnar.py"))
    finally:
        evaluate.eval_vae_mnar = orig
    assert [n for n, _ in calls] == ["train"] and len(rec) == 1
    tr = calls[0][1].arguments
    assert tr["vae_type"] == "reg_notMIWAE1" and tr["obs_dim"] == 6 and tr["p_missingness"] == 50
    assert tr["not_miwae_type"] == "changed" and tr["train_k"] == line["train_k"]["default"]
    ev = rec[0].arguments
    assert ev["data_test"].shape == (30, 6) and ev["mask_test"].dtype.is_floating_point
    assert ev["valid_k"] == line["valid_k"]["default"]


def test_mirror_signatures_equal_reference_signatures():
    sys.path.insert(0, REF)
    import types
    mpl = types.ModuleType("matplotlib"); mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules.setdefault("matplotlib", mpl); sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    import src.experiment_main.evaluate as RE
    import src.experiment_main.train as RT
    import src.models.VAE as RV
    import src.utils.loaders as RL
    from vae_posterior_consistency_b200 import VAE, evaluate, loaders, train
    names = lambda f: list(inspect.signature(f).parameters)
    assert names(train.train) == names(RT.train)
    assert names(evaluate.eval_vae) == names(RE.eval_vae)
    assert names(evaluate.active_learning_func) == names(RE.active_learning_func)
    assert names(evaluate.R_lindley_chain) == names(RE.R_lindley_chain)
    assert names(loaders.model_loader) == names(RL.model_loader)
    assert names(loaders.data_loader) == names(RL.data_loader)
    assert names(loaders.data_loader_mnar) == names(RL.data_loader_mnar)
    assert names(evaluate.eval_vae_mnar) == names(RE.eval_vae_mnar)
    for cls in ("REG_notMIWAE_v2", "notMIWAE_myversion"):
        ours, ref = getattr(VAE, cls), getattr(RV, cls)
        assert names(ours.__init__) == names(ref.__init__), cls
        for meth in ("encoder", "decoder", "forward", "loss"):
            assert names(getattr(ours, meth)) == names(getattr(ref, meth)), (cls, meth)
        dd = lambda f: {k: v.default for k, v in inspect.signature(f).parameters.items()}
        assert dd(ours.loss) == dd(ref.loss), cls
    for cls in ("Reg_VAE", "vanilla_VAE", "Reg_EDDI", "vanilla_EDDI"):
        ours, ref = getattr(VAE, cls), getattr(RV, cls)
        assert names(ours.__init__) == names(ref.__init__), cls
        for meth in ("encoder", "decoder", "forward", "loss"):
            assert names(getattr(ours, meth)) == names(getattr(ref, meth)), (cls, meth)
        # same defaults on loss (alpha etc.)
        d = lambda f: {k: v.default for k, v in inspect.signature(f).parameters.items()}
        assert d(ours.loss) == d(ref.loss), cls
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
