"""Install the UNMODIFIED reference (stschia/VAE-posterior-consistency) into baseline/_ref/ (git-ignored, shipped to the
GPU box with the snapshot) so that `bench.py --impl reference` and the driver-level GPU tests can run the reference's
own code on the box's host cores.  The reference is a bare source tree (no setup.py / pyproject.toml), so it is copied
to a scratch directory, given a three-line pyproject.toml there, and installed with the offline pip command of the
task contract; `Data/*.json` (the drivers' argument files) is copied beside it.  Nothing is written to /root/reference
and no reference source enters the repository history.

    python baseline/install_ref.py [--force]
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("PCVAE_REFERENCE", "/root/reference")

PYPROJECT = """[build-system]
requires = ["setuptools"]
build-backend = "setuptools.build_meta"
[project]
name = "vae-posterior-consistency-reference"
version = "0"
[tool.setuptools.packages.find]
include = ["src*"]
"""


def installed():
    return os.path.exists(os.path.join(DEST, "src", "models", "VAE.py"))


def install(force=False):
    if installed() and not force:
        return DEST
    if not os.path.isdir(os.path.join(SRC, "src")):
        raise FileNotFoundError(f"{SRC}/src not found: the reference is only present in the build container")
    tmp = tempfile.mkdtemp(prefix="pcvae_ref_")
    try:
        work = os.path.join(tmp, "reference")
        shutil.copytree(SRC, work)
        with open(os.path.join(work, "pyproject.toml"), "w") as f:
            f.write(PYPROJECT)
        if os.path.isdir(DEST):
            shutil.rmtree(DEST)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", DEST, work]
        out = subprocess.run(cmd, capture_output=True, text=True)
        if out.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + out.stdout[-2000:] + out.stderr[-2000:])
        shutil.copytree(os.path.join(SRC, "Data"), os.path.join(DEST, "Data"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    assert installed()
    return DEST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
