"""Build libpcvae_b200.so in-tree with nvcc for sm_100a (no torch headers, plain C ABI)."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libpcvae_b200.so")
SOURCES = ["pcvae_common.cu", "pcvae_train.cu", "pcvae_reward.cu", "pcvae_data.cu", "pcvae_dense.cu",
           "pcvae_mnar.cu", "pcvae_miwae.cu", "pcvae_impute.cu", "pcvae_reward_tc.cu", "pcvae_reward_ws.cu", "pcvae_dec_tc.cu", "pcvae_enc_tc.cu", "pcvae_pnp_tc.cu", "pcvae_tc_images.cu", "pcvae_wgrad_tc.cu", "pcvae_dp.cu"]
HEADERS = ["pcvae_internal.cuh", "pcvae_tile.cuh", "pcvae_reward.cuh", "pcvae_tc.cuh", "pcvae_tc_tile.cuh", "pcvae_train.cuh", "pcvae_special.cuh", "pcvae_philox.cuh", os.path.join(ROOT, "include", "pcvae_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(CSRC, s.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {s}")
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "--cudart", "shared", "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
