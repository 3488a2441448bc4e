"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/pcvae_b200.h declares.
No compute calls here (no GPU in the build container)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pcvae_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcvae_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from vae_posterior_consistency_b200 import build, lib
    path = build.build()
    dll = ctypes.CDLL(path)
    declared = header_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(dll, s)]
    assert not missing, missing
    assert sorted(lib.SYMBOLS) == declared           # the Python binding tracks the header
    assert dll.pcvae_abi_version() == 1


def test_only_sm100a_code_is_embedded():
    from vae_posterior_consistency_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_layout_helpers_and_argument_validation_without_gpu():
    from vae_posterior_consistency_b200 import lib as L
    m = L.model(L.FAMILY_MLP, 13)
    assert L.param_count(m) == 14433                 # MLP D=13
    assert L.param_count(L.model(L.FAMILY_MLP, 100)) == 31920
    assert L.param_count(L.model(L.FAMILY_PNP, 13, 20)) == 15866
    assert L.param_count(L.model(L.FAMILY_PNP, 100, 20)) == 26480
    assert L.param_count(L.model(L.FAMILY_MLP_MASK, 13)) == 14433 + 100 * 13     # first layer [100, 2D], VAE.py:526
    assert L.param_offsets(L.model(L.FAMILY_MLP_MASK, 100))[:3] == [0, 20000, 20100]
    offs = L.param_offsets(L.model(L.FAMILY_PNP, 100, 20))
    assert len(offs) == 17 and offs[0] == 0 and offs[-1] == 26480
    assert L.decoder_offset(m) == 7470
    with pytest.raises(L.PcvaeError):
        L.param_count(L.model(L.FAMILY_MLP, 4000))    # obs_dim outside the supported range
    with pytest.raises(L.PcvaeError):
        L.param_count(L.Model(0, 13, 0, 11))          # latent_dim must be 10
    # without a CUDA device every compute entry point refuses loudly (no CPU fallback)
    lib = L.load()
    import torch
    if not torch.cuda.is_available():
        assert lib.pcvae_grid_ctas() == -1
        assert b"no CPU fallback" in lib.pcvae_last_error()
        p = L.EncFwdParams(model=m, rows=4, n_branch=1)
        assert lib.pcvae_enc_fwd(ctypes.byref(p), None) == 2      # PCVAE_EDEVICE


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "vae_posterior_consistency_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            txt = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", txt, flags=re.S).replace("the oracle", ""), fn


def test_sass_every_mma_descriptor_register_is_written_before_its_first_use():
    """Guard against a ptxas 12.9 miscompilation seen in k_enc_fwd_tc: the high word of a tcgen05 shared-memory
    descriptor pair (stride-byte-offset + version bits) was materialised by a UMOV placed BEHIND the first UTCHMMA that
    reads it, inside the work-item loop, so the first item of every CTA ran with SBO = 0.  For every UTCHMMA in the
    library, both uniform registers of its gdesc[URn] pair must have a write earlier in the function's linear order."""
    from vae_posterior_consistency_b200 import build
    out = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if name and m:
            funcs[name].append(m.group(1).strip())
    n_mma, bad = 0, []
    for fn, ins in funcs.items():
        first_write = {}
        for i, s_ in enumerate(ins):
            m = re.match(r"(?:@!?U?P\d+\s+)?(\S+)\s+(UR\d+)", s_)
            if not m or m.group(1).startswith(("UTCHMMA", "UTCBAR", "UBLKCP")):
                continue
            dst = int(m.group(2)[2:])
            for r in ([dst, dst + 1] if (".64" in m.group(1) or ".WIDE" in m.group(1)) else [dst]):
                first_write.setdefault(r, i)
        for i, s_ in enumerate(ins):
            if "UTCHMMA" in s_:
                n_mma += 1
                for m in re.finditer(r"gdesc\[UR(\d+)\]", s_):
                    for r in (int(m.group(1)), int(m.group(1)) + 1):
                        if first_write.get(r, 1 << 30) > i:
                            bad.append((fn, i, f"UR{r}", first_write.get(r)))
    assert n_mma > 100, "no tcgen05 MMAs found in the library: the check did not look at anything"
    assert not bad, bad[:5]


def test_host_side_packing_of_the_streamed_batch_format():
    """kernels.pack_mask_bits / compact_rows (the host side of pcvae_prep_packed): bit j of word w is mask[row][32 w + j];
    vals are the observed entries in row-major order, row_off their exclusive prefix sums."""
    import numpy as np
    import torch
    from vae_posterior_consistency_b200 import kernels as KR
    g = torch.Generator().manual_seed(0)
    for rows, D in ((1, 4), (7, 100), (33, 128), (5, 36)):
        x = torch.rand(rows, D, generator=g)
        m = torch.rand(rows, D, generator=g) < 0.6
        if rows > 2:
            m[0] = False
            m[1] = True
        bits = KR.pack_mask_bits(m)
        assert bits.dtype == torch.int32 and bits.shape == (rows, (D + 31) // 32)
        u = bits.numpy().view(np.uint32)
        back = np.array([[(u[r, d // 32] >> (d % 32)) & 1 for d in range(D)] for r in range(rows)], dtype=bool)
        assert (back == m.numpy()).all()
        assert all(int(u[r, -1]) >> (D - 32 * (u.shape[1] - 1)) == 0 for r in range(rows)) or D % 32 == 0   # padding bits are zero
        vals, row_off, bits2 = KR.compact_rows(x, m)
        assert torch.equal(bits2, bits) and vals.numel() == int(m.sum())
        cnt = m.sum(1)
        assert torch.equal(row_off.long(), torch.cumsum(cnt, 0) - cnt)
        for r in range(rows):
            assert torch.equal(vals[row_off[r]:row_off[r] + cnt[r]], x[r][m[r]])


def test_ctypes_structures_have_the_size_and_field_offsets_of_the_header(tmp_path):
    """The ctypes mirrors in lib.py against include/pcvae_b200.h compiled as plain C (gcc): sizeof and the offset of every
    field, struct by struct -- a field added on one side only (as pcvae_dp_params.step_state was, mid-round) shows here
    on the CPU instead of as a garbage pointer on the GPU."""
    import ctypes as C
    from vae_posterior_consistency_b200 import lib as L
    pairs = {"pcvae_model": L.Model, "pcvae_enc_fwd_params": L.EncFwdParams, "pcvae_enc_bwd_params": L.EncBwdParams,
             "pcvae_dec_params": L.DecParams, "pcvae_loss_params": L.LossParams, "pcvae_reward_params": L.RewardParams,
             "pcvae_dense_fwd_params": L.DenseFwdParams, "pcvae_dense_bwd_params": L.DenseBwdParams,
             "pcvae_mnar_loss_params": L.MnarLossParams, "pcvae_dp_params": L.DpParams,
             "pcvae_miwae_loss_params": L.MiwaeLossParams, "pcvae_mnar_impute_params": L.MnarImputeParams}
    hdr = open(os.path.join(ROOT, "include", "pcvae_b200.h")).read()
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "pcvae_b200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        assert re.search(r"\}\s*" + cname + r"\s*;", hdr), f"{cname} is not a struct of the header"
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    for line in out.strip().splitlines():
        cname, size, *offs = line.split()
        cls = pairs[cname]
        assert int(size) == C.sizeof(cls), f"{cname}: sizeof {size} (header) vs {C.sizeof(cls)} (ctypes)"
        got = [getattr(cls, f).offset for f, _ in cls._fields_]
        assert [int(o) for o in offs] == got, f"{cname}: field offsets differ: {offs} vs {got}"
    # and every struct typedef of the header has a ctypes mirror
    assert set(re.findall(r"\}\s*(pcvae_\w+)\s*;", hdr)) == set(pairs)


def test_ctypes_argtypes_have_the_arity_of_the_header_prototypes():
    """Every prototype of include/pcvae_b200.h against the argtypes lib.py sets: same number of parameters."""
    from vae_posterior_consistency_b200 import lib as L
    lib = L.load()
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "pcvae_b200.h")).read(), flags=re.S)
    protos = re.findall(r"\b(?:int|size_t|long|const char\*)\s+(pcvae_\w+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S)
    assert len(protos) >= 35, len(protos)
    seen = set()
    for name, args in protos:
        seen.add(name)
        args = args.strip()
        n = 0 if args in ("", "void") else args.count(",") + 1
        fn = getattr(lib, name)
        if fn.argtypes is None:
            assert n == 0, f"{name}: {n} parameters in the header, no argtypes in lib.py"
        else:
            assert len(fn.argtypes) == n, f"{name}: {n} parameters in the header, {len(fn.argtypes)} argtypes"
    assert seen == set(L.SYMBOLS), (seen ^ set(L.SYMBOLS))


def test_digamma_of_the_student_t_score_against_scipy(tmp_path):
    """csrc/pcvae_special.cuh (host build, g++) against scipy.special.digamma on the range the MIWAE decoder produces
    (df = softplus + 3 >= 3, so arguments >= 1.5), in double and in float."""
    import numpy as np
    from scipy.special import digamma
    src = tmp_path / "dg.cpp"
    src.write_text('#include <cstdio>\n#include "pcvae_special.cuh"\n'
                   'int main() { for (double x = 1.5; x < 400.0; x *= 1.07) '
                   'printf("%.17g %.17g %.9g %.9g\\n", x, pcvae::digamma_pos<double>(x), (double)pcvae::digamma_pos<float>((float)x), '
                   '(double)pcvae::digamma_half_step<float>((float)(2.0 * x))); return 0; }\n')
    exe = tmp_path / "dg"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "vae_posterior_consistency_b200", "csrc"), str(src), "-o", str(exe)])
    rows = np.array([[float(v) for v in line.split()] for line in
                     subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()])
    x = rows[:, 0]
    assert len(x) > 50
    ref = digamma(x)
    assert np.max(np.abs(rows[:, 1] - ref)) < 5e-10                                  # double: series truncation only
    assert np.max(np.abs(rows[:, 2] - digamma(x.astype(np.float32).astype(np.float64)))) < 2e-6   # float
    step = digamma(np.float32(2 * x).astype(np.float64) / 2 + 0.5) - digamma(np.float32(2 * x).astype(np.float64) / 2)
    assert np.max(np.abs(rows[:, 3] - step) / step) < 2e-3 and np.max(np.abs(rows[:, 3] - step)) < 2e-6
