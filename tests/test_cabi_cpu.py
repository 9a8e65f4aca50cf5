"""CPU-only checks of the C-ABI boundary: the library loads, exports every symbol declared in
include/espm_b200.h, its struct layout matches the ctypes mirror, and the product path refuses to run
without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from espm_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    from espm_b200 import _lib
    header = open(os.path.join(ROOT, "include", "espm_b200.h")).read()
    declared = set(re.findall(r"\b(espm_[a-z_0-9]+)\s*\(", header))
    declared -= {"espm_status", "espm_state"}
    assert declared, "no declarations parsed"
    assert declared == set(_lib.exported_symbols())
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layout_matches_ctypes(lib):
    from espm_b200 import _lib
    out = (ctypes.c_int64 * 8)()
    assert lib.espm_state_layout(out) == 0
    S = _lib.EspmState
    expect = [ctypes.sizeof(S)] + [getattr(S, f).offset for f in
                                   ("p_total", "lambda_L", "mu", "Xt", "H_prev", "numraw", "scalars")]
    assert list(out) == expect


def test_constants_match_header():
    from espm_b200 import _lib
    header = open(os.path.join(ROOT, "include", "espm_b200.h")).read()
    for name, val in (("ESPM_TILE_PX", _lib.TILE_PX), ("ESPM_MAX_K", _lib.MAX_K), ("ESPM_NSCALARS", _lib.NSCALARS),
                      ("ESPM_MAXIT_DICHOTOMY", _lib.MAXIT_DICHOTOMY)):
        m = re.search(r"#define\s+%s\s+(\d+)" % name, header)
        assert m and int(m.group(1)) == val, name
    for name in ("SIMPLEX_H", "SIMPLEX_W", "G_IDENTITY", "CLAMP_Y", "LOSS_DUAL", "FIXED_H", "FIXED_W", "MU",
                 "LAPLACIAN", "HAVE_HPREV", "SIMPLEX_ROWS", "HQ", "FUSED_WREDUCE", "PEER", "BMD", "PG", "L2", "L2_H",
                 "LINESEARCH", "EVAL_ONLY", "LS_PARTIAL", "NO_HSPEC", "TIMING"):
        m = re.search(r"#define\s+ESPM_FLAG_%s\s+\(1u << (\d+)\)" % name, header)
        assert m and (1 << int(m.group(1))) == getattr(_lib, "FLAG_" + name), name
    # no two flags share a bit
    bits = [int(b) for b in re.findall(r"#define\s+ESPM_FLAG_\w+\s+\(1u << (\d+)\)", header)]
    assert len(bits) == len(set(bits))
    # layout of the peer flag block and the record slots added in round 2
    for name, val in (("ESPM_PF_SFLAG", _lib.PF_SFLAG), ("ESPM_PF_MFLAG", _lib.PF_MFLAG), ("ESPM_PF_MASK", _lib.PF_MASK),
                      ("ESPM_PF_TFLAG", _lib.PF_TFLAG), ("ESPM_PF_WORDS", _lib.PF_WORDS),
                      ("ESPM_COOP_BLOCKS", _lib.COOP_BLOCKS)):
        m = re.search(r"#define\s+%s\s+(\d+)" % name, header)
        assert m and int(m.group(1)) == val, name
    max_ranks = int(re.search(r"#define\s+ESPM_MAX_RANKS\s+(\d+)", header).group(1))
    assert _lib.PF_TFLAG + 32 * max_ranks <= _lib.PF_WORDS and _lib.COOP_BLOCKS <= 32      # one flag word per (rank, CTA)
    assert _lib.PF_MASK + 4 * max_ranks <= _lib.PF_TFLAG
    for name, val in (("ESPM_S_T0", _lib.S_T0), ("ESPM_S_STAMP", _lib.S_STAMP)):
        m = re.search(r"%s\s*=\s*(\d+)" % name, header)
        assert m and int(m.group(1)) == val, name
    assert _lib.S_LS_D < _lib.S_T0 and _lib.S_T0 + 8 < _lib.S_STAMP < _lib.NSCALARS


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert lib.espm_device_count() < 0
    from espm_b200 import SmoothNMF, _lib
    X = np.random.default_rng(0).poisson(3.0, size=(20, 12)).astype(float)
    est = SmoothNMF(n_components=2, max_iter=2, verbose=0)
    with pytest.raises(_lib.EspmError):
        est.fit_transform(X)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under espm_b200/ may import or execute it."""
    pkg = os.path.join(ROOT, "espm_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[./]smooth_nmf_oracle|_ref\b", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), (dirpath, f)


def test_header_is_plain_c_and_links(tmp_path):
    """include/espm_b200.h must be consumable from C (the drop-in boundary is a C ABI): compile a C99 translation unit
    with -pedantic, link it against the shared library and call two entry points that need no GPU."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    from espm_b200 import _lib
    src = tmp_path / "abi.c"
    src.write_text(
        '#include <stdio.h>\n#include "espm_b200.h"\n'
        "int main(void) {\n"
        "    long long out[8];\n"
        "    espm_state st;\n"
        "    if (espm_state_layout((int64_t*)out) != 0) return 2;\n"
        "    if (out[0] != (long long)sizeof(st)) return 3;\n"
        '    printf("%d %lld\\n", espm_version(), out[0]);\n'
        "    return 0;\n}\n")
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIBPATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
           "-o", str(exe), "-L", libdir, "-lespm_b200", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    version, size = r.stdout.split()
    assert int(version) > 0 and int(size) == ctypes.sizeof(_lib.EspmState)
