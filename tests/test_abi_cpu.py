"""CPU-side checks of the drop-in boundary: the library builds, loads, exports every
symbol include/c3r_b200.h declares, and refuses to run without a GPU (no fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from clair3_rna_b200 import build, lib as L
    build.build()
    return L.load()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "c3r_b200.h")).read()
    return sorted(set(re.findall(r"\b(c3r_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from clair3_rna_b200 import lib as L
    syms = header_symbols()
    assert set(syms) == set(L.EXPORTS)
    for s in syms:
        assert hasattr(lib, s), s


def test_abi_version_and_defaults(lib):
    from clair3_rna_b200 import lib as L
    assert lib.c3r_abi_version() == 2
    p = L.Params()
    lib.c3r_default_params(C.byref(p))
    assert (p.channels, p.min_coverage, p.min_mq, p.excl_flags, p.max_depth) == (18, 4, 5, 2316, 144)
    assert (p.snp_min_af, p.indel_min_af, p.skip_proportion) == (0.08, 0.15, 0.2)


def test_struct_sizes_match_header(lib):
    from clair3_rna_b200 import lib as L
    assert C.sizeof(L.AltEntry) == 16
    assert C.sizeof(L.Reads) == 3 * 8 + 8 * 8
    assert C.sizeof(L.Params) == 64


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from clair3_rna_b200.engine import Engine, C3RError
    with pytest.raises(C3RError) as e:
        Engine(0, 18)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_bad_channels_rejected(lib):
    from clair3_rna_b200 import lib as L
    p = L.Params()
    lib.c3r_default_params(C.byref(p))
    p.channels = 17
    ctx = C.c_void_p()
    assert lib.c3r_create(C.byref(ctx), 0, C.byref(p)) == -1
