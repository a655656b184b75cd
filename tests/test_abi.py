"""CPU tier: the C-ABI library loads without a GPU, exports every symbol the header declares, and refuses to run without CUDA."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from gw_analysis_tools_b200 import abi, engine
from gw_analysis_tools_b200 import sampler as sampler_binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gwat_b200.h")
SAMPLER_HEADER = os.path.join(ROOT, "include", "gwat_b200_sampler.h")


def declared_functions(header=HEADER):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gwat_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_python_binds():
    assert sorted(engine.EXPORTS) == declared_functions()


def test_sampler_header_declares_what_python_binds():
    assert sorted(sampler_binding.EXPORTS) == declared_functions(SAMPLER_HEADER)


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    for name in declared_functions() + declared_functions(SAMPLER_HEADER):
        assert hasattr(lib, name), name
    assert lib.gwat_b200_abi_version() == abi.ABI_VERSION


def test_struct_layout_matches_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "gwat_b200_sampler.h"\nint main(void){printf("%zu %zu %zu %zu\\n", sizeof(gwat_b200_source), sizeof(gwat_b200_mod), sizeof(gwat_b200_prior), sizeof(gwat_b200_sampler_options));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    a, b, c, d = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(a) == C.sizeof(abi.Source) and int(b) == C.sizeof(abi.Mod)
    assert int(c) == C.sizeof(sampler_binding.Prior) and int(d) == C.sizeof(sampler_binding.Options)


def test_defaults_match_reference_members():
    lib = engine.load_library()
    s = abi.Source()
    lib.gwat_b200_source_init(C.byref(s))
    d = abi.source_defaults()
    for name, _ in abi.Source._fields_:
        a, b = getattr(s, name), getattr(d, name)
        if hasattr(a, "__len__"):
            assert list(a) == list(b), name
        else:
            assert a == b, name
    assert s.tidal1 == -1 and s.chip == -1 and s.shift_time == 1 and s.tidal_love == 1 and s.PNorder == 35


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to create a context (and says why); it never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(engine.GwatB200Error) as e:
        engine.Context(0)
    assert e.value.code == abi.ERR_CUDA
    assert "no CPU path" in str(e.value)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "gw_analysis_tools_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".h", ".cu", ".cpp", ".inc")):
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                path = os.path.join(dirpath, fn)
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
                assert not re.search(r'#include\s+"[^"]*(oracle|host_harness)', text), path
                assert "libgwat_ref" not in text and "libgwat_host_harness" not in text, path
                # the only library the product opens at run time is NCCL (the sharded sampler, gwat_sampler.cu: nccl_api)
                for line in text.splitlines():
                    if re.search(r"\bdlopen\s*\(", line):
                        assert re.search(r"dlopen\(n, RTLD_NOW", line) and "libnccl.so.2" in text, (path, line)


def test_gauss_legendre_grid_matches_reference(oracle):
    """gwat_b200_gauss_legendre_grid against the reference's gauleg (+ pow(10, .)), and against the rule's defining property."""
    for lo, hi, n, lg in ((10.0, 2048.0, 1000, True), (20.0, 1024.0, 257, True), (-1.0, 3.0, 16, False), (5.0, 6.0, 1, False)):
        f, w = engine.gauss_legendre_grid(lo, hi, n, lg)
        rf, rw = oracle.gauleg_grid(lo, hi, n, lg)
        assert np.allclose(f, rf, rtol=1e-14, atol=0) and np.allclose(w, rw, rtol=1e-12, atol=0)
        assert np.all(np.diff(f) > 0) if n > 1 else True
    x, w = engine.gauss_legendre_grid(0.0, 2.0, 12, False)
    # exact for polynomials up to degree 2n - 1 -- to the 1e-10 at which the algorithm (and the reference) stops Newton's
    # iteration: the weights are formed from the derivative at the last-but-one iterate
    for k in range(0, 24):
        assert abs((w * x ** k).sum() - 2.0 ** (k + 1) / (k + 1)) <= 1e-9 * 2.0 ** (k + 1)
    f, w = engine.gauss_legendre_grid(10.0, 1000.0, 64, True)
    assert abs((w * f * np.log(10.0)).sum() - 990.0) <= 1e-9 * 990.0  # int df = int f ln10 dlog10 f
    with pytest.raises(engine.GwatB200Error):
        engine.gauss_legendre_grid(10.0, 5.0, 8, True)
