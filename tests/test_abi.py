"""The C-ABI library loads without a GPU and exports exactly the symbols include/b200nb.h declares; the product
has no CPU fallback (compute calls fail loudly without a device) and never imports the oracle."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "b200nb.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200nb_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported_and_bound():
    import deseq2_b200
    from deseq2_b200 import _lib
    lib = deseq2_b200.lib()
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200nb.h but not exported by libb200nb.so"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in deseq2_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == syms
    out = subprocess.run(["nm", "-D", "--defined-only", deseq2_b200.lib_path()], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\b(b200nb_[a-z0-9_]+)\b", out)))
    assert exported == syms, "exported b200nb_* symbols differ from the header"
    assert lib.b200nb_version().startswith(b"b200nb")
    assert lib.b200nb_kernel_launches() >= 0


def test_no_cpu_fallback():
    import deseq2_b200
    from deseq2_b200 import wrappers
    if deseq2_b200.lib().b200nb_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    y = np.ones((4, 6), dtype=np.int32)
    x = np.c_[np.ones(6), np.r_[0, 0, 0, 1, 1, 1.0]]
    with pytest.raises(deseq2_b200.EngineError):
        wrappers.fitDisp(y, x, np.ones((4, 6)), np.zeros(4), np.zeros(4), 1.0, -20.0, 1.0, 1e-6, 10, False, None, False,
                         1e-2, True)
    with pytest.raises(deseq2_b200.EngineError):
        wrappers.fitBeta(y, x, np.ones((4, 6)), np.ones(4), [1, 0], np.zeros((4, 2)), [1e-6, 1e-6], None, False, 1e-8,
                         10, True, 0.5)


def test_argument_errors_are_reported():
    import deseq2_b200
    lib = deseq2_b200.lib()
    rc = lib.b200nb_fit_disp(None, 0, None, None, None, None, 1.0, -20.0, 1.0, 1e-6, 10, 0, None, 0, 1e-2, 1, 5, 0, 2,
                             *([None] * 9))
    assert rc != 0 and b"bad dimensions" in lib.b200nb_last_error()
    rc = lib.b200nb_fit_beta(None, 0, None, None, None, None, None, None, None, 0, 1e-8, 10, 1, 0.5, 5, 6, 64,
                             *([None] * 8))
    assert rc != 0 and b"not supported" in lib.b200nb_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "deseq2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "nbglm_oracle" not in txt, f


def test_na_screening_mirrors_r_wrappers():
    """R/wrappers.R:30-34,109-113: NA in any argument is an error raised before the native call."""
    from deseq2_b200 import wrappers
    y = np.ones((3, 4))
    x = np.c_[np.ones(4), [0, 0, 1, 1.0]]
    mu = np.ones((3, 4))
    mu[1, 2] = np.nan
    with pytest.raises(ValueError, match="mu_hatSEXP"):
        wrappers.fitDispWrapper(y, x, mu, np.zeros(3), np.zeros(3), 1.0, -20, 1.0, 1e-6, 10, False, None, False, 1e-2, True)
    with pytest.raises(ValueError, match="alpha_hatSEXP"):
        wrappers.fitBetaWrapper(y, x, np.ones((3, 4)), np.array([1.0, np.nan, 1.0]), np.zeros((3, 2)), [1e-6, 1e-6], None,
                                False, 1e-8, 10, True, 0.5)


def test_product_refuses_the_emulated_engine():
    """The SIMT-emulated build of the engine (tests/simt_emu) is a checker, not a way to run: without the test-suite's
    B200NB_TEST_EMULATOR=1 the loader refuses it, so no configuration of the product computes on the CPU."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "simt_emu"))
    import build_emu
    env = {k: v for k, v in os.environ.items() if k != "B200NB_TEST_EMULATOR"}
    env["B200NB_LIB"] = build_emu.build()
    r = subprocess.run([sys.executable, "-c", "import deseq2_b200; deseq2_b200.lib()"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "SIMT-emulated test build" in r.stderr


def test_host_content_hash_variants_agree_and_match_definition():
    """The cache key of the host entry points (hostrt.h::hash_elems): every SIMD variant this CPU can run equals the
    scalar one, the value is independent of how the buffer is split, and it matches a restatement of the definition
    in Python integers (lo32 * hi32 + swap32 of the keyed word, two key streams, summed mod 2^64)."""
    import ctypes as C
    import numpy as np
    import deseq2_b200
    L = deseq2_b200.lib()
    K1, K2, M = 0x9E3779B97F4A7C15, 0xD6E8FEB86659FD93, (1 << 64) - 1
    rng = np.random.Generator(np.random.PCG64(5))

    def py_hash(words, first):
        a = b = 0
        for i, v in enumerate(words):
            for which, K in ((0, K1), (1, K2)):
                x = int(v) ^ (((first + i + 1) * K) & M)
                t = ((x & 0xffffffff) * (x >> 32) + (((x << 32) | (x >> 32)) & M)) & M
                if which == 0:
                    a = (a + t) & M
                else:
                    b = (b + t) & M
        return a, b

    def lib_hash(arr, elem, first):
        out = (C.c_ulonglong * 2)()
        rc = L.b200nb_test_hash(C.c_void_p(arr.ctypes.data), arr.size, elem, first, out)
        assert rc >= 1, rc
        return (out[0], out[1]), rc

    for elem, dt in ((4, np.uint32), (8, np.uint64)):
        for count in (0, 1, 7, 8, 9, 31, 257):
            arr = rng.integers(0, np.iinfo(dt).max, size=count, dtype=dt, endpoint=True)
            h, variants = lib_hash(arr, elem, 1000)
            assert h == py_hash(arr, 1000), (elem, count)
        big = rng.integers(0, np.iinfo(dt).max, size=100003, dtype=dt, endpoint=True)
        whole, variants = lib_hash(big, elem, 0)
        left, _ = lib_hash(big[:40001], elem, 0)
        right, _ = lib_hash(np.ascontiguousarray(big[40001:]), elem, 40001)
        assert whole == ((left[0] + right[0]) & M, (left[1] + right[1]) & M)
        flipped = big.copy()
        flipped[77777] ^= dt(1)
        assert lib_hash(flipped, elem, 0)[0] != whole
    assert variants >= 1
