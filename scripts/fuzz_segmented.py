"""Random grouped designs (5..32 groups, rows of G + 2 .. 700 samples) through
tests/test_parity_gpu.py::test_segmented_general_p_kernels on the emulated engine: segmented kernels vs the oracle and
vs the kernels of fit_generic.cu.
usage: B200NB_LIB=<emulated library> python scripts/fuzz_segmented.py <rng seed> <cases>"""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "simt_emu")]
os.environ["B200NB_TEST_EMULATOR"] = "1"
import numpy as np
import ctypes as C
from deseq2_b200 import _lib, wrappers
from oracle import oracle as O
O.build()
import test_parity_gpu as TP
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
ncase = int(sys.argv[2]) if len(sys.argv) > 2 else 40
bad = 0
for c in range(ncase):
    G = int(rng.integers(5, 33))
    m = int(rng.integers(G + 2, 700)) if rng.random() < 0.7 else int(rng.integers(G + 2, 80))
    seed = int(rng.integers(1, 10000))
    try:
        TP.test_segmented_general_p_kernels(wrappers, O, m, G, seed)
        print(f"ok   m={m} G={G} seed={seed}", flush=True)
    except AssertionError as e:
        bad += 1
        print(f"FAIL m={m} G={G} seed={seed}: {str(e)[:300]}", flush=True)
    except Exception as e:
        bad += 1
        print(f"ERR  m={m} G={G} seed={seed}: {type(e).__name__} {str(e)[:300]}", flush=True)
print("cases", ncase, "failures", bad)
