"""Long fuzz campaign of the emulated engine against the oracle (tests/test_emulated_fuzz.py holds the generator and a
short fixed-seed run).  Usage: python scripts/fuzz_emulated.py <first seed> <count>   (CPU only, ~0.4 s per case)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "simt_emu")]
import build_emu                      # noqa: E402
os.environ["B200NB_TEST_EMULATOR"] = "1"
os.environ.setdefault("B200NB_LIB", build_emu.build())      # an experiment build can be named from outside
from deseq2_b200 import wrappers      # noqa: E402
from oracle import oracle as O        # noqa: E402
import test_emulated_fuzz as F        # noqa: E402
from test_parity_gpu import ROBUST    # noqa: E402

O.build()
first, count = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for s in range(first, first + count):
    P = F.make_problem(s)
    if P is None:
        continue
    try:
        F.check_problem(wrappers, O, P, ROBUST)
    except AssertionError as e:
        bad += 1
        print("FAIL", str(e)[:600], flush=True)
print(f"{count} seeds, {bad} failures")
