"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py with the oracle).
CPU: the oracle still reproduces them (guards the checker against drift).  GPU: the engine matches them."""
import glob
import os

import numpy as np
import pytest

from helpers import DISP_KEYS, rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _load(path):
    z = np.load(path)
    ins = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    outs = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    for k in ("weightsSEXP",):
        ins.setdefault(k, None)
    for k, v in list(ins.items()):
        if isinstance(v, np.ndarray) and v.ndim == 0:
            ins[k] = v.item()
    return ins, outs


def _call(eng, name, ins):
    if name.startswith("fitdispgrid"):
        return eng.fitDispGrid(**ins)
    if name.startswith("fitdisp"):
        return eng.fitDisp(**ins)
    return eng.fitBeta(**ins)


def test_golden_files_present():
    assert len(FILES) >= 5


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_golden(oracle, path):
    name = os.path.basename(path)
    ins, outs = _load(path)
    got = _call(oracle, name, ins)
    for k, v in outs.items():
        if k == "margin":
            continue
        if v.dtype.kind in "iu":
            assert np.array_equal(got[k], v), (name, k)
        else:
            assert np.allclose(got[k], v, rtol=1e-10, atol=1e-12, equal_nan=True), (name, k)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_engine_matches_golden(engine, path):
    name = os.path.basename(path)
    ins, outs = _load(path)
    got = _call(engine, name, ins)
    if name.startswith("fitdispgrid"):
        assert np.mean(np.abs(got["log_alpha"] - outs["log_alpha"]) < 1e-9) > 0.96
        return
    if name.startswith("fitdisp"):
        ok = (outs["margin"] > 64) & (got["iter"] == outs["iter"]) & (got["iter_accept"] == outs["iter_accept"])
        assert np.all((got["iter"] == outs["iter"])[outs["margin"] > 64])
        assert ok.mean() > 0.8
        for k in ("log_alpha", "initial_lp", "last_lp"):
            assert np.max(rel_err(got[k][ok], outs[k][ok])) < 1e-6, (name, k)
        return
    assert np.array_equal(got["iter"], outs["iter"])
    for k in ("beta_mat", "beta_var_mat", "hat_diagonals", "contrast_num", "contrast_denom", "deviance"):
        assert np.nanmax(rel_err(got[k], outs[k], floor=1e-8)) < 1e-6, (name, k)
