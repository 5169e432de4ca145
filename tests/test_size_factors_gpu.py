"""Median-of-ratios size factors on the device (csrc/size_factors.cu) against the numpy
restatement of estimateSizeFactorsForMatrix (R/core.R:535-578), and the device-resident DESeq() started from raw
counts.  The same functions run on the emulated engine in tests/test_emulated_kernels.py (DEV = "cpu")."""
import numpy as np
import pytest

from helpers import rel_err

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _poscounts_reference(counts):
    """R/core.R:544-549 + :560-563: zeros contribute 0 to the row mean of logs; all-zero rows are left out."""
    counts = np.asarray(counts, dtype=np.float64)
    with np.errstate(divide="ignore"):
        lc = np.log(counts)
    lc0 = np.where(np.isfinite(lc), lc, 0.0)
    lg = lc0.mean(axis=1)
    lg[counts.sum(axis=1) == 0] = -np.inf
    sf = np.empty(counts.shape[1])
    for j in range(counts.shape[1]):
        sel = np.isfinite(lg) & (counts[:, j] > 0)
        sf[j] = np.exp(np.median(lc[sel, j] - lg[sel])) if sel.any() else np.nan
    return sf, lg


@pytest.mark.parametrize("n,m,seed,as_double", [(4000, 12, 1, False), (2501, 37, 2, True), (600, 130, 3, False),
                                                (64, 5, 4, False)])
def test_size_factors_match_numpy(engine, n, m, seed, as_double):
    import torch
    from deseq2_b200 import device as D, device_pipeline as DP, pipeline, synth
    d = synth.make_example_counts(n, m, seed=seed, interceptMean=6.0)
    counts = d["counts"].copy()
    counts[::11] = 0                                        # all-zero genes
    counts[1::13, 0] = 0                                    # genes with a single zero: out for "ratio", in for "poscounts"
    counts[5:40] = counts[5]                                # ties: many identical ratios around the median
    y = D.to_gene_major(counts.astype(np.float64) if as_double else counts, torch.device(DEV))
    got = DP.size_factors(y, m)
    ref = pipeline.estimateSizeFactorsForMatrix(counts)
    assert np.max(rel_err(got["sizeFactors"].cpu().numpy(), ref)) < 1e-12
    with np.errstate(divide="ignore"):
        lg = np.log(counts.astype(np.float64)).mean(axis=1)
    glg = got["loggeomeans"].cpu().numpy()
    assert np.array_equal(np.isfinite(glg), np.isfinite(lg))
    assert np.max(rel_err(glg[np.isfinite(lg)], lg[np.isfinite(lg)])) < 1e-13
    got2 = DP.size_factors(y, m, type="poscounts")
    ref2, lg2 = _poscounts_reference(counts)
    assert np.max(rel_err(got2["sizeFactors"].cpu().numpy(), ref2)) < 1e-12
    assert np.array_equal(np.isfinite(got2["loggeomeans"].cpu().numpy()), np.isfinite(lg2))


def test_size_factors_degenerate_inputs(engine):
    import torch
    from deseq2_b200 import device as D, device_pipeline as DP
    dev = torch.device(DEV)
    counts = np.array([[0, 3, 5], [2, 0, 7], [4, 1, 0]], dtype=np.int32)        # every gene has a zero
    with pytest.raises(ValueError, match="every gene contains at least one zero"):
        DP.size_factors(D.to_gene_major(counts, dev), 3)
    sf = DP.size_factors(D.to_gene_major(counts, dev), 3, type="poscounts")["sizeFactors"].cpu().numpy()
    lg = np.array([np.log([3, 5]).sum(), np.log([2, 7]).sum(), np.log([4, 1]).sum()]) / 3
    want = [np.exp(np.median([np.log(2) - lg[1], np.log(4) - lg[2]])), np.exp(np.median([np.log(3) - lg[0], np.log(1) - lg[2]])),
            np.exp(np.median([np.log(5) - lg[0], np.log(7) - lg[1]]))]
    assert np.allclose(sf, want, rtol=1e-13)
    one = np.array([[10, 20, 40, 80]], dtype=np.int32)                          # a single gene, even sample count
    sf1 = DP.size_factors(D.to_gene_major(one, dev), 4)["sizeFactors"].cpu().numpy()
    assert np.allclose(sf1, one[0] / np.exp(np.log(one[0]).mean()), rtol=1e-13)
    col = np.array([[5, 0], [7, 0], [9, 0]], dtype=np.int32)                    # a sample with no usable count
    sfc = DP.size_factors(D.to_gene_major(col, dev), 2, type="poscounts")["sizeFactors"].cpu().numpy()
    assert np.isfinite(sfc[0]) and np.isnan(sfc[1])


def test_deseq_device_from_raw_counts(engine, n=3000):
    """DESeq_device(sizeFactors=None) == DESeq_device(sizeFactors = the numpy median of ratios)."""
    import torch
    from deseq2_b200 import device as D, device_pipeline as DP, pipeline, synth
    m = 24
    x = synth.design_condition(m)
    d = synth.make_example_counts(n, m, x=x, seed=9)
    y = D.to_gene_major(d["counts"], torch.device(DEV))
    sf = pipeline.estimateSizeFactorsForMatrix(d["counts"])
    a = DP.DESeq_device(y, x)
    assert np.max(rel_err(a["sizeFactors"], sf)) < 1e-12
    b = DP.DESeq_device(y, x, sf)
    for k in ("dispGeneEst", "dispersion", "betaMatrix", "betaSE", "WaldPvalue", "maxCooks"):
        e = rel_err(a[k].cpu().numpy(), b[k].cpu().numpy(), floor=1e-9)
        assert np.mean(e < 1e-6) > 0.99, k
