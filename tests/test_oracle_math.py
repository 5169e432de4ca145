"""The oracle's restated nmath (lgamma / digamma / trigamma / dnbinom_mu) against mpmath at 50 digits."""
import mpmath as mp
import numpy as np

mp.mp.dps = 50


def test_special_functions_vs_mpmath(oracle):
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(1e-3, 30, 300), 10 ** rng.uniform(-8, 9, 300), np.arange(1, 60) * 0.5])
    L = oracle.lib()
    for x in xs:
        x = float(x)
        for f, ref in ((L.oracle_lgamma, mp.loggamma), (L.oracle_digamma, mp.digamma),
                       (L.oracle_trigamma, lambda v: mp.polygamma(1, v))):
            r = float(ref(mp.mpf(x)))
            assert abs(f(x) - r) <= 4e-15 * max(1.0, abs(r)), (f, x)


def test_dnbinom_mu_vs_mpmath(oracle):
    rng = np.random.default_rng(1)

    def ref(x, size, mu):
        x, size, mu = mp.mpf(x), mp.mpf(size), mp.mpf(mu)
        return (mp.loggamma(x + size) - mp.loggamma(size) - mp.loggamma(x + 1) + size * mp.log(size / (size + mu))
                + x * mp.log(mu / (size + mu)))

    worst = 0.0
    for _ in range(1500):
        mu = 10 ** rng.uniform(-0.3, 5)
        alpha = 10 ** rng.uniform(-8, 1.3)
        size = 1 / alpha
        x = float(rng.negative_binomial(max(size, 1e-3), size / (size + mu))) if rng.random() < 0.8 else float(
            rng.integers(0, 5))
        a, b = oracle.dnbinom_mu_log(x, size, mu), float(ref(x, size, mu))
        worst = max(worst, abs(a - b) / max(1.0, abs(b)))
    # R's dbinom_raw carries log1p(-x/n) cancellation for huge size: ~1e-9 relative, inherited on purpose
    assert worst < 1e-8
