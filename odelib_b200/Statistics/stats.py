"""Host-side statistics helpers with the reference's names and formulas (ODElib/Statistics/stats.py).

These are the small-array utilities a user calls on host data (``ModelFramework.get_chi(mod_dict)`` etc.).
The hot path never comes through here: sweeps and chains score trajectories inside the CUDA kernels
(odl_kernels.cuh: odl_score) with the same masking rules.
"""
import numpy as np


def chi(O, C, S):
    """sum((O - C)^2 / (2 S^2)) where every non-finite term is dropped (stats.py:22-41, np.ma semantics).

    Returns np.ma.masked when no term is valid, like the reference."""
    O = np.asarray(O, dtype=np.float64)
    C = np.asarray(C, dtype=np.float64)
    S = np.asarray(S, dtype=np.float64)
    with np.errstate(all="ignore"):
        d = O - C
        dd = d * d
        den = 2 * (S * S)
        term = dd / den
        ok = np.isfinite(O) & np.isfinite(dd) & np.isfinite(term) & ~(np.abs(dd) * np.finfo(float).tiny >= np.abs(den))
    if not ok.any():
        return np.ma.masked
    return float(np.sum(term[ok]))


def AIC(chi, num_parameters):
    """2*chi + 2*k (stats.py:44-47)."""
    return -2 * (-chi) + 2 * num_parameters


def Rsqrd(C_dict, O_dict):
    """Linear-space R^2 with NaN residuals skipped and population variance (stats.py:49-56)."""
    sstot = 0.0
    ssres = 0.0
    for sname in C_dict:
        c = np.asarray(C_dict[sname], dtype=np.float64)
        o = np.asarray(O_dict[sname], dtype=np.float64)
        ssres += np.nansum((c - o) ** 2)
        sstot += c.shape[0] * np.var(o)
    return 1 - ssres / sstot


def get_adjusted_rsquared(Rsqrd, num_samples, num_parameters):
    """1 - (1-R^2)(n-1)/(n-p-1) (stats.py:58-63)."""
    n, p = num_samples, num_parameters
    return 1 - (1 - Rsqrd) * (n - 1) / (n - p - 1)


def predict_logsigma(sigma, mean):
    """Log-space standard deviation from linear-space sigma and mean (stats.py:3-20)."""
    return np.log(1.0 + sigma ** 2.0 / mean ** 2.0) ** 0.5
