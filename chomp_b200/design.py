"""Synthetic parameter batches: a Latin-hypercube design over cosmology / halo /
HOD parameters, as the reference's batch caller draws it
(simulation_design.py:17-33: ``P[:, i] = permutation(n); (P + U(0,1)) / n``),
with the flat-universe closure of SimulationDesignFlatUniverse
(simulation_design.py:238-239) and the Wake et al. assumption
log_M_0 = log_M_min (simulation_design.py:291).  Ranges: SURVEY.md section 8(d).
"""
import numpy as np

from . import _lib

SEED = 20261018

COSMO_RANGES = {
    "omega_m0": (0.25, 0.35), "omega_b0": (0.040, 0.052), "h": (0.65, 0.75),
    "sigma_8": (0.70, 0.90), "n_scalar": (0.92, 1.00),
}
HALO_RANGES = {"c0": (7.0, 11.0), "beta": (-0.20, -0.08)}
ZHENG_RANGES = {"log_M_min": (11.8, 12.6), "sigma": (0.10, 0.50),
                "log_M_1p": (13.0, 14.0), "alpha": (0.8, 1.3)}
MANDELBAUM_RANGES = {"log_M_0": (11.8, 12.6), "w": (0.5, 2.0)}


def latin_hypercube(n_points, n_dim, rng):
    """Unit-cube Latin hypercube, one stratum per point and dimension."""
    cells = np.empty((n_points, n_dim))
    for d in range(n_dim):
        cells[:, d] = rng.permutation(n_points)
    return (cells + rng.uniform(size=(n_points, n_dim)))/n_points


def synthetic_batch(n_points, hod="zheng", seed=SEED):
    """(cosmo [B,10], halo [B,6], hod [B,5]) float64 arrays in the column order
    of include/chomp_b200.h."""
    rng = np.random.default_rng(seed)
    hod_ranges = ZHENG_RANGES if hod == "zheng" else MANDELBAUM_RANGES
    names = list(COSMO_RANGES) + list(HALO_RANGES) + list(hod_ranges)
    ranges = dict(COSMO_RANGES)
    ranges.update(HALO_RANGES)
    ranges.update(hod_ranges)
    unit = latin_hypercube(n_points, len(names), rng)
    col = {nm: ranges[nm][0] + (ranges[nm][1] - ranges[nm][0])*unit[:, i]
           for i, nm in enumerate(names)}
    cosmo = np.zeros((n_points, _lib.N_COSMO))
    orad = 4.15e-5/col["h"]**2
    cosmo[:, 0] = col["omega_m0"]
    cosmo[:, 1] = col["omega_b0"]
    cosmo[:, 2] = 1.0 - col["omega_m0"] - orad
    cosmo[:, 3] = orad
    cosmo[:, 4] = 2.726
    cosmo[:, 5] = col["h"]
    cosmo[:, 6] = col["sigma_8"]
    cosmo[:, 7] = col["n_scalar"]
    cosmo[:, 8] = -1.0
    cosmo[:, 9] = 0.0
    halo = np.zeros((n_points, _lib.N_HALO))
    halo[:, 0] = 0.3
    halo[:, 1] = 0.707
    halo[:, 2] = col["c0"]
    halo[:, 3] = col["beta"]
    halo[:, 4] = -1.0
    halo[:, 5] = -1.0
    hodp = np.zeros((n_points, _lib.N_HOD))
    if hod == "zheng":
        hodp[:, 0] = col["log_M_min"]
        hodp[:, 1] = col["sigma"]
        hodp[:, 2] = col["log_M_min"]
        hodp[:, 3] = col["log_M_1p"]
        hodp[:, 4] = col["alpha"]
    else:
        hodp[:, 0] = col["log_M_0"]
        hodp[:, 1] = col["w"]
    return cosmo, halo, hodp


def as_dicts(cosmo, halo, hodp, hod="zheng"):
    """Row i of the arrays as the reference's three parameter dictionaries."""
    keys = _lib.HOD_ZHENG_KEYS if hod == "zheng" else _lib.HOD_MANDELBAUM_KEYS
    out = []
    for c, h, g in zip(cosmo, halo, hodp):
        out.append((dict(zip(_lib.COSMO_KEYS, map(float, c))),
                    dict(zip(_lib.HALO_KEYS, map(float, h))),
                    dict(zip(keys, map(float, g[:len(keys)])))))
    return out


def shard(n_points, rank, world_size):
    """Contiguous slice of the batch owned by `rank`."""
    base, extra = divmod(n_points, world_size)
    start = rank*base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))
