"""Plumbing shared by the drop-in classes: one-point batches through `Engine`."""
import numpy as np

from . import _lib, defaults, engine


def cosmo_row(cosmo_dict):
    return engine.pack_params([cosmo_dict], _lib.COSMO_KEYS)


def halo_row(halo_dict, profile=None):
    """[stq, st_little_a] from the mass function's dictionary, [c0, beta, alpha,
    delta_v] from the halo-profile state (they can differ after Halo.set_halo,
    reference halo.py:196-212)."""
    row = engine.pack_params([halo_dict], _lib.HALO_KEYS)
    if profile is not None:
        row[0, 2], row[0, 3], row[0, 5] = profile["c0"], profile["beta"], profile["delta_v"]
    return row


def hod_row(kind, hod_dict):
    keys = _lib.HOD_ZHENG_KEYS if kind == _lib.HOD_ZHENG else _lib.HOD_MANDELBAUM_KEYS
    row = np.zeros((1, _lib.N_HOD))
    row[0, :len(keys)] = [float(hod_dict[k]) for k in keys]
    return row


def base_config(**kw):
    """A valid configuration from the *current* module-level defaults
    (reference semantics: defaults are read at call time) with harmless
    placeholders for whatever the caller does not use."""
    p, lim, q = defaults.default_precision, defaults.default_limits, defaults.default_quadrature
    c = _lib.Config()
    c.n_cosmo, c.n_mass, c.n_halo = p["cosmo_npoints"], p["mass_npoints"], p["halo_npoints"]
    c.n_window, c.n_kernel = p["window_npoints"], p["kernel_npoints"]
    c.nq_nu, c.nq_hankel, c.nq_limber, c.nq_lens = q["nu"], q["hankel"], q["limber"], q["lens"]
    c.halo_precision, c.cosmo_precision = p["halo_precision"], p["cosmo_precision"]
    c.window_precision = p["window_precision"]
    c.k_min, c.k_max = lim["k_min"], lim["k_max"]
    c.mass_min, c.mass_max = lim["mass_min"], lim["mass_max"]
    c.zk_min, c.zk_max = 0.0, 5.0
    for i in range(2):
        c.window_kind[i], c.dndz_kind[i] = _lib.WINDOW_GALAXY, _lib.DNDZ_GAUSSIAN
        c.dndz_zmin[i], c.dndz_zmax[i] = 0.0, 2.0
        c.dndz_p[i][0], c.dndz_p[i][1], c.dndz_p[i][2] = 1.0, 0.2, 0.0
    c.ktheta_min, c.ktheta_max = 1e-6*engine.DEG_TO_RAD, 100.0*engine.DEG_TO_RAD
    c.bessel_order = 0
    c.bessel_limit = engine.bessel_limit(0, p["kernel_bessel_limit"])
    c.corr_k_min = c.corr_k_max = -1.0
    c.tri_moment = -1
    for name, value in kw.items():
        setattr(c, name, value)
    return c


def set_window(cfg, slot, window):
    """Fill window / dN/dz slot `slot` of a config from a WindowFunction object."""
    d = window._redshift_dist
    cfg.window_kind[slot] = window._kind
    cfg.dndz_kind[slot] = d._kind
    cfg.dndz_zmin[slot], cfg.dndz_zmax[slot] = float(d.z_min), float(d.z_max)
    for j, v in enumerate(d._params()):
        cfg.dndz_p[slot][j] = float(v)
    if d._kind == _lib.DNDZ_TABLE:      # the table itself travels when an engine takes the config
        cfg.__dict__.setdefault("_dndz_uploads", {})[slot] = d


def like_input(x, values):
    """Return `values` shaped like the caller's argument (float in, float out)."""
    values = np.asarray(values, dtype=np.float64)
    if np.ndim(x) == 0:
        return np.float64(values.reshape(-1)[0])
    return values.reshape(np.shape(x))


def flat(x):
    return np.atleast_1d(np.asarray(x, dtype=np.float64)).reshape(-1)


class OnePoint(object):
    """An Engine used with batches of one parameter point."""

    def __init__(self):
        self.eng = engine.Engine()

    def configure(self, cfg):
        up = getattr(cfg, "_dndz_uploads", {})
        if len(up) == 2 and up[0] is up[1]:
            up[0]._upload(self.eng, 2)
        else:
            for slot, d in up.items():
                d._upload(self.eng, slot)
        self.eng.configure(cfg)

    def ev(self, what, x, aux=0.0):
        return self.eng.evaluate(what, flat(x), 0, aux).cpu().numpy()

    def table(self, table_id):
        return self.eng.table(table_id, 1).cpu().numpy()[0]

    def epoch(self):
        return dict(zip(_lib.EPOCH_FIELDS, self.table(_lib.T_EPOCH)))
