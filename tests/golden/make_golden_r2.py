#!/usr/bin/env python
"""Generate tests/golden/reference_r2.json by running the REFERENCE ITSELF (oracle/_ref) on the round-2
boundary cases: Correlation(k_min=, k_max=) (correlation.py:104-112, 242-275) with limits wider and narrower
than the halo's k range, inputs = the dictionaries of unit_test.py:59-119.

    python oracle/make_ref.py && python tests/golden/make_golden_r2.py [section ...]

Sections already present in the JSON file are kept unless named on the command line."""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from common import C_DICT, D2R, H_DICT, HOD_DICT  # noqa: E402

PATH = os.path.join(HERE, "reference_r2.json")


def arr(x):
    return [float(v) for v in np.asarray(x, dtype=float).ravel()]


def make_kernel(R, z0=0.5, sigma=0.1):
    cm = R["cosmology"].MultiEpoch(0.0, 5.0, cosmo_dict=C_DICT)
    dist = R["kernel"].dNdzGaussian(0.0, 2.0, z0, sigma)
    wa, wb = R["kernel"].WindowFunctionGalaxy(dist, cm), R["kernel"].WindowFunctionGalaxy(dist, cm)
    return R["kernel"].Kernel(1e-6*D2R, 100.0*D2R, wa, wb, cm)


def k_limits(R):
    """Correlation(k_min=, k_max=): (a) wider than the halo range on both sides -> extrapolation switched on;
    (b) narrower on both sides; (c) only k_max given (Python 2: None < x, extrapolation on)."""
    out = {}
    kern = make_kernel(R)
    for name, kw in (("wide", dict(k_min=1e-4, k_max=1e3)), ("narrow", dict(k_min=5e-3, k_max=30.0)),
                     ("kmax_only", dict(k_max=300.0))):
        for spec in ("power_gg", "power_mm"):
            h = R["halo"].Halo(input_hod=R["hod"].HODZheng(HOD_DICT),
                               cosmo_single_epoch=R["cosmology"].SingleEpoch(0.0, cosmo_dict=C_DICT), halo_dict=H_DICT)
            corr = R["correlation"].Correlation(0.01, 1.0, kern, bins_per_decade=3.0, input_halo=h, power_spec=spec, **kw)
            corr.compute_correlation()
            out[name + "_" + spec] = {"theta": arr(corr.theta_array), "w": arr(corr.wtheta_array),
                                      "extrapolate": bool(h.get_extrapolation()), "args": {k: float(v) for k, v in kw.items()}}
    return out


HOD_DICT_B = {"log_M_min": 12.5, "sigma": 0.25, "log_M_0": 12.5, "log_M_1p": 13.8, "alpha": 1.1}
CROSS = {"theta_deg": (0.01, 1.0), "tri_z": 0.5, "area_deg2": 25.0, "n_a": [1e10, 1e10], "n_b": [1e10, 1e10], "variance": 1.0,
         "dist_a": (0.0, 2.0, 0.5, 0.1), "dist_b": (0.0, 2.0, 0.6, 0.1), "hod_b": HOD_DICT_B}


def cross_cov(R):
    """covariance.Covariance between two DIFFERENT correlations (covariance.py:60-63 matching_corrs False): galaxy
    clustering of two overlapping redshift slices with different HODs, power_gg, 1-halo trispectrum gggg; and the
    2 x 2 block matrix of CovarianceMulti (covariance.py:794-871; its blocks use the default power_mm)."""
    import importlib
    import time
    covariance = importlib.import_module("covariance")
    t0 = time.time()
    c = CROSS

    def make_corr(dist_args, hod_dict):
        kern = make_kernel(R, dist_args[2], dist_args[3])
        h = R["halo"].Halo(input_hod=R["hod"].HODZheng(hod_dict), cosmo_single_epoch=R["cosmology"].SingleEpoch(0.0, cosmo_dict=C_DICT),
                           halo_dict=H_DICT)
        return R["correlation"].Correlation(c["theta_deg"][0], c["theta_deg"][1], kern, bins_per_decade=5.0, input_halo=h,
                                            power_spec="power_gg")

    def make_tri():
        cs_t = R["cosmology"].SingleEpoch(c["tri_z"], cosmo_dict=C_DICT)
        mf_t = R["mass_function"].MassFunction(c["tri_z"], cs_t, H_DICT)
        return R["halo_trispectrum"].HaloTrispectrumOneHalo(c["tri_z"], cs_t, mf_t, None, H_DICT, R["hod"].HODZheng(HOD_DICT), "power_gggg")

    corr_a, corr_b = make_corr(c["dist_a"], HOD_DICT), make_corr(c["dist_b"], c["hod_b"])
    tri = make_tri()
    cov = covariance.Covariance(corr_a, corr_b, bins_per_decade=5.0, survey_area_deg2=c["area_deg2"], n_a=c["n_a"], n_b=c["n_b"],
                                variance=c["variance"], nongaussian_cov=True, input_halo_trispectrum=tri, power_spec="power_gg")
    bins = cov.annular_bins
    n = len(bins)
    G, NG = np.zeros((n, n)), np.zeros((n, n))
    for i in range(n):
        for j in range(i, n):
            a, b = bins[i], bins[j]
            G[i, j] = G[j, i] = cov.covariance_G(a.center, b.center, a.delta, b.delta)
            NG[i, j] = NG[j, i] = cov.covariance_NG(a.center, b.center)
        print("cross row", i, "of", n, "%.0f s" % (time.time() - t0), flush=True)
    total = np.asarray(cov.get_covariance(), dtype=float)
    out = {"config": {k: (list(v) if isinstance(v, tuple) else v) for k, v in c.items()},
           "matching_corrs": bool(cov.matching_corrs), "equal_windows": [bool(x) for x in cov.equal_windows],
           "cosmic_shear": [bool(x) for x in cov.cosmic_shear], "z_bar_a": float(corr_a.kernel.z_bar), "z_bar_b": float(corr_b.kernel.z_bar),
           "z_bar_NG": float(cov.kernel.z_bar_NG), "bins_center": arr([b.center for b in bins]),
           "kernel_NG_table": arr(np.asarray(cov.kernel._kernel_array, dtype=float)),
           "cov_G": arr(G), "cov_NG": arr(NG), "cov": arr(total),
           "proj": {k: arr(getattr(cov, "_halo_%s_spline" % k)(cov._ln_K_array)) for k in ("a", "b", "ab", "ba")},
           "ln_K": arr(cov._ln_K_array)}
    # CovarianceMulti over the same two correlations, Gaussian + Poisson only (its non-Gaussian blocks repeat the above)
    multi = covariance.CovarianceMulti([corr_a, corr_b], bins_per_decade=5.0, survey_area_deg2=c["area_deg2"], n_a=c["n_a"],
                                       n_b=c["n_b"], variance=c["variance"], nongaussian_cov=False, input_halo_trispectrum=tri)
    out["multi_gaussian"] = arr(np.asarray(multi.get_covariance(), dtype=float))
    out["seconds"] = time.time() - t0
    return out


def bao(R):
    """SingleEpoch(with_bao=True) (cosmology.py:474-538): linear power, sigma(R), the mass function's nu nodes, the halo
    model spectra and w_gg(theta) on top of the wiggle transfer function."""
    out = {}
    k = np.logspace(-3.5, 2.5, 120)
    for z in (0.0, 0.5):
        cs = R["cosmology"].SingleEpoch(z, cosmo_dict=C_DICT, with_bao=True)
        out["z%.1f" % z] = {"k": arr(k), "linear_power": arr(cs.linear_power(k)), "sigma_norm": float(cs._sigma_norm),
                            "sigma_r": arr([cs.sigma_r(r) for r in (0.5, 2.0, 8.0, 30.0)]), "growth": float(cs._growth)}
    cs = R["cosmology"].SingleEpoch(0.0, cosmo_dict=C_DICT, with_bao=True)
    mf = R["mass_function"].MassFunction(0.0, cs, H_DICT)
    out["mass"] = {"nu_nodes": arr(mf._nu_array), "ln_mass_nodes": arr(mf._ln_mass_array), "f_norm": float(mf.f_norm),
                   "bias_norm": float(mf.bias_norm)}
    h = R["halo"].Halo(input_hod=R["hod"].HODZheng(HOD_DICT), cosmo_single_epoch=cs, halo_dict=H_DICT)
    kk = np.logspace(-3, 2, 60)
    out["halo"] = {"k": arr(kk), "power_mm": arr(h.power_mm(kk)), "power_gm": arr(h.power_gm(kk)), "power_gg": arr(h.power_gg(kk))}
    # Correlation moves the halo to z_bar with halo.set_redshift -> SingleEpoch.set_cosmology -> __init__ without
    # with_bao (Q16): the wiggles are gone from w(theta) unless the halo is kept at its own redshift
    kern = make_kernel(R)
    hz = R["halo"].Halo(redshift=float(kern.z_bar), input_hod=R["hod"].HODZheng(HOD_DICT),
                        cosmo_single_epoch=R["cosmology"].SingleEpoch(float(kern.z_bar), cosmo_dict=C_DICT, with_bao=True), halo_dict=H_DICT)
    corr = R["correlation"].Correlation(0.01, 1.0, kern, bins_per_decade=3.0, input_halo=hz, power_spec="power_gg", keep_halo_z_bar=True)
    corr.compute_correlation()
    out["wtheta_keep_z_bar"] = {"theta": arr(corr.theta_array), "w": arr(corr.wtheta_array), "with_bao_after": bool(hz.cosmo._with_bao)}
    h2 = R["halo"].Halo(input_hod=R["hod"].HODZheng(HOD_DICT), cosmo_single_epoch=R["cosmology"].SingleEpoch(0.0, cosmo_dict=C_DICT, with_bao=True),
                        halo_dict=H_DICT)
    corr2 = R["correlation"].Correlation(0.01, 1.0, kern, bins_per_decade=3.0, input_halo=h2, power_spec="power_gg")
    corr2.compute_correlation()
    out["wtheta_moved"] = {"theta": arr(corr2.theta_array), "w": arr(corr2.wtheta_array), "with_bao_after": bool(h2.cosmo._with_bao)}
    return out


def tinker(R):
    """TinkerMassFunction (mass_function.py:436-564) on its own and under a Halo (spectra, w_gg(theta))."""
    from common import H_DICT_2
    out = {}
    nu = np.logspace(-0.9, 1.6, 12)
    for z, hd, key in ((0.0, H_DICT, "z0.0"), (0.5, H_DICT_2, "z0.5_delta_v_200")):
        cs = R["cosmology"].SingleEpoch(z, cosmo_dict=C_DICT)
        mf = R["mass_function"].TinkerMassFunction(z, cs, hd)
        out[key] = {"nu": arr(nu), "f_nu": arr(mf.f_nu(nu)), "bias_nu": arr(mf.bias_nu(nu)), "bias_norm": float(mf.bias_norm),
                    "delta_v": float(mf.delta_v), "nu_nodes": arr(mf._nu_array), "m_star": float(mf.m_star)}
    cs = R["cosmology"].SingleEpoch(0.0, cosmo_dict=C_DICT)
    mf = R["mass_function"].TinkerMassFunction(0.0, cs, H_DICT)
    h = R["halo"].Halo(input_hod=R["hod"].HODZheng(HOD_DICT), cosmo_single_epoch=cs, mass_func=mf, halo_dict=H_DICT)
    kk = np.logspace(-3, 2, 60)
    out["halo"] = {"k": arr(kk), "power_mm": arr(h.power_mm(kk)), "power_gm": arr(h.power_gm(kk)), "power_gg": arr(h.power_gg(kk))}
    kern = make_kernel(R)
    corr = R["correlation"].Correlation(0.01, 1.0, kern, bins_per_decade=3.0, input_halo=h, power_spec="power_gg")
    corr.compute_correlation()
    out["wtheta"] = {"theta": arr(corr.theta_array), "w": arr(corr.wtheta_array)}
    return out


def ssc_halo(R):
    """HaloSuperSampleCovariance (halo.py:1089-1199): I^1_2(k), dln P / d delta_b, power_mm_ssc."""
    out = {}
    kk = np.concatenate([[5e-4], np.logspace(-3, 2, 40), [150.0]])
    for z in (0.0, 0.5):
        cs = R["cosmology"].SingleEpoch(z, cosmo_dict=C_DICT)
        h = R["halo"].HaloSuperSampleCovariance(z, R["hod"].HODZheng(HOD_DICT), cs, None, H_DICT, False, 0.02)
        out["z%.1f" % z] = {"k": arr(kk), "i_1_2": arr(h._i_1_2(kk)) if h._initialize_i_1_2() is None else None,
                            "dln_power_ddelta_b": arr(h.dln_power_ddelta_b(kk)), "power_mm_ssc": arr(h.power_mm_ssc(kk)),
                            "power_mm": arr(h.power_mm(kk)), "delta_b": 0.02}
    base = R["halo"].Halo(0.0, R["hod"].HODZheng(HOD_DICT), R["cosmology"].SingleEpoch(0.0, cosmo_dict=C_DICT), None, H_DICT)
    base.power_mm(1.0)
    h2 = R["halo"].HaloSuperSampleCovariance.init_from_halo(base, 0.01)
    out["init_from_halo"] = {"k": arr(kk), "dln_power_ddelta_b": arr(h2.dln_power_ddelta_b(kk))}
    return out


def xi3d(R):
    """Correlation3d (correlation.py:408-510) with the linear and the halo-model spectra."""
    out = {}
    for spec in ("linear_power", "power_mm", "power_gg"):
        h = R["halo"].Halo(0.3, R["hod"].HODZheng(HOD_DICT), R["cosmology"].SingleEpoch(0.3, cosmo_dict=C_DICT), None, H_DICT)
        c3 = R["correlation"].Correlation3d(0.05, 60.0, 0.3, input_halo=h, powSpec=spec)
        c3.compute_correlation()
        rq = np.array([0.04, 0.07, 1.3, 25.0, 60.0, 80.0])
        out[spec] = {"r": arr(c3.r_array), "xi": arr(c3.xi_array), "r_query": arr(rq), "xi_query": arr(c3.correlation(rq))}
    return out


SECTIONS = {"k_limits": k_limits, "cross_cov": cross_cov, "bao": bao, "tinker": tinker, "ssc_halo": ssc_halo, "xi3d": xi3d}


def main():
    os.chdir(tempfile.mkdtemp(prefix="chomp_golden_"))
    R = oracle.import_ref()
    out = json.load(open(PATH)) if os.path.exists(PATH) else {}
    want = sys.argv[1:] or [s for s in SECTIONS if s not in out]
    for s in want:
        print("running", s)
        out[s] = SECTIONS[s](R)
        with open(PATH, "w") as f:
            json.dump(out, f)
    print("wrote", PATH, os.path.getsize(PATH), "bytes")


if __name__ == "__main__":
    main()
