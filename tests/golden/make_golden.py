#!/usr/bin/env python
"""Generate tests/golden/reference_outputs.json by running the REFERENCE ITSELF
(the generated copy oracle/_ref, see oracle/make_ref.py) in this container.
/root/reference does not exist on the GPU box, so the vectors are committed.

    python oracle/make_ref.py && python tests/golden/make_golden.py

Every entry is the reference at its default tolerances (defaults.py:62-92),
inputs = the dictionaries of unit_test.py:59-119.
"""
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
from common import (C_DICT, C_DICT_2, D2R, H_DICT, H_DICT_2, HOD_DICT,  # noqa: E402
                    HOD_DICT_2)

MAND = {"log_M_0": 12.3, "w": 1.2}


def arr(x):
    return [float(v) for v in np.asarray(x, dtype=float).ravel()]


def main():
    os.chdir(tempfile.mkdtemp(prefix="chomp_golden_"))   # Kernel.__init__ writes debug files into the CWD
    R = oracle.import_ref()
    cosmology, mass_function, hod, halo, kernel, correlation = (
        R[k] for k in ("cosmology", "mass_function", "hod", "halo", "kernel", "correlation"))
    out = {"k": arr(np.logspace(-3, 2, 200)), "masses": arr(np.logspace(9, 16, 8))}
    k = np.array(out["k"])
    M = np.array(out["masses"])

    # ---- cosmology + mass function + halo spectra (config 1, unit_test.py:131-407) -------------
    cases = {"base": (C_DICT, H_DICT, HOD_DICT, None), "cosmo2": (C_DICT_2, H_DICT, HOD_DICT, None),
             "hod2": (C_DICT, H_DICT, HOD_DICT_2, None), "set_halo2": (C_DICT, H_DICT, HOD_DICT, H_DICT_2)}
    out["halo"] = {}
    for name, (cd, hd, gd, set_halo) in cases.items():
        cs = cosmology.SingleEpoch(0.0, cosmo_dict=cd)
        h = halo.Halo(input_hod=hod.HODZheng(gd), cosmo_single_epoch=cs, halo_dict=hd)
        if set_halo is not None:
            h.set_halo(set_halo)          # unit_test.py:388-398: keeps stale profile splines
        entry = {"linear_power": arr(h.linear_power(k)), "power_mm": arr(h.power_mm(k)),
                 "power_gm": arr(h.power_gm(k)), "power_gg": arr(h.power_gg(k)),
                 "nu": arr(h.mass.nu(M)), "f_nu": arr(h.mass.f_nu(h.mass.nu(M))),
                 "bias_nu": arr(h.mass.bias_nu(h.mass.nu(M))), "nu_nodes": arr(h.mass._nu_array),
                 "ln_mass_nodes": arr(h.mass._ln_mass_array), "n_bar_over_rho_bar": float(h.n_bar_over_rho_bar),
                 "sigma_8": float(cs.sigma_r(8.0)), "delta_c": float(cs.delta_c()), "delta_v": float(cs.delta_v()),
                 "rho_bar": float(cs.rho_bar()), "sigma_norm": float(cs._sigma_norm)}
        for nm, sp in (("h_m", h._h_m_spline), ("pp_mm", h._pp_mm_spline), ("h_g", h._h_g_spline),
                       ("pp_gm", h._pp_gm_spline), ("pp_gg", h._pp_gg_spline)):
            entry[nm] = arr(sp(h._ln_k_array))
        out["halo"][name] = entry
    cs = cosmology.SingleEpoch(0.0, cosmo_dict=C_DICT)
    hx = halo.HaloExclusion(input_hod=hod.HODZheng(HOD_DICT), cosmo_single_epoch=cs, halo_dict=H_DICT)
    out["halo"]["exclusion"] = {"power_mm": arr(hx.power_mm(k)), "power_gm": arr(hx.power_gm(k)),
                                "power_gg": arr(hx.power_gg(k)), "h_m": arr(hx._h_m_spline(hx._ln_k_array)),
                                "h_g": arr(hx._h_g_spline(hx._ln_k_array))}
    z = hod.HODZheng(HOD_DICT)
    out["hod_zheng"] = {"first": arr(z.first_moment(M)), "second": arr(z.second_moment(M)),
                        "third": arr(z.nth_moment(M, 3))}
    m = hod.HODMandelbaum(MAND)
    out["hod_mandelbaum"] = {"params": MAND, "first": arr(m.first_moment(M)), "second": arr(m.second_moment(M))}

    # ---- Limber kernels and correlations (configs 1-3) -------------------------------------------
    def corr_case(cd, dist_a, dist_b, wa_cls, wb_cls, kern_cls, hod_obj, spec, bpd):
        cm = cosmology.MultiEpoch(0.0, 5.0, cosmo_dict=cd)
        wa = wa_cls(dist_a, cm)
        wb = wb_cls(dist_b, cm)
        kern = kern_cls(1e-6*D2R, 100.0*D2R, wa, wb, cm)
        cs = cosmology.SingleEpoch(0.0, cosmo_dict=cd)
        h = halo.Halo(input_hod=hod_obj, cosmo_single_epoch=cs, halo_dict=H_DICT)
        c = correlation.Correlation(0.001, 1.0, kern, bins_per_decade=bpd, input_halo=h, power_spec=spec)
        c.compute_correlation()
        return {"theta": arr(c.theta_array), "w": arr(c.wtheta_array), "z_bar": float(kern.z_bar),
                "D_z": float(c.D_z), "chi_nodes": arr(cm._chi_array),
                "kernel_nodes": arr(np.asarray(kern._kernel_array, dtype=float)),
                "window_a": arr(np.asarray(kern.window_function_a._wf_array, dtype=float)),
                "window_b": arr(np.asarray(kern.window_function_b._wf_array, dtype=float)),
                "chi_range": [float(kern.chi_min), float(kern.chi_max)]}

    G = kernel.dNdzGaussian
    out["corr"] = {
        "cfg1_mm": corr_case(C_DICT, G(0.0, 2.0, 1.0, 0.2), G(0.0, 2.0, 1.0, 0.2), kernel.WindowFunctionGalaxy,
                             kernel.WindowFunctionGalaxy, kernel.Kernel, hod.HODZheng(HOD_DICT), "power_mm", 5.0),
        "cfg2_gg": corr_case(C_DICT, G(0.0, 2.0, 0.5, 0.1), G(0.0, 2.0, 0.5, 0.1), kernel.WindowFunctionGalaxy,
                             kernel.WindowFunctionGalaxy, kernel.Kernel, hod.HODZheng(HOD_DICT), "power_gg", 10.0),
        "cfg3_gammat": corr_case(C_DICT, G(0.0, 2.0, 0.4, 0.1), G(0.0, 2.0, 1.0, 0.2), kernel.WindowFunctionGalaxy,
                                 kernel.WindowFunctionConvergence, kernel.GalaxyGalaxyLensingKernel,
                                 hod.HODMandelbaum(MAND), "power_gm", 5.0),
        "maglim_conv": corr_case(C_DICT, kernel.dNdzMagLim(0.0, 2.0, 2, 0.3, 2), G(0.0, 2.0, 1.0, 0.2),
                                 kernel.WindowFunctionGalaxy, kernel.WindowFunctionConvergence, kernel.Kernel,
                                 hod.HODZheng(HOD_DICT), "power_gm", 5.0),
    }
    # ---- config 4: HaloFit shear-shear C(l), magnitude-limited dN/dz (examples/shear_shear_spectrum.py) ----
    cm = cosmology.MultiEpoch(0.0, 5.0, cosmo_dict=C_DICT)
    win = kernel.WindowFunctionConvergence(kernel.dNdzMagLim(0.0, 2.0, 2.0, 0.5, 2.0), cm)
    kern = kernel.Kernel(1e-6*D2R, 100.0*D2R, win, win, cm)
    hf = halo.HaloFit(input_hod=hod.HODZheng(HOD_DICT), cosmo_single_epoch=cosmology.SingleEpoch(0.0, cosmo_dict=C_DICT),
                      halo_dict=H_DICT)
    cf = correlation.CorrelationFourier(10, 1e5, kern, input_halo=hf, powSpec="power_mm")
    ell = np.logspace(1, 5, 30)
    cl = [float(cf.correlation(l)) for l in ell]
    out["cfg4_halofit"] = {
        "ell": arr(ell), "cl": cl, "z_bar": float(kern.z_bar), "k": arr(k),
        "power_mm": arr(hf.power_mm(k)), "power_gm": arr(hf.power_gm(k)), "power_gg": arr(hf.power_gg(k)),
        "fit": {n: float(getattr(hf, "_" + n)) for n in ("k_s", "n_eff", "C", "a_n", "b_n", "c_n", "gamma_n",
                                                        "alpha_n", "beta_n", "nu_n", "f_1", "f_2", "f_3")}}
    cf_lin = correlation.CorrelationFourier(10, 1e5, kern, input_halo=hf, powSpec="linear_power")
    out["cfg4_halofit"]["cl_linear"] = [float(cf_lin.correlation(l)) for l in ell]

    # ---- 1-halo trispectrum (config 5's table; halo_trispectrum.py:58-140) ------------------------
    tri_mod = R["halo_trispectrum"]
    out["trispectrum"] = {}
    kq1 = np.array([1e-4, 0.01, 0.5, 3.0, 30.0, 200.0])
    kq2 = np.array([0.02, 0.02, 7.0, 50.0, 30.0, 1.0])
    for spec, gd in (("power_mmmm", HOD_DICT), ("power_ggmm", HOD_DICT)):
        cs = cosmology.SingleEpoch(0.0, cosmo_dict=C_DICT)
        mf = mass_function.MassFunction(0.0, cs, H_DICT)
        tri = tri_mod.HaloTrispectrumOneHalo(0.0, cs, mf, None, H_DICT, hod.HODZheng(gd), spec)
        tri._initialize_i_0_4()
        out["trispectrum"][spec] = {
            "table": arr(tri._i_0_4_array), "k1": arr(kq1), "k2": arr(kq2),
            "parallelogram": [float(tri.trispectrum_parallelogram(a, b)) for a, b in zip(kq1, kq2)]}
    path = os.path.join(HERE, "reference_outputs.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
