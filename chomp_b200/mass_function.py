"""Drop-in for the reference's mass_function.MassFunction (Sheth-Tormen;
mass_function.py:25-363), MassFunctionSecondOrder (:365-433) and TinkerMassFunction (Tinker et al. 2010, :436-564)."""
import numpy as np

from . import _facade, _lib, cosmology, defaults


class MassFunction(object):
    _mf_kind = _lib.MF_SHETH_TORMEN

    def __init__(self, redshift=0.0, cosmo_single_epoch=None, halo_dict=None, **kws):
        self._redshift = redshift
        if cosmo_single_epoch is None:
            cosmo_single_epoch = cosmology.SingleEpoch(self._redshift)
        self.cosmo = cosmo_single_epoch
        self.cosmo.set_redshift(self._redshift)
        if halo_dict is None:
            halo_dict = defaults.default_halo_dict
        self.halo_dict = halo_dict
        self._gpu = _facade.OnePoint()
        self._refresh()

    def _refresh(self):
        """mass_function.py:160-241: mass limits, nu(M) table, normalisations."""
        h = self.halo_dict
        self.stq, self.st_little_a = h["stq"], h["st_little_a"]
        self.c0 = h["c0"]/(1.0 + self._redshift)
        self._gpu.configure(_facade.base_config(with_bao=int(bool(getattr(self.cosmo, "_with_bao", False))),
                                                mass_function_kind=self._mf_kind))
        self._gpu.eng.mass_tables(_facade.cosmo_row(self.cosmo.cosmo_dict), _facade.halo_row(h),
                                  [self._redshift])
        e = self._gpu.epoch()
        self.delta_c, self.delta_v = e["delta_c"], e["delta_v"]
        self.ln_mass_min, self.ln_mass_max = e["ln_mass_min"], e["ln_mass_max"]
        self.nu_min, self.nu_max = e["nu_min"], e["nu_max"]
        self.f_norm, self.bias_norm = e["f_norm"], e["bias_norm"]
        self.m_star = float(np.exp(e["ln_m_star"]))
        self._ln_mass_array = self._gpu.table(_lib.T_LNM_NODES).copy()
        self._nu_array = self._gpu.table(_lib.T_NU_NODES).copy()

    def get_redshift(self):
        return self._redshift

    def set_redshift(self, redshift):
        self._redshift = redshift
        self.cosmo.set_redshift(redshift)
        self._refresh()

    def get_cosmology(self):
        return self.cosmo.get_cosmology()

    def set_cosmology(self, cosmo_dict, redshift=None):
        if redshift is None:
            redshift = self._redshift
        self._redshift = redshift
        self.cosmo.set_cosmology(cosmo_dict, redshift)
        self._refresh()

    def set_cosmology_object(self, cosmo_single_epoch):
        self._redshift = cosmo_single_epoch.redshift()
        self.cosmo = cosmo_single_epoch
        self._refresh()

    def get_halo(self):
        return self.halo_dict

    def set_halo(self, halo_dict):
        self.halo_dict = halo_dict
        self._refresh()

    def f_nu(self, nu):
        return _facade.like_input(nu, self._gpu.ev(_lib.EVAL_F_NU, nu))

    def f_m(self, mass):
        return self.f_nu(self.nu(mass))

    def bias_nu(self, nu):
        return _facade.like_input(nu, self._gpu.ev(_lib.EVAL_BIAS_NU, nu))

    def bias_m(self, mass):
        return self.bias_nu(self.nu(mass))

    def nu(self, mass):
        return _facade.like_input(mass, self._gpu.ev(_lib.EVAL_NU_OF_MASS, mass))

    def mass(self, nu):
        return _facade.like_input(nu, self._gpu.ev(_lib.EVAL_MASS_OF_NU, nu))

    def ln_mass(self, nu):
        return np.log(self.mass(nu))

    def write(self, output_file_name):
        print("M* = 10^%1.4f M_sun" % np.log10(self.m_star))
        with open(output_file_name, "w") as f:
            f.write("#ttype1 = mass [M_solar/h]\n#ttype2 = nu\n#ttype3 = f(nu)\n#ttype4 = bias(nu)\n")
            fn, bn = self.f_nu(self._nu_array), self.bias_nu(self._nu_array)
            for lm, nu, a, b in zip(self._ln_mass_array, self._nu_array, fn, bn):
                f.write("%1.10f %1.10f %1.10f %1.10f\n" % (np.exp(lm), nu, a, b))


class MassFunctionSecondOrder(MassFunction):
    """mass_function.py:365-433: adds the sigma(nu) spline and the second-order bias b_2(nu)."""

    def _refresh(self):
        MassFunction._refresh(self)
        self.bias_2_norm = float(self._gpu.eng.mass_second_order(1).cpu().numpy()[0])
        self._sigma_array = self.delta_c/np.sqrt(self._nu_array)

    def bias_2_nu(self, nu):
        return _facade.like_input(nu, self._gpu.ev(_lib.EVAL_BIAS_2_NU, nu))

    def bias_2_mass(self, mass):
        return self.bias_2_nu(self.nu(mass))


class TinkerMassFunction(MassFunction):
    """mass_function.py:436-564: f(nu) and b(nu) of Tinker et al. (2010), shape parameters from the reference's cubic
    splines in ln(Delta_v) scaled with redshift; the multiplicity function is not normalised (``f_norm`` = 1 here, the
    reference has no such attribute), the bias is."""
    _mf_kind = _lib.MF_TINKER
