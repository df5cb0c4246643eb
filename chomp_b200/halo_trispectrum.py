"""Drop-in for the reference's halo_trispectrum.HaloTrispectrumOneHalo
(halo_trispectrum.py:13-151): the 1-halo trispectrum table
I^0_4(k1, k1, k2, k2) and its bicubic interpolation.  The full HaloTrispectrum
(2-, 3-, 4-halo terms with perturbation theory) is listed as next in SURVEY.md
section 8(f)."""
import numpy as np

from . import _facade, _lib, halo, hod as hod_module


class HaloTrispectrumOneHalo(halo.Halo):
    def __init__(self, redshift=0.0, single_epoch_cosmo=None, mass_func_second=None, perturbation=None,
                 halo_dict=None, input_hod=None, power_spec="power_mmmm"):
        self.pert = perturbation
        halo.Halo.__init__(self, redshift, None, single_epoch_cosmo, mass_func_second, halo_dict)
        self.power_spec = power_spec
        if input_hod is None:
            input_hod = hod_module.HODZheng()
        self.local_hod = input_hod
        self._initialized_i_0_4 = False

    def set_cosmology(self, cosmo_dict, redshift=None):
        halo.Halo.set_cosmology(self, cosmo_dict, redshift)
        if self.pert is not None:
            self.pert.set_cosmology_object(self.cosmo)
        self._initialized_i_0_4 = False

    def _ensure(self):
        was_dirty = self._dirty
        halo.Halo._ensure(self)
        if was_dirty:
            self._initialized_i_0_4 = False

    def _initialize_i_0_4(self):
        self._dirty = True                       # the moment kind is part of the configuration
        cfg_moment = _lib.TRISPECTRUM_MOMENT.get(self.power_spec, 0)
        self._tri_moment = cfg_moment
        halo.Halo._ensure(self)
        self._i_0_4_array = self._gpu.eng.trispectrum_1h(1).cpu().numpy()[0]
        self._initialized_i_0_4 = True

    def trispectrum_parallelogram(self, k1, k2):
        if not self._initialized_i_0_4:
            self._initialize_i_0_4()
        out = self._gpu.eng.trispectrum_eval(_facade.flat(k1), _facade.flat(k2)).cpu().numpy()
        return _facade.like_input(k1, out)

    def i_0_4_parallelogram(self, k1, k2):
        return self.trispectrum_parallelogram(k1, k2)
