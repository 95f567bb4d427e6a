from typing import Dict, List

import numpy as np

from ...profile import MassProfile


class ScalingRelation(MassProfile):
    """Sum of one mass profile over a galaxy catalogue with luminosity scaling relations
    (reference ``tf/profiles/mass/scaling_relation.py:6-70``).

    The free parameters are the ``scaling_params`` (value at ``L = lum_star``); member ``g`` gets
    ``scale * (L_g / lum_star) ** power``.  Every other parameter of the wrapped profile is a
    catalogue column.  The CUDA kernel sums the members inside one thread per pixel -- the
    ``(N, bs, G)`` intermediate of the reference is never materialised, so ``chunk_size`` is
    accepted and ignored.
    """

    def __init__(self, profile: MassProfile, scaling_params: List, lum_star: float,
                 scaling_params_power: Dict[str, float], galaxy_catalogue: Dict[str, List], chunk_size=None, **kwargs):
        self.profile = profile
        self._name = f"Scaled-{profile.name}"
        self._params = list(scaling_params)
        self._type_id = profile._type_id
        super().__init__(**kwargs)
        self.scaling_params = list(scaling_params)
        self.lum_star = float(lum_star)
        self.power = {k: float(v) for k, v in scaling_params_power.items()}
        self.galaxy_cat = galaxy_catalogue
        self._luminosities = np.asarray(galaxy_catalogue["lum"], dtype=np.float32)
        self.n_galaxy = len(self._luminosities)
        self.chunk_size = self.n_galaxy if chunk_size is None else chunk_size
        self.not_scaling_params = [p for p in profile.params if p not in self.scaling_params]
        for k in self.not_scaling_params:
            if k not in galaxy_catalogue:
                raise KeyError(f"galaxy_catalogue lacks column '{k}' needed by {profile.name}")

    def member_factors(self, raw_order):
        """[n_raw][G] fp32: (L/L*)^power for scaled parameters, the catalogue value otherwise."""
        out = np.ones((len(raw_order), self.n_galaxy), dtype=np.float32)
        for i, k in enumerate(raw_order):
            if k in self.scaling_params:
                # fp32 like the reference (scaling_relation.py:52-54)
                out[i] = (self._luminosities / np.float32(self.lum_star)) ** np.float32(self.power[k])
            else:
                out[i] = np.asarray(self.galaxy_cat[k], dtype=np.float32)
        return out
