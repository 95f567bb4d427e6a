from . import epl, shear, sie, sis, nfw, piemd, scaling_relation, dpie_subhalo, tnfw, piep  # noqa: F401
