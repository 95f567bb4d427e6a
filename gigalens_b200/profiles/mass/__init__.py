from . import epl, shear, sie, sis, nfw, piemd, scaling_relation, dpie_subhalo  # noqa: F401
