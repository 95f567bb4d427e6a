"""TEST INFRASTRUCTURE -- ``lenstronomy...Shapelets.phi_n`` (third party, not vendored): the 1-D Gauss-Hermite function
phi_n(x) = H_n(x) exp(-x^2/2) / sqrt(2^n sqrt(pi) n!), from the oracle's restatement."""
import numpy as np

from oracle.profiles import shapelet_phi_n_np


class Shapelets:
    def __init__(self, *a, **k):
        pass

    def phi_n(self, n, x):
        return shapelet_phi_n_np(int(n), np.asarray(x, dtype=np.float64))
