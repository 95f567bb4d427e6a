"""Oracle restatement of the reference's TF mass and light profiles (TEST INFRASTRUCTURE).

Each class follows one reference file under ``src/gigalens/tf/profiles/`` and keeps the
reference's operation order, clamps and ``where`` conventions.  Parameters are torch
tensors (or python floats) broadcastable against the coordinates, exactly as in the
reference (``tf/simulator.py:75-76``: coordinates ``(N, bs)``, parameters ``(bs,)``).
Autograd through these functions is the check for the CUDA hand adjoints.
"""
import math

import numpy as np
import torch


def _t(v, like):
    """Promote python scalars to tensors of the coordinate dtype."""
    if torch.is_tensor(v):
        return v
    return torch.as_tensor(v, dtype=like.dtype)


def _rotate(x, y, phi):
    # epl.py:59-64 (same body in sie.py:44-49, nfw.py:122-126, piemd.py:148-152)
    cos_phi, sin_phi = torch.cos(phi), torch.sin(phi)
    return x * cos_phi + y * sin_phi, -x * sin_phi + y * cos_phi


# --------------------------------------------------------------------------- mass


class MassBase:
    """``tf/profile.py:6-43``: default ``hessian`` by differentiating ``deriv`` (``GradientTape`` there, autograd
    here, graph kept so that the result can itself be differentiated w.r.t. the parameters), and the
    ``convergence`` / ``shear`` derived from it."""

    def hessian(self, x, y, **kwargs):
        x = x if x.requires_grad else x.detach().clone().requires_grad_(True)
        y = y if y.requires_grad else y.detach().clone().requires_grad_(True)
        fx, fy = self.deriv(x, y, **kwargs)  # tf/profile.py:22-25
        f_xx, f_xy = torch.autograd.grad(fx.sum(), [x, y], create_graph=True)  # :27
        f_yx, f_yy = torch.autograd.grad(fy.sum(), [x, y], create_graph=True)  # :28
        return f_xx, f_xy, f_yx, f_yy

    def convergence(self, x, y, **kwargs):
        f_xx, f_xy, f_yx, f_yy = self.hessian(x, y, **kwargs)
        return (f_xx + f_yy) / 2  # :34-35

    def shear(self, x, y, **kwargs):
        f_xx, f_xy, f_yx, f_yy = self.hessian(x, y, **kwargs)
        return (f_xx - f_yy) / 2, f_xy  # :40-43



class EPL(MassBase):
    """``tf/profiles/mass/epl.py:5-64``."""

    name = "EPL"
    params = ["theta_E", "gamma", "e1", "e2", "center_x", "center_y"]

    def __init__(self, niter=50):
        self.niter = niter  # epl.py:15-17

    def deriv(self, x, y, theta_E, gamma, e1, e2, center_x, center_y):
        theta_E, gamma, e1, e2, center_x, center_y = (_t(v, x) for v in (theta_E, gamma, e1, e2, center_x, center_y))
        phi = torch.atan2(e2, e1) / 2  # epl.py:21
        c = torch.clamp(torch.sqrt(e1 ** 2 + e2 ** 2), 0, 1)  # :22
        q = (1 - c) / (1 + c)  # :23
        theta_E_conv = theta_E / (torch.sqrt((1.0 + q ** 2) / (2.0 * q)))  # :24
        b = theta_E_conv * torch.sqrt((1 + q ** 2) / 2)  # :25
        t = gamma - 1  # :26

        x, y = x - center_x, y - center_y  # :28
        x, y = _rotate(x, y, phi)  # :29

        R = torch.clamp(torch.sqrt((q * x) ** 2 + y ** 2), 1e-10, 1e10)  # :31
        angle = torch.atan2(y, q * x)  # :32
        f = (1 - q) / (1 + q)  # :33
        Cs, Ss = torch.cos(angle), torch.sin(angle)  # :34
        Cs2, Ss2 = torch.cos(2 * angle), torch.sin(2 * angle)  # :35

        # :37  batch-global, under stop_gradient.  tf.math.log(1e-12) is evaluated in the
        # tensor dtype (fp32 in the reference).
        with torch.no_grad():
            niter = torch.log(torch.as_tensor(1e-12, dtype=x.dtype)) / torch.log(torch.max(f)) + 2
            niter = float(niter)

        # :39-54  while_loop(i < niter, maximum_iterations=self.niter), i starts at 1.0 and the
        # body uses the pre-increment value.
        last_x, last_y, f_x, f_y = Cs, Ss, Cs, Ss
        n = 1.0
        trips = 0
        while n < niter and trips < self.niter:
            prefac_ = -f * (2 * n - (2 - t)) / (2 * n + (2 - t))  # :41
            last_x, last_y = prefac_ * (Cs2 * last_x - Ss2 * last_y), prefac_ * (Ss2 * last_x + Cs2 * last_y)
            f_x, f_y = f_x + last_x, f_y + last_y
            n += 1.0
            trips += 1
        self.last_trips = trips
        prefac = (2 * b) / (1 + q) * torch.pow(b / R, t - 1)  # :55
        f_x, f_y = f_x * prefac, f_y * prefac  # :56
        return _rotate(f_x, f_y, -phi)  # :57


class Shear(MassBase):
    """``tf/profiles/mass/shear.py:5-26``."""

    name = "SHEAR"
    params = ["gamma1", "gamma2"]

    def deriv(self, x, y, gamma1, gamma2):
        return gamma1 * x + gamma2 * y, gamma2 * x - gamma1 * y  # shear.py:16

    def hessian(self, x, y, gamma1, gamma2):
        # shear.py:18-26 (kappa = 0)
        one = torch.ones_like(x)
        f_xx, f_yy, f_xy = gamma1 * one, -gamma1 * one, gamma2 * one
        return f_xx, f_xy, f_xy, f_yy


class SIE(MassBase):
    """``tf/profiles/mass/sie.py:5-49`` (core ``s_scale`` is shadowed to 0, ``:15``)."""

    name = "SIE"
    params = ["theta_E", "e1", "e2", "center_x", "center_y"]

    def _param_conv(self, theta_E, e1, e2):
        s_scale = 0  # sie.py:15
        phi = torch.atan2(e2, e1) / 2
        c = torch.clamp(torch.sqrt(e1 ** 2 + e2 ** 2), max=0.9999)  # minimum(.,0.9999) :17
        q = (1 - c) / (1 + c)
        theta_E_conv = theta_E / (torch.sqrt((1.0 + q ** 2) / (2.0 * q)))
        b = theta_E_conv * torch.sqrt((1 + q ** 2) / 2)
        s = s_scale * torch.sqrt((1 + q ** 2) / (2 * q ** 2))
        return b, s, q, phi

    def deriv(self, x, y, theta_E, e1, e2, center_x, center_y):
        theta_E, e1, e2, center_x, center_y = (_t(v, x) for v in (theta_E, e1, e2, center_x, center_y))
        b, s, q, phi = self._param_conv(theta_E, e1, e2)
        x, y = x - center_x, y - center_y
        x, y = _rotate(x, y, phi)
        psi = torch.sqrt(q ** 2 * (s ** 2 + x ** 2) + y ** 2)  # sie.py:29
        fx = b / torch.sqrt(1.0 - q ** 2) * torch.atan(torch.sqrt(1.0 - q ** 2) * x / (psi + s))
        fy = b / torch.sqrt(1.0 - q ** 2) * torch.atanh(torch.sqrt(1.0 - q ** 2) * y / (psi + q ** 2 * s))
        return _rotate(fx, fy, -phi)


class SIS(MassBase):
    """``tf/profiles/mass/sis.py:5-17``."""

    name = "SIS"
    params = ["theta_E", "center_x", "center_y"]

    def deriv(self, x, y, theta_E, center_x, center_y):
        theta_E = _t(theta_E, x)
        x, y = x - center_x, y - center_y
        R = torch.sqrt(x ** 2 + y ** 2)
        # tf.where(R == 0, 0.0, theta_E / R) -- safe denominator so the dead branch cannot
        # poison autograd (the reference has that hazard).
        zero = R == 0
        a = torch.where(zero, torch.zeros_like(R), theta_E / torch.where(zero, torch.ones_like(R), R))
        return a * x, a * y

    def hessian(self, x, y, theta_E, center_x, center_y):
        # sis.py:19-29
        theta_E = _t(theta_E, x)
        x, y = x - center_x, y - center_y
        R = (x ** 2 + y ** 2) ** (3.0 / 2)
        zero = R == 0
        a = torch.where(zero, torch.zeros_like(R), theta_E / torch.where(zero, torch.ones_like(R), R))
        f_xx, f_yy, f_xy = y ** 2 * a, x ** 2 * a, -x * y * a
        return f_xx, f_xy, f_xy, f_yy


class NFW(MassBase):
    """``tf/profiles/mass/nfw.py:5-52`` (``deriv``, ``nfwAlpha``, ``g_``)."""

    name = "NFW"
    params = ["Rs", "alpha_Rs", "center_x", "center_y"]
    _r_min = 0.0000001
    _c = 0.000001

    def deriv(self, x, y, Rs, alpha_Rs, center_x, center_y):
        Rs, alpha_Rs = _t(Rs, x), _t(alpha_Rs, x)
        rho0 = alpha_Rs / (4.0 * Rs ** 2 * (1.0 - math.log(2.0)))  # nfw.py:17
        x, y = x - center_x, y - center_y
        R = torch.sqrt(x ** 2 + y ** 2)
        return self.nfwAlpha(R, Rs, rho0, x, y)

    def nfwAlpha(self, R, Rs, rho0, ax_x, ax_y):
        R = torch.clamp(R, min=self._r_min)  # :26
        Rs = torch.clamp(Rs, min=self._r_min)  # :27
        x = R / Rs
        gx = self.g_(x)
        a = 4 * rho0 * Rs * gx / x ** 2  # :30
        return a * ax_x, a * ax_y

    def g_(self, x):
        # nfw.py:34-52: scatter-update of the x<1 and x>1 entries; x == 1 keeps the initial 1.0.
        x = torch.clamp(x, min=self._c)
        lt, gt = x < 1, x > 1
        x1 = torch.where(lt, x, torch.full_like(x, 0.5))
        x2 = torch.where(gt, x, torch.full_like(x, 2.0))
        a1 = torch.log(x1 / 2.0) + 1 / torch.sqrt(1 - x1 ** 2) * torch.acosh(1.0 / x1)
        a2 = torch.log(x2 / 2.0) + 1 / torch.sqrt(x2 ** 2 - 1) * torch.acos(1.0 / x2)
        a = torch.ones_like(x)
        a = torch.where(lt, a1, a)
        a = torch.where(gt, a2, a)
        return a

    def F_(self, x):
        # nfw.py:54-76: 1/3 at x == 1
        lt, gt = x < 1, x > 1
        x1 = torch.where(lt, x, torch.full_like(x, 0.5))
        x2 = torch.where(gt, x, torch.full_like(x, 2.0))
        a1 = 1 / (x1 ** 2 - 1) * (1 - 2 / torch.sqrt(1 - x1 ** 2) * torch.atanh(torch.sqrt((1 - x1) / (1 + x1))))
        a2 = 1 / (x2 ** 2 - 1) * (1 - 2 / torch.sqrt(x2 ** 2 - 1) * torch.atan(torch.sqrt((x2 - 1) / (1 + x2))))
        a = torch.ones_like(x) / 3
        a = torch.where(lt, a1, a)
        a = torch.where(gt, a2, a)
        return a

    def hessian(self, x, y, Rs, alpha_Rs, center_x, center_y):
        # nfw.py:78-94
        Rs, alpha_Rs = _t(Rs, x), _t(alpha_Rs, x)
        rho0 = alpha_Rs / (4.0 * Rs ** 2 * (1.0 - math.log(2.0)))
        Rs = torch.clamp(Rs, min=self._r_min)
        x, y = x - center_x, y - center_y
        R = torch.clamp(torch.sqrt(x ** 2 + y ** 2), min=self._c)
        X = R / Rs
        gx, Fx = self.g_(X), self.F_(X)
        kappa = 2 * rho0 * Rs * Fx
        a = 2 * rho0 * Rs * (2 * gx / X ** 2 - Fx)
        gamma1, gamma2 = a * (y ** 2 - x ** 2) / R ** 2, -a * 2 * (x * y) / R ** 2
        return kappa + gamma1, gamma2, gamma2, kappa - gamma1


class NFW_ELLIPSE(MassBase):
    """``tf/profiles/mass/nfw.py:97-134``."""

    name = "NFW_ELLIPSE"
    params = ["Rs", "alpha_Rs", "e1", "e2", "center_x", "center_y"]

    def __init__(self):
        self.nfw = NFW()

    def _param_conv(self, e1, e2):
        phi = torch.atan2(e2, e1) / 2
        c = torch.clamp(torch.sqrt(e1 ** 2 + e2 ** 2), max=0.9999)
        q = (1 - c) / (1 + c)
        e = torch.abs(1 - q ** 2) / (1 + q ** 2)
        return e, phi

    def deriv(self, x, y, Rs, alpha_Rs, e1, e2, center_x, center_y):
        Rs, alpha_Rs, e1, e2 = (_t(v, x) for v in (Rs, alpha_Rs, e1, e2))
        rho0 = alpha_Rs / (4.0 * Rs ** 2 * (1.0 - math.log(2.0)))
        e, phi = self._param_conv(e1, e2)
        x, y = x - center_x, y - center_y
        x, y = _rotate(x, y, phi)
        x, y = x * torch.sqrt(1 - e), y * torch.sqrt(1 + e)
        R = torch.sqrt(x ** 2 + y ** 2)
        fx, fy = self.nfw.nfwAlpha(R, Rs, rho0, x, y)
        fx = fx * torch.sqrt(1 - e)
        fy = fy * torch.sqrt(1 + e)
        return _rotate(fx, fy, -phi)


def _sort_ra_rs(r_core, r_cut, r_min):
    # piemd.py:51-60 / :189-199 -- note the second `where` sees the already-updated r_core.
    r_core = torch.where(r_core < r_cut, r_core, r_cut)
    r_cut = torch.where(r_core > r_cut, r_core, r_cut)
    r_core = torch.clamp(r_core, min=r_min)
    r_cut = torch.where(r_cut > r_core + r_min, r_cut, r_cut + r_min)
    return r_core, r_cut


class DPIS(MassBase):
    """``tf/profiles/mass/piemd.py:21-60``."""

    name = "dPIS"
    params = ["theta_E", "r_core", "r_cut", "center_x", "center_y"]
    _r_min = 0.0001

    def deriv(self, x, y, theta_E, r_core, r_cut, center_x, center_y):
        theta_E, r_core, r_cut = (_t(v, x) for v in (theta_E, r_core, r_cut))
        r_core, r_cut = _sort_ra_rs(r_core, r_cut, self._r_min)
        x, y = x - center_x, y - center_y
        r2 = x ** 2 + y ** 2
        scale = theta_E * r_cut / (r_cut - r_core)
        f_a20 = torch.sqrt(r2 + r_core ** 2) - r_core - torch.sqrt(r2 + r_cut ** 2) + r_cut  # :44-49
        alpha_r = scale / r2 * f_a20
        return alpha_r * x, alpha_r * y

    def hessian(self, x, y, theta_E, r_core, r_cut, center_x, center_y):
        # piemd.py:62-83
        theta_E, r_core, r_cut = (_t(v, x) for v in (theta_E, r_core, r_cut))
        r_core, r_cut = _sort_ra_rs(r_core, r_cut, self._r_min)
        x, y = x - center_x, y - center_y
        r = torch.clamp(torch.sqrt(x ** 2 + y ** 2), min=self._r_min)
        scale = theta_E * r_cut / (r_cut - r_core)
        sc, st = torch.sqrt(r_core ** 2 + r ** 2), torch.sqrt(r_cut ** 2 + r ** 2)
        gamma = scale / 2 * (2 * (1.0 / (r_core + sc) - 1.0 / (r_cut + st)) - (1 / sc - 1 / st))
        kappa = scale / 2 * (r_core + r_cut) / r_cut * (1 / sc - 1 / st)
        sin_imphi = -2 * x * y / r ** 2
        cos_imphi = (y ** 2 - x ** 2) / r ** 2
        gamma1, gamma2 = cos_imphi * gamma, sin_imphi * gamma
        return kappa + gamma1, gamma2, gamma2, kappa - gamma1


class DPIE(MassBase):
    """``tf/profiles/mass/piemd.py:97-119,183-255``."""

    name = "dPIE"
    params = ["theta_E", "r_core", "r_cut", "center_x", "center_y", "e1", "e2"]
    _r_min = 0.0001

    def _param_conv(self, e1, e2):
        phi = torch.atan2(e2, e1) / 2
        e = torch.clamp(torch.sqrt(e1 ** 2 + e2 ** 2), max=0.9999)
        q = (1 - e) / (1 + e)
        return e, q, phi

    def deriv(self, x, y, theta_E, r_core, r_cut, e1, e2, center_x=0, center_y=0):
        theta_E, r_core, r_cut, e1, e2 = (_t(v, x) for v in (theta_E, r_core, r_cut, e1, e2))
        e, q, phi = self._param_conv(e1, e2)
        x, y = x - center_x, y - center_y
        x, y = _rotate(x, y, phi)
        r_core, r_cut = _sort_ra_rs(r_core, r_cut, self._r_min)
        scale = theta_E * r_cut / (r_cut - r_core)
        alpha_x, alpha_y = self.complex_deriv_dual(x, y, r_core, r_cut, e, q)
        alpha_x, alpha_y = _rotate(alpha_x, alpha_y, -phi)
        return scale * alpha_x, scale * alpha_y

    def complex_deriv_dual(self, x, y, r_core, r_cut, e, q):
        # piemd.py:201-255
        sqe = torch.sqrt(e)
        rem2 = x ** 2 / (1.0 + e) ** 2 + y ** 2 / (1.0 - e) ** 2
        zci_re = 0
        zci_im = -0.5 * (1.0 - e ** 2) / sqe
        znum_rc_re = q * x
        znum_rc_im = 2.0 * sqe * torch.sqrt(r_core ** 2 + rem2) - y / q
        zden_rc_re = x
        zden_rc_im = 2.0 * r_core * sqe - y
        znum_rcut_im = 2.0 * sqe * torch.sqrt(r_cut ** 2 + rem2) - y / q
        zden_rcut_im = 2.0 * r_cut * sqe - y
        aa = znum_rc_re * zden_rc_re - znum_rc_im * zden_rcut_im
        bb = znum_rc_re * zden_rcut_im + znum_rc_im * zden_rc_re
        cc = znum_rc_re * zden_rc_re - zden_rc_im * znum_rcut_im
        dd = znum_rc_re * zden_rc_im + zden_rc_re * znum_rcut_im
        norm = cc ** 2 + dd ** 2
        aaa = (aa * cc + bb * dd) / norm
        bbb = (bb * cc - aa * dd) / norm
        norm2 = aaa ** 2 + bbb ** 2
        zr_re = torch.log(torch.sqrt(norm2))
        zr_im = torch.atan2(bbb, aaa)
        zres_re = zci_re * zr_re - zci_im * zr_im
        zres_im = zci_im * zr_re + zci_re * zr_im
        return zres_re, zres_im

    def hessian(self, x, y, theta_E, r_core, r_cut, e1, e2, center_x=0, center_y=0):
        # piemd.py:121-138
        theta_E, r_core, r_cut, e1, e2 = (_t(v, x) for v in (theta_E, r_core, r_cut, e1, e2))
        e, q, phi = self._param_conv(e1, e2)
        x, y = x - center_x, y - center_y
        x, y = _rotate(x, y, phi)
        r_core, r_cut = _sort_ra_rs(r_core, r_cut, self._r_min)
        scale = theta_E * r_cut / (r_cut - r_core)
        xx_c, xy_c, yy_c = self.complex_hessian_single(x, y, r_core, e, q)
        xx_t, xy_t, yy_t = self.complex_hessian_single(x, y, r_cut, e, q)
        f_xx, f_xy, f_yy = scale * (xx_c - xx_t), scale * (xy_c - xy_t), scale * (yy_c - yy_t)
        return self._hessian_rotate(f_xx, f_xy, f_xy, f_yy, -phi)

    @staticmethod
    def _hessian_rotate(f_xx, f_xy, f_yx, f_yy, phi):
        # piemd.py:157-181  (R H R^T)
        cos_2phi, sin_2phi = torch.cos(2 * phi), torch.sin(2 * phi)
        a = 0.5 * (f_xx + f_yy)
        b = 0.5 * (f_xx - f_yy) * cos_2phi
        c = f_xy * sin_2phi
        d = f_xy * cos_2phi
        e = 0.5 * (f_xx - f_yy) * sin_2phi
        return a + b + c, d - e, d - e, a - b - c

    @staticmethod
    def complex_hessian_single(x, y, r_w, e, q):
        # piemd.py:257-300
        sqe = torch.sqrt(e)
        qinv = 1.0 / q
        cxro, cyro = (1.0 + e) * (1.0 + e), (1.0 - e) * (1.0 - e)
        ci = 0.5 * (1.0 - e ** 2) / sqe
        wrem = torch.sqrt(r_w ** 2 + x ** 2 / cxro + y ** 2 / cyro)
        den1 = 2.0 * sqe * wrem - y * qinv
        den1 = q ** 2 * x ** 2 + den1 ** 2
        num2 = 2.0 * r_w * sqe - y
        den2 = x ** 2 + num2 ** 2
        didxre = ci * (q * (2.0 * sqe * x ** 2 / cxro / wrem - 2.0 * sqe * wrem + y * qinv) / den1 + num2 / den2)
        didyre = ci * ((2 * sqe * x * y * q / cyro / wrem - x) / den1 + x / den2)
        didyim = ci * ((2 * sqe * wrem * qinv - y * qinv ** 2 - 4 * e * y / cyro
                        + 2 * sqe * y ** 2 / cyro / wrem * qinv) / den1 - num2 / den2)
        return didxre, didyre, didyim


class TNFW(MassBase):
    """``tf/profiles/mass/tnfw.py:10-62``."""

    name = "TNFW"
    params = ["Rs", "alpha_Rs", "r_trunc", "center_x", "center_y"]

    def deriv(self, x, y, Rs, alpha_Rs, r_trunc, center_x, center_y):
        Rs, alpha_Rs, r_trunc = (_t(v, x) for v in (Rs, alpha_Rs, r_trunc))
        rho0 = alpha_Rs / (4.0 * Rs ** 2 * (1.0 + math.log(0.5)))  # :18
        x, y = x - center_x, y - center_y
        R = torch.sqrt(x ** 2 + y ** 2)
        R = torch.maximum(R, 0.001 * Rs * torch.ones_like(R))  # :21
        X = R / Rs
        tau = r_trunc / Rs
        L = torch.log(X / (tau + torch.sqrt(tau ** 2 + X ** 2)))  # :25
        F = self.F(X)
        gx = (tau ** 2) / (tau ** 2 + 1) ** 2 * (
            (tau ** 2 + 1 + 2 * (X ** 2 - 1)) * F + tau * np.pi + (tau ** 2 - 1) * torch.log(tau)
            + torch.sqrt(tau ** 2 + X ** 2) * (-np.pi + L * (tau ** 2 - 1) / tau))  # :27-36
        a = 4 * rho0 * Rs * gx / X ** 2
        return a * x, a * y

    @staticmethod
    def F(x):
        # :42-62  scatter-update of the x<1 / x>1 entries of a ones tensor
        lt, gt = x < 1, x > 1
        x1 = torch.where(lt, x, torch.full_like(x, 0.5))
        x2 = torch.where(gt, x, torch.full_like(x, 2.0))
        a1 = 1 / torch.sqrt(1 - x1 ** 2) * torch.atanh(torch.sqrt(1 - x1 ** 2))
        a2 = 1 / torch.sqrt(x2 ** 2 - 1) * torch.atan(torch.sqrt(x2 ** 2 - 1))
        a = torch.ones_like(x)
        a = torch.where(lt, a1, a)
        a = torch.where(gt, a2, a)
        return a


class DPIEP(MassBase):
    """``tf/profiles/mass/piep.py:17-56``."""

    name = "dPIE"
    params = ["theta_E", "Ra", "Rs", "center_x", "center_y", "e1", "e2"]

    def __init__(self):
        self.spherical = DPIS()

    def deriv(self, x, y, theta_E, Ra, Rs, e1, e2, center_x=0, center_y=0):
        theta_E, Ra, Rs, e1, e2 = (_t(v, x) for v in (theta_E, Ra, Rs, e1, e2))
        phi = torch.atan2(e2, e1) / 2  # _param_conv :51-56
        c = torch.clamp(torch.sqrt(e1 ** 2 + e2 ** 2), max=0.9999)
        q = (1 - c) / (1 + c)
        e = torch.abs(1 - q ** 2) / (1 + q ** 2)
        x, y = x - center_x, y - center_y
        x, y = _rotate(x, y, phi)
        x, y = x * torch.sqrt(1 - e), y * torch.sqrt(1 + e)
        fx, fy = self.spherical.deriv(x, y, theta_E, Ra, Rs, center_x=0, center_y=0)
        fx = fx * torch.sqrt(1 - e)
        fy = fy * torch.sqrt(1 + e)
        return _rotate(fx, fy, -phi)


class ScalingRelation(MassBase):
    """``tf/profiles/mass/scaling_relation.py:6-70``: member-galaxy sum of a wrapped profile."""

    def __init__(self, profile, scaling_params, lum_star, scaling_params_power, galaxy_catalogue,
                 chunk_size=None, dtype=torch.float32):
        self.profile = profile
        self.name = f"Scaled-{profile.name}"
        self.params = list(scaling_params)
        self.scaling_params = list(scaling_params)
        self.dtype = dtype
        self.power = {k: float(v) for k, v in scaling_params_power.items()}
        lum = torch.as_tensor(np.asarray(galaxy_catalogue["lum"], dtype=np.float32)).to(dtype)
        self.n_galaxy = len(lum)
        self.chunk_size = self.n_galaxy if chunk_size is None else chunk_size
        self.not_scaling_params = [p for p in profile.params if p not in self.scaling_params]
        self._galaxy_constants, self._unscaled_params = [], []
        for pos in range(0, self.n_galaxy, self.chunk_size):  # :44-55
            chunk = slice(pos, pos + self.chunk_size)
            self._galaxy_constants.append(
                {k: torch.as_tensor(np.asarray(galaxy_catalogue[k], dtype=np.float32)[chunk]).to(dtype)
                 for k in self.not_scaling_params})
            # fp32 constants in the reference (:52-54); computed in fp32 then promoted so that the fp64
            # arbiter run sees the same catalogue-derived inputs as the fp32 run.
            # (torch float32 pow, like the tensors of the executed reference: numpy's powf differs from it by an ulp for
            # powers other than 1/2)
            lum32 = torch.as_tensor(np.asarray(galaxy_catalogue["lum"], dtype=np.float32)[chunk])
            self._unscaled_params.append(
                {k: ((lum32 / lum_star) ** torch.tensor(self.power[k], dtype=torch.float32)).to(dtype)
                 for k in self.scaling_params})

    def deriv(self, x, y, **scales):
        alpha_x, alpha_y = torch.zeros_like(x), torch.zeros_like(x)
        x, y = x.unsqueeze(-1), y.unsqueeze(-1)  # (N, bs) -> (N, bs, 1)
        for up, c_chunk in zip(self._unscaled_params, self._galaxy_constants):
            p_chunk = {k: up[k] * _t(scales[k], x).unsqueeze(-1) for k in self.scaling_params}  # :57-59
            ax, ay = self.profile.deriv(x, y, **p_chunk, **c_chunk)
            alpha_x = alpha_x + ax.sum(-1)
            alpha_y = alpha_y + ay.sum(-1)
        return alpha_x, alpha_y

    def hessian(self, x, y, **scales):
        # scaling_relation.py:72-83: member sum of the wrapped profile's hessian
        H = [torch.zeros_like(x) for _ in range(4)]
        x, y = x.unsqueeze(-1), y.unsqueeze(-1)
        for up, c_chunk in zip(self._unscaled_params, self._galaxy_constants):
            p_chunk = {k: up[k] * _t(scales[k], x).unsqueeze(-1) for k in self.scaling_params}
            Hc = self.profile.hessian(x, y, **p_chunk, **c_chunk)
            H = [h + hc.sum(-1) for h, hc in zip(H, Hc)]
        return tuple(H)


class DPIESubhalo(ScalingRelation):
    """``tf/profiles/mass/dpie_subhalo.py:6-21``."""

    def __init__(self, lum_star, galaxy_catalogue, scaling_params_power=None, **kwargs):
        if scaling_params_power is None:
            scaling_params_power = {"theta_E": 0.5, "r_core": 0.5, "r_cut": 0.5}
        super().__init__(profile=DPIE(), scaling_params=["theta_E", "r_core", "r_cut"], lum_star=lum_star,
                         scaling_params_power=scaling_params_power, galaxy_catalogue=galaxy_catalogue, **kwargs)


# -------------------------------------------------------------------------- light


def _sersic_distance(x, y, cx, cy, e1=None, e2=None):
    # sersic.py:37-63
    if e1 is None:
        e1 = torch.zeros_like(_t(cx, x))
    if e2 is None:
        e2 = torch.zeros_like(_t(cx, x))
    phi = torch.atan2(e2, e1) / 2
    c = torch.clamp(torch.sqrt(e1 ** 2 + e2 ** 2), max=0.9999)
    q = (1 - c) / (1 + c)
    dx, dy = x - cx, y - cy
    cos_phi, sin_phi = torch.cos(phi), torch.sin(phi)
    xt1 = (cos_phi * dx + sin_phi * dy) * torch.sqrt(q)
    xt2 = (-sin_phi * dx + cos_phi * dy) / torch.sqrt(q)
    return torch.sqrt(xt1 ** 2 + xt2 ** 2)


class Sersic:
    """``tf/profiles/light/sersic.py:9-35``."""

    name = "SERSIC"
    params = ["R_sersic", "n_sersic", "center_x", "center_y"]
    amp = "Ie"
    depth = 1

    def __init__(self, use_lstsq=False):
        self.use_lstsq = use_lstsq

    def light(self, x, y, R_sersic, n_sersic, center_x, center_y, Ie=None):
        R_sersic, n_sersic = _t(R_sersic, x), _t(n_sersic, x)
        Ie = torch.ones_like(R_sersic) if self.use_lstsq else _t(Ie, x)
        R = _sersic_distance(x, y, _t(center_x, x), _t(center_y, x))
        bn = 1.9992 * n_sersic - 0.3271
        ret = Ie * torch.exp(-bn * ((R / R_sersic) ** (1 / n_sersic) - 1.0))
        return ret[None, ...] if self.use_lstsq else ret


class SersicEllipse(Sersic):
    """``tf/profiles/light/sersic.py:66-80``."""

    name = "SERSIC_ELLIPSE"
    params = ["R_sersic", "n_sersic", "e1", "e2", "center_x", "center_y"]

    def light(self, x, y, R_sersic, n_sersic, e1, e2, center_x, center_y, Ie=None):
        R_sersic, n_sersic, e1, e2 = (_t(v, x) for v in (R_sersic, n_sersic, e1, e2))
        Ie = torch.ones_like(R_sersic) if self.use_lstsq else _t(Ie, x)
        R = _sersic_distance(x, y, _t(center_x, x), _t(center_y, x), e1, e2)
        bn = 1.9992 * n_sersic - 0.3271
        ret = Ie * torch.exp(-bn * ((R / R_sersic) ** (1 / n_sersic) - 1.0))
        return ret[None, ...] if self.use_lstsq else ret


class CoreSersic(Sersic):
    """``tf/profiles/light/sersic.py:83-132``, the expression exactly as written there (``R_sersic ** alpha ** 1.0`` is
    ``R_sersic^alpha``; the ``1/(alpha n)`` divides and the ``- 1.0`` sits outside the ``bn`` product -- SURVEY App. B7)."""

    name = "CORE_SERSIC"
    params = ["R_sersic", "n_sersic", "Rb", "alpha", "gamma", "e1", "e2", "center_x", "center_y"]

    def light(self, x, y, R_sersic, n_sersic, Rb, alpha, gamma, e1, e2, center_x, center_y, Ie=None):
        R_sersic, n_sersic, Rb, alpha, gamma = (_t(v, x) for v in (R_sersic, n_sersic, Rb, alpha, gamma))
        Ie = torch.ones_like(R_sersic) if self.use_lstsq else _t(Ie, x)
        R = _sersic_distance(x, y, _t(center_x, x), _t(center_y, x), _t(e1, x), _t(e2, x))
        bn = 1.9992 * n_sersic - 0.3271
        ret = (Ie * (1 + (Rb / R) ** alpha) ** (gamma / alpha)
               * torch.exp(-bn * ((R ** alpha + Rb ** alpha) / R_sersic ** alpha ** 1.0 / (alpha * n_sersic)) - 1.0))
        return ret[None, ...] if self.use_lstsq else ret


def shapelet_phi_n_np(n, x):
    """lenstronomy ``Shapelets.phi_n`` (restated from its published definition; the source is
    not available offline -> parity unpinned): H_n(x) exp(-x^2/2) / sqrt(2^n sqrt(pi) n!)."""
    coef = np.zeros(n + 1)
    coef[n] = 1.0
    prefactor = 1.0 / np.sqrt(2.0 ** n * np.sqrt(np.pi) * math.factorial(n))
    return prefactor * np.polynomial.hermite.hermval(x, coef) * np.exp(-np.asarray(x) ** 2 / 2.0)


def interp_regular_1d_grid(x, x_min, x_max, y_ref, fill_below=0.0, fill_above=0.0):
    """``tfp.math.interp_regular_1d_grid`` along axis -1 of ``y_ref`` (restated; unpinned).

    y_ref: (C, n) table; x: any shape.  Returns (C, *x.shape): linear interpolation between the
    two bracketing table points, ``fill`` outside [x_min, x_max]."""
    n = y_ref.shape[-1]
    x_idx_unclipped = (x - x_min) / (x_max - x_min) * (n - 1)
    x_idx = torch.clamp(x_idx_unclipped, 0, n - 1)
    idx_below = torch.floor(x_idx)
    idx_above = torch.clamp(idx_below + 1, max=n - 1)
    idx_below = torch.clamp(idx_above - 1, min=0)
    ib = idx_below.long().reshape(-1)
    ia = idx_above.long().reshape(-1)
    y_below = y_ref[:, ib].reshape(y_ref.shape[0], *x.shape)
    y_above = y_ref[:, ia].reshape(y_ref.shape[0], *x.shape)
    t = x_idx - idx_below
    y = t * y_above + (1 - t) * y_below
    y = torch.where(x_idx_unclipped < 0, torch.as_tensor(fill_below, dtype=y.dtype), y)
    y = torch.where(x_idx_unclipped > n - 1, torch.as_tensor(fill_above, dtype=y.dtype), y)
    return y


class Shapelets:
    """``tf/profiles/light/shapelets.py:11-85``."""

    name = "SHAPELETS"
    params = ["beta", "center_x", "center_y"]

    def __init__(self, n_max, use_lstsq=False, interpolate=True, dtype=torch.float32):
        self.use_lstsq = use_lstsq
        self.n_max = n_max
        self.n_layers = int((n_max + 1) * (n_max + 2) / 2)
        self.interpolate = interpolate
        self.dtype = dtype
        n1, n2 = 0, 0
        self.N1, self.N2, self.amp_names = [], [], []
        decimal_places = len(str(self.n_layers))
        grid = np.linspace(-5, 5, 6000)
        herm_X, herm_Y = [], []
        for i in range(self.n_layers):  # shapelets.py:34-46
            self.amp_names.append(f"amp{str(i).zfill(decimal_places)}")
            self.N1.append(n1)
            self.N2.append(n2)
            herm_X.append(shapelet_phi_n_np(n1, grid))
            herm_Y.append(shapelet_phi_n_np(n2, grid))
            if n1 == 0:
                n1 = n2 + 1
                n2 = 0
            else:
                n1 -= 1
                n2 += 1
        N = np.arange(0, n_max + 1, dtype=np.float64)
        # :47-48 evaluated in fp32 in the reference
        pref32 = 1.0 / np.sqrt((2.0 ** N).astype(np.float32) * np.float32(np.sqrt(np.float32(np.pi)))
                               * np.exp(np.array([math.lgamma(v + 1) for v in N], dtype=np.float32)))
        self.prefactor = torch.as_tensor(pref32.astype(np.float32)).to(dtype)
        self.depth = self.n_layers
        # tables are cast to fp32 in the reference (:50-51)
        self.herm_X = torch.as_tensor(np.asarray(herm_X, dtype=np.float32)).to(dtype)
        self.herm_Y = torch.as_tensor(np.asarray(herm_Y, dtype=np.float32)).to(dtype)

    def phi_n(self, x):
        # :77-85
        polys = [torch.ones_like(x), 2 * x]
        i = 2.0
        while i < self.n_max + 1:
            polys.append(2 * (x * polys[-1] - (i - 1) * polys[-2]))
            i += 1.0
        polys = torch.stack(polys[: self.n_max + 1], 0)
        return polys * self.prefactor.reshape(-1, *([1] * x.dim()))

    def light(self, x, y, center_x, center_y, beta, **amp):
        beta = _t(beta, x)
        x = (x - center_x) / beta
        y = (y - center_y) / beta
        if self.interpolate:
            ret = interp_regular_1d_grid(x, -5.0, 5.0, self.herm_X) * interp_regular_1d_grid(y, -5.0, 5.0, self.herm_Y)
            if self.use_lstsq:
                return ret
            # tf.nest.flatten(amp): dict -> values in sorted key order
            a = torch.stack([_t(amp[k], x) * torch.ones_like(beta) for k in sorted(amp)], 0)  # (D, bs)
            return (ret * a.reshape(a.shape[0], *([1] * (x.dim() - 1)), -1)).sum(0)
        XX, YY = self.phi_n(x), self.phi_n(y)
        fac = torch.exp(-(x ** 2 + y ** 2) / 2)
        comp = XX[self.N1] * YY[self.N2]
        if self.use_lstsq:
            return fac * comp
        a = torch.stack([_t(amp[k], x) * torch.ones_like(beta) for k in sorted(amp)], 0)
        return fac * (comp * a.reshape(a.shape[0], *([1] * (x.dim() - 1)), -1)).sum(0)
