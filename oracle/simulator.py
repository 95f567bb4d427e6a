"""Oracle restatement of ``LensWCS`` and the TF ``LensSimulator`` (TEST INFRASTRUCTURE).

Follows ``src/gigalens/simulator.py:32-64`` and ``src/gigalens/tf/simulator.py:13-240``.
``lstsq_simulate`` takes its tensor layout from ``src/gigalens/jax/simulator.py:144-195``
because the TF version cannot run as written; semantics are §3.4.
"""
import copy

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------- lenstronomy bits
# Restated from the published behaviour of lenstronomy 1.9.x ``Util.kernel_util``; the
# source is not available offline, so these are PARITY UNPINNED against lenstronomy.  The
# CUDA path and the oracle consume the same restated kernel, so hot-path parity does not
# depend on matching lenstronomy bit for bit.


def _re_size_array(x_in, y_in, values, x_out, y_out):
    """Bilinear re-sampling (degree-1 spline, clamped outside the input range)."""
    from scipy.interpolate import RectBivariateSpline

    spl = RectBivariateSpline(x_in, y_in, values, kx=1, ky=1, s=0)
    return spl(x_out, y_out)


def _kernel_norm(k):
    return k / np.sum(k)


def _averaging_even_kernel(kernel_high_res, subgrid_res):
    n_high_in = len(kernel_high_res)
    n_low = int(round(n_high_in / subgrid_res + 0.5))
    if n_low % 2 == 0:
        n_low += 1
    n_high = int(n_low * subgrid_res - 1)
    if n_high == n_high_in:
        edges = kernel_high_res
    else:
        i_start = int((n_high - n_high_in) / 2)
        edges = np.zeros((n_high, n_high))
        edges[i_start:-i_start, i_start:-i_start] = kernel_high_res
    low = np.zeros((n_low, n_low))
    for i in range(subgrid_res - 1):  # sub-pixels wholly inside one pixel
        for j in range(subgrid_res - 1):
            low += edges[i::subgrid_res, j::subgrid_res]
    i = subgrid_res - 1  # sub-pixels straddling a horizontal edge: half each
    for j in range(subgrid_res - 1):
        low[1:, :] += edges[i::subgrid_res, j::subgrid_res] / 2
        low[:-1, :] += edges[i::subgrid_res, j::subgrid_res] / 2
    j = subgrid_res - 1
    for i in range(subgrid_res - 1):
        low[:, 1:] += edges[i::subgrid_res, j::subgrid_res] / 2
        low[:, :-1] += edges[i::subgrid_res, j::subgrid_res] / 2
    edge = edges[subgrid_res - 1::subgrid_res, subgrid_res - 1::subgrid_res]  # corners: a quarter each
    low[1:, 1:] += edge / 4
    low[:-1, 1:] += edge / 4
    low[1:, :-1] += edge / 4
    low[:-1, :-1] += edge / 4
    return low


def _averaging(grid, num_grid, num_pix):
    fac = int(num_grid / num_pix)
    return grid.reshape(num_pix, fac, num_pix, fac).mean(-1).mean(1)


def subgrid_kernel(kernel, subgrid_res, odd=False, num_iter=100):
    """lenstronomy ``subgrid_kernel`` (call site ``tf/simulator.py:63-65``)."""
    subgrid_res = int(subgrid_res)
    kernel = np.asarray(kernel, dtype=np.float64)
    if subgrid_res == 1:
        return kernel
    nx, ny = kernel.shape
    d_x = 1.0 / nx
    x_in = np.linspace(d_x / 2, 1 - d_x / 2, nx)
    d_y = 1.0 / nx
    y_in = np.linspace(d_y / 2, 1 - d_y / 2, ny)
    nx_new, ny_new = nx * subgrid_res, ny * subgrid_res
    if odd:
        if nx_new % 2 == 0:
            nx_new -= 1
        if ny_new % 2 == 0:
            ny_new -= 1
    d_x_new, d_y_new = 1.0 / nx_new, 1.0 / ny_new
    x_out = np.linspace(d_x_new / 2.0, 1 - d_x_new / 2.0, nx_new)
    y_out = np.linspace(d_y_new / 2.0, 1 - d_y_new / 2.0, ny_new)
    kernel_input = copy.deepcopy(kernel)
    kernel_subgrid = _kernel_norm(_re_size_array(x_in, y_in, kernel_input, x_out, y_out))
    for _ in range(max(num_iter, 1)):
        if subgrid_res % 2 == 0:
            kernel_pixel = _averaging_even_kernel(kernel_subgrid, subgrid_res)
        else:
            kernel_pixel = _averaging(kernel_subgrid, nx_new, nx)
        delta = kernel - kernel_pixel
        temp_kernel = kernel_input + delta
        kernel_subgrid = _kernel_norm(_re_size_array(x_in, y_in, temp_kernel, x_out, y_out))
        kernel_input = temp_kernel
    if subgrid_res % 2 == 0:
        return kernel_subgrid
    kernel_pixel = _kernel_norm(_averaging(kernel_subgrid, nx_new, nx))
    delta_kernel = kernel_pixel - _kernel_norm(kernel)
    delta_kernel_sub = np.kron(delta_kernel, np.ones((subgrid_res, subgrid_res))) / subgrid_res ** 2
    return _kernel_norm(kernel_subgrid - delta_kernel_sub)


# ----------------------------------------------------------------------- LensWCS


class LensWCS:
    """``src/gigalens/simulator.py:32-64`` (numpy, as in the reference)."""

    def __init__(self, n, supersample=1, transform_pix2angle=None, pix_scale=1.0):
        if transform_pix2angle is None:
            transform_pix2angle = np.eye(2) * pix_scale
        transform_pix2angle = np.asarray(transform_pix2angle, dtype=np.float64)
        self.transform_pix2angle = transform_pix2angle / supersample
        self.transform_angle2pix = np.linalg.inv(transform_pix2angle)
        if isinstance(n, int):
            self.n_x, self.n_y = n, n
        else:
            self.n_x, self.n_y = n
        self.supersample = supersample
        low_x = -(self.n_x * self.supersample - 1) / 2
        low_y = -(self.n_y * self.supersample - 1) / 2
        self.radec_at_xy_0 = np.squeeze(self.transform_pix2angle @ ([[low_x], [low_y]]))

    def pix2angle(self, x, y):
        radec = np.einsum("ij,i...->...j", self.transform_pix2angle, np.concatenate([[x], [y]])) + self.radec_at_xy_0
        radec = np.swapaxes(radec, -1, 0).astype(np.float32)
        return radec[0].T, radec[1].T


# ------------------------------------------------------------------ LensSimulator


class OracleSimulator:
    """``tf/simulator.py:13-240`` on torch CPU.  ``dtype`` float32 (reference arithmetic) or
    float64 (arbiter).  ``phys_model`` is any object with ``lenses``, ``lens_light``,
    ``source_light`` lists of oracle profiles and matching ``*_constants`` lists of dicts."""

    def __init__(self, phys_model, delta_pix, num_pix, supersample=1, kernel=None, transform_pix2angle=None,
                 pix_region=None, bs=1, dtype=torch.float32):
        self.phys_model = phys_model
        self.dtype = dtype
        self.bs = bs
        self.supersample = int(supersample)
        self.num_pix = num_pix
        self.wcs = LensWCS(n=num_pix, supersample=supersample, transform_pix2angle=transform_pix2angle,
                           pix_scale=delta_pix)
        T = np.eye(2) * delta_pix if transform_pix2angle is None else np.asarray(transform_pix2angle)
        # :27-29  det of the un-supersampled transform, cast to fp32 (the default transform is a float32 tensor, tf.eye(2) * delta_pix;
        # a user-supplied array keeps its own dtype through tf.linalg.det)
        self.conversion_factor = float(np.float32(np.linalg.det(T.astype(np.float32) if transform_pix2angle is None else T)))
        nss = num_pix * self.supersample
        if pix_region is None:  # :34-42
            region = np.ones((nss, nss), dtype=bool)
            img_region = np.ones((num_pix, num_pix))
        else:
            img_region = np.asarray(pix_region)
            region = np.repeat(np.repeat(img_region.astype(bool), self.supersample, axis=0), self.supersample, axis=1)
        self.region = np.argwhere(region)  # row-major (row, col) list  :43
        self.img_region = torch.as_tensor(img_region.astype(np.float32)).to(dtype)
        img_X, img_Y = self.wcs.pix2angle(self.region[:, 1], self.region[:, 0])  # x <- column  :45
        self.img_X = torch.as_tensor(img_X).to(dtype)[:, None].repeat(1, bs)  # (N, bs)  :46-51
        self.img_Y = torch.as_tensor(img_Y).to(dtype)[:, None].repeat(1, bs)
        self.flat_kernel = None
        self.kernel_np = None
        if kernel is not None:  # :62-70
            k = subgrid_kernel(np.asarray(kernel), self.supersample, odd=True)[::-1, ::-1]
            self.kernel_np = np.ascontiguousarray(k).astype(np.float32)
            self.flat_kernel = torch.as_tensor(self.kernel_np).to(dtype)

    # -- helpers
    def _split(self, params):
        pm = self.phys_model
        lens = params["lens_mass"] if "lens_mass" in params else [{} for _ in pm.lenses]
        ll = params["lens_light"] if "lens_light" in params else [{} for _ in pm.lens_light]
        sl = params["source_light"] if "source_light" in params else [{} for _ in pm.source_light]
        return lens, ll, sl

    def _c(self, d):
        return {k: (v.to(self.dtype) if torch.is_tensor(v) else torch.as_tensor(np.asarray(v, dtype=np.float32)).to(self.dtype))
                for k, v in d.items()}

    def beta(self, x, y, lens_params):
        # :72-78  every deflector is evaluated at the image-plane position (single plane)
        beta_x, beta_y = x, y
        for lens, p, c in zip(self.phys_model.lenses, lens_params, self.phys_model.lenses_constants):
            f_xi, f_yi = lens.deriv(x, y, **p, **self._c(c))
            beta_x, beta_y = beta_x - f_xi, beta_y - f_yi
        return beta_x, beta_y

    def hessian(self, x, y, lens_params):
        # the sum magnification / convergence / shear take over the lens list (tf/simulator.py:80-107)
        H = [torch.zeros_like(x) for _ in range(4)]
        for lens, p, c in zip(self.phys_model.lenses, lens_params, self.phys_model.lenses_constants):
            Hi = lens.hessian(x, y, **p, **self._c(c))
            H = [h + hi for h, hi in zip(H, Hi)]
        return tuple(H)

    def magnification(self, x, y, lens_params):
        # :80-91
        f_xx, f_xy, f_yx, f_yy = self.hessian(x, y, lens_params)
        det_A = (1 - f_xx) * (1 - f_yy) - f_xy * f_yx
        return 1.0 / det_A

    def convergence(self, x, y, lens_params):
        # :93-98
        kappa = torch.zeros_like(x)
        for lens, p, c in zip(self.phys_model.lenses, lens_params, self.phys_model.lenses_constants):
            kappa = kappa + lens.convergence(x, y, **p, **self._c(c))
        return kappa

    def shear(self, x, y, lens_params):
        # :100-107
        g1, g2 = torch.zeros_like(x), torch.zeros_like(x)
        for lens, p, c in zip(self.phys_model.lenses, lens_params, self.phys_model.lenses_constants):
            a, b = lens.shear(x, y, **p, **self._c(c))
            g1, g2 = g1 + a, g2 + b
        return g1, g2

    def _scatter(self, vals):
        nss = self.num_pix * self.supersample
        img = torch.zeros((nss * nss,) + tuple(vals.shape[1:]), dtype=self.dtype)
        flat = torch.as_tensor(self.region[:, 0] * nss + self.region[:, 1])
        img = img.index_add(0, flat, vals)
        return img.reshape((nss, nss) + tuple(vals.shape[1:]))

    def _conv_pool(self, img):
        """img: (B, C, H, W).  conv2d SAME with the flipped kernel (cross-correlation), then
        ss x ss mean pooling  (:142-155, :214-227)."""
        if self.flat_kernel is not None:
            K = self.flat_kernel.shape[0]
            B, C, H, W = img.shape
            img = F.conv2d(img.reshape(B * C, 1, H, W), self.flat_kernel[None, None], padding=K // 2).reshape(B, C, H, W)
        if self.supersample != 1:
            img = F.avg_pool2d(img, self.supersample)
        return img

    def simulate_ss(self, params, no_deflection=False):
        """The supersampled, pre-convolution image (H_ss, W_ss, bs) after the NaN scrub (:124-140)."""
        lens_params, ll_params, sl_params = self._split(params)
        pm = self.phys_model
        beta_x, beta_y = self.beta(self.img_X, self.img_Y, lens_params)
        if no_deflection:
            beta_x, beta_y = self.img_X, self.img_Y
        nss = self.num_pix * self.supersample
        img = torch.zeros((nss, nss, self.bs), dtype=self.dtype)
        for lm, p, c in zip(pm.lens_light, ll_params, pm.lens_light_constants):
            img = img + self._scatter(lm.light(self.img_X, self.img_Y, **p, **self._c(c)))
        for lm, p, c in zip(pm.source_light, sl_params, pm.source_light_constants):
            img = img + self._scatter(lm.light(beta_x, beta_y, **p, **self._c(c)))
        return torch.where(torch.isnan(img), torch.zeros_like(img), img)

    def simulate(self, params, no_deflection=False):
        img = self.simulate_ss(params, no_deflection)
        img = img.permute(2, 0, 1)  # (bs, H, W)  :141
        ret = self._conv_pool(img[:, None])[:, 0]
        return torch.squeeze(ret) * self.conversion_factor  # :156

    def _finish(self, img):
        """NaN scrub, (bs, H, W) transpose, conv + pool, squeeze, conversion factor: the tail every
        ``simulate_*`` variant of the reference repeats verbatim (e.g. :251-266)."""
        img = torch.where(torch.isnan(img), torch.zeros_like(img), img)
        ret = self._conv_pool(img.permute(2, 0, 1)[:, None])[:, 0]
        return torch.squeeze(ret) * self.conversion_factor

    def simulate_source(self, params):
        """tf/simulator.py:242-266: the source light evaluated on the IMAGE-plane grid (no ray-shooting)."""
        pm = self.phys_model
        nss = self.num_pix * self.supersample
        img = torch.zeros((nss, nss, self.bs), dtype=self.dtype)
        for lm, p, c in zip(pm.source_light, params["source_light"], pm.source_light_constants):
            img = img + self._scatter(lm.light(self.img_X, self.img_Y, **p, **self._c(c)))
        return self._finish(img)

    def simulate_lens_light(self, params):
        """tf/simulator.py:268-293."""
        pm = self.phys_model
        nss = self.num_pix * self.supersample
        img = torch.zeros((nss, nss, self.bs), dtype=self.dtype)
        for lm, p, c in zip(pm.lens_light, params["lens_light"], pm.lens_light_constants):
            img = img + self._scatter(lm.light(self.img_X, self.img_Y, **p, **self._c(c)))
        return self._finish(img)

    def simulate_images(self, params):
        """tf/simulator.py:295-328: the lensed source only."""
        pm = self.phys_model
        beta_x, beta_y = self.beta(self.img_X, self.img_Y, params["lens_mass"])
        nss = self.num_pix * self.supersample
        img = torch.zeros((nss, nss, self.bs), dtype=self.dtype)
        for lm, p, c in zip(pm.source_light, params["source_light"], pm.source_light_constants):
            img = img + self._scatter(lm.light(beta_x, beta_y, **p, **self._c(c)))
        return self._finish(img)

    def lstsq_stack(self, params, no_deflection=False):
        """(bs, ny, nx, D): every linear light component, unit amplitude, convolved and pooled."""
        lens_params, ll_params, sl_params = self._split(params)
        pm = self.phys_model
        beta_x, beta_y = self.beta(self.img_X, self.img_Y, lens_params)
        if no_deflection:
            beta_x, beta_y = self.img_X, self.img_Y
        comps = []
        for lm, p, c in zip(pm.lens_light, ll_params, pm.lens_light_constants):
            comps.append(self._scatter(lm.light(self.img_X, self.img_Y, **p, **self._c(c)).permute(1, 2, 0)))
        for lm, p, c in zip(pm.source_light, sl_params, pm.source_light_constants):
            comps.append(self._scatter(lm.light(beta_x, beta_y, **p, **self._c(c)).permute(1, 2, 0)))
        img = torch.cat(comps, dim=-1)  # (H, W, bs, D)
        img = torch.where(torch.isnan(img), torch.zeros_like(img), img)
        img = img.permute(2, 3, 0, 1)  # (bs, D, H, W)
        ret = self._conv_pool(img).permute(0, 2, 3, 1)  # (bs, ny, nx, D)
        return torch.where(torch.isnan(ret), torch.zeros_like(ret), ret)

    def lstsq_simulate(self, params, observed_image, err_map, return_stacked=False, return_coeffs=False,
                       no_deflection=False):
        ret = self.lstsq_stack(params, no_deflection)
        if return_stacked:
            return ret
        depth = ret.shape[-1]
        observed_image = torch.as_tensor(observed_image).to(self.dtype)
        err_map = torch.as_tensor(err_map).to(self.dtype)
        W = (1 / err_map)[..., None]  # :232
        Y = (observed_image * W[..., 0]).reshape(1, -1, 1)
        X = (ret * W).reshape(self.bs, -1, depth)
        Xt = X.transpose(1, 2)
        # tf.linalg.pinv(a, rcond): SVD with singular values <= rcond * max(s) dropped
        coeffs = (torch.linalg.pinv(Xt @ X, rtol=1e-6) @ Xt @ Y)[..., 0]  # :235-236
        if return_coeffs:
            return coeffs
        out = (ret * coeffs[:, None, None, :]).sum(-1)  # :239
        return torch.squeeze(out)
