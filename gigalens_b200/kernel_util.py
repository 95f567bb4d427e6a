"""Host-side PSF preparation: ``subgrid_kernel`` as used by ``LensSimulator.__init__``
(reference call site ``src/gigalens/tf/simulator.py:62-70``; the function itself lives in
lenstronomy 1.9.x ``Util.kernel_util`` which is not available offline, so this is a restatement
of its published algorithm -- bilinear up-sampling followed by 100 fixed-point iterations that
re-bin the proposal to pixel scale and feed the residual back).  One-time setup, numpy.
"""
import numpy as np


def _bilinear_resize(x_in, y_in, values, x_out, y_out):
    """Degree-1 tensor-product interpolation, clamped outside the input range."""

    def weights(src, dst):
        idx = np.clip(np.searchsorted(src, dst) - 1, 0, len(src) - 2)
        t = np.clip((dst - src[idx]) / (src[idx + 1] - src[idx]), 0.0, 1.0)
        return idx, t

    ix, tx = weights(np.asarray(x_in), np.asarray(x_out))
    iy, ty = weights(np.asarray(y_in), np.asarray(y_out))
    v00 = values[np.ix_(ix, iy)]
    v10 = values[np.ix_(ix + 1, iy)]
    v01 = values[np.ix_(ix, iy + 1)]
    v11 = values[np.ix_(ix + 1, iy + 1)]
    tx = tx[:, None]
    ty = ty[None, :]
    return (1 - tx) * (1 - ty) * v00 + tx * (1 - ty) * v10 + (1 - tx) * ty * v01 + tx * ty * v11


def _rebin_even(high, res):
    """Re-bin an odd-sized, even-factor supersampled kernel: sub-pixels inside a pixel count fully,
    those straddling an edge half to each side, corners a quarter to each of four pixels."""
    n_high_in = len(high)
    n_low = int(round(n_high_in / res + 0.5))
    if n_low % 2 == 0:
        n_low += 1
    n_high = n_low * res - 1
    if n_high != n_high_in:
        if (n_high - n_high_in) % 2:
            raise ValueError("even-sized supersampled kernels are not supported (use odd=True)")
        i0 = (n_high - n_high_in) // 2
        full = np.zeros((n_high, n_high))
        full[i0:n_high - i0, i0:n_high - i0] = high
        high = full
    low = np.zeros((n_low, n_low))
    last = res - 1
    for i in range(last):
        for j in range(last):
            low += high[i::res, j::res]
    for j in range(last):
        edge = high[last::res, j::res] / 2
        low[1:, :] += edge
        low[:-1, :] += edge
    for i in range(last):
        edge = high[i::res, last::res] / 2
        low[:, 1:] += edge
        low[:, :-1] += edge
    corner = high[last::res, last::res] / 4
    low[1:, 1:] += corner
    low[:-1, 1:] += corner
    low[1:, :-1] += corner
    low[:-1, :-1] += corner
    return low


def _rebin_odd(high, n_high, n_low):
    f = n_high // n_low
    return high.reshape(n_low, f, n_low, f).mean(axis=(1, 3))


def subgrid_kernel(kernel, subgrid_res, odd=False, num_iter=100):
    subgrid_res = int(subgrid_res)
    kernel = np.asarray(kernel, dtype=np.float64)
    if subgrid_res == 1:
        return kernel
    nx, ny = kernel.shape
    x_in = np.linspace(0.5 / nx, 1 - 0.5 / nx, nx)
    y_in = np.linspace(0.5 / nx, 1 - 0.5 / nx, ny)
    nx_new, ny_new = nx * subgrid_res, ny * subgrid_res
    if odd:
        nx_new -= 1 - nx_new % 2
        ny_new -= 1 - ny_new % 2
    x_out = np.linspace(0.5 / nx_new, 1 - 0.5 / nx_new, nx_new)
    y_out = np.linspace(0.5 / ny_new, 1 - 0.5 / ny_new, ny_new)
    proposal_in = kernel.copy()
    sub = _bilinear_resize(x_in, y_in, proposal_in, x_out, y_out)
    sub /= sub.sum()
    for _ in range(max(num_iter, 1)):
        low = _rebin_even(sub, subgrid_res) if subgrid_res % 2 == 0 else _rebin_odd(sub, nx_new, nx)
        proposal_in = proposal_in + (kernel - low)
        sub = _bilinear_resize(x_in, y_in, proposal_in, x_out, y_out)
        sub /= sub.sum()
    if subgrid_res % 2 == 0:
        return sub
    low = _rebin_odd(sub, nx_new, nx)
    low /= low.sum()
    delta = low - kernel / kernel.sum()
    sub = sub - np.kron(delta, np.ones((subgrid_res, subgrid_res))) / subgrid_res ** 2
    return sub / sub.sum()
