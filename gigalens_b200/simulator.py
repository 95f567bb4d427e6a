"""``SimulatorConfig`` / ``LensWCS`` / ``LensSimulator`` with the reference's API surface
(``src/gigalens/simulator.py:11-127``, ``src/gigalens/tf/simulator.py:13-240``), running on the
CUDA library.  Everything here is one-time host setup or thin argument marshalling; the data
path is ``libgigalens_b200.so``.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import numpy as np

from . import _cabi
from .kernel_util import subgrid_kernel


@dataclass
class SimulatorConfig:
    """Holds parameters for simulation (``src/gigalens/simulator.py:11-29``)."""

    delta_pix: float
    num_pix: int
    supersample: Optional[int] = 1
    kernel: Optional[Any] = None
    transform_pix2angle: Optional[np.ndarray] = None
    pix_region: Optional[np.ndarray] = None


class LensWCS:
    """Pixel <-> angle grid (``src/gigalens/simulator.py:32-64``), including the reference's
    conventions: the supersampled transform is ``T / ss``, the origin puts the grid centre at
    (0, 0), and ``pix2angle`` contracts with the first index of ``T`` (the transpose of what the
    origin uses -- identical for the diagonal transforms every config uses)."""

    def __init__(self, n, supersample=1, transform_pix2angle=None, pix_scale=1.0):
        if transform_pix2angle is None:
            transform_pix2angle = np.eye(2) * pix_scale
        transform_pix2angle = np.asarray(transform_pix2angle, dtype=np.float64)
        self.transform_pix2angle = transform_pix2angle / supersample
        self.transform_angle2pix = np.linalg.inv(transform_pix2angle)
        if isinstance(n, (int, np.integer)):
            self.n_x, self.n_y = int(n), int(n)
        else:
            self.n_x, self.n_y = n
        self.supersample = supersample
        low_x = -(self.n_x * self.supersample - 1) / 2
        low_y = -(self.n_y * self.supersample - 1) / 2
        self.radec_at_xy_0 = np.squeeze(self.transform_pix2angle @ ([[low_x], [low_y]]))

    def pix2angle(self, x, y):
        xy = np.stack([np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)], 0)
        T = self.transform_pix2angle
        ra = T[0, 0] * xy[0] + T[1, 0] * xy[1] + self.radec_at_xy_0[0]
        dec = T[0, 1] * xy[0] + T[1, 1] * xy[1] + self.radec_at_xy_0[1]
        return ra.astype(np.float32), dec.astype(np.float32)

    def angle2pix(self, ra, dec):
        d0 = np.asarray(ra, dtype=np.float64) - self.radec_at_xy_0[0]
        d1 = np.asarray(dec, dtype=np.float64) - self.radec_at_xy_0[1]
        Ti = self.transform_angle2pix
        return np.stack([Ti[0, 0] * d0 + Ti[0, 1] * d1, Ti[1, 0] * d0 + Ti[1, 1] * d1], 0).astype(np.float32)

    def pixel_grid(self):
        x, y = np.arange(self.n_y * self.supersample), np.arange(self.n_y * self.supersample)
        X, Y = np.meshgrid(x, y)
        return self.pix2angle(X, Y)


GROUPS = ("lens_mass", "lens_light", "source_light")


class CompiledModel:
    """A ``PhysicalModel`` flattened to the C ABI's ``gl_model_desc``: slot order is group by
    group, profile by profile, the profile's own parameter order, skipping names fixed in the
    ``*_constants`` dicts (reference ``tf/simulator.py:75-76,129-138`` merges ``**p, **c``)."""

    def __init__(self, phys_model):
        self.phys_model = phys_model
        groups = [
            (phys_model.lenses, phys_model.lenses_constants),
            (phys_model.lens_light, phys_model.lens_light_constants),
            (phys_model.source_light, phys_model.source_light_constants),
        ]
        self.slots: Dict[tuple, int] = {}
        self.slot_keys: List[tuple] = []
        self._keep = []  # numpy buffers referenced by the descriptors
        self.descs = []
        self.depth = 0
        for gi, (profiles, constants) in enumerate(groups):
            arr = (_cabi.ProfileDesc * max(1, len(profiles)))()
            for pi, (prof, const) in enumerate(zip(profiles, constants)):
                self._fill(arr[pi], prof, const, GROUPS[gi], pi)
                if gi > 0:
                    self.depth += getattr(prof, "depth", 1)
            self.descs.append(arr)
        self.n_params = len(self.slot_keys)
        m = _cabi.ModelDesc()
        m.n_lens, m.n_lens_light, m.n_source_light = (len(g[0]) for g in groups)
        m.lens, m.lens_light, m.source_light = self.descs
        m.n_params = self.n_params
        self.desc = m

    def _slot(self, group, pi, name):
        key = (group, pi, name)
        if key not in self.slots:
            self.slots[key] = len(self.slot_keys)
            self.slot_keys.append(key)
        return self.slots[key]

    def _fill(self, d, prof, const, group, pi):
        inner = getattr(prof, "profile", None)  # ScalingRelation wraps another profile
        type_id = prof._type_id
        raw = _cabi.RAW_ORDER[type_id]
        d.type = type_id
        d.flags = 0
        if getattr(prof, "use_lstsq", False):
            d.flags |= _cabi.GL_FLAG_USE_LSTSQ
        if getattr(prof, "interpolate", False):
            d.flags |= _cabi.GL_FLAG_INTERPOLATE
        d.niter = int(getattr(inner or prof, "niter", 0))
        d.n_max = int(getattr(prof, "n_max", 0))
        d.n_members = 0
        free = set(prof.params)
        for k, name in enumerate(raw):
            d.slot[k] = -1
            d.constant[k] = 0.0
            if name in const:
                d.constant[k] = float(np.asarray(const[name]))
            elif inner is not None:
                if name in prof.scaling_params:
                    d.slot[k] = self._slot(group, pi, name)
                else:
                    d.constant[k] = 1.0  # catalogue column lives in member_factor
            elif name in free:
                d.slot[k] = self._slot(group, pi, name)
            elif name == getattr(prof, "_amp", None) and getattr(prof, "use_lstsq", False):
                d.constant[k] = 1.0
            elif type_id in (_cabi.GL_DPIE, _cabi.GL_DPIEP) and name in ("center_x", "center_y"):
                d.constant[k] = 0.0  # DPIE.deriv / DPIEP.deriv defaults (piemd.py:106, piep.py:33)
            else:
                raise KeyError(f"{group}[{pi}] ({prof.name}): parameter '{name}' is neither free nor constant")
        if inner is not None:
            mf = np.ascontiguousarray(prof.member_factors(raw), dtype=np.float32)
            self._keep.append(mf)
            d.n_members = prof.n_galaxy
            d.member_factor = mf.ctypes.data_as(C.POINTER(C.c_float))
        if type_id == _cabi.GL_SHAPELETS and not prof.use_lstsq:
            amp = np.asarray([self._slot(group, pi, n) for n in sorted(prof._amp_names)], dtype=np.int32)
            self._keep.append(amp)
            d.amp_slot = amp.ctypes.data_as(C.POINTER(C.c_int32))

    # -- params pytree <-> [P][bs]
    def flatten(self, params, bs, torch, device, missing_ok=()):
        """``{'lens_mass': [{name: (bs,)}, ...], ...}`` (or the upstream list-of-lists) -> fp32 [P][bs].
        Groups named in ``missing_ok`` may be absent (filled with 1.0; used by ``beta``)."""
        if isinstance(params, (list, tuple)):
            params = dict(zip(GROUPS, params))
        rows = []
        any_tensor = False
        for group, pi, name in self.slot_keys:
            try:
                v = params[group][pi][name]
            except (KeyError, IndexError) as e:
                if group in missing_ok:
                    rows.append(1.0)
                    continue
                raise KeyError(f"params['{group}'][{pi}] lacks '{name}'") from e
            any_tensor = any_tensor or torch.is_tensor(v)
            rows.append(v)
        if not rows:
            return torch.zeros((1, bs), dtype=torch.float32, device=device)
        if any_tensor:
            rows = [torch.as_tensor(v, dtype=torch.float32, device=device).reshape(-1).expand(bs) for v in rows]
            return torch.stack(rows, 0).contiguous()
        host = np.empty((len(rows), bs), dtype=np.float32)
        for i, v in enumerate(rows):
            host[i] = np.asarray(v, dtype=np.float32).reshape(-1)
        return torch.from_numpy(host).to(device)

    def unflatten(self, mat):
        """[P][bs] tensor -> params pytree of (bs,) views."""
        pm = self.phys_model
        out = {"lens_mass": [dict() for _ in pm.lenses], "lens_light": [dict() for _ in pm.lens_light],
               "source_light": [dict() for _ in pm.source_light]}
        for i, (group, pi, name) in enumerate(self.slot_keys):
            out[group][pi][name] = mat[i]
        return out


class LensSimulatorInterface:
    """``src/gigalens/simulator.py:67-127`` (``get_coords`` is legacy, unused by the fork's
    simulators and needs lenstronomy: not provided)."""

    def __init__(self, phys_model, sim_config: SimulatorConfig, bs: int):
        self.phys_model = phys_model
        self.sim_config = sim_config
        self.bs = int(bs)
        self.wcs = LensWCS(n=sim_config.num_pix, supersample=sim_config.supersample,
                           transform_pix2angle=sim_config.transform_pix2angle, pix_scale=sim_config.delta_pix)


class LensSimulator(LensSimulatorInterface):
    """Batched simulator (``src/gigalens/tf/simulator.py:13-240``) on one B200.

    ``simulate`` / ``lstsq_simulate`` / ``beta`` return CUDA ``torch`` tensors.  A simulator is
    immutable after construction and bound to (device, bs) like the reference's.
    """

    def __init__(self, phys_model, sim_config: SimulatorConfig, bs: int, device=None):
        import torch

        super().__init__(phys_model, sim_config, bs)
        self._torch = torch
        self._lib = _cabi.load()
        if not torch.cuda.is_available():
            raise RuntimeError("gigalens_b200.LensSimulator needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.supersample = int(sim_config.supersample)
        n = int(sim_config.num_pix)
        T = np.eye(2) * sim_config.delta_pix if sim_config.transform_pix2angle is None \
            else np.asarray(sim_config.transform_pix2angle)
        self.transform_pix2angle = (T / float(self.supersample)).astype(np.float32)
        # tf/simulator.py:27-29: the default transform is a float32 tensor (tf.eye(2) * delta_pix), a user-supplied array keeps its own
        # dtype through tf.linalg.det; either way the determinant is then cast to float32
        self.conversion_factor = float(np.float32(np.linalg.det(T.astype(np.float32) if sim_config.transform_pix2angle is None else T)))
        nss = n * self.supersample
        if sim_config.pix_region is None:
            img_region = np.ones((n, n), dtype=np.float32)
            mask = None
        else:
            img_region = np.asarray(sim_config.pix_region).astype(np.float32)
            if img_region.shape != (n, n):
                raise ValueError("pix_region must have shape (num_pix, num_pix)")
            mask = np.ascontiguousarray((img_region != 0).astype(np.uint8))
        self.img_region = torch.from_numpy(img_region).to(self.device)
        rows, cols = np.meshgrid(np.arange(nss), np.arange(nss), indexing="ij")
        gx, gy = self.wcs.pix2angle(cols.reshape(-1), rows.reshape(-1))  # x <- column (tf/simulator.py:45)
        self._grid_x = np.ascontiguousarray(gx, dtype=np.float32)
        self._grid_y = np.ascontiguousarray(gy, dtype=np.float32)
        self.numPix = n
        self.kernel = None
        self.flat_kernel = None
        psf = None
        if sim_config.kernel is not None:
            k = subgrid_kernel(np.asarray(sim_config.kernel), self.supersample, odd=True)[::-1, ::-1]
            psf = np.ascontiguousarray(k, dtype=np.float32)
            self.flat_kernel = psf
            self.kernel = psf
        self.compiled = CompiledModel(phys_model)
        self.depth = self.compiled.depth
        sc = _cabi.SimConfig()
        sc.num_pix, sc.supersample = n, self.supersample
        sc.grid_x = self._grid_x.ctypes.data_as(C.POINTER(C.c_float))
        sc.grid_y = self._grid_y.ctypes.data_as(C.POINTER(C.c_float))
        if psf is not None:
            sc.psf = psf.ctypes.data_as(C.POINTER(C.c_float))
            sc.psf_n = psf.shape[0]
        if mask is not None:
            sc.mask = mask.ctypes.data_as(C.POINTER(C.c_uint8))
        sc.conversion_factor = self.conversion_factor
        plan = C.c_void_p()
        _cabi.check(self._lib.gl_plan_create(C.byref(self.compiled.desc), C.byref(sc), self.bs, self.device.index,
                                             C.byref(plan)), self._lib)
        self._plan = plan
        self._like_owner = None
        self._prior_owner = None
        # the plan reserves the lstsq workspace itself when a light profile has use_lstsq; other models on first use
        self._lstsq_reserved = any(getattr(p, "use_lstsq", False) for p in list(phys_model.lens_light) + list(phys_model.source_light))

    def __del__(self):
        plan = getattr(self, "_plan", None)
        if plan:
            self._lib.gl_plan_destroy(plan)
            self._plan = None

    # -- helpers
    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    def _params_matrix(self, params, missing_ok=()):
        torch = self._torch
        if torch.is_tensor(params):
            mat = params.to(device=self.device, dtype=torch.float32).contiguous()
            if mat.shape != (max(1, self.compiled.n_params), self.bs):
                raise ValueError(f"params matrix must have shape ({self.compiled.n_params}, {self.bs})")
            return mat
        return self.compiled.flatten(params, self.bs, torch, self.device, missing_ok)

    def set_option(self, name, value):
        _cabi.check(self._lib.gl_plan_set_option(self._plan, name.encode(), int(value)), self._lib)
        if name == "lstsq_chunk":
            self._lstsq_reserved = True

    def reserve_lstsq(self, chunk=0):
        """``gl_plan_reserve_lstsq``: the component-stack workspace of ``lstsq_simulate`` (a setup call: the data-path
        entry points never allocate)."""
        _cabi.check(self._lib.gl_plan_reserve_lstsq(self._plan, int(chunk)), self._lib)
        self._lstsq_reserved = True

    # -- reference API
    def simulate(self, params, no_deflection=False):
        """``tf/simulator.py:109-156``.  Returns ``(bs, n, n)`` (``(n, n)`` when bs == 1, like ``tf.squeeze``)."""
        torch = self._torch
        mat = self._params_matrix(params)
        n = self.numPix
        img = torch.empty((self.bs, n, n), dtype=torch.float32, device=self.device)
        if no_deflection:
            self.set_option("no_deflection", 1)
        try:
            _cabi.check(self._lib.gl_simulate(self._plan, mat.data_ptr(), img.data_ptr(), self._stream()), self._lib)
        finally:
            if no_deflection:
                self.set_option("no_deflection", 0)
        return img.squeeze()

    def _simulate_variant(self, params, components, no_deflection, missing_ok):
        torch = self._torch
        mat = self._params_matrix(params, missing_ok)
        n = self.numPix
        img = torch.empty((self.bs, n, n), dtype=torch.float32, device=self.device)
        self.set_option("components", components)
        self.set_option("no_deflection", int(no_deflection))
        try:
            _cabi.check(self._lib.gl_simulate(self._plan, mat.data_ptr(), img.data_ptr(), self._stream()), self._lib)
        finally:
            self.set_option("components", 3)
            self.set_option("no_deflection", 0)
        return img.squeeze()

    def simulate_source(self, params):
        """``tf/simulator.py:242-266``: the unlensed source light (evaluated at the image-plane grid)."""
        return self._simulate_variant(params, 2, True, ("lens_mass", "lens_light"))

    def simulate_lens_light(self, params):
        """``tf/simulator.py:268-293``."""
        return self._simulate_variant(params, 1, True, ("lens_mass", "source_light"))

    def simulate_images(self, params):
        """``tf/simulator.py:295-328``: the lensed source only."""
        return self._simulate_variant(params, 2, False, ("lens_light",))

    def simulate_ss(self, params):
        """Supersampled pre-convolution image ``(bs, n*ss, n*ss)`` (``tf/simulator.py:124-141``)."""
        torch = self._torch
        mat = self._params_matrix(params)
        nss = self.numPix * self.supersample
        img = torch.empty((self.bs, nss, nss), dtype=torch.float32, device=self.device)
        _cabi.check(self._lib.gl_simulate_ss(self._plan, mat.data_ptr(), img.data_ptr(), self._stream()), self._lib)
        return img

    def _points(self, x, y):
        """Coordinates of ``beta`` / ``hessian`` / ``magnification``: the kernels evaluate points SHARED by all samples and
        return ``(bs, npts)``.  The reference passes batch-tiled ``(N, bs)`` coordinates (``init_centroids``, ``img_X``) and
        gets ``(N, bs)`` back: that layout is recognised (second dimension == bs, all columns equal) and the results come
        back transposed to ``(N, bs)``; genuinely per-sample coordinates are refused instead of being silently
        flattened."""
        host = lambda v: np.asarray(v.detach().cpu() if hasattr(v, "detach") else v, dtype=np.float32)
        x, y = host(x), host(y)
        if x.shape != y.shape:
            raise ValueError(f"x and y differ in shape: {x.shape} vs {y.shape}")
        ref_layout = x.ndim == 2 and x.shape[1] == self.bs and (self.bs > 1 or x.shape[0] >= 1)
        if ref_layout:
            if not (np.array_equal(x, np.broadcast_to(x[:, :1], x.shape)) and np.array_equal(y, np.broadcast_to(y[:, :1], y.shape))):
                raise ValueError("per-sample (N, bs) coordinates are not supported: the points must be shared by all samples "
                                 "(tile one column over the batch, as the reference's init_centroids does)")
            x, y = x[:, 0], y[:, 0]
        torch = self._torch
        xt = torch.as_tensor(np.ascontiguousarray(x.reshape(-1))).to(self.device)
        yt = torch.as_tensor(np.ascontiguousarray(y.reshape(-1))).to(self.device)
        return xt, yt, ref_layout

    def eval_points(self, params, x, y, mode=0, missing_ok=()):
        """mode 0: beta, 1: total deflection, 2: surface brightness, at points shared by all samples
        (``(bs, npts)``; ``(N, bs)`` when the coordinates came in the reference's tiled ``(N, bs)`` layout);
        mode 3: the unit-amplitude linear components of the ``use_lstsq`` light profiles, ``(bs, depth, npts)``
        (``(depth, N, bs)`` in the reference layout)."""
        torch = self._torch
        mat = self._params_matrix(params, missing_ok)
        xt, yt, ref_layout = self._points(x, y)
        npts = xt.numel()
        if mode == 3:
            o = torch.empty((self.bs, max(1, self.depth), npts), dtype=torch.float32, device=self.device)
            _cabi.check(self._lib.gl_eval_points(self._plan, mat.data_ptr(), npts, xt.data_ptr(), yt.data_ptr(), 3,
                                                 o.data_ptr(), None, self._stream()), self._lib)
            return (o.permute(1, 2, 0),) if ref_layout else (o,)
        o0 = torch.empty((self.bs, npts), dtype=torch.float32, device=self.device)
        o1 = torch.empty((self.bs, npts), dtype=torch.float32, device=self.device)
        _cabi.check(self._lib.gl_eval_points(self._plan, mat.data_ptr(), npts, xt.data_ptr(), yt.data_ptr(), int(mode),
                                             o0.data_ptr(), o1.data_ptr(), self._stream()), self._lib)
        return (o0.T, o1.T) if ref_layout else (o0, o1)

    def beta(self, x, y, lens_params: List[Dict]):
        """``tf/simulator.py:72-78`` at points ``(x, y)`` shared by all samples -> ``(bs, npts)`` each."""
        return self.eval_points({"lens_mass": lens_params}, x, y, mode=0, missing_ok=("lens_light", "source_light"))

    def hessian(self, x, y, lens_params: List[Dict]):
        """``(f_xx, f_xy, f_yx, f_yy)`` of the summed deflection at points ``(x, y)`` shared by all samples,
        ``(bs, npts)`` each -- the sum the reference's ``magnification`` / ``convergence`` / ``shear`` build
        from every profile's ``hessian`` (``tf/simulator.py:80-107``).  FP64 forward-mode duals on the GPU."""
        torch = self._torch
        mat = self._params_matrix({"lens_mass": lens_params}, ("lens_light", "source_light"))
        xt, yt, ref_layout = self._points(x, y)
        npts = xt.numel()
        H = [torch.empty((self.bs, npts), dtype=torch.float32, device=self.device) for _ in range(4)]
        _cabi.check(self._lib.gl_hessian(self._plan, mat.data_ptr(), npts, xt.data_ptr(), yt.data_ptr(),
                                         *(h.data_ptr() for h in H), self._stream()), self._lib)
        return tuple(h.T for h in H) if ref_layout else tuple(H)

    def magnification(self, x, y, lens_params: List[Dict]):
        """``tf/simulator.py:80-91``: ``1 / det(I - H)``; infinite on critical curves, like the reference."""
        f_xx, f_xy, f_yx, f_yy = self.hessian(x, y, lens_params)
        return 1.0 / ((1 - f_xx) * (1 - f_yy) - f_xy * f_yx)

    def convergence(self, x, y, lens_params: List[Dict]):
        """``tf/simulator.py:93-98``: ``kappa = (f_xx + f_yy) / 2``."""
        f_xx, _, _, f_yy = self.hessian(x, y, lens_params)
        return (f_xx + f_yy) / 2

    def shear(self, x, y, lens_params: List[Dict]):
        """``tf/simulator.py:100-107``: ``gamma1 = (f_xx - f_yy) / 2, gamma2 = f_xy``."""
        f_xx, f_xy, _, f_yy = self.hessian(x, y, lens_params)
        return (f_xx - f_yy) / 2, f_xy

    def set_positions(self, centroids_x, centroids_y, centroids_errors_x, centroids_errors_y):
        """Install the image-position data of ``ForwardProbModel`` (``tf/model.py:69-74``) in the plan."""
        n_img = np.asarray([np.size(c) for c in centroids_x], dtype=np.int32)
        cat = [np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.float32).reshape(-1) for c in group]))
               for group in (centroids_x, centroids_y, centroids_errors_x, centroids_errors_y)]
        if any(a.size != int(n_img.sum()) for a in cat):
            raise ValueError("centroids_x / centroids_y / centroids_errors_x / centroids_errors_y differ in length")
        fp = lambda a: a.ctypes.data_as(C.c_void_p)
        _cabi.check(self._lib.gl_plan_set_positions(self._plan, len(n_img), fp(n_img), *(fp(a) for a in cat)), self._lib)

    def positions_loglike(self, params, want_grad=False):
        """``stats_positions`` alone: (log_like, red_chi2[, d log_like / d params [P][bs]])."""
        torch = self._torch
        mat = self._params_matrix(params, ("lens_light", "source_light"))
        ll = torch.empty(self.bs, dtype=torch.float32, device=self.device)
        chi = torch.empty(self.bs, dtype=torch.float32, device=self.device)
        g = torch.empty((self.compiled.n_params, self.bs), dtype=torch.float32, device=self.device) if want_grad else None
        _cabi.check(self._lib.gl_positions_loglike_grad(self._plan, mat.data_ptr(), ll.data_ptr(), chi.data_ptr(),
                                                        g.data_ptr() if want_grad else None, self._stream()), self._lib)
        return (ll, chi, g) if want_grad else (ll, chi)

    def _install_lstsq_data(self, observed_image, err_map):
        """(observed, err_map) of lstsq_simulate live in the plan's likelihood slot."""
        obs = np.ascontiguousarray(np.asarray(observed_image.cpu() if hasattr(observed_image, "cpu") else observed_image),
                                   dtype=np.float32)
        err = np.ascontiguousarray(np.asarray(err_map.cpu() if hasattr(err_map, "cpu") else err_map), dtype=np.float32)
        key = (obs.tobytes(), err.tobytes())
        if self._like_owner != key:
            lc = _cabi.LikeConfig()
            lc.observed = obs.ctypes.data_as(C.POINTER(C.c_float))
            lc.error_map = err.ctypes.data_as(C.POINTER(C.c_float))
            lc.background_rms, lc.exp_time = 0.0, 1.0
            _cabi.check(self._lib.gl_plan_set_likelihood(self._plan, C.byref(lc)), self._lib)
            self._like_owner = key

    def lstsq_simulate(self, params, observed_image, err_map, return_stacked=False, return_coeffs=False,
                       no_deflection=False):
        """``tf/simulator.py:158-240``: solve the linear light amplitudes against ``observed_image``
        with weights ``1/err_map`` and return the best-fitting image ``(bs, n, n)`` (squeezed), or the
        amplitudes ``(bs, D)`` with ``return_coeffs``, or with ``return_stacked`` the convolved, down-sampled
        unit-amplitude components ``(bs, n, n, D)`` (``:203-229``)."""
        torch = self._torch
        self._install_lstsq_data(observed_image, err_map)
        if not self._lstsq_reserved:
            self.reserve_lstsq()
        mat = self._params_matrix(params)
        n = self.numPix
        if return_stacked:
            stack = torch.empty((self.bs, self.depth, n, n), dtype=torch.float32, device=self.device)
            self.set_option("no_deflection", int(no_deflection))
            try:
                _cabi.check(self._lib.gl_lstsq_stack(self._plan, mat.data_ptr(), stack.data_ptr(), self._stream()), self._lib)
            finally:
                self.set_option("no_deflection", 0)
            return stack.permute(0, 2, 3, 1)
        img = torch.empty((self.bs, n, n), dtype=torch.float32, device=self.device)
        coef = torch.empty((self.bs, self.depth), dtype=torch.float32, device=self.device)
        if no_deflection:
            self.set_option("no_deflection", 1)
        try:
            _cabi.check(self._lib.gl_lstsq_simulate(self._plan, mat.data_ptr(), img.data_ptr(), coef.data_ptr(),
                                                    self._stream()), self._lib)
        finally:
            if no_deflection:
                self.set_option("no_deflection", 0)
        return coef if return_coeffs else img.squeeze()
