"""Golden vectors produced by RUNNING THE REFERENCE'S OWN SOURCE FILES (``/root/reference/src/gigalens/**``, unmodified,
imported from where they lie) on the cases of ``reference_cases.py``.

tensorflow / tensorflow_probability / lenstronomy cannot be installed in this container, so the three package names resolve
to ``oracle/tfshim``: a torch-backed stand-in for the ~60 ``tf.*`` functions those files call (its header says exactly
what that does and does not pin).  Everything between those calls -- parameter conversions, clamps, ``where``s, the EPL
series loop, the dPIE complex algebra, the NFW branches, scaling-relation broadcasting, scatter onto the mask, NaN scrub,
conv / pool / flux scale, the pixel and image-position likelihoods, autodiff Hessians -- is the reference's code executing.
Each case runs twice: with ``tf.float32`` = float32 (the reference's arithmetic) and = float64 (arbiter).

Not covered: ``lstsq_simulate`` (``tf/simulator.py:158-240``) cannot execute as written -- it scatters the component values
into a buffer whose leading dimension is 0 (``:183-203``; a real TensorFlow raises the same out-of-range error the stand-in
does) and the JAX variant concatenates 3-D components onto a 4-D buffer -- so the normal-equation / pinv tail stays restated in
the oracle; ``run_lstsq`` below executes everything of it that can run (the reference's own simulator object, ``beta`` and ``light``
calls, and its source lines :204-240 read from the file), restating only the scatter of :183-203; and the TFP prior / bijector arithmetic.

    python tests/golden/make_reference_golden.py [out.npz]       (needs /root/reference; writes tests/golden/reference_golden.npz)

The fixture travels to the GPU box; this script and /root/reference do not need to."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GIGALENS_REFERENCE", "/root/reference")
# the reference package is called `gigalens`, like the alias package at the repo root: the reference must come first
sys.path[:0] = [os.path.join(ROOT, "oracle", "tfshim"), os.path.join(REF, "src")]
sys.path.append(ROOT)      # for `oracle` (the restated third-party pieces the shim borrows: subgrid_kernel, phi_n)
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import tensorflow as tf  # noqa: E402  (oracle/tfshim)

import gigalens  # noqa: E402
assert os.path.realpath(gigalens.__path__[0]).startswith(os.path.realpath(REF)), gigalens.__path__
import gigalens.tf.model as ref_model  # noqa: E402
import gigalens.tf.simulator as ref_sim  # noqa: E402
from gigalens.simulator import SimulatorConfig  # noqa: E402
from gigalens.tf.profiles.light import sersic as r_sersic, shapelets as r_shapelets  # noqa: E402
from gigalens.tf.profiles.mass import (dpie_subhalo as r_sub, epl as r_epl, nfw as r_nfw, piemd as r_piemd, piep as r_piep,  # noqa: E402
                                       scaling_relation as r_sr, shear as r_shear, sie as r_sie, sis as r_sis, tnfw as r_tnfw)

import reference_cases as RC  # noqa: E402

CLASSES = {
    "EPL": r_epl.EPL, "Shear": r_shear.Shear, "SIE": r_sie.SIE, "SIS": r_sis.SIS, "NFW": r_nfw.NFW, "NFW_ELLIPSE": r_nfw.NFW_ELLIPSE,
    "DPIS": r_piemd.DPIS, "DPIE": r_piemd.DPIE, "TNFW": r_tnfw.TNFW, "DPIEP": r_piep.DPIEP, "DPIESubhalo": r_sub.DPIESubhalo,
    "Sersic": r_sersic.Sersic, "SersicEllipse": r_sersic.SersicEllipse, "CoreSersic": r_sersic.CoreSersic,
    "Shapelets": r_shapelets.Shapelets,
    "ScaledSIS": lambda **k: r_sr.ScalingRelation(profile=r_sis.SIS(), scaling_params=["theta_E"], **k),
}


def T(a, dtype, grad=False):
    t = torch.as_tensor(np.asarray(a, dtype=np.float64)).to(dtype)
    return t.requires_grad_(True) if grad else t


def N(t):
    return t.detach().cpu().numpy()


class FakePrior:
    """``ForwardProbModel.__init__`` only needs an example draw (to size its packing bijector) and a bijector object."""

    def sample(self, *a, **k):
        return {"x": torch.zeros(())}

    def experimental_default_event_space_bijector(self):
        return object()


def promote(obj, dtype, _seen=None):
    """float32 tensors held by a reference object (and by the objects, lists and dicts it holds) -> ``dtype``."""
    _seen = set() if _seen is None else _seen
    if isinstance(obj, torch.Tensor):
        return obj.to(dtype) if obj.dtype == torch.float32 else obj
    if isinstance(obj, list):
        return [promote(v, dtype, _seen) for v in obj]
    if isinstance(obj, tuple):
        return tuple(promote(v, dtype, _seen) for v in obj)
    if isinstance(obj, dict):
        return {k: promote(v, dtype, _seen) for k, v in obj.items()}
    if hasattr(obj, "__dict__") and type(obj).__module__.startswith("gigalens") and id(obj) not in _seen:
        _seen.add(id(obj))
        for k, v in list(vars(obj).items()):
            setattr(obj, k, promote(v, dtype, _seen))
    return obj


def in_f32_then(dtype, ctor):
    """Every constant a reference constructor derives is a ``tf.float32`` number in the reference (the (L / L*)^power factors
    and catalogue constants of a scaling relation, the Shapelets tables and prefactors, the supersampled PSF, the grid, the
    flux factor, the observed image and noise constants).  The float64 arbiter run takes those float32 numbers as its
    INPUTS: objects are built with ``tf.float32`` = float32 and their constants promoted, then the methods run in float64."""
    tf.set_float(torch.float32)
    obj = ctor()
    tf.set_float(dtype)
    return promote(obj, dtype)


def make(cls, ctor, dtype):
    return in_f32_then(dtype, lambda: CLASSES[cls](**ctor))


def run_profiles(out, tag, dtype):
    tf.set_float(dtype)
    for key, c in RC.profile_cases().items():
        obj = make(c["cls"], c["ctor"], dtype)
        p = {k: T(v, dtype) for k, v in c["params"].items()}
        x, y = T(c["x"], dtype), T(c["y"], dtype)
        if c["kind"] == "mass":
            out[f"prof/{key}/{tag}/deriv"] = np.stack([N(v) for v in obj.deriv(x, y, **p)])
            # the analytic Hessians (sis, shear, nfw, piemd) and the autodiff default of tf/profile.py:9-30
            h = obj.hessian(x.clone(), y.clone(), **p)
            out[f"prof/{key}/{tag}/hessian"] = np.stack([N(v) for v in h])
        else:
            out[f"prof/{key}/{tag}/light"] = N(obj.light(x, y, **p))


def build(c, dtype, bs):
    m = c["model"]
    mk = lambda lst: [make(cls, ctor, dtype) for cls, ctor in lst]  # noqa: E731
    k = c.get("constants")
    kw = {} if k is None else dict(lenses_constants=k["lens_mass"], lens_light_constants=k["lens_light"], source_light_constants=k["source_light"])
    profs = [mk(m[g]) for g in ("lens_mass", "lens_light", "source_light")]     # (each make() switches the float type back itself)
    phys = in_f32_then(dtype, lambda: ref_model.PhysicalModel(*profs, **kw))
    s = c["sim"]
    cfg = SimulatorConfig(delta_pix=s["delta_pix"], num_pix=s["num_pix"], supersample=s["supersample"], kernel=s["kernel"],
                          pix_region=s["pix_region"], transform_pix2angle=s.get("transform_pix2angle"))
    return phys, in_f32_then(dtype, lambda: ref_sim.LensSimulator(phys, cfg, bs=bs))


def run_simulators(out, tag, dtype, psf, demo):
    tf.set_float(dtype)
    for key, c in RC.simulator_cases(psf, demo).items():
        keys = RC.grad_keys(c["params"])
        bs = len(c["params"][keys[0][0]][keys[0][1]][keys[0][2]])
        phys, sim = build(c, dtype, bs)
        params = {g: [{k: T(v, dtype, grad=True) for k, v in d.items()} for d in c["params"][g]] for g in c["params"]}
        leaves = [params[g][i][k] for g, i, k in keys]
        pre = f"sim/{key}/{tag}"
        img = sim.simulate(params)
        out[f"{pre}/image"] = N(img)
        if not c.get("image_and_likelihood_only"):
            out[f"{pre}/image_no_deflection"] = N(sim.simulate(params, no_deflection=True))
        if c.get("variants"):
            out[f"{pre}/source"] = N(sim.simulate_source(params))
            out[f"{pre}/lens_light"] = N(sim.simulate_lens_light(params))
            out[f"{pre}/images"] = N(sim.simulate_images(params))
        kw = dict(error_map=c["error_map"]) if "error_map" in c else dict(background_rms=c["noise"]["background_rms"],
                                                                        exp_time=c["noise"]["exp_time"])
        cen = c.get("centroids")
        if cen is not None:
            kw.update(centroids_x=cen["x"], centroids_y=cen["y"], centroids_errors_x=cen["ex"], centroids_errors_y=cen["ey"])
        pm = in_f32_then(dtype, lambda: ref_model.ForwardProbModel(FakePrior(), observed_image=c["observed"], include_pixels=True,
                                                                   include_positions=cen is not None, **kw))
        ll, chi2 = pm.stats_pixels(sim, params)
        out[f"{pre}/loglike"], out[f"{pre}/red_chi2"] = N(ll), N(chi2)
        g = torch.autograd.grad(ll.sum(), leaves, allow_unused=True)
        out[f"{pre}/grad"] = np.stack([N(torch.zeros(bs, dtype=dtype) if v is None else v) for v in g])
        if c.get("image_and_likelihood_only"):
            continue
        # points: beta, magnification, convergence, shear on the centroids (or a few fixed points)
        if cen is not None:
            px, py = np.concatenate(cen["x"]), np.concatenate(cen["y"])
        else:
            h = 0.4 * c["sim"]["delta_pix"] * c["sim"]["num_pix"]
            px, py = RC.f32(np.linspace(-h, h, 7) + 0.013), RC.f32(np.linspace(h, -h, 7) * 0.8 - 0.021)
        out[f"sim/{key}/points"] = np.stack([px, py])
        X, Y = T(px[:, None].repeat(bs, axis=1), dtype), T(py[:, None].repeat(bs, axis=1), dtype)
        lens = [{k: v.detach() for k, v in d.items()} for d in params["lens_mass"]]
        out[f"{pre}/beta"] = np.stack([N(v) for v in sim.beta(X, Y, lens)])
        out[f"{pre}/magnification"] = N(sim.magnification(X.clone(), Y.clone(), lens))
        out[f"{pre}/convergence"] = N(sim.convergence(X.clone(), Y.clone(), lens))
        out[f"{pre}/shear"] = np.stack([N(v) for v in sim.shear(X.clone(), Y.clone(), lens)])
        if cen is not None:
            pm.init_centroids(bs)
            llp, chi2p = pm.stats_positions(sim, params)
            out[f"{pre}/pos_loglike"], out[f"{pre}/pos_red_chi2"] = N(llp), N(chi2p)
            gp = torch.autograd.grad(llp.sum(), leaves, allow_unused=True)
            out[f"{pre}/pos_grad"] = np.stack([N(torch.zeros(bs, dtype=dtype) if v is None else v) for v in gp])
            # ForwardProbModel.log_prob / log_like (tf/model.py:126-181): how the two likelihood terms, the reduced chi^2 average over
            # the included terms and the prior are combined.  The TFP objects are replaced by stand-ins returning fixed vectors (their
            # arithmetic is not what is pinned here); z only supplies the batch size.
            from types import SimpleNamespace
            lp = T(RC.FAKE_LOG_PRIOR[:bs], dtype)
            fldj = T(RC.FAKE_FLDJ[:bs], dtype)
            pm.bij = SimpleNamespace(forward=lambda z: params)
            pm.pack_bij = SimpleNamespace(forward=lambda z: z)
            pm.unconstraining_bij = SimpleNamespace(forward_log_det_jacobian=lambda x: fldj)
            pm.prior = SimpleNamespace(log_prob=lambda p: lp)
            z = torch.zeros((bs, len(leaves)), dtype=dtype)
            logp, chi = pm.log_prob(sim, z)
            out[f"{pre}/logprob_total"], out[f"{pre}/red_chi2_total"] = N(logp), N(chi)
            out[f"{pre}/loglike_total"] = N(pm.log_like(sim, z))


def run_lstsq(out, tag, dtype, psf):
    """``lstsq_simulate`` (``tf/simulator.py:158-240``) cannot run as a whole: lines 183-203 scatter the component values into a
    buffer whose leading dimension is 0.  Everything else can.  The reference's own simulator object (its ``__init__`` builds
    ``self.kernel`` / ``self.depth``), its own ``beta`` and its own profile ``light`` calls produce the components; three lines
    here do what :183-203 meant (scatter each component onto the supersampled grid, concatenate); and the source lines from the
    first NaN scrub (:204) to the end -- transpose / reshape, depthwise conv, pooling, second NaN scrub, weights, normal
    equations, ``pinv(rcond=1e-6)``, recombination -- are read from the reference file and executed."""
    import inspect
    import textwrap

    src = inspect.getsource(ref_sim.LensSimulator.lstsq_simulate)
    body = textwrap.dedent(src[src.index("        img = tf.where(tf.math.is_nan(img), tf.zeros_like(img), img)"):])
    ns = {"tf": tf}
    exec("def lstsq_rest(self, img, observed_image, err_map, return_stacked, return_coeffs):\n" + textwrap.indent(body, "    "), ns)
    tf.set_float(dtype)
    c = RC.lstsq_sersic_case(psf)
    bs = 2
    phys, sim = build(c, dtype, bs)
    params = {g: [{k: T(v, dtype) for k, v in d.items()} for d in c["params"][g]] for g in c["params"]}
    beta_x, beta_y = sim.beta(sim.img_X, sim.img_Y, params["lens_mass"])
    nss = sim.wcs.n_x * sim.supersample
    comps = []
    for lm, p in zip(phys.lens_light, params["lens_light"]):
        comps.append(lm.light(sim.img_X, sim.img_Y, **p))
    for lm, p in zip(phys.source_light, params["source_light"]):
        comps.append(lm.light(beta_x, beta_y, **p))
    vals = torch.cat(comps, dim=0)                                                   # (D, N, bs)
    img = torch.zeros((vals.shape[0], nss, nss, bs), dtype=vals.dtype)
    img[:, sim.region[:, 0], sim.region[:, 1], :] = vals                             # what :183-203 meant
    obs, err = T(c["observed"], dtype), T(c["err_map"], dtype)
    out[f"lstsq/{tag}/stack"] = N(ns["lstsq_rest"](sim, img, obs, err, True, False))
    out[f"lstsq/{tag}/coeffs"] = N(ns["lstsq_rest"](sim, img, obs, err, False, True))
    out[f"lstsq/{tag}/image"] = N(ns["lstsq_rest"](sim, img, obs, err, False, False))


def run_lstsq_tail(out, tag, dtype, psf):
    """The tail of ``lstsq_simulate`` (``:231-240``: weights, normal equations, ``pinv(rcond=1e-6)``, recombination), read from the
    reference file and executed on the ORACLE's component stack of a model the reference's own reshape cannot take (a 15-component
    Shapelets set: ``self.depth`` counts profiles)."""
    import inspect
    import textwrap
    from types import SimpleNamespace

    from oracle import profiles as OP
    from oracle.simulator import OracleSimulator

    src = inspect.getsource(ref_sim.LensSimulator.lstsq_simulate)
    tail = textwrap.dedent(src[src.index("        W = (1 / err_map)"):])
    ns = {"tf": tf}
    exec("def lstsq_tail(self, ret, observed_image, err_map, return_coeffs):\n" + textwrap.indent(tail, "    "), ns)
    tf.set_float(dtype)
    c = RC.lstsq_tail_case(psf)
    mk = {"EPL": lambda k: OP.EPL(**k), "Shear": lambda k: OP.Shear(), "SersicEllipse": lambda k: OP.SersicEllipse(**k),
          "Shapelets": lambda k: OP.Shapelets(k["n_max"], k["use_lstsq"], k["interpolate"], dtype=dtype)}
    m = c["model"]
    om = SimpleNamespace(lenses=[mk[a](k) for a, k in m["lens_mass"]], lens_light=[mk[a](k) for a, k in m["lens_light"]],
                         source_light=[mk[a](k) for a, k in m["source_light"]], lenses_constants=[{}, {}], lens_light_constants=[{}],
                         source_light_constants=[{}])
    s = c["sim"]
    bs = 2
    osim = OracleSimulator(om, s["delta_pix"], s["num_pix"], s["supersample"], kernel=s["kernel"], bs=bs, dtype=dtype)
    params = {g: [{k: T(v, dtype) for k, v in d.items()} for d in c["params"][g]] for g in c["params"]}
    stack = osim.lstsq_stack(params)
    me = SimpleNamespace(bs=bs, depth=stack.shape[-1])
    obs, err = T(c["observed"], dtype), T(c["err_map"], dtype)
    out[f"lstsq_tail/{tag}/coeffs"] = N(ns["lstsq_tail"](me, stack, obs, err, True))
    out[f"lstsq_tail/{tag}/image"] = N(ns["lstsq_tail"](me, stack, obs, err, False))


def generate():
    psf = np.load(os.path.join(REF, "src", "gigalens", "assets", "psf.npy")).astype(np.float32)
    demo = np.load(os.path.join(REF, "src", "gigalens", "assets", "demo.npy")).astype(np.float32)
    out = {}
    for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        run_profiles(out, tag, dtype)
        run_simulators(out, tag, dtype, psf, demo)
        run_lstsq(out, tag, dtype, psf)
        run_lstsq_tail(out, tag, dtype, psf)
    tf.set_float(torch.float32)
    return out


if __name__ == "__main__":
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    out = generate()
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **out)
    print(f"{len(out)} arrays -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")
