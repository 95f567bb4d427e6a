"""TEST INFRASTRUCTURE -- checks of the tensorflow stand-in against INDEPENDENT statements of the TensorFlow semantics it
represents (explicit numpy loops, numpy.linalg, numpy.interp), so that "the reference's code ran on the stand-in" means what
it should.  Run as a script (tests/test_reference_golden.py does, in a subprocess: the stand-in must not leak the module name
``tensorflow`` into the test process):  python oracle/tfshim/selfcheck.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.append(os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import tensorflow as tf  # noqa: E402
import tensorflow_probability as tfp  # noqa: E402

rng = np.random.default_rng(0)
tf.set_float(torch.float64)


def same_pad_corr(x, k):
    """tf.nn.conv2d(padding='SAME', strides=1) for one channel: cross-correlation, zero padding, the odd padding element at the end."""
    H, W = x.shape
    kh, kw = k.shape
    pt, pl = (kh - 1) // 2, (kw - 1) // 2
    xp = np.zeros((H + kh - 1, W + kw - 1))
    xp[pt:pt + H, pl:pl + W] = x
    out = np.zeros((H, W))
    for i in range(H):
        for j in range(W):
            out[i, j] = np.sum(xp[i:i + kh, j:j + kw] * k)
    return out


def check(name, ok):
    print(("ok   " if ok else "FAIL ") + name)
    if not ok:
        sys.exit(1)


# conv2d / depthwise_conv2d / avg_pool2d (NHWC)
x = rng.normal(size=(2, 9, 8, 1))
for ksz in ((5, 5), (3, 4)):
    k = rng.normal(size=ksz)
    got = tf.nn.conv2d(tf.constant(x), tf.constant(k[:, :, None, None]), padding="SAME", strides=1).numpy()
    want = np.stack([same_pad_corr(x[b, :, :, 0], k) for b in range(2)])[..., None]
    check(f"conv2d SAME {ksz}", np.allclose(got, want, atol=1e-12))
xc = rng.normal(size=(2, 6, 6, 3))
k = rng.normal(size=(3, 3))
got = tf.nn.depthwise_conv2d(tf.constant(xc), tf.constant(np.repeat(k[:, :, None, None], 3, axis=2)), padding="SAME", strides=[1, 1, 1, 1]).numpy()
want = np.stack([np.stack([same_pad_corr(xc[b, :, :, c], k) for c in range(3)], -1) for b in range(2)])
check("depthwise_conv2d SAME", np.allclose(got, want, atol=1e-12))
got = tf.nn.avg_pool2d(tf.constant(xc), ksize=2, strides=2, padding="SAME").numpy()
want = xc.reshape(2, 3, 2, 3, 2, 3).mean(axis=(2, 4))
check("avg_pool2d 2x2", np.allclose(got, want, atol=1e-13))

# scatter family, one-argument where
mask = rng.uniform(size=(5, 4)) > 0.4
idx = tf.where(mask)
check("where(cond) == argwhere (row-major)", np.array_equal(idx.numpy(), np.argwhere(mask)))
upd = rng.normal(size=(int(mask.sum()), 3))
base = rng.normal(size=(5, 4, 3))
want = base.copy()
for (i, j), u in zip(np.argwhere(mask), upd):
    want[i, j] += u
check("tensor_scatter_nd_add", np.allclose(tf.tensor_scatter_nd_add(tf.constant(base), idx, tf.constant(upd)).numpy(), want))
want = base.copy()
for (i, j), u in zip(np.argwhere(mask), upd):
    want[i, j] = u
check("tensor_scatter_nd_update", np.allclose(tf.tensor_scatter_nd_update(tf.constant(base), idx, tf.constant(upd)).numpy(), want))
want = np.zeros((5, 4, 3))
for (i, j), u in zip(np.argwhere(mask), upd):
    want[i, j] += u
check("scatter_nd", np.allclose(tf.scatter_nd(idx, tf.constant(upd), (5, 4, 3)).numpy(), want))
v = tf.constant(rng.normal(size=7))
g = tf.gather(v, tf.where(v > 0))
check("gather with (K,1) indices", g.shape == (int((v > 0).sum()), 1) and np.array_equal(g.numpy()[:, 0], v.numpy()[v.numpy() > 0]))

# structure / control flow
check("nest.flatten: sorted dict keys, lists in order",
      tf.nest.flatten({"b": [1, {"z": 2, "a": 3}], "a": 4}) == [4, 1, 3, 2])
n, (acc,) = tf.while_loop(lambda i, p: i < 10.0, lambda i, p: (i + 1, (p[0] + i,)), (1.0, (tf.constant(0.0),)), maximum_iterations=4)
check("while_loop honours maximum_iterations", float(n) == 5.0 and float(acc) == 1 + 2 + 3 + 4)
n, (acc,) = tf.while_loop(lambda i, p: i < 3.5, lambda i, p: (i + 1, (p[0] + i,)), (1.0, (tf.constant(0.0),)), maximum_iterations=50)
check("while_loop stops on the condition", float(n) == 4.0 and float(acc) == 6.0)
xx = tf.constant(rng.normal(size=(4, 3)))
with tf.GradientTape(persistent=True) as tape:
    tape.watch(xx)
    yy = xx ** 3
check("GradientTape.gradient of a non-scalar = gradient of its sum", np.allclose(tape.gradient(yy, xx).detach().numpy(), 3 * xx.detach().numpy() ** 2))
check("repeat([bs], axis=-1)", np.array_equal(tf.repeat(np.arange(3.0)[:, None], [4], axis=-1).numpy(), np.repeat(np.arange(3.0)[:, None], 4, axis=-1)))
check("einsum with '...'", np.allclose(tf.einsum("i,i...->i...", tf.constant(np.arange(3.0)), tf.constant(np.ones((3, 2, 2)))).numpy(),
                                         np.arange(3.0)[:, None, None] * np.ones((3, 2, 2))))
check("clip_by_value / maximum with python scalars", float(tf.clip_by_value(tf.constant(2.0), 0, 1)) == 1.0 and float(tf.math.maximum(1e-7, tf.constant(0.0))) == 1e-7)

# linalg.pinv(rcond) against numpy
a = rng.normal(size=(3, 6, 4))
a = a.transpose(0, 2, 1) @ a
a[1, :, 3] = a[1, :, 2]
a[1, 3, :] = a[1, 2, :]          # exactly rank-deficient
check("linalg.pinv(rcond=1e-6)", np.allclose(tf.linalg.pinv(tf.constant(a), rcond=1e-6).numpy(), np.linalg.pinv(a, rcond=1e-6), atol=1e-9))

# tfp.math.interp_regular_1d_grid against numpy.interp (+ fill values outside, batch of tables along axis 0)
tab = rng.normal(size=(3, 50))
xq = rng.uniform(-6, 6, size=(7, 2))
got = tfp.math.interp_regular_1d_grid(tf.constant(xq), -5.0, 5.0, tf.constant(tab), fill_value_below=0.0, fill_value_above=0.0).numpy()
grid = np.linspace(-5, 5, 50)
want = np.stack([np.interp(xq, grid, tab[c], left=0.0, right=0.0) for c in range(3)])
check("interp_regular_1d_grid", got.shape == (3, 7, 2) and np.allclose(got, want, atol=1e-12))

# float switch
tf.set_float(torch.float32)
check("tf.float32 switch", tf.constant(1.0).dtype == torch.float32 and tf.cast(tf.constant(1), tf.float32).dtype == torch.float32)
tf.set_float(torch.float64)
check("tf.float32 := float64", tf.constant(1.0, dtype=tf.float32).dtype == torch.float64 and tf.zeros((2,)).dtype == torch.float64)
print("all checks passed")
