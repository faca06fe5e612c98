"""Nash solver oracle (numpy).  TEST INFRASTRUCTURE — see oracle/__init__.py.
Restates util/projection.py:9-38 ``projection_simplex`` and environments/nash_sampler.py:24-58 ``Game`` /
``get_nash`` (projected gradient on the bilinear game, averaged iterates)."""
import numpy as np


def projection_simplex(x, max_nz):
    x = np.asarray(x)
    n = x.shape[0]
    masked = np.where(np.arange(n) < max_nz, x, -np.inf)
    idx = np.argsort(-masked, kind="stable")                 # descending
    vals = x[idx]
    cs = np.cumsum(vals)
    cs = np.where(np.arange(n) < max_nz, cs, 0)
    ind = np.arange(n) + 1
    with np.errstate(invalid="ignore"):
        cond = np.nan_to_num(1 / ind + (vals - cs / ind)) > 0
    cond = np.where(np.arange(n) < max_nz, cond, False)
    k = int(np.count_nonzero(cond))
    to_relu = 1 / k + (vals - cs[k - 1] / k)
    to_relu = np.where(np.arange(n) < max_nz, to_relu, 0)
    proj = np.maximum(to_relu, 0)
    out = np.zeros_like(x)
    out[idx] = proj
    return out


def get_nash(game, x, y, x_nz, y_nz, num_iters=10000, lr=0.01):
    xs, ys = x.copy(), y.copy()
    for _ in range(num_iters):
        xg = game @ y
        xn = projection_simplex(x - lr * xg, x_nz)
        yg = -(x @ game)
        yn = projection_simplex(y - lr * yg, y_nz)
        x, y = xn, yn
        xs += x; ys += y
    return xs / (num_iters + 1), ys / (num_iters + 1)
