"""Deterministic synthetic problems of the shapes BASELINE.json names. Pure numpy: these build
INPUT DATA only (data columns, start points, boxes); nothing here evaluates an objective for the product."""
import numpy as np


def lorentz_problem(m, K, t_max=5.0):
    """Sum-of-Lorentzians curve fit: r_i = y_i - sum_k a_k / (1 + w (t_i - c_k)^2), x = (a_0, c_0, a_1, c_1, ...).

    K terms on a uniform grid of m abscissae in [0, t_max]; the width is tied to the centre spacing so that J^T J
    stays well conditioned at any K. Returns dict(t, y, w, x_true, x0, n)."""
    K = int(K)
    t = np.linspace(0.0, t_max, int(m))
    spacing = t_max / K
    hw = 0.8 * spacing                      # half width at half maximum
    w = 1.0 / (hw * hw)
    k = np.arange(K)
    a = 1.0 + 0.5 * np.sin(1.0 + k)
    c = (k + 0.5) * spacing
    y = np.zeros_like(t)
    for kk in range(K):                     # data generation only (zero-noise truth)
        y += a[kk] / (1.0 + w * (t - c[kk]) ** 2)
    sgn = np.where(k % 2 == 0, 1.0, -1.0)
    x_true = np.empty(2 * K)
    x_true[0::2], x_true[1::2] = a, c
    x0 = np.empty(2 * K)
    x0[0::2] = a * (1.0 + 0.1 * sgn)
    x0[1::2] = c + 0.2 * hw * sgn
    return dict(t=t, y=y, w=float(w), x_true=x_true, x0=x0, n=2 * K, m=int(m))


def rastrigin_box(n):
    return np.full(n, -5.12), np.full(n, 5.12)
