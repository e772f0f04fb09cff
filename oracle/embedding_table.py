"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the tabulated structure embedding of
gaussreg_b200/csrc/embedding_tab.cu, used to pin its algorithmic claim on the CPU.

The reference (geotransformer/modules/geotransformer/geotransformer.py:57-72 with
modules/transformer/positional_embedding.py:19-35) evaluates, per (i, j) pair,

    emb = proj_d(sinusoid(d / sigma_d)) + max_k proj_a(sinusoid(a_k * 180 / (sigma_a * pi)))

`proj(sinusoid(x))` is a band-limited function of ONE scalar per channel (angular frequencies div_term <= 1).  The
kernel interpolates it from exact fp64 node values: cubic Hermite, step 1/8, on the angle index; quintic Hermite,
step 1/2, on the distance index.  This file builds the same tables and evaluates the same fp32 Hermite expressions;
tests/test_embedding_table_cpu.py compares them with oracle/network.structure_embedding (the fp32 reference
restatement) and with the exact fp64 function.
"""
import numpy as np

INV_H_A = 8      # 1 / step of the angle table        (csrc/embedding_tab.cu: kTabInvHA)
INV_H_D = 2      # 1 / step of the distance table     (kTabInvHD)
X_MAX_D = 1024   # the distance table covers [0, 1024] (kTabNDG)


def nodes_a(sigma_a):
    """Number of angle nodes: indices reach fl(pi_f32 * factor_a); node n + 1 must exist (tab_nodes_a)."""
    factor_a = np.float32(180.0 / (float(sigma_a) * np.pi))
    amax = np.float32(3.14159274101257324) * factor_a
    return int(np.float32(amax * np.float32(INV_H_A))) + 3


def exact(x, div, W, b, order=0):
    """f (order 0), f' (1) or f'' (2) of proj(sinusoid(x)) in fp64; x (n,), div (C/2,), W (C, C), b (C,)."""
    x = np.asarray(x, np.float64).reshape(-1, 1)
    om = np.asarray(div, np.float64).reshape(1, -1)
    s, c = np.sin(x * om), np.cos(x * om)
    if order == 0:
        E = np.stack([s, c], axis=2)
    elif order == 1:
        E = np.stack([om * c, -om * s], axis=2)
    else:
        E = np.stack([-om * om * s, -om * om * c], axis=2)
    out = E.reshape(x.shape[0], -1) @ np.asarray(W, np.float64).T
    return out + np.asarray(b, np.float64) if order == 0 else out


def build_tables(div, W_d, b_d, W_a, b_a, sigma_a, x_max_d=64):
    """(angle table (nA, 2, C), distance table (nD, 3, C)) as fp32: node values and step-scaled derivatives."""
    nA = nodes_a(sigma_a)
    xa = np.arange(nA) / INV_H_A
    ta = np.stack([exact(xa, div, W_a, b_a, 0), exact(xa, div, W_a, b_a, 1) / INV_H_A], axis=1)
    xd = np.arange(int(x_max_d) * INV_H_D + 1) / INV_H_D
    td = np.stack([exact(xd, div, W_d, b_d, 0), exact(xd, div, W_d, b_d, 1) / INV_H_D,
                   exact(xd, div, W_d, b_d, 2) / INV_H_D ** 2], axis=1)
    return ta.astype(np.float32), td.astype(np.float32)


def _f32(x):
    return np.asarray(x, np.float32)


def hermite3(tab, x):
    """Cubic Hermite in fp32, the kernel's expression order; tab (n, 2, C), x (m,) -> (m, C)."""
    u = _f32(x) * np.float32(INV_H_A)
    n = u.astype(np.int32)
    t = (u - n.astype(np.float32))[:, None]
    t2 = t * t
    t3 = t2 * t
    h01 = np.float32(3.0) * t2 - np.float32(2.0) * t3
    h10 = (t3 - np.float32(2.0) * t2) + t
    h11 = t3 - t2
    f0, g0, f1, g1 = tab[n, 0], tab[n, 1], tab[n + 1, 0], tab[n + 1, 1]
    return ((f0 + h01 * (f1 - f0)) + h10 * g0) + h11 * g1


def hermite5(tab, x):
    """Quintic Hermite in fp32; tab (n, 3, C), x (m,) -> (m, C)."""
    u = _f32(x) * np.float32(INV_H_D)
    n = u.astype(np.int32)
    t = (u - n.astype(np.float32))[:, None]
    t2 = t * t
    t3 = t2 * t
    one, two, three = np.float32(1), np.float32(2), np.float32(3)
    H3 = t3 * (np.float32(10) + t * (np.float32(-15) + np.float32(6) * t))
    H1 = t - t3 * (np.float32(6) + t * (np.float32(-8) + three * t))
    H2 = np.float32(0.5) * (t2 - t3 * (three + t * (t - three)))
    H4 = -t3 * (np.float32(4) + t * (np.float32(-7) + three * t))
    H5 = np.float32(0.5) * t3 * (one + t * (t - two))
    f0, g0, q0 = tab[n, 0], tab[n, 1], tab[n, 2]
    f1, g1, q1 = tab[n + 1, 0], tab[n + 1, 1], tab[n + 1, 2]
    return ((((f0 + H3 * (f1 - f0)) + H1 * g0) + H2 * q0) + H4 * g1) + H5 * q1


def structure_embedding(d_idx, a_idx, tables):
    """d_idx (N, N), a_idx (N, N, k) -> (N, N, C): f_d(d) + max_k f_a(a_k) from the tables."""
    ta, td = tables
    N, k = d_idx.shape[0], a_idx.shape[-1]
    fd = hermite5(td, np.asarray(d_idx).reshape(-1))
    fa = hermite3(ta, np.asarray(a_idx).reshape(-1)).reshape(N * N, k, -1).max(axis=1)
    return (fd + fa).reshape(N, N, -1).astype(np.float32)
