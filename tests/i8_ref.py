"""NumPy statement of the integer-tensor-core metric (gamcsampler.jl_b200/csrc/metric_i8.cuh): fixed-point digit planes of
S = diag(sqrt w) X and the truncated sum of exact int8 digit products.  Test infrastructure (used by tests/test_gpu_i8.py and the
CPU test of the scheme's accuracy); mirrors the device arithmetic operation for operation, so the digit planes must agree bit
for bit and the metric up to the rounding of the final FP64 combination."""
import numpy as np


def column_exponents(X):
    """e_j with 2^e_j >= max_n |x_nj| / 2 (k_i8_scales: frexp(m) = (f, e + 1), f in [0.5, 1))."""
    m = np.abs(X).max(axis=0)
    _, ex = np.frexp(m)
    e = np.where((m > 0) & np.isfinite(m), ex - 1, 0).astype(np.int64)
    return e


def digit_planes(X, w, k):
    """[k, p, N] int8 digit planes (plane 0 = most significant) and the column exponents."""
    e = column_exponents(X)
    ebits = 8 * k - 2
    cs = np.ldexp(1.0, (ebits - e).astype(np.int64))
    bias = float(sum(128.0 * 256.0 ** i for i in range(k - 1)))
    sw = np.sqrt(w)
    t = X * sw[:, None]
    q = np.rint(t * cs[None, :] + bias).astype(np.int64)          # the product with a power of two is exact: one rounding, like the device fma
    planes = np.empty((k, X.shape[1], X.shape[0]), dtype=np.int8)
    for s in range(k):
        b = k - 1 - s
        byte = ((q >> (8 * b)) & 0xFF).astype(np.uint8)
        if s > 0:
            byte ^= 0x80
        planes[s] = byte.view(np.int8).T
    return planes, e


def metric_from_planes(planes, e, smax):
    """G_lik = 2^(e_i + e_j) sum_{a + b <= smax} 2^(4 - 8 (a + b)) D_a' D_b, digit products exact (float64 holds the integers)."""
    k, p, _ = planes.shape
    G = np.zeros((p, p))
    D = [planes[s].astype(np.float64) for s in range(k)]
    for s in range(2, smax + 1):          # same association as the device: all products of one weight first
        acc = np.zeros((p, p))
        for a in range(1, k + 1):
            b = s - a
            if 1 <= b <= k:
                acc += D[a - 1] @ D[b - 1].T
        G += acc * 2.0 ** (4 - 8 * s)
    c = np.ldexp(1.0, e)
    return G * np.outer(c, c)


def metric_i8(X, w, k=6, smax=7):
    planes, e = digit_planes(X, w, k)
    return metric_from_planes(planes, e, smax)
