"""numpy's float32 dot product AS EXECUTED on AVX-512 hosts (test infrastructure).

``numpy.linalg.norm(x)`` of a 1-D float32 array is ``sqrt(x.dot(x))`` (numpy/linalg/_linalg.py), and ``x.dot(x)`` is ONE
``cblas_sdot`` call (numpy/_core/src/multiarray/arraytypes.c.src: FLOAT_dot) into the OpenBLAS that numpy bundles
(scipy-openblas 0.3.30 here, DYNAMIC_ARCH -> the SkylakeX kernel on this container's and the GPU box's Xeons).  That
kernel (OpenBLAS kernel/x86_64/sdot.c + sdot_microk_skylakex-2.c, not in /root/reference: a third-party dependency of a
third-party dependency) was re-derived HERE by fitting candidate accumulation structures against ``np.dot`` bit for
bit (tests/test_oracle_golden.py::test_sdot_model_matches_numpy): 64 float32 FMA accumulators over blocks of 64
elements, one extra block of 32 into 4 x 8 accumulators, fold 64 -> 4 x 8 -> 8 -> 4 -> 1, the last n % 32 products added in
double.  scipy's LSMR (lsmr.py:239-340) takes its alpha / beta from this function, so its float32 accumulation error
(biased low, ~n^1.6: -7.6e-6 relative at 6 M elements) is part of what "the reference's result" is at the BASELINE sizes.
"""
import numpy as np

f32 = np.float32


def sdot_skx(x, y=None):
    """OpenBLAS 0.3.30 SkylakeX ``sdot(x, y)`` -> float32, emulated (python loop over n/64 blocks)."""
    x = np.ascontiguousarray(x, dtype=f32)
    y = x if y is None else np.ascontiguousarray(y, dtype=f32)
    n = len(x)
    n1 = n // 32 * 32
    nb = n1 // 64
    P = x[:nb * 64].astype(np.float64).reshape(nb, 64) * y[:nb * 64].astype(np.float64).reshape(nb, 64)
    a5 = np.zeros(64, f32)
    for b in range(nb):  # fma: exact product (48 bits fit a double) + accumulator, one rounding to float32
        a5 = (P[b] + a5.astype(np.float64)).astype(f32)
    a5 = a5.reshape(4, 16)
    acc = (a5[:, :8] + a5[:, 8:]).astype(f32)
    i = nb * 64
    if i < n1:
        p = (x[i:i + 32].astype(np.float64) * y[i:i + 32].astype(np.float64)).reshape(4, 8)
        acc = (p + acc.astype(np.float64)).astype(f32)
    a = ((acc[0] + acc[1]).astype(f32) + acc[2]).astype(f32)
    a = (a + acc[3]).astype(f32)
    h = (a[:4] + a[4:]).astype(f32)
    d = f32(f32(h[0] + h[1]) + f32(h[2] + h[3]))
    t = 0.0
    for i in range(n1, n):
        t += float(f32(x[i] * y[i]))
    return f32(t + float(d))


def chain_sumsq(x):
    """Sum of squares as the CUDA kernel k_chain_sumsq forms it: the same 64 sequential float32 FMA chains, each thread
    consuming 4 consecutive floats per step (the element -> accumulator assignment differs from sdot_skx; the
    accumulation structure and the chain length n/64 are the same)."""
    x = np.ascontiguousarray(x, dtype=f32)
    n = len(x)
    pad = (-n) % 256
    if pad:
        x = np.concatenate([x, np.zeros(pad, f32)])
    X = x.astype(np.float64).reshape(-1, 64, 4)
    X = X * X
    acc = np.zeros(64, f32)
    for t in range(X.shape[0]):
        for k in range(4):
            acc = (X[t, :, k] + acc.astype(np.float64)).astype(f32)
    a = acc.reshape(4, 16)
    q = (a[:, :8] + a[:, 8:]).astype(f32)
    a8 = ((q[0] + q[1]).astype(f32) + q[2]).astype(f32)
    a8 = (a8 + q[3]).astype(f32)
    h = (a8[:4] + a8[4:]).astype(f32)
    return f32(f32(h[0] + h[1]) + f32(h[2] + h[3]))


def has_avx512():
    try:
        with open("/proc/cpuinfo") as fh:
            return "avx512f" in fh.read()
    except OSError:
        return False
