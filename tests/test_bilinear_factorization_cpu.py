"""CPU check of the identity the matrix-free trilinear operator rests on (csrc/hb2_bilinear.cuh), on
the ORACLE's trilinear rows (oracle/denovo3d_oracle.py restates SLR:1403-1510 and is pinned to the reference's matrices
by tests/test_oracle_golden.py): with tilt = psi = dy = 0 the row of (symmetry copy, image column k, ray j) touches two
neighbouring slices only, its two slice parts are the SAME in-plane footprint scaled by (1 - zf) and zf, and that
footprint does not depend on the column k."""
import numpy as np

from oracle import denovo3d_oracle as O


def _image(N, seed=3):
    rng = np.random.default_rng(seed)
    return rng.uniform(0.1, 1.0, (N, N)).astype(np.float32)


def _check(N, L3, twist, rise, csym, inner):
    A, b, pid, used = O.build_A_data_matrix(_image(N), 1.0, twist, rise, csym, 0.0, 0.0, 0.0, N, N, N, inner, L3, 0, "linear",
                                            return_blocks=True)
    A = A.tocsr()
    nd = A.shape[1] // L3
    r0 = 0
    n_rows = n_multi = 0
    worst_prop = worst_same = 0.0
    for (h, c, nrow) in used:
        fp = {}   # ray j -> (footprint of the first column seen, its column)
        for r in range(r0, r0 + nrow):
            k, j = int(pid[r]) // N, int(pid[r]) % N
            row = A.getrow(r)
            # (a copy with an integer h * rise -- h = 0 -- lets every sample truncate Z = n +- 1 ulp to n or n - 1: entries of
            # weight ~1e-16 appear on a third slice; they are below float32 resolution of the row)
            big = np.abs(row.data) > 1e-9 * np.abs(row.data).max()
            row.data, row.indices = row.data[big], row.indices[big]
            z = row.indices // nd
            zs = np.unique(z)
            # one or two NEIGHBOURING slices
            assert len(zs) <= 2 and (len(zs) == 1 or zs[1] == zs[0] + 1), (h, c, k, j, zs)
            dense = np.zeros((2, nd), dtype=np.float64)
            dense[z - zs[0], row.indices % nd] = row.data
            lo, hi = dense[0], dense[1]
            s_lo, s_hi = lo.sum(), hi.sum()
            if abs(s_lo) > 1e-3 and abs(s_hi) > 1e-3:
                # the two slice parts are proportional: hi = (zf / (1 - zf)) lo
                worst_prop = max(worst_prop, float(np.abs(hi - (s_hi / s_lo) * lo).max() / np.abs(lo).max()))
            # blend weights (1 - zf) + zf = 1 -> lo + hi IS the in-plane footprint, the same for every column of the copy
            f = lo + hi
            if h == 0:
                # integer-valued coordinates (angle 0 / 180, Z = n): which samples pass the 8-corner test follows the last-bit
                # noise of the coordinate tables per (column, sample) -- the CUDA path builds EXACT per-column maps for these
                # copies (hb2_bilinear_map.xrow / zrow); the column-independence claim is for the generic copies
                n_rows += 1
                continue
            if j in fp:
                n_multi += 1
                worst_same = max(worst_same, float(np.abs(f - fp[j]).max() / np.abs(fp[j]).max()))
            else:
                fp[j] = f
            n_rows += 1
        r0 += nrow
    return n_rows, n_multi, worst_prop, worst_same


def test_trilinear_rows_factor_into_footprint_times_slice_blend():
    for (N, L3, twist, rise, csym, inner) in ((20, 5, 31.7, 1.9, 1, 0), (18, 4, -47.3, 2.4, 2, 4)):
        n_rows, n_multi, wp, ws = _check(N, L3, twist, rise, csym, inner)
        print(f"N={N} L3={L3} twist={twist} rise={rise} csym={csym}: {n_rows} rows, {n_multi} (copy, ray) column repeats, "
              f"slice parts proportional to {wp:.1e}, footprint column-independent to {ws:.1e}")
        assert n_rows > 500 and n_multi > 200
        # float32 storage of float64 products: a few ulps of the largest weight
        assert wp < 2e-6 and ws < 2e-6
