"""Host logic added in round 2 that needs no GPU: resumable score tiles (checkpoint.py) and the LSQR wrapper's row
masking / right-hand side (lsqr.py) on a stand-in batch whose products are scipy CSR products."""

import os

import numpy as np
from scipy.sparse import random as sprandom
from scipy.sparse.linalg import lsqr

from helicon_b200 import checkpoint, lsqr as hlsqr


def _fp(img, **kw):
    axes = (np.array([1.0]), np.linspace(-3, 3, 7), np.linspace(4, 6, 5))
    return checkpoint.fingerprint(img, axes, apix=1.3, positive_constraint=-1, **kw)


def test_fingerprint_changes_with_every_input():
    rng = np.random.default_rng(0)
    img = rng.standard_normal((8, 8)).astype(np.float32)
    base = _fp(img)
    assert base == _fp(img.copy())
    img2 = img.copy()
    img2[3, 3] += 1e-3
    assert _fp(img2) != base
    assert _fp(img, interpolation="linear") != _fp(img, interpolation="nn")
    assert _fp(img.reshape(4, 16)) != base


def test_store_round_trip_and_resume(tmp_path):
    path = tmp_path / "tiles"
    st = checkpoint.ScoreTileStore(path, "abc", 35, flush_seconds=1e9)
    assert st.n_restored == 0
    st.add([3, 4, 10], [0.5, 0.7, 0.6], [11, 12, 13], [0, 1, 0])
    st.add([20], [0.9], [14], [2])
    assert not os.path.exists(st.path)  # nothing written before the flush interval / close
    st.close()
    assert os.path.exists(st.path) and not [f for f in os.listdir(tmp_path) if ".tmp" in f]

    st2 = checkpoint.ScoreTileStore(path, "abc", 35)
    assert st2.n_restored == 4 and st2.is_done(10) and not st2.is_done(11)
    assert np.array_equal(np.flatnonzero(st2.restored), [3, 4, 10, 20])
    assert np.allclose(st2.scores[[3, 4, 10, 20]], [0.5, 0.7, 0.6, 0.9]) and list(st2.itn[[3, 20]]) == [11, 14]
    assert list(st2.flags[[4, 20]]) == [1, 2]
    # a second interrupted run adds to the same file and keeps the first run's entries
    st2.add([11], [0.8], [9], [0])
    st2.close()
    st3 = checkpoint.ScoreTileStore(path, "abc", 35)
    assert st3.n_restored == 5

    # maps read back from the device hold only the new run's entries: overlay fills the restored ones in
    sc = np.full(35, np.nan, dtype=np.float32)
    it = np.zeros(35, dtype=np.int32)
    fl = np.zeros(35, dtype=np.uint32)
    sc[30], it[30] = 0.95, 7
    st3.overlay(sc, it, fl)
    assert np.isfinite(sc).sum() == 6 and sc[30] == np.float32(0.95) and it[20] == 14 and fl[20] == 2
    tsc, tix = st3.merge_topk([0.95], [30], 3)
    assert list(tix) == [30, 20, 11] and np.allclose(tsc, [0.95, 0.9, 0.8])


def test_store_rejects_foreign_and_broken_files(tmp_path):
    path = tmp_path / "tiles.npz"
    st = checkpoint.ScoreTileStore(path, "grid-A", 10)
    st.add([1], [0.5], [3], [0])
    st.close()
    other = checkpoint.ScoreTileStore(path, "grid-B", 10)  # other image / grid / parameters: recompute everything
    assert other.n_restored == 0 and other.rejected == [str(path)]
    assert checkpoint.ScoreTileStore(path, "grid-A", 11).n_restored == 0  # other grid size
    with open(path, "wb") as fh:
        fh.write(b"PK\x03\x04 truncated")
    broken = checkpoint.ScoreTileStore(path, "grid-A", 10)
    assert broken.n_restored == 0 and broken.rejected
    broken.add([2], [0.25], [1], [0])
    broken.close()  # the broken file is replaced atomically
    assert checkpoint.ScoreTileStore(path, "grid-A", 10).n_restored == 1


def test_store_per_rank_files_are_all_read_on_resume(tmp_path):
    path = tmp_path / "tiles"
    for r in range(2):
        st = checkpoint.ScoreTileStore(path, "fp", 12, rank=r, world=2)
        st.add([r, r + 6], [0.1 * (r + 1), 0.3], [5, 6], [0, 0])
        st.close()
        assert st.path.endswith(f".rank{r}.npz")
    for r in range(2):  # every rank sees what ALL ranks had finished, and rewrites only its own entries
        st = checkpoint.ScoreTileStore(path, "fp", 12, rank=r, world=2)
        assert st.n_restored == 4 + r  # rank 1 resumes after rank 0 has already added one entry
        st.add([10 + r], [0.9], [1], [0])
        st.close()
    one = checkpoint.ScoreTileStore(path, "fp", 12)  # resumed on ONE GPU: reads the rank files too
    assert one.n_restored == 6
    with np.load(str(path) + ".rank0.npz") as z:
        assert sorted(z["ti"].tolist()) == [0, 6, 10]


def test_keep_only_drops_what_another_rank_did_not_see(tmp_path):
    st = checkpoint.ScoreTileStore(tmp_path / "t", "fp", 6)
    st.add([0, 1, 4], [0.1, 0.2, 0.3], [1, 2, 3], [0, 0, 1])
    st.close()
    st = checkpoint.ScoreTileStore(tmp_path / "t", "fp", 6)
    agreed = np.array([1, 0, 0, 0, 1, 1], dtype=bool)  # the AND over ranks: entry 1 was not restored everywhere
    st.keep_only(agreed)
    assert st.n_restored == 2 and not st.is_done(1) and np.isnan(st.scores[1]) and st.is_done(4)
    st.add([1], [0.25], [2], [0])  # solved again in this run
    st.close()
    assert checkpoint.ScoreTileStore(tmp_path / "t", "fp", 6).n_restored == 3


def test_merge_topk_orders_like_the_device_kernel(tmp_path):
    st = checkpoint.ScoreTileStore(tmp_path / "t", "fp", 8)
    st.add([5, 2], [0.5, 0.5], [1, 1], [0, 0])
    st.close()
    st = checkpoint.ScoreTileStore(tmp_path / "t", "fp", 8)
    tsc, tix = st.merge_topk([0.5, 0.4], [3, 7], 4)  # equal scores: the lower task index first
    assert list(tix) == [2, 3, 5, 7]
    tsc, tix = st.merge_topk([0.5, 0.5], [2, 3], 2)  # an index present on both sides counts once
    assert list(tix) == [2, 3]


class _FakeBatch:
    """The padded row layout of engine.Batch on top of a scipy matrix: m real data rows scattered over nd_pad slots,
    then the symmetry rows; padded slots carry junk that the wrapper must mask."""

    def __init__(self, seed=0, m=60, ms=25, n=30, nd_pad=80):
        rng = np.random.default_rng(seed)
        self.n, self.nd_pad, self.tot = n, nd_pad, nd_pad + ms
        self.A = sprandom(m + ms, n, density=0.3, random_state=seed, dtype=np.float32, format="csr")
        self.b = np.concatenate([rng.standard_normal(m), np.zeros(ms)]).astype(np.float32)
        self.slots = np.sort(rng.choice(nd_pad, m, replace=False))
        self.rows = np.concatenate([self.slots, np.arange(nd_pad, self.tot)])
        self.junk = rng.standard_normal(self.tot).astype(np.float32)

    def rows_padded(self, c):
        return self.nd_pad, self.tot

    def rhs_padded(self, c):
        out = self.junk[: self.nd_pad].copy()
        out[self.slots] = self.b[: len(self.slots)]
        return out

    def data_row_index(self, c):
        return self.slots, None, None

    def apply_forward(self, c, x):
        y = self.junk.copy()
        y[self.rows] = self.A @ x
        return y

    def apply_adjoint(self, c, y):
        full = np.ones(self.tot, dtype=bool)
        full[self.rows] = False
        return self.A.T @ y[self.rows] + np.float32(y[full].sum())  # junk rows would leak into x if not masked


def test_lsqr_wrapper_equals_scipy_on_the_real_rows():
    fb = _FakeBatch()
    info = {}
    x = hlsqr.solve_lsqr(fb, 0, atol=1e-6, btol=1e-6, info=info)
    ref = lsqr(fb.A, fb.b, atol=1e-6, btol=1e-6)
    assert info["itn"] == ref[2] and info["istop"] == ref[1]
    assert x.dtype == np.float64 and np.allclose(x, ref[0], rtol=1e-5, atol=1e-6)
    assert info["forward_products"] == ref[2] and info["adjoint_products"] == ref[2] + 1
    real = hlsqr.real_row_mask(fb)
    assert real.sum() == fb.A.shape[0] and np.array_equal(np.flatnonzero(real), fb.rows)
