"""Resumable score tiles of a grid search (the reference keeps finished candidates in its on-disk
function cache, lib/cache.py:132-209, so that an interrupted ``denovo3DBatch`` / app run does not redo them).

A *tile* is what one solved batch contributes to the (csym, twist, rise) maps: flat task indices with their score,
LSMR iteration count and status flags.  ``ScoreTileStore`` appends tiles as batches finish and rewrites ONE ``.npz``
atomically (temporary file + ``os.replace``), at most every ``flush_seconds`` and at close; a new search with the same
fingerprint (image bytes, grid axes and every parameter that changes a score) starts from the stored entries and hands
only the missing candidates to the GPU.  Several ranks write ``<path>.rank<r>`` each and read all of them on resume.
Host-side bookkeeping only -- nothing on the solve path touches it.
"""

from __future__ import annotations

import glob
import hashlib
import os
import time

import numpy as np

FORMAT = 1


def fingerprint(image, axes, **params):
    """SHA-256 over the prepared image, the grid axes and the score-relevant parameters of the search."""
    h = hashlib.sha256()
    img = np.ascontiguousarray(image, dtype=np.float32)
    h.update(np.asarray(img.shape, dtype=np.int64).tobytes())
    h.update(img.tobytes())
    for a in axes:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
        h.update(b"|")
    for k in sorted(params):
        h.update(f"{k}={params[k]!r};".encode())
    return h.hexdigest()


class ScoreTileStore:
    def __init__(self, path, fp, n_total, rank=0, world=1, flush_seconds=30.0):
        self.base = str(path)
        self.path = self.base if world == 1 else f"{self.base}.rank{int(rank)}"
        if not self.path.endswith(".npz"):
            self.path += ".npz"
        self.fp, self.n_total = str(fp), int(n_total)
        self.flush_seconds = float(flush_seconds)
        self.scores = np.full(self.n_total, np.nan, dtype=np.float32)
        self.itn = np.zeros(self.n_total, dtype=np.int32)
        self.flags = np.zeros(self.n_total, dtype=np.uint32)
        self.done = np.zeros(self.n_total, dtype=bool)
        self.restored = np.zeros(self.n_total, dtype=bool)
        self._mine = np.zeros(self.n_total, dtype=bool)  # what THIS store wrote or owns (kept in its own file)
        self._dirty, self._last = False, time.monotonic()
        self.rejected = []
        self._load()

    # ---- resume -------------------------------------------------------------------------------------------------
    def _files(self):
        stem = self.base[:-4] if self.base.endswith(".npz") else self.base
        cands = {stem + ".npz", self.path}
        cands.update(glob.glob(glob.escape(stem) + ".rank*.npz"))
        return sorted(f for f in cands if os.path.exists(f))

    def _load(self):
        for f in self._files():
            try:
                with np.load(f, allow_pickle=False) as z:
                    ok = (int(z["format"]) == FORMAT and str(z["fingerprint"]) == self.fp
                          and int(z["n_total"]) == self.n_total)
                    if not ok:
                        self.rejected.append(f)
                        continue
                    ti = z["ti"].astype(np.int64)
                    sc, it, fl = z["score"], z["itn"], z["flags"]
            except Exception:  # truncated / foreign file: ignore it, the search recomputes
                self.rejected.append(f)
                continue
            if len(ti) and (ti.min() < 0 or ti.max() >= self.n_total or not (len(sc) == len(it) == len(fl) == len(ti))):
                self.rejected.append(f)
                continue
            self.scores[ti], self.itn[ti], self.flags[ti] = sc, it, fl
            self.done[ti] = True
            self.restored[ti] = True
            if os.path.abspath(f) == os.path.abspath(self.path):
                self._mine[ti] = True

    def keep_only(self, mask):
        """Several ranks: keep what EVERY rank restored (``mask`` = the AND of the ranks' ``done`` maps), so that all
        ranks cut the same task list into the same chunks; anything else is solved again."""
        drop = self.done & ~np.asarray(mask, dtype=bool)
        self.scores[drop], self.itn[drop], self.flags[drop] = np.nan, 0, 0
        self.done[drop] = self.restored[drop] = self._mine[drop] = False

    @property
    def n_restored(self):
        return int(self.restored.sum())

    def is_done(self, ti):
        return bool(self.done[int(ti)])

    # ---- recording ----------------------------------------------------------------------------------------------
    def add(self, ti, score, itn, flags):
        ti = np.asarray(ti, dtype=np.int64)
        self.scores[ti] = np.asarray(score, dtype=np.float32)
        self.itn[ti] = np.asarray(itn, dtype=np.int32)
        self.flags[ti] = np.asarray(flags, dtype=np.uint32)
        self.done[ti] = True
        self._mine[ti] = True
        self._dirty = True
        if time.monotonic() - self._last >= self.flush_seconds:
            self.flush()

    def flush(self):
        if not self._dirty:
            return
        ti = np.flatnonzero(self._mine)
        tmp = f"{self.path}.tmp{os.getpid()}"
        with open(tmp, "wb") as fh:
            np.savez(fh, format=np.int64(FORMAT), fingerprint=np.str_(self.fp), n_total=np.int64(self.n_total), ti=ti,
                     score=self.scores[ti], itn=self.itn[ti], flags=self.flags[ti])
            fh.flush()
            os.fsync(fh.fileno())
        os.replace(tmp, self.path)
        self._dirty, self._last = False, time.monotonic()

    def close(self):
        self.flush()

    # ---- host-side inspection of tile files (no GPU): search_grid itself restores into the DEVICE maps ------------
    def overlay(self, scores, itn, flags):
        """Fill the entries restored from disk into the maps read back from the device (which hold this run's)."""
        r = self.restored & ~np.isfinite(scores)
        scores[r], itn[r], flags[r] = self.scores[r], self.itn[r], self.flags[r]
        return scores, itn, flags

    def merge_topk(self, top_scores, top_index, k):
        """Global top-K of (device top-K of this run) + (restored entries): score descending, ties by task index --
        the order of the device kernel."""
        ti = np.concatenate([np.asarray(top_index, dtype=np.int64), np.flatnonzero(self.restored)])
        sc = np.concatenate([np.asarray(top_scores, dtype=np.float32), self.scores[self.restored]])
        keep = np.isfinite(sc) & (ti >= 0)
        ti, sc = ti[keep], sc[keep]
        ti, first = np.unique(ti, return_index=True)
        sc = sc[first]
        order = np.lexsort((ti, -sc.astype(np.float64)))[: int(k)]
        return sc[order], ti[order]


def agree_across_ranks(store, dist, device=0):
    """One all-reduce (MIN = logical AND) of the ranks' ``done`` maps: every rank then filters the SAME tasks, which
    the chunk queue needs (it deals indices of a chunk list that all ranks cut identically).  A tile file another rank
    could not read (or had not been flushed when that rank looked) costs a re-solve, never a wrong map."""
    import torch

    on = f"cuda:{int(device)}" if dist.get_backend() == "nccl" else "cpu"
    agreed = torch.from_numpy(store.done.astype(np.uint8)).to(on)
    dist.all_reduce(agreed, op=dist.ReduceOp.MIN)
    store.keep_only(agreed.cpu().numpy().astype(bool))
    return store
