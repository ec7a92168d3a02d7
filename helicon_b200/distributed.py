"""Multi-GPU sharding of the candidate grid.

Candidates are independent, so ranks take disjoint candidates and there is NO
data-path collective; the only exchange is the all-gather of each rank's score
tile and local top-K at the end.  ``torch.distributed`` is the plumbing (NCCL on
GPUs, gloo in the CPU tests); nothing here touches the CUDA library.
"""

from __future__ import annotations

import numpy as np


def shard_tasks(tasks, rank, world):
    """Round-robin deal of the (twist-major) task list: neighbouring candidates cost
    about the same, so every rank gets a balanced share (``tasks[rank::world]``)."""
    return tasks[rank::world] if world > 1 else list(tasks)


def merge_topk(entries, k):
    """Deterministic merge: by score descending, ties by task index ascending."""
    entries = sorted(entries, key=lambda e: (-e["score"], e["ti"]))
    return entries[:k]


def gather_grid_results(out, top_k=10, dist=None, device="cpu"):
    """All-gather the per-rank results of ``grid.search_grid(shard=(rank, world))``.

    ``out["scores"]`` / ``itn`` / ``flags`` hold values only for the tasks this rank
    solved (NaN / 0 elsewhere).  Every rank returns the merged arrays and the merged
    global top-K.  ``dist`` is ``torch.distributed`` (initialised) or None.
    """
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        res = dict(out)
        res["top"] = merge_topk(out["top"], top_k)
        return res
    import torch

    world = dist.get_world_size()
    shape = out["scores"].shape
    # one tile per rank: [score | itn | flags] as float32/int32 views of the full grid
    sc = torch.from_numpy(np.ascontiguousarray(out["scores"], dtype=np.float32).ravel()).to(device)
    meta = torch.from_numpy(np.stack([out["itn"].ravel().astype(np.int32),
                                      out["flags"].ravel().astype(np.int64).astype(np.int32)])).to(device)
    sc_all = [torch.empty_like(sc) for _ in range(world)]
    meta_all = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(sc_all, sc)
    dist.all_gather(meta_all, meta)
    sc_all = torch.stack(sc_all).cpu().numpy()
    meta_all = torch.stack(meta_all).cpu().numpy()
    owned = ~np.isnan(sc_all)
    if np.any(owned.sum(axis=0) > 1):
        raise RuntimeError("a grid candidate was solved by more than one rank")
    scores = np.full(sc_all.shape[1], np.nan, dtype=np.float32)
    itn = np.zeros(sc_all.shape[1], dtype=np.int32)
    flags = np.zeros(sc_all.shape[1], dtype=np.uint32)
    for r in range(world):
        m = owned[r]
        scores[m] = sc_all[r][m]
        itn[m] = meta_all[r][0][m]
        flags[m] = meta_all[r][1][m].astype(np.uint32)
    # local top-K as fixed-size (score, ti) tiles
    k = int(top_k)
    loc = merge_topk(out["top"], k)
    t_sc = torch.full((k,), float("-inf"), dtype=torch.float32)
    t_id = torch.full((k,), -1, dtype=torch.int64)
    for i, e in enumerate(loc):
        t_sc[i] = e["score"]
        t_id[i] = e["ti"]
    t_sc, t_id = t_sc.to(device), t_id.to(device)
    g_sc = [torch.empty_like(t_sc) for _ in range(world)]
    g_id = [torch.empty_like(t_id) for _ in range(world)]
    dist.all_gather(g_sc, t_sc)
    dist.all_gather(g_id, t_id)
    ents = []
    nt, nr = shape[-2], shape[-1]
    for s_r, i_r in zip(g_sc, g_id):
        for s, ti in zip(s_r.cpu().tolist(), i_r.cpu().tolist()):
            if ti >= 0:
                ents.append(dict(score=float(s), ti=int(ti)))
    res = dict(out)
    res.update(scores=scores.reshape(shape), itn=itn.reshape(shape), flags=flags.reshape(shape),
               top=merge_topk(ents, k), n_candidates=int(owned.sum()))
    return res
