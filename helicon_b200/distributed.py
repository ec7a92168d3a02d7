"""Multi-GPU sharding of the candidate grid.

Candidates are independent, so ranks take disjoint CHUNKS of candidates (grid.make_chunks,
cost-sorted; dealt round-robin or dynamically through grid.ChunkQueue's atomic counter) and
there is NO data-path collective; the only exchange is ONE all-gather of each rank's score
tile and local top-K at the end of a search.  ``torch.distributed`` is the plumbing (NCCL on
GPUs, gloo in the CPU tests); nothing here touches the CUDA library.
"""

from __future__ import annotations

import numpy as np


def shard_tasks(tasks, rank, world):
    """Round-robin deal of a task (or chunk) list that is already sorted by cost: neighbours cost
    about the same, so every rank gets a balanced share (``tasks[rank::world]``)."""
    return tasks[rank::world] if world > 1 else list(tasks)


def entry_from_index(ti, score, axes):
    """Top-K entry of grid candidate ``ti`` (flat index over (csym, twist, rise), the order of
    ``grid.build_tasks``) with its parameters recovered from the grid axes."""
    csyms, twists, rises = axes
    c, a, b = np.unravel_index(int(ti), (len(csyms), len(twists), len(rises)))
    return dict(score=float(score), ti=int(ti), twist=float(twists[a]), rise=float(rises[b]), csym=int(csyms[c]))


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
    ntot = int(np.prod(shape))
    k = int(top_k)
    loc = merge_topk(out["top"], k)
    # ONE collective per search: [scores | itn | flags | top-K scores | top-K ids] of this rank as one int32 tile
    tile = np.empty(3 * ntot + 2 * k, dtype=np.int32)
    tile[:ntot] = np.ascontiguousarray(out["scores"], dtype=np.float32).ravel().view(np.int32)
    tile[ntot:2 * ntot] = out["itn"].ravel().astype(np.int32)
    tile[2 * ntot:3 * ntot] = out["flags"].ravel().astype(np.uint32).view(np.int32)
    tk_sc = np.full(k, -np.inf, dtype=np.float32)
    tk_id = np.full(k, -1, dtype=np.int32)
    for i, e in enumerate(loc):
        tk_sc[i], tk_id[i] = e["score"], e["ti"]
    tile[3 * ntot:3 * ntot + k] = tk_sc.view(np.int32)
    tile[3 * ntot + k:] = tk_id
    t = torch.from_numpy(tile).to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    allt = torch.stack(parts).cpu().numpy()  # [world, 3 ntot + 2 k]
    sc_all = allt[:, :ntot].view(np.float32)
    owned = ~np.isnan(sc_all)
    if np.any(owned.sum(axis=0) > 1):
        raise RuntimeError("a grid candidate was solved by more than one rank")
    scores = np.full(ntot, np.nan, dtype=np.float32)
    itn = np.zeros(ntot, dtype=np.int32)
    flags = np.zeros(ntot, dtype=np.uint32)
    for r in range(world):
        m = owned[r]
        scores[m] = sc_all[r][m]
        itn[m] = allt[r, ntot:2 * ntot][m]
        flags[m] = allt[r, 2 * ntot:3 * ntot][m].view(np.uint32)
    g_sc = [torch.from_numpy(allt[r, 3 * ntot:3 * ntot + k].view(np.float32).copy()) for r in range(world)]
    g_id = [torch.from_numpy(allt[r, 3 * ntot + k:].copy()) for r in range(world)]
    ents = []
    axes = out.get("axes")
    mine = {e["ti"]: e for e in loc}
    for s_r, i_r in zip(g_sc, g_id):
        for s, ti in zip(s_r.cpu().tolist(), i_r.cpu().tolist()):
            if ti < 0:
                continue
            if ti in mine:
                ents.append(mine[ti])  # keeps this rank's extras (e.g. rec3d)
            elif axes is not None:
                ents.append(entry_from_index(ti, s, axes))
            else:
                ents.append(dict(score=float(s), ti=int(ti)))
    res = dict(out)
    res.update(scores=scores.reshape(shape), itn=itn.reshape(shape), flags=flags.reshape(shape),
               top=merge_topk(ents, k), n_candidates=int(owned.sum()))
    return res
