"""world_size-2 gloo test of the multi-GPU host logic (sharding + score/top-K all-gather).
No CUDA: the per-rank solve is replaced by a deterministic fake score."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

from helicon_b200 import distributed as D  # noqa: E402
from helicon_b200.grid import build_tasks  # noqa: E402


def _fake_local(rank, world, twists, rises):
    tasks, ntot = build_tasks(64, 64, 5.0, twists, rises, csyms=(1,), reconstruct_length_rise=3)
    mine = D.shard_tasks(tasks, rank, world)
    scores = np.full(ntot, np.nan, np.float32)
    itn = np.zeros(ntot, np.int32)
    flags = np.zeros(ntot, np.uint32)
    for t in mine:
        scores[t.ti] = np.float32(np.cos(0.37 * t.ti) * 0.5 + 0.5)
        itn[t.ti] = 100 + t.ti % 7
        flags[t.ti] = t.ti % 3
    top = sorted((dict(score=float(scores[t.ti]), ti=t.ti, twist=t.twist, rise=t.rise, csym=t.csym) for t in mine),
                 key=lambda e: (-e["score"], e["ti"]))[:5]
    shape = (1, len(twists), len(rises))
    return dict(scores=scores.reshape(shape), itn=itn.reshape(shape), flags=flags.reshape(shape), top=top,
                n_candidates=len(mine), axes=((1,), twists, rises)), tasks


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    twists = np.linspace(-3.0, 3.0, 13)  # includes |twist| < 0.01 -> skipped task (NaN on every rank)
    rises = np.linspace(4.0, 6.0, 5)
    out, tasks = _fake_local(rank, world, twists, rises)
    res = D.gather_grid_results(out, top_k=5, dist=dist)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, res["scores"], res["itn"], res["flags"],
           [(e["score"], e["ti"], e["twist"], e["rise"], e["csym"]) for e in res["top"]], res["n_candidates"]))


def test_two_rank_gather_equals_single_process():
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    twists = np.linspace(-3.0, 3.0, 13)
    rises = np.linspace(4.0, 6.0, 5)
    ref, tasks = _fake_local(0, 1, twists, rises)
    ref = D.gather_grid_results(ref, top_k=5, dist=None)
    for rank, scores, itn, flags, top, ncand in got:
        assert np.array_equal(np.isnan(scores), np.isnan(ref["scores"]))
        assert np.array_equal(np.nan_to_num(scores), np.nan_to_num(ref["scores"]))
        assert np.array_equal(itn, ref["itn"]) and np.array_equal(flags, ref["flags"])
        # every merged entry carries its parameters (denovo3DBatch reads twist / rise / csym on rank 0)
        assert top == [(e["score"], e["ti"], e["twist"], e["rise"], e["csym"]) for e in ref["top"]]
        assert ncand == len(tasks)
    # the skipped tasks (|twist| < 0.01) are NaN everywhere
    assert np.isnan(ref["scores"]).sum() == 5


def test_shards_are_disjoint_and_cover():
    tasks, ntot = build_tasks(64, 64, 5.0, np.linspace(-2, -1, 7), np.linspace(4, 5, 3))
    for world in (1, 2, 3, 8):
        seen = []
        for r in range(world):
            seen += [t.ti for t in D.shard_tasks(tasks, r, world)]
        assert sorted(seen) == [t.ti for t in tasks]
        sizes = [len(D.shard_tasks(tasks, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def _img_worker(rank, world, port, q):
    import torch.distributed as dist

    from helicon_b200 import grid as G

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    twists, rises = np.array([-2.0, -1.0]), np.array([4.5, 5.0, 5.5])

    def fake_search_grid(image, apix, tw, ri, csyms=(1,), **kw):  # stands in for the GPU solve
        sc = (np.float32(image.mean()) + 0.01 * np.arange(len(tw) * len(ri), dtype=np.float32)).reshape(1, len(tw), len(ri))
        a, b = np.unravel_index(np.argmax(sc[0]), sc[0].shape)
        return dict(scores=sc, n_candidates=sc.size,
                    top=[dict(score=float(sc[0, a, b]), twist=float(tw[a]), rise=float(ri[b]), csym=1)])

    G.search_grid, keep = fake_search_grid, G.search_grid
    images = [np.full((8, 8), float(i), np.float32) for i in range(5)]
    out = G.search_images(images, 5.0, twists, rises, shard=(rank, world), dist=dist, gather_device="cpu")
    G.search_grid = keep
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, out["scores"], out["best"], out["n_candidates"]))


def test_per_image_search_sharded_by_image_and_gathered():
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_img_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, scores, best, ncand in got:
        assert scores.shape == (5, 1, 2, 3) and np.all(np.isfinite(scores)) and ncand == 30
        for i in range(5):
            assert np.isclose(scores[i].max(), i + 0.05) and best[i][1:] == (-1.0, 5.5, 1)


def _queue_worker(rank, world, port, q):
    import time

    import torch.distributed as dist

    from helicon_b200.grid import ChunkQueue

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    got = []
    for n in (23, 5):  # two queues in a row (one per search): the counter keys follow the creation order
        mine = []
        for i in ChunkQueue(n, shard=(rank, world), dist=dist):
            mine.append(i)
            time.sleep(0.002 * (1 + 3 * rank))  # rank 1 is the slow one: it must end up with fewer chunks
        got.append(mine)
        dist.barrier()
    dist.destroy_process_group()
    q.put((rank, got))


def test_dynamic_chunk_queue_deals_every_chunk_once():
    """grid.ChunkQueue under gloo, world size 2: the atomic counter of the process group's store hands every chunk to
    exactly one rank, in order, and the faster rank takes more of them (SURVEY 8e dynamic deal)."""
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_queue_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for qi, n in enumerate((23, 5)):
        a, b = got[0][qi], got[1][qi]
        assert sorted(a + b) == list(range(n)) and a == sorted(a) and b == sorted(b)
    assert len(got[0][0]) > len(got[1][0])


def test_static_chunk_deal_without_process_group():
    from helicon_b200.grid import ChunkQueue

    for world in (1, 2, 3):
        seen = []
        for r in range(world):
            seen += list(ChunkQueue(10, shard=(r, world)))
        assert sorted(seen) == list(range(10))


def test_make_chunks_cost_sorted_and_complete():
    from helicon_b200.grid import make_chunks

    tasks, ntot = build_tasks(384, 384, 1.3, np.linspace(-30, 30, 5), np.array([21.0, 33.0, 47.0]), csyms=(1, 2))
    chunks = make_chunks(tasks, lambda key: 1000, batch_candidates=4, positive_constraint=0)
    assert sorted(t.ti for _, ch, _ in chunks for t in ch) == sorted(t.ti for t in tasks)
    costs = [c for _, _, c in chunks]
    assert costs == sorted(costs, reverse=True)
    for key, ch, _ in chunks:  # one chunk = one batch shape
        assert len({(t.geom["L3"], t.geom["D3"]) for t in ch}) == 1 and len(ch) <= 4


def _resume_worker(rank, world, port, q, base):
    import torch.distributed as dist

    from helicon_b200 import checkpoint

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # first "run": every rank writes its own tile file
    st = checkpoint.ScoreTileStore(base, "fp", 16, rank=rank, world=world)
    st.add([rank, 8 + rank], [0.5 + 0.1 * rank, 0.7], [3, 4], [0, 0])
    st.close()
    dist.barrier()
    if rank == 1:  # rank 1 loses sight of rank 0's file (stale directory listing / unreadable file)
        os.rename(base + ".rank0.npz", base + ".hidden")
    st = checkpoint.ScoreTileStore(base, "fp", 16, rank=rank, world=world) if rank == 1 else None
    dist.barrier()
    if rank == 1:
        os.rename(base + ".hidden", base + ".rank0.npz")
    dist.barrier()
    if rank == 0:
        st = checkpoint.ScoreTileStore(base, "fp", 16, rank=rank, world=world)
    before = st.n_restored
    checkpoint.agree_across_ranks(st, dist)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, before, np.flatnonzero(st.done).tolist()))


def test_two_rank_resume_agrees_on_the_restored_set(tmp_path):
    """Resumable searches under several ranks: a tile file one rank could not read is dropped by ALL ranks (one
    all-reduce), so the task lists -- and the chunk lists the queue deals from -- stay identical."""
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    base = str(tmp_path / "tiles")
    procs = [ctx.Process(target=_resume_worker, args=(r, world, port, q, base)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict((r, (b, d)) for r, b, d in (q.get(timeout=120) for _ in range(world)))
    for p in procs:
        p.join(timeout=60)
    assert got[0][0] == 4 and got[1][0] == 2  # rank 0 saw both files, rank 1 only its own
    assert got[0][1] == got[1][1] == [1, 9]   # after the all-reduce both keep exactly rank 1's entries
