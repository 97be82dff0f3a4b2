"""The N>1 path on CPU: two gloo ranks shard the read stream exactly like bench.py does on GPUs
(own reference copy, own reads, no data-path collective) and only reduce counters / times.  The
oracle stands in for the device as the per-shard worker, which is enough to test the sharding and
reduction logic."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, per_rank, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world_size))
    import torch.distributed as dist
    from fade_b200 import shard, sim
    from oracle import oracle as orc
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    r, lr, w = shard.world()
    assert (r, w) == (rank, world_size)
    names, contigs, cfg, _ = sim.config_c1()            # every rank builds its own reference copy
    first, last = shard.weak_range(per_rank, r)
    rd = sim.make_reads(cfg, first, last - first, contigs, with_records=False)
    res, _ = orc.align_batch(rd.seq4, rd.seq_off, rd.l_qseq, rd.tid, rd.pos, rd.aligned_len, rd.clip_left,
                             rd.clip_right, [contigs[0].tobytes()], n_threads=2)
    red = shard.Reducer(w)
    tot_reads = red.sum(rd.n)
    tot_art = red.sum(int(((res["art_left"] | res["art_right"]) == 1).sum()))
    t_max = red.max(1.0 + rank)
    np.save(os.path.join(out_dir, f"art_{rank}.npy"), (res["art_left"] | (res["art_right"] << 1)).astype(np.int8))
    if rank == 0:
        np.save(os.path.join(out_dir, "totals.npy"), np.array([tot_reads, tot_art, t_max]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    sys.path.insert(0, ROOT)
    from fade_b200 import shard, sim
    from oracle import oracle as orc
    per_rank, world_size = 1500, 2
    mp.spawn(_worker, args=(world_size, _free_port(), per_rank, str(tmp_path)), nprocs=world_size, join=True)
    names, contigs, cfg, _ = sim.config_c1()
    rd = sim.make_reads(cfg, 0, per_rank * world_size, contigs, with_records=False)
    res, _ = orc.align_batch(rd.seq4, rd.seq_off, rd.l_qseq, rd.tid, rd.pos, rd.aligned_len, rd.clip_left,
                             rd.clip_right, [contigs[0].tobytes()])
    whole = (res["art_left"] | (res["art_right"] << 1)).astype(np.int8)
    parts = np.concatenate([np.load(tmp_path / f"art_{r}.npy") for r in range(world_size)])
    assert np.array_equal(parts, whole)          # any rank count sees identical reads and results
    tot = np.load(tmp_path / "totals.npy")
    assert tot[0] == per_rank * world_size and tot[1] == int((whole != 0).sum()) and tot[2] == 2.0
    # strong-scaling helper: contiguous, balanced, exhaustive
    for total, w in ((10, 3), (7, 8), (1000, 4)):
        rs = [shard.shard_range(total, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == total and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1


def _strong_worker(rank, world_size, port, total, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world_size))
    import argparse
    import torch.distributed as dist
    import bench
    from fade_b200 import shard
    from oracle import oracle as orc
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    bench.C3_BASES = 3.0e6                                  # the 24-contig geometry of BASELINE configs[2], scaled down
    args = argparse.Namespace(workload="c3", reads=total, chunk=700, group=2, ref_len=0)
    wl = bench.Workload(args, rank, world_size)             # strong: the read stream is split over the ranks
    assert wl.scaling == "strong" and len(wl.contigs) == 24 and (wl.first, wl.last) == shard.shard_range(total, rank, world_size)
    assert sum(n for _, n in wl.groups) == wl.n and all(n <= 2 * 700 for _, n in wl.groups)
    assert [a for a, _ in wl.groups] == [wl.first + 1400 * k for k in range(len(wl.groups))]
    parts = []
    for first, n in wl.groups:                              # group by group, as the bench generates them
        rd = wl.reads(first, n)
        res, _ = orc.align_batch(rd.seq4, rd.seq_off, rd.l_qseq, rd.tid, rd.pos, rd.aligned_len, rd.clip_left, rd.clip_right,
                                 wl.contigs, n_threads=2, simd=True)
        parts.append(np.stack([res["aligned"], res["score"], res["art_left"] | (res["art_right"] << 1), rd.tid], axis=1))
    np.save(os.path.join(out_dir, f"strong_{rank}.npy"), np.concatenate(parts))
    tot = shard.Reducer(world_size).sum(wl.n)
    assert tot == total
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_strong_split_of_the_c3_workload(tmp_path):
    """bench.py --workload c3 on two gloo ranks: the 100 M-read stream of BASELINE configs[2] (here 5,000 reads, a 3 Mbp
    24-contig reference) is SPLIT over the ranks and generated group by group; read k derives its RNG stream from
    (seed, k), so the ranks' results concatenated equal a single process's over the whole stream."""
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    from oracle import oracle as orc
    total, world_size = 5000, 2
    mp.spawn(_strong_worker, args=(world_size, _free_port(), total, str(tmp_path)), nprocs=world_size, join=True)
    bench.C3_BASES = 3.0e6
    wl = bench.Workload(argparse.Namespace(workload="c3", reads=total, chunk=700, group=2, ref_len=0), 0, 1)
    rd = wl.reads(0, total)
    res, _ = orc.align_batch(rd.seq4, rd.seq_off, rd.l_qseq, rd.tid, rd.pos, rd.aligned_len, rd.clip_left, rd.clip_right,
                             wl.contigs, simd=True)
    whole = np.stack([res["aligned"], res["score"], res["art_left"] | (res["art_right"] << 1), rd.tid], axis=1)
    parts = np.concatenate([np.load(tmp_path / f"strong_{r}.npy") for r in range(world_size)])
    assert np.array_equal(parts, whole) and len(np.unique(rd.tid)) == 24 and (whole[:, 2] != 0).sum() > 200
