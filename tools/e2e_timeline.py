#!/usr/bin/env python
"""Per-chunk timeline of the e2e loop of bench.py (c2 workload): host wall clock of submit / wait per chunk and the
device events of every batch (uploads start, inputs ready, kernels done, results home), to see where an e2e step
spends the time it spends above the kernel-only time.
    python tools/e2e_timeline.py [--chunk N] [--depth D] [--path compact|view] [--reads M]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from fade_b200 import Context, api, default_params, sim  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--chunk", type=int, default=1_000_000)
    ap.add_argument("--depth", type=int, default=3)
    ap.add_argument("--path", default="compact")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--consume", type=int, default=1)
    a = ap.parse_args()
    sim.set_threads(a.threads)
    ref = sim.make_contig(1002, 0, 100_000_000, 0, 0, 0.0)
    cfg = sim.default_cfg(read_seed=2002)
    rd = sim.make_reads(cfg, 0, a.reads, [ref], with_records=False)
    stride = 75
    ctx = Context(0, default_params(host_threads=a.threads, flags=api.F_NO_SCATTER))
    ctx.load_reference(["chrS"], [ref])
    bounds = [(x, min(rd.n, x + a.chunk)) for x in range(0, rd.n, a.chunk)]
    bs = [ctx.alloc_batch(a.chunk, a.chunk * stride) for _ in bounds]
    for b, (x, e) in zip(bs, bounds):
        args = (rd.seq4[x * stride: e * stride], rd.seq_off[x: e + 1] - rd.seq_off[x], rd.l_qseq[x:e], rd.tid[x:e], rd.pos[x:e],
                rd.aligned_len[x:e], rd.clip_left[x:e], rd.clip_right[x:e])
        (b.fill_compact if a.path == "compact" else b.fill)(*args)
    sub = (lambda b: b.submit_compact()) if a.path == "compact" else (lambda b: b.submit())

    def step(log):
        t0 = time.perf_counter()
        issued = 0
        for i in range(len(bs)):
            while issued < min(len(bs), i + a.depth):
                ts = time.perf_counter()
                sub(bs[issued])
                log.append(("submit", issued, ts - t0, time.perf_counter() - t0))
                issued += 1
            ts = time.perf_counter()
            bs[i].wait()
            tw = time.perf_counter()
            if a.consume:
                rec, ws, ridx = bs[i].results()
                _ = int(bs[i].flags[: bs[i].n].view(np.uint64).sum(dtype=np.uint64)) + int(rec["score"].sum())
            log.append(("wait", i, ts - t0, tw - t0, time.perf_counter() - t0))
        return time.perf_counter() - t0

    for _ in range(3):
        step([])
    ms_k = ctx.replay_batches(bs, 1)
    log = []
    dt = step(log)
    print(f"kernel-only {ms_k:.2f} ms; e2e step {dt * 1e3:.2f} ms ({a.path}, chunk {a.chunk}, depth {a.depth})")
    for ev in log:
        if ev[0] == "wait":
            tl = bs[ev[1]].timeline(bs[0])
            st = bs[ev[1]].stats()
            print(f"chunk {ev[1]:2d}: wait called {ev[2] * 1e3:7.2f} returned {ev[3] * 1e3:7.2f} consumed {ev[4] * 1e3:7.2f} | device: up {tl[0]:7.2f} inputs {tl[4]:7.2f} "
                  f"fill {tl[1]:7.2f} - {tl[5]:7.2f} kernels done {tl[2]:7.2f} home {tl[3]:7.2f} | host submit {st.host_submit_ms:.2f} "
                  f"(classify wait {st.host_classify_ms:.2f}) wait-scatter {st.host_wait_ms:.2f}")
        else:
            print(f"   submit {ev[1]:2d} at {ev[2] * 1e3:7.2f} ({(ev[3] - ev[2]) * 1e3:.3f} ms)")
    for b in bs:
        b.close()
    ctx.close()


if __name__ == "__main__":
    main()
