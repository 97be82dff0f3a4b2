#!/usr/bin/env python
"""bench.py -- headline benchmark of the `fade annotate` realignment hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--reads M]

One "step" = one pass of the hot path over the whole workload (BASELINE.json configs[1]: 10 M
simulated 2x150 paired reads against a synthetic 100 Mbp chromosome, default window-size /
min-length), processed as chunks of --chunk reads through the C ABI (libfadegpu.so).

  value  reads/s with every chunk's inputs already resident in HBM: one CUDA-event interval around
         ALL kernel launches of the step's chunks (binning, fill, traceback rounds, generic kernel,
         result index), queued back to back as consecutive submits queue them
         (fadegpu_replay_batches); no host work, no copies.
  e2e    reads/s through the public C ABI with HOST buffers.  --e2e-path view (default): every
         chunk's records sit in the pinned host view of its batch (where the caller's BAM reader
         writes them); per chunk fadegpu_submit copies them to the device,
         bins them there, runs the kernels and copies the results back; fadegpu_wait +
         fadegpu_get_results hand them to the host, which reads them.  --e2e-path arrays:
         fadegpu_submit_inputs on pageable caller arrays (host binning + gather into staging).
  roofline  the INT16x2 ALU roofline of SURVEY.md 8(d): cells/s against 2*R_alu/9 with R_alu
         measured live by fadegpu_measure_alu_peak (packed VIADDMNMX.S16x2 issue rate).
  cpu_baseline  a CPU port of the path (oracle/fade_oracle_simd.c: AVX2, 16 alignments per vector,
         trace table + traceback, OpenMP on the host cores; validated against the scalar oracle)
         timed on a bounded sample of the same reads.  `--impl reference` prints that arm alone.

Multi-GPU: launched under torchrun, one rank per GPU; every rank holds its own reference copy
and its own shard of reads (weak scaling, no data-path collective); value = all ranks' reads /
max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REF_LEN = 100_000_000
REF_SEED = 1002
READ_SEED = 2002
N_READS = 10_000_000
CHUNK = 1_000_000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fade_b200", choices=["fade_b200", "reference"])
    ap.add_argument("--reads", type=int, default=N_READS, help="reads per GPU per step")
    ap.add_argument("--ref-len", type=int, default=REF_LEN)
    ap.add_argument("--chunk", type=int, default=CHUNK)
    ap.add_argument("--cpu-sample", type=int, default=4_000_000,
                    help="reads in the cpu_baseline sample (4 M reads = about 30 core-seconds of the AVX2 port)")
    ap.add_argument("--host-threads", type=int, default=0, help="host threads per rank (0 = cores / ranks)")
    ap.add_argument("--depth", type=int, default=3, help="chunks in flight in the e2e loop (view path)")
    ap.add_argument("--e2e-path", default="view", choices=["view", "arrays"],
                    help="view = fadegpu_submit from the pinned batch views (binning on the device); "
                         "arrays = fadegpu_submit_inputs from pageable arrays (binning on the host)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"],
                    help="c2 = BASELINE configs[1] (default, the bench line); c4 = stress sweep configs[3]: 2x250 reads, "
                         "--window-size 1000, clip law U{1..40} (use with --reads 2000000)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def make_workload(args, rank: int):
    from fade_b200 import sim
    t0 = time.time()
    ref = sim.make_contig(REF_SEED, 0, args.ref_len, 0, 0, 0.0)
    if getattr(args, "workload", "c2") == "c4":
        cfg = sim.default_cfg(read_seed=2004, read_len=250, window=1000, frag_mean=600, frag_sd=80, short_clip_law=1)
    else:
        cfg = sim.default_cfg(read_seed=READ_SEED)
    from fade_b200 import shard
    first, last = shard.weak_range(args.reads, rank)      # every rank owns its slice of the read stream
    rd = sim.make_reads(cfg, first, last - first, [ref], with_records=False)
    return ref, cfg, rd, time.time() - t0


CPU_KIND = ("AVX2 16-lane inter-sequence port of the path (oracle/fade_oracle_simd.c: SW fill with trace table, "
            "traceback, predicates; in-memory reference), validated against the scalar oracle")


def cpu_baseline(ref, rd, n_sample: int, threads: int, window: int = 300):
    """CPU port of the path on the host cores over the first n_sample reads: reads/s, GCUPS."""
    from oracle import oracle as orc
    prm = orc.default_params(window_size=window)
    n = min(n_sample, rd.n)
    sl = slice(0, n)
    off = rd.seq_off[: n + 1]
    contigs = [ref.tobytes()]
    t0 = time.perf_counter()
    res, _ = orc.align_batch(rd.seq4[: int(off[n])], off, rd.l_qseq[sl], rd.tid[sl], rd.pos[sl], rd.aligned_len[sl],
                             rd.clip_left[sl], rd.clip_right[sl], contigs, params=prm, n_threads=threads, simd=True)
    dt = time.perf_counter() - t0
    al = res["aligned"] == 1
    cells = int((rd.l_qseq[sl][al].astype(np.int64) * res["tlen"][al]).sum())
    return n / dt, cells / dt / 1e9, n, dt


def run_reference(args, rank: int, world: int):
    """Reference arm: the reference's CPU implementation of the path.  The real `fade` binary cannot
    be built here (D + dparasail/parasail + dhtslib/htslib are absent), so this times the oracle
    port on all host cores (kind "port"), one bounded sample per step."""
    if rank != 0:
        return
    ref, cfg, rd, _ = make_workload(argparse.Namespace(**{**vars(args), "reads": max(args.cpu_sample, 1)}), 0)
    threads = os.cpu_count() or 1          # rank 0 alone runs this arm: it may use every host core
    from fade_b200 import sim as _sim
    _sim.set_threads(threads)
    vals, gc = [], []
    for i in range(args.warmup + args.steps):
        v, g, n, dt = cpu_baseline(ref, rd, args.cpu_sample, threads, 1000 if args.workload == "c4" else 300)
        if i >= args.warmup:
            vals.append(v); gc.append(g)
    v = statistics.mean(vals)
    sample = f"first {args.cpu_sample} reads of the workload per step; {CPU_KIND}; OpenMP {threads} threads"
    line = {
        "impl": "reference", "metric": "annotate_reads_per_sec", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * args.cpu_sample / v,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": workload_config(args), "gcups": statistics.mean(gc),
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=args.out, flush=True)


def workload_config(args) -> dict:
    if getattr(args, "workload", "c2") == "c4":
        desc = (f"{args.reads} simulated 2x250 paired reads (seed 2004, clip lengths U(1..40)) vs synthetic "
                f"{args.ref_len} bp chromosome (seed {REF_SEED}), window-size 1000, min-length 5, per GPU")
    else:
        desc = (f"{args.reads} simulated 2x150 paired reads (seed {READ_SEED}) vs synthetic "
                f"{args.ref_len} bp chromosome (seed {REF_SEED}), window-size 300, min-length 5, per GPU")
    return {"workload": desc,
            "chunk_reads": args.chunk,
            "l2": "inputs + checkpoint scratch per chunk (>2 GB) exceed the 126 MB L2; no explicit flush"}


def claim_stdout():
    """Keep stdout clean for the ONE JSON line: everything else that writes to fd 1 (NCCL's version
    banner, library chatter) goes to stderr; returns a writer bound to the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    args = parse_args()
    out = claim_stdout()
    from fade_b200 import shard as _shard
    rank, local_rank, world = _shard.world()
    # torchrun exports OMP_NUM_THREADS=1; the host side of the path (binning / gather / scatter) and
    # the generators are told their thread count explicitly instead: the node's cores split by ranks
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    host_threads = args.host_threads or max(1, (os.cpu_count() or 1) // max(1, local_world))
    args.host_threads = host_threads
    args.out = out
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    from fade_b200 import sim as _sim
    _sim.set_threads(host_threads)

    import torch
    import torch.distributed as dist
    from fade_b200 import Context

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize(local_rank)
        if world > 1:
            dist.barrier()

    from fade_b200 import shard
    red = shard.Reducer(world, device=f"cuda:{local_rank}")
    max_over_ranks, sum_over_ranks = red.max, red.sum

    ref, cfg, rd, gen_s = make_workload(args, rank)
    from fade_b200 import default_params
    from fade_b200 import api as _api
    # compact results (fadegpu_get_results): fadegpu_wait scatters only flags[] and the result index
    ctx = Context(local_rank, default_params(host_threads=host_threads, flags=_api.F_NO_SCATTER,
                                             window_size=1000 if args.workload == "c4" else 300))
    ctx.load_reference(["chrS"], [ref.tobytes()])
    alu_ops, max_mhz = ctx.measure_alu_peak()

    n = rd.n
    chunk = min(args.chunk, n)
    bounds = [(a, min(n, a + chunk)) for a in range(0, n, chunk)]
    stride = (cfg.read_len + 1) // 2
    view_path = args.e2e_path == "view"
    batches = [ctx.alloc_batch(chunk, chunk * stride) for _ in range(len(bounds) if view_path else 2)]

    def load_chunk(b, a, e):
        m = e - a
        b.seq4[: m * stride] = rd.seq4[a * stride: e * stride]
        b.seq_off[: m + 1] = rd.seq_off[a: e + 1] - rd.seq_off[a]
        b.l_qseq[:m] = rd.l_qseq[a:e]
        b.tid[:m] = rd.tid[a:e]
        b.pos[:m] = rd.pos[a:e]
        b.aligned_len[:m] = rd.aligned_len[a:e]
        b.clip_left[:m] = rd.clip_left[a:e]
        b.clip_right[:m] = rd.clip_right[a:e]
        b.n = m

    agg = {"aligned": 0, "cells": 0, "h2d": 0, "d2h": 0, "launches": 0, "generic": 0, "art": 0}

    if view_path:       # the host side of every chunk: its records in the pinned view of its batch
        for b, (a, e) in zip(batches, bounds):
            load_chunk(b, a, e)

    stage = {"fill": 0.0, "trace": 0.0, "gen": 0.0}

    def collect_stats(b):
        st = b.stats()
        agg["aligned"] += st.n_aligned; agg["cells"] += st.cells; agg["launches"] += st.kernel_launches
        agg["generic"] += st.n_generic
        agg["h2d"] += st.h2d_bytes; agg["d2h"] += st.d2h_bytes
        agg["art"] += int(((b.flags[: b.n] & 6) != 0).sum())

    def kernel_step(collect: bool):
        """inputs resident: uploads untimed, then only the kernels, timed with CUDA events on the device."""
        if view_path:
            # every chunk is resident (its own batch); one timed region over the kernels of all of them,
            # queued back to back exactly as consecutive submits queue them
            if collect:
                for b in batches:
                    b.run()
                    collect_stats(b)
                    b.replay_kernels(1)          # serialised per-stage timers (fill / traceback / generic)
                    st = b.stats()
                    stage["fill"] += st.fill_ms; stage["trace"] += st.trace_ms; stage["gen"] += st.generic_ms
            return ctx.replay_batches(batches, 1)
        ms = 0.0
        for (a, e) in bounds:
            b = batches[0]
            load_chunk(b, a, e)
            b.run()                          # untimed: makes the chunk resident in HBM
            ms += b.replay_kernels(1)        # timed on the ctx stream with CUDA events
            if collect:
                collect_stats(b)
                st = b.stats()
                stage["fill"] += st.fill_ms; stage["trace"] += st.trace_ms; stage["gen"] += st.generic_ms
        return ms

    def consume(b):
        """read the step's results on the host: rs-relevant flags of every read + the compact records"""
        rec, ws, ridx = b.results()
        m8 = b.n & ~7                        # every flag byte, eight at a time
        return int(b.flags[:m8].view(np.uint64).sum(dtype=np.uint64) & 0xffff) + int(b.flags[m8: b.n].sum()) \
            + int(rec["score"].sum()) + len(ws)

    def e2e_step():
        """host buffers -> C ABI -> host results, several chunks in flight; wall clock around the whole step."""
        t0 = time.perf_counter()
        pending = None
        acc = 0
        if view_path:
            # pinned host views -> device (binning there) -> host results; fadegpu_submit only queues,
            # so up to --depth chunks are in flight while the host reads the results of the oldest
            issued = 0
            for i in range(len(bounds)):
                while issued < min(len(bounds), i + args.depth):
                    batches[issued].submit(bounds[issued][1] - bounds[issued][0])
                    issued += 1
                batches[i].wait()
                acc += consume(batches[i])
                note_copies(batches[i])
            return time.perf_counter() - t0, acc
        # --e2e-path arrays: pageable numpy arrays -> fadegpu_submit_inputs, two batches alternating
        for i, (a, e) in enumerate(bounds):
            b = batches[i & 1]
            b.submit_arrays(e - a, rd.seq4, rd.seq_off[a:], rd.l_qseq[a:], rd.tid[a:], rd.pos[a:],
                            rd.aligned_len[a:], rd.clip_left[a:], rd.clip_right[a:])   # seq_off: absolute offsets
            if pending is not None:
                pending.wait()
                acc += consume(pending)
                note_copies(pending)
            pending = b
        pending.wait()
        acc += consume(pending)
        note_copies(pending)
        return time.perf_counter() - t0, acc

    copies = {"h2d": 0, "d2h": 0}

    def note_copies(b):
        st = b.stats()
        copies["h2d"] += st.h2d_bytes; copies["d2h"] += st.d2h_bytes

    host_ms = {}

    def note_host(b):
        st = b.stats()
        for kx in ("host_submit_ms", "host_wait_ms", "host_classify_ms", "host_sort_ms", "host_gather_ms"):
            host_ms[kx] = host_ms.get(kx, 0.0) + getattr(st, kx)

    # ---- warm-up ----
    kernel_step(True)                    # makes the chunks resident, collects the per-batch statistics
    for _ in range(args.warmup):
        kernel_step(False)
    e2e_step()

    # ---- timed: kernels with resident inputs ----
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    k_ms = 0.0
    for s in range(args.steps):
        k_ms += kernel_step(False)
    barrier()
    # per-stage times of one step, from the serialised pass (stages overlap in the timed passes)
    k_fill, k_trace, k_gen = (stage[x] * args.steps for x in ("fill", "trace", "gen"))
    # ---- timed: end to end through the C ABI ----
    e_s = 0.0
    for s in range(args.steps):
        barrier()
        copies["h2d"] = copies["d2h"] = 0
        dt, _ = e2e_step()
        e_s += dt
    barrier()
    for b in batches:
        note_host(b)
    clocks = sampler.stop()

    k_ms_max = max_over_ranks(k_ms)
    e_s_max = max_over_ranks(e_s)
    total_reads = sum_over_ranks(float(n))
    total_cells = sum_over_ranks(float(agg["cells"]))
    total_aligned = sum_over_ranks(float(agg["aligned"]))
    ms_per_step = k_ms_max / args.steps
    value = total_reads / (ms_per_step * 1e-3)
    e2e_value = total_reads / (e_s_max / args.steps)

    if rank == 0:
        cells_rank = float(agg["cells"])
        achieved = cells_rank / (k_ms / args.steps * 1e-3) / 1e9            # GCUPS, all kernels of the path
        fill_gcups = cells_rank / (k_fill / args.steps * 1e-3) / 1e9 if k_fill > 0 else None
        peak = 2.0 * alu_ops / 9.0 / 1e9                                    # 9 packed instr / 2 cells
        # algorithmic HBM bytes per alignment: window 2-bit + N mask, query, metadata, result
        n_al = max(agg["aligned"], 1)
        hbm_bytes = agg["h2d"] + agg["d2h"] + n_al * 270
        cb_v, cb_g, cb_n, cb_dt = cpu_baseline(ref, rd, args.cpu_sample, host_threads, 1000 if args.workload == "c4" else 300)
        line = {
            "metric": "annotate_reads_per_sec", "value": value, "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16",
            "data": "synthetic", "config": workload_config(args),
            "gcups": total_cells / (ms_per_step * 1e-3) / 1e9,
            "aligned_reads_per_step": total_aligned, "artifact_reads_rank0": agg["art"],
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": copies["h2d"],
                    "d2h_bytes_per_step": copies["d2h"], "ms_per_step": 1e3 * e_s_max / args.steps,
                    "host_threads_per_rank": host_threads, "path": args.e2e_path,
                    "host_ms_last_chunks": {kx: round(v, 3) for kx, v in host_ms.items()}},
            "gpu_launches": agg["launches"] * args.steps,
            "roofline": {"bound": "int_alu", "achieved": achieved, "peak": peak, "unit": "GCUPS",
                         "frac": achieved / peak if peak > 0 else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel (sw_fill_kernel<19>)
                         # per launch, from the ncu --set full capture of a 1 M-read chunk of this workload
                         # (profiles/r01_prof_fill_5blk_raw.csv: 0.173 GB read + 2.293 GB written); only meaningful for the default configuration
                         "traffic": 2.466e9 if (args.workload == "c2" and chunk == 1_000_000) else None,
                         "kernel": "dominant: sw_fill_kernel<19> (80 % of the path); achieved = cells / time of ALL "
                                   "kernels of the path (binning + fill + traceback rounds + generic + result index)",
                         "fill_only_gcups": fill_gcups,
                         "fill_ms": k_fill / args.steps, "trace_ms": k_trace / args.steps,
                         "generic_ms": k_gen / args.steps,
                         "r_alu_thread_instr_per_s": alu_ops, "peak_source": "measured live (fadegpu_measure_alu_peak)",
                         "hbm": {"algorithmic_gbs": hbm_bytes / (k_ms / args.steps * 1e-3) / 1e9,
                                 "peak_gbs": peak_hbm()}},
            "cpu_baseline": {"value": cb_v, "unit": "reads/s", "cores": host_threads, "kind": "port",
                             "gcups": cb_g,
                             "sample": f"first {cb_n} reads of the workload; {CPU_KIND}; OpenMP {host_threads} threads, {cb_dt:.1f} s"},
            "gen_seconds": gen_s,
        }
        print(json.dumps(line), file=out, flush=True)
    for b in batches:
        b.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f).get("hbm_gbs")
    except Exception:
        return 6650.0


if __name__ == "__main__":
    main()
