"""GPU parity tests proper: libfadegpu (through the C ABI) against the oracle, bit-exact.
Run with `pytest -m gpu` on a B200."""
import numpy as np
import pytest

from fade_b200 import Context, default_params, sim
from parity_util import compare, oracle_params, run_gpu

pytestmark = pytest.mark.gpu


def test_c1_config_bit_exact(gpu_ctx):
    """BASELINE.json configs[0]: 1 Mbp reference, 10k simulated 2x150 reads, defaults."""
    names, contigs, cfg, n = sim.config_c1()
    gpu_ctx.load_reference(names, [c.tobytes() for c in contigs])
    rd = sim.make_reads(cfg, 0, n, contigs)
    b = run_gpu(gpu_ctx, rd)
    n_al = compare(b, rd, contigs, oracle_params(gpu_ctx.params))
    assert n_al > 1000
    st = b.stats()
    assert st.n_aligned == n_al and st.kernel_launches >= 2 and st.n_generic == 0
    b.close()


def test_empty_and_no_clip_batches(gpu_ctx):
    names, contigs, cfg, n = sim.config_c1()
    gpu_ctx.load_reference(names, [c.tobytes() for c in contigs])
    rd = sim.make_reads(cfg, 0, 64, contigs)
    b = gpu_ctx.alloc_batch(64, 64 * 75)
    b.run(0)                                     # empty batch
    rd.clip_left[:] = 0
    rd.clip_right[:] = 0
    b.fill(rd.seq4, rd.seq_off, rd.l_qseq, rd.tid, rd.pos, rd.aligned_len, rd.clip_left, rd.clip_right).run()
    assert not b.flags[:64].any() and b.stats().n_aligned == 0
    b.close()
