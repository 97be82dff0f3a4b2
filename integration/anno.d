/**
 * anno.d -- replacement of blachlylab/fade's source/anno.d: `fade annotate` with the per-record
 * Smith-Waterman realignment on the GPU through libfadegpu (module fadegpu, integration/fadegpu.d).
 *
 * What stays as in the reference: the CLI (app.d), SAMReader / SAMWriter / header handling and the @PG line
 * (anno.d:16-33), the tag schema (rs always; am / as / ar / ab on artifact records, anno.d:94-107).
 * What changes: the body of `foreach(rec; parallel(bam.allRecords))` (anno.d:44-50).  Records are collected
 * into batches; for every record the host keeps anno.d:61-74 (early-outs, parse_clips, sc / sup bits --
 * fadehost_prepare), steps a-d of align_clip (analysis.d:34-80,98-104) run on the GPU for the whole batch,
 * and anno.d:94-107 (rs byte, tag strings -- fadehost_finish) is applied to the same records afterwards.
 * dparasail / IndexedFastaFile per-read fetches / the two mutexes are not needed any more.
 *
 * Output order: input order (the reference's is unspecified, anno.d:19).
 * GPUs: FADE_GPUS=N in the environment deals the batches round-robin to N devices (default 1).
 *
 * Not compiled in the development container (no D toolchain there); the same call sequence is what
 * fade_b200/csrc/host/bamfast.hpp (C++) runs and tests/test_gpu_cli.py checks record by record.
 */
module anno;
import std.conv : to;
import std.exception : enforce;
import std.process : environment;
import std.string : fromStringz, toStringz;
import core.stdc.string : memcpy, strlen;
import dhtslib;
import htslib.hts_log;
import fadegpu;
import util;

private string lastError(fadegpu_ctx* ctx)
{
    return fadegpu_last_error(ctx).fromStringz.idup;
}

/// the records of one batch and what annotateTask computed for them before the device call
private struct Slot
{
    fadegpu_ctx* ctx;
    fadegpu_batch* bt;
    fadegpu_batch_view v;
    SAMRecord[] recs;
    ubyte[] rsBase;          // sc / sup bits, anno.d:69-74
    int[] alignedLen, clipLeft, clipRight;
    bool inFlight;
}

private fadehost_record hostRecord(SAMRecord rec)
{
    auto b = rec.b;
    fadehost_record hr;
    hr.flag = b.core.flag;
    hr.has_sa = rec["SA"].exists ? 1 : 0;                                     // anno.d:73
    hr.cigar = cast(const(uint)*)(b.data + b.core.l_qname);
    hr.n_cigar = b.core.n_cigar;
    hr.seq4 = b.data + b.core.l_qname + (b.core.n_cigar << 2);               // as util.d:25
    hr.qual = hr.seq4 + ((b.core.l_qseq + 1) >> 1);
    hr.l_qseq = b.core.l_qseq;
    hr.tid = b.core.tid;
    hr.pos = b.core.pos;
    return hr;
}

int annotate(string cl, string[] args, ubyte con, int artifact_floor_length, int align_buffer_size)
{
    hts_set_log_level(htsLogLevel.HTS_LOG_INFO);
    hts_log_warning("fade annotate", "Output SAM/BAM keeps the input order");
    // open bam read and writer, also modify header (anno.d:22-33, unchanged)
    auto bam = SAMReader(args[1]);
    auto fai = IndexedFastaFile(args[2]);
    auto header = bam.header.dup;
    header.addLine(RecordType.PG, "ID", "fade-annotate", "PN", "fade", "VN", VERSION, "PP",
            header.valueByPos(RecordType.PG, header.numRecords(RecordType.PG) - 1, "ID"), "CL", cl);
    auto out_bam = getWriter(con, header);

    // ---- anno.d:36 (Parasail profile) + anno.d:23 (reference): one context per GPU, the reference resident in HBM ----
    fadegpu_params prm;
    fadegpu_default_params(&prm);                       // open 10, extend 2, match 2, mismatch -3 = Parasail("ACTGN",10,2,2,-3)
    prm.window_size = align_buffer_size;
    prm.min_length = artifact_floor_length;
    prm.flags = FADEGPU_F_NO_SCATTER | FADEGPU_F_TAGS_ONLY;   // only rs / am / as / ar / ab leave this loop
    int nGpus = environment.get("FADE_GPUS", "1").to!int;
    int nDev;
    enforce(fadegpu_device_count(&nDev) == 0 && nDev >= nGpus && nGpus >= 1, "fade annotate: " ~ lastError(null));
    auto ctxs = new fadegpu_ctx*[nGpus];
    immutable nTargets = bam.header.nTargets;
    string[] seqs;
    const(char)*[] namez, seqz;
    long[] lens;
    foreach (tid; 0 .. nTargets)                        // same order as the BAM header: rec.tid indexes it
    {
        auto name = bam.header.targetName(tid);
        immutable len = bam.header.targetLength(tid);
        seqs ~= fai.fetchSequence(name, ZBHO(0, len));  // any case, any letters: upper-cased on the device (analysis.d:63)
        namez ~= name.toStringz;
        seqz ~= seqs[$ - 1].ptr;
        lens ~= len;
    }
    foreach (g; 0 .. nGpus)
    {
        enforce(fadegpu_create(g, &prm, &ctxs[g]) == 0, "fade annotate: " ~ lastError(null));
        if (nTargets == 0)
            continue;
        immutable rc = g == 0 ? fadegpu_load_reference(ctxs[0], cast(int) lens.length, namez.ptr, lens.ptr, seqz.ptr)
            : fadegpu_share_reference(ctxs[g], ctxs[0]);      // GPU-to-GPU copy of the packed reference
        enforce(rc == 0, "fade annotate: " ~ lastError(ctxs[g]));
    }
    seqs = null;

    enum long BATCH = 1 << 20;
    enum long MAX_SEQ = BATCH * 160;
    auto slots = new Slot[2 * nGpus];                   // two batches per GPU: one computing, one being read / written
    foreach (i, ref s; slots)
    {
        s.ctx = ctxs[i % nGpus];
        enforce(fadegpu_alloc_batch(s.ctx, BATCH, MAX_SEQ, &s.bt) == 0 && fadegpu_get_batch_view(s.bt, &s.v) == 0,
                "fade annotate: " ~ lastError(s.ctx));
    }

    // anno.d:94-107 on the records of a finished batch, then the writer (anno.d:47-49; one thread: no mutex)
    void emit(ref Slot s)
    {
        s.inFlight = false;
        fadegpu_results_view rv;
        enforce(fadegpu_wait(s.ctx, s.bt) == 0 && fadegpu_get_results(s.bt, &rv) == 0, "fade annotate: " ~ lastError(s.ctx));
        char[] am, as_, ar, ab;
        static immutable uint[1] noOps = [0];
        foreach (i, rec; s.recs)
        {
            auto hr = hostRecord(rec);
            hr.has_sa = (s.rsBase[i] & FADE_RS_SUP) != 0;
            const(char)* cname = hr.tid >= 0 ? bam.header.targetName(hr.tid).toStringz : "".ptr;
            immutable cap = 4 * cast(size_t) hr.l_qseq + 512 + strlen(cname);
            if (am.length < cap)
            {
                am.length = cap; as_.length = cap; ar.length = cap; ab.length = cap;
            }
            immutable ri = rv.result_index[i];
            const(fadegpu_result)* res = ri >= 0 ? &rv.results[ri] : null;
            ubyte rs;
            immutable rc = fadehost_finish(&hr, cname, s.rsBase[i], s.clipLeft[i], s.clipRight[i], s.alignedLen[i],
                    s.v.flags[i], res ? rv.win_start[ri] : 0, res ? res.beg_ref : 0, res ? res.n_ops : 0,
                    res ? res.ops.ptr : noOps.ptr, &rs, am.ptr, as_.ptr, ar.ptr, ab.ptr, cap);
            enforce(rc >= 0, "fade annotate: tag buffer too small");
            rec["rs"] = rs;                                                   // anno.d:94
            if (rc == 1)                                                      // anno.d:98-107
            {
                rec["am"] = am.ptr.fromStringz.idup;
                rec["as"] = as_.ptr.fromStringz.idup;
                rec["ar"] = ar.ptr.fromStringz.idup;
                rec["ab"] = ab.ptr.fromStringz.idup;
            }
            out_bam.write(rec);
        }
        s.recs.length = 0;
        s.rsBase.length = 0;
        s.alignedLen.length = 0;
        s.clipLeft.length = 0;
        s.clipRight.length = 0;
    }

    // the compact layout of a batch (include/fadegpu.h): gate byte + 32-byte record per read + the bases as they are
    long n = 0, off = 0;
    size_t cur = 0;
    void submit()
    {
        auto s = &slots[cur];
        enforce(fadegpu_submit_compact(s.ctx, s.bt, n, off) == 0, "fade annotate: " ~ lastError(s.ctx));
        s.inFlight = true;
        n = 0;
        off = 0;
        // the ring is filled and emitted in order: the slot after this one holds the oldest batch in flight
        cur = (cur + 1) % slots.length;
        if (slots[cur].inFlight)
            emit(slots[cur]);
    }

    foreach (rec; bam.allRecords)                       // was: foreach(rec; parallel(bam.allRecords))
    {
        auto hr = hostRecord(rec);
        immutable nb = (hr.l_qseq + 1) >> 1;
        enforce(nb <= MAX_SEQ, "fade annotate: a read does not fit a batch");
        if (n == BATCH || off + nb > MAX_SEQ)
            submit();
        auto s = &slots[cur];
        int al, cl_, cr;
        ubyte rsb;
        fadehost_prepare(&hr, &al, &cl_, &cr, &rsb);    // anno.d:61-74: early-out records get clips 0 and rs 0
        memcpy(s.v.seq4 + off, hr.seq4, nb);
        fadegpu_read_meta* m = &s.v.meta[n];
        m.pos = hr.pos;
        m.seq_off = cast(uint) off;
        m.l_qseq = hr.l_qseq;
        m.tid = hr.tid;
        m.aligned_len = al;
        m.clip_left = cast(uint) cl_;
        m.clip_right = cast(uint) cr;
        immutable uint big = m.clip_left > m.clip_right ? m.clip_left : m.clip_right;
        s.v.gate[n] = cast(ubyte)(big > 255 ? 255 : big);
        s.recs ~= rec;
        s.rsBase ~= rsb;
        s.alignedLen ~= al;
        s.clipLeft ~= cl_;
        s.clipRight ~= cr;
        off += nb;
        ++n;
    }
    if (n > 0)
        submit();
    foreach (k; 0 .. slots.length)                      // drain in ring order, oldest first
    {
        auto s = &slots[(cur + k) % slots.length];
        if (s.inFlight)
            emit(*s);
    }
    foreach (ref s; slots)
        fadegpu_free_batch(s.bt);
    foreach (c; ctxs)
        fadegpu_destroy(c);
    return 0;
}
