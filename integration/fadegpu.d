/**
 * fadegpu.d -- D binding of libfadegpu's C ABI (include/fadegpu.h, ABI version 2) and of the host helpers
 * (include/fadehost.h).  Drop into blachlylab/fade's source/ next to anno.d (integration/anno.d replaces
 * source/anno.d) and link with -L-lfadegpu (see integration/README.md).
 *
 * Every declaration mirrors the C header field for field; static asserts below pin the struct sizes the
 * C side checks as well (tests/test_abi_and_host.py).  Not compiled in the development container (no D
 * toolchain there): compile with ldc2 / dmd >= 2.090.
 */
module fadegpu;

extern (C) @nogc nothrow:

enum FADEGPU_ABI_VERSION = 2;
enum FADEGPU_MAX_OPS = 10;

enum : int
{
    FADEGPU_OK = 0,
    FADEGPU_E_ARG = -1,
    FADEGPU_E_CUDA = -2,
    FADEGPU_E_OOM = -3,
    FADEGPU_E_STATE = -4,
    FADEGPU_E_NODEV = -5
}

struct fadegpu_ctx;
struct fadegpu_batch;

struct fadegpu_params
{
    int window_size;    /// --window-size, app.d:18
    int min_length;     /// --min-length,  app.d:17
    int gap_open;       /// 10, anno.d:36
    int gap_extend;     /// 2
    int match;          /// 2
    int mismatch;       /// -3
    uint flags;         /// FADEGPU_F_*
    long scratch_bytes; /// 0 = default
    int host_threads;   /// 0 = all cores
    int reserved;
}

enum : uint
{
    FADEGPU_F_FORCE_GENERIC = 1,
    FADEGPU_F_NO_SCATTER = 2,
    FADEGPU_F_TAGS_ONLY = 4,
    FADEGPU_F_NO_SHORTCUT = 8,
    FADEGPU_F_HOST_BINNING = 16,
    FADEGPU_F_SYNC_SUBMIT = 32
}

/// per-read result flags
enum : uint
{
    FADEGPU_R_ALIGNED = 1,
    FADEGPU_R_ART_LEFT = 2,   /// status.art_left,  analysis.d:82
    FADEGPU_R_ART_RIGHT = 4,  /// status.art_right, analysis.d:106
    FADEGPU_R_OPS_TRUNC = 8,
    FADEGPU_R_GENERIC = 16,
    FADEGPU_R_OVERSIZE = 32,
    FADEGPU_R_SCORE_ONLY = 64
}

/// one read of the compact input layout (32 bytes)
struct fadegpu_read_meta
{
    long pos;         /// rec.pos, 0-based
    uint seq_off;     /// byte offset of the read's bases inside seq4
    int l_qseq;       /// rec.length
    int tid;          /// rec.tid
    int aligned_len;  /// rec.cigar.alignedLength, analysis.d:53
    uint clip_left;   /// parse_clips(rec.cigar)[0].length
    uint clip_right;  /// parse_clips(rec.cigar)[1].length
}

struct fadegpu_batch_view
{
    long max_reads, max_seq_bytes;
    // inputs (seven-array layout, fadegpu_submit)
    ubyte* seq4;
    long* seq_off;
    int* l_qseq;
    int* tid;
    long* pos;
    int* aligned_len;
    int* clip_left;
    int* clip_right;
    // outputs
    ubyte* flags;
    int* score;
    int* beg_query;
    int* end_query;
    int* beg_ref;
    int* end_ref;
    long* win_start;
    int* n_ops;
    uint* ops;
    // inputs (compact layout, fadegpu_submit_compact)
    ubyte* gate;
    fadegpu_read_meta* meta;
}

struct fadegpu_stats
{
    long n_reads, n_aligned, n_generic, cells, h2d_bytes, d2h_bytes;
    int kernel_launches;
    float kernel_ms, fill_ms, trace_ms, generic_ms, total_ms;
    long scratch_bytes;
    float host_submit_ms, host_wait_ms, host_classify_ms, host_sort_ms, host_gather_ms;
    int host_threads;
    int reserved;
    long n_oversize;
}

struct fadegpu_inputs
{
    const(ubyte)* seq4;
    const(long)* seq_off;
    const(int)* l_qseq, tid;
    const(long)* pos;
    const(int)* aligned_len, clip_left, clip_right;
}

struct fadegpu_result
{
    int score;
    int end_query, end_ref;
    int beg_query, beg_ref;   /// beg_ref == res.position
    int n_ops;
    uint flags;
    int read;
    uint[FADEGPU_MAX_OPS] ops;
}

struct fadegpu_results_view
{
    long n_results;
    const(fadegpu_result)* results;
    const(long)* win_start;
    const(int)* result_index;
}

static assert(fadegpu_params.sizeof == 48);
static assert(fadegpu_read_meta.sizeof == 32);
static assert(fadegpu_batch_view.sizeof == 168);
static assert(fadegpu_result.sizeof == 72);
static assert(fadegpu_stats.sizeof == 120);

int fadegpu_abi_version();
int fadegpu_device_count(int* n);
int fadegpu_default_params(fadegpu_params* p);
int fadegpu_create(int device, const(fadegpu_params)* p, fadegpu_ctx** ctx);
void fadegpu_destroy(fadegpu_ctx* ctx);
const(char)* fadegpu_last_error(const(fadegpu_ctx)* ctx);
int fadegpu_load_reference(fadegpu_ctx* ctx, int n_contigs, const(char*)* names, const(long)* lengths, const(char*)* seqs);
int fadegpu_share_reference(fadegpu_ctx* dst, const(fadegpu_ctx)* src);
int fadegpu_reference_info(const(fadegpu_ctx)* ctx, int* n_contigs, long* total_bases, long* device_bytes);
int fadegpu_alloc_batch(fadegpu_ctx* ctx, long max_reads, long max_seq_bytes, fadegpu_batch** b);
int fadegpu_get_batch_view(fadegpu_batch* b, fadegpu_batch_view* v);
void fadegpu_free_batch(fadegpu_batch* b);
int fadegpu_submit(fadegpu_ctx* ctx, fadegpu_batch* b, long n_reads);
int fadegpu_submit_compact(fadegpu_ctx* ctx, fadegpu_batch* b, long n_reads, long seq_bytes);
int fadegpu_submit_inputs(fadegpu_ctx* ctx, fadegpu_batch* b, long n_reads, const(fadegpu_inputs)* inp);
int fadegpu_wait(fadegpu_ctx* ctx, fadegpu_batch* b);
int fadegpu_get_results(const(fadegpu_batch)* b, fadegpu_results_view* r);
int fadegpu_get_stats(const(fadegpu_batch)* b, fadegpu_stats* s);

// ---- include/fadehost.h: annotateTask's bookkeeping around the device call (anno.d:61-74, 94-107) ----
enum : ubyte
{
    FADE_RS_SC = 1,
    FADE_RS_ART_LEFT = 2,
    FADE_RS_ART_RIGHT = 4,
    FADE_RS_SUP = 32
}

struct fadehost_record
{
    int flag;            /// SAM FLAG
    int has_sa;          /// rec["SA"].exists
    const(uint)* cigar;  /// BAM-encoded ops
    int n_cigar;
    const(ubyte)* seq4;  /// BAM 4-bit packed bases
    const(ubyte)* qual;  /// raw phred values
    int l_qseq;
    int tid;
    long pos;            /// 0-based
}

void fadehost_parse_clips(const(uint)* cigar, int n_cigar, uint* clips);
long fadehost_aligned_length(const(uint)* cigar, int n_cigar);
int fadehost_prepare(const(fadehost_record)* rec, int* aligned_len, int* clip_left, int* clip_right, ubyte* rs_base);
int fadehost_finish(const(fadehost_record)* rec, const(char)* contig_name, ubyte rs_base, int clip_left, int clip_right,
        int aligned_len, ubyte flags, long win_start, int beg_ref, int n_ops, const(uint)* ops, ubyte* rs_out,
        char* am, char* as_, char* ar, char* ab, size_t cap);
