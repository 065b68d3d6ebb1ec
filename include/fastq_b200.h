/*
 * fastq_b200.h -- C ABI of the B200-native FASTQ delimit + per-record statistics engine.
 *
 * This is the drop-in boundary for ONE path of the `fastq` crate (aseyboldt/fastq-rs 0.6.0):
 * record delimiting with '@'/'+' validation over raw bytes, plus the per-position base and
 * quality histograms a stats closure would compute over Record::seq()/qual().  Everything
 * here is `extern "C"`, plain pointers and sizes -- what a Rust `extern "C"` block (or cgo /
 * ctypes) binds.  No callbacks cross this boundary: closures run on the caller's side over
 * the record index this library returns (see INTEGRATION.md and rust/).
 *
 * Each entry point cites the reference interface it replaces (paths relative to the crate).
 * The CUDA extension is the only implementation: there is no CPU fallback behind this ABI.
 */
#ifndef FASTQ_B200_H
#define FASTQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FQB_ABI_VERSION 3u

/* ---- status codes --------------------------------------------------------------------
 * 1..5 are the reference's grammar errors; the host shim maps each to
 * io::Error::new(InvalidData, <message>) with the exact reference text. */
enum {
    FQB_OK = 0,
    FQB_E_HEADER = 1,    /* "Fastq headers must start with '@'"        src/records.rs:143-146 */
    FQB_E_SEP = 2,       /* "Sequence and quality not separated by +"  src/records.rs:157-160 */
    FQB_E_LENGTH = 3,    /* "Sequence and quality length mismatch"     src/records.rs:233-238 */
    FQB_E_TOO_LONG = 4,  /* "Fastq record is too long"                 src/lib.rs:278-283     */
    FQB_E_TRUNCATED = 5, /* "Possibly truncated input file"            src/lib.rs:286-291     */
    FQB_E_IO = 6,        /* reader error passed through                src/buffer.rs:86-96    */
    FQB_E_PHASE = 7,     /* FQB_F_INFER_START: the first record of the shard could not be inferred
                            from its bytes; parse again with the exact line_base              */
    FQB_E_RETRY = 8,     /* only ever seen in fqb_device_result / the outcome slots, never returned by fqb_fetch: a
                            record in the middle of the shard is bad, every record in front of it is verified, but
                            the device-resident numbers still include records behind it.  fqb_fetch completes the
                            parse (the bytes in front of the bad record once more, then its classification);
                            callers that read outcomes on the device call it when they see this status. */
    FQB_E_ARG = 50,      /* bad argument (null pointer, misaligned buffer, ...) */
    FQB_E_STATE = 51,    /* call out of order (e.g. submit without acquire) */
    FQB_E_NOMEM = 52,
    FQB_E_CANCELLED = 53, /* batch mode: fqb_batch_cancel was called (the closure returned false) */
    FQB_E_CUDA = 100,    /* CUDA runtime failure; fqb_last_error() has the text */
    FQB_E_NCCL = 101     /* NCCL missing or failing (multi-rank entry points only); fqb_last_error() */
};

/* The reference refuses records that do not fit its 68 KiB window (src/lib.rs:129,278-283).  Whether a
 * record of 69 617 .. 69 632 bytes fits depends on where Buffer::clean parks the incomplete record
 * (src/buffer.rs:51-72: so that the next read is 16-byte aligned).  With a reader that fills every read()
 * (Cursor, File, &[u8] -- the readers of the reference's own tests) every refill ends on a 16-byte boundary
 * of the stream, the record is parked at buffer offset (stream offset mod 16), and the rule is exact:
 *     a record starting at stream offset p fits  iff  (p mod 16) + length <= FQB_MAX_RECORD_BYTES;
 *     an incomplete record is FQB_E_TOO_LONG iff the stream holds >= FQB_MAX_RECORD_BYTES - (p mod 16)
 *     bytes from p on, else FQB_E_TRUNCATED.
 * That is the rule implemented here (tests/test_gpu_parity.py::test_too_long_band checks it against the
 * oracle's restatement of Buffer); the stream offsets of the shards / chunks therefore matter mod 16. */
#define FQB_MAX_RECORD_BYTES (68u * 1024u)

/* ---- flags for fqb_shard.flags -------------------------------------------------------- */
#define FQB_F_HIST        0x01u /* accumulate the per-position histograms (stats closure)   */
#define FQB_F_INDEX       0x02u /* write the line-end index (record offsets)                */
#define FQB_F_LINE_START  0x04u /* byte 0 of the buffer is the first byte of a line (stream
                                   start, or the byte before it is '\n')                   */
#define FQB_F_EOF         0x08u /* no byte of the stream exists beyond n_avail              */
#define FQB_F_FRONT16     0x10u /* d_bytes[-16..0) is readable and holds the 16 stream bytes
                                   before the shard (lets the kernel see whether the shard
                                   starts right after a '\n'); excludes FQB_F_LINE_START    */
#define FQB_F_INFER_START 0x20u /* line_base is not known yet (a later shard of a multi-GPU job):
                                   infer where the first record of the shard starts from the
                                   grammar of the following records; fqb_result.line_phase reports
                                   the line_base mod 4 that start implies.  The caller checks it
                                   against the exact prefix of the shards' n_lines (which does not
                                   depend on the phase) and parses again with the exact line_base
                                   on a mismatch or on FQB_E_PHASE.                             */

#define FQB_F_PARTIAL     0x40u /* fqb_parse_host only: more bytes of the stream follow the bytes of this
                                   call (a refill is pending, src/lib.rs:262-294): a trailing record
                                   that is incomplete -- and shorter than FQB_MAX_RECORD_BYTES -- is
                                   not an error; fqb_result.tail_offset reports where it starts so
                                   that the caller carries it over in front of the next bytes, as
                                   Buffer::clean does (src/buffer.rs:51-72)                      */

typedef struct fqb_ctx fqb_ctx;

typedef struct {
    uint32_t abi_version;   /* FQB_ABI_VERSION */
    int32_t device;         /* CUDA device ordinal */
    uint32_t max_len;       /* P: positions tracked per read (150, 300, ...), 1..4096 */
    uint32_t reserved0;
    uint64_t slot_bytes;    /* streaming: bytes per pinned/device ring slot (0 = 64 MiB) */
    uint32_t n_slots;       /* streaming: ring depth (0 = 4); mirrors thread_reader's queuelen
                               (src/thread_reader.rs:13-32) */
    uint32_t reserved1;
} fqb_config;

/* One contiguous piece of a byte stream resident in device memory.
 * Records are OWNED by the shard in which their first byte lies ([0, n_own)); their tails
 * may extend into [n_own, n_avail) -- the device-side counterpart of the carry-over that
 * Buffer::clean / replace_buffer perform (src/buffer.rs:30-72). */
typedef struct {
    const uint8_t *d_bytes; /* device pointer, 16-byte aligned */
    uint64_t n_own;
    uint64_t n_avail;       /* >= n_own; bytes readable */
    uint64_t stream_offset; /* stream offset of d_bytes[0] (reported in err_offset, index) */
    uint64_t line_base;     /* number of '\n' in the stream before d_bytes[0]: fixes which
                               lines are headers (strict 4-line cadence, src/records.rs:201-247) */
    uint32_t flags;         /* FQB_F_* */
    uint32_t reserved;
    uint32_t *d_index;      /* FQB_F_INDEX: device array receiving, for the i-th '\n' of
                               [0, n_own), the low 32 bits of its stream offset (the four line
                               ends of IdxRecord, src/records.rs:56-63, laid out densely:
                               record k of the shard = entries 4k'..4k'+3 after the phase) */
    uint64_t index_cap;     /* entries available in d_index */
} fqb_shard;

/* Outcome of a parse: what Parser::each returns (src/lib.rs:221-238) plus where it stopped. */
typedef struct {
    int32_t status;        /* FQB_OK or FQB_E_* of the FIRST bad record in stream order */
    int32_t finished;      /* 1 = reached the end of the shard/stream cleanly */
    uint64_t n_records;    /* records delivered = all records before the first bad one */
    uint64_t n_lines;      /* '\n' bytes in the owned range */
    uint64_t err_offset;   /* stream offset of the record at which the error was raised */
    uint64_t tail_offset;  /* stream offset of the first owned record that is incomplete within
                              n_avail without being an error (only without FQB_F_EOF);
                              UINT64_MAX if none */
    uint32_t line_phase;   /* line_base mod 4 implied by where the first record of the shard was
                              found (== line_base & 3 unless FQB_F_INFER_START) */
    uint32_t reserved;
} fqb_result;

/* Statistics block: a flat array of uint64_t, fqb_stats_words(P) long, laid out
 *   [0] n_records  [1] n_bases  [2] clip_seq  [3] clip_qual  [4..7] reserved
 *   [8 .. 8+P+2)                len_hist[min(len, P+1)]
 *   [.. + 6P)                   base_hist[pos][A,C,G,T,N,other]
 *   [.. + 256P)                 qual_hist[pos][raw byte]
 * computed over Record::seq()/qual() (src/records.rs:82-90: one trailing '\r' trimmed),
 * base classes as validate_dnan's alphabet (src/records.rs:29-33).  One block = one
 * ncclAllReduce(sum, u64) payload. */
#define FQB_STATS_HDR 8u
size_t fqb_stats_words(uint32_t max_len);
size_t fqb_stats_len_hist_off(uint32_t max_len);
size_t fqb_stats_base_hist_off(uint32_t max_len);
size_t fqb_stats_qual_hist_off(uint32_t max_len);

/* ---- lifecycle ------------------------------------------------------------------------ */
/* replaces Parser::new (src/lib.rs:200-205): owns all device/pinned resources */
int fqb_create(const fqb_config *cfg, fqb_ctx **out);
void fqb_destroy(fqb_ctx *ctx);
const char *fqb_strerror(int status);     /* reference message for 1..5 */
const char *fqb_last_error(fqb_ctx *ctx); /* detail for FQB_E_CUDA etc. */
uint32_t fqb_abi_version(void);

/* ---- in-HBM path (bytes already resident on the device) --------------------------------
 * replaces Parser::each + the stats closure over one shard (src/lib.rs:221-238,
 * src/records.rs:201-247).  Asynchronous on `stream` (a cudaStream_t, NULL = default). */
int fqb_parse_device(fqb_ctx *ctx, const fqb_shard *shard, void *stream);
/* count '\n' in d_bytes[0..n) (the line-phase exchange between shards; also what
 * `wc -l` measures, README.md:41).  Asynchronous; result via fqb_fetch_line_count. */
int fqb_count_lines_device(fqb_ctx *ctx, const uint8_t *d_bytes, uint64_t n, void *stream);
int fqb_fetch_line_count(fqb_ctx *ctx, void *stream, uint64_t *n_lines);
/* wait for `stream`, copy out the outcome and (if host_stats != NULL) the stats block */
int fqb_fetch(fqb_ctx *ctx, void *stream, fqb_result *res, uint64_t *host_stats);
/* device address of the stats block of the last parse (for an in-place allreduce) */
uint64_t *fqb_device_stats(fqb_ctx *ctx);
/* device address of the outcome of the last fqb_parse_device as 8 words, valid once the kernels of
 * that parse have run (stream order): [0] status [1] finished [2] n_records [3] n_lines
 * [4] err_offset [5] tail_offset (UINT64_MAX if none) [6] line_phase [7] reserved -- lets an N-rank
 * driver all-gather the outcomes without a host round trip */
uint64_t *fqb_device_result(fqb_ctx *ctx);
/* which path produced the result the last fqb_fetch returned (diagnostics; tests use it to make sure clean
 * inputs stay on the fast path): out[0] = 1 if the speculative pass was abandoned and the exact kernel redid the
 * shard, out[1] = windows the speculative kernel predicted, out[2] = windows it scanned */
int fqb_last_path(fqb_ctx *ctx, uint64_t out[3]);
/* number of kernels this library launched on ctx so far (bench accounting) */
uint64_t fqb_launch_count(fqb_ctx *ctx);
/* CUDA-event time of the main scan kernel of the last fqb_parse_device call, in ms
 * (waits for it); <0 on error */
float fqb_last_scan_ms(fqb_ctx *ctx);
/* the same for the kernel that turns the scan kernel's staged line ends / window descriptors into the dense
 * line-end index (0 if the last call wrote no index through it) */
float fqb_last_index_ms(fqb_ctx *ctx);

/* ---- record filter: the step after the path (validate_dna / validate_dnan + Record::write) --------
 * Keeps the records whose seq() passes the predicate and writes their raw bytes ('@' .. final '\n',
 * RefRecord::write, src/records.rs:93-96) to d_out densely and in stream order.
 *   d_bytes / stream_offset : the shard that was parsed (same pointer and stream_offset)
 *   d_index / n_records     : the line-end index fqb_parse_device wrote, starting at the first line end
 *                             of record 0, and the number of records to consider (fqb_result.n_records)
 *   first_offset            : stream offset of the first byte of record 0 (= stream_offset for a shard
 *                             that starts at a record start)
 * Asynchronous on `stream`; fqb_fetch_filter waits and returns the totals.  Records that do not fit in
 * out_cap are not written (out_bytes > out_cap tells). */
#define FQB_KEEP_ALL  0u
#define FQB_KEEP_DNA  1u /* Record::validate_dna,  src/records.rs:19-23: seq() only A C T G       */
#define FQB_KEEP_DNAN 2u /* Record::validate_dnan, src/records.rs:29-33: seq() only A C T G N     */
int fqb_filter_device(fqb_ctx *ctx, const uint8_t *d_bytes, uint64_t stream_offset, const uint32_t *d_index,
                      uint64_t n_records, uint64_t first_offset, uint32_t mode, uint8_t *d_out,
                      uint64_t out_cap, void *stream);
int fqb_fetch_filter(fqb_ctx *ctx, void *stream, uint64_t *n_kept, uint64_t *out_bytes);

/* ---- host path: bytes in host memory, staged through the pinned ring --------------------
 * replaces Parser::new(reader).each(stats closure) end to end.  Synchronous.
 * host_index (optional): receives the low 32 bits of the stream offset of every '\n'
 * before the first bad record, up to index_cap entries; *n_index = entries written.  A host_index in
 * pinned memory (fqb_host_alloc) is written by the device in stream order, with no synchronisation
 * per chunk; pageable memory costs one per chunk.
 * stream_offset: stream offset of bytes[0]; err_offset, tail_offset and the index are stream offsets.
 * flags: FQB_F_HIST | FQB_F_INDEX | FQB_F_PARTIAL.  With FQB_F_PARTIAL the call is one refill of a
 * longer stream (bounded-memory each()/record_sets(): src/lib.rs:262-294, 364-425): `bytes` must
 * start at a record start. */
int fqb_parse_host(fqb_ctx *ctx, const uint8_t *bytes, uint64_t n, uint64_t stream_offset, uint32_t flags,
                   fqb_result *res, uint64_t *host_stats,
                   uint32_t *host_index, uint64_t index_cap, uint64_t *n_index);

/* ---- streaming ring: the thread_reader protocol on pinned slots --------------------------
 * (src/thread_reader.rs:13-50: `empty` -> reader fills -> `full` -> consumer -> `empty`)   */
int fqb_stream_begin(fqb_ctx *ctx, uint32_t flags);
/* blocks until a pinned slot is free (empty_recv.recv(), src/thread_reader.rs:44) */
int fqb_stream_acquire(fqb_ctx *ctx, uint8_t **pinned, uint64_t *cap);
/* hand the filled slot over (full_send.send(), src/thread_reader.rs:46): enqueues the
 * H2D copy on the copy stream and the kernels on the compute stream */
int fqb_stream_submit(fqb_ctx *ctx, uint64_t n_valid);
/* end of input: drains the ring, returns outcome + stats */
int fqb_stream_finish(fqb_ctx *ctx, fqb_result *res, uint64_t *host_stats);

/* ---- batch mode: the generic-closure path, asynchronous --------------------------------------------
 * replaces RecordSetIter::next / RecordSet (src/lib.rs:306-426) and what parallel_each's producer loop does with
 * them (src/lib.rs:509-566): the GPU delimits chunk after chunk while the caller's closures run over the
 * chunks already delimited.
 *   producer thread: fqb_batch_begin; then, per read() of the input, fqb_stream_acquire -> fill -> fqb_stream_submit
 *                    (the thread_reader protocol, as in the streaming ring); fqb_batch_close at the end of the input
 *   consumer thread: fqb_next_batch -> the records of one chunk: their bytes (borrowed from the pinned host ring:
 *                    valid until fqb_release_batch, the lifetime rule of RecordSet's own buffer) and, per record,
 *                    the four line ends of IdxRecord (src/records.rs:56-63) -- low 32 bits of stream offsets;
 *                    offset within `bytes` = (uint32_t)(line_end - (uint32_t)stream_offset).
 * The shim deals sub-ranges of a batch to its worker threads exactly as src/lib.rs:535 deals RecordSets.
 * At most n_slots - 3 batches may be held (not yet released) at a time; fqb_config.n_slots >= 4.
 * A batch with status != FQB_OK is the last one: its n_records records precede the bad record (each() delivers
 * them, src/lib.rs:226-237), err_offset is the stream offset of the bad one. */
typedef struct {
    const uint8_t *bytes;      /* first byte of the first record of the batch (pinned host memory) */
    uint64_t n_bytes;          /* bytes[0 .. n_bytes) = the records, back to back */
    uint64_t n_avail;          /* >= n_bytes: stream bytes readable from `bytes` on (what follows the records: the
                                  bad record of a batch with status != FQB_OK, for diagnostics) */
    uint64_t stream_offset;    /* stream offset of bytes[0] */
    const uint32_t *line_ends; /* 4 x n_records entries */
    uint64_t n_records;
    uint64_t first_record;     /* records handed out before this batch */
    uint64_t err_offset;
    uint64_t token;            /* for fqb_release_batch (UINT64_MAX: an empty end-of-input marker, nothing to release) */
    int32_t status;
    int32_t last;              /* 1 = no batch follows (end of input, or status != FQB_OK) */
} fqb_batch;
int fqb_batch_begin(fqb_ctx *ctx, uint32_t flags);    /* flags: 0 or FQB_F_HIST (statistics alongside, via fqb_batch_end) */
int fqb_batch_close(fqb_ctx *ctx);                    /* producer: end of input */
int fqb_next_batch(fqb_ctx *ctx, fqb_batch *out);     /* consumer: blocks until the next batch is ready */
int fqb_release_batch(fqb_ctx *ctx, uint64_t token);  /* consumer (any thread): RecordSet dropped */
int fqb_batch_cancel(fqb_ctx *ctx);                   /* either side: stop early; blocked calls return FQB_E_CANCELLED */
int fqb_batch_end(fqb_ctx *ctx, fqb_result *res);     /* after the last batch / cancel, producer thread joined */

/* ---- N ranks, one byte shard per GPU: the one collective of the path ---------------------------
 * The reference has nothing to shard (parallel_each delimits on ONE thread, src/lib.rs:535); this is the
 * B200-side answer to "the byte stream shards by chunk across the GPUs of one box" (SURVEY.md 8(e)).
 * One process per GPU.  Every rank parses its shard at once (fqb_parse_device with FQB_F_INFER_START for all
 * shards but the first), then ONE all-reduce(sum, u64) over NVLink combines
 *     [ statistics block | world x 8 outcome words ]
 * -- every rank has written its outcome (the words of fqb_device_result) into its own slot and zeros into the
 * others, so the sum of the slots is the gather of the outcomes.  Every rank then holds the global statistics
 * and all outcomes: exact line numbers = prefix of the n_lines, first error in stream order, and whether every
 * inferred shard start was right (line_phase == prefix & 3) -- if not, the caller parses those shards again
 * with the exact line_base and reduces once more (never observed on well-formed input).
 * NCCL is bound at run time (dlopen of libnccl.so.2): none is needed for world == 1. */
#define FQB_MAX_WORLD 64
#define FQB_COMM_ID_BYTES 128
/* rank 0: a fresh ncclUniqueId, to be handed to the other ranks by whatever the caller has (a file, MPI, ...) */
int fqb_comm_unique_id(uint8_t out[FQB_COMM_ID_BYTES]);
/* collective over all ranks (ncclCommInitRank); world == 1 needs no id and no NCCL */
int fqb_comm_init(fqb_ctx *ctx, int rank, int world, const uint8_t id[FQB_COMM_ID_BYTES]);
int fqb_comm_destroy(fqb_ctx *ctx);
int fqb_comm_rank(fqb_ctx *ctx);
int fqb_comm_world(fqb_ctx *ctx);
/* enqueue the all-reduce of the last parse's [block | slots] on `stream` (after the parse, stream order) */
int fqb_allreduce(fqb_ctx *ctx, void *stream);
/* wait for it; host_stats (fqb_stats_words(P) words, may be NULL) = the global block, outcomes
 * (8 x world words, may be NULL) = per rank: status, finished, n_records, n_lines, err_offset, tail_offset,
 * line_phase, 0.  One device-to-host copy, one synchronisation. */
int fqb_fetch_reduced(fqb_ctx *ctx, void *stream, uint64_t *host_stats, uint64_t *outcomes);
/* the send buffer itself -- fqb_exchange_words(ctx) = fqb_stats_words(P) + 8 * world device-resident words --
 * for callers that run the collective with their own library */
uint64_t *fqb_device_exchange(fqb_ctx *ctx);
size_t fqb_exchange_words(fqb_ctx *ctx);

/* ---- pinned host memory helpers ------------------------------------------------------- */
int fqb_host_alloc(uint64_t bytes, void **out); /* cudaHostAlloc */
void fqb_host_free(void *p);

/* ---- synthetic FASTQ; device twin of the oracle generator -------------- */
#define FQB_SYNTH_SEED 0xFA57A11CE5EED001ull
/* bytes [byte_off, byte_off+n) of the infinite fixed-length stream (read length L) */
int fqb_synth_fixed_device(uint8_t *d_out, uint64_t n, uint64_t byte_off, uint32_t L,
                           uint64_t seed, void *stream);
/* variable-length records [first, first+count); d_rec_off[i] = byte offset of record
 * first+i relative to d_out (count+1 entries, exclusive prefix sum of record sizes) */
int fqb_synth_var_device(uint8_t *d_out, const uint64_t *d_rec_off, uint64_t first,
                         uint64_t count, uint64_t seed, void *stream);
/* record sizes (bytes) of variable-length records [first, first+count) into d_sizes */
int fqb_synth_var_sizes_device(uint64_t *d_sizes, uint64_t first, uint64_t count,
                               uint64_t seed, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FASTQ_B200_H */
