/*
 * fastq_oracle.h -- CPU ORACLE for the fastq-rs hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the algorithm in the reference crate
 * (aseyboldt/fastq-rs, crate `fastq` 0.6.0).  It is NOT part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it, and only as the checker / the timed CPU baseline.  The product path
 * (fastq_rs_b200/) never links, imports or executes anything in this directory.
 *
 * Parity pin: the restatement is checked against every unit test and the executable
 * doc-test the reference holds for this path (src/lib.rs:616-810, src/lib.rs:474-508);
 * see tests/golden/ and tests/test_oracle_golden.py.  The Rust crate itself cannot be
 * built in this environment (no rustc/cargo, deps not vendored), so oracle/_ref does
 * not exist.  The per-position histogram closure has no counterpart in the reference
 * (it is defined in SURVEY.md 8(a) row 15 on top of Record::seq()/qual()); for that
 * part the oracle is the definition and parity is "unpinned by the reference".
 *
 * Each function cites the reference file:line it follows.
 */
#ifndef FASTQ_ORACLE_H
#define FASTQ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status / error kinds (shared numbering with include/fastq_b200.h) */
enum {
    FQO_OK = 0,
    FQO_E_HEADER = 1,   /* "Fastq headers must start with '@'"        src/records.rs:145 */
    FQO_E_SEP = 2,      /* "Sequence and quality not separated by +"  src/records.rs:159 */
    FQO_E_LENGTH = 3,   /* "Sequence and quality length mismatch"     src/records.rs:236 */
    FQO_E_TOO_LONG = 4, /* "Fastq record is too long"                 src/lib.rs:281,401 */
    FQO_E_TRUNCATED = 5,/* "Possibly truncated input file"            src/lib.rs:289,409 */
    FQO_E_IO = 6
};

#define FQO_BUFSIZE_DEFAULT (68u * 1024u) /* src/lib.rs:129 */

/* ---- in-memory reader (stands in for std::io::Cursor, src/lib.rs:618 etc.) ---------- */
typedef struct {
    const uint8_t *data;
    size_t len;
    size_t pos;
    size_t max_read; /* 0 = no limit; else a read() returns at most this many bytes */
} fqo_reader;

/* ---- sliding byte buffer (src/buffer.rs:3-112) -------------------------------------- */
typedef struct {
    uint8_t *data;
    size_t cap;
    size_t start;
    size_t end;
} fqo_buffer;

/* ---- record index (src/records.rs:56-63): offsets of the four '\n', record-relative -- */
typedef struct {
    size_t head, seq, sep, qual;
    size_t data0, data1; /* byte range; data1 = qual + 1 */
} fqo_idx_record;

/* borrowed view (src/records.rs:36-45) */
typedef struct {
    const uint8_t *data; /* record bytes, data[0] == '@' */
    size_t len;
    size_t head, seq, sep, qual;
} fqo_ref_record;

enum { FQO_RES_RECORD = 0, FQO_RES_INCOMPLETE = 1, FQO_RES_EMPTY = 2 };

/* src/records.rs:201-247.  Returns FQO_OK and *kind, or an FQO_E_* grammar error. */
int fqo_from_buffer(const uint8_t *buf, size_t len, int *kind, fqo_idx_record *out);

/* Record accessors with one trailing '\r' removed (src/records.rs:65-90). */
void fqo_rec_head(const fqo_ref_record *r, const uint8_t **p, size_t *n);
void fqo_rec_seq(const fqo_ref_record *r, const uint8_t **p, size_t *n);
void fqo_rec_qual(const fqo_ref_record *r, const uint8_t **p, size_t *n);
void fqo_rec_sepline(const fqo_ref_record *r, const uint8_t **p, size_t *n); /* src/records.rs:172 */
int fqo_validate_dna(const fqo_ref_record *r);  /* src/records.rs:19-23 */
int fqo_validate_dnan(const fqo_ref_record *r); /* src/records.rs:29-33 */

/* ---- Parser::each (src/lib.rs:221-238 over RecordRefIter::advance, :255-303) -------- */
/* callback returns nonzero to continue, 0 to stop (closure -> bool) */
typedef int (*fqo_each_fn)(void *user, const fqo_ref_record *rec, uint64_t stream_offset);

typedef struct {
    int status;           /* FQO_OK or FQO_E_* */
    int finished;         /* 1 = Ok(true) reached EOF, 0 = Ok(false) closure stopped */
    uint64_t n_delivered; /* records handed to the closure */
    uint64_t err_offset;  /* stream offset of the record at which the error was raised */
} fqo_each_result;

void fqo_each(fqo_reader *rd, size_t bufsize, fqo_each_fn fn, void *user, fqo_each_result *res);

/* ---- stats closure (built on seq()/qual(), src/records.rs:82-90) */
typedef struct {
    uint32_t max_len;    /* P */
    uint64_t n_records;
    uint64_t n_bases;
    uint64_t clip_seq;   /* seq bytes at positions >= P  */
    uint64_t clip_qual;  /* qual bytes at positions >= P */
    uint64_t *base_hist; /* [P][6]   A C G T N other */
    uint64_t *qual_hist; /* [P][256] raw byte        */
    uint64_t *len_hist;  /* [P+2]    index min(len, P+1) */
} fqo_stats;

fqo_stats *fqo_stats_new(uint32_t max_len);
void fqo_stats_free(fqo_stats *s);
void fqo_stats_clear(fqo_stats *s);
void fqo_stats_add(fqo_stats *dst, const fqo_stats *src);
void fqo_stats_record(fqo_stats *s, const fqo_ref_record *r);

/* each() + stats closure over a memory block */
void fqo_each_stats(const uint8_t *data, size_t len, size_t bufsize, size_t max_read,
                    fqo_stats *stats, fqo_each_result *res);

/* each() collecting the absolute offsets of every delivered record:
 * out[5*i+0..4] = record start, head '\n', seq '\n', sep '\n', qual '\n' (stream offsets).
 * At most cap records are stored; res->n_delivered is always the true count. */
/* each() + a closure writing the records that pass validate_dna (1) / validate_dnan (2) / all (0)
 * verbatim (src/records.rs:19-33, 93-96) */
void fqo_each_filter(const uint8_t *data, size_t len, size_t bufsize, size_t max_read, int mode,
                     uint8_t *out, size_t cap, uint64_t *n_kept, uint64_t *n_bytes, fqo_each_result *res);
void fqo_each_index(const uint8_t *data, size_t len, size_t bufsize, size_t max_read,
                    uint64_t *out, size_t cap, fqo_each_result *res);

/* ---- record_sets / parallel_each (src/lib.rs:306-566) -------------------------------- */
typedef struct {
    uint8_t *buffer; /* owned, bufsize bytes */
    size_t bufsize;
    fqo_idx_record *records; /* data0/data1 absolute in buffer (src/lib.rs:417-418) */
    size_t n_records;
    size_t cap_records;
} fqo_record_set;

void fqo_record_set_free(fqo_record_set *s);
void fqo_record_set_get(const fqo_record_set *s, size_t i, fqo_ref_record *out);

typedef struct fqo_set_iter fqo_set_iter;
fqo_set_iter *fqo_record_sets_new(fqo_reader *rd, size_t bufsize);
/* returns 1 + *set (caller frees), 0 at end, or -(FQO_E_*) on error (the partially built
 * set is dropped, src/lib.rs:375,399-410) */
int fqo_record_sets_next(fqo_set_iter *it, fqo_record_set **set);
void fqo_record_sets_free(fqo_set_iter *it);

/* parallel_each(n_threads, closure) with the stats closure run by every worker on every
 * record of every set it receives; per-worker stats are summed into `total` after join.
 * sets_per_worker[i] receives the number of sets worker i consumed (may be NULL). */
int fqo_parallel_each_stats(const uint8_t *data, size_t len, size_t bufsize, size_t max_read,
                            int n_threads, fqo_stats *total, uint64_t *sets_per_worker);

/* parallel_each counting records only (examples/fastq-count-thread.rs:15-25) */
int fqo_parallel_each_count(const uint8_t *data, size_t len, size_t bufsize, size_t max_read,
                            int n_threads, uint64_t *n_records);

/* ---- synthetic FASTQ; byte-exact twin of the device generator K0 ---- */
#define FQO_SYNTH_SEED 0xFA57A11CE5EED001ull
uint64_t fqo_splitmix64(uint64_t x);
uint32_t fqo_synth_var_len(uint64_t seed, uint64_t rec);       /* 50..300 */
size_t fqo_synth_fixed_record_bytes(uint32_t L);               /* 17 + 2(L+1) + 2 */
/* bytes [byte_off, byte_off + n) of the infinite fixed-length stream */
void fqo_synth_fixed(uint64_t seed, uint32_t L, uint64_t byte_off, size_t n, uint8_t *out);
/* records [first, first+count) of the variable-length stream; returns bytes written
 * (call with out == NULL to get the size) */
size_t fqo_synth_var(uint64_t seed, uint64_t first, uint64_t count, uint8_t *out);
/* one record (fixed L or variable when L == 0) rendered at out; returns its size */
size_t fqo_synth_record(uint64_t seed, uint64_t rec, uint32_t L, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
