/*
 * fastq_oracle.c -- CPU ORACLE (test infrastructure, never shipped, never on the product
 * path).  Plain-C restatement of aseyboldt/fastq-rs' record delimiting, the each /
 * record_sets / parallel_each drivers and the sliding buffer, plus the stats closure and
 * the synthetic generators of SURVEY.md 8(a)/8(d).  See fastq_oracle.h for the parity pin.
 */
#define _GNU_SOURCE
#include "fastq_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================================
 * reader: std::io::Cursor stand-in.  max_read lets tests emulate readers that return
 * short reads (results must not depend on it, SURVEY 8(a) "independent of ... chunks").
 * ==================================================================================== */
static size_t reader_read(fqo_reader *rd, uint8_t *dst, size_t want)
{
    size_t left = rd->len - rd->pos;
    size_t n = want < left ? want : left;
    if (rd->max_read && n > rd->max_read)
        n = rd->max_read;
    if (n)
        memcpy(dst, rd->data + rd->pos, n);
    rd->pos += n;
    return n;
}

/* ======================================================================================
 * Buffer (src/buffer.rs)
 * ==================================================================================== */
static int buffer_init(fqo_buffer *b, size_t size) /* Buffer::new, src/buffer.rs:10-16 */
{
    b->data = (uint8_t *)calloc(size ? size : 1, 1);
    b->cap = size;
    b->start = b->end = 0;
    return b->data != NULL;
}

static size_t buffer_len(const fqo_buffer *b) { return b->end - b->start; }   /* :18-20 */
static size_t buffer_free(const fqo_buffer *b) { return b->cap - b->end; }    /* :22-24 */

/* Where leftover bytes go so that the NEXT read lands on a 16-byte boundary
 * (src/buffer.rs:32-33 and :57-58). */
static void aligned_tail_slot(size_t n_left, size_t *new_start, size_t *new_end)
{
    *new_end = (n_left + 15) & ~(size_t)15;
    *new_start = *new_end - n_left;
}

/* Buffer::clean, src/buffer.rs:51-72 */
static void buffer_clean(fqo_buffer *b)
{
    size_t n, ns, ne;
    if (b->start == 0)
        return;
    n = buffer_len(b);
    aligned_tail_slot(n, &ns, &ne);
    if (ns >= b->start)
        return;
    memmove(b->data + ns, b->data + b->start, n);
    b->start = ns;
    b->end = ne;
}

/* Buffer::replace_buffer, src/buffer.rs:30-48: leftover moves into `fresh`, the old box is
 * handed back to the caller. */
static uint8_t *buffer_replace(fqo_buffer *b, uint8_t *fresh)
{
    size_t n = buffer_len(b), ns, ne;
    uint8_t *old = b->data;
    aligned_tail_slot(n, &ns, &ne);
    /* assert!(buffer.len() >= new_end): with cap a multiple of 16 this always holds */
    memcpy(fresh + ns, old + b->start, n);
    b->data = fresh;
    b->start = ns;
    b->end = ne;
    return old;
}

/* Buffer::read_into, src/buffer.rs:74-100 (the Interrupted retry has no analogue for a
 * memory reader). */
static size_t buffer_read_into(fqo_buffer *b, fqo_reader *rd)
{
    size_t n_free = buffer_free(b);
    size_t want = n_free < 4096 ? n_free : n_free - n_free % 4096;
    size_t got = reader_read(rd, b->data + b->end, want);
    b->end += got;
    return got;
}

/* ======================================================================================
 * Record grammar (src/records.rs)
 * ==================================================================================== */
/* read_header :137-149 / read_sep :151-163 share a shape: empty -> "none"; wrong first
 * byte -> error; else position of the first '\n' (or "none").  Returns 1 found, 0 none,
 * -1 grammar error. */
static int line_after_marker(const uint8_t *p, size_t n, uint8_t marker, size_t *nl)
{
    const uint8_t *q;
    if (n == 0)
        return 0;
    if (p[0] != marker)
        return -1;
    q = (const uint8_t *)memchr(p, '\n', n);
    if (!q)
        return 0;
    *nl = (size_t)(q - p);
    return 1;
}

int fqo_from_buffer(const uint8_t *buf, size_t len, int *kind, fqo_idx_record *out)
{
    size_t head_end, seq_end, sep_end, qual_end, pos, rel;
    const uint8_t *q;
    int r;

    if (len == 0) { /* :203-205 */
        *kind = FQO_RES_EMPTY;
        return FQO_OK;
    }
    r = line_after_marker(buf, len, '@', &head_end); /* :207-210 */
    if (r < 0)
        return FQO_E_HEADER;
    if (r == 0) {
        *kind = FQO_RES_INCOMPLETE;
        return FQO_OK;
    }
    pos = head_end + 1;

    q = (const uint8_t *)memchr(buf + pos, '\n', len - pos); /* :213-217 */
    if (!q) {
        *kind = FQO_RES_INCOMPLETE;
        return FQO_OK;
    }
    seq_end = (size_t)(q - buf);
    pos = seq_end + 1;

    r = line_after_marker(buf + pos, len - pos, '+', &rel); /* :220-224 */
    if (r < 0)
        return FQO_E_SEP;
    if (r == 0) {
        *kind = FQO_RES_INCOMPLETE;
        return FQO_OK;
    }
    sep_end = pos + rel;
    pos = sep_end + 1;

    q = (const uint8_t *)memchr(buf + pos, '\n', len - pos); /* :227-231 */
    if (!q) {
        *kind = FQO_RES_INCOMPLETE;
        return FQO_OK;
    }
    qual_end = (size_t)(q - buf);

    if (qual_end - sep_end != seq_end - head_end) /* :233-238, RAW lengths */
        return FQO_E_LENGTH;

    out->data0 = 0; /* :240-246 */
    out->data1 = qual_end + 1;
    out->head = head_end;
    out->seq = seq_end;
    out->sep = sep_end;
    out->qual = qual_end;
    *kind = FQO_RES_RECORD;
    return FQO_OK;
}

/* trim_winline, src/records.rs:65-73 */
static void trim_cr(const uint8_t *p, size_t n, const uint8_t **op, size_t *on)
{
    if (n && p[n - 1] == '\r')
        n--;
    *op = p;
    *on = n;
}

void fqo_rec_head(const fqo_ref_record *r, const uint8_t **p, size_t *n) /* :77-80 */
{
    trim_cr(r->data + 1, r->head - 1, p, n);
}
void fqo_rec_seq(const fqo_ref_record *r, const uint8_t **p, size_t *n) /* :83-85 */
{
    trim_cr(r->data + r->head + 1, r->seq - r->head - 1, p, n);
}
void fqo_rec_qual(const fqo_ref_record *r, const uint8_t **p, size_t *n) /* :88-90 */
{
    trim_cr(r->data + r->sep + 1, r->qual - r->sep - 1, p, n);
}
void fqo_rec_sepline(const fqo_ref_record *r, const uint8_t **p, size_t *n) /* :172 */
{
    trim_cr(r->data + r->seq + 1, r->sep - r->seq - 1, p, n);
}

int fqo_validate_dna(const fqo_ref_record *r) /* :19-23 */
{
    const uint8_t *p;
    size_t n, i;
    fqo_rec_seq(r, &p, &n);
    for (i = 0; i < n; i++)
        if (!(p[i] == 'A' || p[i] == 'C' || p[i] == 'T' || p[i] == 'G'))
            return 0;
    return 1;
}

int fqo_validate_dnan(const fqo_ref_record *r) /* :29-33 */
{
    const uint8_t *p;
    size_t n, i;
    fqo_rec_seq(r, &p, &n);
    for (i = 0; i < n; i++)
        if (!(p[i] == 'A' || p[i] == 'C' || p[i] == 'T' || p[i] == 'G' || p[i] == 'N'))
            return 0;
    return 1;
}

/* IdxRecord::to_ref_record, src/records.rs:178-199 */
static void to_ref_record(const fqo_idx_record *ix, const uint8_t *buffer, fqo_ref_record *out)
{
    out->data = buffer + ix->data0;
    out->len = ix->data1 - ix->data0;
    out->head = ix->head;
    out->seq = ix->seq;
    out->sep = ix->sep;
    out->qual = ix->qual;
}

/* ======================================================================================
 * Parser::each  (src/lib.rs:221-238) over RecordRefIter::advance (src/lib.rs:255-303)
 * ==================================================================================== */
void fqo_each(fqo_reader *rd, size_t bufsize, fqo_each_fn fn, void *user, fqo_each_result *res)
{
    fqo_buffer b;
    size_t pending = 0;   /* current_length, :258-260 */
    uint64_t consumed = 0;/* stream offset of buffer.start */
    memset(res, 0, sizeof *res);
    if (!buffer_init(&b, bufsize)) {
        res->status = FQO_E_IO;
        return;
    }
    for (;;) {
        /* ---- advance() ---- */
        int have = 0;
        fqo_idx_record ix;
        if (pending) {
            b.start += pending; /* Buffer::consume, src/buffer.rs:107-111 */
            consumed += pending;
            pending = 0;
        }
        for (;;) {
            int kind = 0;
            int rc = fqo_from_buffer(b.data + b.start, buffer_len(&b), &kind, &ix);
            if (rc != FQO_OK) { /* :263 */
                res->status = rc;
                res->err_offset = consumed;
                goto done;
            }
            if (kind == FQO_RES_EMPTY) { /* :264-275 */
                buffer_clean(&b);
                if (buffer_read_into(&b, rd) == 0)
                    break; /* clean EOF: current = None */
                continue;
            }
            if (kind == FQO_RES_INCOMPLETE) { /* :276-294 */
                buffer_clean(&b);
                if (buffer_free(&b) == 0) {
                    res->status = FQO_E_TOO_LONG;
                    res->err_offset = consumed;
                    goto done;
                }
                if (buffer_read_into(&b, rd) == 0) {
                    res->status = FQO_E_TRUNCATED;
                    res->err_offset = consumed;
                    goto done;
                }
                continue;
            }
            have = 1; /* :295-300 */
            pending = ix.data1 - ix.data0;
            break;
        }
        /* ---- each(): get() + closure, :228-236 ---- */
        if (!have) {
            res->finished = 1;
            goto done;
        } else {
            fqo_ref_record rec;
            to_ref_record(&ix, b.data + b.start, &rec);
            res->n_delivered++;
            if (!fn(user, &rec, consumed)) {
                res->finished = 0;
                goto done;
            }
        }
    }
done:
    free(b.data);
}

/* ======================================================================================
 * stats closure
 * ==================================================================================== */
fqo_stats *fqo_stats_new(uint32_t max_len)
{
    fqo_stats *s = (fqo_stats *)calloc(1, sizeof *s);
    if (!s)
        return NULL;
    s->max_len = max_len;
    s->base_hist = (uint64_t *)calloc((size_t)max_len * 6 + 1, sizeof(uint64_t));
    s->qual_hist = (uint64_t *)calloc((size_t)max_len * 256 + 1, sizeof(uint64_t));
    s->len_hist = (uint64_t *)calloc((size_t)max_len + 2, sizeof(uint64_t));
    if (!s->base_hist || !s->qual_hist || !s->len_hist) {
        fqo_stats_free(s);
        return NULL;
    }
    return s;
}

void fqo_stats_free(fqo_stats *s)
{
    if (!s)
        return;
    free(s->base_hist);
    free(s->qual_hist);
    free(s->len_hist);
    free(s);
}

void fqo_stats_clear(fqo_stats *s)
{
    s->n_records = s->n_bases = s->clip_seq = s->clip_qual = 0;
    memset(s->base_hist, 0, (size_t)s->max_len * 6 * sizeof(uint64_t));
    memset(s->qual_hist, 0, (size_t)s->max_len * 256 * sizeof(uint64_t));
    memset(s->len_hist, 0, ((size_t)s->max_len + 2) * sizeof(uint64_t));
}

void fqo_stats_add(fqo_stats *d, const fqo_stats *s)
{
    size_t i, P = d->max_len;
    d->n_records += s->n_records;
    d->n_bases += s->n_bases;
    d->clip_seq += s->clip_seq;
    d->clip_qual += s->clip_qual;
    for (i = 0; i < P * 6; i++) d->base_hist[i] += s->base_hist[i];
    for (i = 0; i < P * 256; i++) d->qual_hist[i] += s->qual_hist[i];
    for (i = 0; i < P + 2; i++) d->len_hist[i] += s->len_hist[i];
}

/* base alphabet of validate_dnan (src/records.rs:29-33), uppercase only */
static inline unsigned base_class(uint8_t c)
{
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    case 'N': return 4;
    default:  return 5;
    }
}

void fqo_stats_record(fqo_stats *s, const fqo_ref_record *r)
{
    const uint8_t *sq, *ql;
    size_t ns, nq, i, P = s->max_len, lim;
    fqo_rec_seq(r, &sq, &ns);
    fqo_rec_qual(r, &ql, &nq);
    s->n_records++;
    s->n_bases += ns;
    s->len_hist[ns <= P ? ns : P + 1]++;
    lim = ns < P ? ns : P;
    for (i = 0; i < lim; i++)
        s->base_hist[i * 6 + base_class(sq[i])]++;
    s->clip_seq += ns - lim;
    lim = nq < P ? nq : P;
    for (i = 0; i < lim; i++)
        s->qual_hist[i * 256 + ql[i]]++;
    s->clip_qual += nq - lim;
}

static int stats_cb(void *user, const fqo_ref_record *rec, uint64_t off)
{
    (void)off;
    fqo_stats_record((fqo_stats *)user, rec);
    return 1;
}

void fqo_each_stats(const uint8_t *data, size_t len, size_t bufsize, size_t max_read,
                    fqo_stats *stats, fqo_each_result *res)
{
    fqo_reader rd = {data, len, 0, max_read};
    fqo_each(&rd, bufsize, stats_cb, stats, res);
}

typedef struct {
    uint64_t *out;
    size_t cap, n;
} index_sink;

static int index_cb(void *user, const fqo_ref_record *rec, uint64_t off)
{
    index_sink *k = (index_sink *)user;
    if (k->n < k->cap) {
        uint64_t *o = k->out + 5 * k->n;
        o[0] = off;
        o[1] = off + rec->head;
        o[2] = off + rec->seq;
        o[3] = off + rec->sep;
        o[4] = off + rec->qual;
    }
    k->n++;
    return 1;
}

void fqo_each_index(const uint8_t *data, size_t len, size_t bufsize, size_t max_read,
                    uint64_t *out, size_t cap, fqo_each_result *res)
{
    fqo_reader rd = {data, len, 0, max_read};
    index_sink k = {out, cap, 0};
    fqo_each(&rd, bufsize, index_cb, &k, res);
}

/* each() with a closure that writes the records passing validate_dna (mode 1) / validate_dnan (mode 2)
 * (src/records.rs:19-33; mode 0 = every record) with RefRecord::write (src/records.rs:93-96: raw bytes) */
typedef struct {
    uint8_t *out;
    size_t cap, n_bytes, n_kept;
    int mode;
} filter_sink;

static int filter_cb(void *user, const fqo_ref_record *rec, uint64_t off)
{
    filter_sink *k = (filter_sink *)user;
    (void)off;
    int ok = k->mode == 1 ? fqo_validate_dna(rec) : k->mode == 2 ? fqo_validate_dnan(rec) : 1;
    if (ok) {
        if (k->n_bytes + rec->len <= k->cap)
            memcpy(k->out + k->n_bytes, rec->data, rec->len);
        k->n_bytes += rec->len;
        k->n_kept++;
    }
    return 1;
}

void fqo_each_filter(const uint8_t *data, size_t len, size_t bufsize, size_t max_read, int mode,
                     uint8_t *out, size_t cap, uint64_t *n_kept, uint64_t *n_bytes, fqo_each_result *res)
{
    fqo_reader rd = {data, len, 0, max_read};
    filter_sink k = {out, cap, 0, 0, mode};
    fqo_each(&rd, bufsize, filter_cb, &k, res);
    *n_kept = k.n_kept;
    *n_bytes = k.n_bytes;
}

/* ======================================================================================
 * record_sets (src/lib.rs:355-436)
 * ==================================================================================== */
struct fqo_set_iter {
    fqo_reader *rd;
    fqo_buffer buf;
    size_t bufsize;
    size_t guess;  /* num_records_guess, :434 */
    int at_end;    /* reader_at_end */
};

fqo_set_iter *fqo_record_sets_new(fqo_reader *rd, size_t bufsize)
{
    fqo_set_iter *it = (fqo_set_iter *)calloc(1, sizeof *it);
    if (!it)
        return NULL;
    it->rd = rd;
    it->bufsize = bufsize;
    it->guess = 100;
    if (!buffer_init(&it->buf, bufsize)) {
        free(it);
        return NULL;
    }
    return it;
}

void fqo_record_sets_free(fqo_set_iter *it)
{
    if (!it)
        return;
    free(it->buf.data);
    free(it);
}

void fqo_record_set_free(fqo_record_set *s)
{
    if (!s)
        return;
    free(s->buffer);
    free(s->records);
    free(s);
}

void fqo_record_set_get(const fqo_record_set *s, size_t i, fqo_ref_record *out)
{
    to_ref_record(&s->records[i], s->buffer, out); /* RecordSetItems::next, :347-352 */
}

static fqo_record_set *emit_set(fqo_set_iter *it, fqo_idx_record *recs, size_t n, size_t cap)
{
    /* vec![0u8; BUFSIZE] + replace_buffer, :384-385 / :396-397 */
    fqo_record_set *s = (fqo_record_set *)calloc(1, sizeof *s);
    uint8_t *fresh = (uint8_t *)calloc(it->bufsize ? it->bufsize : 1, 1);
    s->buffer = buffer_replace(&it->buf, fresh);
    s->bufsize = it->bufsize;
    s->records = recs;
    s->n_records = n;
    s->cap_records = cap;
    return s;
}

int fqo_record_sets_next(fqo_set_iter *it, fqo_record_set **out) /* :364-425 */
{
    fqo_idx_record *recs;
    size_t n = 0, cap;
    *out = NULL;
    if (it->at_end)
        return 0;
    cap = it->guess ? it->guess : 1;
    recs = (fqo_idx_record *)malloc(cap * sizeof *recs); /* Vec::with_capacity, :369 */
    for (;;) {
        int kind = 0;
        fqo_idx_record ix;
        int rc = fqo_from_buffer(it->buf.data + it->buf.start, buffer_len(&it->buf), &kind, &ix);
        if (rc != FQO_OK) { /* :375 */
            free(recs);
            return -rc;
        }
        if (kind == FQO_RES_EMPTY) { /* :381-392 */
            it->guess = n + 1;
            *out = emit_set(it, recs, n, cap);
            if (buffer_read_into(&it->buf, it->rd) == 0)
                it->at_end = 1;
            return 1;
        }
        if (kind == FQO_RES_INCOMPLETE) { /* :393-415 */
            fqo_record_set *s;
            it->guess = n + 1;
            s = emit_set(it, recs, n, cap);
            if (buffer_free(&it->buf) == 0) {
                fqo_record_set_free(s);
                return -FQO_E_TOO_LONG;
            }
            if (buffer_read_into(&it->buf, it->rd) == 0) {
                fqo_record_set_free(s);
                return -FQO_E_TRUNCATED;
            }
            *out = s;
            return 1;
        }
        /* Record, :416-422 */
        ix.data0 += it->buf.start;
        ix.data1 += it->buf.start;
        if (n == cap) {
            cap *= 2;
            recs = (fqo_idx_record *)realloc(recs, cap * sizeof *recs);
        }
        recs[n++] = ix;
        it->buf.start += ix.data1 - ix.data0;
    }
}

/* ======================================================================================
 * parallel_each (src/lib.rs:509-565): one bounded channel (depth 10) per worker, sets
 * dealt round-robin by the caller thread, workers joined at the end.
 * ==================================================================================== */
#define CHAN_DEPTH 10 /* sync_channel(10), :522 */

typedef struct {
    pthread_mutex_t mu;
    pthread_cond_t not_empty, not_full;
    fqo_record_set *slot[CHAN_DEPTH];
    int head, count, closed, rx_gone;
} chan;

static void chan_init(chan *c)
{
    memset(c, 0, sizeof *c);
    pthread_mutex_init(&c->mu, NULL);
    pthread_cond_init(&c->not_empty, NULL);
    pthread_cond_init(&c->not_full, NULL);
}

static int chan_send(chan *c, fqo_record_set *s) /* SyncSender::send, :540 */
{
    pthread_mutex_lock(&c->mu);
    while (c->count == CHAN_DEPTH && !c->rx_gone)
        pthread_cond_wait(&c->not_full, &c->mu);
    if (c->rx_gone) {
        pthread_mutex_unlock(&c->mu);
        return 0;
    }
    c->slot[(c->head + c->count) % CHAN_DEPTH] = s;
    c->count++;
    pthread_cond_signal(&c->not_empty);
    pthread_mutex_unlock(&c->mu);
    return 1;
}

static fqo_record_set *chan_recv(chan *c) /* rx.into_iter().next(), :527 */
{
    fqo_record_set *s = NULL;
    pthread_mutex_lock(&c->mu);
    while (c->count == 0 && !c->closed)
        pthread_cond_wait(&c->not_empty, &c->mu);
    if (c->count) {
        s = c->slot[c->head];
        c->head = (c->head + 1) % CHAN_DEPTH;
        c->count--;
        pthread_cond_signal(&c->not_full);
    }
    pthread_mutex_unlock(&c->mu);
    return s;
}

static void chan_close(chan *c) /* drop(senders), :551 */
{
    pthread_mutex_lock(&c->mu);
    c->closed = 1;
    pthread_cond_broadcast(&c->not_empty);
    pthread_mutex_unlock(&c->mu);
}

typedef struct {
    chan ch;
    pthread_t th;
    fqo_stats *stats;   /* NULL for count-only */
    uint64_t n_records; /* count-only closure */
    uint64_t n_sets;
} worker;

static void *worker_main(void *arg) /* the closure body, e.g. examples/fastq-count-thread.rs */
{
    worker *w = (worker *)arg;
    fqo_record_set *s;
    while ((s = chan_recv(&w->ch)) != NULL) {
        w->n_sets++;
        if (w->stats) {
            size_t i;
            for (i = 0; i < s->n_records; i++) {
                fqo_ref_record rec;
                fqo_record_set_get(s, i, &rec);
                fqo_stats_record(w->stats, &rec);
            }
        } else {
            w->n_records += s->n_records; /* record_set.len() */
        }
        fqo_record_set_free(s);
    }
    return NULL;
}

static int parallel_each_impl(const uint8_t *data, size_t len, size_t bufsize, size_t max_read,
                              int n_threads, uint32_t max_len, int want_stats, fqo_stats *total,
                              uint64_t *n_records, uint64_t *sets_per_worker)
{
    fqo_reader rd = {data, len, 0, max_read};
    fqo_set_iter *it = fqo_record_sets_new(&rd, bufsize);
    worker *ws = (worker *)calloc((size_t)n_threads, sizeof *ws);
    int i, next = 0, err = FQO_OK;

    for (i = 0; i < n_threads; i++) { /* :521-532 */
        chan_init(&ws[i].ch);
        ws[i].stats = want_stats ? fqo_stats_new(max_len) : NULL;
        pthread_create(&ws[i].th, NULL, worker_main, &ws[i]);
    }
    for (;;) { /* self.record_sets().zip(senders.iter().cycle()), :535 */
        fqo_record_set *s = NULL;
        int rc;
        if (n_threads == 0)
            break; /* zip with an empty cycle yields nothing */
        rc = fqo_record_sets_next(it, &s);
        if (rc == 0)
            break;
        if (rc < 0) { /* :544-547 */
            err = -rc;
            break;
        }
        if (!chan_send(&ws[next].ch, s)) { /* :540-542 */
            fqo_record_set_free(s);
            break;
        }
        next = (next + 1) % n_threads;
    }
    for (i = 0; i < n_threads; i++)
        chan_close(&ws[i].ch);
    for (i = 0; i < n_threads; i++) { /* join, :553-559 */
        pthread_join(ws[i].th, NULL);
        if (want_stats) {
            if (err == FQO_OK)
                fqo_stats_add(total, ws[i].stats);
            fqo_stats_free(ws[i].stats);
        } else if (n_records) {
            *n_records += ws[i].n_records;
        }
        if (sets_per_worker)
            sets_per_worker[i] = ws[i].n_sets;
    }
    free(ws);
    fqo_record_sets_free(it);
    return err; /* Err(e) discards worker results, :561-564 */
}

int fqo_parallel_each_stats(const uint8_t *data, size_t len, size_t bufsize, size_t max_read,
                            int n_threads, fqo_stats *total, uint64_t *sets_per_worker)
{
    return parallel_each_impl(data, len, bufsize, max_read, n_threads, total->max_len, 1, total,
                              NULL, sets_per_worker);
}

int fqo_parallel_each_count(const uint8_t *data, size_t len, size_t bufsize, size_t max_read,
                            int n_threads, uint64_t *n_records)
{
    *n_records = 0;
    return parallel_each_impl(data, len, bufsize, max_read, n_threads, 0, 0, NULL, n_records,
                              NULL);
}

/* ======================================================================================
 * synthetic FASTQ
 * ==================================================================================== */
uint64_t fqo_splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

uint32_t fqo_synth_var_len(uint64_t seed, uint64_t rec)
{
    return 50u + (uint32_t)(fqo_splitmix64(seed ^ 0x4C454Eull ^ rec) % 251u);
}

size_t fqo_synth_fixed_record_bytes(uint32_t L) { return 17 + 2 * ((size_t)L + 1) + 2; }

/* byte k of record `rec` whose read length is L */
static inline uint8_t synth_byte(uint64_t seed, uint64_t rec, uint32_t L, uint32_t k)
{
    uint32_t j;
    uint64_t h;
    if (k < 17) {
        static const uint64_t pow10[13] = {1000000000000ull, 100000000000ull, 10000000000ull,
                                           1000000000ull, 100000000ull, 10000000ull, 1000000ull,
                                           100000ull, 10000ull, 1000ull, 100ull, 10ull, 1ull};
        if (k == 0) return '@';
        if (k == 1) return 'F';
        if (k == 2) return 'Q';
        if (k == 16) return '\n';
        return (uint8_t)('0' + (rec / pow10[k - 3]) % 10);
    }
    k -= 17;
    if (k < L) { /* base j */
        j = k;
        h = fqo_splitmix64(seed ^ ((rec << 10) | j));
        if ((h & 0xFF) < 2)
            return 'N';
        return (uint8_t)"ACGT"[(h >> 8) & 3];
    }
    if (k == L) return '\n';
    if (k == L + 1) return '+';
    if (k == L + 2) return '\n';
    k -= L + 3;
    if (k < L) { /* quality j */
        uint32_t span;
        j = k;
        h = fqo_splitmix64(seed ^ ((rec << 10) | j));
        span = 40u - (20u * j) / L;
        return (uint8_t)(33u + 2u + (uint32_t)((h >> 16) % span));
    }
    return '\n';
}

size_t fqo_synth_record(uint64_t seed, uint64_t rec, uint32_t L, uint8_t *out)
{
    uint32_t len = L ? L : fqo_synth_var_len(seed, rec);
    size_t n = fqo_synth_fixed_record_bytes(len), k;
    if (out)
        for (k = 0; k < n; k++)
            out[k] = synth_byte(seed, rec, len, (uint32_t)k);
    return n;
}

void fqo_synth_fixed(uint64_t seed, uint32_t L, uint64_t byte_off, size_t n, uint8_t *out)
{
    uint64_t rb = fqo_synth_fixed_record_bytes(L);
    uint64_t rec = byte_off / rb;
    uint32_t k = (uint32_t)(byte_off % rb);
    size_t i;
    for (i = 0; i < n; i++) {
        out[i] = synth_byte(seed, rec, L, k);
        if (++k == rb) {
            k = 0;
            rec++;
        }
    }
}

size_t fqo_synth_var(uint64_t seed, uint64_t first, uint64_t count, uint8_t *out)
{
    size_t total = 0;
    uint64_t r;
    for (r = first; r < first + count; r++)
        total += fqo_synth_record(seed, r, 0, out ? out + total : NULL);
    return total;
}
