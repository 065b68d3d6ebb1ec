/* batch_each.c -- the generic-closure path of the C ABI (include/fastq_b200.h) driven WITHOUT Python:
 * what a Rust / C caller of the drop-in does for Parser::new(file).each(closure)  (src/lib.rs:221-238).
 *
 *   producer thread  read(2) into the pinned ring slots: fqb_stream_acquire -> read -> fqb_stream_submit
 *                    (thread_reader's protocol, src/thread_reader.rs:40-50), fqb_batch_close at EOF
 *   main thread      fqb_next_batch -> the "closure" over every record of the batch (here: count, sum of seq()
 *                    lengths with the '\r' trim of src/records.rs:65-73, a position-weighted checksum of the raw record bytes)
 *                    -> fqb_release_batch
 *
 *   usage: batch_each FILE [slot_kib]      prints: status n_records n_bases checksum err_offset n_batches
 * Test infrastructure (tests/test_gpu_parity.py::test_c_program_drives_the_batch_mode compiles and runs it).
 */
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "fastq_b200.h"

struct producer {
    fqb_ctx *ctx;
    int fd;
    int rc;
};

static void *pump(void *arg)
{
    struct producer *p = (struct producer *)arg;
    for (;;) {
        uint8_t *slot;
        uint64_t cap;
        int rc = fqb_stream_acquire(p->ctx, &slot, &cap);
        if (rc == FQB_E_CANCELLED) return NULL;          /* the consumer stopped */
        if (rc != FQB_OK) { p->rc = rc; fqb_batch_cancel(p->ctx); return NULL; }
        ssize_t n = read(p->fd, slot, cap);
        if (getenv("FQB_DEBUG")) fprintf(stderr, "slot %p cap %llu read %lld\n", (void *)slot, (unsigned long long)cap, (long long)n);
        if (n < 0) { p->rc = FQB_E_IO; fqb_stream_submit(p->ctx, 0); fqb_batch_cancel(p->ctx); return NULL; }
        rc = fqb_stream_submit(p->ctx, (uint64_t)n);
        if (rc != FQB_OK) { p->rc = rc; fqb_batch_cancel(p->ctx); return NULL; }
        if (n == 0) { p->rc = fqb_batch_close(p->ctx); return NULL; }
    }
}

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s FILE [slot_kib]\n", argv[0]); return 2; }
    fqb_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = FQB_ABI_VERSION;
    cfg.max_len = 150;
    cfg.slot_bytes = (argc > 2 ? strtoull(argv[2], NULL, 10) : 4096) * 1024;
    cfg.n_slots = 4;
    fqb_ctx *ctx = NULL;
    int rc = fqb_create(&cfg, &ctx);
    if (rc != FQB_OK) { fprintf(stderr, "fqb_create: %s\n", fqb_strerror(rc)); return 1; }
    struct producer p = {ctx, open(argv[1], O_RDONLY), FQB_OK};
    if (p.fd < 0) { perror(argv[1]); return 1; }
    rc = fqb_batch_begin(ctx, 0);
    if (rc != FQB_OK) { fprintf(stderr, "fqb_batch_begin: %s %s\n", fqb_strerror(rc), fqb_last_error(ctx)); return 1; }
    pthread_t th;
    pthread_create(&th, NULL, pump, &p);

    uint64_t n_records = 0, n_bases = 0, s1 = 0, s2 = 0, pos = 0, n_batches = 0, err_offset = 0;
    int status = FQB_OK;
    for (;;) {
        fqb_batch b;
        rc = fqb_next_batch(ctx, &b);
        if (rc != FQB_OK) { status = rc; break; }        /* cancelled by the producer (I/O error) or a CUDA failure */
        ++n_batches;
        const uint32_t base = (uint32_t)b.stream_offset;
        uint64_t start = 0;
        for (uint64_t k = 0; k < b.n_records; ++k) {     /* the closure */
            const uint32_t *le = b.line_ends + 4 * k;
            const uint64_t head = (uint32_t)(le[0] - base), seq = (uint32_t)(le[1] - base), qual = (uint32_t)(le[3] - base);
            uint64_t ls = seq - head - 1;
            if (ls && b.bytes[seq - 1] == '\r') --ls;    /* trim_winline */
            n_bases += ls;
            for (uint64_t i = start; i <= qual; ++i) {   /* sum of bytes, sum of byte x (1 + its rank among the record bytes) */
                s1 += b.bytes[i];
                s2 += (uint64_t)b.bytes[i] * ++pos;
            }
            start = qual + 1;
        }
        n_records += b.n_records;
        status = b.status;
        err_offset = b.err_offset;
        const int last = b.last;
        if (b.token != UINT64_MAX) fqb_release_batch(ctx, b.token);
        if (last) break;
    }
    fqb_batch_cancel(ctx);                               /* (a no-op after a clean end: stops a producer still reading) */
    pthread_join(th, NULL);
    fqb_result res;
    fqb_batch_end(ctx, &res);
    if (status == FQB_OK && p.rc != FQB_OK) status = p.rc;
    printf("%d %llu %llu %llu %llu %llu\n", status, (unsigned long long)n_records, (unsigned long long)n_bases,
           (unsigned long long)(s1 * 1000003ull + s2), (unsigned long long)err_offset, (unsigned long long)n_batches);
    close(p.fd);
    fqb_destroy(ctx);
    return 0;
}
