// fq_api.cu -- the C ABI (include/fastq_b200.h) over the sm_100a kernels.
//
// Host-side counterpart of the reference's drivers:
//   Parser::new / each            src/lib.rs:198-238   -> fqb_create / fqb_parse_device / fqb_parse_host
//   Buffer (sliding window)       src/buffer.rs:3-112  -> contiguous device ring: a record is owned by
//                                                        the chunk it starts in and read through into the next
//   thread_reader (2-queue ring)  src/thread_reader.rs:8-200 -> pinned slots: acquire / submit / recycle on event
#include "../../include/fastq_b200.h"
#include "fq_common.cuh"
#include "fq_device.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace fq;

namespace {

// Reset of the device-resident outcome + a look at the head of the shard (one warp): do its reads vary in
// length?  The first PROBE_BYTES are cut into lines; with the strict 4-line cadence (src/records.rs:201-247)
// line j and line j + 4 belong to the same kind of line whatever the phase, so on fixed-length reads at least
// three of the four kinds keep their length from record to record (the id line may vary: instrument
// coordinates).  If two or more kinds vary, the launch goes to the speculative kernel's variable-length
// variant -- and so does a shard whose head holds bytes >= 0x80.  A hint only: either variant delivers the
// exact result on any input.
constexpr int PROBE_BYTES = 8192, PROBE_LINES = 256, PROBE_THREADS = 512;
__global__ void __launch_bounds__(PROBE_THREADS) fq_init_kernel(DevResult* r, int spec_fail, int line_phase, const uint8_t* data,
                                                                unsigned long long n, int probe)
{
    __shared__ unsigned short pos[PROBE_LINES + 1];
    __shared__ unsigned int warp_cnt[PROBE_THREADS / 32];
    __shared__ unsigned int s_varies, s_hi;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    int shape_var = 0;
    if (probe && n >= (unsigned long long)PROBE_BYTES) {
        // newline positions of the first PROBE_BYTES, in order: 16 bytes per thread, block prefix of the counts
        if (t == 0) {
            s_varies = 0;
            s_hi = 0;
        }
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(data) + t);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        unsigned m = 0;                                          // bit i: byte i of the piece is '\n'
        for (int i = 0; i < 16; ++i) m |= (((w[i >> 2] >> (8 * (i & 3))) & 0xFFu) == '\n' ? 1u : 0u) << i;
        const unsigned hi = (v.x | v.y | v.z | v.w) & 0x80808080u;
        const int c = __popc(m);
        int incl = c;
        for (int k = 1; k < 32; k <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, k);
            if (lane >= k) incl += u;
        }
        if (lane == 31) warp_cnt[warp] = incl;
        __syncthreads();
        int before = 0, total = 0;
        for (int k = 0; k < PROBE_THREADS / 32; ++k) {
            if (k < warp) before += warp_cnt[k];
            total += warp_cnt[k];
        }
        int rank = before + incl - c;
        while (m && rank < PROBE_LINES) {
            pos[rank++] = (unsigned short)(16 * t + __ffs(m) - 1);
            m &= m - 1;
        }
        if (hi) atomicOr(&s_hi, 1u);
        __syncthreads();
        const int nl = total < PROBE_LINES ? total : PROBE_LINES;
        // line j = (pos[j], pos[j + 1]]; compare the lengths of line j and line j + 4
        unsigned varies = 0;                                     // bit k: some line of kind k changed its length
        for (int j = t; j + 5 < nl; j += PROBE_THREADS)
            if (pos[j + 1] - pos[j] != pos[j + 5] - pos[j + 4]) varies |= 1u << (j & 3);
        if (varies) atomicOr(&s_varies, varies);
        __syncthreads();
        // (probe == 2, no histograms: the predicting variant needs every line length fixed there -- it checks a
        // window by its '\n' count --, so any kind that varies sends the launch to the variable-length variant)
        shape_var = (nl >= 16 && __popc(s_varies) >= (probe == 2 ? 1 : 2)) ? 1 : 0;
        // bytes >= 0x80 (UTF-8 in the id lines, say): with histograms the predicting variant leaves the fast path at
        // the first window it has to scan that holds one; the variable-length variant only minds them in the
        // sequence and quality lines themselves
        if (s_hi && probe == 1) shape_var = 1;
    }
    if (t != 0) return;
    r->shape_var = shape_var;
    r->first_bad = NONE64;
    r->tail_start = NONE64;
    r->n_lines = 0;
    r->line_end = 0;
    r->err_offset = 0;
    r->n_records = 0;
    r->status = 0;
    r->finished = 0;
    r->spec_fail = spec_fail;   // 1: no speculative launch, go straight to the exact path
    r->line_phase = line_phase;
    r->n_win_pred = 0;
    r->n_win_scan = 0;
    r->spec_bad = NONE64;
    r->spec_retry = 0;
    r->tail_err = 0;
}

// streaming with FQB_F_INDEX into PINNED caller memory: the line ends of one chunk go straight to the mapped
// host buffer, at the running line count the carry holds -- no host round trip per chunk
__global__ void __launch_bounds__(256) fq_index_out_kernel(const uint32_t* __restrict__ idx, const DevResult* r,
                                                           const DevCarry* carry, uint32_t* out, unsigned long long cap)
{
    const unsigned long long n = r->n_lines;
    const unsigned long long off = carry->n_lines - n;          // fq_finalize_kernel has added this chunk already
    const unsigned long long m = off >= cap ? 0ull : (n < cap - off ? n : cap - off);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride)
        out[off + i] = idx[i];
}

// batch mode: the outcome of ONE chunk and its line ends, written by the device into pinned host memory right
// behind the chunk's kernels (stream order) -- the consumer thread waits for an event, not for the stream
constexpr int BATCH_INFO_WORDS = 8;   // status, n_records, n_lines, err_offset, line_base before the chunk, tail + 1, -, -
__global__ void __launch_bounds__(256) fq_batch_out_kernel(const uint32_t* __restrict__ idx, const DevResult* r,
                                                           unsigned long long* info, uint32_t* out, unsigned long long cap)
{
    const unsigned long long n = r->n_lines;
    const unsigned long long m = n < cap ? n : cap;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) out[i] = idx[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        info[0] = (unsigned long long)r->status;
        info[1] = r->n_records;
        info[2] = n;
        info[3] = r->err_offset;
        info[4] = r->line_end - n;
        info[5] = r->tail_start == NONE64 ? 0ull : r->tail_start + 1ull;
        info[6] = 0;
        info[7] = 0;
    }
}

// ---- a bad record in the middle of a shard (DevResult::spec_retry): after the bytes in front of it have been parsed
// again on their own, these put the outcome of the WHOLE shard together: the bad record is classified
// (fq_diagnose_kernel), the '\n' behind it still count as lines of the shard, the outcome is published anew
__global__ void fq_retry_mark_kernel(DevResult* r, unsigned long long x)
{
    r->first_bad = x;
    r->tail_err = 1;     // (nothing behind it was counted: no restricted second pass needed)
}
__global__ void __launch_bounds__(256) fq_tail_lines_kernel(const uint8_t* __restrict__ d, unsigned long long x,
                                                            unsigned long long n_own, DevResult* r)
{
    const unsigned long long xa = min(n_own, (x + 15ull) & ~15ull), ne = xa + ((n_own - xa) & ~15ull);
    unsigned long long cnt = 0;
    const unsigned long long gt = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = xa / 16 + gt; i < ne / 16; i += stride) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(d) + i);
        cnt += __popc(nlbits(v.x)) + __popc(nlbits(v.y)) + __popc(nlbits(v.z)) + __popc(nlbits(v.w));
    }
    if (gt < xa - x) cnt += d[x + gt] == '\n';                   // the bytes in front of the first aligned piece
    if (gt < n_own - ne) cnt += d[ne + gt] == '\n';              // ... and behind the last one
    cnt = warp_sum_u64(cnt);
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicAdd(&r->n_lines, cnt);
        atomicAdd(&r->line_end, cnt);
    }
}
__global__ void fq_republish_kernel(const ScanParams p, unsigned long long* pub, unsigned long long* slot)
{
    DevResult* r = p.res;
    r->finished = 0;
    r->tail_err = 0;
    unsigned long long w[8] = {(unsigned long long)r->status, 0ull, r->n_records, r->n_lines, r->err_offset,
                               NONE64, (unsigned long long)r->line_phase, 0ull};
    for (int k = 0; k < 8; ++k) {
        if (pub) pub[k] = w[k];
        if (slot) slot[k] = w[k];
    }
}

struct Slot {  // one stage of the streaming ring
    uint8_t* h_pinned = nullptr;
    cudaEvent_t copied = nullptr;  // H2D of the chunk in this host slot finished
    bool copied_pending = false;
};

struct DevSlot {
    cudaEvent_t done = nullptr;  // kernels that read this device slot finished
    bool done_pending = false;
    uint32_t* d_index = nullptr;
    uint64_t bytes = 0;
};

}  // namespace

struct fqb_ctx {
    int device = 0;
    uint32_t P = 0;
    int nchunk = 5;
    int num_sms = 148;
    int grid = 148;
    size_t nwords = 0;
    // one-shot / per-chunk device state
    uint64_t* d_stats = nullptr;      // [statistics block (nwords) | FQB_MAX_WORLD x 8 outcome words, one slot per rank]
    uint64_t* d_reduced = nullptr;    // the same layout after fqb_allreduce
    uint64_t* h_reduced = nullptr;    // pinned
    int rank = 0, world = 1;          // fqb_comm_init
    void* nccl_comm = nullptr;
    uint64_t* d_seqraw = nullptr;
    DevResult* d_res = nullptr;
    unsigned long long* d_pub = nullptr;   // outcome of the last parse as 8 device words (fqb_device_result)
    RangeInfo* d_ranges = nullptr;
    StreamRange* d_sranges = nullptr;
    uint32_t* d_index_stage = nullptr;  // speculative launch: per-range staging of the line ends
    size_t index_stage_cap = 0;
    unsigned long long* d_linecount = nullptr;
    // record filter workspace (grown on demand)
    unsigned long long* d_fblk = nullptr;    // look-back descriptors, one per 256 records
    size_t fblk_cap = 0;
    unsigned long long* d_fmisc = nullptr;   // wraps [1 + FILTER_MAX_WRAPS], result [4], ticket [1]
    unsigned long long* h_fres = nullptr;    // pinned [4]
    DevResult* h_res = nullptr;  // pinned
    unsigned long long* h_linecount = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;   // scan kernel; index compaction
    bool ev_valid = false, ev_index_valid = false;
    uint64_t launches = 0;
    unsigned long long* d_trace = nullptr;  // FQB_TRACE=<file>: kernel timeline, dumped by fqb_fetch
    std::string trace_path;
    uint64_t last_off = 0;
    fqb_shard last_shard = {};     // of the last fqb_parse_device (fqb_fetch may have to parse a part of it again)
    uint64_t retries = 0;          // 1 if fqb_fetch did
    std::string err;
    // streaming
    uint64_t slot_bytes = 0;
    uint32_t n_slots = 0;
    std::vector<Slot> slots;
    std::vector<DevSlot> dslots;
    uint8_t* d_ring = nullptr;  // n_dev * slot_bytes + halo
    uint32_t n_dev = 0;
    uint64_t* d_total = nullptr;
    DevCarry* d_carry = nullptr;
    DevCarry* h_carry = nullptr;  // pinned
    cudaStream_t s_copy = nullptr, s_comp = nullptr;
    bool streaming = false;
    uint32_t stream_flags = 0;
    bool stream_partial = false; // FQB_F_PARTIAL: the last chunk is not the end of the stream
    uint64_t chunk_no = 0;       // chunks submitted to the device
    uint64_t fill = 0;           // bytes filled in the current host slot
    uint64_t stream_pos = 0;     // stream offset of the current host slot's first byte
    bool acquired = false;
    bool have_pending = false;   // a copied chunk is waiting for its successor (or EOF)
    uint64_t pending_bytes = 0, pending_chunk = 0, pending_off = 0;
    bool pending_line_start = true;  // the byte before the pending chunk is '\n' (or stream start)
    bool last_byte_nl = true;        // last byte copied so far is '\n' (true before the first byte)
    uint8_t* h_ring = nullptr;   // n_slots pinned slots back to back + MAXREC mirror of slot 0
    // batch mode (fqb_batch_begin .. fqb_next_batch): one producer thread (acquire / submit / close), one
    // consumer thread (next / release); everything below is guarded by `mu`
    struct BChunk {
        uint64_t no = ~0ull;         // chunk number held by this host slot
        uint64_t off = 0, n_bytes = 0;
        bool line_start = false;     // the byte in front of the chunk is '\n' (or the stream starts here)
        bool launched = false;       // its parse has been enqueued (event `parsed` recorded)
        bool released = true;        // the consumer has given the batch back
        cudaEvent_t parsed = nullptr;
        unsigned long long* h_info = nullptr;   // pinned, written by the device: BATCH_INFO_WORDS words
        uint32_t* h_index = nullptr;            // pinned: the chunk's line ends
        uint32_t* h_index_dev = nullptr;        // its device alias
        unsigned long long* h_info_dev = nullptr;
        size_t index_cap = 0;
        const uint32_t* d_index = nullptr;      // device copy (valid until the device slot is reused)
    };
    std::vector<BChunk> bchunks;
    std::mutex mu;
    std::condition_variable cv;
    bool batch_mode = false, batch_closed = false, batch_cancel = false, batch_failed = false;
    uint64_t b_total = ~0ull;        // chunks of the stream (known once the producer has closed it)
    uint64_t b_next = 0;             // next chunk to deliver
    uint64_t b_held = 0;             // batches delivered and not yet released
    uint64_t b_records = 0;          // records delivered so far
    int b_status = 0;                // status that ended the stream
    // host index collection (generic closure path)
    uint32_t* host_index = nullptr;
    uint32_t* host_index_dev = nullptr;   // device alias of host_index when the caller's buffer is pinned
    uint64_t host_index_cap = 0, host_index_n = 0;
};

static int fail(fqb_ctx* c, cudaError_t e, const char* what)
{
    if (c) {
        char buf[256];
        snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
        c->err = buf;
    }
    return FQB_E_CUDA;
}
#define CK(call)                                         \
    do {                                                 \
        cudaError_t e_ = (call);                         \
        if (e_ != cudaSuccess) return fail(ctx, e_, #call); \
    } while (0)

extern "C" {

uint32_t fqb_abi_version(void) { return FQB_ABI_VERSION; }
size_t fqb_stats_words(uint32_t P) { return stats_words(P); }
size_t fqb_stats_len_hist_off(uint32_t P) { return stats_len_off(P); }
size_t fqb_stats_base_hist_off(uint32_t P) { return stats_base_off(P); }
size_t fqb_stats_qual_hist_off(uint32_t P) { return stats_qual_off(P); }

const char* fqb_strerror(int status)
{
    switch (status) {
    case FQB_OK: return "ok";
    case FQB_E_HEADER: return "Fastq headers must start with '@'";
    case FQB_E_SEP: return "Sequence and quality not separated by +";
    case FQB_E_LENGTH: return "Sequence and quality length mismatch";
    case FQB_E_TOO_LONG: return "Fastq record is too long";
    case FQB_E_TRUNCATED: return "Possibly truncated input file";
    case FQB_E_IO: return "I/O error";
    case FQB_E_PHASE: return "the first record of the shard could not be inferred";
    case FQB_E_ARG: return "invalid argument";
    case FQB_E_STATE: return "call out of order";
    case FQB_E_NOMEM: return "out of memory";
    case FQB_E_CUDA: return "CUDA error";
    default: return "unknown status";
    }
}

const char* fqb_last_error(fqb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int fqb_create(const fqb_config* cfg, fqb_ctx** out)
{
    if (!cfg || !out || cfg->abi_version != FQB_ABI_VERSION || cfg->max_len == 0 || cfg->max_len > 4096)
        return FQB_E_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return FQB_E_CUDA;  // no CPU fallback, by design
    if (cfg->device < 0 || cfg->device >= ndev) return FQB_E_ARG;
    fqb_ctx* ctx = new fqb_ctx();
    ctx->device = cfg->device;
    ctx->P = cfg->max_len;
    ctx->nchunk = cfg->max_len <= 160 ? 5 : 10;
    ctx->nwords = stats_words(ctx->P);
    ctx->slot_bytes = cfg->slot_bytes ? (cfg->slot_bytes + 4095) / 4096 * 4096 : (64ull << 20);
    if (ctx->slot_bytes < 2 * (uint64_t)MAXREC) ctx->slot_bytes = 2 * (uint64_t)MAXREC;
    ctx->n_slots = cfg->n_slots ? cfg->n_slots : 4;
    if (ctx->n_slots < 2) ctx->n_slots = 2;
    cudaError_t e;
#define CKC(call)                        \
    if ((e = (call)) != cudaSuccess) {   \
        fail(ctx, e, #call);             \
        fprintf(stderr, "fastq_b200: %s\n", ctx->err.c_str()); \
        fqb_destroy(ctx);                \
        return FQB_E_CUDA;               \
    }
    CKC(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, ctx->device));
    ctx->num_sms = prop.multiProcessorCount;
    CKC(scan_configure());
    CKC(stream_configure());
    ctx->grid = ctx->num_sms * scan_blocks_per_sm(ctx->nchunk);
    CKC(cudaMalloc(&ctx->d_stats, (ctx->nwords + 8 * (size_t)FQB_MAX_WORLD) * 8));   // [block | 8 outcome words per rank]
    CKC(cudaMalloc(&ctx->d_reduced, (ctx->nwords + 8 * (size_t)FQB_MAX_WORLD) * 8));
    CKC(cudaHostAlloc(&ctx->h_reduced, (ctx->nwords + 8 * (size_t)FQB_MAX_WORLD) * 8, cudaHostAllocDefault));
    CKC(cudaMalloc(&ctx->d_seqraw, (size_t)ctx->P * 256 * 8));
    CKC(cudaMalloc(&ctx->d_res, sizeof(DevResult)));
    CKC(cudaMalloc(&ctx->d_ranges, sizeof(RangeInfo) * ctx->grid));
    CKC(cudaMalloc(&ctx->d_pub, 8 * sizeof(unsigned long long)));
    CKC(cudaMemset(ctx->d_pub, 0, 8 * sizeof(unsigned long long)));
    CKC(cudaMalloc(&ctx->d_sranges, sizeof(StreamRange) * ctx->grid * 32));   // (<= 32 warp ranges per CTA)
    CKC(cudaMalloc(&ctx->d_linecount, 8));
    CKC(cudaMalloc(&ctx->d_carry, sizeof(DevCarry)));
    CKC(cudaHostAlloc(&ctx->h_res, sizeof(DevResult), cudaHostAllocDefault));
    CKC(cudaHostAlloc(&ctx->h_linecount, 8, cudaHostAllocDefault));
    CKC(cudaHostAlloc(&ctx->h_carry, sizeof(DevCarry), cudaHostAllocDefault));
    if (const char* tp = getenv("FQB_TRACE")) {
        ctx->trace_path = tp;
        CKC(cudaMalloc(&ctx->d_trace, (size_t)ctx->num_sms * TRACE_K * 16 * 8));
    }
    CKC(cudaEventCreate(&ctx->ev0));
    CKC(cudaEventCreate(&ctx->ev1));
    CKC(cudaEventCreate(&ctx->ev2));
    CKC(cudaEventCreate(&ctx->ev3));
#undef CKC
    *out = ctx;
    return FQB_OK;
}

static void stream_free(fqb_ctx* ctx)
{
    for (auto& s : ctx->slots)
        if (s.copied) cudaEventDestroy(s.copied);
    ctx->slots.clear();
    if (ctx->h_ring) cudaFreeHost(ctx->h_ring);
    ctx->h_ring = nullptr;
    for (auto& b : ctx->bchunks) {
        if (b.parsed) cudaEventDestroy(b.parsed);
        if (b.h_info) cudaFreeHost(b.h_info);
        if (b.h_index) cudaFreeHost(b.h_index);
    }
    ctx->bchunks.clear();
    for (auto& d : ctx->dslots) {
        if (d.done) cudaEventDestroy(d.done);
        if (d.d_index) cudaFree(d.d_index);
    }
    ctx->dslots.clear();
    if (ctx->d_ring) cudaFree(ctx->d_ring);
    ctx->d_ring = nullptr;
    if (ctx->d_total) cudaFree(ctx->d_total);
    ctx->d_total = nullptr;
    if (ctx->s_copy) cudaStreamDestroy(ctx->s_copy);
    if (ctx->s_comp) cudaStreamDestroy(ctx->s_comp);
    ctx->s_copy = ctx->s_comp = nullptr;
}

void fqb_destroy(fqb_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    stream_free(ctx);
    fqb_comm_destroy(ctx);
    cudaFree(ctx->d_stats);
    cudaFree(ctx->d_reduced);
    if (ctx->h_reduced) cudaFreeHost(ctx->h_reduced);
    cudaFree(ctx->d_seqraw);
    cudaFree(ctx->d_res);
    cudaFree(ctx->d_ranges);
    cudaFree(ctx->d_pub);
    cudaFree(ctx->d_sranges);
    cudaFree(ctx->d_index_stage);
    cudaFree(ctx->d_linecount);
    cudaFree(ctx->d_carry);
    cudaFree(ctx->d_fblk);
    cudaFree(ctx->d_fmisc);
    if (ctx->h_fres) cudaFreeHost(ctx->h_fres);
    cudaFree(ctx->d_trace);
    if (ctx->h_res) cudaFreeHost(ctx->h_res);
    if (ctx->h_linecount) cudaFreeHost(ctx->h_linecount);
    if (ctx->h_carry) cudaFreeHost(ctx->h_carry);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->ev3) cudaEventDestroy(ctx->ev3);
    delete ctx;
}

// Enqueue the whole parse of one shard on `st`:
//   reset -> speculative kernel (fq_stream.cu) + verification of its record chains
//   -> [only if that did not deliver: reset, newline counts of the CTA ranges, exact kernel (fq_scan.cu)]
//   -> diagnose -> [only if a bad record was found: reset + exact kernel again, restricted to the
//   records before it: each() delivers exactly those, src/lib.rs:226-237]
//   -> index compaction (speculative kernel only) -> finalize
static int enqueue_parse(fqb_ctx* ctx, const fqb_shard* sh, cudaStream_t st, DevCarry* carry, uint64_t* total,
                         bool timed, bool allow_retry = false)
{
    if (!sh || sh->n_avail < sh->n_own) return FQB_E_ARG;
    if (sh->n_avail && (!sh->d_bytes || (reinterpret_cast<uintptr_t>(sh->d_bytes) & 15))) return FQB_E_ARG;
    if ((sh->flags & FQB_F_INDEX) && sh->index_cap && !sh->d_index) return FQB_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const uint64_t tile_bytes = scan_tile_bytes(ctx->nchunk);
    const uint64_t ntiles64 = (sh->n_own + tile_bytes - 1) / tile_bytes;
    if (ntiles64 > 0xFFFFFFF0ull) return FQB_E_ARG;
    const bool want_index = (sh->flags & FQB_F_INDEX) && sh->d_index && sh->index_cap;
    // the speculative kernel: 32 warp ranges per CTA, at least 8 KiB each; every range stages its line
    // ends in its own share of a staging area sized for lines of >= 16 bytes on average (a range that
    // needs more gives up and the exact path writes the index)
    const bool fast = ctx->nchunk <= 10 && !getenv("FQB_NO_FAST");
    uint64_t srange_bytes, stage_share = 0, n_sranges, desc_cap = 0;
    {
        // equal ranges, rounded DOWN to 16 bytes; the last range takes the remainder (a little more work
        // for one warp, but a range too short to hold a few records could not infer its start)
        const uint64_t nr = (uint64_t)ctx->grid * stream_warps(ctx->nchunk, (sh->flags & FQB_F_HIST) != 0);
        srange_bytes = sh->n_own / nr / 16 * 16;
        if (srange_bytes < 8192) srange_bytes = 8192;
        const uint64_t live = std::max<uint64_t>(1, std::min<uint64_t>(nr, sh->n_own / srange_bytes));
        n_sranges = live;
        if (fast && want_index) {
            // (lines of >= 16 bytes on average; the last range is longer by the remainder)
            const uint64_t rem = sh->n_own > live * srange_bytes ? sh->n_own - live * srange_bytes : 0;
            stage_share = ((srange_bytes + rem) / 16 + 64 + 3) / 4 * 4;   // (a multiple of 4: the descriptors behind are uint4)
            // window descriptors (4 words each): one per window; a window that consumes less than 256 bytes on
            // average (records of a few bytes) overflows them -- the exact path writes the index then
            desc_cap = (srange_bytes + rem) / 256 + 64;
            const uint64_t need = live * (stage_share + 4 * desc_cap);
            if (need > ctx->index_stage_cap) {
                if (ctx->d_index_stage) {
                    // (an earlier parse of this context may still be using the staging area on ANOTHER stream)
                    CK(cudaDeviceSynchronize());
                    CK(cudaFree(ctx->d_index_stage));
                    ctx->d_index_stage = nullptr;
                    ctx->index_stage_cap = 0;
                }
                CK(cudaMalloc(&ctx->d_index_stage, need * 4 + 64));
                ctx->index_stage_cap = need;
            }
        }
    }
    ScanParams p;
    memset(&p, 0, sizeof p);
    p.data = sh->d_bytes;
    p.n_own = sh->n_own;
    p.n_avail = sh->n_avail;
    p.stream_offset = sh->stream_offset;
    p.line_base = sh->line_base;
    p.carry = carry;
    p.flags = (sh->flags & (F_HIST | F_INDEX | F_LINE_START | F_EOF | F_FRONT16 | F_INFER_START)) | (carry ? F_CARRY : 0);
    if (allow_retry && !carry && !getenv("FQB_NO_RETRY")) p.flags |= F_CAN_RETRY;
    if ((p.flags & F_INFER_START) && carry) return FQB_E_ARG;   // a stream knows its line numbers
    if ((p.flags & F_FRONT16) && (p.flags & F_LINE_START)) return FQB_E_ARG;
    p.max_len = ctx->P;
    p.ntiles = (uint32_t)ntiles64;
    p.tiles_per_cta = (uint32_t)((ntiles64 + ctx->grid - 1) / ctx->grid);
    p.ranges = ctx->d_ranges;
    p.nranges = (uint32_t)ctx->grid;
    p.index_stage = ctx->d_index_stage;
    p.desc = ctx->d_index_stage ? ctx->d_index_stage + n_sranges * stage_share : nullptr;
    p.desc_cap = desc_cap;
    p.sranges = ctx->d_sranges;
    p.srange_bytes = srange_bytes;
    p.n_sranges = (uint32_t)n_sranges;
    p.stage_share = stage_share;
    p.index = sh->d_index;
    p.index_cap = sh->index_cap;
    p.res = ctx->d_res;
    p.stats = reinterpret_cast<unsigned long long*>(ctx->d_stats);
    p.seqraw = reinterpret_cast<unsigned long long*>(ctx->d_seqraw);
    p.trace = ctx->d_trace;
    if (ctx->d_trace) CK(cudaMemsetAsync(ctx->d_trace, 0, (size_t)ctx->num_sms * TRACE_K * 16 * 8, st));

    fq_init_kernel<<<1, PROBE_THREADS, 0, st>>>(ctx->d_res, fast ? 0 : 1, (int)(sh->line_base & 3), sh->d_bytes, sh->n_avail,
                                     (fast && !getenv("FQB_NO_VAR")) ? ((sh->flags & FQB_F_HIST) ? 1 : 2) : 0);
    CK(cudaGetLastError());
    CK(cudaMemsetAsync(ctx->d_stats, 0, (ctx->nwords + 8 * (size_t)FQB_MAX_WORLD) * 8, st));   // block + outcome slots
    CK(cudaMemsetAsync(ctx->d_seqraw, 0, (size_t)ctx->P * 256 * 8, st));
    ctx->launches += 1;
    if (p.ntiles) {
        // speculative kernel + the check of its record chains (the events bracket the dominant kernel:
        // the speculative one, or the exact one when the speculative one is not used for this shape)
        if (fast) {
            if (timed) CK(cudaEventRecord(ctx->ev0, st));
            CK(launch_stream(p, ctx->nchunk, ctx->grid, st));
            if (timed) CK(cudaEventRecord(ctx->ev1, st));
            CK(launch_stream_verify(p, carry, st));
            ctx->launches += 3;   // (both variants of the speculative kernel + the chain check)
        }
        // exact path (every launch of it returns at once unless res->spec_fail is set): newline counts
        // of the CTA ranges -> exact line numbers -> exact kernel.  It needs the exact line_base: with
        // FQB_F_INFER_START a speculative launch that did not deliver ends in FQB_E_PHASE instead.
        if (!(p.flags & F_INFER_START)) {
            CK(launch_rerun_reset(p, 0, st));
            CK(launch_range_count(p, carry, ctx->grid, (unsigned long long)p.tiles_per_cta * tile_bytes, st));
            if (timed && !fast) CK(cudaEventRecord(ctx->ev0, st));
            CK(launch_scan(p, ctx->nchunk, ctx->grid, st));
            if (timed && !fast) CK(cudaEventRecord(ctx->ev1, st));
            // classify the first bad record; redo restricted to the records before it (each() delivers those)
            CK(launch_diagnose(p, carry, st));
            CK(launch_rerun_reset(p, 1, st));
            ScanParams p2 = p;
            p2.flags |= F_RERUN;
            p2.trace = nullptr;
            CK(launch_scan(p2, ctx->nchunk, ctx->grid, st));
            ctx->launches += 7;
        }
        if (timed) ctx->ev_valid = fast || !(p.flags & F_INFER_START);
        if (timed) ctx->ev_index_valid = false;
        if (fast && want_index) {
            if (timed) CK(cudaEventRecord(ctx->ev2, st));
            CK(launch_stream_compact(p, carry, ctx->grid, st));
            if (timed) {
                CK(cudaEventRecord(ctx->ev3, st));
                ctx->ev_index_valid = true;
            }
            ctx->launches += 1;
        }
        if (fast && (p.flags & F_EOF) && !(p.flags & F_INFER_START)) {
            // a bad record the speculative kernel found at the end of the stream: the line ends behind it
            CK(launch_tail_index(p, carry, st));
            ctx->launches += 1;
        }
    }
    CK(launch_finalize(p, carry, reinterpret_cast<unsigned long long*>(total), carry ? nullptr : ctx->d_pub,
                       carry ? nullptr : reinterpret_cast<unsigned long long*>(ctx->d_stats) + ctx->nwords + 8 * (size_t)ctx->rank, st));
    ctx->launches += 1;
    return FQB_OK;
}

int fqb_parse_device(fqb_ctx* ctx, const fqb_shard* shard, void* stream)
{
    if (!ctx || !shard) return FQB_E_ARG;
    ctx->last_off = shard->stream_offset;
    ctx->last_shard = *shard;
    ctx->retries = 0;
    return enqueue_parse(ctx, shard, static_cast<cudaStream_t>(stream), nullptr, nullptr, true, true);
}

static void fill_result(const DevResult* r, uint64_t stream_offset_of_tail_base, fqb_result* res)
{
    res->status = r->status;
    res->finished = r->finished;
    res->n_records = r->n_records;
    res->n_lines = r->n_lines;
    res->err_offset = r->err_offset;
    res->tail_offset = r->tail_start == NONE64 ? UINT64_MAX : stream_offset_of_tail_base + r->tail_start;
    res->line_phase = (uint32_t)r->line_phase;
    res->reserved = 0;
}

int fqb_fetch(fqb_ctx* ctx, void* stream, fqb_result* res, uint64_t* host_stats)
{
    if (!ctx || !res) return FQB_E_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->h_res, ctx->d_res, sizeof(DevResult), cudaMemcpyDeviceToHost, st));
    if (host_stats) CK(cudaMemcpyAsync(host_stats, ctx->d_stats, ctx->nwords * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (ctx->h_res->spec_retry && !ctx->h_res->spec_fail) {
        // A record in the middle of the shard fails validation, and the speculative kernel has proved that everything
        // in front of it (offset X) is delimited as a sequential parse would.  Its counters, though, include records
        // BEHIND the bad one, which each() never delivers (src/lib.rs:226-237).  So: the bytes [0, X) once more, as a
        // shard of their own (a clean parse of a shorter input), then the bad record is classified in the reference's
        // check order and the '\n' behind it are counted.  Twice a clean parse at most, instead of the exact path.
        const unsigned long long X = ctx->h_res->spec_bad;
        fqb_shard sh = ctx->last_shard;
        sh.n_own = sh.n_avail = X;
        int rc = enqueue_parse(ctx, &sh, st, nullptr, nullptr, false, false);
        if (rc) return rc;
        ScanParams p;
        memset(&p, 0, sizeof p);
        p.data = ctx->last_shard.d_bytes;
        p.n_own = ctx->last_shard.n_own;
        p.n_avail = ctx->last_shard.n_avail;
        p.stream_offset = ctx->last_shard.stream_offset;
        p.flags = ctx->last_shard.flags & (F_EOF | F_INFER_START);
        p.res = ctx->d_res;
        fq_retry_mark_kernel<<<1, 1, 0, st>>>(ctx->d_res, X);
        CK(launch_diagnose(p, nullptr, st));
        fq_tail_lines_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(p.data, X, p.n_own, ctx->d_res);
        fq_republish_kernel<<<1, 1, 0, st>>>(p, ctx->d_pub, reinterpret_cast<unsigned long long*>(ctx->d_stats) + ctx->nwords + 8 * (size_t)ctx->rank);
        CK(cudaGetLastError());
        ctx->launches += 4;
        ctx->retries = 1;
        CK(cudaMemcpyAsync(ctx->h_res, ctx->d_res, sizeof(DevResult), cudaMemcpyDeviceToHost, st));
        if (host_stats) CK(cudaMemcpyAsync(host_stats, ctx->d_stats, ctx->nwords * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    if (ctx->d_trace) {
        std::vector<unsigned long long> h((size_t)ctx->num_sms * TRACE_K * 16);
        CK(cudaMemcpy(h.data(), ctx->d_trace, h.size() * 8, cudaMemcpyDeviceToHost));
        if (FILE* f = fopen(ctx->trace_path.c_str(), "wb")) {
            fwrite(h.data(), 8, h.size(), f);
            fclose(f);
        }
    }
    if (getenv("FQB_DEBUG"))
        fprintf(stderr, "fastq_b200: spec_fail=%d status=%d windows predicted=%llu scanned=%llu\n", ctx->h_res->spec_fail,
                ctx->h_res->status, ctx->h_res->n_win_pred, ctx->h_res->n_win_scan);
    fill_result(ctx->h_res, ctx->last_off, res);
    return FQB_OK;
}

uint64_t* fqb_device_stats(fqb_ctx* ctx) { return ctx ? ctx->d_stats : nullptr; }
uint64_t* fqb_device_result(fqb_ctx* ctx) { return ctx ? reinterpret_cast<uint64_t*>(ctx->d_pub) : nullptr; }
uint64_t fqb_launch_count(fqb_ctx* ctx) { return ctx ? ctx->launches : 0; }

float fqb_last_scan_ms(fqb_ctx* ctx)
{
    if (!ctx || !ctx->ev_valid) return -1.f;
    if (cudaEventSynchronize(ctx->ev1) != cudaSuccess) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) != cudaSuccess) return -1.f;
    return ms;
}

float fqb_last_index_ms(fqb_ctx* ctx)
{
    if (!ctx || !ctx->ev_index_valid) return 0.f;
    if (cudaEventSynchronize(ctx->ev3) != cudaSuccess) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3) != cudaSuccess) return -1.f;
    return ms;
}

int fqb_count_lines_device(fqb_ctx* ctx, const uint8_t* d_bytes, uint64_t n, void* stream)
{
    if (!ctx || (n && (!d_bytes || (reinterpret_cast<uintptr_t>(d_bytes) & 15)))) return FQB_E_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->d_linecount, 0, 8, st));
    if (n) {
        CK(launch_count(d_bytes, n, ctx->d_linecount, ctx->num_sms * 8, st));
        ctx->launches += 1;
    }
    return FQB_OK;
}

int fqb_fetch_line_count(fqb_ctx* ctx, void* stream, uint64_t* n_lines)
{
    if (!ctx || !n_lines) return FQB_E_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaMemcpyAsync(ctx->h_linecount, ctx->d_linecount, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *n_lines = *ctx->h_linecount;
    return FQB_OK;
}

int fqb_filter_device(fqb_ctx* ctx, const uint8_t* d_bytes, uint64_t stream_offset, const uint32_t* d_index,
                      uint64_t n_records, uint64_t first_offset, uint32_t mode, uint8_t* d_out, uint64_t out_cap,
                      void* stream)
{
    if (!ctx || mode > FQB_KEEP_DNAN || first_offset < stream_offset) return FQB_E_ARG;
    if (n_records && (!d_bytes || !d_index || (reinterpret_cast<uintptr_t>(d_bytes) & 15) ||
                      (reinterpret_cast<uintptr_t>(d_index) & 3) || (out_cap && !d_out)))
        return FQB_E_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(ctx->device));
    const size_t nblk = (size_t)((n_records + 255) / 256);
    if (!ctx->d_fmisc) {
        CK(cudaMalloc(&ctx->d_fmisc, (1 + FILTER_MAX_WRAPS + 4 + 1) * 8));
        CK(cudaHostAlloc(&ctx->h_fres, 4 * 8, cudaHostAllocDefault));
    }
    if (nblk > ctx->fblk_cap) {   // the previous call may still be running on another stream: wait for it
        CK(cudaDeviceSynchronize());
        cudaFree(ctx->d_fblk);
        ctx->d_fblk = nullptr;
        ctx->fblk_cap = 0;
        CK(cudaMalloc(&ctx->d_fblk, nblk * 8));
        ctx->fblk_cap = nblk;
    }
    FilterParams p;
    p.data = d_bytes;
    p.index = d_index;
    p.n_records = n_records;
    p.stream_offset = stream_offset;
    p.first_offset = first_offset;
    p.mode = mode;
    p.blk = ctx->d_fblk;
    p.ticket = reinterpret_cast<unsigned int*>(ctx->d_fmisc + 1 + FILTER_MAX_WRAPS + 4);
    p.wraps = ctx->d_fmisc;
    p.out = d_out;
    p.out_cap = out_cap;
    p.result = ctx->d_fmisc + 1 + FILTER_MAX_WRAPS;
    CK(launch_filter(p, ctx->num_sms, st));
    ctx->launches += filter_launches(n_records);
    return FQB_OK;
}

int fqb_fetch_filter(fqb_ctx* ctx, void* stream, uint64_t* n_kept, uint64_t* out_bytes)
{
    if (!ctx || !n_kept || !out_bytes || !ctx->d_fmisc) return FQB_E_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaMemcpyAsync(ctx->h_fres, ctx->d_fmisc + 1 + FILTER_MAX_WRAPS, 4 * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *n_kept = ctx->h_fres[0];
    *out_bytes = ctx->h_fres[1];
    if (ctx->h_fres[2] > FILTER_MAX_WRAPS) return FQB_E_ARG;   // not an index of increasing offsets
    return FQB_OK;
}

int fqb_last_path(fqb_ctx* ctx, uint64_t out[3])
{
    if (!ctx || !out || !ctx->h_res) return FQB_E_ARG;
    out[0] = (ctx->h_res->spec_fail ? 1 : 0) | (ctx->retries ? 2 : 0);   // bit 1: the part in front of a bad record was parsed twice
    out[1] = ctx->h_res->n_win_pred;
    out[2] = ctx->h_res->n_win_scan;
    return FQB_OK;
}

int fqb_host_alloc(uint64_t bytes, void** out)
{
    if (!out) return FQB_E_ARG;
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return FQB_E_NOMEM;
    }
    *out = p;
    return FQB_OK;
}

void fqb_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

int fqb_synth_fixed_device(uint8_t* d_out, uint64_t n, uint64_t byte_off, uint32_t L, uint64_t seed, void* stream)
{
    if ((n && !d_out) || L == 0 || L > 1023) return FQB_E_ARG;
    return launch_synth_fixed(d_out, n, byte_off, L, seed, static_cast<cudaStream_t>(stream)) == cudaSuccess ? FQB_OK
                                                                                                             : FQB_E_CUDA;
}

int fqb_synth_var_device(uint8_t* d_out, const uint64_t* d_rec_off, uint64_t first, uint64_t count, uint64_t seed,
                         void* stream)
{
    if (count && (!d_out || !d_rec_off)) return FQB_E_ARG;
    return launch_synth_var(d_out, reinterpret_cast<const unsigned long long*>(d_rec_off), first, count, seed,
                            static_cast<cudaStream_t>(stream)) == cudaSuccess
               ? FQB_OK
               : FQB_E_CUDA;
}

int fqb_synth_var_sizes_device(uint64_t* d_sizes, uint64_t first, uint64_t count, uint64_t seed, void* stream)
{
    if (count && !d_sizes) return FQB_E_ARG;
    return launch_synth_var_sizes(reinterpret_cast<unsigned long long*>(d_sizes), first, count, seed,
                                  static_cast<cudaStream_t>(stream)) == cudaSuccess
               ? FQB_OK
               : FQB_E_CUDA;
}

// ==========================================================================================
// streaming ring
// ==========================================================================================
// Device ring: n_dev slots of slot_bytes laid out back to back, plus MAXREC bytes behind the last
// slot that mirror the head of slot 0, so that every chunk is followed in memory by the head of
// its successor.  Chunk i is parsed once chunk i+1 has landed (or EOF is known): records are
// owned by the chunk they start in and are read through into the next one -- no partial-record
// memmove as in Buffer::clean (src/buffer.rs:51-72), no host round trip between chunks.

static int stream_alloc(fqb_ctx* ctx)
{
    if (ctx->d_ring) return FQB_OK;
    CK(cudaSetDevice(ctx->device));
    ctx->n_dev = ctx->n_slots + 1;
    CK(cudaMalloc(&ctx->d_ring, (size_t)ctx->n_dev * ctx->slot_bytes + MAXREC + 64));
    CK(cudaMalloc(&ctx->d_total, ctx->nwords * 8));
    CK(cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->s_comp, cudaStreamNonBlocking));
    // the pinned host slots lie back to back like the device slots, with MAXREC bytes behind the last one that
    // mirror the head of slot 0: a record that starts in one slot and ends in the next is contiguous in host
    // memory too (batch mode hands out records that borrow the ring, fqb_next_batch)
    CK(cudaHostAlloc(&ctx->h_ring, (size_t)ctx->n_slots * ctx->slot_bytes + MAXREC + 64, cudaHostAllocDefault));
    ctx->slots.resize(ctx->n_slots);
    for (uint32_t k = 0; k < ctx->n_slots; ++k) {
        ctx->slots[k].h_pinned = ctx->h_ring + (size_t)k * ctx->slot_bytes;
        CK(cudaEventCreateWithFlags(&ctx->slots[k].copied, cudaEventDisableTiming));
    }
    ctx->dslots.resize(ctx->n_dev);
    for (auto& d : ctx->dslots) CK(cudaEventCreateWithFlags(&d.done, cudaEventDisableTiming));
    return FQB_OK;
}

int fqb_stream_begin(fqb_ctx* ctx, uint32_t flags)
{
    if (!ctx) return FQB_E_ARG;
    if (ctx->streaming) return FQB_E_STATE;
    int rc = stream_alloc(ctx);
    if (rc) return rc;
    CK(cudaMemsetAsync(ctx->d_total, 0, ctx->nwords * 8, ctx->s_comp));
    CK(cudaMemsetAsync(ctx->d_carry, 0, sizeof(DevCarry), ctx->s_comp));
    ctx->streaming = true;
    ctx->stream_flags = flags & (FQB_F_HIST | FQB_F_INDEX);
    ctx->stream_partial = false;
    ctx->chunk_no = 0;
    ctx->fill = 0;
    ctx->stream_pos = 0;
    ctx->acquired = false;
    ctx->have_pending = false;
    ctx->last_byte_nl = true;
    ctx->host_index_n = 0;
    for (auto& s : ctx->slots) s.copied_pending = false;
    for (auto& d : ctx->dslots) d.done_pending = false;
    return FQB_OK;
}

// launch the parse of the pending chunk; `next_bytes` = bytes of its successor already resident
static int launch_pending(fqb_ctx* ctx, uint64_t next_bytes, bool eof)
{
    const uint64_t i = ctx->pending_chunk;
    const uint32_t ds = (uint32_t)(i % ctx->n_dev);
    fqb_shard sh;
    memset(&sh, 0, sizeof sh);
    sh.d_bytes = ctx->d_ring + (size_t)ds * ctx->slot_bytes;
    sh.n_own = ctx->pending_bytes;
    sh.n_avail = ctx->pending_bytes + std::min<uint64_t>(next_bytes, MAXREC);
    sh.stream_offset = ctx->pending_off;
    sh.flags = ctx->stream_flags | (ctx->pending_line_start ? FQB_F_LINE_START : 0) |
               ((eof && next_bytes <= MAXREC && !ctx->stream_partial) ? FQB_F_EOF : 0);
    if (ctx->stream_flags & FQB_F_INDEX) {
        DevSlot& d = ctx->dslots[ds];
        if (!d.d_index) CK(cudaMalloc(&d.d_index, (size_t)ctx->slot_bytes * 4));
        sh.d_index = d.d_index;
        sh.index_cap = ctx->slot_bytes;
    }
    int rc = enqueue_parse(ctx, &sh, ctx->s_comp, ctx->d_carry, ctx->d_total, false);
    if (rc) return rc;
    if (ctx->batch_mode) {
        // batch mode: outcome + line ends of this chunk into the pinned buffers of its host slot, then the event the
        // consumer waits for
        fqb_ctx::BChunk& bc = ctx->bchunks[i % ctx->n_slots];
        fq_batch_out_kernel<<<ctx->num_sms * 2, 256, 0, ctx->s_comp>>>(sh.d_index, ctx->d_res, bc.h_info_dev, bc.h_index_dev,
                                                                       bc.index_cap);
        CK(cudaGetLastError());
        ctx->launches += 1;
        CK(cudaEventRecord(bc.parsed, ctx->s_comp));
        {
            std::lock_guard<std::mutex> lk(ctx->mu);
            bc.d_index = sh.d_index;
            bc.launched = true;
        }
        ctx->cv.notify_all();
    } else if ((ctx->stream_flags & FQB_F_INDEX) && ctx->host_index_dev) {
        // generic-closure path, pinned index buffer: written by the device, in stream order
        fq_index_out_kernel<<<ctx->num_sms * 2, 256, 0, ctx->s_comp>>>(sh.d_index, ctx->d_res, ctx->d_carry,
                                                                       ctx->host_index_dev, ctx->host_index_cap);
        CK(cudaGetLastError());
        ctx->launches += 1;
    } else if (ctx->stream_flags & FQB_F_INDEX) {
        // pageable index buffer: bring this chunk's line ends to the host (needs the count first)
        CK(cudaMemcpyAsync(ctx->h_res, ctx->d_res, sizeof(DevResult), cudaMemcpyDeviceToHost, ctx->s_comp));
        CK(cudaStreamSynchronize(ctx->s_comp));
        uint64_t nl = ctx->h_res->n_lines;
        uint64_t room = ctx->host_index_cap > ctx->host_index_n ? ctx->host_index_cap - ctx->host_index_n : 0;
        uint64_t take = std::min(nl, room);
        if (take && ctx->host_index)
            CK(cudaMemcpy(ctx->host_index + ctx->host_index_n, sh.d_index, take * 4, cudaMemcpyDeviceToHost));
        ctx->host_index_n += take;
    }
    CK(cudaEventRecord(ctx->dslots[ds].done, ctx->s_comp));
    ctx->dslots[ds].done_pending = true;
    ctx->have_pending = false;
    return FQB_OK;
}

// copy one chunk (host pointer must stay valid until `copied` fires) into the device ring
static int submit_chunk(fqb_ctx* ctx, const uint8_t* h_src, uint64_t n, cudaEvent_t copied)
{
    const uint64_t i = ctx->chunk_no;
    const uint32_t ds = (uint32_t)(i % ctx->n_dev);
    uint8_t* dst = ctx->d_ring + (size_t)ds * ctx->slot_bytes;
    // the device slot is reused every n_dev chunks: wait for the kernels that read it, and for the
    // kernel of the chunk before it (it reads this slot's head as its halo)
    if (ctx->dslots[ds].done_pending) CK(cudaStreamWaitEvent(ctx->s_copy, ctx->dslots[ds].done, 0));
    const uint32_t prev = (uint32_t)((i + ctx->n_dev - 1) % ctx->n_dev);
    if (i >= ctx->n_dev && ctx->dslots[prev].done_pending)
        CK(cudaStreamWaitEvent(ctx->s_copy, ctx->dslots[prev].done, 0));
    CK(cudaMemcpyAsync(dst, h_src, n, cudaMemcpyHostToDevice, ctx->s_copy));
    if (ds == 0 && i > 0) {
        // mirror the head of slot 0 behind the last slot so chunk i-1 sees its successor contiguously
        uint64_t m = std::min<uint64_t>(n, MAXREC);
        CK(cudaMemcpyAsync(ctx->d_ring + (size_t)ctx->n_dev * ctx->slot_bytes, dst, m, cudaMemcpyDeviceToDevice,
                           ctx->s_copy));
    }
    CK(cudaEventRecord(copied, ctx->s_copy));
    CK(cudaStreamWaitEvent(ctx->s_comp, copied, 0));
    if (ctx->have_pending) {
        // only the last chunk of a stream is ever shorter than a slot
        int rc = launch_pending(ctx, n, n < ctx->slot_bytes);
        if (rc) return rc;
    }
    if (ctx->batch_mode) {
        fqb_ctx::BChunk& bc = ctx->bchunks[i % ctx->n_slots];
        std::lock_guard<std::mutex> lk(ctx->mu);
        bc.no = i;
        bc.off = ctx->stream_pos;
        bc.n_bytes = n;
        bc.line_start = ctx->last_byte_nl;
        bc.launched = false;
        bc.released = false;
        // the head of slot 0 again behind the last slot: the last record of the chunk before it reads on there
        if (i % ctx->n_slots == 0 && i > 0)
            memcpy(ctx->h_ring + (size_t)ctx->n_slots * ctx->slot_bytes, ctx->h_ring, (size_t)std::min<uint64_t>(n, MAXREC));
    }
    ctx->have_pending = true;
    ctx->pending_line_start = ctx->last_byte_nl;
    ctx->last_byte_nl = h_src[n - 1] == '\n';
    ctx->pending_chunk = i;
    ctx->pending_bytes = n;
    ctx->pending_off = ctx->stream_pos;
    ctx->chunk_no = i + 1;
    ctx->stream_pos += n;
    return FQB_OK;
}

int fqb_stream_acquire(fqb_ctx* ctx, uint8_t** pinned, uint64_t* cap)
{
    if (!ctx || !pinned || !cap) return FQB_E_ARG;
    if (!ctx->streaming || ctx->acquired) return FQB_E_STATE;
    Slot& s = ctx->slots[ctx->chunk_no % ctx->n_slots];
    if (ctx->fill == 0 && s.copied_pending) {
        CK(cudaEventSynchronize(s.copied));  // empty_recv.recv(): wait until the slot has been drained
        s.copied_pending = false;
    }
    if (ctx->batch_mode && ctx->fill == 0) {
        // the slot still backs the batch of the chunk it held before: that batch must have been given back
        // (fqb_release_batch).  The batch of the chunk in front of THAT one reads its last record on into this
        // slot (or into the mirror of slot 0); it was given back before the slot in front of this one was refilled --
        // slots are filled in order -- so this one condition covers it.
        std::unique_lock<std::mutex> lk(ctx->mu);
        const uint32_t k = (uint32_t)(ctx->chunk_no % ctx->n_slots);
        ctx->cv.wait(lk, [&] { return ctx->batch_cancel || ctx->bchunks[k].released; });
        if (ctx->batch_cancel) return FQB_E_CANCELLED;
    }
    *pinned = s.h_pinned + ctx->fill;
    *cap = ctx->slot_bytes - ctx->fill;
    ctx->acquired = true;
    return FQB_OK;
}

int fqb_stream_submit(fqb_ctx* ctx, uint64_t n_valid)
{
    if (!ctx) return FQB_E_ARG;
    if (!ctx->streaming || !ctx->acquired) return FQB_E_STATE;
    if (n_valid > ctx->slot_bytes - ctx->fill) return FQB_E_ARG;
    ctx->acquired = false;
    ctx->fill += n_valid;
    if (ctx->fill < ctx->slot_bytes) return FQB_OK;  // short read: the next acquire tops the slot up
    Slot& s = ctx->slots[ctx->chunk_no % ctx->n_slots];
    int rc = submit_chunk(ctx, s.h_pinned, ctx->fill, s.copied);
    s.copied_pending = true;
    ctx->fill = 0;
    return rc;
}

static int stream_drain(fqb_ctx* ctx, fqb_result* res, uint64_t* host_stats)
{
    if (ctx->have_pending) {
        int rc = launch_pending(ctx, 0, true);
        if (rc) return rc;
    }
    CK(cudaMemcpyAsync(ctx->h_carry, ctx->d_carry, sizeof(DevCarry), cudaMemcpyDeviceToHost, ctx->s_comp));
    if (host_stats) CK(cudaMemcpyAsync(host_stats, ctx->d_total, ctx->nwords * 8, cudaMemcpyDeviceToHost, ctx->s_comp));
    CK(cudaStreamSynchronize(ctx->s_comp));
    CK(cudaStreamSynchronize(ctx->s_copy));
    res->status = ctx->h_carry->status;
    res->finished = ctx->h_carry->status == 0 && ctx->h_carry->tail_plus1 == 0;
    res->n_records = ctx->h_carry->n_records;
    res->n_lines = ctx->h_carry->n_lines;
    res->err_offset = ctx->h_carry->err_offset;
    res->tail_offset = ctx->h_carry->tail_plus1 ? ctx->h_carry->tail_plus1 - 1 : UINT64_MAX;
    res->line_phase = 0;
    res->reserved = 0;
    if (ctx->host_index_dev) ctx->host_index_n = std::min<uint64_t>(ctx->h_carry->n_lines, ctx->host_index_cap);
    ctx->streaming = false;
    return FQB_OK;
}

int fqb_stream_finish(fqb_ctx* ctx, fqb_result* res, uint64_t* host_stats)
{
    if (!ctx || !res) return FQB_E_ARG;
    if (!ctx->streaming || ctx->acquired) return FQB_E_STATE;
    if (ctx->fill) {  // partially filled last slot
        Slot& s = ctx->slots[ctx->chunk_no % ctx->n_slots];
        int rc = submit_chunk(ctx, s.h_pinned, ctx->fill, s.copied);
        s.copied_pending = true;
        ctx->fill = 0;
        if (rc) {
            ctx->streaming = false;
            return rc;
        }
    }
    int rc = stream_drain(ctx, res, host_stats);
    ctx->streaming = false;
    return rc;
}

int fqb_parse_host(fqb_ctx* ctx, const uint8_t* bytes, uint64_t n, uint64_t stream_offset, uint32_t flags,
                   fqb_result* res, uint64_t* host_stats, uint32_t* host_index, uint64_t index_cap, uint64_t* n_index)
{
    if (!ctx || !res || (n && !bytes)) return FQB_E_ARG;
    if ((flags & FQB_F_INDEX) && index_cap && !host_index) return FQB_E_ARG;
    int rc = fqb_stream_begin(ctx, flags);
    if (rc) return rc;
    ctx->stream_pos = stream_offset;   // offsets reported (and the too-long rule, rec_window()) are stream offsets
    ctx->host_index = host_index;
    ctx->host_index_cap = host_index ? index_cap : 0;
    ctx->stream_partial = (flags & FQB_F_PARTIAL) != 0;
    ctx->host_index_dev = nullptr;
    if (ctx->host_index_cap) {
        cudaPointerAttributes ia;
        void* dev = nullptr;
        if (cudaPointerGetAttributes(&ia, host_index) == cudaSuccess && ia.type == cudaMemoryTypeHost &&
            cudaHostGetDevicePointer(&dev, host_index, 0) == cudaSuccess)
            ctx->host_index_dev = static_cast<uint32_t*>(dev);
        else
            cudaGetLastError();
    }
    // pinned caller memory is copied from directly; pageable memory goes through the pinned slots
    cudaPointerAttributes attr;
    bool pinned = false;
    if (n && cudaPointerGetAttributes(&attr, bytes) == cudaSuccess)
        pinned = attr.type == cudaMemoryTypeHost;
    else
        cudaGetLastError();
    uint64_t off = 0;
    while (off < n && rc == FQB_OK) {
        const uint64_t len = std::min<uint64_t>(ctx->slot_bytes, n - off);
        Slot& s = ctx->slots[ctx->chunk_no % ctx->n_slots];
        if (s.copied_pending) {
            cudaError_t e = cudaEventSynchronize(s.copied);
            if (e != cudaSuccess) {
                rc = fail(ctx, e, "cudaEventSynchronize");
                break;
            }
            s.copied_pending = false;
        }
        const uint8_t* src = bytes + off;
        if (!pinned) {
            // stage with a few threads: one memcpy stream cannot keep PCIe busy
            const int nt = 4;
            std::vector<std::thread> th;
            for (int k = 0; k < nt; ++k) {
                uint64_t a = len * k / nt, b = len * (k + 1) / nt;
                th.emplace_back([=, &s] { memcpy(s.h_pinned + a, src + a, b - a); });
            }
            for (auto& t : th) t.join();
            src = s.h_pinned;
        }
        rc = submit_chunk(ctx, src, len, s.copied);
        s.copied_pending = true;
        off += len;
    }
    if (rc == FQB_OK) rc = stream_drain(ctx, res, host_stats);
    ctx->streaming = false;
    if (n_index) *n_index = ctx->host_index_n;
    ctx->host_index = nullptr;
    ctx->host_index_dev = nullptr;
    ctx->host_index_cap = 0;
    return rc;
}


// ==========================================================================================
// batch mode: the generic-closure path, asynchronous (RecordSet hand-off, src/lib.rs:306-426, 509-566)
// ==========================================================================================
// Producer thread:  fqb_batch_begin, then fqb_stream_acquire / fqb_stream_submit per read() as in the streaming
//                   ring (thread_reader's protocol), fqb_batch_close at the end of the input.
// Consumer thread:  fqb_next_batch hands out, chunk by chunk, the records the GPU has delimited -- bytes borrowed
//                   from the pinned ring + their line ends -- while the producer keeps reading and the GPU keeps
//                   delimiting the chunks behind; fqb_release_batch gives the memory back (RecordSet dropped).
// A batch holds the records that START in its chunk; the last of them may end in the next chunk (contiguous in the
// host ring), and its last line ends then come from that chunk's index -- which is why batch i is handed out once
// chunk i + 1 has been delimited as well (or the input has ended).

static int batch_alloc(fqb_ctx* ctx)
{
    if (!ctx->bchunks.empty()) return FQB_OK;
    ctx->bchunks.resize(ctx->n_slots);
    for (auto& b : ctx->bchunks) {
        CK(cudaEventCreateWithFlags(&b.parsed, cudaEventDisableTiming));
        CK(cudaHostAlloc(&b.h_info, BATCH_INFO_WORDS * 8, cudaHostAllocMapped));
        b.index_cap = (size_t)(ctx->slot_bytes / 8) + 64;       // lines of >= 8 bytes on average; grown on demand
        CK(cudaHostAlloc(&b.h_index, b.index_cap * 4, cudaHostAllocMapped));
        void* d = nullptr;
        CK(cudaHostGetDevicePointer(&d, b.h_info, 0));
        b.h_info_dev = static_cast<unsigned long long*>(d);
        CK(cudaHostGetDevicePointer(&d, b.h_index, 0));
        b.h_index_dev = static_cast<uint32_t*>(d);
    }
    return FQB_OK;
}

int fqb_batch_begin(fqb_ctx* ctx, uint32_t flags)
{
    if (!ctx) return FQB_E_ARG;
    if (ctx->streaming || ctx->batch_mode) return FQB_E_STATE;
    if (ctx->n_slots < 4) {
        ctx->err = "batch mode needs fqb_config.n_slots >= 4";
        return FQB_E_STATE;
    }
    int rc = fqb_stream_begin(ctx, (flags & FQB_F_HIST) | FQB_F_INDEX);
    if (rc) return rc;
    rc = batch_alloc(ctx);
    if (rc) {
        ctx->streaming = false;
        return rc;
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (auto& b : ctx->bchunks) {
        b.no = ~0ull;
        b.launched = false;
        b.released = true;
    }
    ctx->batch_mode = true;
    ctx->batch_closed = ctx->batch_cancel = ctx->batch_failed = false;
    ctx->b_total = ~0ull;
    ctx->b_next = ctx->b_held = ctx->b_records = 0;
    ctx->b_status = 0;
    return FQB_OK;
}

int fqb_batch_close(fqb_ctx* ctx)
{
    if (!ctx) return FQB_E_ARG;
    if (!ctx->batch_mode || ctx->acquired || ctx->batch_closed) return FQB_E_STATE;
    int rc = FQB_OK;
    if (ctx->fill) {   // partially filled last slot
        Slot& s = ctx->slots[ctx->chunk_no % ctx->n_slots];
        rc = submit_chunk(ctx, s.h_pinned, ctx->fill, s.copied);
        s.copied_pending = true;
        ctx->fill = 0;
    }
    if (rc == FQB_OK && ctx->have_pending) rc = launch_pending(ctx, 0, true);
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->b_total = ctx->chunk_no;
        ctx->batch_closed = true;
        if (rc != FQB_OK) ctx->batch_failed = true;
    }
    ctx->cv.notify_all();
    return rc;
}

int fqb_next_batch(fqb_ctx* ctx, fqb_batch* out)
{
    if (!ctx || !out) return FQB_E_ARG;
    if (!ctx->batch_mode) return FQB_E_STATE;
    memset(out, 0, sizeof *out);
    std::unique_lock<std::mutex> lk(ctx->mu);
    if (ctx->b_status != 0) return FQB_E_STATE;                       // the stream has ended with an error batch
    if (ctx->b_held + 3 > ctx->n_slots) {
        ctx->err = "fqb_next_batch: release a batch first (at most n_slots - 3 may be held)";
        return FQB_E_STATE;
    }
    const uint64_t i = ctx->b_next;
    // chunk i delimited, and chunk i + 1 too unless chunk i is the last one
    auto ready = [&](uint64_t c) { return ctx->bchunks[c % ctx->n_slots].no == c && ctx->bchunks[c % ctx->n_slots].launched; };
    ctx->cv.wait(lk, [&] {
        if (ctx->batch_cancel || ctx->batch_failed) return true;
        if (ctx->batch_closed && i >= ctx->b_total) return true;
        if (!ready(i)) return false;
        return ready(i + 1) || (ctx->batch_closed && i + 1 >= ctx->b_total);
    });
    if (ctx->batch_cancel) return FQB_E_CANCELLED;
    if (ctx->batch_failed) return FQB_E_CUDA;
    if (i >= ctx->b_total) {                                          // an empty stream, or everything handed out
        out->last = 1;
        out->first_record = ctx->b_records;
        out->token = UINT64_MAX;                                      // nothing to release
        return FQB_OK;
    }
    fqb_ctx::BChunk& bc = ctx->bchunks[i % ctx->n_slots];
    const bool have_next = i + 1 < ctx->b_total;
    fqb_ctx::BChunk* bn = have_next ? &ctx->bchunks[(i + 1) % ctx->n_slots] : nullptr;
    lk.unlock();
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(bc.parsed));
    if (bn) CK(cudaEventSynchronize(bn->parsed));
    const unsigned long long status = bc.h_info[0], n_rec = bc.h_info[1], n_lines = bc.h_info[2], lb = bc.h_info[4];
    // the chunk has more line ends than its pinned index holds: fetch them from the device copy (still valid: the
    // device slot is reused only after this batch has been handed out) into a larger buffer
    if (n_lines + 8 > bc.index_cap) {
        uint32_t* bigger = nullptr;
        const size_t cap = (size_t)n_lines + 64;
        CK(cudaHostAlloc(&bigger, cap * 4, cudaHostAllocMapped));
        CK(cudaMemcpy(bigger, bc.d_index, (size_t)n_lines * 4, cudaMemcpyDeviceToHost));
        cudaFreeHost(bc.h_index);
        bc.h_index = bigger;
        bc.index_cap = cap;
        void* d = nullptr;
        CK(cudaHostGetDevicePointer(&d, bigger, 0));
        bc.h_index_dev = static_cast<uint32_t*>(d);
    }
    // line ends in front of the first record that starts in the chunk (they end a record of the chunk before)
    const uint64_t lead = (bc.line_start && (lb & 3) == 0) ? 0 : 4 - (lb & 3);
    const uint64_t need = 4 * n_rec;
    uint64_t have = n_lines > lead ? n_lines - lead : 0;
    if (have < need) {
        // the last record ends in the next chunk: its remaining line ends lead that chunk's index
        if (!bn) return FQB_E_STATE;
        const uint64_t miss = need - have;
        if (miss > 4 || bn->h_info[2] < miss || bc.index_cap < lead + need) return FQB_E_STATE;
        for (uint64_t k = 0; k < miss; ++k) bc.h_index[lead + have + k] = bn->h_index[k];
        have = need;
    }
    // where the first record starts: right behind the `lead`-th line end of the chunk (or at its first byte)
    uint64_t first_off = bc.off;
    if (lead) {
        if (n_lines < lead) {
            first_off = bc.off + bc.n_bytes;                          // no record starts in this chunk
        } else {
            const uint32_t e = bc.h_index[lead - 1];
            first_off = bc.off + (uint32_t)(e - (uint32_t)bc.off) + 1;
        }
    }
    uint64_t end_off = first_off;
    if (n_rec) {
        const uint32_t e = bc.h_index[lead + need - 1];
        end_off = first_off + (uint32_t)(e - (uint32_t)first_off) + 1;
    }
    out->bytes = ctx->h_ring + (size_t)(i % ctx->n_slots) * ctx->slot_bytes + (first_off - bc.off);
    out->n_bytes = end_off - first_off;
    out->n_avail = bc.off + bc.n_bytes + (bn ? std::min<uint64_t>(bn->n_bytes, MAXREC) : 0) - first_off;
    out->stream_offset = first_off;
    out->line_ends = bc.h_index + lead;
    out->n_records = n_rec;
    out->status = (int32_t)status;
    out->err_offset = bc.h_info[3];
    out->token = i;
    lk.lock();
    out->first_record = ctx->b_records;
    ctx->b_records += n_rec;
    ctx->b_next = i + 1;
    ctx->b_held += 1;
    if (status != 0) ctx->b_status = (int)status;
    out->last = (status != 0 || (ctx->batch_closed && i + 1 >= ctx->b_total)) ? 1 : 0;
    return FQB_OK;
}

int fqb_release_batch(fqb_ctx* ctx, uint64_t token)
{
    if (!ctx) return FQB_E_ARG;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        fqb_ctx::BChunk& bc = ctx->bchunks[token % ctx->n_slots];
        if (!ctx->batch_mode || bc.no != token || bc.released) return FQB_E_STATE;
        bc.released = true;
        ctx->b_held -= 1;
    }
    ctx->cv.notify_all();
    return FQB_OK;
}

int fqb_batch_cancel(fqb_ctx* ctx)
{
    if (!ctx) return FQB_E_ARG;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->batch_cancel = true;
    }
    ctx->cv.notify_all();
    return FQB_OK;
}

int fqb_batch_end(fqb_ctx* ctx, fqb_result* res)
{
    if (!ctx) return FQB_E_ARG;
    if (!ctx->batch_mode) return FQB_E_STATE;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->s_copy);
    cudaStreamSynchronize(ctx->s_comp);
    if (res) {
        memset(res, 0, sizeof *res);
        if (cudaMemcpy(ctx->h_carry, ctx->d_carry, sizeof(DevCarry), cudaMemcpyDeviceToHost) == cudaSuccess) {
            res->status = ctx->h_carry->status;
            res->finished = ctx->h_carry->status == 0 && ctx->batch_closed && !ctx->batch_cancel;
            res->n_records = ctx->h_carry->n_records;
            res->n_lines = ctx->h_carry->n_lines;
            res->err_offset = ctx->h_carry->err_offset;
            res->tail_offset = UINT64_MAX;
        }
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->batch_mode = false;
    ctx->streaming = false;
    ctx->acquired = false;
    ctx->fill = 0;
    ctx->have_pending = false;
    return FQB_OK;
}

// ==========================================================================================
// N ranks, one byte shard each: the ONE collective of the path behind the ABI
// ==========================================================================================
// NCCL is bound at run time (dlopen): a process that already carries a libnccl.so.2 (PyTorch's) shares it, a
// plain C / Rust caller gets the system one, and single-GPU users need none at all.  Only entry points whose
// signatures have been stable across NCCL 2.x are used.
namespace {
struct NcclApi {
    struct Id128 {                      // ncclUniqueId (passed by value)
        char b[FQB_COMM_ID_BYTES];
    };
    void* h = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
const char* nccl_load()
{
    if (g_nccl.h) return nullptr;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return "libnccl.so.2 not found";
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(h, "ncclAllReduce"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) return "libnccl: missing symbols";
    g_nccl.h = h;
    return nullptr;
}
int nccl_fail(fqb_ctx* ctx, int rc, const char* what)
{
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error");
    ctx->err = buf;
    return FQB_E_NCCL;
}
}  // namespace

int fqb_comm_unique_id(uint8_t out[FQB_COMM_ID_BYTES])
{
    if (!out) return FQB_E_ARG;
    if (nccl_load()) return FQB_E_NCCL;
    return g_nccl.GetUniqueId(out) == 0 ? FQB_OK : FQB_E_NCCL;
}

int fqb_comm_init(fqb_ctx* ctx, int rank, int world, const uint8_t id[FQB_COMM_ID_BYTES])
{
    if (!ctx || world < 1 || world > FQB_MAX_WORLD || rank < 0 || rank >= world) return FQB_E_ARG;
    if (ctx->nccl_comm) return FQB_E_STATE;
    if (world > 1) {
        if (!id) return FQB_E_ARG;
        if (const char* e = nccl_load()) {
            ctx->err = e;
            return FQB_E_NCCL;
        }
        CK(cudaSetDevice(ctx->device));
        NcclApi::Id128 uid;
        memcpy(uid.b, id, sizeof uid.b);
        void* comm = nullptr;
        const int rc = g_nccl.CommInitRank(&comm, world, uid, rank);
        if (rc != 0) return nccl_fail(ctx, rc, "ncclCommInitRank");
        ctx->nccl_comm = comm;
    }
    ctx->rank = rank;
    ctx->world = world;
    return FQB_OK;
}

int fqb_comm_destroy(fqb_ctx* ctx)
{
    if (!ctx) return FQB_E_ARG;
    if (ctx->nccl_comm && g_nccl.CommDestroy) {
        cudaSetDevice(ctx->device);
        cudaDeviceSynchronize();
        g_nccl.CommDestroy(ctx->nccl_comm);
    }
    ctx->nccl_comm = nullptr;
    ctx->rank = 0;
    ctx->world = 1;
    return FQB_OK;
}

int fqb_comm_rank(fqb_ctx* ctx) { return ctx ? ctx->rank : -1; }
int fqb_comm_world(fqb_ctx* ctx) { return ctx ? ctx->world : -1; }

uint64_t* fqb_device_exchange(fqb_ctx* ctx) { return ctx ? ctx->d_stats : nullptr; }
size_t fqb_exchange_words(fqb_ctx* ctx) { return ctx ? ctx->nwords + 8 * (size_t)ctx->world : 0; }

int fqb_allreduce(fqb_ctx* ctx, void* stream)
{
    if (!ctx) return FQB_E_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(ctx->device));
    const size_t n = ctx->nwords + 8 * (size_t)ctx->world;
    if (ctx->world == 1) {
        CK(cudaMemcpyAsync(ctx->d_reduced, ctx->d_stats, n * 8, cudaMemcpyDeviceToDevice, st));
        return FQB_OK;
    }
    if (!ctx->nccl_comm) return FQB_E_STATE;
    const int rc = g_nccl.AllReduce(ctx->d_stats, ctx->d_reduced, n, 5 /* ncclUint64 */, 0 /* ncclSum */, ctx->nccl_comm, st);
    if (rc != 0) return nccl_fail(ctx, rc, "ncclAllReduce");
    return FQB_OK;
}

int fqb_fetch_reduced(fqb_ctx* ctx, void* stream, uint64_t* host_stats, uint64_t* outcomes)
{
    if (!ctx) return FQB_E_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaSetDevice(ctx->device));
    const size_t n = ctx->nwords + 8 * (size_t)ctx->world;
    // (the outcome slots first: a caller that only wants them -- the common case checks them before it looks at
    // the block -- still pays one copy, one synchronisation)
    CK(cudaMemcpyAsync(ctx->h_reduced, ctx->d_reduced, n * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (host_stats) memcpy(host_stats, ctx->h_reduced, ctx->nwords * 8);
    if (outcomes) memcpy(outcomes, ctx->h_reduced + ctx->nwords, 8 * (size_t)ctx->world * 8);
    return FQB_OK;
}

}  // extern "C"
