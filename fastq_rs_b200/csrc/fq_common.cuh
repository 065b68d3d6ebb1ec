// fq_common.cuh -- shared definitions between the sm_100a kernels and the C-ABI host layer.
//
// Path replaced (reference = aseyboldt/fastq-rs 0.6.0):
//   IdxRecord::from_buffer      src/records.rs:201-247   (4 x memchr('\n') + '@'/'+'/length checks)
//   RecordRefIter::advance      src/lib.rs:255-303       (refill / carry-over state machine)
//   Record::seq()/qual()        src/records.rs:82-90     (views the stats closure reads)
// Design notes live in DESIGN.md.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fq {

constexpr int HALO = 1024;                 // bytes staged after the tile (tails of records that start in it)
constexpr int FRONT = 16;                  // bytes staged before the tile (only the last one matters)
constexpr int UNIT = 512;                  // one warp x 16 B
constexpr int HIST_ROWS = 128;             // byte values with a shared-memory counter row (ASCII)
constexpr uint32_t MAXREC = 68u * 1024u;   // src/lib.rs:129 BUFSIZE
constexpr unsigned long long NONE64 = ~0ull;

// Bytes of the stream the reference can hold of a record that starts at stream offset `off`
// (src/lib.rs:276-283 + src/buffer.rs:51-100).  With a reader that fills every read() (Cursor, File, &[u8])
// all buffer refills end on a 16-byte boundary of the STREAM: the first fill is BUFSIZE = 17 * 4096 bytes,
// clean()/replace_buffer() park the leftover so that it ENDS at a multiple of 16 and read_into() adds a
// multiple of 4096 (or all of n_free, itself a multiple of 16).  Buffer offsets and stream offsets therefore
// stay congruent mod 16, an incomplete record is parked at buffer offset `off mod 16`, and "record too long"
// (n_free() == 0 while still incomplete) is raised iff the record needs more than BUFSIZE - off mod 16 bytes.
__host__ __device__ inline unsigned long long rec_window(unsigned long long off) { return MAXREC - (off & 15ull); }

constexpr uint32_t F_HIST = 0x01, F_INDEX = 0x02, F_LINE_START = 0x04, F_EOF = 0x08, F_FRONT16 = 0x10;
constexpr uint32_t F_INFER_START = 0x20;   // the phase of line_base is unknown: range 0 infers its first record too
constexpr uint32_t F_RERUN = 0x100;        // internal: second pass restricted to records before first_bad
constexpr uint32_t F_CARRY = 0x200;        // internal: streaming, line_base comes from the carry block
constexpr uint32_t F_CAN_RETRY = 0x400;    // internal: a record that fails validation in the middle of the shard ends its
                                           // range there (spec_bad) instead of voiding the launch: the host parses the bytes
                                           // in front of it again (fqb_fetch), see fq_stream_verify_kernel


// device-resident outcome of one parse (mirrors fqb_result, plus scratch)
struct DevResult {
    unsigned long long first_bad;   // buffer-relative offset of the first bad record (atomicMin)
    unsigned long long tail_start;  // buffer-relative offset of the first incomplete-not-bad record
    unsigned long long n_lines;     // '\n' in [0, n_own)
    unsigned long long line_end;    // line_base + n_lines
    unsigned long long err_offset;  // stream offset
    unsigned long long n_records;
    int status;
    int finished;
    int spec_fail;                  // the speculative kernel did not deliver: the exact path redoes the shard
    int line_phase;                 // line_base mod 4 implied by the first record start of the shard
    unsigned long long n_win_pred;  // windows of the speculative kernel that were predicted / scanned (FQB_DEBUG)
    unsigned long long n_win_scan;
    int tail_err;                   // the speculative kernel itself found the first bad record: it lies in the last
    unsigned long long spec_bad;    // smallest offset at which a range of the speculative kernel stopped at a record that
                                    // fails validation (F_CAN_RETRY); spec_retry: fq_stream_verify_kernel has confirmed
    int spec_retry;                 // that every record in front of it is delimited as a sequential parse would
    int pad3;
    int shape_var;                  // range of an EOF shard, so nothing behind it was counted; first_bad holds it
                                    // shape_var: the head of the shard holds reads of varying length (fq_init_kernel):
                                    // the speculative kernel's variant for such input does the work
};

// one contiguous range of tiles = the work of one CTA
struct RangeInfo {
    unsigned long long count;       // '\n' in the owned bytes of the range (phase independent)
    unsigned long long base;        // exact stream line number at the range start (fq_verify_kernel)
    uint32_t spec_phase;            // line number mod 4 the CTA inferred from the first records of its range
    uint32_t flags;                 // 1 = inferred, 2 = inference failed (ambiguous or nothing to test)
};

// one contiguous byte range = the work of one warp of the speculative kernel (fq_stream.cu)
struct StreamRange {
    unsigned long long first;       // buffer offset of the first record that starts in the range (inferred)
    unsigned long long end;         // cursor after the last such record = start of the first one beyond the range
    unsigned long long n_lines;     // '\n' of those records inside the owned bytes (+ the ones in front of range 0's first record)
    unsigned long long rank0;       // prefix of n_lines (fq_stream_verify_kernel): where the staged line ends go
    uint32_t flags;                 // 1 = delivered, 2 = gave up, 3 = stopped at a bad record (F_CAN_RETRY)
    uint32_t n_desc;                // window descriptors the range wrote (DESC_RAW_ONLY: none, its line ends are all staged)
};
constexpr uint32_t DESC_RAW_ONLY = 0xFFFFFFFFu;

// streaming carry block (device resident, lives across chunk launches)
struct DevCarry {
    unsigned long long line_base;   // lines before the next chunk
    unsigned long long n_records;   // records delivered so far
    unsigned long long err_offset;
    unsigned long long n_lines;
    int status;                     // sticky first error
    int pad;
    unsigned long long tail_plus1;  // 1 + stream offset of the first incomplete-not-bad record (0 = none)
};

struct ScanParams {
    const uint8_t* data;
    unsigned long long n_own, n_avail;
    unsigned long long stream_offset;
    unsigned long long line_base;
    const DevCarry* carry;          // F_CARRY
    uint32_t flags;
    uint32_t max_len;               // P
    uint32_t ntiles;
    uint32_t tiles_per_cta;         // CTA b owns the tiles [b * tiles_per_cta, (b + 1) * tiles_per_cta)
    RangeInfo* ranges;              // [nranges] CTA ranges of the exact kernel
    uint32_t nranges;
    uint32_t n_sranges;             // live warp ranges: r covers [r * srange_bytes, (r + 1) * srange_bytes), the last one up to n_own
    StreamRange* sranges;           // [32 * grid] warp ranges of the speculative kernel
    unsigned long long srange_bytes;
    uint32_t* index_stage;          // speculative kernel: range r stages its line ends at index_stage + r * stage_share
    unsigned long long stage_share;
    // ... or, window by window, DESCRIBES them: a window of predicted records is an arithmetic sequence -- 16 bytes
    // instead of 16 bytes per record.  desc + 4 * (r * desc_cap + k) = the k-th window of range r:
    //   [0] line ends of the window | kind << 24   kind 0: they are staged at index_stage + r * stage_share + [1]
    //   [1] kind 1: low 32 bits of the stream offset of the window's first record
    //   [2] kind 1: offsets of the 1st | 2nd << 16 line end within a record    [3] 3rd | record length << 16
    // fq_stream_compact_kernel turns descriptors (and staged runs) into the dense index.
    uint32_t* desc;
    unsigned long long desc_cap;
    uint32_t* index;
    unsigned long long index_cap;
    DevResult* res;
    unsigned long long* stats;      // stats block
    unsigned long long* seqraw;     // [P][256] raw byte histogram of the sequence lines
    unsigned long long* trace;      // debug timeline (FQB_TRACE), normally null: [cta][TRACE_K][16] clock64 stamps
};
constexpr int TRACE_K = 2048;

__host__ __device__ inline size_t stats_len_off(uint32_t) { return 8; }
__host__ __device__ inline size_t stats_base_off(uint32_t P) { return 8 + (size_t)P + 2; }
__host__ __device__ inline size_t stats_qual_off(uint32_t P) { return 8 + (size_t)P + 2 + 6 * (size_t)P; }
__host__ __device__ inline size_t stats_words(uint32_t P) { return 8 + (size_t)P + 2 + 6 * (size_t)P + 256 * (size_t)P; }

// record filter (fq_filter.cu)
constexpr unsigned long long FILTER_MAX_WRAPS = 64;   // 32-bit offset wraps per call = 256 GiB of input
struct FilterParams {
    const uint8_t* data;               // shard bytes
    const uint32_t* index;             // 4 line ends per record (low 32 bits of stream offsets)
    unsigned long long n_records;
    unsigned long long stream_offset;  // stream offset of data[0]
    unsigned long long first_offset;   // stream offset of the first byte of record 0
    uint32_t mode;                     // 0 keep all, 1 validate_dna, 2 validate_dnan
    unsigned long long* blk;           // [blocks] look-back descriptors: status | kept bytes (total, then prefix)
    unsigned int* ticket;              // order in which the blocks enter the look-back chain
    unsigned long long* wraps;         // [0] count, [1..] records at which the 32-bit offsets wrap; result follows
    uint8_t* out;
    unsigned long long out_cap;
    unsigned long long* result;        // [0] records kept [1] bytes kept [2] wraps seen
};
cudaError_t launch_filter(const FilterParams& p, int num_sms, cudaStream_t st);
int filter_launches(unsigned long long n_records);

// launchers (fq_kernels.cu)
size_t scan_smem_bytes(int nchunk);
uint32_t scan_tile_bytes(int nchunk);
cudaError_t scan_configure();
int scan_blocks_per_sm(int nchunk);
cudaError_t launch_scan(const ScanParams& p, int nchunk, int grid, cudaStream_t st);
cudaError_t launch_diagnose(const ScanParams& p, DevCarry* carry, cudaStream_t st);
cudaError_t launch_rerun_reset(const ScanParams& p, int mode, cudaStream_t st);
cudaError_t launch_range_count(const ScanParams& p, DevCarry* carry, int nranges, unsigned long long range_bytes,
                               cudaStream_t st);
int stream_warps(int nchunk, bool hist);
cudaError_t stream_configure();
cudaError_t launch_stream(const ScanParams& p, int nchunk, int grid, cudaStream_t st);
cudaError_t launch_stream_verify(const ScanParams& p, DevCarry* carry, cudaStream_t st);
cudaError_t launch_stream_compact(const ScanParams& p, DevCarry* carry, int grid, cudaStream_t st);
cudaError_t launch_tail_index(const ScanParams& p, DevCarry* carry, cudaStream_t st);
cudaError_t launch_finalize(const ScanParams& p, DevCarry* carry, unsigned long long* total, unsigned long long* pub,
                            unsigned long long* slot, cudaStream_t st);
cudaError_t launch_count(const uint8_t* d, unsigned long long n, unsigned long long* out, int grid, cudaStream_t st);
cudaError_t launch_synth_fixed(uint8_t* out, unsigned long long n, unsigned long long byte_off, uint32_t L,
                               unsigned long long seed, cudaStream_t st);
cudaError_t launch_synth_var(uint8_t* out, const unsigned long long* rec_off, unsigned long long first,
                             unsigned long long count, unsigned long long seed, cudaStream_t st);
cudaError_t launch_synth_var_sizes(unsigned long long* sizes, unsigned long long first, unsigned long long count,
                                   unsigned long long seed, cudaStream_t st);

}  // namespace fq
