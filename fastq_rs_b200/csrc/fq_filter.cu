// fq_filter.cu -- the step after the path: keep the records whose seq() passes
// Record::validate_dna / validate_dnan and write them out verbatim, densely, in stream order.
//
// Reference behaviour reproduced (aseyboldt/fastq-rs 0.6.0):
//   src/records.rs:19-23   validate_dna : every byte of seq() is one of A C T G
//   src/records.rs:29-33   validate_dnan: ... or N          (uppercase only; an empty seq() passes)
//   src/records.rs:82-85   seq() = bytes between the header '\n' and the sequence '\n', one trailing '\r' dropped
//   src/records.rs:93-96   RefRecord::write = the record's raw bytes, '@' .. final '\n', unchanged
//
// Input: the shard's bytes and the line-end index a finished fqb_parse_device wrote (4 x u32 per record =
// low 32 bits of the stream offsets of its four '\n').  HBM-bound byte work: the input is read once, the
// survivors are written once.
//   fq_filter_wraps_kernel  the (very few) records at which the 32-bit offsets wrap (one per 4 GiB)
//   fq_filter_kernel        one CTA per 256 records, taken in ticket order:
//                           (1) one thread per record: predicate over the sequence line (aligned 16-byte loads,
//                               SWAR membership test) -> bytes this record contributes (0 if dropped);
//                           (2) block prefix of those bytes; the block total is published and the block's
//                               offset in the output comes from a decoupled look-back over the totals /
//                               prefixes of the blocks before it (single pass, no second read of the input);
//                           (3) consecutive survivors are adjacent in the input and in the output: each warp
//                               copies whole runs with dst-aligned 16-byte stores (source words funnel-shifted
//                               into place, all loads of a run issued before its first store) -- the bytes
//                               were just touched by (1), so most of them come from L1/L2.
#include "fq_common.cuh"
#include "fq_device.cuh"

namespace fq {

static constexpr int FBLOCK = 256;  // records per CTA in mark / copy

// buffer position of record k's first byte: E(k-1) + 1 - stream_offset, E(-1) = first_offset - 1
__device__ __forceinline__ unsigned long long record_pos(const FilterParams& p, unsigned long long k, uint32_t prev_lo,
                                                         const unsigned long long* __restrict__ wr, uint32_t nwr)
{
    long long hi = (long long)(p.first_offset - 1) >> 32;
    for (uint32_t j = 0; j < nwr; ++j) hi += wr[1 + j] < k ? 1 : 0;
    const long long e_prev = hi * 4294967296ll + (long long)prev_lo;
    return (unsigned long long)(e_prev + 1) - p.stream_offset;
}

__global__ void __launch_bounds__(256) fq_filter_wraps_kernel(const FilterParams p)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const uint32_t lo_m1 = (uint32_t)(p.first_offset - 1);
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < p.n_records; k += stride) {
        const uint32_t cur = __ldg(p.index + 4 * k + 3);
        const uint32_t prev = k ? __ldg(p.index + 4 * k - 1) : lo_m1;
        if (cur < prev) {
            const unsigned long long slot = atomicAdd(p.wraps, 1ull);
            if (slot < FILTER_MAX_WRAPS) p.wraps[1 + slot] = k;
        }
    }
}

// every byte of [a0, a1) is in the alphabet whose (byte & 31) bits are set in M and whose top 3 bits are 010.
// Aligned 16-byte loads; the bytes of the first / last vector that lie outside the line count as 'A'.
__device__ __forceinline__ uint32_t word_bad(uint32_t x, uint32_t M)
{
    const uint32_t ok = (M >> (x & 31)) & (M >> ((x >> 8) & 31)) & (M >> ((x >> 16) & 31)) & (M >> ((x >> 24) & 31));
    return ((x & 0xE0E0E0E0u) ^ 0x40404040u) | (~ok & 1u);
}

__device__ __forceinline__ uint32_t clip_word(uint32_t x, long long lo, long long hi)   // valid bytes [lo, hi) of the word
{
    lo = lo < 0 ? 0 : lo;
    hi = hi > 4 ? 4 : hi;
    if (lo >= hi) return 0x41414141u;
    const uint32_t m = (0xFFFFFFFFu << (8 * (uint32_t)lo)) & (0xFFFFFFFFu >> (8 * (4 - (uint32_t)hi)));
    return (x & m) | (0x41414141u & ~m);
}

__device__ __forceinline__ bool seq_ok(const uint8_t* __restrict__ d, unsigned long long a0, unsigned long long a1, uint32_t M)
{
    uint32_t bad = 0;
    unsigned long long a = a0 & ~15ull;
    const uint4* v = reinterpret_cast<const uint4*>(d + a);
    for (; a < a1; a += 16, ++v) {
        uint4 x = __ldg(v);
        if (a < a0 || a + 16 > a1) {
            const long long lo = (long long)a0 - (long long)a, hi = (long long)a1 - (long long)a;
            x.x = clip_word(x.x, lo, hi);
            x.y = clip_word(x.y, lo - 4, hi - 4);
            x.z = clip_word(x.z, lo - 8, hi - 8);
            x.w = clip_word(x.w, lo - 12, hi - 12);
        }
        bad |= word_bad(x.x, M) | word_bad(x.y, M) | word_bad(x.z, M) | word_bad(x.w, M);
    }
    return bad == 0;
}

// copy n bytes with all 32 lanes: dst-aligned 16-byte stores, two chunks per lane in flight; source words are
// funnel-shifted into place unless source and destination are congruent mod 16 (then plain 16-byte loads)
__device__ __forceinline__ uint4 load_shifted(const uint32_t* __restrict__ sa, uint32_t sh)
{
    const uint32_t w0 = __ldcs(sa), w1 = __ldcs(sa + 1), w2 = __ldcs(sa + 2), w3 = __ldcs(sa + 3);
    // the fifth word holds bytes of this run whenever sh != 0, so the aligned load stays inside its page
    const uint32_t w4 = sh ? __ldcs(sa + 4) : 0u;
    return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                      __funnelshift_r(w3, w4, sh));
}

__device__ __forceinline__ void copy_run(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, unsigned long long n, int lane)
{
    unsigned long long h = (0 - reinterpret_cast<uintptr_t>(dst)) & 15;   // bytes until dst is 16-aligned
    if (h > n) h = n;
    const unsigned long long body = (n - h) >> 4, t = (n - h) & 15, o = h + 16 * body;
    const uint8_t* s = src + h;
    uint4* d = reinterpret_cast<uint4*>(dst + h);
    const bool aligned = (reinterpret_cast<uintptr_t>(s) & 15) == 0;
    const uint4* sv = reinterpret_cast<const uint4*>(s);
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(s) & 3) * 8;
    const uint32_t* sa = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(s) & ~(uintptr_t)3);
    // first 32 chunks + head + tail: every load is issued before the first store (one memory latency per run)
    uint8_t hb = 0, tb = 0;
    uint4 x = make_uint4(0, 0, 0, 0);
    if ((unsigned long long)lane < h) hb = __ldcs(src + lane);
    if ((unsigned long long)lane < t) tb = __ldcs(src + o + lane);
    if ((unsigned long long)lane < body) x = aligned ? __ldcs(sv + lane) : load_shifted(sa + 4 * lane, sh);
    if ((unsigned long long)lane < h) dst[lane] = hb;
    if ((unsigned long long)lane < t) dst[o + lane] = tb;
    if ((unsigned long long)lane < body) __stcs(d + lane, x);
    if (body <= 32) return;
    unsigned long long c = lane + 32;
    if (aligned) {
        for (; c + 96 < body; c += 128) {
            const uint4 x0 = __ldcs(sv + c), x1 = __ldcs(sv + c + 32), x2 = __ldcs(sv + c + 64), x3 = __ldcs(sv + c + 96);
            __stcs(d + c, x0);
            __stcs(d + c + 32, x1);
            __stcs(d + c + 64, x2);
            __stcs(d + c + 96, x3);
        }
        for (; c < body; c += 32) __stcs(d + c, __ldcs(sv + c));
    } else {
        for (; c + 32 < body; c += 64) {
            const uint4 x0 = load_shifted(sa + 4 * c, sh), x1 = load_shifted(sa + 4 * (c + 32), sh);
            __stcs(d + c, x0);
            __stcs(d + c + 32, x1);
        }
        for (; c < body; c += 32) __stcs(d + c, load_shifted(sa + 4 * c, sh));
    }
}

// descriptor of a block in the look-back chain: status in the two top bits, bytes below
static constexpr unsigned long long F_AGG = 1ull << 62, F_PRE = 2ull << 62, F_VAL = (1ull << 62) - 1ull;

__global__ void __launch_bounds__(FBLOCK) fq_filter_kernel(const FilterParams p)
{
    __shared__ uint32_t s_warp[FBLOCK / 32];
    __shared__ uint32_t s_cnt[FBLOCK / 32];
    __shared__ unsigned long long s_excl;
    __shared__ unsigned int s_block;
    if (threadIdx.x == 0) s_block = atomicAdd(p.ticket, 1u);   // blocks enter the chain in the order they start
    __syncthreads();
    const unsigned long long blk = s_block;
    const unsigned long long k = blk * FBLOCK + threadIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nwr = (uint32_t)min(p.wraps[0], (unsigned long long)FILTER_MAX_WRAPS);

    // (1) predicate
    uint32_t len = 0;
    unsigned long long pos = 0;
    if (k < p.n_records) {
        uint4 e;
        if ((reinterpret_cast<uintptr_t>(p.index) & 15) == 0) {
            e = __ldg(reinterpret_cast<const uint4*>(p.index) + k);
        } else {   // an index that starts at the shard's phase (not a multiple of 4 entries in)
            e.x = __ldg(p.index + 4 * k);
            e.y = __ldg(p.index + 4 * k + 1);
            e.z = __ldg(p.index + 4 * k + 2);
            e.w = __ldg(p.index + 4 * k + 3);
        }
        const uint32_t prev_lo = k ? __ldg(p.index + 4 * k - 1) : (uint32_t)(p.first_offset - 1);
        pos = record_pos(p, k, prev_lo, p.wraps, nwr);
        const uint32_t s_lo = prev_lo + 1;
        const uint32_t h = e.x - s_lo, sq = e.y - s_lo, q = e.w - s_lo;   // record-relative line ends
        unsigned long long a0 = pos + h + 1, a1 = pos + sq;
        if (a1 > a0 && p.data[a1 - 1] == '\r') --a1;                      // trim_winline, src/records.rs:65-73
        bool ok = true;
        if (p.mode == 1)
            ok = seq_ok(p.data, a0, a1, (1u << 1) | (1u << 3) | (1u << 7) | (1u << 20));                // A C G T
        else if (p.mode == 2)
            ok = seq_ok(p.data, a0, a1, (1u << 1) | (1u << 3) | (1u << 7) | (1u << 20) | (1u << 14));   // + N
        len = ok ? q + 1 : 0;
    }

    // (2) block prefix, then the block's place in the output
    uint32_t incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    const uint32_t cnt = (uint32_t)__popc(__ballot_sync(0xffffffffu, len != 0));
    if (lane == 31) s_warp[warp] = incl;
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    uint32_t before = 0;
    for (int i = 0; i < warp; ++i) before += s_warp[i];
    if (warp == 0) {
        unsigned long long total = 0, kept = 0;
        for (int i = 0; i < FBLOCK / 32; ++i) {
            total += s_warp[i];
            kept += s_cnt[i];
        }
        volatile unsigned long long* desc = p.blk;
        if (lane == 0) {
            desc[blk] = (blk == 0 ? F_PRE : F_AGG) | total;
            if (kept) atomicAdd(p.result + 0, kept);
            if (total) atomicAdd(p.result + 1, total);
        }
        // decoupled look-back: lane l looks at block blk - 1 - l of the current group of 32
        unsigned long long excl = 0;
        long long j = (long long)blk - 1;
        while (j >= 0) {
            const long long mine = j - lane;
            unsigned long long d = F_PRE;                       // blocks in front of block 0: prefix 0
            if (mine >= 0) {
                do {
                    d = desc[mine];
                } while ((d & ~F_VAL) == 0);
            }
            const unsigned pre = __ballot_sync(0xffffffffu, (d & F_PRE) != 0);
            const int stop = __ffs(pre) - 1;                    // nearest block that already knows its prefix
            unsigned long long v = (stop < 0 || lane <= stop) ? (d & F_VAL) : 0ull;
            excl += warp_sum_u64(v);
            if (stop >= 0) break;
            j -= 32;
        }
        if (lane == 0) {
            if (blk != 0) desc[blk] = F_PRE | (excl + total);
            s_excl = excl;
        }
    }
    __syncthreads();

    // (3) copy the survivors
    const unsigned long long dst = s_excl + before + (incl - len);
    if (dst + len > p.out_cap) len = 0;   // what does not fit is not written (out_bytes tells)
    const uint32_t kept = __ballot_sync(0xffffffffu, len != 0);
    uint32_t starts = kept & ~(kept << 1);
    while (starts) {
        const int b = __ffs(starts) - 1;
        starts &= starts - 1;
        const uint32_t gap = ~kept & (0xFFFFFFFFu << b);          // first dropped record behind the run
        const int last = (gap ? __ffs(gap) - 1 : 32) - 1;
        const unsigned long long src_b = __shfl_sync(0xffffffffu, pos, b), dst_b = __shfl_sync(0xffffffffu, dst, b);
        const unsigned long long end = __shfl_sync(0xffffffffu, dst + len, last);
        copy_run(p.data + src_b, p.out + dst_b, end - dst_b, lane);
    }
}

__global__ void fq_filter_finish_kernel(const FilterParams p)
{
    p.result[2] = p.wraps[0];
}

cudaError_t launch_filter(const FilterParams& p, int num_sms, cudaStream_t st)
{
    const unsigned long long nblk = (p.n_records + FBLOCK - 1) / FBLOCK;
    cudaError_t e = cudaMemsetAsync(p.wraps, 0, (1 + FILTER_MAX_WRAPS + 4) * sizeof(unsigned long long), st);   // + result
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(p.ticket, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    if (nblk) {
        e = cudaMemsetAsync(p.blk, 0, nblk * sizeof(unsigned long long), st);
        if (e != cudaSuccess) return e;
        const unsigned long long want = (p.n_records + 255) / 256;
        const int grid = (int)(want < (unsigned long long)num_sms * 8 ? want : (unsigned long long)num_sms * 8);
        fq_filter_wraps_kernel<<<grid, 256, 0, st>>>(p);
        fq_filter_kernel<<<(unsigned)nblk, FBLOCK, 0, st>>>(p);
    }
    fq_filter_finish_kernel<<<1, 1, 0, st>>>(p);
    return cudaGetLastError();
}

int filter_launches(unsigned long long n_records) { return n_records ? 3 : 1; }

}  // namespace fq
