// fq_kernels.cu -- sm_100a kernels of the FASTQ delimit + per-position histogram path.
//
// K1+K2 fused: fq_scan_kernel  (one pass over the bytes: newline scan, line-end index, record
//              validation, per-position byte histograms)
// helpers    : fq_diagnose_kernel (classifies the first bad record exactly in the reference's
//              check order), fq_rerun_reset_kernel, fq_finalize_kernel, fq_count_kernel,
//              synthetic generators (K0).
//
// Reference behaviour reproduced (aseyboldt/fastq-rs 0.6.0):
//   src/records.rs:201-247  IdxRecord::from_buffer: 4 line ends per record, '@' at the record
//                           start, '+' right after the sequence line end, raw line lengths equal
//   src/records.rs:65-90    seq()/qual(): one trailing '\r' trimmed
//   src/lib.rs:221-238      each(): every record before the first bad one is delivered
//   src/lib.rs:255-303      too long / truncated mapping
#include "fq_common.cuh"
#include "fq_device.cuh"

namespace fq {

// ------------------------------------------------------------------------------------------
// diagnose: classify the first bad record in the reference's check order
//   '@' (even if the record is incomplete, src/records.rs:139-147) -> header '\n' -> seq '\n'
//   -> '+' as soon as that byte exists (:153-161) -> sep '\n' -> qual '\n' -> raw length equality
//   incomplete: window full -> too long (src/lib.rs:278-283), else truncated (:286-291)
// one warp; only runs real work when something was flagged
// ------------------------------------------------------------------------------------------
__device__ unsigned long long find_nl(const uint8_t* d, unsigned long long from, unsigned long long end, int lane)
{
    for (unsigned long long q = from; q < end; q += 32) {
        unsigned long long a = q + lane;
        unsigned m = __ballot_sync(0xffffffffu, a < end && d[a] == '\n');
        if (m) return q + (__ffs(m) - 1);
    }
    return NONE64;
}

__global__ void fq_diagnose_kernel(const ScanParams p, DevCarry* carry)
{
    const int lane = threadIdx.x;
    DevResult* r = p.res;
    if (carry && carry->status != 0) return;
    const unsigned long long s = r->first_bad;
    if (s == NONE64) return;
    const uint8_t* d = p.data;
    const unsigned long long avail = p.n_avail - s;
    const unsigned long long W = rec_window(p.stream_offset + s);   // what the reference's buffer holds of it
    const unsigned long long end = s + (avail < W ? avail : W);
    int status = 0;
    bool incomplete = false;
    unsigned long long n0, n1 = 0, n2 = 0, n3 = 0;
    if (d[s] != '@') {
        status = 1;
    } else {
        n0 = find_nl(d, s, end, lane);
        if (n0 == NONE64) incomplete = true;
        if (!incomplete) {
            n1 = find_nl(d, n0 + 1, end, lane);
            if (n1 == NONE64) incomplete = true;
        }
        if (!incomplete) {
            if (n1 + 1 >= end)
                incomplete = true;
            else if (d[n1 + 1] != '+')
                status = 2;
        }
        if (!incomplete && !status) {
            n2 = find_nl(d, n1 + 1, end, lane);
            if (n2 == NONE64) incomplete = true;
        }
        if (!incomplete && !status) {
            n3 = find_nl(d, n2 + 1, end, lane);
            if (n3 == NONE64) incomplete = true;
        }
        if (!incomplete && !status) status = (n3 - n2) != (n1 - n0) ? 3 : 51 /* flagged but valid: internal */;
        if (incomplete) status = avail >= W ? 4 : 5;
    }
    if (lane == 0) {
        r->status = status;
        r->err_offset = p.stream_offset + s;
    }
}

// zero the accumulators before a conditional relaunch -- only when there is something to redo
//   mode 0: before the exact path (runs when the speculative kernel did not deliver): everything the
//           speculative launch produced is void, including its error flags
//   mode 1: before the launch restricted to the records in front of the first bad one
__global__ void fq_rerun_reset_kernel(const ScanParams p, int mode)
{
    // mode 0: the speculative pass was abandoned; mode 1: the exact pass found a bad record (a bad record the
    // speculative pass found itself at the end of the shard needs no second pass: nothing behind it was counted)
    if (mode == 0 ? p.res->spec_fail == 0 : (p.res->first_bad == NONE64 || p.res->tail_err)) return;
    const size_t n_stats = stats_words(p.max_len);
    const size_t n_seq = (size_t)p.max_len * 256;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = i0; i < n_stats; i += stride) p.stats[i] = 0;
    for (size_t i = i0; i < n_seq; i += stride) p.seqraw[i] = 0;
    if (mode == 0) {
        for (size_t i = i0; i < p.nranges; i += stride) p.ranges[i].count = 0;
        if (i0 == 0) {
            p.res->first_bad = NONE64;
            p.res->tail_start = NONE64;
            p.res->tail_err = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// exact path, step 1: '\n' count of every CTA range of the exact kernel (grid = ranges x 8 pieces)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fq_range_count_kernel(const ScanParams p, const DevCarry* carry,
                                                             unsigned long long range_bytes)
{
    if (!p.res->spec_fail) return;
    if (carry && carry->status != 0) return;
    const unsigned long long r0 = (unsigned long long)blockIdx.x * range_bytes;
    if (r0 >= p.n_own) return;
    const unsigned long long r1 = min(p.n_own, r0 + range_bytes);
    // 16-byte pieces of the range, dealt to the 8 blocks of the range
    const unsigned long long n16 = (r1 - r0) / 16;
    const uint4* v = reinterpret_cast<const uint4*>(p.data + r0);   // range_bytes is a multiple of 16
    unsigned long long cnt = 0;
    const unsigned long long stride = (unsigned long long)gridDim.y * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.y * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 a = __ldg(v + i);
        const uint32_t s = (nlbits(a.x) >> 7) + (nlbits(a.y) >> 7) + (nlbits(a.z) >> 7) + (nlbits(a.w) >> 7);
        cnt += (s * 0x01010101u) >> 24;   // per-byte sums <= 4: no carries
    }
    if (blockIdx.y == 0 && threadIdx.x < ((r1 - r0) & 15)) cnt += p.data[r0 + n16 * 16 + threadIdx.x] == '\n';
    cnt = warp_sum_u64(cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&p.ranges[blockIdx.x].count, cnt);
}

// exact path, step 2: line number of every range start = prefix of the counts.  One warp.
__global__ void fq_range_prefix_kernel(const ScanParams p, const DevCarry* carry, int nranges)
{
    const int lane = threadIdx.x;
    if (!p.res->spec_fail) return;
    if (carry && carry->status != 0) return;
    const unsigned long long line_base = carry ? carry->line_base : p.line_base;
    unsigned long long run = line_base;
    for (int b0 = 0; b0 < nranges; b0 += 32) {
        const int b = b0 + lane;
        const unsigned long long c = b < nranges ? p.ranges[b].count : 0ull;
        unsigned long long incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (b < nranges) p.ranges[b].base = run + incl - c;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) {
        p.res->n_lines = run - line_base;
        p.res->line_end = run;
    }
}

// fold the raw sequence-byte histogram into the six base classes (validate_dnan's alphabet,
// src/records.rs:29-33), publish the outcome, and (streaming) add the chunk into the totals
__global__ void fq_finalize_kernel(const ScanParams p, DevCarry* carry, unsigned long long* total,
                                   unsigned long long* pub, unsigned long long* slot)
{
    const uint32_t P = p.max_len;
    const bool skip = carry && carry->status != 0;  // stream failed in an earlier chunk
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long* base = p.stats + stats_base_off(P);
    if (!skip) {
        // one warp per position: the 256 raw counters of the row are summed by the lanes (8 each)
        const int lane = threadIdx.x & 31;
        const size_t wstride = stride / 32;
        for (size_t pos = i0 / 32; pos < P; pos += wstride) {
            const unsigned long long* row = p.seqraw + pos * 256;
            unsigned long long all = 0;
            for (int b = lane; b < 256; b += 32) all += row[b];
            all = warp_sum_u64(all);
            if (lane == 0) {
                const unsigned long long a = row['A'], c = row['C'], g = row['G'], t = row['T'], n = row['N'];
                base[pos * 6 + 0] = a;
                base[pos * 6 + 1] = c;
                base[pos * 6 + 2] = g;
                base[pos * 6 + 3] = t;
                base[pos * 6 + 4] = n;
                base[pos * 6 + 5] = all - a - c - g - t - n;
                if (total)   // (streaming: the chunk's base classes into the running totals, by the lane that made them)
                    for (int k = 0; k < 6; ++k) total[stats_base_off(P) + pos * 6 + k] += base[pos * 6 + k];
            }
        }
    }
    // grid-wide ordering is not needed: totals only read words this thread itself finalized or
    // words written by earlier kernels, except base_hist -> handle it in a second sweep below
    if (total && !skip) {
        const size_t nb = stats_base_off(P), nq = stats_qual_off(P), nw = stats_words(P);
        for (size_t i = i0; i < nw; i += stride) {
            if (i >= nb && i < nq) continue;  // base_hist: added above, by the lane that folded the row
            total[i] += p.stats[i];
        }
    }
    if (i0 == 0) {
        DevResult* r = p.res;
        if (!skip) {
            // an inferred shard start the speculative kernel could not stand by: the caller must come
            // back with the exact line_base (FQB_E_PHASE)
            if ((p.flags & F_INFER_START) && r->spec_fail) r->status = 7;
            // a bad record in the middle of the shard, everything in front of it verified: the numbers of this
            // launch include records behind it -- the caller fetches (fqb_fetch parses the bytes in front of it again)
            if (r->spec_retry) r->status = 8;
            r->n_records = p.stats[0];
            r->finished = r->status == 0 ? 1 : 0;
            if (pub) {   // the outcome as 8 device-resident words (fqb_device_result)
                pub[0] = (unsigned long long)r->status;
                pub[1] = (unsigned long long)r->finished;
                pub[2] = r->n_records;
                pub[3] = r->n_lines;
                pub[4] = r->err_offset;
                pub[5] = r->tail_start == NONE64 ? NONE64 : p.stream_offset + r->tail_start;
                pub[6] = (unsigned long long)r->line_phase;
                pub[7] = 0;
                // the same words in this rank's slot behind the statistics block: all other slots are zero, so the
                // all-reduce(sum) of [block | slots] is at the same time the all-gather of the outcomes
                if (slot)
                    for (int k = 0; k < 8; ++k) slot[k] = pub[k];
            }
            if (carry) {
                carry->n_records += p.stats[0];
                carry->n_lines += r->n_lines;
                carry->line_base = r->line_end;
                if (r->tail_start != NONE64 && carry->tail_plus1 == 0)
                    carry->tail_plus1 = p.stream_offset + r->tail_start + 1;
                if (r->status != 0) {
                    carry->status = r->status;
                    carry->err_offset = r->err_offset;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// tail_err: the speculative kernel found the first bad record itself, at the end of a stream; the line ends of
// the bytes from that record on still belong to the shard's line count and index.  One CTA, 16 bytes per
// thread and step (that tail is a truncated record or a few stray lines; the rare long one just takes longer).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) fq_tail_index_kernel(const ScanParams p, const DevCarry* carry)
{
    __shared__ unsigned int s_warp[32];
    __shared__ unsigned long long s_base;
    DevResult* r = p.res;
    if (!r->tail_err || r->spec_fail) return;
    if (carry && carry->status != 0) return;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool want_index = (p.flags & F_INDEX) && p.index != nullptr && p.index_cap != 0;
    if (t == 0) s_base = r->n_lines;
    __syncthreads();
    for (unsigned long long a = r->first_bad; a < p.n_own; a += 16384) {
        const unsigned long long mine = a + 16ull * t;
        unsigned m = 0;                                   // bit i = byte mine + i is a '\n' of the owned bytes
        for (int i = 0; i < 16; ++i)
            if (mine + i < p.n_own && p.data[mine + i] == '\n') m |= 1u << i;
        const unsigned c = __popc(m);
        unsigned incl = c;
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned before = 0, total = 0;
        for (int w = 0; w < 32; ++w) {
            if (w < warp) before += s_warp[w];
            total += s_warp[w];
        }
        unsigned long long rank = s_base + before + (incl - c);
        while (m) {
            const int i = __ffs(m) - 1;
            m &= m - 1;
            if (want_index && rank < p.index_cap) p.index[rank] = (uint32_t)(p.stream_offset + mine + i);
            ++rank;
        }
        __syncthreads();
        if (t == 0) s_base += total;
        __syncthreads();
    }
    if (t == 0) {
        r->line_end += s_base - r->n_lines;
        r->n_lines = s_base;
    }
}

cudaError_t launch_tail_index(const ScanParams& p, DevCarry* carry, cudaStream_t st)
{
    fq_tail_index_kernel<<<1, 1024, 0, st>>>(p, carry);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// count '\n' (line-phase exchange between shards; `wc -l`)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fq_count_kernel(const uint8_t* __restrict__ d, unsigned long long n,
                                                       unsigned long long* out)
{
    const unsigned long long n16 = n / 16;
    const uint4* v = reinterpret_cast<const uint4*>(d);
    unsigned long long cnt = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 a = __ldg(v + i), b = __ldg(v + i + stride), c = __ldg(v + i + 2 * stride), e = __ldg(v + i + 3 * stride);
        uint32_t s0 = nlbits(a.x) >> 7, s1 = nlbits(a.y) >> 7, s2 = nlbits(a.z) >> 7, s3 = nlbits(a.w) >> 7;
        s0 += nlbits(b.x) >> 7; s1 += nlbits(b.y) >> 7; s2 += nlbits(b.z) >> 7; s3 += nlbits(b.w) >> 7;
        s0 += nlbits(c.x) >> 7; s1 += nlbits(c.y) >> 7; s2 += nlbits(c.z) >> 7; s3 += nlbits(c.w) >> 7;
        s0 += nlbits(e.x) >> 7; s1 += nlbits(e.y) >> 7; s2 += nlbits(e.z) >> 7; s3 += nlbits(e.w) >> 7;
        uint32_t s = s0 + s1 + s2 + s3;  // per-byte sums <= 16, no carries
        cnt += (s * 0x01010101u) >> 24;
    }
    for (; i < n16; i += stride) {
        uint4 a = __ldg(v + i);
        cnt += __popc(nlbits(a.x)) + __popc(nlbits(a.y)) + __popc(nlbits(a.z)) + __popc(nlbits(a.w));
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 15)) cnt += d[n16 * 16 + threadIdx.x] == '\n';
    cnt = warp_sum_u64(cnt);
    __shared__ unsigned long long part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int w = 0; w < 8; ++w) s += part[w];
        if (s) atomicAdd(out, s);
    }
}

// ------------------------------------------------------------------------------------------
// K0: synthetic FASTQ, byte-exact twin of the oracle generator
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__device__ __forceinline__ uint8_t synth_byte(unsigned long long seed, unsigned long long rec, uint32_t L, uint32_t k)
{
    if (k < 17) {
        if (k == 0) return '@';
        if (k == 1) return 'F';
        if (k == 2) return 'Q';
        if (k == 16) return '\n';
        unsigned long long pw = 1;
        for (uint32_t i = k; i < 15; ++i) pw *= 10;  // 10^(15-k): k=3 -> 10^12 ... k=15 -> 1
        return (uint8_t)('0' + (rec / pw) % 10);
    }
    k -= 17;
    if (k < L) {
        unsigned long long h = splitmix64(seed ^ ((rec << 10) | k));
        if ((h & 0xFF) < 2) return 'N';
        const char acgt[4] = {'A', 'C', 'G', 'T'};
        return (uint8_t)acgt[(h >> 8) & 3];
    }
    if (k == L) return '\n';
    if (k == L + 1) return '+';
    if (k == L + 2) return '\n';
    k -= L + 3;
    if (k < L) {
        unsigned long long h = splitmix64(seed ^ ((rec << 10) | k));
        uint32_t span = 40u - (20u * k) / L;
        return (uint8_t)(35u + (uint32_t)((h >> 16) % span));
    }
    return '\n';
}

__global__ void fq_synth_fixed_kernel(uint8_t* out, unsigned long long n, unsigned long long byte_off, uint32_t L,
                                      unsigned long long seed)
{
    const unsigned long long rb = 17ull + 2ull * (L + 1) + 2ull;
    const unsigned long long npieces = (n + 15) / 16;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long pc = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; pc < npieces; pc += stride) {
        const unsigned long long o = pc * 16;
        unsigned long long rec = (byte_off + o) / rb;
        uint32_t k = (uint32_t)((byte_off + o) % rb);
        uint32_t w[4] = {0, 0, 0, 0};
        const int cnt = (int)((n - o) < 16 ? (n - o) : 16);
        for (int i = 0; i < cnt; ++i) {
            w[i >> 2] |= (uint32_t)synth_byte(seed, rec, L, k) << (8 * (i & 3));
            if (++k == rb) {
                k = 0;
                ++rec;
            }
        }
        if (cnt == 16 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
            *reinterpret_cast<uint4*>(out + o) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            for (int i = 0; i < cnt; ++i) out[o + i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
        }
    }
}

__device__ __forceinline__ uint32_t synth_var_len(unsigned long long seed, unsigned long long rec)
{
    return 50u + (uint32_t)(splitmix64(seed ^ 0x4C454Eull ^ rec) % 251ull);
}

__global__ void fq_synth_var_sizes_kernel(unsigned long long* sizes, unsigned long long first, unsigned long long count,
                                          unsigned long long seed)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        sizes[i] = 17ull + 2ull * (synth_var_len(seed, first + i) + 1) + 2ull;
}

// one warp per record
__global__ void fq_synth_var_kernel(uint8_t* out, const unsigned long long* rec_off, unsigned long long first,
                                    unsigned long long count, unsigned long long seed)
{
    const int lane = threadIdx.x & 31;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    for (unsigned long long i = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < count;
         i += nwarps) {
        const unsigned long long rec = first + i;
        const uint32_t L = synth_var_len(seed, rec);
        const uint32_t nb = 17u + 2u * (L + 1u) + 2u;
        uint8_t* o = out + rec_off[i];
        for (uint32_t k = lane; k < nb; k += 32) o[k] = synth_byte(seed, rec, L, k);
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
cudaError_t launch_diagnose(const ScanParams& p, DevCarry* carry, cudaStream_t st)
{
    fq_diagnose_kernel<<<1, 32, 0, st>>>(p, carry);
    return cudaGetLastError();
}

cudaError_t launch_rerun_reset(const ScanParams& p, int mode, cudaStream_t st)
{
    fq_rerun_reset_kernel<<<256, 256, 0, st>>>(p, mode);
    return cudaGetLastError();
}

cudaError_t launch_range_count(const ScanParams& p, DevCarry* carry, int nranges, unsigned long long range_bytes,
                               cudaStream_t st)
{
    fq_range_count_kernel<<<dim3(nranges, 8), 256, 0, st>>>(p, carry, range_bytes);
    if (cudaGetLastError() != cudaSuccess) return cudaGetLastError();
    fq_range_prefix_kernel<<<1, 32, 0, st>>>(p, carry, nranges);
    return cudaGetLastError();
}

cudaError_t launch_finalize(const ScanParams& p, DevCarry* carry, unsigned long long* total, unsigned long long* pub,
                            unsigned long long* slot, cudaStream_t st)
{
    fq_finalize_kernel<<<64, 256, 0, st>>>(p, carry, total, pub, slot);
    return cudaGetLastError();
}

cudaError_t launch_count(const uint8_t* d, unsigned long long n, unsigned long long* out, int grid, cudaStream_t st)
{
    fq_count_kernel<<<grid, 256, 0, st>>>(d, n, out);
    return cudaGetLastError();
}

cudaError_t launch_synth_fixed(uint8_t* out, unsigned long long n, unsigned long long byte_off, uint32_t L,
                               unsigned long long seed, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    unsigned long long np = (n + 15) / 16;
    int grid = (int)((np + 255) / 256 < 148ull * 32 ? (np + 255) / 256 : 148ull * 32);
    fq_synth_fixed_kernel<<<grid, 256, 0, st>>>(out, n, byte_off, L, seed);
    return cudaGetLastError();
}

cudaError_t launch_synth_var(uint8_t* out, const unsigned long long* rec_off, unsigned long long first,
                             unsigned long long count, unsigned long long seed, cudaStream_t st)
{
    if (count == 0) return cudaSuccess;
    int grid = (int)((count + 7) / 8 < 148ull * 32 ? (count + 7) / 8 : 148ull * 32);
    fq_synth_var_kernel<<<grid, 256, 0, st>>>(out, rec_off, first, count, seed);
    return cudaGetLastError();
}

cudaError_t launch_synth_var_sizes(unsigned long long* sizes, unsigned long long first, unsigned long long count,
                                   unsigned long long seed, cudaStream_t st)
{
    if (count == 0) return cudaSuccess;
    int grid = (int)((count + 255) / 256 < 148ull * 16 ? (count + 255) / 256 : 148ull * 16);
    fq_synth_var_sizes_kernel<<<grid, 256, 0, st>>>(sizes, first, count, seed);
    return cudaGetLastError();
}

}  // namespace fq
