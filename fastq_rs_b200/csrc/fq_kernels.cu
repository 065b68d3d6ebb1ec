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

namespace fq {

// ------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// 0x80 in every byte of w that equals '\n' (exact: no false positives, no cross-byte carries)
__device__ __forceinline__ uint32_t nlbits(uint32_t w)
{
    uint32_t y = (w ^ 0x0A0A0A0Au) & 0x7F7F7F7Fu;
    uint32_t t = y + 0x7F7F7F7Fu;
    return ~(t | w) & 0x80808080u;
}
// 16-bit mask, bit i = byte i of the 16-byte piece is '\n'
__device__ __forceinline__ uint32_t nlmask16(const uint4& v)
{
    uint32_t m0 = nlbits(v.x), m1 = nlbits(v.y), m2 = nlbits(v.z), m3 = nlbits(v.w);
    uint32_t c01 = m1 | (m0 >> 4);
    uint32_t c23 = m3 | (m2 >> 4);
    // multiply-gather: bits {3,11,19,27} -> 24..27, bits {7,15,23,31} -> 28..31
    uint32_t b01 = (c01 * 0x00204081u) >> 24;
    uint32_t b23 = (c23 * 0x00204081u) >> 24;
    return b01 | (b23 << 8);
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// ------------------------------------------------------------------------------------------
// per-CTA state in shared memory
// ------------------------------------------------------------------------------------------
struct TileShared {
    unsigned long long mbar;
    unsigned long long base;       // exclusive line count (stream-global) at the tile start
    uint32_t tile;                 // ticket
    uint32_t front;                // 1 if the byte before the tile is (or acts as) '\n'
    uint32_t own_count;            // '\n' in the owned range
    uint32_t total_count;          // '\n' in owned range + halo
    uint32_t nonascii;             // some byte >= 0x80 staged
    uint32_t unit_all[NUNITS + 2];
    uint32_t unit_own[NUNITS + 2];
    uint32_t unit_base[NUNITS + 2];
};

struct Acc {  // per-warp accumulators kept in registers of lane 0
    unsigned long long n_records, n_bases, clip_seq, clip_qual;
};

template <int NCHUNK>
struct Smem {
    static constexpr int HIST_BYTES = NCHUNK * CHUNK_WORDS * 4;
    static constexpr int TILE_BYTES = (SM_TILE + 127) / 128 * 128;
    static constexpr int LIST_BYTES = LIST_CAP * 2;
    static constexpr int LENH_WORDS = (32 * NCHUNK + 2 + 63) / 64 * 64;
    static constexpr int TOTAL = HIST_BYTES + TILE_BYTES + LIST_BYTES + LENH_WORDS * 4;
};

// add one byte observation at read position c (< 32*NCHUNK): conflict-free because lane == c % 32
// owns bank `lane` of every row; `inc` = 1 (sequence line) or 0x10000 (quality line)
template <bool ASCII>
__device__ __forceinline__ void hist_add(uint32_t* hp /* hist + chunk*CHUNK_WORDS + lane */, uint32_t b, uint32_t inc,
                                         unsigned long long* gfallback /* &g[c*256] */)
{
    if (ASCII || b < (uint32_t)HIST_ROWS)
        atomicAdd(hp + (b << 5), inc);
    else
        atomicAdd(gfallback + b, 1ull);
}

// flush the u16-pair shared counters into the u64 global histograms
template <int NCHUNK>
__device__ void flush_hist(uint32_t* hist, uint32_t* lenh, const ScanParams& p)
{
    const uint32_t P = p.max_len;
    unsigned long long* qual = p.stats + stats_qual_off(P);
    unsigned long long* lenh_g = p.stats + stats_len_off(P);
    __syncthreads();
    for (int i = threadIdx.x; i < NCHUNK * CHUNK_WORDS; i += THREADS) {
        uint32_t v = hist[i];
        if (v) {
            hist[i] = 0;
            uint32_t pos = (uint32_t)(i / CHUNK_WORDS) * 32u + (uint32_t)(i & 31);
            uint32_t b = (uint32_t)(i >> 5) & (HIST_ROWS - 1);
            uint32_t lo = v & 0xFFFFu, hi = v >> 16;
            if (pos < P) {
                if (lo) atomicAdd(p.seqraw + (size_t)pos * 256 + b, (unsigned long long)lo);
                if (hi) atomicAdd(qual + (size_t)pos * 256 + b, (unsigned long long)hi);
            }
        }
    }
    for (int i = threadIdx.x; i < 32 * NCHUNK + 2; i += THREADS) {
        uint32_t v = lenh[i];
        if (v) {
            lenh[i] = 0;
            atomicAdd(lenh_g + i, (unsigned long long)v);
        }
    }
    __syncthreads();
}

// Book-keeping for one valid record whose trimmed lengths are Ls / Lq (lane 0 only).
template <int NCHUNK>
__device__ __forceinline__ void account_record(Acc& acc, uint32_t* lenh, const ScanParams& p, uint32_t Ls, uint32_t Lq)
{
    const uint32_t P = p.max_len;
    acc.n_bases += Ls;
    if (Ls > P) acc.clip_seq += Ls - P;
    if (Lq > P) acc.clip_qual += Lq - P;
    uint32_t lb = Ls <= P ? Ls : P + 1;
    if (lb < 32u * NCHUNK + 2u)
        atomicAdd(lenh + lb, 1u);
    else
        atomicAdd(p.stats + stats_len_off(P) + lb, 1ull);
}

// Slow path: a record that is not completely staged in shared memory (longer than the halo,
// or in a tile with more newlines than LIST_CAP).  One warp walks it in global memory.
// Returns the buffer-relative offset of its final '\n', or NONE64 when the record was flagged
// (bad / incomplete) or skipped.
template <int NCHUNK>
__device__ unsigned long long record_global(const ScanParams& p, unsigned long long s, unsigned long long limit,
                                            uint32_t* hist, uint32_t* lenh, Acc& acc, int lane)
{
    if (s >= limit) return NONE64;
    const uint8_t* __restrict__ d = p.data;
    const unsigned long long navail = p.n_avail;
    unsigned long long win_end = s + MAXREC;
    const bool window_full = win_end <= navail;
    if (win_end > navail) win_end = navail;
    unsigned long long nl[4];
    int found = 0;
    for (unsigned long long q = s; q < win_end && found < 4; q += 32) {
        unsigned long long a = q + lane;
        bool isnl = a < win_end && d[a] == '\n';
        unsigned m = __ballot_sync(0xffffffffu, isnl);
        while (m && found < 4) {
            int b = __ffs(m) - 1;
            m &= m - 1;
            nl[found++] = q + b;
        }
    }
    bool bad = false, tail = false;
    if (found < 4) {
        // incomplete within the window: too long if the window was full; otherwise the data
        // ended -- an error at EOF, a tail to carry over when more data will follow.
        // (src/lib.rs:276-293)
        if (window_full || (p.flags & F_EOF))
            bad = true;
        else
            tail = true;
    } else {
        bad = d[s] != '@' || d[nl[1] + 1] != '+' || (nl[3] - nl[2]) != (nl[1] - nl[0]);
    }
    if (bad || tail) {
        if (lane == 0) {
            if (bad)
                atomicMin(&p.res->first_bad, s);
            else
                atomicMin(&p.res->tail_start, s);
        }
        return NONE64;
    }
    if (lane == 0) acc.n_records++;
    if (p.flags & F_HIST) {
        const uint32_t P = p.max_len;
        const uint32_t Pm = P < 32u * NCHUNK ? P : 32u * NCHUNK;
        uint32_t Lr = (uint32_t)(nl[1] - nl[0] - 1);
        uint32_t Ls = Lr - ((Lr > 0 && d[nl[1] - 1] == '\r') ? 1u : 0u);
        uint32_t Lq = Lr - ((Lr > 0 && d[nl[3] - 1] == '\r') ? 1u : 0u);
        const uint8_t* sq = d + nl[0] + 1;
        const uint8_t* ql = d + nl[2] + 1;
        unsigned long long* qualg = p.stats + stats_qual_off(P);
        uint32_t ns = Ls < P ? Ls : P, nq = Lq < P ? Lq : P;
        for (uint32_t c = lane; c < ns; c += 32) {
            uint32_t b = sq[c];
            if (c < Pm)
                hist_add<false>(hist + (c >> 5) * CHUNK_WORDS + lane, b, 1u, p.seqraw + (size_t)c * 256);
            else
                atomicAdd(p.seqraw + (size_t)c * 256 + b, 1ull);
        }
        for (uint32_t c = lane; c < nq; c += 32) {
            uint32_t b = ql[c];
            if (c < Pm)
                hist_add<false>(hist + (c >> 5) * CHUNK_WORDS + lane, b, 0x10000u, qualg + (size_t)c * 256);
            else
                atomicAdd(qualg + (size_t)c * 256 + b, 1ull);
        }
        if (lane == 0) account_record<NCHUNK>(acc, lenh, p, Ls, Lq);
    }
    return nl[3];
}

// Fast path: all four line ends of the record are in the shared-memory list.
template <int NCHUNK, bool ASCII>
__device__ __forceinline__ void record_smem(const ScanParams& p, const uint8_t* tile, const uint16_t* list, uint32_t j,
                                            unsigned long long abs_s, uint32_t* hist, uint32_t* lenh, Acc& acc,
                                            int lane)
{
    const uint32_t s = (uint32_t)list[j] + 1u;
    const uint32_t h = list[j + 1], q = list[j + 2], pp = list[j + 3], e = list[j + 4];
    // src/records.rs:137-149 ('@'), :151-163 ('+'), :233-238 (raw length equality)
    const bool ok = tile[s] == '@' && tile[q + 1] == '+' && (e - pp) == (q - h);
    if (!ok) {
        if (lane == 0) atomicMin(&p.res->first_bad, abs_s);
        return;
    }
    if (lane == 0) acc.n_records++;
    if (!(p.flags & F_HIST)) return;
    const uint32_t P = p.max_len;
    const uint32_t Pm = P < 32u * NCHUNK ? P : 32u * NCHUNK;
    const uint32_t Lr = q - h - 1u;
    // seq()/qual() drop one trailing '\r' (src/records.rs:65-73,82-90)
    const uint32_t Ls = Lr - ((Lr > 0 && tile[q - 1] == '\r') ? 1u : 0u);
    const uint32_t Lq = Lr - ((Lr > 0 && tile[e - 1] == '\r') ? 1u : 0u);
    const uint8_t* sq = tile + h + 1;
    const uint8_t* ql = tile + pp + 1;
    unsigned long long* qualg = p.stats + stats_qual_off(P);
    const uint32_t ns = Ls < Pm ? Ls : Pm, nq = Lq < Pm ? Lq : Pm;
    const uint32_t nboth = ns < nq ? ns : nq;
    uint32_t* hp = hist + lane;
    uint32_t c = lane;
    // full 32-position chunks of both lines
    for (; c + (31 - lane) < nboth; c += 32, hp += CHUNK_WORDS) {
        hist_add<ASCII>(hp, sq[c], 1u, p.seqraw + (size_t)c * 256);
        hist_add<ASCII>(hp, ql[c], 0x10000u, qualg + (size_t)c * 256);
    }
    // ragged remainder (warp-uniform trip count)
    const uint32_t nmax = ns > nq ? ns : nq;
    for (; c - lane < nmax; c += 32, hp += CHUNK_WORDS) {
        if (c < ns) hist_add<ASCII>(hp, sq[c], 1u, p.seqraw + (size_t)c * 256);
        if (c < nq) hist_add<ASCII>(hp, ql[c], 0x10000u, qualg + (size_t)c * 256);
    }
    // positions beyond the shared-memory histogram but below P: straight to global
    if (P > Pm) {
        const uint32_t gs = Ls < P ? Ls : P, gq = Lq < P ? Lq : P;
        for (uint32_t g = Pm + lane; g < gs; g += 32) atomicAdd(p.seqraw + (size_t)g * 256 + sq[g], 1ull);
        for (uint32_t g = Pm + lane; g < gq; g += 32) atomicAdd(qualg + (size_t)g * 256 + ql[g], 1ull);
    }
    if (lane == 0) account_record<NCHUNK>(acc, lenh, p, Ls, Lq);
}

template <int NCHUNK, bool ASCII>
__device__ __forceinline__ void tile_records(const ScanParams& p, const TileShared& sh, const uint8_t* tile,
                                             const uint16_t* list, unsigned long long ts, uint32_t own_len,
                                             unsigned long long limit, uint32_t* hist, uint32_t* lenh, Acc& acc,
                                             int warp, int lane)
{
    const uint32_t f = sh.front;
    const uint32_t nown = f + sh.own_count;                       // entries that may precede a record start
    const uint32_t nstored = min(f + sh.total_count, (uint32_t)LIST_CAP);
    // list entry j terminates global line (base - f + j); a record starts after every line = 3 (mod 4)
    const uint32_t gb = (uint32_t)((sh.base - f) & 3ull);
    const uint32_t j0 = (3u - gb) & 3u;
    const uint32_t own_end = FRONT + own_len;
    for (uint32_t j = j0 + 4u * warp; j < nown; j += 4u * NWARPS) {
        const uint32_t s = (uint32_t)list[j] + 1u;
        if (s >= own_end) break;                                   // starts in the next tile
        const unsigned long long abs_s = ts + s - FRONT;
        if (abs_s >= limit) break;
        if (j + 4 < nstored)
            record_smem<NCHUNK, ASCII>(p, tile, list, j, abs_s, hist, lenh, acc, lane);
        else
            record_global<NCHUNK>(p, abs_s, limit, hist, lenh, acc, lane);
    }
}

// ------------------------------------------------------------------------------------------
// the fused scan kernel: persistent CTAs, tiles handed out by ticket, line numbering chained
// across tiles with a decoupled look-back
// ------------------------------------------------------------------------------------------
template <int NCHUNK>
__global__ void __launch_bounds__(THREADS, (NCHUNK <= 5 ? 2 : 1)) fq_scan_kernel(const ScanParams p)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    using L = Smem<NCHUNK>;
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw);
    uint8_t* tile = smem_raw + L::HIST_BYTES;
    uint16_t* list = reinterpret_cast<uint16_t*>(tile + L::TILE_BYTES);
    uint32_t* lenh = reinterpret_cast<uint32_t*>(tile + L::TILE_BYTES + L::LIST_BYTES);
    __shared__ TileShared sh;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;

    unsigned long long limit = NONE64;
    if (p.flags & F_RERUN) {
        limit = p.res->first_bad;  // written by the first pass, stable during this launch
        if (limit == NONE64) return;
    }
    if ((p.flags & F_CARRY) && p.carry->status != 0) return;  // stream already failed
    const unsigned long long line_base = (p.flags & F_CARRY) ? p.carry->line_base : p.line_base;

    for (int i = tid; i < NCHUNK * CHUNK_WORDS; i += THREADS) hist[i] = 0;
    for (int i = tid; i < L::LENH_WORDS; i += THREADS) lenh[i] = 0;
    if (tid == 0) {
        mbar_init(&sh.mbar, 1);
        fence_mbar_init();
    }
    Acc acc = {0, 0, 0, 0};
    uint32_t parity = 0;
    uint32_t recs_since_flush = 0;
    __syncthreads();

    for (;;) {
        if (tid == 0) {
            sh.tile = atomicAdd(p.ticket, 1u);
            sh.nonascii = 0;
        }
        __syncthreads();  // (A) previous tile fully consumed; ticket visible
        const uint32_t t = sh.tile;
        if (t >= p.ntiles) break;
        const unsigned long long ts = (unsigned long long)t * TILE;
        if (ts >= limit) break;
        const uint32_t own_len = (uint32_t)min((unsigned long long)TILE, p.n_own - ts);
        const uint32_t data_len = (uint32_t)min((unsigned long long)(TILE + HALO), p.n_avail - ts);

        // ---- stage [ts-16, ts+data_len) into shared memory (TMA bulk copy) -----------------
        const uint32_t front = (ts || (p.flags & F_FRONT16)) ? FRONT : 0;
        const uint32_t span = front + data_len;
        const uint32_t bulk = span & ~15u;
        const uint8_t* src = p.data + ts - front;
        uint8_t* dst = tile + FRONT - front;
        if (tid == 0) {
            fence_proxy_async();
            mbar_arrive_expect_tx(&sh.mbar, bulk);
            if (bulk) bulk_g2s(dst, src, bulk, &sh.mbar);
        }
        if (span != (uint32_t)SM_TILE) {
            // ragged first / last tiles: leading zeros, trailing bytes and zero fill by hand
            const bool virt_nl = ts == 0 && front == 0 && (p.flags & F_LINE_START);  // stream/line start acts as a '\n' before byte 0
            for (uint32_t i = tid; i < FRONT - front; i += THREADS) tile[i] = (virt_nl && i == FRONT - 1) ? '\n' : 0;
            for (uint32_t i = bulk + tid; i < span; i += THREADS) dst[i] = src[i];
            for (uint32_t i = FRONT + data_len + tid; i < (uint32_t)SM_TILE; i += THREADS) tile[i] = 0;
        }
        mbar_wait(&sh.mbar, parity);
        parity ^= 1u;
        __syncthreads();  // (B) hand-written bytes visible

        // ---- pass 1: newline masks, per-unit counts --------------------------------------
        uint32_t mask[ITERS];
        uint32_t hib = 0;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int u = it * NWARPS + warp;
            mask[it] = 0;
            if (u < NUNITS) {
                const uint32_t off = (uint32_t)u * UNIT + (uint32_t)lane * 16u;  // relative to the tile start
                const uint4 v = *reinterpret_cast<const uint4*>(tile + FRONT + off);
                hib |= v.x | v.y | v.z | v.w;
                const uint32_t m = nlmask16(v);
                mask[it] = m;
                const int rem = (int)own_len - (int)off;
                const uint32_t ownm = rem >= 16 ? 0xFFFFu : (rem > 0 ? ((1u << rem) - 1u) : 0u);
                const uint32_t call = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(m));
                const uint32_t cown = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(m & ownm));
                if (lane == 0) {
                    sh.unit_all[u] = call;
                    sh.unit_own[u] = cown;
                }
            }
        }
        if (__any_sync(0xffffffffu, (hib & 0x80808080u) != 0) && lane == 0) sh.nonascii = 1;
        __syncthreads();  // (C)

        // ---- warp 0: unit prefix, publish aggregate, look back ------------------------------
        if (warp == 0) {
            const uint32_t f = tile[FRONT - 1] == '\n' ? 1u : 0u;
            uint32_t a0 = lane < NUNITS ? sh.unit_all[lane] : 0u;
            uint32_t a1 = (lane + 32) < NUNITS ? sh.unit_all[lane + 32] : 0u;
            uint32_t o0 = lane < NUNITS ? sh.unit_own[lane] : 0u;
            uint32_t o1 = (lane + 32) < NUNITS ? sh.unit_own[lane + 32] : 0u;
            uint32_t inc0 = a0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, inc0, d);
                if (lane >= d) inc0 += v;
            }
            const uint32_t tot0 = __shfl_sync(0xffffffffu, inc0, 31);
            uint32_t inc1 = a1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, inc1, d);
                if (lane >= d) inc1 += v;
            }
            const uint32_t tot1 = __shfl_sync(0xffffffffu, inc1, 31);
            if (lane < NUNITS) sh.unit_base[lane] = f + inc0 - a0;
            if (lane + 32 < NUNITS) sh.unit_base[lane + 32] = f + tot0 + inc1 - a1;
            const uint32_t own_count = __reduce_add_sync(0xffffffffu, o0 + o1);

            unsigned long long excl;
            if (t == 0) {
                excl = line_base;
            } else {
                if (lane == 0) st_volatile_u64(p.tile_status + t, ST_AGG | own_count);
                excl = 0;
                long long pos = (long long)t - 1;
                for (;;) {
                    const long long i = pos - lane;
                    unsigned long long v;
                    do {
                        v = i >= 0 ? ld_volatile_u64(p.tile_status + i) : ST_INC;
                    } while (__any_sync(0xffffffffu, (v >> 62) == 0));
                    const unsigned incm = __ballot_sync(0xffffffffu, (v >> 62) == 2);
                    const unsigned long long val = v & ST_VAL;
                    if (incm) {
                        const int first = __ffs(incm) - 1;
                        excl += warp_sum_u64(lane <= first ? val : 0ull);
                        break;
                    }
                    excl += warp_sum_u64(val);
                    pos -= 32;
                }
            }
            if (lane == 0) {
                st_volatile_u64(p.tile_status + t, ST_INC | (excl + own_count));
                sh.base = excl;
                sh.front = f;
                sh.own_count = own_count;
                sh.total_count = tot0 + tot1;
                if (f) list[0] = FRONT - 1;
                if (t == p.ntiles - 1 && !(p.flags & F_RERUN)) {
                    p.res->n_lines = excl + own_count - line_base;
                    p.res->line_end = excl + own_count;
                }
            }
        }
        __syncthreads();  // (D)

        const uint32_t f = sh.front;
        const uint32_t own_count = sh.own_count;
        const bool nonascii = sh.nonascii != 0;  // read between (D) and (E): thread 0 resets it at the loop top
        const bool overflow = f + sh.total_count > (uint32_t)LIST_CAP;
        const unsigned long long idx_base = sh.base - line_base;  // buffer-local line number of the first own '\n'
        const bool want_index = (p.flags & F_INDEX) && !(p.flags & F_RERUN) && p.index != nullptr;
        const unsigned long long off_base = p.stream_offset + ts;

        // ---- pass 2: rank every newline, fill the position list -----------------------------
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int u = it * NWARPS + warp;
            if (u < NUNITS) {
                uint32_t m = mask[it];
                const int c = __popc(m);
                int pre = 0;
                for (int k = 0;; ++k) {
                    const unsigned b = __ballot_sync(0xffffffffu, c > k);
                    if (!b) break;
                    pre += __popc(b & lt_mask);
                }
                uint32_t rank = sh.unit_base[u] + (uint32_t)pre;
                const uint32_t off = (uint32_t)u * UNIT + (uint32_t)lane * 16u;
                while (m) {
                    const uint32_t bit = (uint32_t)__ffs(m) - 1u;
                    m &= m - 1u;
                    if (rank < (uint32_t)LIST_CAP) list[rank] = (uint16_t)(FRONT + off + bit);
                    if (overflow && want_index && rank >= f && rank < f + own_count) {
                        const unsigned long long gi = idx_base + (rank - f);
                        if (gi < p.index_cap) p.index[gi] = (uint32_t)(off_base + off + bit);
                    }
                    ++rank;
                }
            }
        }
        __syncthreads();  // (E) list complete

        // ---- line-end index: coalesced copy of the owned part of the list ---------------------
        if (want_index && !overflow) {
            for (uint32_t i = tid; i < own_count; i += THREADS) {
                const unsigned long long gi = idx_base + i;
                if (gi < p.index_cap) p.index[gi] = (uint32_t)(off_base + (uint32_t)list[f + i] - FRONT);
            }
        }

        // ---- records that start in this tile --------------------------------------------------
        const uint32_t nrec_upper = (f + own_count) / 4u + 2u;
        if (recs_since_flush + nrec_upper > 65535u) {  // u16 counters could overflow: flush first
            flush_hist<NCHUNK>(hist, lenh, p);
            recs_since_flush = 0;
        }
        recs_since_flush += nrec_upper;

        if (!overflow) {
            if (nonascii)
                tile_records<NCHUNK, false>(p, sh, tile, list, ts, own_len, limit, hist, lenh, acc, warp, lane);
            else
                tile_records<NCHUNK, true>(p, sh, tile, list, ts, own_len, limit, hist, lenh, acc, warp, lane);
        } else if (warp == 0) {
            // dense-newline tile: walk its records one after the other in global memory
            const uint32_t gb = (uint32_t)((sh.base - f) & 3ull);
            const uint32_t j0 = (3u - gb) & 3u;  // < 4 <= LIST_CAP, always stored
            if (j0 < f + own_count) {
                unsigned long long s = ts + (uint32_t)list[j0] + 1u - FRONT;
                const unsigned long long tend = ts + own_len;
                // (a tile holds < TILE/6 records, so the u16 counters cannot overflow in here)
                while (s < tend && s < limit) {
                    const unsigned long long e = record_global<NCHUNK>(p, s, limit, hist, lenh, acc, lane);
                    if (e == NONE64) break;
                    s = e + 1;
                }
            }
        }
        // loop: barrier (A) at the top protects tile/list reuse
    }

    // ---- drain -----------------------------------------------------------------------------
    flush_hist<NCHUNK>(hist, lenh, p);
    if (lane == 0) {
        if (acc.n_records) atomicAdd(p.stats + 0, acc.n_records);
        if (acc.n_bases) atomicAdd(p.stats + 1, acc.n_bases);
        if (acc.clip_seq) atomicAdd(p.stats + 2, acc.clip_seq);
        if (acc.clip_qual) atomicAdd(p.stats + 3, acc.clip_qual);
    }
}

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
    const unsigned long long end = s + (avail < MAXREC ? avail : (unsigned long long)MAXREC);
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
        if (incomplete) status = avail >= MAXREC ? 4 : 5;
    }
    if (lane == 0) {
        r->status = status;
        r->err_offset = p.stream_offset + s;
    }
}

// zero the accumulators for the second pass -- only when there is something to redo
__global__ void fq_rerun_reset_kernel(const ScanParams p)
{
    if (p.res->first_bad == NONE64) return;
    const size_t n_stats = stats_words(p.max_len);
    const size_t n_seq = (size_t)p.max_len * 256;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = i0; i < n_stats; i += stride) p.stats[i] = 0;
    for (size_t i = i0; i < n_seq; i += stride) p.seqraw[i] = 0;
    for (size_t i = i0; i < p.ntiles; i += stride) p.tile_status[i] = 0;
    if (i0 == 0) *p.ticket = 0;
}

// fold the raw sequence-byte histogram into the six base classes (validate_dnan's alphabet,
// src/records.rs:29-33), publish the outcome, and (streaming) add the chunk into the totals
__global__ void fq_finalize_kernel(const ScanParams p, DevCarry* carry, unsigned long long* total)
{
    const uint32_t P = p.max_len;
    const bool skip = carry && carry->status != 0;  // stream failed in an earlier chunk
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long* base = p.stats + stats_base_off(P);
    if (!skip) {
        for (size_t pos = i0; pos < P; pos += stride) {
            const unsigned long long* row = p.seqraw + pos * 256;
            unsigned long long all = 0;
            for (int b = 0; b < 256; ++b) all += row[b];
            const unsigned long long a = row['A'], c = row['C'], g = row['G'], t = row['T'], n = row['N'];
            base[pos * 6 + 0] = a;
            base[pos * 6 + 1] = c;
            base[pos * 6 + 2] = g;
            base[pos * 6 + 3] = t;
            base[pos * 6 + 4] = n;
            base[pos * 6 + 5] = all - a - c - g - t - n;
        }
    }
    // grid-wide ordering is not needed: totals only read words this thread itself finalized or
    // words written by earlier kernels, except base_hist -> handle it in a second sweep below
    if (total && !skip) {
        const size_t nb = stats_base_off(P), nq = stats_qual_off(P), nw = stats_words(P);
        for (size_t i = i0; i < nw; i += stride) {
            if (i >= nb && i < nq) continue;  // base_hist: added by the owning thread below
            total[i] += p.stats[i];
        }
        for (size_t pos = i0; pos < P; pos += stride)
            for (int k = 0; k < 6; ++k) total[nb + pos * 6 + k] += base[pos * 6 + k];
    }
    if (i0 == 0) {
        DevResult* r = p.res;
        if (!skip) {
            r->n_records = p.stats[0];
            r->finished = r->status == 0 ? 1 : 0;
            if (carry) {
                carry->n_records += p.stats[0];
                carry->n_lines += r->n_lines;
                carry->line_base = r->line_end;
                if (r->status != 0) {
                    carry->status = r->status;
                    carry->err_offset = r->err_offset;
                }
            }
        }
    }
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
size_t scan_smem_bytes(int nchunk) { return nchunk <= 5 ? (size_t)Smem<5>::TOTAL : (size_t)Smem<10>::TOTAL; }

cudaError_t scan_configure()
{
    cudaError_t e = cudaFuncSetAttribute(fq_scan_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<5>::TOTAL);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(fq_scan_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<10>::TOTAL);
}

int scan_blocks_per_sm(int nchunk)
{
    int n = 0;
    cudaError_t e;
    if (nchunk <= 5)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fq_scan_kernel<5>, THREADS, Smem<5>::TOTAL);
    else
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fq_scan_kernel<10>, THREADS, Smem<10>::TOTAL);
    if (e != cudaSuccess || n < 1) n = 1;
    return n;
}

cudaError_t launch_scan(const ScanParams& p, int nchunk, int grid, cudaStream_t st)
{
    if (nchunk <= 5)
        fq_scan_kernel<5><<<grid, THREADS, Smem<5>::TOTAL, st>>>(p);
    else
        fq_scan_kernel<10><<<grid, THREADS, Smem<10>::TOTAL, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_diagnose(const ScanParams& p, DevCarry* carry, cudaStream_t st)
{
    fq_diagnose_kernel<<<1, 32, 0, st>>>(p, carry);
    return cudaGetLastError();
}

cudaError_t launch_rerun_reset(const ScanParams& p, cudaStream_t st)
{
    fq_rerun_reset_kernel<<<256, 256, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_finalize(const ScanParams& p, DevCarry* carry, unsigned long long* total, cudaStream_t st)
{
    fq_finalize_kernel<<<64, 256, 0, st>>>(p, carry, total);
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
