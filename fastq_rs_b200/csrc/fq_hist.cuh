// fq_hist.cuh -- device code shared by the two scan kernels (fq_stream.cu: the speculative
// warp-autonomous kernel; fq_scan.cu: the exact tile-pipelined kernel): shared-memory histogram
// layout and bump, newline masks, the per-record slow path, counter drain.
//
// The template parameter C of the helpers is the kernel configuration; it provides NCHUNK, PPAD,
// CHUNK_WORDS and HIST_WORDS.
#pragma once
#include "fq_common.cuh"
#include "fq_device.cuh"

namespace fq {

__device__ __forceinline__ void named_bar(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int OFF>
__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ void red_add(uint32_t addr, uint32_t v)
{
    asm volatile("red.shared.add.u32 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "r"(v));
}
__device__ __forceinline__ uint32_t dp4a_u(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// debug timeline: one clock64 stamp per (tile, event); no-op unless the host set p.trace (FQB_TRACE)
__device__ __forceinline__ void trace_ev(const ScanParams& p, int k, int ev, bool use_max = false)
{
    if (p.trace && k < TRACE_K) {
        unsigned long long* slot = p.trace + ((size_t)blockIdx.x * TRACE_K + k) * 16 + ev;
        if (use_max)
            atomicMax(slot, (unsigned long long)clock64());
        else
            *slot = (unsigned long long)clock64();
    }
}

// word index of hist[chunk][byte][position % 32]
template <class C>
__device__ __forceinline__ uint32_t hist_word(uint32_t byte, uint32_t pos)
{
    return (pos >> 5) * (uint32_t)C::CHUNK_WORDS + byte * 32u + (pos & 31u);
}

// newline mask of a 16-byte piece, shifted left by 7: bit 7 + i = byte i is '\n'
// (the four 0x80 flags of each word are gathered by dp4a with weights 1,2,4,8 / 16,32,64,128)
__device__ __forceinline__ uint32_t nlmask16s7(const uint4& v)
{
    const uint32_t m0 = nlbits(v.x), m1 = nlbits(v.y), m2 = nlbits(v.z), m3 = nlbits(v.w);
    const uint32_t lo = dp4a_u(m1, 0x80402010u, dp4a_u(m0, 0x08040201u, 0u));
    const uint32_t hi = dp4a_u(m3, 0x80402010u, dp4a_u(m2, 0x08040201u, 0u));
    return lo + (hi << 8);
}

// ------------------------------------------------------------------------------------------
// lock-free drain of the u16-pair counters: atomicExch leaves concurrent increments of the other
// warps intact, so a slice may be flushed whenever the CTA-wide record counter says a half could
// approach 65535
// ------------------------------------------------------------------------------------------
template <class C>
__device__ void flush_hist(uint32_t* hist, const ScanParams& p, int first, int last, int tid, int nthreads)
{
    const uint32_t P = p.max_len;
    unsigned long long* qual = p.stats + stats_qual_off(P);
    for (int i = first + tid; i < last; i += nthreads) {
        if (hist[i] == 0) continue;
        const uint32_t chunk = (uint32_t)i / C::CHUNK_WORDS, r = (uint32_t)i % C::CHUNK_WORDS;
        const uint32_t b = (r >> 5) + (uint32_t)C::ROW0, pos = chunk * 32u + (r & 31u);   // (first row = byte ROW0)
        if (b >= (uint32_t)HIST_ROWS) continue;      // spare row of the speculative kernel's table: never read back
        const uint32_t v = atomicExch(hist + i, 0u);
        const uint32_t lo = v & 0xFFFFu, hi = v >> 16;
        // a '\n' counted INSIDE a sequence or quality line: the lines were not what the (predicting)
        // delimiter took them for -- the exact path redoes the shard.  (Lines delimited by a scan end
        // at their first '\n', so this row stays empty there.)
        if (b == '\n') atomicExch(&p.res->spec_fail, 1);
        if (pos < P) {
            if (lo) atomicAdd(p.seqraw + (size_t)pos * 256 + b, (unsigned long long)lo);
            if (hi) atomicAdd(qual + (size_t)pos * 256 + b, (unsigned long long)hi);
        }
    }
}

template <class C>
__device__ __forceinline__ void account_record(Acc& acc, uint32_t* lenh, const ScanParams& p, uint32_t Ls, uint32_t Lq)
{
    const uint32_t P = p.max_len;
    acc.n_bases += Ls;
    if (Ls > P) acc.clip_seq += Ls - P;
    if (Lq > P) acc.clip_qual += Lq - P;
    const uint32_t lb = Ls <= P ? Ls : P + 1;
    if (lb < (uint32_t)C::PPAD + 2u)
        atomicAdd(lenh + lb, 1u);
    else
        atomicAdd(p.stats + stats_len_off(P) + lb, 1ull);
}

// ------------------------------------------------------------------------------------------
// slow path: one warp walks a record in global memory (longer than the halo, or inside a tile
// with more newlines than LIST_CAP).  Returns the offset of its final '\n', or NONE64 when the
// record was flagged (bad / incomplete) or lies beyond `limit`.
// ------------------------------------------------------------------------------------------
template <class C>
__device__ __noinline__ unsigned long long record_global(const ScanParams& p, unsigned long long s,
                                                         unsigned long long limit, uint32_t* hist, uint32_t* lenh,
                                                         int lane)
{
    if (s >= limit) return NONE64;
    const uint8_t* __restrict__ d = p.data;
    const unsigned long long navail = p.n_avail;
    // the reference's 68 KiB buffer keeps the record at buffer offset (stream offset mod 16) when it refills
    // (Buffer::clean, src/buffer.rs:51-72: the leftover is parked so that the next read is 16-byte aligned), so
    // a record fits iff (offset mod 16) + length <= BUFSIZE -- see rec_window()
    unsigned long long win_end = s + rec_window(p.stream_offset + s);
    const bool window_full = win_end <= navail;
    if (win_end > navail) win_end = navail;
    unsigned long long nl[4] = {0, 0, 0, 0};
    int found = 0;
    for (unsigned long long q = s; q < win_end && found < 4; q += 32) {
        const unsigned long long a = q + lane;
        const bool isnl = a < win_end && d[a] == '\n';
        unsigned m = __ballot_sync(0xffffffffu, isnl);
        while (m && found < 4) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            nl[found++] = q + b;
        }
    }
    bool bad = false, tail = false;
    if (found < 4) {
        // incomplete inside the window: too long if the window was full; otherwise the data ended --
        // an error at EOF, a tail to carry over when more bytes will follow (src/lib.rs:276-293)
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
    if (lane == 0) atomicAdd(p.stats + 0, 1ull);   // rare path: straight to the global counters
    if (p.flags & F_HIST) {
        const uint32_t P = p.max_len;
        const uint32_t Pm = P < (uint32_t)C::PPAD ? P : (uint32_t)C::PPAD;
        const uint32_t Lr = (uint32_t)(nl[1] - nl[0] - 1);
        const uint32_t Ls = Lr - ((Lr > 0 && d[nl[1] - 1] == '\r') ? 1u : 0u);
        const uint32_t Lq = Lr - ((Lr > 0 && d[nl[3] - 1] == '\r') ? 1u : 0u);
        const uint8_t* sq = d + nl[0] + 1;
        const uint8_t* ql = d + nl[2] + 1;
        unsigned long long* qualg = p.stats + stats_qual_off(P);
        const uint32_t ns = Ls < P ? Ls : P, nq = Lq < P ? Lq : P;
        for (uint32_t c = lane; c < ns; c += 32) {
            const uint32_t b = sq[c];
            if (c < Pm && b < (uint32_t)HIST_ROWS)
                atomicAdd(hist + hist_word<C>(b, c), 1u);
            else
                atomicAdd(p.seqraw + (size_t)c * 256 + b, 1ull);
        }
        for (uint32_t c = lane; c < nq; c += 32) {
            const uint32_t b = ql[c];
            if (c < Pm && b < (uint32_t)HIST_ROWS)
                atomicAdd(hist + hist_word<C>(b, c), 0x10000u);
            else
                atomicAdd(qualg + (size_t)c * 256 + b, 1ull);
        }
        if (lane == 0) {
            Acc a = {0, 0, 0, 0};
            account_record<C>(a, lenh, p, Ls, Lq);
            if (a.n_bases) atomicAdd(p.stats + 1, a.n_bases);
            if (a.clip_seq) atomicAdd(p.stats + 2, a.clip_seq);
            if (a.clip_qual) atomicAdd(p.stats + 3, a.clip_qual);
        }
    }
    return nl[3];
}

// ------------------------------------------------------------------------------------------
// records: one pass = 4 records per warp, 8 lanes each
// lane = 8*sub + i.  In round T lane (sub,i) owns the 4-byte group g = i + 8T of its record's
// sequence and quality lines and visits its bytes in the order (k + sub) & 3, k = 0..3, so that
// the k-th ATOMS of the round touches position 4g + ((k+sub)&3): over the 32 lanes these are 32
// different residues mod 32 = 32 different banks of hist[chunk][byte][position % 32].
// ------------------------------------------------------------------------------------------
struct LaneConst {         // fixed per lane for the whole kernel
    uint32_t hk[4];        // shared address of hist[0][0][pk[k]]
    uint32_t wsel[4];      // dp4a weights: 128 in the byte lane visited k-th
    uint32_t pk[4];        // position visited by the k-th bump in round 0
};

struct RoundCtx {
    uint32_t as0, aq0;     // shared addresses of the aligned words holding position 4i of seq / qual
    uint32_t shs, shq;     // funnel shifts that realign them
    uint32_t ns, nq;       // bytes of seq / qual that have a shared-memory column
    uint32_t nmax_w, nmin_w;
    unsigned long long *gseq, *gqual;   // global rows (non-ASCII bytes only)
};

template <class C, bool ASCII, int T>
struct Rounds {
    static __device__ __forceinline__ void run(const RoundCtx& c, const LaneConst& lc)
    {
        if (32u * T >= c.nmax_w) return;                                  // warp-uniform
        const uint32_t s0 = lds32<32 * T>(c.as0), s1 = lds32<32 * T + 4>(c.as0);
        const uint32_t q0 = lds32<32 * T>(c.aq0), q1 = lds32<32 * T + 4>(c.aq0);
        const uint32_t vs = __funnelshift_r(s0, s1, c.shs);
        const uint32_t vq = __funnelshift_r(q0, q1, c.shq);
        constexpr int CO = 4 * C::CHUNK_WORDS * T;                        // byte offset of chunk T
        if (ASCII && 32u * (T + 1) <= c.nmin_w) {                         // every lane's group lies inside both lines
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                red_add<CO>(dp4a_u(vs, lc.wsel[k], lc.hk[k]), 1u);
                red_add<CO>(dp4a_u(vq, lc.wsel[k], lc.hk[k]), 0x10000u);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t pos = lc.pk[k] + 32u * T;
                if (pos < c.ns) {
                    if (ASCII) {
                        red_add<CO>(dp4a_u(vs, lc.wsel[k], lc.hk[k]), 1u);
                    } else {
                        const uint32_t bs = (vs >> (8u * (lc.pk[k] & 3u))) & 0xFFu;
                        if (bs < (uint32_t)HIST_ROWS)
                            red_add<CO>(lc.hk[k] + bs * 128u, 1u);
                        else
                            atomicAdd(c.gseq + (size_t)pos * 256 + bs, 1ull);
                    }
                }
                if (pos < c.nq) {
                    if (ASCII) {
                        red_add<CO>(dp4a_u(vq, lc.wsel[k], lc.hk[k]), 0x10000u);
                    } else {
                        const uint32_t bq = (vq >> (8u * (lc.pk[k] & 3u))) & 0xFFu;
                        if (bq < (uint32_t)HIST_ROWS)
                            red_add<CO>(lc.hk[k] + bq * 128u, 0x10000u);
                        else
                            atomicAdd(c.gqual + (size_t)pos * 256 + bq, 1ull);
                    }
                }
            }
        }
        Rounds<C, ASCII, T + 1>::run(c, lc);
    }
};
template <class C, bool ASCII>
struct Rounds<C, ASCII, C::NCHUNK> {
    static __device__ __forceinline__ void run(const RoundCtx&, const LaneConst&) {}
};

// warp-exclusive prefix of small per-lane counts (most are 0, a few 1 or 2): ballot levels
__device__ __forceinline__ uint32_t small_prefix(int c, uint32_t lt_mask)
{
    const unsigned b1 = __ballot_sync(0xffffffffu, c > 0);
    const unsigned b2 = __ballot_sync(0xffffffffu, c > 1);
    const unsigned b3 = __ballot_sync(0xffffffffu, c > 2);
    uint32_t pre = (uint32_t)__popc(b1 & lt_mask) + (uint32_t)__popc(b2 & lt_mask);
    if (b3) {
        pre += (uint32_t)__popc(b3 & lt_mask);
        for (int lvl = 3;; ++lvl) {
            const unsigned b = __ballot_sync(0xffffffffu, c > lvl);
            if (!b) break;
            pre += (uint32_t)__popc(b & lt_mask);
        }
    }
    return pre;
}

}  // namespace fq
