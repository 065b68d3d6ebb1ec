// fq_stream.cu -- the SPECULATIVE sm_100a scan kernel (K1 delimit + K2 per-position histograms):
// 32 autonomous warps per SM, each streaming through its own contiguous byte range the way the
// reference's Buffer does (src/buffer.rs:30-100, src/lib.rs:255-303) -- load a window at the cursor,
// delimit the complete records in it, consume them, move the cursor to the first record that did not
// fit -- with no synchronisation between warps at all:
//
//   window      one TMA bulk copy (UBLKCP, 4 KB, 16-byte aligned source) into the warp's own buffer,
//               completing on the warp's own mbarrier
//   scan        16-byte SWAR newline masks (3 ops / word + dp4a bit gather), warp prefix of the
//               counts, line starts into a u16 list in shared memory
//   records     8 lanes per record: '@' / '+' / raw-length validation (src/records.rs:201-247), then
//               the histogram rounds of fq_hist.cuh (one dp4a + one ATOMS per byte, bank-conflict free)
//   index       the line ends of the window, ranked relative to the range, into a staging area
//
// What makes it speculative: a range other than the first does not know where its first record
// starts.  It INFERS it -- of four consecutive line starts exactly one is a record start; the warp
// tests the four candidates against the grammar over the following records and takes the only one
// that holds -- and fq_stream_verify_kernel afterwards checks that the record chain of every range
// ends exactly where the next range started.  Chained from the first range, whose start is known,
// that proves the ranges saw the very records a sequential parse delivers.  Anything else --
// a record that fails validation, a record longer than the window, bytes >= 0x80, an ambiguous or
// mis-inferred start, a staging area too small -- only raises res->spec_fail, and the exact path
// (fq_scan.cu) redoes the shard.  Results never depend on the inference.
//
// Second speculation, same rule (verify, else fall back): PREDICTED windows.  While the records keep
// the shape of the last record a scan delimited, a window is not scanned; its records are located by
// arithmetic and verified -- with histograms by per-record byte checks plus the histogram's own '\n'
// row (pred_pass), without by byte checks plus the '\n' count of the window (win_count_newlines).
// A window that does not verify is scanned.  This is what takes the kernel from ~2.2 to ~3.3 TB/s with
// histograms and from ~4 to ~5.7 TB/s without, on fixed-length reads; reads of varying length never
// predict and pay two failed attempts per 34 windows.
#include "fq_hist.cuh"

#include <cstdlib>

namespace fq {

template <int NCHUNK_, int NWARPS_, int WIN_, int SPARE_ROW_ = 0, int ROW0_ = 0>
struct SCfg {
    static constexpr int NCHUNK = NCHUNK_;
    static constexpr int NWARPS = NWARPS_;                 // autonomous warps per CTA
    static constexpr int NTHREADS = 32 * NWARPS_;
    static constexpr int PPAD = 32 * NCHUNK;
    // byte rows ROW0 .. 127 (+ a spare row in the variable-length variant, see mask_to_trash).  ROW0 = 32 leaves
    // out the control characters: 25 % less table, which buys the P <= 320 variant 4 KiB windows.  A byte
    // below ROW0 inside a sequence or quality line is noticed (line_steps) and sends the shard to the exact
    // path; its bump lands ROW0 rows below its chunk -- in the previous chunk, or in the PRE bytes in front
    // of the table (which is where the length histogram lives then): never outside the CTA's shared memory.
    static constexpr int ROW0 = ROW0_;
    static constexpr int ROWS = HIST_ROWS - ROW0_;             // rows read back (the spare row follows them)
    static constexpr int CHUNK_WORDS = (ROWS + SPARE_ROW_) * 32;
    static constexpr int HIST_WORDS = NCHUNK * CHUNK_WORDS;
    static constexpr int LENH_WORDS = (PPAD + 2 + 31) / 32 * 32;
    static constexpr int PRE = ROW0_ ? (ROW0_ * 128 > LENH_WORDS * 4 ? ROW0_ * 128 : LENH_WORDS * 4) : 0;
    static constexpr int WIN = WIN_;                       // window bytes (a multiple of 512)
    static constexpr int NU = WIN / UNIT;                  // 512-byte units per window
    static constexpr int LIST_N = WIN_ >= 4096 ? 192 : 128;   // u16 entries: [0] = cursor, [j] = start of the line after the j-th '\n'
    static constexpr int MAXR = (LIST_N - 12) / 4;         // records consumed per window at most
    static constexpr int LIST_DUMMY = LIST_N - 1;          // writes beyond the capacity land here
    static constexpr int WARP_BYTES = WIN + 16 + LIST_N * 2;
    static constexpr int TAIL_PAD = 256;                   // word loads of the rounds may run past the last buffer
    static constexpr int BUF0 = PRE + HIST_WORDS * 4 + (ROW0_ ? 0 : LENH_WORDS * 4);   // offset of the warp buffers
    static constexpr int TOTAL = BUF0 + NWARPS * WARP_BYTES + TAIL_PAD;
    static_assert(WIN % UNIT == 0 && WIN <= 65520, "window");
    static_assert(TOTAL <= 232448 - 1024, "shared memory per CTA");
    static_assert(WARP_BYTES % 16 == 0, "TMA destination alignment");
};

// The two masks of the newline test live in constant memory so that ptxas keeps them in registers /
// constant-bank operands: (w ^ A) & B is then ONE LOP3 instead of two with immediates.
static __constant__ uint32_t fq_kmask[2] = {0x0A0A0A0Au, 0x7F7F7F7Fu};

// 0x80 in every byte of w that equals '\n' (exact), 3 instructions: LOP3, IADD, LOP3
__device__ __forceinline__ uint32_t nlbits3(uint32_t w, uint32_t kA, uint32_t kB)
{
    uint32_t y;
    asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(y) : "r"(w), "r"(kA), "r"(kB));   // (w ^ A) & B
    const uint32_t t = y + kB;
    return ~(t | w) & 0x80808080u;
}
// newline mask of a 16-byte piece, shifted left by 7 (see nlmask16s7)
__device__ __forceinline__ uint32_t nlmask16k(const uint4& v, uint32_t kA, uint32_t kB)
{
    const uint32_t m0 = nlbits3(v.x, kA, kB), m1 = nlbits3(v.y, kA, kB), m2 = nlbits3(v.z, kA, kB), m3 = nlbits3(v.w, kA, kB);
    const uint32_t lo = dp4a_u(m1, 0x80402010u, dp4a_u(m0, 0x08040201u, 0u));
    const uint32_t hi = dp4a_u(m3, 0x80402010u, dp4a_u(m2, 0x08040201u, 0u));
    return lo + (hi << 8);
}

// shared-memory accesses by 32-bit shared address (no generic-pointer arithmetic in the hot loops)
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v));
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <int OFF>
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

struct Window {
    long long src;            // buffer-relative offset of the window (16-byte aligned; -16 for the window that
                              // looks at the byte in front of the shard)
    uint32_t pad;             // cursor - src
    uint32_t vlen;            // valid bytes in the window
};

// ------------------------------------------------------------------------------------------
// window load: the analogue of Buffer::clean + read_into, src/buffer.rs:51-100 -- nothing is moved,
// the next window simply starts at the 16-byte boundary below the first record that did not fit
// ------------------------------------------------------------------------------------------
template <class C>
__device__ __forceinline__ Window win_load(const ScanParams& p, uint8_t* buf, unsigned long long* bar, uint32_t& parity,
                                           long long cur, int lane)
{
    Window w;
    w.src = cur & ~15ll;
    w.pad = (uint32_t)(cur - w.src);
    w.vlen = (uint32_t)min((long long)C::WIN, (long long)p.n_avail - w.src);
    const uint32_t bulk = w.vlen & ~15u;
    __syncwarp();
    if (lane == 0) {
        fence_proxy_async();
        mbar_arrive_expect_tx(bar, bulk);
        if (bulk) bulk_g2s(buf, p.data + w.src, bulk, bar);
#ifndef FQ_NO_L2_PREFETCH
        // the window after this one starts somewhere in the last few hundred bytes of it: have L2 fetch what follows
        // while this window is being worked on (the warp has one buffer; this turns the next load's DRAM latency
        // into an L2 hit)
        if (w.vlen == (uint32_t)C::WIN && w.src + 2 * C::WIN <= (long long)p.n_avail)
            bulk_prefetch_l2(p.data + w.src + C::WIN, (uint32_t)C::WIN);
#endif
    }
    // the bytes the 16-byte-granular bulk copy leaves out (last window of the shard only)
    if ((w.vlen & 15u) && (uint32_t)lane < (w.vlen & 15u)) buf[bulk + lane] = p.data[w.src + bulk + lane];
    mbar_wait(bar, parity);
    parity ^= 1u;
    __syncwarp();
    return w;
}

// One pass over the window: newline masks, ranks and the list of line starts
//   list[0] = cursor, list[j] = position after the j-th '\n' of [pad, vlen)
// The ranks need no second pass: the running count is warp-local.  Per unit (one piece per lane) the per-lane
// counts (0, 1, rarely 2) are prefix-summed with two ballots; a piece with three or more newlines
// sends the unit through the generic path.  Returns the newline count; HIB: hib = OR of all words.
// RAGGED: the window is shorter than WIN (the end of the shard) -- stale bytes are masked out.

// ranks + list entries of one unit: bit b of this lane's mask mm = a '\n' whose list entry is pos1 + b
template <class C>
__device__ __forceinline__ void scan_rank_store(uint32_t mm, uint32_t pos1, uint32_t& ubase, uint32_t list_s,
                                                uint16_t* list, uint32_t lt_mask)
{
    const int c = __popc(mm);
    const unsigned b1 = __ballot_sync(0xffffffffu, mm != 0);
    const unsigned b2 = __ballot_sync(0xffffffffu, c > 1);
    const unsigned b3 = __ballot_sync(0xffffffffu, c > 2);
    if (!b3) {
        const uint32_t rank = min(ubase + (uint32_t)__popc(b1 & lt_mask) + (uint32_t)__popc(b2 & lt_mask),
                                  (uint32_t)C::LIST_DUMMY - 1u);
        const uint32_t a = list_s + 2u * rank;
        // the first and the last newline of the piece: lowest and highest set bit
        if (mm != 0) sts16(a, pos1 + (uint32_t)__clz(__brev(mm)));
        if (c > 1) sts16(a + 2u, pos1 + 31u - (uint32_t)__clz(mm));
        ubase += (uint32_t)__popc(b1) + (uint32_t)__popc(b2);
    } else {
        uint32_t rank = ubase + small_prefix(c, lt_mask);
        while (mm) {
            list[min(rank, (uint32_t)C::LIST_DUMMY)] = (uint16_t)(pos1 + (uint32_t)__ffs(mm) - 1u);
            mm &= mm - 1u;
            ++rank;
        }
        ubase += __reduce_add_sync(0xffffffffu, (uint32_t)c);
    }
}

// WIDE: units of 1 KiB, 32 contiguous bytes per lane (two 16-byte loads -- lanes 4..7 of every 8 take their
// second piece first, so that each quarter warp's LDS.128 touches all 32 banks once), one set of ballots,
// ranks and list stores per 32 bytes instead of per 16; what is left of the window (WIN % 1024) goes in
// 512-byte units.  The variable-length variant scans every window: this is ~1/3 off its scan.
template <class C, bool RAGGED, bool HIB, bool WIDE>
__device__ __forceinline__ uint32_t win_scan_t(uint32_t buf_s, uint16_t* list, const Window& w, uint32_t& hib, int lane,
                                               uint32_t lt_mask)
{
    const uint32_t kA = fq_kmask[0], kB = fq_kmask[1];
    const uint32_t list_s = buf_s + (uint32_t)(C::WIN + 16);
    if (lane == 0) sts16(list_s, w.pad);
    uint32_t ubase = 1;
    hib = 0;
    constexpr int NW = WIDE ? C::WIN / 1024 : 0;                           // 1 KiB units
    constexpr int N16 = (C::WIN - 1024 * NW) / UNIT;                       // 512-byte units behind them
    if (WIDE) {
        const uint32_t odd = ((uint32_t)lane >> 2) & 1u;
#pragma unroll
        for (int it = 0; it < NW; ++it) {
            const uint32_t off = (uint32_t)it * 1024u + (uint32_t)lane * 32u;
            const uint4 va = lds_v4(buf_s + off + 16u * odd);
            const uint4 vb = lds_v4(buf_s + off + 16u * (odd ^ 1u));
            if (HIB) hib |= va.x | va.y | va.z | va.w | vb.x | vb.y | vb.z | vb.w;
            const uint32_t ma = nlmask16k(va, kA, kB), mb = nlmask16k(vb, kA, kB);   // bit 7 + i = byte i of the piece
            uint32_t mm = ((odd ? mb : ma) >> 7) | ((odd ? ma : mb) << 9);           // bit i = byte i of the 32 bytes
            if (it == 0 && lane == 0) mm &= ~((1u << w.pad) - 1u);                   // bytes before the cursor
            if (RAGGED) {                                                            // stale bytes beyond the data
                const int rem = (int)w.vlen - (int)off;
                mm &= rem >= 32 ? 0xFFFFFFFFu : (rem > 0 ? ((1u << rem) - 1u) : 0u);
            }
            scan_rank_store<C>(mm, off + 1u, ubase, list_s, list, lt_mask);
        }
    }
#pragma unroll
    for (int it = 0; it < N16; ++it) {
        const uint32_t off = (uint32_t)(1024 * NW + it * UNIT) + (uint32_t)lane * 16u;
        const uint4 v = lds_v4(buf_s + off);
        if (HIB) hib |= v.x | v.y | v.z | v.w;
        uint32_t mm = nlmask16k(v, kA, kB);                                 // bit 7 + i = byte i of the piece
        if (NW == 0 && it == 0 && lane == 0) mm &= ~(((1u << w.pad) - 1u) << 7);      // bytes before the cursor
        if (RAGGED) {                                                       // stale bytes beyond the data
            const int rem = (int)w.vlen - (int)off;
            const uint32_t m = rem >= 16 ? 0xFFFFu : (rem > 0 ? ((1u << rem) - 1u) : 0u);
            mm &= m << 7;
        }
        scan_rank_store<C>(mm, off - 7u + 1u, ubase, list_s, list, lt_mask);
    }
    __syncwarp();
    return ubase - 1u;
}

template <class C, bool HIB, bool WIDE = false>
__device__ __forceinline__ uint32_t win_scan(uint32_t buf_s, uint16_t* list, const Window& w, uint32_t& hib, int lane,
                                             uint32_t lt_mask)
{
    if (w.vlen < (uint32_t)C::WIN) return win_scan_t<C, true, HIB, WIDE>(buf_s, list, w, hib, lane, lt_mask);
    return win_scan_t<C, false, HIB, WIDE>(buf_s, list, w, hib, lane, lt_mask);
}

// ------------------------------------------------------------------------------------------
// where does the first record of the range start?  list[cf .. cf + 3] are four consecutive line
// starts: lane = 8 * (c - cf) + r tests record r of candidate c ('@' at its start, '+' after its
// sequence line, equal raw lengths).  Accepted only if exactly one candidate passes every record it
// could test, at least two.  Returns c, or NO_START when the start is ambiguous.
// ------------------------------------------------------------------------------------------
constexpr uint32_t NO_START = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t infer_start(const uint8_t* buf, const uint16_t* list, uint32_t nstored, uint32_t cf,
                                                int lane)
{
    const uint32_t cand = cf + ((uint32_t)lane >> 3), r = (uint32_t)lane & 7u;
    const uint32_t j = cand + 4u * r;
    const bool testable = j + 4u <= nstored;
    bool good = true;
    if (testable) {
        const uint32_t s = list[j], h = list[j + 1] - 1u, q = list[j + 2] - 1u, pp = list[j + 3] - 1u, e = list[j + 4] - 1u;
        good = buf[s] == '@' && buf[q + 1] == '+' && (e - pp) == (q - h);
    }
    const unsigned tested = __ballot_sync(0xffffffffu, testable);
    const unsigned bad = __ballot_sync(0xffffffffu, testable && !good);
    uint32_t pass = 0, npass = 0;
#pragma unroll
    for (uint32_t c = 0; c < 4; ++c) {
        const unsigned m = 0xFFu << (8 * c);
        if (__popc(tested & m) >= 2 && !(bad & m)) {
            pass = cf + c;
            ++npass;
        }
    }
    return npass == 1 ? pass : NO_START;
}

// ------------------------------------------------------------------------------------------
// one pass = 4 records, 8 lanes each (same lane mapping as the exact kernel; see fq_hist.cuh)
// ------------------------------------------------------------------------------------------
struct WinAcc {   // per-lane sums over one window, folded into the CTA counters once per window
    uint32_t n_records, n_bases;
};

struct LaneK {    // fixed per lane for the whole kernel (see LaneConst; positions are implied by hk)
    uint32_t hk[4];        // shared address of hist[0][0][position visited k-th in round 0]
    uint32_t wsel[4];      // dp4a weights: 128 in the byte lane visited k-th
};

template <class C, int T>
struct SRounds {
    // inc_s / inc_q: 1 / 0x10000 for lanes with a record, 0 for the others (their bumps add nothing)
    static __device__ __forceinline__ void run(uint32_t as0, uint32_t aq0, uint32_t shs, uint32_t shq, uint32_t ns,
                                               uint32_t nq, uint32_t nmax_w, uint32_t nmin_w, uint32_t inc_s,
                                               uint32_t inc_q, uint32_t hist_s, const LaneK& lc, uint32_t& hib)
    {
        if (32u * T >= nmax_w) return;                                    // warp-uniform
        const uint32_t s0 = lds32<32 * T>(as0), s1 = lds32<32 * T + 4>(as0);
        const uint32_t q0 = lds32<32 * T>(aq0), q1 = lds32<32 * T + 4>(aq0);
        const uint32_t vs = __funnelshift_r(s0, s1, shs);
        const uint32_t vq = __funnelshift_r(q0, q1, shq);
        hib |= vs | vq;                                                   // (see pred_pass: bytes >= 0x80)
        constexpr int CO = 4 * C::CHUNK_WORDS * T;                        // byte offset of chunk T
        if (32u * (T + 1) <= nmin_w) {                                    // every record's group lies inside both lines
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                red_add<CO>(dp4a_u(vs, lc.wsel[k], lc.hk[k]), inc_s);
                red_add<CO>(dp4a_u(vq, lc.wsel[k], lc.hk[k]), inc_q);
            }
        } else {
            // position visited k-th = (hk[k] - hist_s) / 4; it lies in the line iff it is below the bytes
            // the line has left for this round
            const int ts = (int)hist_s + 4 * ((int)ns - 32 * T), tq = (int)hist_s + 4 * ((int)nq - 32 * T);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // branch-free: a byte beyond its line bumps by zero (a branch around the asm costs a
                // divergence region per bump; the address stays inside the table for bytes < 0x80)
                red_add<CO>(dp4a_u(vs, lc.wsel[k], lc.hk[k]), (int)lc.hk[k] < ts ? inc_s : 0u);
                red_add<CO>(dp4a_u(vq, lc.wsel[k], lc.hk[k]), (int)lc.hk[k] < tq ? inc_q : 0u);
            }
        }
        SRounds<C, T + 1>::run(as0, aq0, shs, shq, ns, nq, nmax_w, nmin_w, inc_s, inc_q, hist_s, lc, hib);
    }
};
template <class C>
struct SRounds<C, C::NCHUNK> {
    static __device__ __forceinline__ void run(uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t,
                                               uint32_t, uint32_t, uint32_t, uint32_t, const LaneK&, uint32_t&)
    {
    }
};

// The same rounds for records of a predicted shape with seq() and qual() of one length n, NR = ceil(n / 32) rounds
// known at compile time (the caller dispatches once per window): straight-line code, no branch between the rounds,
// and the words of round T + 1 are loaded before the bumps of round T (shared-memory atomics and loads stay in
// program order, so the loads would otherwise wait behind them and the funnel shift behind the loads).  Only the last
// round can be partial, and which of a lane's four bumps of that round lie inside the line is the same for every
// record of the shape -- byte k of gm is 1 where the k-th bump counts (0 for lanes without a record), so the guard is
// one PRMT per bump instead of a compare and a select.
template <int OFF>
__device__ __forceinline__ uint32_t lds32_ordered(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <class C, int NR, int T>
struct FRounds {
    static __device__ __forceinline__ void run(uint32_t as0, uint32_t aq0, uint32_t shs, uint32_t shq, uint32_t s0,
                                               uint32_t s1, uint32_t q0, uint32_t q1, uint32_t inc_s, uint32_t inc_q,
                                               uint32_t gm, const LaneK& lc, uint32_t& hib)
    {
        const uint32_t vs = __funnelshift_r(s0, s1, shs);
        const uint32_t vq = __funnelshift_r(q0, q1, shq);
        if (T + 1 < NR) {
            s0 = lds32_ordered<32 * (T + 1)>(as0), s1 = lds32_ordered<32 * (T + 1) + 4>(as0);
            q0 = lds32_ordered<32 * (T + 1)>(aq0), q1 = lds32_ordered<32 * (T + 1) + 4>(aq0);
        }
        hib |= vs | vq;
        constexpr int CO = 4 * C::CHUNK_WORDS * T;
        if (T + 1 < NR) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                red_add<CO>(dp4a_u(vs, lc.wsel[k], lc.hk[k]), inc_s);
                red_add<CO>(dp4a_u(vq, lc.wsel[k], lc.hk[k]), inc_q);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                red_add<CO>(dp4a_u(vs, lc.wsel[k], lc.hk[k]), __byte_perm(gm, 0u, 0x4440u + (uint32_t)k));
                red_add<CO>(dp4a_u(vq, lc.wsel[k], lc.hk[k]), __byte_perm(gm, 0u, 0x4044u + ((uint32_t)k << 8)));
            }
        }
        FRounds<C, NR, T + 1>::run(as0, aq0, shs, shq, s0, s1, q0, q1, inc_s, inc_q, gm, lc, hib);
    }
};
template <class C, int NR>
struct FRounds<C, NR, NR> {
    static __device__ __forceinline__ void run(uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t,
                                               uint32_t, uint32_t, uint32_t, uint32_t, const LaneK&, uint32_t&)
    {
    }
};
template <class C, int NR>
__device__ __forceinline__ void frounds(uint32_t sa, uint32_t qa, uint32_t inc_s, uint32_t inc_q, uint32_t gm,
                                        const LaneK& lc, uint32_t& hib)
{
    if (NR == 0) return;
    const uint32_t as0 = sa & ~3u, aq0 = qa & ~3u;
    const uint32_t s0 = lds32_ordered<0>(as0), s1 = lds32_ordered<4>(as0);
    const uint32_t q0 = lds32_ordered<0>(aq0), q1 = lds32_ordered<4>(aq0);
    FRounds<C, NR, 0>::run(as0, aq0, sa << 3, qa << 3, s0, s1, q0, q1, inc_s, inc_q, gm, lc, hib);
}

// bytes [from, to) of the window (to - from <= 32) hold no '\n': lane i of the 8-lane group looks at
// the 4 bytes from + 4i .. (two aligned words, funnel shift); lane-local answer, the caller votes
__device__ __forceinline__ bool no_newline32(uint32_t buf_s, uint32_t from, uint32_t to, uint32_t i, uint32_t kA, uint32_t kB)
{
    const uint32_t a = buf_s + from + 4u * i;
    const uint32_t w0 = lds32<0>(a & ~3u), w1 = lds32<4>(a & ~3u);
    const uint32_t v = __funnelshift_r(w0, w1, a << 3);
    const int left = (int)to - (int)(from + 4u * i);                      // bytes of the range in this lane's word
    const uint32_t m = left >= 4 ? 0xFFFFFFFFu : (left > 0 ? ((1u << (8 * left)) - 1u) : 0u);
    return (nlbits3(v, kA, kB) & m) == 0u;
}

// One pass over 4 records whose line starts come from the scan's list.  Returns the window-relative
// number of the first record of the pass that failed validation (NO_START if none did).
template <class C, bool HIST>
__device__ __forceinline__ uint32_t stream_pass(const ScanParams& p, const uint8_t* buf, uint32_t buf_s,
                                                const LaneK& lc, uint32_t hist_s, uint32_t* lenh, uint32_t Pm,
                                                uint32_t n_rec, uint32_t pass, WinAcc& wa, uint32_t sub, uint32_t i)
{
    const uint32_t r = 4u * pass + sub;
    const bool valid = r < n_rec;
    // lanes without a record look at record 0 of the window (real data, harmless) and are masked out below
    const uint32_t lp = buf_s + (uint32_t)(C::WIN + 16) + (valid ? 8u * r : 0u);
    const uint32_t s = lds_u16<0>(lp), h = lds_u16<2>(lp) - 1u, q = lds_u16<4>(lp) - 1u, pp = lds_u16<6>(lp) - 1u,
                   e = lds_u16<8>(lp) - 1u;
    const uint32_t c_at = lds_u8(buf_s + s), c_plus = lds_u8(buf_s + q + 1u), c_sr = lds_u8(buf_s + q - 1u),
                   c_qr = lds_u8(buf_s + e - 1u);
    // src/records.rs:137-149 ('@'), :151-163 ('+'), :233-238 (raw line lengths equal)
    const bool good = c_at == '@' && c_plus == '+' && (e - pp) == (q - h);
    bool ok = valid && good;
    uint32_t first_bad = NO_START;
    {
        const unsigned nok = __ballot_sync(0xffffffffu, valid && !good);
        if (nok) {
            // nothing behind the first bad record is counted (Parser::each delivers the records before it)
            const uint32_t fsub = ((uint32_t)__ffs(nok) - 1u) >> 3;
            first_bad = 4u * pass + fsub;
            ok = ok && sub < fsub;
        }
    }
    if (HIST) {
        uint32_t Ls = 0, Lq = 0;
        if (ok) {
            const uint32_t Lr = q - h - 1u;
            // seq()/qual() drop one trailing '\r' (src/records.rs:65-73,82-90)
            Ls = Lr - ((Lr > 0 && c_sr == '\r') ? 1u : 0u);
            Lq = Lr - ((Lr > 0 && c_qr == '\r') ? 1u : 0u);
            if (i == 0) {
                wa.n_records++;
                wa.n_bases += Ls;
                if (max(Ls, Lq) > p.max_len) {                              // longer than the tracked positions: rare
                    if (Ls > p.max_len) atomicAdd(p.stats + 2, (unsigned long long)(Ls - p.max_len));
                    if (Lq > p.max_len) atomicAdd(p.stats + 3, (unsigned long long)(Lq - p.max_len));
                }
                if (Ls > p.max_len)
                    atomicAdd(p.stats + stats_len_off(p.max_len) + p.max_len + 1, 1ull);
                else
                    atomicAdd(lenh + Ls, 1u);
            }
        }
        const uint32_t ns = min(Ls, Pm), nq = min(Lq, Pm);
        const uint32_t nmax_w = __reduce_max_sync(0xffffffffu, max(ns, nq));
        // lanes without a record do not hold the fast rounds back: their bumps add zero
        const uint32_t nmin_w = __reduce_min_sync(0xffffffffu, ok ? min(ns, nq) : 0xFFFFFFFFu);
        const uint32_t sa = buf_s + h + 1u + 4u * i;       // shared address of position 4i of the sequence line
        const uint32_t qa = buf_s + pp + 1u + 4u * i;      // (buf_s is 16-byte aligned; the funnel shift takes sa * 8 mod 32)
        uint32_t hib_unused = 0;   // (the scan has looked at every byte of the window already)
        SRounds<C, 0>::run(sa & ~3u, qa & ~3u, sa << 3, qa << 3, ns, nq, nmax_w, nmin_w, ok ? 1u : 0u, ok ? 0x10000u : 0u,
                           hist_s, lc, hib_unused);
        // positions beyond the shared-memory columns but below P: straight to global (P > PPAD only)
        if (p.max_len > Pm && ok) {
            const uint32_t gs = min(Ls, p.max_len), gq = min(Lq, p.max_len);
            unsigned long long* gqual = p.stats + stats_qual_off(p.max_len);
            for (uint32_t g = Pm + i; g < gs; g += 8) atomicAdd(p.seqraw + (size_t)g * 256 + buf[h + 1u + g], 1ull);
            for (uint32_t g = Pm + i; g < gq; g += 8) atomicAdd(gqual + (size_t)g * 256 + buf[pp + 1u + g], 1ull);
        }
    } else if (ok && i == 0) {
        wa.n_records++;
    }
    return first_bad;
}

// ------------------------------------------------------------------------------------------
// SCANNED windows: records of any shape, delimited by the scan's list.
//   validate_block  lane j <-> record j of the window (32 at a time): '@' / '+' / raw-length validation
//                   (src/records.rs:201-247), '\r' trim, totals, length histogram -- all lane-parallel
//   line_steps      then ONE record at a time for the whole warp: lanes 0..15 take the sequence line, lanes
//                   16..31 the quality line, 4 bytes per lane and step = 64 positions of both lines per step.
//                   Reads of any length keep all lanes busy up to the last step of each record (no waiting
//                   for the longest of four records as with 8 lanes per record), and only that last step
//                   is guarded.  Bank = position % 32 as everywhere: lane (sub = lane >> 3, i = lane & 7)
//                   visits byte (k + sub) & 3 of its word in its k-th bump, so the 32 lanes of every ATOMS hit
//                   32 different banks (the two halves use the lo / hi half of the same counters).
// ------------------------------------------------------------------------------------------
struct StepK {    // per-lane constants of line_steps (made once per stretch of variable-shape windows)
    uint32_t lane_base;    // buf_s + 4 * (lane & 15)
    uint32_t lane_pos;     // 4 * (lane & 15): position of this lane's word in step 0
    uint32_t inc;          // 1 for the sequence half, 0x10000 for the quality half
    uint32_t gofs;         // byte offset of the chunk of the lane's word within a step (0 or one chunk)
    uint32_t hk[4];        // shared address of hist[chunk of the lane's word in step 0][0][position visited k-th]
    uint32_t wsel[4];      // dp4a weights: 128 in the byte lane visited k-th
};

// bytes outside the line become 0x80: their bumps land in the spare row 128 of the chunk (never read back),
// at the lane's usual bank -- no per-bump guard, no select
__device__ __forceinline__ uint32_t mask_to_trash(uint32_t v, uint32_t m)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE2;" : "=r"(r) : "r"(v), "r"(m), "r"(0x80808080u));   // (v & m) | (~m & K)
    return r;
}
// 0xFF in the low `n` bytes (n clamped to 0..4)
__device__ __forceinline__ uint32_t low_bytes_mask(int n)
{
    const uint32_t l = (uint32_t)max(min(n, 4), 0);
    uint32_t m;
    asm("shr.b32 %0, %1, %2;" : "=r"(m) : "r"(0xFFFFFFFFu), "r"(32u - 8u * l));          // shift amounts >= 32 give 0
    return m;
}

// one step in which every lane's word lies inside its line: no guard (S: compile-time step number)
// (lob: AND of (byte | byte << 1) over the bytes counted: bit 6 stays set iff all of them are >= 32 -- only
// kept when the table leaves out the rows below ROW0 = 32)
template <class C, int S>
__device__ __forceinline__ void full_step(uint32_t a0, uint32_t sh, const StepK& sk, uint32_t& hib, uint32_t& lob)
{
    const uint32_t w0 = lds32<64 * S>(a0), w1 = lds32<64 * S + 4>(a0);
    const uint32_t v = __funnelshift_r(w0, w1, sh);
    constexpr int CO = 4 * C::CHUNK_WORDS * 2 * S;                        // byte offset of chunk 2 S
    hib |= v;
    if (C::ROW0) lob &= v | (v << 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) red_add<CO>(dp4a_u(v, sk.wsel[k], sk.hk[k]), sk.inc);
}
// the step in which the lines end: bytes beyond the line go to the spare row (left0: bytes of the line from
// this lane's word of step 0 on)
template <class C, int S>
__device__ __forceinline__ void last_step(uint32_t a0, uint32_t sh, int left0, const StepK& sk, uint32_t& hib,
                                          uint32_t& lob)
{
    const uint32_t w0 = lds32<64 * S>(a0), w1 = lds32<64 * S + 4>(a0);
    uint32_t v = __funnelshift_r(w0, w1, sh);
    const uint32_t m = low_bytes_mask(left0 - 64 * S);                    // bytes of the line in this lane's word
    hib |= v & m;
    if (C::ROW0) lob &= v | (v << 1) | ~m;
    v = mask_to_trash(v, m);
    // a lane beyond the last chunk of the table (PPAD % 64 != 0: it never has a byte of the line) bumps the spare
    // row of the last chunk instead
    constexpr int CO = 4 * C::CHUNK_WORDS * 2 * S;
    if (2 * S + 1 < C::NCHUNK) {
#pragma unroll
        for (int k = 0; k < 4; ++k) red_add<CO>(dp4a_u(v, sk.wsel[k], sk.hk[k]), sk.inc);
    } else {
        constexpr int CLAST = 4 * C::CHUNK_WORDS * (C::NCHUNK - 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) red_add<CLAST>(dp4a_u(v, sk.wsel[k], sk.hk[k] - sk.gofs), sk.inc);
    }
}
// (S known only at run time: the rare second step of a record whose two lines end in different steps)
template <class C>
__device__ __noinline__ void last_step_any(uint32_t S, uint32_t a0, uint32_t sh, int left0, uint32_t hk0, uint32_t hk1,
                                           uint32_t hk2, uint32_t hk3, uint32_t w0k, uint32_t w1k, uint32_t w2k,
                                           uint32_t w3k, uint32_t gofs, uint32_t inc, uint32_t* hib, uint32_t* lob)
{
    const uint32_t ao = a0 + 64u * S;
    uint32_t v = __funnelshift_r(lds32<0>(ao), lds32<4>(ao), sh);
    const uint32_t m = low_bytes_mask(left0 - 64 * (int)S);
    *hib |= v & m;
    if (C::ROW0) *lob &= v | (v << 1) | ~m;
    v = mask_to_trash(v, m);
    const uint32_t adj = min(2u * S * (uint32_t)(4 * C::CHUNK_WORDS) + gofs, (uint32_t)((C::NCHUNK - 1) * 4 * C::CHUNK_WORDS)) - gofs;
    red_add<0>(dp4a_u(v, w0k, hk0 + adj), inc);
    red_add<0>(dp4a_u(v, w1k, hk1 + adj), inc);
    red_add<0>(dp4a_u(v, w2k, hk2 + adj), inc);
    red_add<0>(dp4a_u(v, w3k, hk3 + adj), inc);
}

// ps / pq: (window offset of the line) | (bytes of it with a counter) << 16, of ONE record, warp-uniform
template <class C>
__device__ __forceinline__ void line_steps(uint32_t ps, uint32_t pq, bool qhalf, const StepK& sk, uint32_t& hib,
                                           uint32_t& lob)
{
    constexpr int NSTEP = (C::PPAD + 63) / 64;
    static_assert(NSTEP <= 6, "line_steps: at most 6 steps of 64 positions");
    const uint32_t ns = ps >> 16, nq = pq >> 16;
    const uint32_t pk = qhalf ? pq : ps;
    const uint32_t addr = sk.lane_base + (pk & 0xFFFFu);                  // shared address of this lane's word in step 0
    const uint32_t a0 = addr & ~3u, sh = addr << 3;
    const int left0 = (int)(pk >> 16) - (int)sk.lane_pos;
    const uint32_t nmax = max(ns, nq);
    // steps 0 .. nfull - 1 lie inside both lines; the lines end in step nfull (nothing of them is left there when
    // the length is a multiple of 64: that step then only feeds the spare row).  One computed jump per record:
    // case k = the guarded step k, then the unguarded steps k - 1 .. 0 as straight-line code.
    const uint32_t nfull = min(min(ns, nq) >> 6, (uint32_t)(NSTEP - 1));  // (warp-uniform)
    switch (nfull) {
    case 5: if (NSTEP > 5) { last_step<C, 5>(a0, sh, left0, sk, hib, lob); goto f4; }
    case 4: if (NSTEP > 4) { last_step<C, 4>(a0, sh, left0, sk, hib, lob); goto f3; }
    case 3: if (NSTEP > 3) { last_step<C, 3>(a0, sh, left0, sk, hib, lob); goto f2; }
    case 2: if (NSTEP > 2) { last_step<C, 2>(a0, sh, left0, sk, hib, lob); goto f1; }
    case 1: if (NSTEP > 1) { last_step<C, 1>(a0, sh, left0, sk, hib, lob); goto f0; }
    default: last_step<C, 0>(a0, sh, left0, sk, hib, lob); goto done;
    }
f4: if (NSTEP > 5) full_step<C, 4>(a0, sh, sk, hib, lob);
f3: if (NSTEP > 4) full_step<C, 3>(a0, sh, sk, hib, lob);
f2: if (NSTEP > 3) full_step<C, 2>(a0, sh, sk, hib, lob);
f1: if (NSTEP > 2) full_step<C, 1>(a0, sh, sk, hib, lob);
f0: if (NSTEP > 1) full_step<C, 0>(a0, sh, sk, hib, lob);
done:
    // seq() and qual() end in different steps ('\r' trimmed from one of them only, at a multiple of 64): rare
    if (nmax > 64u * (nfull + 1u))
        last_step_any<C>(nfull + 1u, a0, sh, left0, sk.hk[0], sk.hk[1], sk.hk[2], sk.hk[3], sk.wsel[0], sk.wsel[1], sk.wsel[2],
                         sk.wsel[3], sk.gofs, sk.inc, &hib, &lob);
}

struct RecSink {   // where validate_block accounts the records of a scanned window
    unsigned long long* stats;
    unsigned long long* seqraw;
    uint32_t* lenh;
    uint32_t max_len;
};

template <class C, bool HIST>
__device__ __forceinline__ uint32_t validate_block(const RecSink& p, const uint8_t* buf, uint32_t buf_s, uint32_t Pm,
                                                   uint32_t n_rec, uint32_t base, int lane, uint32_t& ps, uint32_t& pq,
                                                   WinAcc& wa)
{
    uint32_t* const lenh = p.lenh;
    const uint32_t r = base + (uint32_t)lane;
    const bool valid = r < n_rec;
    // lanes without a record look at record 0 of the window (real data, harmless) and are masked out below
    const uint32_t lp = buf_s + (uint32_t)(C::WIN + 16) + (valid ? 8u * r : 0u);
    // (the bound keeps every address derived from the list inside this warp's buffer even if a byte >= 0x80
    // in another warp's sequence line made a bump of that warp land in this list -- the launch is void then)
    constexpr uint32_t LIM = (uint32_t)C::WIN + 15u;
    const uint32_t s = min(lds_u16<0>(lp), LIM), h = min(lds_u16<2>(lp) - 1u, LIM), q = min(lds_u16<4>(lp) - 1u, LIM),
                   pp = min(lds_u16<6>(lp) - 1u, LIM), e = min(lds_u16<8>(lp) - 1u, LIM);
    const uint32_t c_at = lds_u8(buf_s + s), c_plus = lds_u8(buf_s + q + 1u), c_sr = lds_u8(buf_s + q - 1u),
                   c_qr = lds_u8(buf_s + e - 1u);
    // src/records.rs:137-149 ('@'), :151-163 ('+'), :233-238 (raw line lengths equal)
    const bool good = c_at == '@' && c_plus == '+' && (e - pp) == (q - h);
    const unsigned nok = __ballot_sync(0xffffffffu, valid && !good);
    const uint32_t first_bad = nok ? base + (uint32_t)__ffs(nok) - 1u : NO_START;
    const bool ok = valid && r < first_bad;
    ps = 0;
    pq = 0;
    if (HIST) {
        if (ok) {
            const uint32_t P = p.max_len;
            const uint32_t Lr = q - h - 1u;
            // seq()/qual() drop one trailing '\r' (src/records.rs:65-73,82-90)
            const uint32_t Ls = Lr - ((Lr > 0 && c_sr == '\r') ? 1u : 0u);
            const uint32_t Lq = Lr - ((Lr > 0 && c_qr == '\r') ? 1u : 0u);
            wa.n_records++;
            wa.n_bases += Ls;
            if (max(Ls, Lq) > P) {                                        // longer than the tracked positions: rare
                if (Ls > P) atomicAdd(p.stats + 2, (unsigned long long)(Ls - P));
                if (Lq > P) atomicAdd(p.stats + 3, (unsigned long long)(Lq - P));
            }
            const uint32_t lb = Ls <= P ? Ls : P + 1u;
            if (lb < (uint32_t)C::PPAD + 2u)
                atomicAdd(lenh + lb, 1u);
            else
                atomicAdd(p.stats + stats_len_off(P) + lb, 1ull);
            ps = (h + 1u) | (min(Ls, Pm) << 16);
            pq = (pp + 1u) | (min(Lq, Pm) << 16);
            // positions beyond the shared-memory columns but below P: straight to global (P > PPAD only)
            if (P > Pm) {
                const uint32_t gs = min(Ls, P), gq = min(Lq, P);
                unsigned long long* gqual = p.stats + stats_qual_off(P);
                for (uint32_t g = Pm; g < gs; ++g) atomicAdd(p.seqraw + (size_t)g * 256 + buf[h + 1u + g], 1ull);
                for (uint32_t g = Pm; g < gq; ++g) atomicAdd(gqual + (size_t)g * 256 + buf[pp + 1u + g], 1ull);
            }
        }
    } else if (ok) {
        wa.n_records++;
    }
    return first_bad;
}

// ------------------------------------------------------------------------------------------
// PREDICTED windows.  While the records keep the shape of the last record a scan delimited (line
// lengths Lh, Lsq, Lp, Lsq with their '\n'; '\r' before the '\n' of the sequence / quality line or
// not), a window is not scanned at all: record r starts at pad + r * reclen and is VERIFIED instead:
//   lane i of its 8 lanes checks one byte: '@', the four '\n', '+', and the two bytes that decide
//   the '\r' trimming; all lanes check that header and separator line hold no other '\n';
//   that the sequence and quality lines hold no '\n' is checked by the histogram itself -- every
//   byte of them is counted (Lsq - 1 <= positions with a counter), and a count in row '\n' raises
//   spec_fail when the counters are drained (flush_hist).
// A record that does not verify ends the prediction: it and everything behind it is left to the
// next window, which scans.  Per record this costs ~1/4 of the scan and no list.
// ------------------------------------------------------------------------------------------
struct Shape {
    uint32_t Lh, Lsq, Lp;     // header / sequence (= quality) / separator line length, each with its '\n'
    uint32_t cr_s, cr_q;      // 1 if the byte before the '\n' of the sequence / quality line is '\r'
    uint32_t reclen;
};

template <class C, int NR>
__device__ __forceinline__ uint32_t pred_pass(uint32_t buf_s, const Shape& sh, uint32_t pad, uint32_t chk_off,
                                              uint32_t chk_exp, uint32_t chk_neg, uint32_t gm,
                                              const LaneK& lc, uint32_t hist_s, uint32_t n_rec, uint32_t pass,
                                              uint32_t sub, uint32_t i, uint32_t kA, uint32_t kB, uint32_t& hib)
{
    const uint32_t r = 4u * pass + sub;
    const bool valid = r < n_rec;
    const uint32_t s = buf_s + pad + (valid ? r : 0u) * sh.reclen;        // shared address of the record
    bool lane_ok = ((lds_u8(s + chk_off) == chk_exp) ? 1u : 0u) != chk_neg;
    lane_ok = lane_ok && no_newline32(s, 0u, min(sh.Lh - 1u, 32u), i, kA, kB);
    if (sh.Lh > 33u) lane_ok = lane_ok && no_newline32(s, 32u, sh.Lh - 1u, i, kA, kB);
    if (sh.Lp > 2u) lane_ok = lane_ok && no_newline32(s, sh.Lh + sh.Lsq + 1u, sh.Lh + sh.Lsq + sh.Lp - 1u, i, kA, kB);
    const unsigned nok = __ballot_sync(0xffffffffu, valid && !lane_ok);
    uint32_t first_bad = NO_START;
    bool ok = valid;
    if (nok) {
        const uint32_t fsub = ((uint32_t)__ffs(nok) - 1u) >> 3;             // first group with a lane that objects
        first_bad = 4u * pass + fsub;
        ok = valid && sub < fsub;
    }
    const uint32_t sa = s + sh.Lh + 4u * i;                    // position 4i of the sequence line
    const uint32_t qa = sa + sh.Lsq + sh.Lp;                   // ... of the quality line
    // lanes without a (holding) record do not hold the fast rounds back: their bumps add zero
    // hib: OR of every word the rounds look at.  A byte >= 0x80 makes the dp4a address leave its row --
    // still inside this CTA's shared memory (at most 32 KB above the table: the length histogram and
    // window buffers), so nothing faults; the caller raises spec_fail and the exact path redoes the shard.
    // (the line length is the same for every record of the window: the guards of the last, partial round do
    // not depend on the pass, only the increments do)
    frounds<C, NR>(sa, qa, ok ? 1u : 0u, ok ? 0x10000u : 0u, ok ? gm : 0u, lc, hib);
    return first_bad;
}

// The rounds for records whose HEADER length varies (instrument coordinates in the id line).  The caller has scanned
// the window (all its '\n' are in the list) and checked every record against the shape, one lane per record: lane r
// holds the window offset of the sequence line of record r; nothing is left to verify here.
template <class C, int NR>
__device__ __forceinline__ uint32_t flex_pass(uint32_t buf_s, const Shape& sh, uint32_t my_body, uint32_t gm, const LaneK& lc,
                                              uint32_t n_rec, uint32_t pass, uint32_t sub, uint32_t i, uint32_t& hib)
{
    const uint32_t r = 4u * pass + sub;
    const bool ok = r < n_rec;
    const uint32_t sa = buf_s + __shfl_sync(0xffffffffu, my_body, ok ? r : 0u) + 4u * i;
    const uint32_t qa = sa + sh.Lsq + sh.Lp;
    frounds<C, NR>(sa, qa, ok ? 1u : 0u, ok ? 0x10000u : 0u, ok ? gm : 0u, lc, hib);
    return NO_START;
}

// All passes of a predicted window, dispatched once on the number of rounds NR = ceil(n / 32) of its shape
template <class C, bool FLEX, int NR>
struct PredWindow {
    static __device__ __forceinline__ uint32_t run(uint32_t nr, uint32_t buf_s, const Shape& sh, uint32_t pad,
                                                   uint32_t my_start, uint32_t my_lh, uint32_t chk_off, uint32_t chk_exp,
                                                   uint32_t chk_neg, uint32_t gm, const LaneK& lc, uint32_t hist_s,
                                                   uint32_t n_rec, uint32_t sub, uint32_t i, uint32_t kA, uint32_t kB,
                                                   uint32_t& hib)
    {
        if (nr != (uint32_t)NR)
            return PredWindow<C, FLEX, NR - 1>::run(nr, buf_s, sh, pad, my_start, my_lh, chk_off, chk_exp, chk_neg, gm, lc,
                                                    hist_s, n_rec, sub, i, kA, kB, hib);
        uint32_t first_bad = NO_START;
        for (uint32_t pass = 0; 4u * pass < n_rec && first_bad == NO_START; ++pass)
            first_bad = FLEX ? flex_pass<C, NR>(buf_s, sh, my_start, gm, lc, n_rec, pass, sub, i, hib)
                             : pred_pass<C, NR>(buf_s, sh, pad, chk_off, chk_exp, chk_neg, gm, lc, hist_s, n_rec, pass, sub,
                                                i, kA, kB, hib);
        return first_bad;
    }
};
template <class C, bool FLEX>
struct PredWindow<C, FLEX, -1> {
    static __device__ __forceinline__ uint32_t run(uint32_t, uint32_t, const Shape&, uint32_t, uint32_t, uint32_t, uint32_t,
                                                   uint32_t, uint32_t, uint32_t, const LaneK&, uint32_t, uint32_t, uint32_t,
                                                   uint32_t, uint32_t, uint32_t, uint32_t&)
    {
        return 0u;   // (not reached: nr <= NCHUNK by the predict condition)
    }
};

// '\n' count of a full window at positions >= pad (no list, no ranks: ~1/3 of the scan)
template <class C>
__device__ __forceinline__ uint32_t win_count_newlines(uint32_t buf_s, uint32_t pad, int lane, uint32_t kA, uint32_t kB)
{
    uint32_t cnt = 0;
#pragma unroll
    for (int it = 0; it < C::NU; ++it) {
        const uint4 v = lds_v4(buf_s + (uint32_t)(it * UNIT + lane * 16));
        uint32_t m0 = nlbits3(v.x, kA, kB), m1 = nlbits3(v.y, kA, kB), m2 = nlbits3(v.z, kA, kB), m3 = nlbits3(v.w, kA, kB);
        if (it == 0 && lane == 0) {                                       // bytes before the cursor
            const uint32_t pw = pad >> 2, pb = (pad & 3u) * 8u;           // whole words / bytes of the next word to drop
            const uint32_t part = ~((1u << pb) - 1u);
            m0 = pw > 0 ? 0u : m0 & part;
            m1 = pw > 1 ? 0u : (pw == 1 ? m1 & part : m1);
            m2 = pw > 2 ? 0u : (pw == 2 ? m2 & part : m2);
            m3 = pw == 3 ? m3 & part : m3;
        }
        // the flags sit at bit 7 of every byte: one popc over the four words, shifted apart
        cnt += (uint32_t)__popc(m0 | (m1 >> 1) | (m2 >> 2) | (m3 >> 3));
    }
    return __reduce_add_sync(0xffffffffu, cnt);
}

// one window descriptor of the warp's range (see ScanParams::desc): lanes 0..3 hold one word each (w01: lane 0 / 1,
// w23: lanes 2 / 3 -- for predicted windows a per-shape constant); false if the range's share is full
struct DescOut {
    uint32_t* ptr;      // this lane's word of the next descriptor
    uint32_t left;      // descriptors the range may still write
};
__device__ __forceinline__ bool desc_put(DescOut& d, int lane, uint32_t w0, uint32_t w1, uint32_t w23)
{
    if (d.left == 0) return false;
    if (lane < 4) *d.ptr = lane == 0 ? w0 : lane == 1 ? w1 : w23;
    d.ptr += 4;
    --d.left;
    return true;
}

struct StreamCta {
    uint32_t n_records, n_bases;   // totals of the CTA (u32: a CTA sees < 4 G bases per launch; native shared atomics)
    uint32_t recs;          // records consumed by the CTA (drives the drain of the u16 counter halves)
    uint32_t flush_epoch;   // bumped by the warp that pushes `recs` over a multiple of DRAIN_MARK
    uint32_t orphans;       // bit w: warp w is not (or no longer) inside its range loop -- it cannot drain its slice
};
constexpr uint32_t DRAIN_MARK = 24000u;   // records a CTA consumes between two drains of its u16 counter halves

// u16 counter halves: the warp that pushes the CTA-wide record count over a multiple of the mark starts a drain
// epoch.  Every warp still inside its range loop drains ITS slice of the table when it notices (at the end of
// its current window) -- lock-free: atomicExch leaves the other warps' concurrent bumps intact.  The slices of
// warps that have left their loop (or never had a range) are ORPHANS: the warp that starts the epoch drains
// those as well, there and then.  So every counter is drained within one window's time of every epoch,
// whatever the other warps are doing: between two drains of a counter the CTA consumes DRAIN_MARK records plus
// the few hundred of one window per warp -- far below 65 535 per half, for any input.
template <class C>
__device__ __forceinline__ void drain_tick(uint32_t* hist, const ScanParams& p, StreamCta& cta, uint32_t n_rec,
                                           uint32_t& my_epoch, int warp, int lane)
{
    constexpr int SLICE = (C::HIST_WORDS + C::NWARPS - 1) / C::NWARPS;
    uint32_t trip = 0;
    if (lane == 0) {
        const uint32_t before = atomicAdd(&cta.recs, n_rec);
        if (before / DRAIN_MARK != (before + n_rec) / DRAIN_MARK) {
            atomicAdd(&cta.flush_epoch, 1u);
            trip = 1;
        }
    }
    const uint32_t ep = *reinterpret_cast<volatile uint32_t*>(&cta.flush_epoch);
    if (ep != my_epoch) {
        my_epoch = ep;
        flush_hist<C>(hist, p, warp * SLICE, min((warp + 1) * SLICE, C::HIST_WORDS), lane, 32);
    }
    if (__shfl_sync(0xffffffffu, trip, 0)) {
        uint32_t m = *reinterpret_cast<volatile uint32_t*>(&cta.orphans);
        while (m) {
            const int w = __ffs(m) - 1;
            m &= m - 1u;
            flush_hist<C>(hist, p, w * SLICE, min((w + 1) * SLICE, C::HIST_WORDS), lane, 32);
        }
    }
}
// a warp leaves its range loop (or has no range): its slice is an orphan from now on; one last drain of its own
template <class C>
__device__ __forceinline__ void drain_leave(uint32_t* hist, const ScanParams& p, StreamCta& cta, int warp, int lane)
{
    constexpr int SLICE = (C::HIST_WORDS + C::NWARPS - 1) / C::NWARPS;
    if (lane == 0) atomicOr(&cta.orphans, 1u << warp);
    __syncwarp();
    flush_hist<C>(hist, p, warp * SLICE, min((warp + 1) * SLICE, C::HIST_WORDS), lane, 32);
}

// ------------------------------------------------------------------------------------------
// VARIABLE-shape stretch of a range (reads of varying length: nothing to predict).  Every window is scanned;
// its records are validated 32 at a time, one per lane, and counted one record at a time by the whole warp
// (validate_block / line_steps).  A separate loop -- not a branch of the loop over predicted windows -- so
// that neither weighs on the other's registers (the two live in different instantiations of the kernel).
// ------------------------------------------------------------------------------------------
struct RangeState {          // what the two loops over a range hand to each other
    unsigned long long cur;      // cursor: buffer offset of the next record
    unsigned long long lrank;    // line ends of the range staged so far
    unsigned long long tail_x;   // offset of the bad / incomplete record that ends an EOF shard (NONE64: none)
    uint32_t parity;             // phase of the warp's mbarrier
    uint32_t epoch;              // last drain epoch this warp has served
    uint32_t dbg_scan;
    bool failed;
    bool open_tail;              // the bytes of the shard end inside the record at `cur`, and more of the stream follows
};

template <class C, bool HIST>
__device__ __forceinline__ void var_loop(const ScanParams& p, uint8_t* buf, uint32_t buf_s, uint16_t* list,
                                         unsigned long long* bar, uint32_t* hist, uint32_t* lenh, uint32_t hist_s,
                                         StreamCta& cta, uint32_t rid, unsigned long long R1, bool last_eof,
                                         bool want_index, RangeState& rs, int warp, int lane, uint32_t lt_mask)
{
    const bool last_open = rid + 1u == p.n_sranges && !(p.flags & F_EOF) && p.n_avail == p.n_own;
    const uint32_t Pm = p.max_len < (uint32_t)C::PPAD ? p.max_len : (uint32_t)C::PPAD;
    const RecSink sink = {p.stats, p.seqraw, lenh, p.max_len};
    const bool qhalf = lane >= 16;
    StepK sk = {};
    if (HIST) {
        const uint32_t sub = (uint32_t)lane >> 3, i = (uint32_t)lane & 7u;
        sk.lane_pos = 4u * ((uint32_t)lane & 15u);
        sk.lane_base = buf_s + sk.lane_pos;
        sk.inc = qhalf ? 0x10000u : 1u;
        sk.gofs = (sub & 1u) * (uint32_t)(4 * C::CHUNK_WORDS);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const uint32_t bytek = ((uint32_t)kk + sub) & 3u;
            sk.hk[kk] = hist_s - (uint32_t)(C::ROW0 * 128) + 4u * (4u * i + bytek) + sk.gofs;   // (row of byte b: b - ROW0)
            sk.wsel[kk] = 128u << (8u * bytek);
            asm volatile("" : "+r"(sk.hk[kk]), "+r"(sk.wsel[kk]));
        }
        // (kept in registers: recomputing them for every record costs more than it saves)
        asm volatile("" : "+r"(sk.lane_pos), "+r"(sk.lane_base), "+r"(sk.inc), "+r"(sk.gofs));
    }
    while (!rs.failed && rs.cur < R1 && rs.cur < p.n_avail) {
        const Window w = win_load<C>(p, buf, bar, rs.parity, (long long)rs.cur, lane);
        const unsigned long long room = R1 - w.src;          // > pad: window bytes inside the range
        ++rs.dbg_scan;
        uint32_t hib_unused;
        const uint32_t total = win_scan<C, false, true>(buf_s, list, w, hib_unused, lane, lt_mask);
        const uint32_t n_win = min(total / 4u, (uint32_t)C::MAXR);   // complete records in the window
        if (n_win == 0 && last_eof && w.vlen < (uint32_t)C::WIN) {
            rs.tail_x = rs.cur;   // the stream ends inside this record (or in garbage): the first bad record
            break;
        }
        if (n_win == 0 && last_open && w.vlen < (uint32_t)C::WIN) {
            rs.open_tail = true;  // the bytes END inside this record and more will follow (a refill): carried over
            break;
        }
        if (n_win == 0) {
            rs.failed = true;     // a record longer than the window, data ending inside a record
            break;
        }
        // records of the window that start inside the range (their starts increase)
        uint32_t n_rec = n_win;
        if (room < (unsigned long long)C::WIN) {              // the range ends inside this window
            const uint32_t ra = (uint32_t)lane, rb = (uint32_t)lane + 32u;
            const bool va = ra < n_win && list[4u * ra] < (uint32_t)room;
            const bool vb = rb < n_win && list[4u * min(rb, (uint32_t)C::MAXR)] < (uint32_t)room;
            n_rec = (uint32_t)__popc(__ballot_sync(0xffffffffu, va)) + (uint32_t)__popc(__ballot_sync(0xffffffffu, vb));
        }
        WinAcc wa = {0, 0};
        uint32_t first_bad = NO_START, hib = 0, lob = 0xFFFFFFFFu;
        for (uint32_t b0 = 0; b0 < n_rec && first_bad == NO_START; b0 += 32u) {
            uint32_t ps, pq;
            first_bad = validate_block<C, HIST>(sink, buf, buf_s, Pm, n_rec, b0, lane, ps, pq, wa);
            if (HIST) {
                const uint32_t cnt = min(min(n_rec, first_bad) - b0, 32u);
                for (uint32_t r = 0; r < cnt; ++r)
                    line_steps<C>(__shfl_sync(0xffffffffu, ps, r), __shfl_sync(0xffffffffu, pq, r), qhalf, sk, hib, lob);
            }
        }
        // bytes >= 0x80 (or below the first row of the table) in a sequence or quality line -- id and separator
        // lines may hold anything --: their bumps left the table rows, the exact path redoes the shard
        if (HIST && C::ROW0) hib |= ~lob << 1;                 // bit 7 of a byte lane: some byte there was < 32
        if (HIST && __any_sync(0xffffffffu, (hib & 0x80808080u) != 0)) {
            rs.failed = true;
            break;
        }
        if (first_bad != NO_START) {
            if (!last_eof && !(p.flags & F_CAN_RETRY)) {
                rs.failed = true;   // a record that fails validation: the exact path finds and classifies it
                break;
            }
            n_rec = first_bad;      // the records in front of it stand; the range ends at the bad one
            rs.tail_x = (unsigned long long)(w.src + (long long)list[4u * first_bad]);
        }
        {
            const uint32_t nr = __reduce_add_sync(0xffffffffu, wa.n_records), nb = __reduce_add_sync(0xffffffffu, wa.n_bases);
            if (lane == 0) {
                atomicAdd(&cta.n_records, nr);
                atomicAdd(&cta.n_bases, nb);
            }
        }
        // line ends of the consumed records that lie in the owned bytes of the shard
        uint32_t n_lines = 4u * n_rec;
        const uint32_t next = list[n_lines];                  // start of the first record not consumed
        if (next <= w.pad && rs.tail_x == NONE64) {           // (only a list another warp's stray bump damaged)
            rs.failed = true;
            break;
        }
        if (w.src + next - 1u >= p.n_own) {                   // the shard's last record reaches beyond n_own
            const bool in = lane < 4 && w.src + list[n_lines - 3u + (uint32_t)lane] - 1u < p.n_own;
            n_lines = n_lines - 4u + (uint32_t)__popc(__ballot_sync(0xffffffffu, in));
        }
        if (want_index) {
            if (rs.lrank + n_lines > p.stage_share) {
                rs.failed = true;   // staging share too small: the exact path writes the index
                break;
            }
            const uint32_t off = (uint32_t)(p.stream_offset + w.src) - 1u;   // low 32 bits are what the index holds
            uint32_t* out = p.index_stage + (size_t)rid * p.stage_share + rs.lrank;
            const uint32_t ls = buf_s + (uint32_t)(C::WIN + 16) + 2u;
            for (uint32_t j = lane; j < n_lines; j += 32) out[j] = off + lds_u16<0>(ls + 2u * j);
        }
        rs.lrank += n_lines;
        rs.cur = w.src + next;
        if (rs.tail_x != NONE64) break;
        if (HIST) drain_tick<C>(hist, p, cta, n_rec, rs.epoch, warp, lane);   // u16 counter halves
    }
}

// HIST: the launch accumulates the per-position histograms (FQB_F_HIST); the two variants share no
// hot code (rounds + '\n'-row check vs. newline count), so each is compiled without the other's registers
// VAR: the variant for reads of varying length (every window scanned, records counted by line_steps); the
// other variant predicts.  Both are launched; fq_init_kernel's look at the head of the shard (res->shape_var)
// decides which of them does the work -- either one is correct on any input, they differ in what they are fast on.
template <class C, bool HIST, bool VAR>
__global__ void __launch_bounds__(C::NTHREADS, 1) fq_stream_kernel(const __grid_constant__ ScanParams p)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + C::PRE);
    uint32_t* lenh = C::ROW0 ? reinterpret_cast<uint32_t*>(smem_raw) : hist + C::HIST_WORDS;
    __shared__ unsigned long long bars[C::NWARPS];
    __shared__ StreamCta cta;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;

    if (p.res->spec_fail) return;                              // the host sent this shard to the exact path
    if ((p.res->shape_var != 0) != VAR) return;                // the other variant's kind of input
    if ((p.flags & F_CARRY) && p.carry->status != 0) return;   // the stream already failed
    const unsigned long long line_base = (p.flags & F_CARRY) ? p.carry->line_base : p.line_base;

    for (int i = tid; i < C::BUF0 / 4; i += C::NTHREADS) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
    if (lane == 0) mbar_init(&bars[warp], 1);
    if (tid == 0) {
        cta.n_records = 0;
        cta.n_bases = 0;
        cta.recs = 0;
        cta.flush_epoch = 0;
        cta.orphans = 0;
    }
    fence_mbar_init();
    __syncthreads();

    uint8_t* buf = smem_raw + C::BUF0 + warp * C::WARP_BYTES;
    uint16_t* list = reinterpret_cast<uint16_t*>(buf + C::WIN + 16);
    const uint32_t buf_s = smem_u32(buf);
    unsigned long long* bar = &bars[warp];
    uint32_t parity = 0;

    const uint32_t hist_s = smem_u32(hist);
    uint32_t sub = (uint32_t)lane >> 3, li = (uint32_t)lane & 7u;
    asm volatile("" : "+r"(sub), "+r"(li));
    LaneK lc;
    {
        const uint32_t i = li;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const uint32_t bytek = ((uint32_t)kk + sub) & 3u;
            lc.hk[kk] = hist_s + 4u * (4u * i + bytek);
            lc.wsel[kk] = 128u << (8u * bytek);
            // keep them in registers: recomputing them inside every pass costs more than it saves
            asm volatile("" : "+r"(lc.hk[kk]), "+r"(lc.wsel[kk]));
        }
    }

    // the range of this warp: records that START in [R0, R1)
    const uint32_t rid = blockIdx.x * (uint32_t)C::NWARPS + (uint32_t)warp;
    const unsigned long long R0 = (unsigned long long)rid * p.srange_bytes;
    const bool live = rid < p.n_sranges;
    // (the last range takes the remainder: a range too short to hold a few records could not infer its start)
    const unsigned long long R1 = live ? (rid + 1u == p.n_sranges ? p.n_own : R0 + p.srange_bytes) : 0ull;
    const bool want_index = (p.flags & F_INDEX) && p.index != nullptr && p.index_cap != 0;

    unsigned long long cur = 0, lrank = 0;
    bool failed = false;
    uint32_t my_epoch = 0;
    // window descriptors of this range (the predicting variant; see ScanParams::desc)
    const uint32_t rid0 = blockIdx.x * (uint32_t)C::NWARPS + (uint32_t)warp;
    DescOut dout = {p.desc ? p.desc + ((size_t)rid0 * p.desc_cap) * 4 + (lane & 3) : nullptr, p.desc ? (uint32_t)p.desc_cap : 0u};

    if (live) {
        // ---- where the first record of the range starts -------------------------------------------
        {
            // the window starts one byte early where that byte exists: it may be the '\n' that makes
            // the range start itself a line start
            const bool front = rid != 0 || (p.flags & F_FRONT16);
            const Window w = win_load<C>(p, buf, bar, parity, (long long)R0 - (front ? 1 : 0), lane);
            uint32_t hib;
            const uint32_t total = win_scan<C, HIST && !VAR>(buf_s, list, w, hib, lane, lt_mask);
            const uint32_t nstored = min(total, (uint32_t)C::LIST_DUMMY - 1u);
            const uint32_t neg = (front && nstored >= 1u && list[1] == 16u) ? 1u : 0u;   // a '\n' right in front of the range
            uint32_t c;
            if (rid != 0 || (p.flags & F_INFER_START)) {
                c = infer_start(buf, list, nstored, (rid == 0 && (p.flags & F_LINE_START)) ? 0u : 1u, lane);
            } else {
                // the caller's line number decides: K owned '\n' lie in front of the first record
                const bool line_start = (p.flags & F_LINE_START) || neg;
                const uint32_t phase = (uint32_t)(line_base & 3ull);
                const uint32_t K = (line_start && phase == 0) ? 0u : 4u - phase;
                c = K + neg <= nstored ? K + neg : NO_START;
            }
            // (bytes >= 0x80 only matter to the histogram addressing: without histograms any byte will do; the
            // variable-length variant checks the sequence and quality lines themselves, see line_steps)
            if (c == NO_START || (HIST && !VAR && __any_sync(0xffffffffu, (hib & 0x80808080u) != 0))) {
                failed = true;
            } else {
                cur = (unsigned long long)(w.src + (long long)list[c]);
                if (rid == 0) {
                    // the K line ends in front of the first record belong to the shard all the same
                    const uint32_t K = c - neg;
                    if (want_index && (uint32_t)lane < K) {
                        if ((unsigned long long)lane < p.stage_share)
                            p.index_stage[lane] = (uint32_t)(p.stream_offset + (unsigned long long)(w.src + (long long)list[neg + 1u + lane] - 1));
                        else
                            failed = true;
                    }
                    failed = __any_sync(0xffffffffu, failed);
                    if (K && (unsigned long long)(w.src + (long long)list[c] - 1) >= p.n_own) failed = true;   // (a shard inside one record)
                    if (!VAR && want_index && K && !desc_put(dout, lane, K, 0u, 0u)) failed = true;
                    lrank = K;
                    if (lane == 0) p.res->line_phase = (int)((4u - K) & 3u);
                }
            }
        }
        if (lane == 0) p.sranges[rid].first = cur;

        // ---- stream through the range ------------------------------------------------------------
        const uint32_t Pm = p.max_len < (uint32_t)C::PPAD ? p.max_len : (uint32_t)C::PPAD;
        // shape of the last record delimited by a scan: while it keeps predicting the following
        // records, windows are not scanned at all (see pred_pass)
        Shape sh = {0, 0, 0, 0, 0, 1};
        // per-lane constants of the shape (set with it): the byte this lane verifies in every record,
        // the line end this lane writes to the index, and how many records fit a window
        uint32_t chk_off = 0, chk_exp = 0, chk_neg = 0, idx_le = 0, nfit_max = 0, nfit_rem = 0, gmask = 0, desc_w23 = 0;
        bool predict = false, flex = false;   // flex: header lengths vary, see flex_pass
        // a bad record in the LAST range of a shard that ends the stream needs no exact pass: every record in
        // front of it has been counted, nothing behind it has (tail_x = its offset)
        // (not with an inferred shard start: that protocol has no classification step behind this kernel)
        const bool last_eof = rid + 1u == p.n_sranges && (p.flags & F_EOF) && !(p.flags & F_INFER_START) && p.n_avail == p.n_own;
        unsigned long long tail_x = NONE64;
        // the shard's bytes end inside its last record and the stream goes on (FQB_F_PARTIAL refills, a shard without
        // halo): no error and no reason for the exact path -- the record is reported as the tail to carry over
        const bool last_open = rid + 1u == p.n_sranges && !(p.flags & F_EOF) && p.n_avail == p.n_own;
        bool open_tail = false;
        uint32_t strikes = 0, cooldown = 0;
        uint32_t dbg_pred = 0, dbg_scan = 0;
        uint32_t acc_rec = 0;   // records of predicted windows not yet in the CTA's totals
        const uint32_t kA = fq_kmask[0], kB = fq_kmask[1];
        if (VAR) {
            // ---- reads of varying length: every window is scanned (var_loop) ---------------------------
            RangeState rs = {cur, lrank, tail_x, parity, my_epoch, dbg_scan, failed, false};
            var_loop<C, HIST>(p, buf, buf_s, list, bar, hist, lenh, hist_s, cta, rid, R1, last_eof, want_index, rs, warp, lane,
                              lt_mask);
            cur = rs.cur;
            lrank = rs.lrank;
            tail_x = rs.tail_x;
            dbg_scan = rs.dbg_scan;
            failed = rs.failed;
            open_tail = rs.open_tail;
        }
        while (!VAR && !failed && cur < R1 && cur < p.n_avail) {
            const Window w = win_load<C>(p, buf, bar, parity, (long long)cur, lane);
            const unsigned long long room = R1 - w.src;          // > pad: window bytes inside the range
            uint32_t n_rec, n_lines, next;
            // ---- predicted window: full, inside the owned bytes, at least one record ------------------
            const uint32_t n_fit = nfit_max - (w.pad > nfit_rem ? 1u : 0u);   // = (WIN - pad) / reclen
            const bool full = w.vlen == (uint32_t)C::WIN && w.src + C::WIN <= (long long)p.n_own;
            const bool can_predict = predict && !flex && n_fit && full;
            if (can_predict && !HIST) {
                // ---- no histograms: the window holds n_fit predicted records and the head of the next.
                // All their predicted line ends (and '@', '+') are verified byte by byte, and the window
                // must hold exactly that many '\n': together that leaves no room for a stray one.
                const uint32_t E = w.pad + n_fit * sh.reclen;                 // start of the partial record
                const uint32_t o2 = sh.Lh + sh.Lsq;
                const uint32_t tail = (uint32_t)C::WIN - E;                   // its bytes in the window
                const uint32_t n_tail = (tail >= sh.Lh ? 1u : 0u) + (tail >= o2 ? 1u : 0u) + (tail >= o2 + sh.Lp ? 1u : 0u);
                const uint32_t total = win_count_newlines<C>(buf_s, w.pad, lane, kA, kB);
                bool good = total == 4u * n_fit + n_tail;
                const uint32_t n_chk = 6u * n_fit + n_tail;                   // '@', 4 x '\n', '+' per record; '\n' of the tail
                for (uint32_t j = lane; j < n_chk; j += 32) {
                    uint32_t rec = j / 6u, k = j - 6u * rec;
                    if (rec >= n_fit) {                                       // the line ends of the partial record
                        k = 1u + (j - 6u * n_fit) + ((j - 6u * n_fit) >= 2u ? 1u : 0u);   // -> k = 1, 2, 4
                        rec = n_fit;
                    }
                    const uint32_t off = k == 0 ? 0u : k == 1 ? sh.Lh - 1u : k == 2 ? o2 - 1u : k == 3 ? o2
                                       : k == 4 ? o2 + sh.Lp - 1u : sh.reclen - 1u;
                    const uint32_t want = k == 0 ? '@' : k == 3 ? '+' : '\n';
                    good = good && lds_u8(buf_s + w.pad + rec * sh.reclen + off) == want;
                }
                if (!__all_sync(0xffffffffu, good)) {
                    predict = false;                                          // scan this window instead
                    if (++strikes >= 2) cooldown = 32;
                    continue;
                }
                strikes = 0;
                ++dbg_pred;
                n_rec = n_fit;                                                // records that start inside the range
                if (room < (unsigned long long)C::WIN) n_rec = min(n_fit, ((uint32_t)room - w.pad + sh.reclen - 1u) / sh.reclen);
                acc_rec += n_rec;    // (into the CTA's totals when the shape changes or the range ends)
                n_lines = 4u * n_rec;
                next = w.pad + n_rec * sh.reclen;
                if (want_index) {
                    // the line ends of a predicted window are an arithmetic sequence: 16 bytes describe them
                    if (!desc_put(dout, lane, n_lines | (1u << 24), (uint32_t)(p.stream_offset + w.src) + w.pad, desc_w23)) {
                        failed = true;   // descriptor share too small: the exact path writes the index
                        break;
                    }
                }
            } else if (can_predict) {
                // records that start inside the range
                n_rec = n_fit;
                if (room < (unsigned long long)C::WIN) n_rec = min(n_fit, ((uint32_t)room - w.pad + sh.reclen - 1u) / sh.reclen);
                const uint32_t Lr = sh.Lsq - 1u;
                const uint32_t Ls = Lr - sh.cr_s;                           // seq()/qual() drop one trailing '\r' (cr_s == cr_q)
                uint32_t hib = 0;
                const uint32_t first_bad = PredWindow<C, false, C::NCHUNK>::run((Ls + 31u) >> 5, buf_s, sh, w.pad, 0u, 0u, chk_off,
                                                                                chk_exp, chk_neg, gmask, lc, hist_s, n_rec, sub,
                                                                                li, kA, kB, hib);
                if (__any_sync(0xffffffffu, (hib & 0x80808080u) != 0)) {
                    failed = true;   // bytes >= 0x80 reached the rounds: the exact path
                    break;
                }
                if (first_bad != NO_START) {
                    // the prediction stops holding at this record: consume what came before it, scan next time
                    n_rec = first_bad;
                    predict = false;
                    // (reads or headers of varying length: after two predictions that did not even get
                    // through half of their window, stop trying for a while)
                    if (2u * first_bad < n_fit && ++strikes >= 2) cooldown = 32;
                } else {
                    strikes = 0;
                }
                ++dbg_pred;
                if (n_rec == 0) continue;
                acc_rec += n_rec;    // records of the current shape: counted when it changes or the range ends
                n_lines = 4u * n_rec;
                next = w.pad + n_rec * sh.reclen;
                if (want_index) {
                    // the line ends of a predicted window are an arithmetic sequence: 16 bytes describe them
                    if (!desc_put(dout, lane, n_lines | (1u << 24), (uint32_t)(p.stream_offset + w.src) + w.pad, desc_w23)) {
                        failed = true;   // descriptor share too small: the exact path writes the index
                        break;
                    }
                }
            } else if (HIST && predict && flex && full) {
                // ---- predicted window, header lengths vary: the window is scanned (32 bytes per lane and step, the
                // scan of the variable-length variant) so that every '\n' is in the list; then one lane per record
                // checks it against the shape -- the three line lengths behind the header, '@', '+', the '\r' flags --
                // and the rounds run as for predicted records, from the sequence line the list gives
                uint32_t hib_unused;
                const uint32_t total = win_scan<C, false, true>(buf_s, list, w, hib_unused, lane, lt_mask);
                const uint32_t n_win = min(min(total / 4u, (uint32_t)C::MAXR), 32u);     // complete records looked at
                const uint32_t lim = room < (unsigned long long)C::WIN ? (uint32_t)room : (uint32_t)C::WIN;
                uint32_t my_body = 0;
                bool mine = false, good = false;
                {
                    const uint32_t j = 4u * min((uint32_t)lane, n_win ? n_win - 1u : 0u);
                    const uint32_t a = list[j], b = list[j + 1u], c = list[j + 2u], d = list[j + 3u], e = list[j + 4u];
                    mine = (uint32_t)lane < n_win && a < lim;                            // starts inside the range
                    good = c - b == sh.Lsq && d - c == sh.Lp && e - d == sh.Lsq && buf[a] == '@' && buf[c] == '+' &&
                           ((sh.Lsq > 1u && buf[c - 2u] == '\r') ? 1u : 0u) == sh.cr_s &&
                           ((sh.Lsq > 1u && buf[e - 2u] == '\r') ? 1u : 0u) == sh.cr_q;
                    my_body = b;
                }
                const uint32_t n_in = (uint32_t)__popc(__ballot_sync(0xffffffffu, mine));
                const unsigned nok = __ballot_sync(0xffffffffu, mine && !good);
                n_rec = nok ? min(n_in, (uint32_t)__ffs(nok) - 1u) : n_in;
                if (nok || n_win == 0) {
                    // the prediction stops holding (or the window holds no complete record): scan next time
                    predict = false;
                    if ((n_win == 0 || 2u * n_rec < n_in) && ++strikes >= 2) cooldown = 32;
                } else {
                    strikes = 0;
                }
                if (n_rec == 0) continue;
                const uint32_t Lr = sh.Lsq - 1u;
                const uint32_t Ls = Lr - sh.cr_s;
                uint32_t hib = 0;
                PredWindow<C, true, C::NCHUNK>::run((Ls + 31u) >> 5, buf_s, sh, 0u, my_body, 0u, 0u, 0u, 0u, gmask, lc, hist_s, n_rec,
                                                    sub, li, kA, kB, hib);
                if (__any_sync(0xffffffffu, (hib & 0x80808080u) != 0)) {
                    failed = true;
                    break;
                }
                ++dbg_pred;
                acc_rec += n_rec;    // records of the current shape: counted when it changes or the range ends
                n_lines = 4u * n_rec;
                next = list[n_lines];
                if (want_index) {
                    if (lrank + n_lines > p.stage_share || !desc_put(dout, lane, n_lines, (uint32_t)lrank, 0u)) {
                        failed = true;
                        break;
                    }
                    const uint32_t off = (uint32_t)(p.stream_offset + w.src) - 1u;       // low 32 bits are what the index holds
                    uint32_t* out = p.index_stage + (size_t)rid * p.stage_share + lrank;
                    const uint32_t ls = buf_s + (uint32_t)(C::WIN + 16) + 2u;
                    for (uint32_t j = lane; j < n_lines; j += 32) out[j] = off + lds_u16<0>(ls + 2u * j);
                }
            } else {
                // ---- scanned window -------------------------------------------------------------------
                ++dbg_scan;
                // the records predicted so far go into the CTA's totals before the shape they were counted with changes
                if (acc_rec) {
                    if (lane == 0) {
                        atomicAdd(&cta.n_records, acc_rec);
                        if (HIST) {
                            const uint32_t Ls = sh.Lsq - 1u - sh.cr_s;
                            atomicAdd(&cta.n_bases, acc_rec * Ls);
                            atomicAdd(lenh + Ls, acc_rec);
                        }
                    }
                    acc_rec = 0;
                }
                uint32_t hib;
                const uint32_t total = win_scan<C, HIST>(buf_s, list, w, hib, lane, lt_mask);
                const uint32_t n_win = min(total / 4u, (uint32_t)C::MAXR);   // complete records in the window
                if (n_win == 0 && last_eof && w.vlen < (uint32_t)C::WIN &&
                    !(HIST && __any_sync(0xffffffffu, (hib & 0x80808080u) != 0))) {
                    tail_x = cur;    // the stream ends inside this record (or in garbage): the first bad record
                    break;
                }
                if (n_win == 0 && last_open && w.vlen < (uint32_t)C::WIN) {
                    open_tail = true;   // the bytes END inside this record and more will follow (a refill): carried over
                    break;
                }
                if (n_win == 0 || (HIST && __any_sync(0xffffffffu, (hib & 0x80808080u) != 0))) {
                    failed = true;   // a record longer than the window, data ending inside a record, bytes >= 0x80
                    break;
                }
                // records of the window that start inside the range (their starts increase)
                n_rec = n_win;
                if (room < (unsigned long long)C::WIN) {              // the range ends inside this window
                    const uint32_t ra = (uint32_t)lane, rb = (uint32_t)lane + 32u;
                    const bool va = ra < n_win && list[4u * ra] < (uint32_t)room;
                    const bool vb = rb < n_win && list[4u * min(rb, (uint32_t)C::MAXR)] < (uint32_t)room;
                    n_rec = (uint32_t)__popc(__ballot_sync(0xffffffffu, va)) + (uint32_t)__popc(__ballot_sync(0xffffffffu, vb));
                }
                WinAcc wa = {0, 0};
                uint32_t first_bad = NO_START;
                for (uint32_t pass = 0; 4u * pass < n_rec && first_bad == NO_START; ++pass)
                    first_bad = stream_pass<C, HIST>(p, buf, buf_s, lc, hist_s, lenh, Pm, n_rec, pass, wa, sub, li);
                if (first_bad != NO_START) {
                    if (!last_eof && !(p.flags & F_CAN_RETRY)) {
                        failed = true;   // a record that fails validation: the exact path finds and classifies it
                        break;
                    }
                    n_rec = first_bad;   // the records in front of it stand; the range ends at the bad one
                    tail_x = (unsigned long long)(w.src + (long long)list[4u * first_bad]);
                }
                {
                    const uint32_t nr = __reduce_add_sync(0xffffffffu, wa.n_records), nb = __reduce_add_sync(0xffffffffu, wa.n_bases);
                    if (lane == 0) {
                        atomicAdd(&cta.n_records, nr);
                        atomicAdd(&cta.n_bases, nb);
                    }
                }
                // the shape the next windows are predicted with -- only while the histogram checks the
                // sequence and quality lines for stray '\n': every byte of them must have a counter
                {
                    const uint32_t l0 = list[0], l1 = list[1], l2 = list[2], l3 = list[3], l4 = list[4];
                    sh.Lh = l1 - l0;
                    sh.Lsq = l2 - l1;
                    sh.Lp = l3 - l2;
                    sh.reclen = l4 - l0;
                    sh.cr_s = (sh.Lsq > 1u && buf[l2 - 2u] == '\r') ? 1u : 0u;
                    sh.cr_q = (sh.Lsq > 1u && buf[l4 - 2u] == '\r') ? 1u : 0u;
                    if (cooldown) --cooldown;
                    // do the records of this window (the first 32) share the shape?  Reads of varying length are
                    // not predicted at all; headers of varying length are, with a search for every header end
                    bool all_sq, all_h;
                    {
                        const uint32_t j = 4u * min((uint32_t)lane, n_win - 1u);
                        const uint32_t a = list[j], b = list[j + 1u], c = list[j + 2u], d = list[j + 3u];
                        all_sq = __all_sync(0xffffffffu, c - b == sh.Lsq && d - c == sh.Lp);
                        all_h = __all_sync(0xffffffffu, b - a == sh.Lh);
                    }
                    // with histograms: every byte of the sequence / quality lines must have a counter (the
                    // '\n' row is what checks them) and header / separator are checked 32 bytes at a time;
                    // without: the '\n' count of the window checks every line (a per-record header search was
                    // measured there too: 3.2 TB/s against the 3.8 TB/s of simply scanning, so it is not used)
                    flex = HIST && (!all_h || sh.Lh > 64u);
                    predict = cooldown == 0 && sh.Lh >= 2u && sh.Lp >= 2u && all_sq &&
                              (HIST ? (sh.Lsq - 1u <= Pm && sh.Lp <= 34u && sh.cr_s == sh.cr_q) : all_h);
                    if (HIST && predict) {
                        // which bumps of the last, partial round lie inside seq() / qual() (see FRounds)
                        const uint32_t n = sh.Lsq - 1u - sh.cr_s, base = ((n - 1u) & ~31u) + 4u * li;   // (n == 0: no round)
                        gmask = 0;
#pragma unroll
                        for (uint32_t k = 0; k < 4; ++k)
                            if (base + ((k + sub) & 3u) < n) gmask |= 1u << (8u * k);
                    }
                    if (predict && flex) {
                        const uint32_t rest = 2u * sh.Lsq + sh.Lp;
                        // relative to the first byte of the sequence line (lane 0 looks at the record start)
                        chk_off = li == 1 ? 0u - 1u : li == 2 ? sh.Lsq - 1u : li == 3 ? sh.Lsq : li == 4 ? sh.Lsq + sh.Lp - 1u
                                : li == 5 ? rest - 1u : li == 6 ? sh.Lsq - 2u : rest - 2u;
                        chk_exp = li == 0 ? '@' : li == 3 ? '+' : li >= 6 ? '\r' : '\n';
                        chk_neg = (li == 6 && !sh.cr_s) || (li == 7 && !sh.cr_q) ? 1u : 0u;
                        const uint32_t k = (uint32_t)lane & 3u;
                        idx_le = k == 0 ? 0u - 1u : k == 1 ? sh.Lsq - 1u : k == 2 ? sh.Lsq + sh.Lp - 1u : rest - 1u;
                    } else if (predict) {
                        const uint32_t o2 = sh.Lh + sh.Lsq;
                        // the byte lane li of a record's 8 lanes verifies (relative to the record start)
                        chk_off = li == 0 ? 0u : li == 1 ? sh.Lh - 1u : li == 2 ? o2 - 1u : li == 3 ? o2
                                : li == 4 ? o2 + sh.Lp - 1u : li == 5 ? sh.reclen - 1u : li == 6 ? o2 - 2u : sh.reclen - 2u;
                        chk_exp = li == 0 ? '@' : li == 3 ? '+' : li >= 6 ? '\r' : '\n';
                        chk_neg = (li == 6 && !sh.cr_s) || (li == 7 && !sh.cr_q) ? 1u : 0u;
                        const uint32_t k = (uint32_t)lane & 3u;
                        idx_le = k == 0 ? sh.Lh - 1u : k == 1 ? o2 - 1u : k == 2 ? o2 + sh.Lp - 1u : sh.reclen - 1u;
                        nfit_max = (uint32_t)C::WIN / sh.reclen;
                        nfit_rem = (uint32_t)C::WIN - nfit_max * sh.reclen;
                        // this lane's constant word of the window descriptors of this shape (lanes 2 and 3 write them)
                        desc_w23 = (lane & 1) ? (o2 + sh.Lp - 1u) | (sh.reclen << 16) : (sh.Lh - 1u) | ((o2 - 1u) << 16);
                    }
                }
                // line ends of the consumed records that lie in the owned bytes of the shard
                n_lines = 4u * n_rec;
                next = list[n_lines];                             // start of the first record not consumed
                if (w.src + next - 1u >= p.n_own) {               // the shard's last record reaches beyond n_own
                    const bool in = lane < 4 && w.src + list[n_lines - 3u + (uint32_t)lane] - 1u < p.n_own;
                    n_lines = n_lines - 4u + (uint32_t)__popc(__ballot_sync(0xffffffffu, in));
                }
                if (want_index) {
                    if (lrank + n_lines > p.stage_share || !desc_put(dout, lane, n_lines, (uint32_t)lrank, 0u)) {
                        failed = true;   // staging share too small: the exact path writes the index
                        break;
                    }
                    const uint32_t off = (uint32_t)(p.stream_offset + w.src) - 1u;   // low 32 bits are what the index holds
                    uint32_t* out = p.index_stage + (size_t)rid * p.stage_share + lrank;
                    const uint32_t ls = buf_s + (uint32_t)(C::WIN + 16) + 2u;
                    for (uint32_t j = lane; j < n_lines; j += 32) out[j] = off + lds_u16<0>(ls + 2u * j);
                }
            }
            lrank += n_lines;
            cur = w.src + next;
            if (tail_x != NONE64) break;
            if (HIST) drain_tick<C>(hist, p, cta, n_rec, my_epoch, warp, lane);   // u16 counter halves
        }
        if (!VAR && acc_rec && lane == 0) {                       // (see the scanned branch)
            atomicAdd(&cta.n_records, acc_rec);
            if (HIST) {
                const uint32_t Ls = sh.Lsq - 1u - sh.cr_s;
                atomicAdd(&cta.n_bases, acc_rec * Ls);
                atomicAdd(lenh + Ls, acc_rec);
            }
        }
        bool stopped = false;
        if (tail_x != NONE64 && !failed) {
            if (last_eof) {
                // (the line ends of the bytes behind it are counted and indexed by fq_tail_index_kernel)
                if (lane == 0) {
                    atomicMin(&p.res->first_bad, tail_x);
                    p.res->tail_err = 1;
                }
                cur = p.n_avail;
            } else {
                // a bad record in the middle of the shard: this range ends in front of it; the ranges behind it have
                // counted records that each() never delivers -- the bytes in front of the smallest such offset are
                // parsed again once the chain of the ranges up to here has been verified (fq_stream_verify_kernel)
                if (lane == 0) atomicMin(&p.res->spec_bad, tail_x);
                cur = tail_x;
                stopped = true;
            }
        }
        if (open_tail && !failed && lane == 0) atomicMin(&p.res->tail_start, cur);
        // data that ends inside the owned bytes without a record boundary: not a clean shard
        if (!failed && !stopped && !open_tail && cur < R1) failed = true;
        if (lane == 0) {
            StreamRange& sr = p.sranges[rid];
            sr.end = cur;
            sr.n_lines = lrank;
            sr.flags = failed ? 2u : (stopped ? 3u : 1u);
            // (no window predicted: every line end of the range is staged, back to back -- one plain copy)
            sr.n_desc = (VAR || dbg_pred == 0) ? DESC_RAW_ONLY : (uint32_t)p.desc_cap - dout.left;
            atomicAdd(&p.res->n_win_pred, (unsigned long long)dbg_pred);
            atomicAdd(&p.res->n_win_scan, (unsigned long long)dbg_scan);
            if (failed) atomicExch(&p.res->spec_fail, 1);
        }
    }

    // ---- drain -----------------------------------------------------------------------------
    if (HIST) drain_leave<C>(hist, p, cta, warp, lane);       // (warps still at work drain this warp's slice from now on)
    __syncthreads();
    flush_hist<C>(hist, p, 0, C::HIST_WORDS, tid, C::NTHREADS);
    {
        unsigned long long* lenh_g = p.stats + stats_len_off(p.max_len);
        for (int i = tid; i < C::PPAD + 2; i += C::NTHREADS) {
            const uint32_t v = lenh[i];
            if (v) atomicAdd(lenh_g + i, (unsigned long long)v);
        }
    }
    if (tid == 0) {
        if (cta.n_records) atomicAdd(p.stats + 0, (unsigned long long)cta.n_records);
        if (cta.n_bases) atomicAdd(p.stats + 1, (unsigned long long)cta.n_bases);
    }
}

// ------------------------------------------------------------------------------------------
// verify: the record chain of every range must end exactly where the next range started, and the
// last one at (or, with more of the stream following, beyond) the end of the owned bytes.  Also the
// prefix of the per-range line counts = where each range's staged line ends go.  One 1024-thread block.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) fq_stream_verify_kernel(const ScanParams p, const DevCarry* carry)
{
    __shared__ unsigned long long part[1024];
    __shared__ unsigned int first_fail_s, first_stop_s;
    if (p.res->spec_fail) return;
    if (carry && carry->status != 0) return;
    const unsigned long long line_base = carry ? carry->line_base : p.line_base;
    const int t = threadIdx.x;
    const uint32_t nlive = p.n_sranges;
    const uint32_t per = (nlive + 1023u) / 1024u;
    if (t == 0) {
        first_fail_s = 0xFFFFFFFFu;
        first_stop_s = 0xFFFFFFFFu;
    }
    __syncthreads();
    unsigned long long sum = 0;
    // first range that did not deliver (or whose chain to its predecessor is broken), first range that stopped at
    // a bad record (F_CAN_RETRY): everything in front of the smaller of the two is verified
    uint32_t my_fail = 0xFFFFFFFFu, my_stop = 0xFFFFFFFFu;
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t r = (uint32_t)t * per + k;
        if (r >= nlive) break;
        const StreamRange sr = p.sranges[r];
        sum += sr.n_lines;
        if (sr.flags == 3u) my_stop = min(my_stop, r);
        if (sr.flags != 1u && sr.flags != 3u) my_fail = min(my_fail, r);
        if (sr.flags == 1u) {
            if (r + 1 < nlive) {
                if (p.sranges[r + 1].first != sr.end) my_fail = min(my_fail, r + 1u);
            } else if ((sr.end < p.n_own && sr.end != p.res->tail_start) || ((p.flags & F_EOF) && sr.end != p.n_avail)) {
                my_fail = min(my_fail, r);     // (tail_start: the incomplete last record of a shard that is carried over)
            }
        }
    }
    part[t] = sum;
    if (my_fail != 0xFFFFFFFFu) atomicMin(&first_fail_s, my_fail);
    if (my_stop != 0xFFFFFFFFu) atomicMin(&first_stop_s, my_stop);
    __syncthreads();
    const bool retry = first_stop_s != 0xFFFFFFFFu && first_stop_s < first_fail_s;   // (a stopped range's own start is
                                                                                    // covered by its predecessor's chain check)
    const bool fail_any = !retry && first_fail_s != 0xFFFFFFFFu;
    // exclusive prefix over the 1024 partial sums (Hillis-Steele)
    for (int d = 1; d < 1024; d <<= 1) {
        const unsigned long long v = t >= d ? part[t - d] : 0ull;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned long long run = part[t] - sum;
    for (uint32_t k = 0; k < per; ++k) {
        const uint32_t r = (uint32_t)t * per + k;
        if (r >= nlive) break;
        p.sranges[r].rank0 = run;
        run += p.sranges[r].n_lines;
    }
    if (t == 1023) {
        p.res->n_lines = part[1023];
        p.res->line_end = line_base + part[1023];
    }
    if (t == 0) {
        if (retry) {
            // every record in front of sranges[first_stop].end is what a sequential parse delivers, and the record
            // there fails validation: the host parses [0, that offset) again and classifies the record (fqb_fetch)
            p.res->spec_retry = 1;
            p.res->spec_bad = p.sranges[first_stop_s].end;
        } else if (fail_any) {
            p.res->spec_fail = 1;
        }
    }
}

// the line ends of every range to their place in the caller's index (only when the speculative launch stands;
// the exact kernel writes the index directly): staged runs are copied, window descriptors expanded
__global__ void __launch_bounds__(256) fq_stream_compact_kernel(const ScanParams p, const DevCarry* carry)
{
    __shared__ uint32_t warp_tot[8];
    __shared__ uint4 d_first;
    if (p.res->spec_fail) return;
    if (carry && carry->status != 0) return;
    const uint32_t nlive = p.n_sranges;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (uint32_t r = blockIdx.x; r < nlive; r += gridDim.x) {
        const StreamRange sr = p.sranges[r];
        const uint32_t* src = p.index_stage + (size_t)r * p.stage_share;
        const unsigned long long n = min(sr.n_lines, p.index_cap > sr.rank0 ? p.index_cap - sr.rank0 : 0ull);
        uint32_t* dst = p.index + sr.rank0;
        if (sr.n_desc == DESC_RAW_ONLY) {
            // (source and destination are shifted against each other by an arbitrary number of entries,
            // so this stays a 4-byte copy; four loads in flight per thread keep the memory system busy)
            unsigned long long i = t;
            for (; i + 3ull * blockDim.x < n; i += 4ull * blockDim.x) {
                const uint32_t a = __ldg(src + i), b = __ldg(src + i + blockDim.x), c = __ldg(src + i + 2 * blockDim.x),
                               d = __ldg(src + i + 3 * blockDim.x);
                dst[i] = a;
                dst[i + blockDim.x] = b;
                dst[i + 2 * blockDim.x] = c;
                dst[i + 3 * blockDim.x] = d;
            }
            for (; i < n; i += blockDim.x) dst[i] = __ldg(src + i);
            continue;
        }
        // window descriptors, 256 at a time: block prefix of their line counts, then every warp expands its 32
        const uint4* D = reinterpret_cast<const uint4*>(p.desc) + (size_t)r * p.desc_cap;
        unsigned long long base = 0;
        for (uint32_t c0 = 0; c0 < sr.n_desc; c0 += 256u) {
            const uint4 d = c0 + t < sr.n_desc ? __ldg(D + c0 + t) : make_uint4(0, 0, 0, 0);
            const uint32_t nl = d.x & 0xFFFFFFu;
            uint32_t incl = nl;
#pragma unroll
            for (int k = 1; k < 32; k <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, k);
                if (lane >= k) incl += v;
            }
            __syncthreads();                        // (warp_tot of the previous round has been read)
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            uint32_t before = 0, total = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                before += k < warp ? warp_tot[k] : 0u;
                total += warp_tot[k];
            }
            const unsigned long long my_off = base + before + incl - nl;
            const uint32_t k4 = (uint32_t)lane & 3u;
            // The usual case on fixed-length reads: the 256 windows continue one another -- same record shape, each
            // starting where the one before ended -- so the whole round is ONE arithmetic sequence and every thread
            // writes every 256th entry of it (coalesced, no shuffles).
            if (t == 0) d_first = d;
            __syncthreads();
            {
                const uint4 f = d_first;
                const uint32_t reclen = f.w >> 16;
                const bool pad = c0 + t >= sr.n_desc;
                const bool cont = pad || ((d.x >> 24) == 1u && d.z == f.z && d.w == f.w &&
                                          d.y == f.y + (uint32_t)((my_off - base) >> 2) * reclen);
                // ... or, on reads of varying length, 256 scanned windows whose staged line ends lie back to back
                const bool raw_run = pad || ((d.x >> 24) == 0u && d.y == f.y + (uint32_t)(my_off - base));
                const bool all_cont = __syncthreads_and(cont), all_raw = __syncthreads_and(raw_run);
                if (all_raw && (f.x >> 24) == 0u) {
                    const uint32_t* s0 = src + f.y;
                    for (uint32_t e = t; e < total; e += 256u)
                        if (base + e < n) dst[base + e] = __ldg(s0 + e);
                    base += total;
                    continue;
                }
                if (all_cont && (f.x >> 24) == 1u && ((reinterpret_cast<uintptr_t>(dst + base) & 15) == 0) && base + total <= n) {
                    // (the usual alignment: whole records in front) one 16-byte store per record
                    const uint32_t l0 = f.z & 0xFFFFu, l1 = f.z >> 16, l2 = f.w & 0xFFFFu, l3 = reclen - 1u;
                    uint4* d4 = reinterpret_cast<uint4*>(dst + base);
                    uint32_t v = f.y + (uint32_t)t * reclen;
                    for (uint32_t r = t; r < total / 4u; r += 256u, v += 256u * reclen)
                        d4[r] = make_uint4(v + l0, v + l1, v + l2, v + l3);
                    base += total;
                    continue;
                }
                if (all_cont && (f.x >> 24) == 1u) {
                    const uint32_t le = k4 == 0 ? (f.z & 0xFFFFu) : k4 == 1 ? (f.z >> 16) : k4 == 2 ? (f.w & 0xFFFFu) : reclen - 1u;
                    uint32_t v = f.y + ((uint32_t)t >> 2) * reclen + le;
                    for (uint32_t e = t; e < total; e += 256u, v += 64u * reclen)
                        if (base + e < n) dst[base + e] = v;
                    base += total;
                    continue;
                }
            }
            for (int i = 0; i < 32; ++i) {
                const uint32_t w0 = __shfl_sync(0xffffffffu, d.x, i);
                const uint32_t n_i = w0 & 0xFFFFFFu;
                if (n_i == 0) continue;
                const uint32_t w1 = __shfl_sync(0xffffffffu, d.y, i), w2 = __shfl_sync(0xffffffffu, d.z, i),
                               w3 = __shfl_sync(0xffffffffu, d.w, i);
                const unsigned long long o_i = __shfl_sync(0xffffffffu, my_off, i);
                if (w0 >> 24) {
                    // predicted records: entry j = first record + (j / 4) * record length + offset of line end j % 4
                    const uint32_t reclen = w3 >> 16;
                    const uint32_t le = k4 == 0 ? (w2 & 0xFFFFu) : k4 == 1 ? (w2 >> 16) : k4 == 2 ? (w3 & 0xFFFFu) : reclen - 1u;
                    uint32_t v = w1 + ((uint32_t)lane >> 2) * reclen + le;
                    for (uint32_t j = lane; j < n_i; j += 32, v += 8u * reclen)
                        if (o_i + j < n) dst[o_i + j] = v;
                } else {
                    for (uint32_t j = lane; j < n_i; j += 32)
                        if (o_i + j < n) dst[o_i + j] = __ldg(src + w1 + j);
                }
            }
            base += total;
        }
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
using SCfg5 = SCfg<5, 32, 4096>;     // P <= 160, no histograms: 32 warps x 4 KiB windows (bandwidth- and latency-bound: warps help)
using SCfg5H = SCfg<5, 28, 4096>;    // P <= 160 with histograms: 80 KB of counters, 28 warps x 4 KiB windows -- the rounds are
                                     // issue-bound and want registers: 72 per thread (no spills) beat 32 warps at 64 (+1 %)
using SCfg10 = SCfg<10, 22, 2560>;   // P <= 320: 160 KB of counters, 22 warps x 2.5 KiB windows (measured best of
                                     // 16 x 3584 / 22 x 2560 / 26 x 2048 on fixed 300 bp and on 50..300 bp reads)
// the variable-length variants: same ranges (the host cuts the shard into grid x NWARPS of them), a spare row per chunk
using VCfg5 = SCfg<5, 28, 4096, 1>;
using VCfg10 = SCfg<10, 22, 4096, 1, 32>;   // rows 32 .. 127 only: 124 KB of counters leave room for 4 KiB windows
static_assert(VCfg5::NWARPS == SCfg5H::NWARPS && VCfg10::NWARPS == SCfg10::NWARPS, "both variants walk the same ranges");

// (without histograms there is no counter table: 32 warps x 4 KiB windows whatever the maximum read length)
int stream_warps(int nchunk, bool hist) { return !hist ? SCfg5::NWARPS : nchunk <= 5 ? SCfg5H::NWARPS : SCfg10::NWARPS; }

template <class C, class H, class V>
static cudaError_t configure_set()
{
    cudaError_t e = cudaFuncSetAttribute(fq_stream_kernel<H, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, H::TOTAL);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fq_stream_kernel<V, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, V::TOTAL);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fq_stream_kernel<C, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(fq_stream_kernel<C, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL);
}

cudaError_t stream_configure()
{
    cudaError_t e = configure_set<SCfg5, SCfg5H, VCfg5>();
    if (e != cudaSuccess) return e;
    return configure_set<SCfg10, SCfg10, VCfg10>();
}

template <class C, class H, class V>
static void launch_set(const ScanParams& p, int grid, cudaStream_t st)
{
    if (p.flags & F_HIST) {
        // (one of the two returns at once: res->shape_var, set by fq_init_kernel's look at the head of the shard)
        fq_stream_kernel<H, true, false><<<grid, H::NTHREADS, H::TOTAL, st>>>(p);
        fq_stream_kernel<V, true, true><<<grid, V::NTHREADS, V::TOTAL, st>>>(p);
    } else {
        fq_stream_kernel<C, false, false><<<grid, C::NTHREADS, C::TOTAL, st>>>(p);
        fq_stream_kernel<C, false, true><<<grid, C::NTHREADS, C::TOTAL, st>>>(p);
    }
}

cudaError_t launch_stream(const ScanParams& p, int nchunk, int grid, cudaStream_t st)
{
    if (nchunk <= 5 || !(p.flags & F_HIST))
        launch_set<SCfg5, SCfg5H, VCfg5>(p, grid, st);
    else
        launch_set<SCfg10, SCfg10, VCfg10>(p, grid, st);
    return cudaGetLastError();
}

cudaError_t launch_stream_verify(const ScanParams& p, DevCarry* carry, cudaStream_t st)
{
    fq_stream_verify_kernel<<<1, 1024, 0, st>>>(p, carry);
    return cudaGetLastError();
}

cudaError_t launch_stream_compact(const ScanParams& p, DevCarry* carry, int grid, cudaStream_t st)
{
    // one block per range (they are equally long): no block is left with a range more than the others
    (void)grid;
    fq_stream_compact_kernel<<<p.n_sranges ? p.n_sranges : 1u, 256, 0, st>>>(p, carry);
    return cudaGetLastError();
}

}  // namespace fq
