// fq_scan.cu -- the fused sm_100a scan kernel (K1 delimit + K2 per-position histograms), v6.
//
// One pass over the bytes.  Persistent grid, one 1024-thread CTA per SM; CTA b owns a CONTIGUOUS
// range of tiles, so the stream line number of a tile is the line number of the range start plus a
// running count the CTA keeps itself -- no look-back, no coupling between CTAs while they run.
// The line number of a range start (mod 4 it decides which lines are headers) is not known before
// the ranges in front of it have been counted, so the first launch INFERS it: the CTA tests the four
// possible phases against the grammar over the first records of its range ('@', '+', equal raw
// lengths) and proceeds with the only one that holds.  fq_verify_kernel then compares every
// inferred phase with the exact prefix of the per-range newline counts (which do not depend on the
// phase); any mismatch or ambiguity makes the host-enqueued second launch redo the shard with the
// exact bases (F_BASES).  Results therefore never depend on the inference.
//
// The 32 warps of a CTA are SPECIALISED and hand tiles to each other through a ring of NSTAGE
// shared-memory stages guarded by mbarriers -- no CTA-wide barrier in the steady state:
//
//   TMA warp (1)       keeps the ring full: UBLKCP bulk copies completing on full[s]
//   scan warps (SW)    16-byte SWAR newline masks (3 ops / word + dp4a bit gather) -> per-unit counts
//                      -> ranks -> position list of the tile (u16, shared memory); scan warp 0 keeps
//                      the running line number of the range
//   record warps (HW)  8 lanes per record: '@' / '+' / raw-length validation
//                      (src/records.rs:201-247), then each lane walks 4-byte groups of the sequence
//                      and quality lines and bumps hist[chunk][byte][position % 32] with one dp4a
//                      (address = lane base + byte * 128) and one ATOMS per byte; bank = position % 32
//                      and the (group, byte) rotation make the 32 lanes of every ATOMS hit 32 banks.
//                      They also copy the line-end list to the global index.
//
//   full[s]  TMA -> scan warps     scanned[s]  scan warps -> record warps     freed[s]  record warps -> TMA
//
// Reference behaviour reproduced: see fq_kernels.cu header.
#include "fq_common.cuh"
#include "fq_device.cuh"

namespace fq {

template <int NCHUNK_, int TILE_, int NSTAGE_, int SCANW_>
struct Cfg {
    static constexpr int NCHUNK = NCHUNK_, TILE = TILE_, NSTAGE = NSTAGE_;
    static constexpr int SW = SCANW_;                      // scan warps
    static constexpr int HW = 31 - SCANW_;                 // record / histogram warps (warp 0: TMA)
    static constexpr int PPAD = 32 * NCHUNK;               // positions with a shared-memory counter column
    static constexpr int SM_TILE = FRONT + TILE + HALO;
    static constexpr int TILE_PAD = (SM_TILE + 16 + 127) / 128 * 128;
    static constexpr int NUNITS = (TILE + HALO) / UNIT;
    static constexpr int OWN_UNITS = TILE / UNIT;
    static constexpr int ITERS = (NUNITS + SW - 1) / SW;
    static constexpr int UPL = (NUNITS + 31) / 32;         // unit counts per lane
    static constexpr int LIST_CAP = TILE / 8;              // line ends the position list holds
    static constexpr int LIST_DUMMY = LIST_CAP + 16;       // writes beyond the capacity land here
    static constexpr int LIST_BYTES = (LIST_CAP * 2 + 64 + 127) / 128 * 128;
    static constexpr int CHUNK_WORDS = HIST_ROWS * 32;     // one 32-position chunk: [byte][32] u32
    static constexpr int HIST_WORDS = NCHUNK * CHUNK_WORDS;   // u32 = lo16 seq | hi16 qual
    static constexpr int LENH_WORDS = (PPAD + 2 + 31) / 32 * 32;
    // word loads of the record pass may run up to PPAD + 8 bytes past a tile buffer: keep them inside
    static constexpr int TAIL_PAD = (PPAD + 8 + 127) / 128 * 128;
    static constexpr int TOTAL = HIST_WORDS * 4 + LENH_WORDS * 4 + NSTAGE * (TILE_PAD + LIST_BYTES) + TAIL_PAD;
    static_assert(SM_TILE < 65536, "list entries are u16");
    static_assert(UPL <= 2, "unit counts: <= 64 units");
    static_assert(SW >= 1 && HW >= 1, "roles");
};

struct TileMeta {
    unsigned long long base;   // line number at the tile start (exact, or exact mod 4 when inferred)
    unsigned long long lrank;  // '\n' of the range before the tile
    unsigned long long ts;     // buffer-relative offset of the tile
    uint32_t front;            // 1 if the byte before the tile is (or acts as) '\n'
    uint32_t own_count;        // '\n' in the owned range
    uint32_t total_count;      // '\n' staged (owned + halo)
    uint32_t own_len;
};

template <int NUNITS>
struct StageCtl {
    unsigned long long full;      // mbarrier: tile bytes landed (TMA)
    unsigned long long scanned;   // mbarrier: position list + meta complete (SW arrivals)
    unsigned long long freed;     // mbarrier: stage may be overwritten (HW arrivals)
    TileMeta meta;
    uint32_t unit_all[NUNITS + 2];
    uint32_t unit_own[NUNITS + 2];
    uint32_t nonascii;            // tile holds a byte >= 0x80
};

struct CtaCtl {
    uint32_t recs_since_flush;
    uint32_t flush_epoch;
};

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
        const uint32_t v = atomicExch(hist + i, 0u);
        const uint32_t chunk = (uint32_t)i / C::CHUNK_WORDS, r = (uint32_t)i % C::CHUNK_WORDS;
        const uint32_t b = r >> 5, pos = chunk * 32u + (r & 31u);
        const uint32_t lo = v & 0xFFFFu, hi = v >> 16;
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
    unsigned long long win_end = s + MAXREC;
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

struct TileView {          // warp-uniform view of the tile being consumed
    const uint8_t* tile;
    const uint16_t* list;
    unsigned long long ts;   // buffer-relative offset of the tile
    uint32_t tile_s;         // shared address of `tile`
    uint32_t f;              // 1 if list[0] is the line end before the tile
    uint32_t nown;           // list entries that end a line inside the owned range (+f)
    uint32_t nstored;        // list entries stored
    uint32_t j0;             // first list entry after which a record starts
    uint32_t own_end;        // tile offset of the end of the owned range
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

template <class C, bool ASCII>
__device__ __forceinline__ void records_pass(const ScanParams& p, const TileView& tv, const LaneConst& lc,
                                             uint32_t* hist, uint32_t* lenh, unsigned long long limit,
                                             uint32_t pass, Acc& acc, int lane)
{
    const uint32_t sub = (uint32_t)lane >> 3, i = (uint32_t)lane & 7u;
    const uint8_t* tile = tv.tile;
    const uint32_t j = tv.j0 + 4u * (4u * pass + sub);
    // the five line ends around the record, loaded together (entries past the stored ones are stale
    // values that the predicates below never let through)
    const uint16_t* lp = tv.list + min(j, (uint32_t)C::LIST_CAP);
    const uint32_t l0 = lp[0], l1 = lp[1], l2 = lp[2], l3 = lp[3], l4 = lp[4];
    const uint32_t s = l0 + 1u;
    bool valid = j < tv.nown && s < tv.own_end;            // else it starts in the next tile
    if (limit != NONE64) valid = valid && (tv.ts + s - FRONT) < limit;
    const bool complete = valid && (j + 4u < tv.nstored);
    // stale entries must not turn into wild shared-memory addresses in the rounds below
    const uint32_t h = complete ? l1 : (uint32_t)FRONT, q = l2, pp = complete ? l3 : (uint32_t)FRONT, e = l4;
    bool ok = false;
    uint32_t c_at = 0, c_plus = 0, c_sr = 0, c_qr = 0;
    if (complete) {
        c_at = tile[s];
        c_plus = tile[q + 1];
        c_sr = tile[q - 1];
        c_qr = tile[e - 1];
        // src/records.rs:137-149 ('@'), :151-163 ('+'), :233-238 (raw line lengths equal)
        ok = c_at == '@' && c_plus == '+' && (e - pp) == (q - h);
        if (!ok && i == 0) atomicMin(&p.res->first_bad, tv.ts + s - FRONT);
    }
    if (ok && i == 0) acc.n_records++;

    if (p.flags & F_HIST) {
        const uint32_t P = p.max_len;
        const uint32_t Pm = P < (uint32_t)C::PPAD ? P : (uint32_t)C::PPAD;
        uint32_t Ls = 0, Lq = 0;
        if (ok) {
            const uint32_t Lr = q - h - 1u;
            // seq()/qual() drop one trailing '\r' (src/records.rs:65-73,82-90)
            Ls = Lr - ((Lr > 0 && c_sr == '\r') ? 1u : 0u);
            Lq = Lr - ((Lr > 0 && c_qr == '\r') ? 1u : 0u);
            if (i == 0) account_record<C>(acc, lenh, p, Ls, Lq);
        }
        RoundCtx c;
        c.ns = min(Ls, Pm);
        c.nq = min(Lq, Pm);
        c.nmax_w = __reduce_max_sync(0xffffffffu, max(c.ns, c.nq));
        c.nmin_w = __reduce_min_sync(0xffffffffu, min(c.ns, c.nq));
        const uint32_t sa = h + 1u + 4u * i;                // shared offset of position 4i of the sequence line
        const uint32_t qa = pp + 1u + 4u * i;
        c.as0 = tv.tile_s + (sa & ~3u);
        c.aq0 = tv.tile_s + (qa & ~3u);
        c.shs = (sa & 3u) * 8u;
        c.shq = (qa & 3u) * 8u;
        c.gseq = p.seqraw;
        c.gqual = p.stats + stats_qual_off(P);
        Rounds<C, ASCII, 0>::run(c, lc);
        // positions beyond the shared-memory columns but below P: straight to global (P > PPAD only)
        if (P > Pm && ok) {
            const uint32_t gs = min(Ls, P), gq = min(Lq, P);
            for (uint32_t g = Pm + i; g < gs; g += 8) atomicAdd(c.gseq + (size_t)g * 256 + tile[h + 1u + g], 1ull);
            for (uint32_t g = Pm + i; g < gq; g += 8) atomicAdd(c.gqual + (size_t)g * 256 + tile[pp + 1u + g], 1ull);
        }
    }

    // records that are not completely staged: whole warp, one at a time
    unsigned slow = __ballot_sync(0xffffffffu, valid && !complete && i == 0);
    while (slow) {
        const int src = __ffs(slow) - 1;
        slow &= slow - 1;
        const unsigned long long a = __shfl_sync(0xffffffffu, tv.ts + s - FRONT, src);
        record_global<C>(p, a, limit, hist, lenh, lane);
    }
}

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

// line ends of the owned range of a dense-newline tile (more than the list holds; never a healthy
// FASTQ), ranked straight into the index by one warp
template <class C>
__device__ __noinline__ void index_dense(uint32_t* idx_out, unsigned long long idx_cap, const uint8_t* tile,
                                         uint32_t own_count, unsigned long long idx_base,
                                         unsigned long long off_base, int lane)
{
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t run = 0;
    for (int u = 0; u < C::OWN_UNITS && run < own_count; ++u) {
        const uint32_t off = (uint32_t)u * UNIT + (uint32_t)lane * 16u;
        uint32_t mm = nlmask16s7(*reinterpret_cast<const uint4*>(tile + FRONT + off)) >> 7;
        uint32_t rank = run + small_prefix(__popc(mm), lt_mask);
        run += __reduce_add_sync(0xffffffffu, (uint32_t)__popc(mm));
        while (mm) {
            const uint32_t bit = (uint32_t)__ffs(mm) - 1u;
            mm &= mm - 1u;
            if (rank < own_count && idx_base + rank < idx_cap)
                idx_out[idx_base + rank] = (uint32_t)(off_base + off + bit);
            ++rank;
        }
    }
}

// ------------------------------------------------------------------------------------------
// phase inference at the start of a CTA range (one warp): which list entry j0 in 0..3 is followed by
// a record start?  lane = 8 * candidate + r tests record r of the candidate: '@' after entry j,
// '+' after entry j + 2, equal raw lengths.  Accepted only if exactly one candidate passes all the
// records it could test (at least two).  Returns j0, or 4 when the range start is ambiguous.
// ------------------------------------------------------------------------------------------
template <class C>
__device__ __forceinline__ uint32_t infer_j0(const uint8_t* tile, const uint16_t* list, uint32_t nstored, int lane)
{
    const uint32_t cand = (uint32_t)lane >> 3, r = (uint32_t)lane & 7u;
    const uint32_t j = cand + 4u * r;
    const bool testable = j + 4u < nstored;
    bool good = true;
    if (testable) {
        const uint32_t s = (uint32_t)list[j] + 1u, h = list[j + 1], q = list[j + 2], pp = list[j + 3], e = list[j + 4];
        good = tile[s] == '@' && tile[q + 1] == '+' && (e - pp) == (q - h);
    }
    const unsigned tested = __ballot_sync(0xffffffffu, testable);
    const unsigned bad = __ballot_sync(0xffffffffu, testable && !good);
    uint32_t pass = 0, npass = 0;
#pragma unroll
    for (uint32_t c = 0; c < 4; ++c) {
        const unsigned m = 0xFFu << (8 * c);
        if (__popc(tested & m) >= 2 && !(bad & m)) {
            pass = c;
            ++npass;
        }
    }
    return npass == 1 ? pass : 4u;
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(1024, 1) fq_scan_kernel(const __grid_constant__ ScanParams p)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* lenh = hist + C::HIST_WORDS;
    __shared__ StageCtl<C::NUNITS> stage_ctl[C::NSTAGE];
    __shared__ CtaCtl cta;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;                    // 0 = TMA, 1..SW = scan, rest = records
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint8_t* stage_mem = smem_raw + C::HIST_WORDS * 4 + C::LENH_WORDS * 4;

    // launch 1 infers the range phases; launch 2 (F_BASES) runs only when the inference failed;
    // launch 3 (F_BASES | F_RERUN) only when a bad record was found, restricted to the records before it
    unsigned long long limit = NONE64;
    if (p.flags & F_RERUN) {
        limit = p.res->first_bad;                 // written by the earlier launches, stable during this one
        if (limit == NONE64) return;
    } else if ((p.flags & F_BASES) && !p.res->spec_fail) {
        return;
    }
    if ((p.flags & F_CARRY) && p.carry->status != 0) return;   // the stream already failed
    const unsigned long long line_base = (p.flags & F_CARRY) ? p.carry->line_base : p.line_base;

    for (int i = tid; i < C::HIST_WORDS + C::LENH_WORDS; i += 1024) hist[i] = 0;
    if (tid == 0) {
        for (int s = 0; s < C::NSTAGE; ++s) {
            mbar_init(&stage_ctl[s].full, 1);
            mbar_init(&stage_ctl[s].scanned, C::SW);
            mbar_init(&stage_ctl[s].freed, C::HW);
            stage_ctl[s].nonascii = 0;
        }
        fence_mbar_init();
        cta.recs_since_flush = 0;
        cta.flush_epoch = 0;
    }
    __syncthreads();

    // static schedule: CTA b handles the tiles [b * tiles_per_cta, (b + 1) * tiles_per_cta)
    uint32_t ntiles_eff = p.ntiles;
    if (limit != NONE64) {
        const unsigned long long lt = limit / C::TILE + 1;   // tiles that start below the limit
        if (lt < ntiles_eff) ntiles_eff = (uint32_t)lt;
    }
    const unsigned long long tile0 = (unsigned long long)blockIdx.x * p.tiles_per_cta;
    const int K = tile0 < ntiles_eff ? (int)min((unsigned long long)p.tiles_per_cta, ntiles_eff - tile0) : 0;
    // the range of CTA 0 starts at the shard start, whose line number the caller gave; the others
    // infer theirs unless the exact bases are there
    const bool infer = blockIdx.x != 0 && !(p.flags & F_BASES);
    const bool staged_index = !(p.flags & F_BASES);          // ranks relative to the range -> index_stage
    const bool want_index = (p.flags & F_INDEX) && !(p.flags & F_RERUN) && p.index != nullptr && p.index_cap != 0;

    auto tile_no = [&](int k) -> uint32_t { return (uint32_t)tile0 + (uint32_t)k; };
    auto tile_buf = [&](int k) -> uint8_t* { return stage_mem + (k % C::NSTAGE) * (C::TILE_PAD + C::LIST_BYTES); };
    auto tile_list = [&](int k) -> uint16_t* { return reinterpret_cast<uint16_t*>(tile_buf(k) + C::TILE_PAD); };

    Acc acc = {0, 0, 0, 0};

    if (warp == 0) {
        // =====================================================================================
        // TMA warp: one bulk copy per tile; ragged edges are filled by the scan warps
        // =====================================================================================
        for (int k = 0; k < K; ++k) {
            StageCtl<C::NUNITS>& sc = stage_ctl[k % C::NSTAGE];
            if (k >= C::NSTAGE) mbar_wait(&sc.freed, (uint32_t)(k / C::NSTAGE - 1) & 1u);
            if (lane == 0) {
                const unsigned long long ts = (unsigned long long)tile_no(k) * C::TILE;
                const uint32_t data_len = (uint32_t)min((unsigned long long)(C::TILE + HALO), p.n_avail - ts);
                const uint32_t front = (ts || (p.flags & F_FRONT16)) ? FRONT : 0;
                const uint32_t bulk = (front + data_len) & ~15u;
                sc.nonascii = 0;
                fence_proxy_async();
                mbar_arrive_expect_tx(&sc.full, bulk);
                if (bulk) bulk_g2s(tile_buf(k) + FRONT - front, p.data + ts - front, bulk, &sc.full);
                trace_ev(p, k, 0);
            }
            __syncwarp();
        }
    } else if (warp <= C::SW) {
        // =====================================================================================
        // scan warps: newline masks -> unit counts -> ranks -> position list
        // warp sw owns the units [sw * ITERS, (sw + 1) * ITERS) of every tile
        // =====================================================================================
        const int sw = warp - 1;
        unsigned long long lbase = blockIdx.x == 0 ? line_base : ((p.flags & F_BASES) ? p.ranges[blockIdx.x].base : 0ull);
        unsigned long long lrank = 0;
        uint32_t spec_phase = 0, spec_flags = 0;
        const int sthreads = C::SW * 32;
        const int stid = sw * 32 + lane;
        const int u0 = sw * C::ITERS;
        for (int k = 0; k < K; ++k) {
            StageCtl<C::NUNITS>& sc = stage_ctl[k % C::NSTAGE];
            uint8_t* tile = tile_buf(k);
            uint16_t* list = tile_list(k);
            const uint32_t tn = tile_no(k);
            const unsigned long long ts = (unsigned long long)tn * C::TILE;
            const uint32_t own_len = (uint32_t)min((unsigned long long)C::TILE, p.n_own - ts);
            const uint32_t data_len = (uint32_t)min((unsigned long long)(C::TILE + HALO), p.n_avail - ts);
            const uint32_t front = (ts || (p.flags & F_FRONT16)) ? FRONT : 0;
            const uint32_t span = front + data_len;
            mbar_wait(&sc.full, (uint32_t)(k / C::NSTAGE) & 1u);
            if (stid == 0) trace_ev(p, k, 1);
            if (span != (uint32_t)C::SM_TILE) {
                // ragged first / last tiles: leading zeros (or the virtual '\n' of a line start), the
                // bytes the 16-byte-granular bulk copy leaves out, and zero fill
                const uint32_t bulk = span & ~15u;
                const uint8_t* src = p.data + ts - front;
                uint8_t* dst = tile + FRONT - front;
                const bool virt_nl = ts == 0 && front == 0 && (p.flags & F_LINE_START);
                for (uint32_t i = stid; i < FRONT - front; i += sthreads) tile[i] = (virt_nl && i == FRONT - 1) ? '\n' : 0;
                for (uint32_t i = bulk + stid; i < span; i += sthreads) dst[i] = src[i];
                for (uint32_t i = FRONT + data_len + stid; i < (uint32_t)C::SM_TILE; i += sthreads) tile[i] = 0;
                named_bar(1, sthreads);
            }
            // ---- pass 1: newline masks, per-unit counts ------------------------------------------
            uint32_t mask[C::ITERS], call[C::ITERS];
            uint32_t hib = 0;
#pragma unroll
            for (int it = 0; it < C::ITERS; ++it) {
                const int u = u0 + it;
                mask[it] = 0;
                call[it] = 0;
                if (u < C::NUNITS) {
                    const uint32_t off = (uint32_t)u * UNIT + (uint32_t)lane * 16u;
                    const uint4 v = *reinterpret_cast<const uint4*>(tile + FRONT + off);
                    hib |= v.x | v.y | v.z | v.w;
                    const uint32_t mm = nlmask16s7(v);
                    mask[it] = mm;
                    call[it] = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(mm));
                    uint32_t cown;
                    if (own_len == (uint32_t)C::TILE) {
                        cown = u < C::OWN_UNITS ? call[it] : 0u;
                    } else {
                        const int rem = (int)own_len - (int)off;
                        const uint32_t ownm = rem >= 16 ? 0xFFFFu : (rem > 0 ? ((1u << rem) - 1u) : 0u);
                        cown = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(mm & (ownm << 7)));
                    }
                    if (lane == 0) {
                        sc.unit_all[u] = call[it];
                        sc.unit_own[u] = cown;
                    }
                }
            }
            if (__any_sync(0xffffffffu, (hib & 0x80808080u) != 0) && lane == 0) atomicOr(&sc.nonascii, 1u);
            named_bar(1, sthreads);   // unit counts of the tile are visible to all scan warps
            if (stid == 0) trace_ev(p, k, 2);
            // ---- totals, published for the look-backs of the other CTAs ---------------------------
            const uint32_t f = tile[FRONT - 1] == '\n' ? 1u : 0u;
            uint32_t cnt[C::UPL];
#pragma unroll
            for (int r = 0; r < C::UPL; ++r) cnt[r] = (lane + 32 * r) < C::NUNITS ? sc.unit_all[lane + 32 * r] : 0u;
            uint32_t own_count_t = 0;
            if (sw == 0) {
                uint32_t a = 0, o = 0;
#pragma unroll
                for (int r = 0; r < C::UPL; ++r) {
                    a += cnt[r];
                    o += (lane + 32 * r) < C::NUNITS ? sc.unit_own[lane + 32 * r] : 0u;
                }
                const uint32_t total = __reduce_add_sync(0xffffffffu, a);
                own_count_t = __reduce_add_sync(0xffffffffu, o);
                if (lane == 0) {
                    sc.meta.base = lbase;
                    sc.meta.lrank = lrank;
                    sc.meta.ts = ts;
                    sc.meta.front = f;
                    sc.meta.own_count = own_count_t;
                    sc.meta.total_count = total;
                    sc.meta.own_len = own_len;
                    if (f) list[0] = FRONT - 1;
                }
            }
            // ---- pass 2: rank every newline, fill the position list --------------------------------
            uint32_t x = 0;
#pragma unroll
            for (int r = 0; r < C::UPL; ++r) x += (lane + 32 * r < u0) ? cnt[r] : 0u;
            uint32_t ubase = f + __reduce_add_sync(0xffffffffu, x);   // rank of the first newline of unit u0
#pragma unroll
            for (int it = 0; it < C::ITERS; ++it) {
                const int u = u0 + it;
                if (u < C::NUNITS) {
                    const uint32_t mm = mask[it];                     // bit 7 + i = byte i
                    const int c = __popc(mm);
                    const uint32_t rank = ubase + small_prefix(c, lt_mask);
                    const uint32_t pos0 = FRONT + (uint32_t)u * UNIT + (uint32_t)lane * 16u - 7u;
                    // the first and the last newline of the piece, no loop; a third one is rare
                    if (c > 0) list[min(rank, (uint32_t)C::LIST_DUMMY)] = (uint16_t)(pos0 + (uint32_t)__ffs(mm) - 1u);
                    if (c > 1) list[min(rank + (uint32_t)c - 1u, (uint32_t)C::LIST_DUMMY)] = (uint16_t)(pos0 + 31u - (uint32_t)__clz(mm));
                    if (__any_sync(0xffffffffu, c > 2) && c > 2) {
                        uint32_t m2 = mm & (mm - 1u), r2 = rank + 1u;
                        while (m2 & (m2 - 1u)) {
                            list[min(r2, (uint32_t)C::LIST_DUMMY)] = (uint16_t)(pos0 + (uint32_t)__ffs(m2) - 1u);
                            m2 &= m2 - 1u;
                            ++r2;
                        }
                    }
                    ubase += call[it];
                }
            }
            if (infer && k == 0) {
                // the first tile of the range: its complete list decides the phase of the whole range
                __syncwarp();
                named_bar(1, sthreads);
                if (sw == 0) {
                    const uint32_t nstored = min(f + sc.meta.total_count, (uint32_t)C::LIST_CAP);
                    const uint32_t j0 = infer_j0<C>(tile, list, nstored, lane);
                    spec_flags = j0 < 4u ? 1u : 2u;
                    // entry j ends line base - f + j; a record starts after every line = 3 (mod 4)
                    spec_phase = (3u + f - (j0 & 3u)) & 3u;
                    lbase = spec_phase;
                    if (lane == 0) {
                        sc.meta.base = lbase;
                        if (j0 >= 4u) atomicExch(&p.res->spec_fail, 1);
                    }
                }
            }
            if (sw == 0) {
                lbase += own_count_t;
                lrank += own_count_t;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sc.scanned);
            if (stid == 0) trace_ev(p, k, 3);
            if (lane == 0) trace_ev(p, k, 10, true);
        }
        if (sw == 0 && lane == 0 && !(p.flags & F_RERUN)) {
            // what fq_verify_kernel needs: the newline count of the range and the inferred phase
            RangeInfo& ri = p.ranges[blockIdx.x];
            ri.count = lrank;
            if (!(p.flags & F_BASES)) {
                ri.spec_phase = spec_phase;
                ri.flags = spec_flags;
            }
        }
    } else {
        // =====================================================================================
        // record warps: validation + per-position histograms + index copy; the items of a tile
        // (record passes, then one index item) are dealt round-robin, rotated from tile to tile
        // =====================================================================================
        const int hw = warp - 1 - C::SW;
        uint32_t my_epoch = 0;
        constexpr int SLICE = (C::HIST_WORDS + C::HW - 1) / C::HW;
        LaneConst lc;
        {
            const uint32_t sub = (uint32_t)lane >> 3, i = (uint32_t)lane & 7u;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint32_t bytek = ((uint32_t)kk + sub) & 3u;
                lc.pk[kk] = 4u * i + bytek;
                lc.hk[kk] = smem_u32(hist) + 4u * lc.pk[kk];
                lc.wsel[kk] = 128u << (8u * bytek);
            }
        }
        for (int k = 0; k < K; ++k) {
            StageCtl<C::NUNITS>& sc = stage_ctl[k % C::NSTAGE];
            const uint32_t par = (uint32_t)(k / C::NSTAGE) & 1u;
            mbar_wait(&sc.scanned, par);
            if (hw == 0 && lane == 0) trace_ev(p, k, 7);
            const TileMeta mk = sc.meta;
            TileView tv;
            tv.tile = tile_buf(k);
            tv.list = tile_list(k);
            tv.ts = mk.ts;
            tv.tile_s = smem_u32(tv.tile);
            tv.f = mk.front;
            tv.nown = mk.front + mk.own_count;
            tv.nstored = min(mk.front + mk.total_count, (uint32_t)C::LIST_CAP);
            tv.j0 = (3u - (uint32_t)((mk.base - mk.front) & 3ull)) & 3u;   // entry j ends global line base - f + j
            tv.own_end = FRONT + mk.own_len;
            const bool overflow = mk.front + mk.total_count > (uint32_t)C::LIST_CAP;
            const uint32_t nrec = tv.nown > tv.j0 ? (tv.nown - tv.j0 + 3u) / 4u : 0u;
            const uint32_t npass = (nrec + 3u) / 4u;
            const uint32_t first = (uint32_t)((hw + C::HW - (k % C::HW)) % C::HW);
            // where the line ends of this tile go: ranks relative to the range into the staging area of
            // this CTA, or (exact bases) straight into the caller's index
            uint32_t* const idx_out = staged_index ? p.index_stage + (size_t)blockIdx.x * p.stage_share : p.index;
            const unsigned long long idx_cap = staged_index ? p.stage_share : p.index_cap;
            const unsigned long long idx_base = staged_index ? mk.lrank : mk.base - line_base;
            const unsigned long long off_base = p.stream_offset + mk.ts;
            if (hw == 0 && lane == 0 && nrec) {
                // u16 counter halves: when the CTA-wide record count passes the mark, every record warp
                // drains its slice of the table before its next tile
                const uint32_t tot = cta.recs_since_flush + nrec;
                if (tot >= 24000u) {
                    cta.recs_since_flush = 0;
                    atomicAdd(&cta.flush_epoch, 1u);
                } else {
                    cta.recs_since_flush = tot;
                }
            }
            if (!overflow) {
                const bool nonascii = sc.nonascii != 0;
                const uint32_t nitems = npass + (want_index ? 1u : 0u);
                for (uint32_t item = first; item < nitems; item += C::HW) {
                    if (item < npass) {
                        if (nonascii)
                            records_pass<C, false>(p, tv, lc, hist, lenh, limit, item, acc, lane);
                        else
                            records_pass<C, true>(p, tv, lc, hist, lenh, limit, item, acc, lane);
                    } else {
                        for (uint32_t i = lane; i < mk.own_count; i += 32) {
                            const unsigned long long gi = idx_base + i;
                            if (gi < idx_cap) idx_out[gi] = (uint32_t)(off_base + (uint32_t)tv.list[tv.f + i] - FRONT);
                        }
                        // a staging share too small for this range: the exact second launch writes the index
                        if (staged_index && lane == 0 && idx_base + mk.own_count > idx_cap) atomicExch(&p.res->spec_fail, 1);
                    }
                }
            } else {
                // dense-newline tile: one warp walks the records one by one, another ranks the index
                if (first == 0 && tv.j0 < tv.nown) {
                    unsigned long long s = mk.ts + (uint32_t)tv.list[tv.j0] + 1u - FRONT;   // j0 < 4 <= LIST_CAP: stored
                    const unsigned long long tend = mk.ts + mk.own_len;
                    while (s < tend && s < limit) {
                        const unsigned long long e = record_global<C>(p, s, limit, hist, lenh, lane);
                        if (e == NONE64) break;
                        s = e + 1;
                    }
                }
                if (want_index && first == (uint32_t)(1 % C::HW))
                {
                    index_dense<C>(idx_out, idx_cap, tv.tile, mk.own_count, idx_base, off_base, lane);
                    if (staged_index && lane == 0 && idx_base + mk.own_count > idx_cap) atomicExch(&p.res->spec_fail, 1);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sc.freed);
            if (hw == 0 && lane == 0) trace_ev(p, k, 8);
            if (lane == 0) trace_ev(p, k, 9, true);
            const uint32_t ep = *reinterpret_cast<volatile uint32_t*>(&cta.flush_epoch);
            if (ep != my_epoch) {
                my_epoch = ep;
                flush_hist<C>(hist, p, hw * SLICE, min((hw + 1) * SLICE, C::HIST_WORDS), lane, 32);
            }
        }
    }

    // ---- drain -----------------------------------------------------------------------------
    __syncthreads();
    flush_hist<C>(hist, p, 0, C::HIST_WORDS, tid, 1024);
    {
        unsigned long long* lenh_g = p.stats + stats_len_off(p.max_len);
        for (int i = tid; i < C::PPAD + 2; i += 1024) {
            const uint32_t v = lenh[i];
            if (v) atomicAdd(lenh_g + i, (unsigned long long)v);
        }
    }
    acc.n_records = warp_sum_u64(acc.n_records);
    acc.n_bases = warp_sum_u64(acc.n_bases);
    acc.clip_seq = warp_sum_u64(acc.clip_seq);
    acc.clip_qual = warp_sum_u64(acc.clip_qual);
    if (lane == 0) {
        if (acc.n_records) atomicAdd(p.stats + 0, acc.n_records);
        if (acc.n_bases) atomicAdd(p.stats + 1, acc.n_bases);
        if (acc.clip_seq) atomicAdd(p.stats + 2, acc.clip_seq);
        if (acc.clip_qual) atomicAdd(p.stats + 3, acc.clip_qual);
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
using Cfg5 = Cfg<5, 16384, 4, 12>;    // P <= 160: 34 units over 12 scan warps (3 each), 18 record warps
using Cfg10 = Cfg<10, 8192, 3, 9>;    // P <= 320 (longer reads: positions >= 320 go to global atomics)

size_t scan_smem_bytes(int nchunk) { return nchunk <= 5 ? (size_t)Cfg5::TOTAL : (size_t)Cfg10::TOTAL; }
uint32_t scan_tile_bytes(int nchunk) { return nchunk <= 5 ? (uint32_t)Cfg5::TILE : (uint32_t)Cfg10::TILE; }

cudaError_t scan_configure()
{
    cudaError_t e =
        cudaFuncSetAttribute(fq_scan_kernel<Cfg5>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg5::TOTAL);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(fq_scan_kernel<Cfg10>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg10::TOTAL);
}

// the static tile schedule needs every CTA of the grid resident: one CTA per SM
int scan_blocks_per_sm(int) { return 1; }

cudaError_t launch_scan(const ScanParams& p, int nchunk, int grid, cudaStream_t st)
{
    if (nchunk <= 5)
        fq_scan_kernel<Cfg5><<<grid, 1024, Cfg5::TOTAL, st>>>(p);
    else
        fq_scan_kernel<Cfg10><<<grid, 1024, Cfg10::TOTAL, st>>>(p);
    return cudaGetLastError();
}

}  // namespace fq
