// fq_scan.cu -- the EXACT sm_100a scan kernel (K1 delimit + K2 per-position histograms).
//
// This is the kernel that needs no assumption about the input: it is given the exact stream line
// number of every CTA range (fq_range_count_kernel + fq_range_prefix_kernel count the newlines of
// the ranges first) and therefore knows which lines are headers whatever the bytes look like.  It
// runs (a) when the speculative warp-autonomous kernel (fq_stream.cu) reported that it could not
// stand by its result -- malformed input, records longer than its window, ambiguous range starts --,
// (b) for read lengths its shared-memory layout does not cover, and (c) as the relaunch restricted
// to the records in front of the first bad one (each() delivers exactly those).
//
// Persistent grid, one 1024-thread CTA per SM; CTA b owns a CONTIGUOUS range of tiles, so the line
// number of a tile is the line number of the range start plus a running count.  The 32 warps of a
// CTA are SPECIALISED and hand tiles to each other through a ring of NSTAGE shared-memory stages
// guarded by mbarriers -- no CTA-wide barrier in the steady state:
//
//   TMA warp (1)       keeps the ring full: UBLKCP bulk copies completing on full[s]
//   scan warps (SW)    16-byte SWAR newline masks (3 ops / word + dp4a bit gather) -> per-unit counts
//                      -> ranks -> position list of the tile (u16, shared memory); scan warp 0 keeps
//                      the running line number of the range
//   record warps (HW)  8 lanes per record: '@' / '+' / raw-length validation
//                      (src/records.rs:201-247), then the per-position histogram rounds of
//                      fq_hist.cuh; they also copy the line-end list to the global index.
//
//   full[s]  TMA -> scan warps     scanned[s]  scan warps -> record warps     freed[s]  record warps -> TMA
//
// Reference behaviour reproduced: see fq_kernels.cu header.
#include "fq_hist.cuh"

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
    static constexpr int ROW0 = 0;                         // (first byte value with a row: all of them)
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
    unsigned long long base;   // stream line number at the tile start
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

    // launch 1 runs only when the speculative kernel did not deliver; launch 2 (F_RERUN) only when a
    // bad record was found, restricted to the records before it
    unsigned long long limit = NONE64;
    if (p.flags & F_RERUN) {
        limit = p.res->first_bad;                 // written by the earlier launches, stable during this one
        if (limit == NONE64 || p.res->tail_err) return;
    } else if (!p.res->spec_fail) {
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
        unsigned long long lbase = K ? p.ranges[blockIdx.x].base : 0ull;   // exact line number of the range start
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
            if (sw == 0) lbase += own_count_t;
            __syncwarp();
            if (lane == 0) mbar_arrive(&sc.scanned);
            if (stid == 0) trace_ev(p, k, 3);
            if (lane == 0) trace_ev(p, k, 10, true);
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
            uint32_t* const idx_out = p.index;
            const unsigned long long idx_cap = p.index_cap;
            const unsigned long long idx_base = mk.base - line_base;   // buffer-local number of the first own line
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
                    index_dense<C>(idx_out, idx_cap, tv.tile, mk.own_count, idx_base, off_base, lane);
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
