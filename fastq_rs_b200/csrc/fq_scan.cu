// fq_scan.cu -- the fused sm_100a scan kernel (K1 delimit + K2 per-position histograms).
//
// One pass over the bytes.  Persistent grid, one 1024-thread CTA per SM, split into TEAMS
// independent teams that share one shared-memory histogram.  Every team runs a software
// pipeline over its statically assigned tiles (tile = team + k * n_teams):
//
//   TMA ring      cp.async.bulk of tile k+NBUF-1 is in flight                  (UBLKCP + mbarrier)
//   scan(k+1)     16-byte SWAR newline masks -> per-unit counts -> ranks -> position list;
//                 the tile's newline count is published right away (decoupled look-back)
//   look-back(k+1) warp 0 sums the predecessors' counts with one wide window of loads,
//                 hidden behind ...
//   records(k)    8 lanes per record: '@' / '+' / raw-length validation (src/records.rs:201-247),
//                 then each lane walks 4-byte groups of the sequence and quality lines and
//                 bumps hist[byte][position] -- bank = position % 32, and the (group, byte)
//                 rotation below makes the 32 lanes of every ATOMS hit 32 different banks
//
// Reference behaviour reproduced: see fq_kernels.cu header.
#include "fq_common.cuh"
#include "fq_device.cuh"

namespace fq {

template <int NCHUNK_, int TEAMS_, int TILE_, int NBUF_>
struct Cfg {
    static constexpr int NCHUNK = NCHUNK_, TEAMS = TEAMS_, TILE = TILE_, NBUF = NBUF_;
    static constexpr int PPAD = 32 * NCHUNK;               // positions with a shared-memory counter column
    static constexpr int TW = 32 / TEAMS;                  // warps per team
    static constexpr int TT = TW * 32;                     // threads per team
    static constexpr int SM_TILE = FRONT + TILE + HALO;
    static constexpr int TILE_PAD = (SM_TILE + 16 + 127) / 128 * 128;  // slack: word loads may run past the end
    static constexpr int NUNITS = (TILE + HALO) / UNIT;
    static constexpr int OWN_UNITS = TILE / UNIT;
    static constexpr int ITERS = (NUNITS + TW - 1) / TW;
    static constexpr int UPL = (NUNITS + 31) / 32;         // unit counts per lane in the warp scan
    static constexpr int LIST_CAP = TILE / 4;
    static constexpr int HIST_WORDS = HIST_ROWS * PPAD;    // hist[byte][position], u32 = lo16 seq | hi16 qual
    static constexpr int LENH_WORDS = (PPAD + 2 + 31) / 32 * 32;
    static constexpr int TEAM_BYTES = NBUF * TILE_PAD + 2 * LIST_CAP * 2;
    static constexpr int TOTAL = HIST_WORDS * 4 + LENH_WORDS * 4 + TEAMS * TEAM_BYTES;
    static constexpr uint32_t ROW_BYTES = PPAD * 4;
    static_assert(SM_TILE < 65536, "list entries are u16");
    static_assert(UPL <= 3, "warp scan handles <= 96 units");
};

struct TileMeta {
    unsigned long long base;   // stream-global exclusive line count at the tile start
    unsigned long long ts;     // buffer-relative offset of the tile
    uint32_t front;            // 1 if the byte before the tile is (or acts as) '\n'
    uint32_t own_count;        // '\n' in the owned range
    uint32_t total_count;      // '\n' staged (owned + halo)
    uint32_t own_len;
    uint32_t nonascii;
    uint32_t pad;
};

template <int NBUF, int NUNITS>
struct TeamCtl {
    unsigned long long full[NBUF];   // mbarriers: tile bytes landed
    TileMeta meta[2];
    uint32_t unit_all[NUNITS + 2];
    uint32_t unit_own[NUNITS + 2];
    uint32_t pass_counter[2];   // records(k) hands out passes from pass_counter[k & 1]
    int flush_iter;             // iteration whose end this team drains the shared counters at
    uint32_t pad;
};

struct CtaCtl {
    uint32_t recs_since_flush;
};

template <int ID_BASE, int NTHREADS>
__device__ __forceinline__ void team_bar(int team)
{
    asm volatile("bar.sync %0, %1;" ::"r"(ID_BASE + team), "r"(NTHREADS) : "memory");
}

__device__ __forceinline__ void red_shared_add(uint32_t addr, uint32_t v)
{
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------
// lock-free drain of the u16-pair counters: atomicExch leaves concurrent increments of the other
// teams intact, so a team may flush whenever the CTA-wide record counter says the halves could
// approach 65535
// ------------------------------------------------------------------------------------------
template <class C>
__device__ void flush_hist(uint32_t* hist, const ScanParams& p, int tid, int nthreads)
{
    const uint32_t P = p.max_len;
    unsigned long long* qual = p.stats + stats_qual_off(P);
    for (int i = tid; i < C::HIST_WORDS; i += nthreads) {
        if (hist[i] == 0) continue;
        const uint32_t v = atomicExch(hist + i, 0u);
        const uint32_t b = (uint32_t)i / C::PPAD, pos = (uint32_t)i % C::PPAD;
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
__device__ unsigned long long record_global(const ScanParams& p, unsigned long long s, unsigned long long limit,
                                            uint32_t* hist, uint32_t* lenh, Acc& acc, int lane)
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
    if (lane == 0) acc.n_records++;
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
                atomicAdd(hist + b * C::PPAD + c, 1u);
            else
                atomicAdd(p.seqraw + (size_t)c * 256 + b, 1ull);
        }
        for (uint32_t c = lane; c < nq; c += 32) {
            const uint32_t b = ql[c];
            if (c < Pm && b < (uint32_t)HIST_ROWS)
                atomicAdd(hist + b * C::PPAD + c, 0x10000u);
            else
                atomicAdd(qualg + (size_t)c * 256 + b, 1ull);
        }
        if (lane == 0) account_record<C>(acc, lenh, p, Ls, Lq);
    }
    return nl[3];
}

// one byte observation of a 4-byte group; hk = shared address of hist[0][position]
template <bool ASCII>
__device__ __forceinline__ void bump(uint32_t hk, uint32_t b, uint32_t inc, uint32_t row_bytes,
                                     unsigned long long* grow /* &g[pos*256] */)
{
    if (ASCII || b < (uint32_t)HIST_ROWS)
        red_shared_add(hk + b * row_bytes, inc);
    else
        atomicAdd(grow + b, 1ull);
}

// ------------------------------------------------------------------------------------------
// records: one pass = 4 records per warp, 8 lanes each
// lane = 8*sub + i.  In round t lane (sub,i) owns the 4-byte group g = i + 8t of its record's
// sequence and quality lines and visits its bytes in the order (k + sub) & 3, k = 0..3, so that
// the k-th ATOMS of the round touches position 4g + ((k+sub)&3): over the 32 lanes these are 32
// different residues mod 32 = 32 different banks of hist[byte][position].
// ------------------------------------------------------------------------------------------
template <class C, bool ASCII>
__device__ __forceinline__ void records_pass(const ScanParams& p, const TileMeta& m, const uint8_t* tile,
                                             const uint16_t* list, uint32_t* hist, uint32_t* lenh,
                                             unsigned long long limit, uint32_t pass, Acc& acc, int lane)
{
    const uint32_t sub = (uint32_t)lane >> 3, i = (uint32_t)lane & 7u;
    const uint32_t f = m.front;
    const uint32_t nown = f + m.own_count;
    const uint32_t nstored = min(f + m.total_count, (uint32_t)C::LIST_CAP);
    const uint32_t gb = (uint32_t)((m.base - f) & 3ull);   // list entry j ends global line (base - f + j)
    const uint32_t j0 = (3u - gb) & 3u;                     // a record starts after every line = 3 (mod 4)
    const uint32_t own_end = FRONT + m.own_len;
    const uint32_t j = j0 + 4u * (4u * pass + sub);

    bool valid = j < nown;
    uint32_t s = FRONT;
    if (valid) {
        s = (uint32_t)list[j] + 1u;
        valid = s < own_end;                                // else it starts in the next tile
    }
    const unsigned long long abs_s = m.ts + s - FRONT;
    valid = valid && abs_s < limit;
    const bool complete = valid && (j + 4u < nstored);
    uint32_t h = FRONT, q = FRONT, pp = FRONT, e = FRONT;
    bool ok = false;
    if (complete) {
        h = list[j + 1];
        q = list[j + 2];
        pp = list[j + 3];
        e = list[j + 4];
        // src/records.rs:137-149 ('@'), :151-163 ('+'), :233-238 (raw line lengths equal)
        ok = tile[s] == '@' && tile[q + 1] == '+' && (e - pp) == (q - h);
        if (!ok && i == 0) atomicMin(&p.res->first_bad, abs_s);
    }
    if (ok && i == 0) acc.n_records++;

    if (p.flags & F_HIST) {
        const uint32_t P = p.max_len;
        const uint32_t Pm = P < (uint32_t)C::PPAD ? P : (uint32_t)C::PPAD;
        uint32_t Ls = 0, Lq = 0;
        if (ok) {
            const uint32_t Lr = q - h - 1u;
            // seq()/qual() drop one trailing '\r' (src/records.rs:65-73,82-90)
            Ls = Lr - ((Lr > 0 && tile[q - 1] == '\r') ? 1u : 0u);
            Lq = Lr - ((Lr > 0 && tile[e - 1] == '\r') ? 1u : 0u);
            if (i == 0) account_record<C>(acc, lenh, p, Ls, Lq);
        }
        const uint32_t ns = min(Ls, Pm), nq = min(Lq, Pm);
        const uint32_t nmax_w = __reduce_max_sync(0xffffffffu, max(ns, nq));
        const uint32_t nmin_w = __reduce_min_sync(0xffffffffu, min(ns, nq));
        unsigned long long* qualg = p.stats + stats_qual_off(P);

        uint32_t sa = h + 1u + 4u * i;                      // shared offset of position 4i of the sequence line
        uint32_t qa = pp + 1u + 4u * i;
        const uint32_t rot = 8u * sub;
        const uint32_t tile_s = smem_u32(tile);
        uint32_t hk[4];
        uint32_t pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            pk[k] = 4u * i + (((uint32_t)k + sub) & 3u);    // position visited by the k-th bump in round 0
            hk[k] = smem_u32(hist) + 4u * pk[k];
        }
        constexpr uint32_t MAXOFF = C::TILE_PAD - 8;
        for (uint32_t t = 0; 32u * t < nmax_w; ++t) {
            const uint32_t as = min(sa & ~3u, MAXOFF), aq = min(qa & ~3u, MAXOFF);
            uint32_t s0, s1, q0, q1;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(s0) : "r"(tile_s + as));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(s1) : "r"(tile_s + as + 4u));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(q0) : "r"(tile_s + aq));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(q1) : "r"(tile_s + aq + 4u));
            uint32_t vs = __funnelshift_r(s0, s1, (sa & 3u) * 8u);
            uint32_t vq = __funnelshift_r(q0, q1, (qa & 3u) * 8u);
            vs = __funnelshift_r(vs, vs, rot);               // byte k of vs = byte (k+sub)&3 of the group
            vq = __funnelshift_r(vq, vq, rot);
            if (32u * (t + 1u) <= nmin_w) {                  // every lane's group is inside both lines
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    bump<ASCII>(hk[k], __byte_perm(vs, 0, 0x4440 + k), 1u, C::ROW_BYTES,
                                p.seqraw + (size_t)(pk[k] + 32u * t) * 256);
                    bump<ASCII>(hk[k], __byte_perm(vq, 0, 0x4440 + k), 0x10000u, C::ROW_BYTES,
                                qualg + (size_t)(pk[k] + 32u * t) * 256);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t pos = pk[k] + 32u * t;
                    if (pos < ns)
                        bump<ASCII>(hk[k], __byte_perm(vs, 0, 0x4440 + k), 1u, C::ROW_BYTES,
                                    p.seqraw + (size_t)pos * 256);
                    if (pos < nq)
                        bump<ASCII>(hk[k], __byte_perm(vq, 0, 0x4440 + k), 0x10000u, C::ROW_BYTES,
                                    qualg + (size_t)pos * 256);
                }
            }
            sa += 32u;
            qa += 32u;
#pragma unroll
            for (int k = 0; k < 4; ++k) hk[k] += 128u;
        }
        // positions beyond the shared-memory columns but below P: straight to global (P > PPAD only)
        if (P > Pm && ok) {
            const uint32_t gs = min(Ls, P), gq = min(Lq, P);
            for (uint32_t g = Pm + i; g < gs; g += 8) atomicAdd(p.seqraw + (size_t)g * 256 + tile[h + 1u + g], 1ull);
            for (uint32_t g = Pm + i; g < gq; g += 8) atomicAdd(qualg + (size_t)g * 256 + tile[pp + 1u + g], 1ull);
        }
    }

    // records that are not completely staged: whole warp, one at a time
    unsigned slow = __ballot_sync(0xffffffffu, valid && !complete && i == 0);
    while (slow) {
        const int src = __ffs(slow) - 1;
        slow &= slow - 1;
        const unsigned long long a = __shfl_sync(0xffffffffu, abs_s, src);
        record_global<C>(p, a, limit, hist, lenh, acc, lane);
    }
}

// ------------------------------------------------------------------------------------------
// decoupled look-back with a wide window: all loads of a 128-entry chunk are in flight together
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long look_back(const unsigned long long* status, uint32_t t, int lane)
{
    unsigned long long excl = 0;
    long long pos = (long long)t - 1;   // nearest predecessor
    for (;;) {
        unsigned long long v[4];
        bool ready;
        do {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const long long idx = pos - 32 * r - lane;
                v[r] = idx >= 0 ? ld_volatile_u64(status + idx) : ST_INC;
            }
            // entries are needed up to the nearest inclusive one
            ready = true;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const unsigned inc = __ballot_sync(0xffffffffu, (v[r] >> 62) == 2);
                const unsigned nil = __ballot_sync(0xffffffffu, (v[r] >> 62) == 0);
                const unsigned below = inc ? ((inc & (0u - inc)) - 1u) : 0xffffffffu;  // lanes nearer than the first inclusive
                if (nil & below) ready = false;
                if (inc || !ready) break;
            }
        } while (!ready);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const unsigned inc = __ballot_sync(0xffffffffu, (v[r] >> 62) == 2);
            const unsigned long long val = v[r] & ST_VAL;
            if (inc) {
                const int first = __ffs(inc) - 1;
                excl += warp_sum_u64(lane <= first ? val : 0ull);
                return excl;
            }
            excl += warp_sum_u64(val);
        }
        pos -= 128;
    }
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(1024, 1) fq_scan_kernel(const ScanParams p)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* lenh = hist + C::HIST_WORDS;
    __shared__ TeamCtl<C::NBUF, C::NUNITS> ctl_all[C::TEAMS];
    __shared__ CtaCtl cta;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int team = tid / C::TT;
    const int ttid = tid - team * C::TT;          // thread within the team
    const int warp = ttid >> 5;                   // warp within the team
    const uint32_t lt_mask = (1u << lane) - 1u;
    TeamCtl<C::NBUF, C::NUNITS>& ctl = ctl_all[team];
    uint8_t* team_mem = smem_raw + C::HIST_WORDS * 4 + C::LENH_WORDS * 4 + team * C::TEAM_BYTES;
    uint16_t* lists = reinterpret_cast<uint16_t*>(team_mem + C::NBUF * C::TILE_PAD);

    unsigned long long limit = NONE64;
    if (p.flags & F_RERUN) {
        limit = p.res->first_bad;                 // written by the first pass, stable during this launch
        if (limit == NONE64) return;
    }
    if ((p.flags & F_CARRY) && p.carry->status != 0) return;   // the stream already failed
    const unsigned long long line_base = (p.flags & F_CARRY) ? p.carry->line_base : p.line_base;

    for (int i = tid; i < C::HIST_WORDS + C::LENH_WORDS; i += 1024) hist[i] = 0;
    if (ttid == 0) {
        for (int b = 0; b < C::NBUF; ++b) mbar_init(&ctl.full[b], 1);
        fence_mbar_init();
        ctl.pass_counter[0] = ctl.pass_counter[1] = 0;
        ctl.flush_iter = -2;
    }
    if (tid == 0) cta.recs_since_flush = 0;
    __syncthreads();

    // static schedule: team gt of n_teams handles tiles gt, gt + n_teams, ...
    const uint32_t n_teams = gridDim.x * C::TEAMS;
    const uint32_t gt = blockIdx.x * C::TEAMS + team;
    uint32_t ntiles_eff = p.ntiles;
    if (limit != NONE64) {
        const unsigned long long lt = limit / C::TILE + 1;   // tiles that start below the limit
        if (lt < ntiles_eff) ntiles_eff = (uint32_t)lt;
    }
    const int K = gt < ntiles_eff ? (int)((ntiles_eff - gt + n_teams - 1) / n_teams) : 0;

    Acc acc = {0, 0, 0, 0};
    const bool want_index = (p.flags & F_INDEX) && !(p.flags & F_RERUN) && p.index != nullptr;

    auto tile_no = [&](int k) -> uint32_t { return gt + (uint32_t)k * n_teams; };
    auto tile_buf = [&](int k) -> uint8_t* { return team_mem + (k % C::NBUF) * C::TILE_PAD; };

    // issue the bulk copy of tile k (one thread); ragged edges are filled by hand in scan()
    auto issue = [&](int k) {
        const unsigned long long ts = (unsigned long long)tile_no(k) * C::TILE;
        const uint32_t data_len = (uint32_t)min((unsigned long long)(C::TILE + HALO), p.n_avail - ts);
        const uint32_t front = (ts || (p.flags & F_FRONT16)) ? FRONT : 0;
        const uint32_t bulk = (front + data_len) & ~15u;
        fence_proxy_async();
        mbar_arrive_expect_tx(&ctl.full[k % C::NBUF], bulk);
        if (bulk) bulk_g2s(tile_buf(k) + FRONT - front, p.data + ts - front, bulk, &ctl.full[k % C::NBUF]);
    };

    if (ttid == 0)
        for (int k = 0; k < K && k < C::NBUF - 1; ++k) issue(k);

    for (int k = -1; k < K; ++k) {
        // ---- keep the TMA ring full: the buffer of tile k-1 is free since the barrier below ----
        if (k >= 0 && ttid == 0 && k + C::NBUF - 1 < K) issue(k + C::NBUF - 1);

        // =====================================================================================
        // scan(k+1)
        // =====================================================================================
        const int kn = k + 1;
        const bool have_next = kn < K;
        uint32_t mask[C::ITERS];
        uint32_t ubase[C::ITERS];
        TileMeta& mn = ctl.meta[kn & 1];
        uint8_t* tile_n = tile_buf(kn);
        uint16_t* list_n = lists + (kn & 1) * C::LIST_CAP;
        bool overflow_n = false;
        uint32_t f_n = 0, own_count_n = 0, total_n = 0;
        if (have_next) {
            const uint32_t t = tile_no(kn);
            const unsigned long long ts = (unsigned long long)t * C::TILE;
            const uint32_t own_len = (uint32_t)min((unsigned long long)C::TILE, p.n_own - ts);
            const uint32_t data_len = (uint32_t)min((unsigned long long)(C::TILE + HALO), p.n_avail - ts);
            const uint32_t front = (ts || (p.flags & F_FRONT16)) ? FRONT : 0;
            const uint32_t span = front + data_len;
            if (span != (uint32_t)C::SM_TILE) {
                // ragged first / last tiles: leading zeros (or the virtual '\n' of a line start),
                // the bytes the 16-byte-granular bulk copy leaves out, and zero fill
                const uint32_t bulk = span & ~15u;
                const uint8_t* src = p.data + ts - front;
                uint8_t* dst = tile_n + FRONT - front;
                const bool virt_nl = ts == 0 && front == 0 && (p.flags & F_LINE_START);
                for (uint32_t i = ttid; i < FRONT - front; i += C::TT) tile_n[i] = (virt_nl && i == FRONT - 1) ? '\n' : 0;
                for (uint32_t i = bulk + ttid; i < span; i += C::TT) dst[i] = src[i];
                for (uint32_t i = FRONT + data_len + ttid; i < (uint32_t)C::SM_TILE; i += C::TT) tile_n[i] = 0;
            }
            mbar_wait(&ctl.full[kn % C::NBUF], (uint32_t)(kn / C::NBUF) & 1u);

            // ---- pass 1: newline masks, per-unit counts --------------------------------------
            uint32_t hib = 0;
            if (span != (uint32_t)C::SM_TILE) team_bar<1, C::TT>(team);   // hand-written bytes visible
#pragma unroll
            for (int it = 0; it < C::ITERS; ++it) {
                const int u = it * C::TW + warp;
                mask[it] = 0;
                if (u < C::NUNITS) {
                    const uint32_t off = (uint32_t)u * UNIT + (uint32_t)lane * 16u;
                    const uint4 v = *reinterpret_cast<const uint4*>(tile_n + FRONT + off);
                    hib |= v.x | v.y | v.z | v.w;
                    const uint32_t mm = nlmask16(v);
                    mask[it] = mm;
                    const uint32_t call = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(mm));
                    uint32_t cown;
                    if (own_len == (uint32_t)C::TILE) {
                        cown = u < C::OWN_UNITS ? call : 0u;
                    } else {
                        const int rem = (int)own_len - (int)off;
                        const uint32_t ownm = rem >= 16 ? 0xFFFFu : (rem > 0 ? ((1u << rem) - 1u) : 0u);
                        cown = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(mm & ownm));
                    }
                    if (lane == 0) {
                        ctl.unit_all[u] = call;
                        ctl.unit_own[u] = cown;
                    }
                }
            }
            const bool na = __any_sync(0xffffffffu, (hib & 0x80808080u) != 0);
            if (ttid == 0) mn.nonascii = 0;
            team_bar<1, C::TT>(team);                      // BAR1: unit counts visible
            if (na && lane == 0) mn.nonascii = 1;          // (read after BAR2)

            // ---- every warp: exclusive prefix over the unit counts ----------------------------
            f_n = tile_n[FRONT - 1] == '\n' ? 1u : 0u;
            uint32_t run = f_n;
            uint32_t cnt[C::UPL], inc[C::UPL], own_sum = 0;
#pragma unroll
            for (int r = 0; r < C::UPL; ++r) {
                const int u = lane + 32 * r;
                cnt[r] = u < C::NUNITS ? ctl.unit_all[u] : 0u;
                own_sum += u < C::NUNITS ? ctl.unit_own[u] : 0u;
                uint32_t x = cnt[r];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                    if (lane >= d) x += y;
                }
                inc[r] = x + run - cnt[r];                  // exclusive prefix (incl. the front entry)
                run += __shfl_sync(0xffffffffu, x, 31);
            }
            total_n = run - f_n;
            own_count_n = __reduce_add_sync(0xffffffffu, own_sum);
#pragma unroll
            for (int it = 0; it < C::ITERS; ++it) {
                const int u = it * C::TW + warp;
                ubase[it] = __shfl_sync(0xffffffffu, inc[(u >> 5) < C::UPL ? (u >> 5) : 0], u & 31);
            }
            overflow_n = f_n + total_n > (uint32_t)C::LIST_CAP;

            // ---- warp 0: meta + publish this tile's count right away -----------------------------
            if (warp == 0) {
                if (lane == 0) {
                    mn.ts = ts;
                    mn.front = f_n;
                    mn.own_count = own_count_n;
                    mn.total_count = total_n;
                    mn.own_len = own_len;
                    if (f_n) list_n[0] = FRONT - 1;
                    if (t == 0)
                        st_volatile_u64(p.tile_status, ST_INC | (line_base + own_count_n));
                    else
                        st_volatile_u64(p.tile_status + t, ST_AGG | own_count_n);
                }
                if (overflow_n) {
                    // dense tile: its index entries are written straight from pass 2, which needs the
                    // line number now -> look back immediately (rare, not overlapped)
                    const unsigned long long excl = t == 0 ? line_base : look_back(p.tile_status, t, lane);
                    if (lane == 0) {
                        mn.base = excl;
                        if (t) st_volatile_u64(p.tile_status + t, ST_INC | (excl + own_count_n));
                    }
                }
            }
            if (overflow_n) team_bar<1, C::TT>(team);

            // ---- pass 2: rank every newline, fill the position list -----------------------------
            const unsigned long long idx_base_n = overflow_n ? mn.base - line_base : 0;
            const unsigned long long off_base_n = p.stream_offset + ts;
#pragma unroll
            for (int it = 0; it < C::ITERS; ++it) {
                const int u = it * C::TW + warp;
                if (u < C::NUNITS) {
                    uint32_t mm = mask[it];
                    const int c = __popc(mm);
                    int pre = 0;
                    for (int lvl = 0;; ++lvl) {
                        const unsigned b = __ballot_sync(0xffffffffu, c > lvl);
                        if (!b) break;
                        pre += __popc(b & lt_mask);
                    }
                    uint32_t rank = ubase[it] + (uint32_t)pre;
                    const uint32_t off = (uint32_t)u * UNIT + (uint32_t)lane * 16u;
                    while (mm) {
                        const uint32_t bit = (uint32_t)__ffs(mm) - 1u;
                        mm &= mm - 1u;
                        if (rank < (uint32_t)C::LIST_CAP) list_n[rank] = (uint16_t)(FRONT + off + bit);
                        if (overflow_n && want_index && rank >= f_n && rank < f_n + own_count_n) {
                            const unsigned long long gi = idx_base_n + (rank - f_n);
                            if (gi < p.index_cap) p.index[gi] = (uint32_t)(off_base_n + off + bit);
                        }
                        ++rank;
                    }
                }
            }
        }
        if (ttid == 0) ctl.pass_counter[kn & 1] = 0;   // used by records(k+1), after the barrier below

        // =====================================================================================
        // look-back(k+1) by warp 0, hidden behind records(k) of the other warps
        // =====================================================================================
        if (have_next && warp == 0 && !overflow_n) {
            const uint32_t t = tile_no(kn);
            const unsigned long long excl = t == 0 ? line_base : look_back(p.tile_status, t, lane);
            if (lane == 0) {
                mn.base = excl;
                if (t) st_volatile_u64(p.tile_status + t, ST_INC | (excl + own_count_n));
            }
        }
        if (have_next && warp == 0 && lane == 0 && tile_no(kn) == p.ntiles - 1 && !(p.flags & F_RERUN)) {
            // (mn.base was written by this very thread)
            p.res->n_lines = mn.base + own_count_n - line_base;
            p.res->line_end = mn.base + own_count_n;
        }

        // =====================================================================================
        // records(k): index copy, then passes handed out dynamically inside the team
        // =====================================================================================
        if (k >= 0) {
            const TileMeta& m = ctl.meta[k & 1];
            const uint8_t* tile = tile_buf(k);
            const uint16_t* list = lists + (k & 1) * C::LIST_CAP;
            const uint32_t f = m.front, own_count = m.own_count;
            const bool overflow = f + m.total_count > (uint32_t)C::LIST_CAP;
            if (want_index && !overflow) {
                const unsigned long long idx_base = m.base - line_base;
                const unsigned long long off_base = p.stream_offset + m.ts;
                for (uint32_t i = ttid; i < own_count; i += C::TT) {
                    const unsigned long long gi = idx_base + i;
                    if (gi < p.index_cap) p.index[gi] = (uint32_t)(off_base + (uint32_t)list[f + i] - FRONT);
                }
            }
            const uint32_t gb = (uint32_t)((m.base - f) & 3ull);
            const uint32_t j0 = (3u - gb) & 3u;
            const uint32_t nown = f + own_count;
            const uint32_t nrec = nown > j0 ? (nown - j0 + 3u) / 4u : 0u;
            const uint32_t npass = (nrec + 3u) / 4u;
            if (ttid == 0 && nrec) {
                // u16 counter halves: whoever pushes the CTA-wide record count over the mark drains
                const uint32_t before = atomicAdd(&cta.recs_since_flush, nrec);
                if (before + nrec >= 24000u) {
                    atomicExch(&cta.recs_since_flush, 0u);
                    ctl.flush_iter = k;
                }
            }
            if (!overflow) {
                for (;;) {
                    uint32_t pass = 0;
                    if (lane == 0) pass = atomicAdd(&ctl.pass_counter[k & 1], 1u);
                    pass = __shfl_sync(0xffffffffu, pass, 0);
                    if (pass >= npass) break;
                    if (m.nonascii)
                        records_pass<C, false>(p, m, tile, list, hist, lenh, limit, pass, acc, lane);
                    else
                        records_pass<C, true>(p, m, tile, list, hist, lenh, limit, pass, acc, lane);
                }
            } else if (warp == C::TW - 1) {
                // dense-newline tile: walk its records one after the other in global memory
                if (j0 < nown) {
                    unsigned long long s = m.ts + (uint32_t)list[j0] + 1u - FRONT;
                    const unsigned long long tend = m.ts + m.own_len;
                    while (s < tend && s < limit) {
                        const unsigned long long e = record_global<C>(p, s, limit, hist, lenh, acc, lane);
                        if (e == NONE64) break;
                        s = e + 1;
                    }
                }
            }
        }
        team_bar<1, C::TT>(team);                          // BAR2: tile k consumed, tile k+1 fully described
        if (ctl.flush_iter == k) flush_hist<C>(hist, p, ttid, C::TT);
    }

    // ---- drain -----------------------------------------------------------------------------
    __syncthreads();
    flush_hist<C>(hist, p, tid, 1024);
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
using Cfg5 = Cfg<5, 2, 16384, 3>;    // P <= 160
using Cfg10 = Cfg<10, 2, 8192, 2>;   // P <= 320 (longer reads: positions >= 320 go to global atomics)

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
