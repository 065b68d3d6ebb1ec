// fq_scan.cu -- the fused sm_100a scan kernel (K1 delimit + K2 per-position histograms).
//
// One pass over the bytes.  Persistent grid, one 1024-thread CTA per SM, split into TEAMS
// independent teams that share one shared-memory histogram.  Every team runs a software
// pipeline over its statically assigned tiles (tile = team + k * n_teams); in iteration k
//
//   control warp   issues the TMA bulk copy of tile k+NBUF-1 (UBLKCP + mbarrier), then resolves
//                  the line number of tile k with a decoupled look-back over the newline counts
//                  the other teams published ONE ITERATION EARLIER (so it never waits for a
//                  team that runs in lock step with this one)
//   scan(k+1)      other warps: 16-byte SWAR newline masks -> per-unit counts -> ranks ->
//                  position list of tile k+1; its count is published for the look-backs to come
//   records(k)     all warps, 8 lanes per record: '@' / '+' / raw-length validation
//                  (src/records.rs:201-247), then each lane walks 4-byte groups of the sequence
//                  and quality lines and bumps hist[byte][position] -- bank = position % 32, and
//                  the (group, byte) rotation makes the 32 lanes of every ATOMS hit 32 banks
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
    static constexpr int SW = TW - 1;                      // warps that scan (warp 0 is the control warp)
    static constexpr int SM_TILE = FRONT + TILE + HALO;
    static constexpr int TILE_PAD = (SM_TILE + 16 + 127) / 128 * 128;
    static constexpr int NUNITS = (TILE + HALO) / UNIT;
    static constexpr int OWN_UNITS = TILE / UNIT;
    static constexpr int ITERS = (NUNITS + SW - 1) / SW;
    static constexpr int UPL = (NUNITS + 31) / 32;         // unit counts per lane
    static constexpr int LIST_CAP = TILE / 4;
    static constexpr int HIST_WORDS = HIST_ROWS * PPAD;    // hist[byte][position], u32 = lo16 seq | hi16 qual
    static constexpr int LENH_WORDS = (PPAD + 2 + 31) / 32 * 32;
    static constexpr int TEAM_BYTES = NBUF * TILE_PAD + 2 * LIST_CAP * 2;
    // word loads of the record pass may run up to PPAD + 8 bytes past a tile buffer: keep them inside
    static constexpr int TAIL_PAD = (PPAD + 8 + 127) / 128 * 128;
    static constexpr int TOTAL = HIST_WORDS * 4 + LENH_WORDS * 4 + TEAMS * TEAM_BYTES + TAIL_PAD;
    static constexpr uint32_t ROW_BYTES = PPAD * 4;
    static_assert(SM_TILE < 65536, "list entries are u16");
    static_assert(UPL <= 3, "unit counts: <= 96 units");
};

struct TileMeta {
    unsigned long long base;   // stream-global exclusive line count at the tile start
    unsigned long long ts;     // buffer-relative offset of the tile
    uint32_t front;            // 1 if the byte before the tile is (or acts as) '\n'
    uint32_t own_count;        // '\n' in the owned range
    uint32_t total_count;      // '\n' staged (owned + halo)
    uint32_t own_len;
};

template <int NBUF, int NUNITS>
struct TeamCtl {
    unsigned long long full[NBUF];   // mbarriers: tile bytes landed
    TileMeta meta[2];
    uint32_t unit_all[2][NUNITS + 2];
    uint32_t unit_own[NUNITS + 2];
    uint32_t pass_counter[2];   // records(k) hands out passes from pass_counter[k & 1]
    int nonascii_iter[2];       // == k + 1 when tile k holds a byte >= 0x80
    int flush_iter;             // iteration at whose end this team drains the shared counters
    uint32_t pad;
};

struct CtaCtl {
    uint32_t recs_since_flush;
};

template <int NTHREADS>
__device__ __forceinline__ void team_bar(int team)
{
    asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(NTHREADS) : "memory");
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
    asm volatile("red.shared.add.u32 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------
// lock-free drain of the u16-pair counters: atomicExch leaves concurrent increments of the other
// team intact, so a team may flush whenever the CTA-wide record counter says a half could
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
// different residues mod 32 = 32 different banks of hist[byte][position].
// ------------------------------------------------------------------------------------------
struct RoundCtx {
    uint32_t as0, aq0;     // shared addresses of the aligned words holding position 4i of seq / qual
    uint32_t shs, shq;     // funnel shifts that realign them
    uint32_t rot;          // 8 * sub
    uint32_t hk[4];        // shared address of hist[0][pk[k]]
    uint32_t pk[4];        // position visited by the k-th bump in round 0
    uint32_t ns, nq;       // bytes of seq / qual that have a shared-memory column
    uint32_t nmax_w, nmin_w;
    unsigned long long *gseq, *gqual;   // global rows (non-ASCII bytes only)
};

template <class C, bool ASCII, int T>
struct Rounds {
    static __device__ __forceinline__ void run(const RoundCtx& c)
    {
        if (32u * T >= c.nmax_w) return;                                  // warp-uniform
        const uint32_t s0 = lds32<32 * T>(c.as0), s1 = lds32<32 * T + 4>(c.as0);
        const uint32_t q0 = lds32<32 * T>(c.aq0), q1 = lds32<32 * T + 4>(c.aq0);
        uint32_t vs = __funnelshift_r(s0, s1, c.shs);
        uint32_t vq = __funnelshift_r(q0, q1, c.shq);
        vs = __funnelshift_r(vs, vs, c.rot);                              // byte k = byte (k+sub)&3 of the group
        vq = __funnelshift_r(vq, vq, c.rot);
        if (ASCII && 32u * (T + 1) <= c.nmin_w) {                         // every lane's group lies inside both lines
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                red_add<128 * T>(c.hk[k] + __byte_perm(vs, 0, 0x4440 + k) * C::ROW_BYTES, 1u);
                red_add<128 * T>(c.hk[k] + __byte_perm(vq, 0, 0x4440 + k) * C::ROW_BYTES, 0x10000u);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t pos = c.pk[k] + 32u * T;
                const uint32_t bs = __byte_perm(vs, 0, 0x4440 + k), bq = __byte_perm(vq, 0, 0x4440 + k);
                if (pos < c.ns) {
                    if (ASCII || bs < (uint32_t)HIST_ROWS)
                        red_add<128 * T>(c.hk[k] + bs * C::ROW_BYTES, 1u);
                    else
                        atomicAdd(c.gseq + (size_t)pos * 256 + bs, 1ull);
                }
                if (pos < c.nq) {
                    if (ASCII || bq < (uint32_t)HIST_ROWS)
                        red_add<128 * T>(c.hk[k] + bq * C::ROW_BYTES, 0x10000u);
                    else
                        atomicAdd(c.gqual + (size_t)pos * 256 + bq, 1ull);
                }
            }
        }
        Rounds<C, ASCII, T + 1>::run(c);
    }
};
template <class C, bool ASCII>
struct Rounds<C, ASCII, C::NCHUNK> {
    static __device__ __forceinline__ void run(const RoundCtx&) {}
};

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
        RoundCtx c;
        c.ns = min(Ls, Pm);
        c.nq = min(Lq, Pm);
        c.nmax_w = __reduce_max_sync(0xffffffffu, max(c.ns, c.nq));
        c.nmin_w = __reduce_min_sync(0xffffffffu, min(c.ns, c.nq));
        const uint32_t sa = h + 1u + 4u * i;                // shared offset of position 4i of the sequence line
        const uint32_t qa = pp + 1u + 4u * i;
        const uint32_t tile_s = smem_u32(tile);
        c.as0 = tile_s + (sa & ~3u);
        c.aq0 = tile_s + (qa & ~3u);
        c.shs = (sa & 3u) * 8u;
        c.shq = (qa & 3u) * 8u;
        c.rot = 8u * sub;
        c.gseq = p.seqraw;
        c.gqual = p.stats + stats_qual_off(P);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            c.pk[k] = 4u * i + (((uint32_t)k + sub) & 3u);
            c.hk[k] = smem_u32(hist) + 4u * c.pk[k];
        }
        Rounds<C, ASCII, 0>::run(c);
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
        const unsigned long long a = __shfl_sync(0xffffffffu, abs_s, src);
        record_global<C>(p, a, limit, hist, lenh, lane);
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

// exclusive prefix of the unit counts for unit u (u warp-uniform): f + sum_{u' < u} cnt[u']
template <int UPL>
__device__ __forceinline__ uint32_t unit_prefix(const uint32_t (&cnt)[UPL], int u, int lane, uint32_t f)
{
    uint32_t x = 0;
#pragma unroll
    for (int r = 0; r < UPL; ++r) x += (lane + 32 * r < u) ? cnt[r] : 0u;
    return f + __reduce_add_sync(0xffffffffu, x);
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
    __shared__ TeamCtl<C::NBUF, C::NUNITS> ctl_all[C::TEAMS];
    __shared__ CtaCtl cta;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int team = tid / C::TT;
    const int ttid = tid - team * C::TT;          // thread within the team
    const int warp = ttid >> 5;                   // warp within the team; 0 = control warp
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
        ctl.nonascii_iter[0] = ctl.nonascii_iter[1] = -5;
        ctl.flush_iter = -5;
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

    // issue the bulk copy of tile k (one thread); ragged edges are filled by hand before the scan
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
        const int kn = k + 1;
        const bool have_next = kn < K;
        TileMeta& mk = ctl.meta[k & 1];
        TileMeta& mn = ctl.meta[kn & 1];
        uint8_t* tile_n = tile_buf(kn);
        uint16_t* list_n = lists + (kn & 1) * C::LIST_CAP;
        uint32_t* unit_all_n = ctl.unit_all[kn & 1];

        // tile k+1 geometry
        uint32_t tn = 0, own_len_n = 0;
        unsigned long long ts_n = 0;
        bool ragged_n = false;
        if (have_next) {
            tn = tile_no(kn);
            ts_n = (unsigned long long)tn * C::TILE;
            own_len_n = (uint32_t)min((unsigned long long)C::TILE, p.n_own - ts_n);
            const uint32_t data_len = (uint32_t)min((unsigned long long)(C::TILE + HALO), p.n_avail - ts_n);
            const uint32_t front = (ts_n || (p.flags & F_FRONT16)) ? FRONT : 0;
            const uint32_t span = front + data_len;
            ragged_n = span != (uint32_t)C::SM_TILE;
            if (ragged_n) {
                // ragged first / last tiles: leading zeros (or the virtual '\n' of a line start), the
                // bytes the 16-byte-granular bulk copy leaves out, and zero fill -- by the whole team
                const uint32_t bulk = span & ~15u;
                const uint8_t* src = p.data + ts_n - front;
                uint8_t* dst = tile_n + FRONT - front;
                const bool virt_nl = ts_n == 0 && front == 0 && (p.flags & F_LINE_START);
                for (uint32_t i = ttid; i < FRONT - front; i += C::TT) tile_n[i] = (virt_nl && i == FRONT - 1) ? '\n' : 0;
                for (uint32_t i = bulk + ttid; i < span; i += C::TT) dst[i] = src[i];
                for (uint32_t i = FRONT + data_len + ttid; i < (uint32_t)C::SM_TILE; i += C::TT) tile_n[i] = 0;
                team_bar<C::TT>(team);
            }
        }

        uint32_t mask[C::ITERS];
        if (warp == 0) {
            // =================================================================================
            // control warp: keep the TMA ring full, resolve the line number of tile k
            // =================================================================================
            if (lane == 0 && k >= 0 && k + C::NBUF - 1 < K) issue(k + C::NBUF - 1);   // buffer of tile k-1: free
            if (k >= 0) {
                const uint32_t t = tile_no(k);
                const unsigned long long excl = t == 0 ? line_base : look_back(p.tile_status, t, lane);
                if (lane == 0) {
                    mk.base = excl;
                    if (t) st_volatile_u64(p.tile_status + t, ST_INC | (excl + mk.own_count));
                    if (t == p.ntiles - 1 && !(p.flags & F_RERUN)) {
                        p.res->n_lines = excl + mk.own_count - line_base;
                        p.res->line_end = excl + mk.own_count;
                    }
                }
            }
        } else if (have_next) {
            // =================================================================================
            // scan(k+1), pass 1: newline masks, per-unit counts
            // =================================================================================
            mbar_wait(&ctl.full[kn % C::NBUF], (uint32_t)(kn / C::NBUF) & 1u);
            uint32_t hib = 0;
#pragma unroll
            for (int it = 0; it < C::ITERS; ++it) {
                const int u = it * C::SW + (warp - 1);
                mask[it] = 0;
                if (u < C::NUNITS) {
                    const uint32_t off = (uint32_t)u * UNIT + (uint32_t)lane * 16u;
                    const uint4 v = *reinterpret_cast<const uint4*>(tile_n + FRONT + off);
                    hib |= v.x | v.y | v.z | v.w;
                    const uint32_t mm = nlmask16(v);
                    mask[it] = mm;
                    const uint32_t call = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(mm));
                    uint32_t cown;
                    if (own_len_n == (uint32_t)C::TILE) {
                        cown = u < C::OWN_UNITS ? call : 0u;
                    } else {
                        const int rem = (int)own_len_n - (int)off;
                        const uint32_t ownm = rem >= 16 ? 0xFFFFu : (rem > 0 ? ((1u << rem) - 1u) : 0u);
                        cown = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(mm & ownm));
                    }
                    if (lane == 0) {
                        unit_all_n[u] = call;
                        ctl.unit_own[u] = cown;
                    }
                }
            }
            if (__any_sync(0xffffffffu, (hib & 0x80808080u) != 0) && lane == 0) ctl.nonascii_iter[kn & 1] = kn + 1;
        }
        team_bar<C::TT>(team);   // BAR1: unit counts of tile k+1 and the line number of tile k are visible

        if (have_next) {
            if (warp == 0) {
                // ---- control warp: describe tile k+1 and publish its count for later look-backs ----
                mbar_wait(&ctl.full[kn % C::NBUF], (uint32_t)(kn / C::NBUF) & 1u);   // (already complete)
                const uint32_t f_n = tile_n[FRONT - 1] == '\n' ? 1u : 0u;
                uint32_t a = 0, o = 0;
#pragma unroll
                for (int r = 0; r < C::UPL; ++r) {
                    const int u = lane + 32 * r;
                    a += u < C::NUNITS ? unit_all_n[u] : 0u;
                    o += u < C::NUNITS ? ctl.unit_own[u] : 0u;
                }
                const uint32_t total_n = __reduce_add_sync(0xffffffffu, a);
                const uint32_t own_count_n = __reduce_add_sync(0xffffffffu, o);
                if (lane == 0) {
                    mn.ts = ts_n;
                    mn.front = f_n;
                    mn.own_count = own_count_n;
                    mn.total_count = total_n;
                    mn.own_len = own_len_n;
                    if (f_n) list_n[0] = FRONT - 1;
                    if (tn == 0)
                        st_volatile_u64(p.tile_status, ST_INC | (line_base + own_count_n));
                    else
                        st_volatile_u64(p.tile_status + tn, ST_AGG | own_count_n);
                    ctl.pass_counter[kn & 1] = 0;   // used by records(k+1), after BAR2
                }
            } else {
                // ---- scan(k+1), pass 2: rank every newline, fill the position list ---------------
                const uint32_t f_n = tile_n[FRONT - 1] == '\n' ? 1u : 0u;
                uint32_t cnt[C::UPL];
#pragma unroll
                for (int r = 0; r < C::UPL; ++r) cnt[r] = (lane + 32 * r) < C::NUNITS ? unit_all_n[lane + 32 * r] : 0u;
#pragma unroll
                for (int it = 0; it < C::ITERS; ++it) {
                    const int u = it * C::SW + (warp - 1);
                    if (u < C::NUNITS) {
                        uint32_t mm = mask[it];
                        const int c = __popc(mm);
                        int pre = 0;
                        for (int lvl = 0;; ++lvl) {
                            const unsigned b = __ballot_sync(0xffffffffu, c > lvl);
                            if (!b) break;
                            pre += __popc(b & lt_mask);
                        }
                        uint32_t rank = unit_prefix<C::UPL>(cnt, u, lane, f_n) + (uint32_t)pre;
                        const uint32_t pos0 = FRONT + (uint32_t)u * UNIT + (uint32_t)lane * 16u;
                        while (mm) {
                            const uint32_t bit = (uint32_t)__ffs(mm) - 1u;
                            mm &= mm - 1u;
                            if (rank < (uint32_t)C::LIST_CAP) list_n[rank] = (uint16_t)(pos0 + bit);
                            ++rank;
                        }
                    }
                }
            }
        }

        // =====================================================================================
        // records(k): index copy, then passes handed out dynamically inside the team
        // =====================================================================================
        if (k >= 0) {
            const uint8_t* tile = tile_buf(k);
            const uint16_t* list = lists + (k & 1) * C::LIST_CAP;
            const uint32_t f = mk.front, own_count = mk.own_count;
            const bool overflow = f + mk.total_count > (uint32_t)C::LIST_CAP;
            const unsigned long long idx_base = mk.base - line_base;     // buffer-local number of the first own line
            const unsigned long long off_base = p.stream_offset + mk.ts;
            const uint32_t gb = (uint32_t)((mk.base - f) & 3ull);
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
                if (want_index) {
                    for (uint32_t i = ttid; i < own_count; i += C::TT) {
                        const unsigned long long gi = idx_base + i;
                        if (gi < p.index_cap) p.index[gi] = (uint32_t)(off_base + (uint32_t)list[f + i] - FRONT);
                    }
                }
                const bool nonascii = ctl.nonascii_iter[k & 1] == k + 1;
                for (;;) {
                    uint32_t pass = 0;
                    if (lane == 0) pass = atomicAdd(&ctl.pass_counter[k & 1], 1u);
                    pass = __shfl_sync(0xffffffffu, pass, 0);
                    if (pass >= npass) break;
                    if (nonascii)
                        records_pass<C, false>(p, mk, tile, list, hist, lenh, limit, pass, acc, lane);
                    else
                        records_pass<C, true>(p, mk, tile, list, hist, lenh, limit, pass, acc, lane);
                }
            } else {
                // dense-newline tile (more line ends than the list holds; never a healthy FASTQ):
                // redo the ranking straight into the index, then walk the records one by one
                if (want_index && warp > 0) {
                    const uint32_t* unit_all_k = ctl.unit_all[k & 1];
                    uint32_t cnt[C::UPL];
#pragma unroll
                    for (int r = 0; r < C::UPL; ++r) cnt[r] = (lane + 32 * r) < C::NUNITS ? unit_all_k[lane + 32 * r] : 0u;
                    for (int u = warp - 1; u < C::NUNITS; u += C::SW) {
                        const uint32_t off = (uint32_t)u * UNIT + (uint32_t)lane * 16u;
                        uint32_t mm = nlmask16(*reinterpret_cast<const uint4*>(tile + FRONT + off));
                        const int c = __popc(mm);
                        int pre = 0;
                        for (int lvl = 0;; ++lvl) {
                            const unsigned b = __ballot_sync(0xffffffffu, c > lvl);
                            if (!b) break;
                            pre += __popc(b & lt_mask);
                        }
                        uint32_t rank = unit_prefix<C::UPL>(cnt, u, lane, 0u) + (uint32_t)pre;   // rank among own+halo
                        while (mm) {
                            const uint32_t bit = (uint32_t)__ffs(mm) - 1u;
                            mm &= mm - 1u;
                            if (rank < own_count && idx_base + rank < p.index_cap)
                                p.index[idx_base + rank] = (uint32_t)(off_base + off + bit);
                            ++rank;
                        }
                    }
                }
                if (warp == 0 && j0 < nown) {
                    unsigned long long s = mk.ts + (uint32_t)list[j0] + 1u - FRONT;   // j0 < 4 <= LIST_CAP: stored
                    const unsigned long long tend = mk.ts + mk.own_len;
                    while (s < tend && s < limit) {
                        const unsigned long long e = record_global<C>(p, s, limit, hist, lenh, lane);
                        if (e == NONE64) break;
                        s = e + 1;
                    }
                }
            }
        }
        team_bar<C::TT>(team);   // BAR2: tile k consumed, tile k+1 described
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
