// Barcode stage, packed kernels (included from kernels_fast.cuh; structs FastDev / FastGroup live there).
//
//   k_context       per (window, set) task: the two shared-context columns F and G and the base codes of the region,
//                   packed one u32 per DP row (W space):  profile-row offset of the base code (bits 4..9) |
//                   F (11 bits, from bit 10) | G (11 bits, from bit 21).
//   k_barcode_fast  lane = window x two barcodes, profile from shared memory (any set size).
// (textually included inside namespace qcb by kernels_fast.cuh)

constexpr int kRowTile = 32;                       // tasks per row-info tile

// Row-info word.  The base code is stored as the byte offset of its profile row inside a pair's profile block:
// code * kProfRowBytes = code * 144 = (code << 4) | (code << 7) for code < 8, so `block address | (info & kRowCodeMask)`
// is the row's shared-memory address (blocks are 1 KB aligned) -- one LOP3 instead of a multiply-add chain.
constexpr uint32_t kRowCodeMask = 0x3f0u;
constexpr int kRowFShift = 10, kRowGShift = 21;
constexpr uint32_t kRowFMask = 0x7ffu;
__device__ __forceinline__ uint32_t rowinfo_code(int code) { return ((uint32_t)code << 4) | ((uint32_t)code << 7); }

__device__ __forceinline__ long long rowinfo_base(long long task)
{
    return (task >> 5) * (long long)(kRows * kRowTile) + (task & 31);
}

// Task order: epi2me t = window; dual t = k * n_windows + w (all first-set tasks, then all second-set tasks).
__device__ __forceinline__ void decode_task(long long t, long long n_windows, int dual, long long &w, int &k)
{
    if (dual && t >= n_windows) { w = t - n_windows; k = 1; } else { w = t; k = 0; }
}

// Task order of the barcode stage.  A tile of 32 tasks runs as many rows as its longest region, so tasks are bucketed by
// region length first: short regions (extract_barcode_region: barcode + 2 x extension + 1 rows) from slot 0 upwards,
// long ones (the full-window branch, or a Python slice that wraps around) from the last slot downwards.  Slots are
// handed out with warp-aggregated atomics; the order inside a bucket is irrelevant (every task writes its own scores).
__global__ void k_task_order(DevTables t, const WindowSel *__restrict__ sel, long long n_windows, int dual, int short_rows,
                             unsigned int *__restrict__ counters, uint32_t *__restrict__ perm)
{
    const long long n_tasks = dual ? 2 * n_windows : n_windows;
    const long long task = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = task < n_tasks;
    long long w; int k;
    decode_task(valid ? task : 0, n_windows, dual, w, k);
    const WindowSel s = sel[w];
    int n = k ? s.hi1 - s.lo1 : s.hi0 - s.lo0;
    if (t.group[s.layout * 2 + k] < 0 || n < 0) n = 0;
    const bool is_long = n >= short_rows;
    const unsigned lane = threadIdx.x & 31;
    const unsigned m_short = __ballot_sync(0xffffffffu, valid && !is_long), m_long = __ballot_sync(0xffffffffu, valid && is_long);
    unsigned base_short = 0, base_long = 0;
    if (lane == 0) {
        if (m_short) base_short = atomicAdd(&counters[0], (unsigned)__popc(m_short));
        if (m_long) base_long = atomicAdd(&counters[1], (unsigned)__popc(m_long));
    }
    base_short = __shfl_sync(0xffffffffu, base_short, 0);
    base_long = __shfl_sync(0xffffffffu, base_long, 0);
    if (!valid) return;
    const unsigned below = (1u << lane) - 1u;
    const long long slot = is_long ? n_tasks - 1 - (long long)(base_long + __popc(m_long & below))
                                   : (long long)(base_short + __popc(m_short & below));
    perm[slot] = (uint32_t)task;
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Shared-context columns.  F[i] = H[i][u] + (i+u) g for the forward DP of region rows vs the shared prefix;
// G[i] = H'[n-i][d] + (n-i+d) g for the DP of the reversed region vs the reversed shared suffix, whose left border
// is 0 except -g at its last row (node (0, m) is not a valid end).  See DESIGN.md section 4.
//
// Packed: lane = one task, low u16 half = the F problem, high half = the G problem, both right-aligned in NCOL
// registers (dead columns in front keep the border value).  Step s feeds region row s to F and row n-s+1 to G.
// PAIR = true: one table word per column holds both shifted scores, ctx_tab[grp][code F][code G][NCOL], so a step costs
// NCOL / 4 LDS.128 and a cell is one add and one VIMNMX3.U16x2; PAIR = false (too many groups for that table):
// ctx_tab[grp][F | G][code][NCOL], two lookups and a three-input add per cell.
// Output: row r of the packed row info needs F[r] (step r) and G[r] (step n - r).  The value that comes first waits in
// shared memory, indexed by its step; the step that brings the second one assembles the word and stores it -- every row
// is written once, nothing is read back from global memory.
constexpr int kCtxWarps = 4;
constexpr int kCtxEarlyRows = kRows / 2 + 1;                  // steps s with 2 s < n, plus slot 0 (the constant border values)

template <int NCOL, bool PAIR>
__global__ void __launch_bounds__(kCtxWarps * 32)
k_context(FastDev f, DevTables t, const uint32_t *__restrict__ ctx_tab, const uint8_t *__restrict__ codes, int stride,
          const int32_t *__restrict__ wlen, long long n_windows, const WindowSel *__restrict__ sel, int dual, const uint32_t *__restrict__ perm,
          uint32_t *__restrict__ rowinfo, int4 *__restrict__ taskmeta)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nc = f.n_codes, g = f.gap;
    const int grp_words = (PAIR ? nc * nc : 2 * nc) * NCOL;
    const int tab_words = f.n_groups * grp_words;
    uint32_t *s_tab = (uint32_t *)smem;
    uint32_t *s_early = s_tab + tab_words + warp * (kCtxEarlyRows * kRowTile);
    uint8_t *s_code = (uint8_t *)(s_tab + tab_words + kCtxWarps * (kCtxEarlyRows * kRowTile)) + (size_t)warp * (kRows * kRowTile);
    for (int i = threadIdx.x; i < tab_words; i += blockDim.x) s_tab[i] = ctx_tab[i];
    __syncthreads();

    const long long n_tasks = dual ? 2 * n_windows : n_windows;
    const long long n_tiles = (n_tasks + kRowTile - 1) / kRowTile;
    constexpr uint32_t kLowMask = (1u << kRowGShift) - 1u;     // code | F of a row-info word
    for (long long tile = (long long)blockIdx.x * kCtxWarps + warp; tile < n_tiles; tile += (long long)gridDim.x * kCtxWarps) {
        const long long slot = tile * kRowTile + lane;          // position in the bucketed task order (k_task_order)
        const bool valid = slot < n_tasks;
        const long long task = valid ? (long long)perm[slot] : 0;
        long long w; int k;
        decode_task(task, n_windows, dual, w, k);
        const WindowSel s = sel[w];
        const int lo = k ? s.lo1 : s.lo0, hi = k ? s.hi1 : s.hi0;
        const int grp = t.group[s.layout * 2 + k];
        int n = hi - lo;
        if (!valid || grp < 0 || n < 0) n = 0;
        const FastGroup G = f.groups[grp < 0 ? 0 : grp];
        const int u = G.u, d = G.d;
        __syncwarp();
        {   // the lane's window: barcode-matrix code = high nibble of the packed code byte.  3' windows (odd w) are stored
            // unreversed by k_map_codes: base p of the stored window is position wl - 1 - p of the oriented one
            const uint4 *src = (const uint4 *)(codes + w * stride);
            const int wl = wlen[w >> 1];
            const bool rev = (w & 1) != 0;
            const int first = n > 0 ? (rev ? wl - (lo + n) : lo) : 0, last = n > 0 ? (rev ? wl - lo : lo + n) : 0;
            for (int ch = max(first, 0) >> 4; ch * 16 < last; ++ch) {
                const uint4 v = src[ch];
                const uint32_t words[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int b = 0; b < 16; ++b) {
                    const int p = ch * 16 + b;
                    const int pos = rev ? wl - 1 - p : p;                // position in the oriented window
                    if (pos >= 0 && pos < kRows - 1) s_code[(pos + 1) * kRowTile + lane] = (uint8_t)((words[b >> 2] >> ((b & 3) * 8 + 4)) & 15u);
                }
            }
        }
        // slot 0 of the early values: F[0] (row 0, no base code) and G[n] (row 0 of the reversed problem)
        s_early[lane] = ((uint32_t)(u * g) << kRowFShift) | ((uint32_t)(d * g) << kRowGShift);
        if (n <= 0) s_code[kRowTile + lane] = 0;                     // idle lanes read row 1 in the step loop
        __syncwarp();
        const int nmax = __reduce_max_sync(0xffffffffu, n);
        // shared-memory byte addresses (explicit shared-space loads: no generic-address arithmetic in the step loop)
        const uint32_t tab_addr = (uint32_t)__cvta_generic_to_shared(s_tab) + (uint32_t)(grp < 0 ? 0 : grp) * grp_words * 4u;
        const uint32_t early_addr = (uint32_t)__cvta_generic_to_shared(s_early) + lane * 4u;
        const uint32_t code_addr = (uint32_t)__cvta_generic_to_shared(s_code) + lane;
        uint32_t Wc[NCOL];
#pragma unroll
        for (int c = 0; c < NCOL; ++c) {
            const int jf = c - (NCOL - u) + 1, jg = c - (NCOL - d) + 1;
            Wc[c] = (uint32_t)(max(jf, 0) * g) | ((uint32_t)(max(jg, 0) * g) << 16);
        }
        uint32_t *out = rowinfo + tile * (long long)(kRows * kRowTile) + lane;
        uint32_t *out_hi = out + (long long)((n + 1) >> 1) * kRowTile;      // row st of the first step with 2 st >= n
        uint32_t *out_lo = out + (long long)(n >> 1) * kRowTile;            // row n - st of that step
        const uint32_t gdup = dup16((uint32_t)g);
        uint32_t border = 0;
        int rup = INT32_MIN / 2;
        // a lane past its own last row keeps stepping with its last row's codes: whatever it computes is never stored
        // (idle lanes, n == 0, may carry any lo: they read row 1, zeroed above)
        uint32_t pf = code_addr + (uint32_t)((n > 0 ? lo : 0) + 1) * kRowTile;
        uint32_t pg = code_addr + (uint32_t)(n > 0 ? lo + n : 1) * kRowTile;
        // software pipeline: the table words of step st + 1 and the base codes of step st + 2 are fetched while step st's
        // max chain runs (the chain, not the loads, is then the critical path of a step)
        auto fetch_words = [&](uint32_t (&e)[NCOL], uint32_t cf, uint32_t cg) {
            if (PAIR) {
                const uint32_t pe = tab_addr + (cf * nc + cg) * (NCOL * 4);
#pragma unroll
                for (int c = 0; c < NCOL; c += 4) {
                    const uint4 v = lds128(pe + c * 4);
                    e[c] = v.x; e[c + 1] = v.y; e[c + 2] = v.z; e[c + 3] = v.w;
                }
            } else {
                const uint32_t pfa = tab_addr + cf * (NCOL * 4), pga = tab_addr + (nc + cg) * (NCOL * 4);
#pragma unroll
                for (int c = 0; c < NCOL; c += 4) {
                    const uint4 ef = lds128(pfa + c * 4), eg = lds128(pga + c * 4);
                    e[c] = ef.x + eg.x; e[c + 1] = ef.y + eg.y; e[c + 2] = ef.z + eg.z; e[c + 3] = ef.w + eg.w;
                }
            }
        };
        uint32_t cf_cur = lds_u8(pf), cfn, cgn;
        uint32_t en[NCOL];
        fetch_words(en, cf_cur, lds_u8(pg));
        if (1 < n) { pf += kRowTile; pg -= kRowTile; }
        cfn = lds_u8(pf); cgn = lds_u8(pg);
        if (2 < n) { pf += kRowTile; pg -= kRowTile; }
#pragma unroll 2
        for (int st = 1; st <= nmax; ++st) {
            const uint32_t cf = cf_cur;
            uint32_t diag = border;
            border += gdup;
            uint32_t left = border - (st == n ? ((uint32_t)g << 16) : 0u);      // G's left border at its last row is -g
            uint32_t tt[NCOL];
#pragma unroll
            for (int c = 0; c < NCOL; ++c) tt[c] = (c == 0 ? diag : Wc[c - 1]) + en[c];
            fetch_words(en, cfn, cgn);
            cf_cur = cfn;
            cfn = lds_u8(pf); cgn = lds_u8(pg);
            if (st + 2 < n) { pf += kRowTile; pg -= kRowTile; }
#pragma unroll
            for (int c = 0; c < NCOL; ++c) {
                left = __vimax3_u16x2(tt[c], Wc[c], left);
                Wc[c] = left;
            }
            if (st <= n) {
                // this step's values: F[st] (with the base code of row st, as its profile-row offset code * 144) and
                // G[n - st].  While 2 st <= n both wait in shared memory for their partners; from 2 st >= n on the
                // partners are there (for 2 st == n: the word just written) and two finished rows go out.
                const uint32_t word = cf * 144u + ((left & 0xffffu) << kRowFShift) + ((left >> 16) << kRowGShift);
                if (2 * st <= n) sts32(early_addr + (uint32_t)st * (kRowTile * 4), word);
                if (2 * st >= n) {
                    const uint32_t e = lds32(early_addr + (uint32_t)(n - st) * (kRowTile * 4));   // code | F[n - st], G[st]
                    *out_hi = (word & kLowMask) | (e & ~kLowMask);
                    *out_lo = (e & kLowMask) | (word & ~kLowMask);
                    out_hi += kRowTile;
                    out_lo -= kRowTile;
                }
                if (st == n) {
#pragma unroll
                    for (int c = 0; c < NCOL; ++c) {
                        const int jf = c - (NCOL - u) + 1;
                        if (jf >= 1) rup = max(rup, (int)(Wc[c] & 0xffffu) - (n + jf) * g);
                    }
                }
            }
        }
        __syncwarp();
        if (valid) taskmeta[slot] = make_int4(n, grp, rup, (int)task);
    }
}

// Result of one (window, barcode pair): score = max(R over prefix columns, R over core columns, join with G).
__device__ __forceinline__ void store_pair_scores(const uint32_t (&Wc)[kCore], uint32_t acc, const FastGroup &G, int n, int g, int rup,
                                                  int pr, int32_t *dst)
{
    const int CB = (G.u + kCore) * g;
    uint32_t rm = 0;
#pragma unroll
    for (int c = 0; c < kCore - 1; ++c)
        rm = __viaddmax_u16x2(Wc[c], dup16((uint32_t)(CB - (G.u + max(0, c + 1 - G.pad)) * g)), rm);
    const int bias_r = n * g + CB, bias_j = (n + G.tlen) * g;
    const int s0 = max(max((int)(rm & 0xffffu) - bias_r, (int)(acc & 0xffffu) - bias_j), rup);
    const int s1 = max(max((int)(rm >> 16) - bias_r, (int)(acc >> 16) - bias_j), rup);
    // one 8-byte store: a window's score slots are padded to a multiple of four (api.cu), so slot 2 pr + 1 exists even
    // when the group has an odd number of barcodes (it is never read then)
    *(int2 *)(dst + 2 * pr) = make_int2(s0, s1);
}

// ---- bulk asynchronous copies (TMA's 1-D form: cp.async.bulk, completion counted on an mbarrier) --------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Lane = window x two barcodes (u16 halves); warp = 32 windows x one barcode pair; profile rows via multicast LDS.
// The kernel runs at the issue limit of its instruction mix (add + VIMNMX3 per cell pair, ~2.7 warp-instructions per
// clock per SM), so the row loop is written to need as few instructions as possible around the 48 essential ones:
// one LOP3 for the profile address (see kRowCodeMask), a pointer-compare loop, diagonal terms issued one column ahead
// of the in-place max chain so that no register copies are needed.
// Row-info tiles are double buffered: while the warps work on tile k, warp 0 has the tile the CTA takes next brought
// into the other buffer by one bulk asynchronous copy (cp.async.bulk, completion on an mbarrier) and publishes its slot
// metadata next to it, so a tile switch costs one __syncthreads() and no exposed global-memory latency.
// CTA shape, measured on B200 (profiles/r02_barcode_shapes.txt; PBC096 / NBD104 / dual, ms per 262 144 reads):
// 8 warps x 2 CTAs at <= 128 registers 10.52 / 1.32 / 4.96; 12 x 1 at 139 registers 10.65 / 1.30 / 5.05; 16 x 1 at 127
// 10.55 / 1.32 / 5.11; 12 x 2 at 80 registers 10.86 / 1.37 / 5.24 (ptxas pays for the 80-register cap with 5 - 9 register
// copies per row).  The kernel sits on the shared-memory gather limit either way; fewer, fatter warps win by a little.
#ifndef QCB_BC_MAXWARPS
#define QCB_BC_MAXWARPS 8           // warps per CTA (upper bound; the launch picks a divisor-friendly count <= this)
#endif
#ifndef QCB_BC_MINBLOCKS
#define QCB_BC_MINBLOCKS 2          // CTAs per SM the register allocation is sized for
#endif
constexpr int kBarcodeMaxWarps = QCB_BC_MAXWARPS;

constexpr int kMaxTilesPerIter = 2;
struct BarcodeTileSlot {            // per row-tile buffer (one group of tiles_per_iter consecutive tiles), written by warp 0
    int4 meta[kMaxTilesPerIter][kRowTile];   // taskmeta of the tiles' 32 slots
    long long group;                // tile-group index, -1 = no more groups for this CTA
    unsigned long long bar;         // mbarrier the bulk copies complete on
    int nmax[kMaxTilesPerIter];     // longest region of each tile, -1 = the tile is not taken by this launch
};
struct BarcodeStage {               // warp 0's look-ahead: the taskmeta of the next candidate group, copied asynchronously
    int4 meta[kMaxTilesPerIter][kRowTile];
    long long cand;                 // the next group index to look at
};
constexpr size_t kBarcodeSlotBytes = 2 * sizeof(BarcodeTileSlot) + sizeof(BarcodeStage);

__global__ void __launch_bounds__(kBarcodeMaxWarps * 32, QCB_BC_MINBLOCKS)
k_barcode_fast(FastDev f, long long n_windows, int dual, int bmax0, int bslots, int rows_min, int rows_cap, int one_set,
               int tiles_per_iter, int smem_profile_bytes, const uint32_t *__restrict__ rowinfo, const int4 *__restrict__ taskmeta,
               int32_t *__restrict__ bc_score, const unsigned int *__restrict__ bucket_counts, int pass)
{
    // rows_cap = DP rows (0..n) one shared-memory row tile of this launch holds; a tile is taken when its longest region
    // satisfies rows_min <= n < rows_cap, so a plan whose regions are almost always short (dual mode) can run them with
    // small tiles (more CTAs per SM) and leave the rare long ones to a second launch with full tiles.
    // one_set = 0: the profiles of all core sets of the plan stay in shared memory (one or two sets: explicit kits, dual).
    // one_set = 1 (many kits, `-k auto`): only the set of the tile at hand is resident -- after the kit vote a chunk's
    // tiles nearly always share one set, so it is loaded once per CTA; a tile that mixes sets runs one pass per set.
    extern __shared__ __align__(1024) uint8_t smem_bc[];
    uint8_t *smem = smem_bc;
    uint32_t *s_prof = (uint32_t *)smem;                             // [pair][code][kProfRowBytes], 1 KB per pair
    const int tile_bytes = rows_cap * kRowTile * 4;
    uint8_t *s_rows = smem + smem_profile_bytes;                     // row tiles: [2][tiles_per_iter][rows_cap][32] row-info words
    BarcodeTileSlot *s_slot = (BarcodeTileSlot *)(s_rows + 2 * tiles_per_iter * tile_bytes);
    BarcodeStage *s_stage = (BarcodeStage *)(s_slot + 2);
    const uint32_t prof_addr = (uint32_t)__cvta_generic_to_shared(s_prof);
    const uint32_t rows_addr = (uint32_t)__cvta_generic_to_shared(s_rows);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (!one_set)
        for (int i = threadIdx.x; i < f.profile_bytes / 4; i += blockDim.x) s_prof[i] = f.profile[i];
    int resident = -1;                         // one_set: byte offset of the core set whose profile is in shared memory

    const long long n_tasks = dual ? 2 * n_windows : n_windows;
    const int g = f.gap;
    // Two-launch plans (pass 0 = short regions, pass 1 = long ones): k_task_order put the short tasks in the first
    // bucket_counts[0] slots and the long ones in the last bucket_counts[1], so each launch only walks its own tiles.
    long long tile_begin = 0, n_tiles = (n_tasks + kRowTile - 1) / kRowTile;
    if (pass == 0) n_tiles = min(n_tiles, ((long long)bucket_counts[0] + kRowTile - 1) / kRowTile);
    if (pass == 1) tile_begin = (n_tasks - (long long)bucket_counts[1]) / kRowTile;
    // A CTA iteration takes a group of tiles_per_iter consecutive tiles (1 or 2): with two, a 12-pair set gives the 8 warps
    // 24 items (three full rounds instead of 8 + 4) and the tile switch is paid half as often.
    const int tpi = tiles_per_iter;
    const long long group_begin = tile_begin / tpi, n_groups = (n_tiles + tpi - 1) / tpi;
    if (group_begin + blockIdx.x >= n_groups) return;          // nothing for this CTA (before it loads any profile)

    if (threadIdx.x == 0) {
        mbar_init((uint32_t)__cvta_generic_to_shared(&s_slot[0].bar), 1);
        mbar_init((uint32_t)__cvta_generic_to_shared(&s_slot[1].bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // warp 0: start the asynchronous copy of a group's taskmeta into the staging area (no registers held meanwhile)
    auto stage_meta = [&](long long group) {
#pragma unroll
        for (int j = 0; j < kMaxTilesPerIter; ++j) {
            const long long tile = group * tpi + j;
            const long long slot = tile * kRowTile + lane;
            if (j < tpi && group < n_groups && tile < n_tiles && slot < n_tasks)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;"
                             ::"r"((uint32_t)__cvta_generic_to_shared(&s_stage->meta[j][lane])), "l"(taskmeta + slot) : "memory");
            else
                s_stage->meta[j][lane] = make_int4(0, -1, 0, 0);
        }
    };
    // warp 0: find the next group of this CTA at or after the staged candidate in which this launch takes a tile, start the
    // bulk copies into buffer b, publish the metadata, and stage the candidate after it
    auto fetch = [&](int b) {
        long long cand = s_stage->cand;
        for (;;) {
            if (cand >= n_groups) {
                if (lane == 0) s_slot[b].group = -1;
                break;
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            int4 meta_c[kMaxTilesPerIter];
            int nm[kMaxTilesPerIter];
            bool any = false;
#pragma unroll
            for (int j = 0; j < kMaxTilesPerIter; ++j) {
                meta_c[j] = s_stage->meta[j][lane];
                const int n_c = meta_c[j].y < 0 ? 0 : meta_c[j].x;
                const int nmax_c = __reduce_max_sync(0xffffffffu, n_c);
                const bool take = j < tpi && nmax_c >= rows_min && nmax_c < rows_cap;
                nm[j] = take ? nmax_c : -1;
                any = any || take;
            }
            if (any) {
#pragma unroll
                for (int j = 0; j < kMaxTilesPerIter; ++j)
                    if (j < tpi) s_slot[b].meta[j][lane] = meta_c[j];
                __syncwarp();
                if (lane == 0) {
                    s_slot[b].group = cand;
                    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_slot[b].bar);
                    uint32_t bytes = 0;
#pragma unroll
                    for (int j = 0; j < kMaxTilesPerIter; ++j) {
                        s_slot[b].nmax[j] = nm[j];
                        if (nm[j] >= 0) bytes += (uint32_t)(nm[j] + 1) * kRowTile * 4;
                    }
                    mbar_expect_tx(bar, bytes);
#pragma unroll
                    for (int j = 0; j < kMaxTilesPerIter; ++j)
                        if (nm[j] >= 0)
                            bulk_load(rows_addr + (uint32_t)((b * tpi + j) * tile_bytes),
                                      rowinfo + (cand * tpi + j) * (long long)(kRows * kRowTile), (uint32_t)(nm[j] + 1) * kRowTile * 4, bar);
                }
                cand += gridDim.x;
                stage_meta(cand);
                break;
            }
            cand += gridDim.x;                                       // nothing for this launch in it: look at the next one
            stage_meta(cand);
        }
        __syncwarp();
        if (lane == 0) s_stage->cand = cand;
        __syncwarp();
    };

    if (warp == 0) {
        if (lane == 0) s_stage->cand = group_begin + blockIdx.x;
        __syncwarp();
        stage_meta(group_begin + blockIdx.x);
        fetch(0);
    }
    __syncthreads();

    const int n_warps = (int)(blockDim.x >> 5);
    for (unsigned it = 0;; ++it) {
        const int b = (int)(it & 1);
        if (s_slot[b].group < 0) break;
        // buffer b ^ 1 was last read in the previous iteration, which every warp has left: refill it
        if (warp == 0) fetch(b ^ 1);
        mbar_wait((uint32_t)__cvta_generic_to_shared(&s_slot[b].bar), (it >> 1) & 1u);
        int item_off = 0;                      // items (tile, pair) handed out so far in this iteration, modulo the warp count
        for (int j = 0; j < tpi; ++j) {
            if (s_slot[b].nmax[j] < 0) continue;   // uniform over the CTA
            const int4 meta = s_slot[b].meta[j][lane];
            const uint32_t row_addr = rows_addr + (uint32_t)((b * tpi + j) * tile_bytes);
            const uint32_t *s_row = (const uint32_t *)(s_rows + (b * tpi + j) * tile_bytes);
            const FastGroup G = f.groups[meta.y < 0 ? 0 : meta.y];
            const int n_task = meta.y < 0 ? 0 : meta.x;
            long long w; int k;
            decode_task((long long)(uint32_t)meta.w, n_windows, dual, w, k);      // the task behind this slot
            const int rup = meta.z;
            const int npairs = (G.nb + 1) >> 1;
            const int v = G.u + (kCore - G.pad);            // last core column (template coordinates)
            int32_t *dst = bc_score + w * bslots + (k ? bmax0 : 0);
            unsigned todo = __ballot_sync(0xffffffffu, n_task > 0);               // the same in every warp of the CTA
            while (todo) {
                int n = n_task;
                uint32_t set_base = (uint32_t)G.prof_off;
                if (one_set) {
                    const int set = __reduce_min_sync(0xffffffffu, (todo >> lane) & 1u ? G.prof_off : INT32_MAX);
                    const bool mine = ((todo >> lane) & 1u) && G.prof_off == set;
                    const unsigned m_mine = __ballot_sync(0xffffffffu, mine);
                    if (set != resident) {                                         // uniform over the CTA
                        const int set_pairs = __shfl_sync(0xffffffffu, npairs, __ffs(m_mine) - 1);
                        __syncthreads();                                           // every warp is done with the old set
                        const uint32_t *src = f.profile + set / 4;
                        for (int i = threadIdx.x; i < set_pairs * (kProfPairBytes / 4); i += blockDim.x) s_prof[i] = src[i];
                        resident = set;
                        __syncthreads();
                    }
                    if (!mine) n = 0;
                    todo &= ~m_mine;
                    set_base = 0;
                } else {
                    todo = 0;
                }
                const int nmax = __reduce_max_sync(0xffffffffu, n);
                const int npairs_max = __reduce_max_sync(0xffffffffu, n > 0 ? npairs : 0);
                int first = warp - item_off;                                      // this warp's first pair of the tile
                if (first < 0) first += n_warps;
                item_off = (item_off + npairs_max) % n_warps;
                for (int pr = first; pr < npairs_max; pr += n_warps) {
                    const int pcl = min(pr, npairs - 1);
                    const uint32_t block = prof_addr + set_base + (uint32_t)pcl * kProfPairBytes;
                    uint32_t Wc[kCore];
    #pragma unroll
                    for (int c = 0; c < kCore; ++c) Wc[c] = dup16((uint32_t)((G.u + max(0, c + 1 - G.pad)) * g));
                    const uint32_t info0 = s_row[lane];
                    uint32_t Fprev = dup16((info0 >> kRowFShift) & kRowFMask);
                    uint32_t acc = dup16((uint32_t)(v * g)) + dup16(info0 >> kRowGShift);  // join term of row 0
                    int i = 1;
                    while (i <= nmax) {
                        // rows up to the next row at which some lane's region ends run without per-lane branching; a lane's
                        // scores are taken at its own last row, so whatever it computes afterwards is never used
                        const int ev = __reduce_min_sync(0xffffffffu, n >= i ? n : INT32_MAX);
                        uint32_t rp = row_addr + (uint32_t)(i * kRowTile + lane) * 4u;
                        const uint32_t rp_end = row_addr + (uint32_t)((ev + 1) * kRowTile + lane) * 4u;
                        i = ev + 1;
    #pragma unroll 1
                        do {
                            const uint32_t info = lds32(rp);
                            rp += kRowTile * 4;
                            const uint32_t prow = block | (info & kRowCodeMask);
                            const uint32_t Fi = dup16((info >> kRowFShift) & kRowFMask);
                            const uint32_t Gi = dup16(info >> kRowGShift);
                            uint32_t e[kCore];
    #pragma unroll
                            for (int c = 0; c < kCore; c += 4) {
                                const uint4 q = lds128(prow + c * 4);
                                e[c] = q.x; e[c + 1] = q.y; e[c + 2] = q.z; e[c + 3] = q.w;
                            }
                            uint32_t left = Fi;
                            uint32_t t = e[0] + Fprev;
    #pragma unroll
                            for (int c = 0; c < kCore; ++c) {
                                const uint32_t tn = c + 1 < kCore ? e[c + 1] + Wc[c] : 0u;   // next column's diagonal term first
                                left = __vimax3_u16x2(t, Wc[c], left);
                                Wc[c] = left;
                                t = tn;
                            }
                            Fprev = Fi;
                            acc = __viaddmax_u16x2(left, Gi, acc);
                        } while (rp != rp_end);
                        if (n == ev && pr < npairs) store_pair_scores(Wc, acc, G, n, g, rup, pr, dst);
                    }
                }
            }
        }
        __syncthreads();                       // every warp is done with buffer b; warp 0's slot b ^ 1 is published
    }
}
