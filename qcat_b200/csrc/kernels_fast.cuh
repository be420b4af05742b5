// Packed (u16x2) linear-gap DP kernels for sm_100a -- the fast path of the barcode and adapter stages.
//
// Formulation (exact in integers; see DESIGN.md "Kernels"):
//   * linear gaps (open == extend == g): H[i][j] = max(H[i-1][j-1] + M, H[i-1][j] - g, H[i][j-1] - g).
//     With W[i][j] = H[i][j] + (i + j) g this becomes W = max3(Wdiag + s', Wup, Wleft), s' = M + 2g >= 0,
//     W >= 0 and small (< 2^15): two DP problems share one 32-bit lane as u16x2 and one cell costs one
//     integer add plus one VIMNMX3.U16x2.
//   * barcode stage: all templates of a group are up_context + barcode + down_context.  The columns of
//     the shared prefix are computed once per window (column F), the shared suffix is folded into one
//     backward column G (best continuation from the end of the barcode to any free end), so each barcode
//     only costs its own <= 24 core columns plus one add-max per row for the join.
//   * mapping: one lane = one window x two barcodes (the two u16 halves); a warp = 32 windows x one
//     barcode pair, so the pair's substitution profile is read from shared memory with multicast loads
//     (one 32-bit word per cell pair, rows selected by the lane's own base code).
#pragma once

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "plan.h"

namespace qcb {

constexpr int kCore = 24;                 // core (barcode-specific) columns held in registers
constexpr int kProfWords = 28;            // words per (pair, code) profile row: 24 used + 4 pad (bank spread)
constexpr int kRows = kFastMaxStride + 1; // DP rows 0..160
constexpr int kTile = 32;                 // windows per tile (one per lane)
constexpr int kBarcodeWarps = 8;
constexpr int kMaxCtx = 16;               // longest shared prefix / suffix handled by the prologue
constexpr int kMaxFastGroups = 64;

struct FastGroup {          // one template group (layout, set k)
    int32_t ok;             // group usable by the packed kernel
    int32_t u, d;           // shared prefix / suffix length (columns)
    int32_t pad;            // dead columns in front of the core (kCore - core length)
    int32_t tlen;           // template length (all templates of the group)
    int32_t nb;             // barcodes in the group
    int32_t prof_off;       // byte offset of the group's core-set profile inside the profile image
    int32_t up_off, down_off;  // offsets of the context codes in ctx_codes
};

struct FastDev {
    const uint32_t *profile;     // [set][pair][code][kProfWords]
    int32_t profile_bytes;
    const FastGroup *groups;     // [n_groups]
    const uint8_t *ctx_codes;    // barcode-matrix codes of the shared prefix / suffix columns
    const int32_t *sprime;       // [bmat_size * bmat_size] M + 2g
    int32_t n_codes;             // bmat_size
    int32_t gap;                 // g
    int32_t n_groups;
};

struct FastPlan {
    bool adapter_ok = false;
    bool barcode_ok = false;
    std::string error;
    FastDev dev{};
    void *slab = nullptr;
    size_t slab_bytes = 0;
    int sm_count = 0;
    size_t barcode_smem = 0;
    size_t workspace_bytes() const { return slab_bytes; }
};

// ---------------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t dup16(uint32_t v) { return v * 0x00010001u; }

// Last column of the semi-global DP (zero borders, H space, linear gap g) of `n` region rows against `len`
// context columns, for every row: out[pos(i)] = H[i][len] + (i + len) g   (W space, >= 0).
// reverse = false: rows i = 1..n read region code i, columns read ctx[0..len), pos(i) = i            (column F)
// reverse = true : rows i' = 1..n read region code n - i' + 1, columns read ctx[len-1..0], the left border of
//                  row n is -g instead of 0 and pos(i') = n - i'                                      (column G)
// Returns max_{1<=j<=len} H[n][j] of the forward problem (INT_MIN/2 when len == 0).
__device__ __forceinline__ int context_column(const uint8_t *s_code, int lane, int n, const uint8_t *ctx, int len,
                                              const int32_t *s_sp, int n_codes, int g, bool reverse, uint16_t *s_out)
{
    int H[kMaxCtx + 1];
#pragma unroll
    for (int j = 0; j <= kMaxCtx; ++j) H[j] = 0;
    int cc[kMaxCtx];
#pragma unroll
    for (int j = 0; j < kMaxCtx; ++j) cc[j] = j < len ? ctx[reverse ? len - 1 - j : j] : 0;
    s_out[(reverse ? n : 0) * kTile + lane] = (uint16_t)(len * g);          // row 0: H = 0
    for (int i = 1; i <= n; ++i) {
        int code = s_code[(reverse ? n - i + 1 : i) * kTile + lane];
        const int32_t *sp = s_sp + code * n_codes;
        int diag = 0;
        int left = (reverse && i == n) ? -g : 0;
        H[0] = left;
#pragma unroll
        for (int j = 1; j <= kMaxCtx; ++j) {
            if (j <= len) {
                int up = H[j];
                int h = max(max(diag + sp[cc[j - 1]] - 2 * g, up - g), left - g);
                diag = up; H[j] = h; left = h;
            }
        }
        int last = len > 0 ? left : H[0];
        s_out[(reverse ? n - i : i) * kTile + lane] = (uint16_t)(last + (i + len) * g);
    }
    int rmax = INT32_MIN / 2;
#pragma unroll
    for (int j = 1; j <= kMaxCtx; ++j)
        if (j <= len) rmax = max(rmax, H[j]);
    return rmax;
}

// Barcode stage.  Persistent CTAs; each iteration takes one tile of 32 window tasks.  Task t = window (epi2me) or
// (window, set) (dual: t = 2 w + k).  Output: bc_score[w * bslots + (k ? bmax0 : 0) + b] for every barcode b.
__global__ void __launch_bounds__(kBarcodeWarps * 32, 3)
k_barcode_fast(FastDev f, DevTables t, const uint8_t *__restrict__ wins, int stride, long long n_windows,
               const WindowSel *__restrict__ sel, int dual, int bmax0, int bslots, int32_t *__restrict__ bc_score)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *s_prof = (uint32_t *)smem;
    uint8_t *p = smem + f.profile_bytes;
    uint16_t *s_off = (uint16_t *)p;  p += kRows * kTile * 2;      // profile row byte offset of the row's base code
    uint16_t *s_F = (uint16_t *)p;    p += kRows * kTile * 2;      // F column (W space)
    uint16_t *s_G = (uint16_t *)p;    p += kRows * kTile * 2;      // G column (W space)
    uint8_t *s_code = p;              p += kRows * kTile;          // barcode-matrix code of region row i
    int32_t *s_sp = (int32_t *)p;     p += kMaxMatrix * kMaxMatrix * 4;
    int32_t *s_rup = (int32_t *)p;    p += kTile * 4;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < f.profile_bytes / 4; i += blockDim.x) s_prof[i] = f.profile[i];
    for (int i = threadIdx.x; i < f.n_codes * f.n_codes; i += blockDim.x) s_sp[i] = f.sprime[i];

    const long long n_tasks = dual ? 2 * n_windows : n_windows;
    const long long n_tiles = (n_tasks + kTile - 1) / kTile;
    const int g = f.gap;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();                       // previous tile fully consumed (also covers the table loads above)
        // ---- per-lane task description (every warp computes the same values for its lane) ----
        long long task = tile * kTile + lane;
        bool valid = task < n_tasks;
        long long w = valid ? (dual ? task >> 1 : task) : 0;
        int k = (valid && dual) ? (int)(task & 1) : 0;
        WindowSel s = sel[w];
        int lo = k ? s.lo1 : s.lo0, hi = k ? s.hi1 : s.hi0;
        int n = valid ? hi - lo : 0;
        int grp = t.group[s.layout * 2 + k];
        FastGroup G = f.groups[grp < 0 ? 0 : grp];
        if (grp < 0) n = 0;

        // ---- phase 0: region codes -> shared memory (thread x handles rows x/32, x/32 + 8, ... of its lane) ----
        {
            const uint8_t *src = wins + w * stride + lo;
            for (int i = warp; i < kRows; i += kBarcodeWarps) {
                int code = 0;
                if (i >= 1 && i <= n) code = t.bmap[src[i - 1]];
                s_code[i * kTile + lane] = (uint8_t)code;
                s_off[i * kTile + lane] = (uint16_t)(code * (kProfWords * 4));
            }
        }
        __syncthreads();
        // ---- phase 1: shared-prefix column F (warp 0) and shared-suffix column G (warp 1) ----
        if (warp == 0) {
            int rup = context_column(s_code, lane, n, f.ctx_codes + G.up_off, G.u, s_sp, f.n_codes, g, false, s_F);
            s_rup[lane] = rup;
        } else if (warp == 1) {
            context_column(s_code, lane, n, f.ctx_codes + G.down_off, G.d, s_sp, f.n_codes, g, true, s_G);
        }
        __syncthreads();

        // ---- phase 2: core columns, one barcode pair per warp iteration ----
        const int nmax = __reduce_max_sync(0xffffffffu, n);
        const int npairs = (G.nb + 1) >> 1;
        const int npairs_max = __reduce_max_sync(0xffffffffu, n > 0 ? npairs : 0);
        const int m = G.tlen;
        const int v = G.u + (kCore - G.pad);            // last core column (template coordinates)
        for (int pr = warp; pr < npairs_max; pr += kBarcodeWarps) {
            const int pcl = min(pr, npairs - 1);
            const uint8_t *prow = (const uint8_t *)s_prof + G.prof_off + pcl * (f.n_codes * kProfWords * 4);
            uint32_t Wc[kCore];
#pragma unroll
            for (int c = 0; c < kCore; ++c) Wc[c] = dup16((uint32_t)((G.u + max(0, c + 1 - G.pad)) * g));
            uint32_t Fprev = dup16(s_F[lane]);
            uint32_t acc = dup16((uint32_t)(v * g)) + dup16(s_G[lane]);          // join term of row 0
            for (int i = 1; i <= nmax; ++i) {
                const uint4 *prow_i = (const uint4 *)(prow + s_off[i * kTile + lane]);
                uint32_t e[kCore];
#pragma unroll
                for (int c = 0; c < kCore; c += 4) {
                    uint4 q = prow_i[c >> 2];
                    e[c] = q.x; e[c + 1] = q.y; e[c + 2] = q.z; e[c + 3] = q.w;
                }
                const uint32_t Fi = dup16(s_F[i * kTile + lane]);
                const uint32_t Gi = dup16(s_G[i * kTile + lane]);
                uint32_t diag = Fprev, left = Fi;
#pragma unroll
                for (int c = 0; c < kCore; ++c) {
                    uint32_t tt = diag + e[c];
                    diag = Wc[c];
                    left = __vimax3_u16x2(tt, diag, left);
                    Wc[c] = left;
                }
                Fprev = Fi;
                if (i <= n) acc = __viaddmax_u16x2(left, Gi, acc);
                if (__any_sync(0xffffffffu, i == n)) {
                    if (i == n && pr < npairs) {
                        // last row over the core columns: max_c (W[n][col_c] - (n + col_c) g), col_c = u + max(0, c+1-pad)
                        const int CB = (G.u + kCore) * g;
                        uint32_t rm = 0;
#pragma unroll
                        for (int c = 0; c < kCore - 1; ++c)
                            rm = __viaddmax_u16x2(Wc[c], dup16((uint32_t)(CB - (G.u + max(0, c + 1 - G.pad)) * g)), rm);
                        const int rup = s_rup[lane];
                        const int bias_r = n * g + CB, bias_j = (n + m) * g;
                        int s0 = max(max((int)(rm & 0xffffu) - bias_r, (int)(acc & 0xffffu) - bias_j), rup);
                        int s1 = max(max((int)(rm >> 16) - bias_r, (int)(acc >> 16) - bias_j), rup);
                        int32_t *dst = bc_score + w * bslots + (k ? bmax0 : 0) + 2 * pr;
                        dst[0] = s0;
                        if (2 * pr + 1 < G.nb) dst[1] = s1;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// host side: build the packed-kernel tables
// ---------------------------------------------------------------------------------------------------

inline void fast_plan_free(FastPlan &fp)
{
    if (fp.slab) cudaFree(fp.slab);
    fp.slab = nullptr; fp.slab_bytes = 0;
}

inline int fast_plan_build(FastPlan &fp, const qcb_tables *h, int sm_count)
{
    fp.sm_count = sm_count;
    fp.adapter_ok = false;
    fp.barcode_ok = false;
    const int g = h->barcode_open;
    const int nc = h->bmat_size;
    // --- scoring preconditions: linear gap, non-negative shifted scores, values fit 15 bits ---
    if (h->barcode_open != h->barcode_extend || g <= 0) return 0;
    int smax = 0;
    for (int i = 0; i < nc * nc; ++i) {
        if (h->bmat[i] + 2 * g < 0) return 0;
        smax = std::max(smax, h->bmat[i]);
    }
    const int ng = h->n_groups;
    if (ng <= 0 || ng > kMaxFastGroups) return 0;
    if (h->max_align_length > kFastMaxStride) return 0;

    std::vector<FastGroup> groups(ng);
    std::vector<uint8_t> ctx;
    std::vector<uint32_t> profile;
    std::vector<std::vector<uint8_t>> set_keys;      // de-duplicated core sets
    std::vector<int> set_off;
    for (int gi = 0; gi < ng; ++gi) {
        FastGroup &G = groups[gi];
        memset(&G, 0, sizeof G);
        const int b0 = h->group_off[gi], b1 = h->group_off[gi + 1];
        const int nb = b1 - b0;
        if (nb <= 0) return 0;
        const int tlen = h->tmpl_off[b0 + 1] - h->tmpl_off[b0];
        for (int b = b0; b < b1; ++b)
            if (h->tmpl_off[b + 1] - h->tmpl_off[b] != tlen) return 0;       // ragged templates -> generic kernels
        const uint8_t *first = h->tmpl_seq + h->tmpl_off[b0];
        int u = tlen, d = tlen;
        for (int b = b0 + 1; b < b1; ++b) {
            const uint8_t *sq = h->tmpl_seq + h->tmpl_off[b];
            int a = 0; while (a < u && sq[a] == first[a]) ++a; u = a;
            int z = 0; while (z < d && sq[tlen - 1 - z] == first[tlen - 1 - z]) ++z; d = z;
        }
        u = std::min(u, kMaxCtx); d = std::min(d, kMaxCtx);
        if (u + d > tlen - 1) {                      // single-template groups: keep at least one core column
            d = std::min(d, std::max(0, tlen - 1 - u));
            if (u + d > tlen - 1) u = tlen - 1 - d;
        }
        int core = tlen - u - d;
        while (core > kCore && u < kMaxCtx && nb == 1) { ++u; --core; }
        if (core < 1 || core > kCore) return 0;
        if ((smax * tlen + (kFastMaxStride + tlen) * g) * 2 + 64 >= 65536) return 0;
        G.ok = 1; G.u = u; G.d = d; G.pad = kCore - core; G.tlen = tlen; G.nb = nb;
        G.up_off = (int32_t)ctx.size();
        for (int j = 0; j < u; ++j) ctx.push_back(h->bmap[first[j]]);
        G.down_off = (int32_t)ctx.size();
        for (int j = 0; j < d; ++j) ctx.push_back(h->bmap[first[tlen - d + j]]);
        // core set key: pad + the core columns of every barcode
        std::vector<uint8_t> key;
        key.push_back((uint8_t)G.pad);
        for (int b = b0; b < b1; ++b) {
            const uint8_t *sq = h->tmpl_seq + h->tmpl_off[b];
            for (int c = 0; c < core; ++c) key.push_back(h->bmap[sq[u + c]]);
        }
        int found = -1;
        for (size_t si = 0; si < set_keys.size(); ++si)
            if (set_keys[si] == key) { found = (int)si; break; }
        if (found < 0) {
            found = (int)set_keys.size();
            set_keys.push_back(key);
            set_off.push_back((int)(profile.size() * 4));
            const int npairs = (nb + 1) / 2;
            for (int pr = 0; pr < npairs; ++pr)
                for (int code = 0; code < nc; ++code)
                    for (int c = 0; c < kProfWords; ++c) {
                        uint32_t word = 0;
                        if (c >= G.pad && c < kCore) {
                            int ba = 2 * pr, bb = std::min(2 * pr + 1, nb - 1);
                            int ca = key[1 + ba * core + (c - G.pad)], cb = key[1 + bb * core + (c - G.pad)];
                            uint32_t lo = (uint32_t)(h->bmat[code * nc + ca] + 2 * g);
                            uint32_t hi = (uint32_t)(h->bmat[code * nc + cb] + 2 * g);
                            word = lo | (hi << 16);
                        }
                        profile.push_back(word);
                    }
        }
        G.prof_off = set_off[found];
    }
    const size_t profile_bytes = profile.size() * 4;
    if (profile_bytes + 4 * kRows * kTile * 2 > 60 * 1024) {
        // keep three CTAs per SM: profile + per-tile columns must stay under ~72 KB
        if (profile_bytes > 150 * 1024) return 0;
    }
    std::vector<int32_t> sprime(nc * nc);
    for (int i = 0; i < nc * nc; ++i) sprime[i] = h->bmat[i] + 2 * g;

    // one device slab: profile | groups | ctx | sprime
    size_t o_prof = 0;
    size_t o_grp = (o_prof + profile_bytes + 255) / 256 * 256;
    size_t o_ctx = (o_grp + groups.size() * sizeof(FastGroup) + 255) / 256 * 256;
    size_t o_sp = (o_ctx + ctx.size() + 16 + 255) / 256 * 256;
    size_t total = o_sp + sprime.size() * 4;
    std::vector<uint8_t> slab(total, 0);
    memcpy(slab.data() + o_prof, profile.data(), profile_bytes);
    memcpy(slab.data() + o_grp, groups.data(), groups.size() * sizeof(FastGroup));
    if (!ctx.empty()) memcpy(slab.data() + o_ctx, ctx.data(), ctx.size());
    memcpy(slab.data() + o_sp, sprime.data(), sprime.size() * 4);
    if (cudaMalloc(&fp.slab, total) != cudaSuccess) { fp.error = "cudaMalloc failed"; return 1; }
    fp.slab_bytes = total;
    if (cudaMemcpy(fp.slab, slab.data(), total, cudaMemcpyHostToDevice) != cudaSuccess) { fp.error = "cudaMemcpy failed"; return 1; }
    const uint8_t *dptr = (const uint8_t *)fp.slab;
    fp.dev.profile = (const uint32_t *)(dptr + o_prof);
    fp.dev.profile_bytes = (int32_t)profile_bytes;
    fp.dev.groups = (const FastGroup *)(dptr + o_grp);
    fp.dev.ctx_codes = dptr + o_ctx;
    fp.dev.sprime = (const int32_t *)(dptr + o_sp);
    fp.dev.n_codes = nc;
    fp.dev.gap = g;
    fp.dev.n_groups = ng;
    fp.barcode_smem = profile_bytes + (size_t)kRows * kTile * (2 + 2 + 2 + 1) + kMaxMatrix * kMaxMatrix * 4 + kTile * 4;
    if (fp.barcode_smem > 220 * 1024) return 0;
    if (cudaFuncSetAttribute(k_barcode_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fp.barcode_smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    fp.barcode_ok = true;
    return 0;
}

inline int fast_adapter_stage(FastPlan &, const DevTables &, const uint8_t *, int, const int32_t *, long long,
                              const int32_t *, const int32_t *, int, int32_t *, int32_t *, cudaStream_t, long long *)
{
    return 1;
}

inline int fast_barcode_stage(FastPlan &fp, const DevTables &t, const uint8_t *wins, int stride, long long n_windows,
                              const WindowSel *sel, int bmax0, int bslots, int32_t *bc_score, cudaStream_t st,
                              long long *launches)
{
    const int dual = t.mode == QCB_MODE_DUAL ? 1 : 0;
    long long n_tasks = dual ? 2 * n_windows : n_windows;
    long long n_tiles = (n_tasks + kTile - 1) / kTile;
    int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(3, (220 * 1024) / std::max<size_t>(1, fp.barcode_smem)));
    int grid = (int)std::min<long long>(n_tiles, (long long)fp.sm_count * ctas_per_sm);
    if (grid <= 0) return 0;
    k_barcode_fast<<<grid, kBarcodeWarps * 32, fp.barcode_smem, st>>>(fp.dev, t, wins, stride, n_windows, sel, dual, bmax0, bslots,
                                                                  bc_score);
    ++*launches;
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace qcb
