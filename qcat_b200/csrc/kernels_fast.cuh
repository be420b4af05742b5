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
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "plan.h"

namespace qcb {

constexpr int kCore = 24;                 // core (barcode-specific) columns held in registers
constexpr int kProfWords = 36;            // words per (pair, code) profile row: 24 used + 12 pad; 144 B = (1 << 4) | (1 << 7)
constexpr int kProfPairBytes = 1024;      // one pair's profile block (<= 7 codes x 144 B), 1 KB aligned: see kRowCodeMask
constexpr int kRows = kFastMaxStride + 1; // DP rows 0..160
constexpr int kTile = 32;                 // windows per tile (one per lane)
constexpr int kMaxCtx = 16;               // longest shared prefix / suffix handled by the prologue
constexpr int kMaxFastGroups = 64;

struct FastGroup {          // one template group (layout, set k)
    int32_t u, d;           // shared prefix / suffix length (columns)
    int32_t pad;            // dead columns in front of the core (kCore - core length)
    int32_t tlen;           // template length (all templates of the group)
    int32_t nb;             // barcodes in the group
    int32_t prof_off;       // byte offset of the group's core-set profile inside the profile image
    int32_t up_off, down_off;  // offsets of the context codes in the host-side ctx list (table construction only)
};

struct FastDev {
    const uint32_t *profile;     // [set][pair] blocks of kProfPairBytes: [code][kProfWords]
    int32_t profile_bytes;
    const FastGroup *groups;     // [n_groups]
    int32_t n_codes;             // bmat_size
    int32_t gap;                 // g
    int32_t n_groups;
};

struct AdapterClass {        // the templates of one length class (NC register columns), two per lane
    int nc_cols = 0, npairs = 0, row_words = 0;
    size_t profile_bytes = 0;    // profile words; followed by npairs int4 {len lo, len hi, has hi, output slots lo | hi << 16}
    size_t offset = 0;           // byte offset inside AdapterSubset::dev
};

struct AdapterSubset {       // packed adapter profiles for one layout subset, grouped by template length
    std::vector<int32_t> key;
    void *dev = nullptr;
    size_t bytes = 0;
    std::vector<AdapterClass> classes;
    std::vector<uint8_t> host;   // staging copy (kept alive for the async upload)
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
    int max_pairs = 1;               // barcode pairs of the largest template group
    std::vector<int> group_pairs;    // barcode pairs of every template group
    bool one_set = false;            // k_barcode_fast keeps one core set's profile in shared memory at a time (many kits)
    size_t profile_smem = 0;         // bytes of the profile region of k_barcode_fast's shared memory
    int short_rows = 0;              // > 0: row-tile size of the first of two k_barcode_fast launches (dual mode)
    int bucket_rows = kRows;         // regions shorter than this go to the front of the task order (k_task_order)
    const uint32_t *ctx_tab = nullptr;   // device: k_context score tables
    int ctx_ncol = 12;
    bool ctx_pair = true;                // ctx_tab holds one word per (F code, G code, column) instead of two tables
    size_t context_smem = 0;
    void *rowinfo = nullptr;         // [tiles][kRows][32] u32: code row offset | F << 10 | G << 21 (kernels_barcode_fast.cuh)
    void *taskmeta = nullptr;        // [slots] int4 {region length, group, R over prefix columns, task}
    void *perm = nullptr;            // [slots] u32 task of every slot (k_task_order), followed by the two bucket counters
    size_t perm_bytes = 0;
    size_t rowinfo_bytes = 0, taskmeta_bytes = 0;
    // adapter stage: host copies used to build per-subset profiles
    int a_gap = 0, a_codes = 0;
    std::vector<int32_t> a_mat;
    std::vector<uint8_t> a_map;
    std::vector<std::vector<uint8_t>> a_seq;       // per layout adapter sequence (ASCII)
    std::vector<AdapterSubset *> subsets;
    bool adapter_smem_configured[4] = {false, false, false, false};   // k_adapter_fast<NC> opted into full dynamic shared memory
    int smem_optin = 0;              // cudaDevAttrMaxSharedMemoryPerBlockOptin of the plan's device
    size_t workspace_bytes() const
    {
        size_t b = slab_bytes + rowinfo_bytes + taskmeta_bytes + perm_bytes;
        for (auto *sub : subsets) b += sub->bytes;
        return b;
    }
};

// ---------------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t dup16(uint32_t v) { return v * 0x00010001u; }

#include "kernels_barcode_fast.cuh"

// ---------------------------------------------------------------------------------------------------
// Adapter stage.  One lane = one window x two adapter templates (u16 halves); a warp = 32 windows x one
// template pair; the whole template lives in NC registers, right-aligned (dead columns in front keep the
// left border value i*g, so the last register is always the last template column).  Per row the last
// column feeds a (value, first row) key; at row n the last row is scanned for (value, first column) and the
// parasail end-cell rule is applied (see sg_affine in kernels_generic.cuh / qo_sg in the oracle).
// ---------------------------------------------------------------------------------------------------
constexpr int kAdapterWarps = 4;
constexpr int kAdapterPhase = 24;       // columns whose profile words are in registers at a time

// CTAs per SM the register allocation of each column variant is sized for
#ifndef QCB_AD48_BLOCKS
#define QCB_AD48_BLOCKS 5
#endif
#ifndef QCB_AD56_BLOCKS
#define QCB_AD56_BLOCKS 4
#endif
#ifndef QCB_AD64_BLOCKS
#define QCB_AD64_BLOCKS 4
#endif
#ifndef QCB_AD104_BLOCKS
#define QCB_AD104_BLOCKS 3
#endif
constexpr int adapter_blocks(int nc) { return nc <= 48 ? QCB_AD48_BLOCKS : (nc <= 56 ? QCB_AD56_BLOCKS : (nc <= 64 ? QCB_AD64_BLOCKS : QCB_AD104_BLOCKS)); }
template <int NC>
__global__ void __launch_bounds__(kAdapterWarps * 32, adapter_blocks(NC))
k_adapter_fast(const uint32_t *__restrict__ profile, int profile_words, int row_words, int n_codes,
               const int4 *__restrict__ pair_meta, int npairs,
               const uint8_t *__restrict__ codes, int stride, const int32_t *__restrict__ wlen, int wshift,
               long long n_windows, int n_subset, int g, int32_t *__restrict__ ad_score, int32_t *__restrict__ ad_end)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *s_prof = (uint32_t *)smem;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t *s_code = smem + (size_t)profile_words * 4 + (size_t)warp * (kRows * kTile);
    for (int i = threadIdx.x; i < profile_words; i += blockDim.x) s_prof[i] = profile[i];
    // lanes whose window is shorter than their tile's longest keep stepping through rows they never wrote: make sure
    // whatever is there is a valid base code from the start
    for (int i = lane; i < kRows * kTile / 4; i += 32) ((uint32_t *)s_code)[i] = 0u;
    __syncthreads();
    const uint32_t prof_addr = (uint32_t)__cvta_generic_to_shared(s_prof);
    const uint32_t code_addr = (uint32_t)__cvta_generic_to_shared(s_code) + lane;
    const uint32_t row_bytes = (uint32_t)row_words * 4u;

    const long long n_tiles = (n_windows + kTile - 1) / kTile;
    const long long n_tasks = n_tiles * npairs;
    const long long warp_stride = (long long)gridDim.x * kAdapterWarps;
    for (long long task = (long long)blockIdx.x * kAdapterWarps + warp; task < n_tasks; task += warp_stride) {
        const long long tile = task / npairs;
        const int q = (int)(task % npairs);
        const int4 meta = pair_meta[q];                 // x = len lo, y = len hi, z = has hi, w = output slots (lo | hi << 16)
        const int m_lo = meta.x, m_hi = meta.y;
        const long long w = tile * kTile + lane;
        const bool valid = w < n_windows;
        const int n = valid ? wlen[w >> wshift] : 0;
        // ---- stage the lane's window codes (low nibble = adapter-matrix code) into shared memory; the 3' window of a
        // read (odd w; wshift = 1) is stored unreversed by k_map_codes: its base p is row n - p of the DP ----
        __syncwarp();
        {
            const uint4 *src = (const uint4 *)(codes + (valid ? w : 0) * stride);
            const bool rev = wshift && (w & 1);
            for (int ch = 0; ch * 16 < n; ++ch) {
                uint4 v = src[ch];
                uint32_t words[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int b = 0; b < 16; ++b) {
                    const int p = ch * 16 + b;
                    const int row = rev ? n - p : p + 1;                 // rows beyond the window (p >= n) are never used
                    if (row >= 1) s_code[row * kTile + lane] = (uint8_t)((words[b >> 2] >> ((b & 3) * 8)) & 15u);
                }
            }
        }
        __syncwarp();
        const int nmax = __reduce_max_sync(0xffffffffu, n);
        const uint32_t pbase = prof_addr + (uint32_t)(q * n_codes) * row_bytes;
        uint32_t Wc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            int jl = c - (NC - m_lo) + 1, jh = c - (NC - m_hi) + 1;
            Wc[c] = (uint32_t)(max(jl, 0) * g) | ((uint32_t)(max(jh, 0) * g) << 16);
        }
        const uint32_t gdup = dup16((uint32_t)g);
        uint32_t border = 0;                            // (i - 1) * g in both halves
        int best_lo = 0, best_hi = 0;                   // ((W - i g) << 8) | (255 - i), max over rows
        int rowc = 255;                                 // (255 - i) - ((i g) << 8) of the row at hand
        const int rowc_step = 1 + (g << 8);
        uint32_t cp = code_addr + kTile;                // row 1
        int i = 1;
        while (i <= nmax) {
            // rows up to the next row at which some lane's window ends run without per-lane branches; a lane's result is
            // taken at its own last row, so whatever it computes afterwards is never used
            const int ev = __reduce_min_sync(0xffffffffu, n >= i ? n : INT32_MAX);
            const uint32_t cp_end = code_addr + (uint32_t)(ev + 1) * kTile;
            i = ev + 1;
#pragma unroll 1
            do {
                const uint32_t prow = pbase + lds_u8(cp) * row_bytes;
                cp += kTile;
                // Columns go in phases of kAdapterPhase: a phase's profile words are loaded first (volatile loads: ptxas
                // keeps them ahead of the max chain), every diagonal term is formed one column ahead of the in-place
                // update, and only the last column of a phase is copied (its old value starts the next phase).
                const uint32_t border_prev = border;
                border += gdup;
                uint32_t left = border, t = 0, carry = border_prev;
#pragma unroll
                for (int p0 = 0; p0 < NC; p0 += kAdapterPhase) {
                    constexpr int kMaxCols = kAdapterPhase;
                    const int cols = NC - p0 < kMaxCols ? NC - p0 : kMaxCols;
                    uint32_t e[kMaxCols];
#pragma unroll
                    for (int c = 0; c < kMaxCols; c += 4) {
                        if (c < cols) {
                            const uint4 q = lds128(prow + (uint32_t)(p0 + c) * 4u);
                            e[c] = q.x; e[c + 1] = q.y; e[c + 2] = q.z; e[c + 3] = q.w;
                        }
                    }
                    t = e[0] + carry;
#pragma unroll
                    for (int c = 0; c < kMaxCols; ++c) {
                        if (c < cols) {
                            const uint32_t tn = c + 1 < cols ? e[c + 1] + Wc[p0 + c] : 0u;
                            if (c == cols - 1) carry = Wc[p0 + c];
                            left = __vimax3_u16x2(t, Wc[p0 + c], left);
                            Wc[p0 + c] = left;
                            t = tn;
                        }
                    }
                }
                rowc -= rowc_step;
                best_lo = max(best_lo, (int)(left & 0xffffu) * 256 + rowc);
                best_hi = max(best_hi, (int)(left >> 16) * 256 + rowc);
            } while (cp != cp_end);
            if (n == ev) {
                int rb_lo = -1, rb_hi = -1;         // ((W - j g) << 8) | (255 - j), max over real columns
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const int jl = c - (NC - m_lo) + 1, jh = c - (NC - m_hi) + 1;
                    if (jl >= 1) rb_lo = max(rb_lo, ((int)(Wc[c] & 0xffffu) - jl * g) * 256 + (255 - jl));
                    if (jh >= 1) rb_hi = max(rb_hi, ((int)(Wc[c] >> 16) - jh * g) * 256 + (255 - jh));
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h == 1 && !meta.z) break;
                    const int m = h ? m_hi : m_lo;
                    const int best = h ? best_hi : best_lo, rb = h ? rb_hi : rb_lo;
                    const int C = (best >> 8) - m * g, iC = 255 - (best & 255);
                    const int R = (rb >> 8) - n * g, jR = 255 - (rb & 255);
                    int score, end;
                    if (C > R) { score = C; end = iC - 1; }
                    else { score = R; end = n - 1; if (jR == m) end = iC - 1; }
                    const long long o = w * n_subset + ((pair_meta[q].w >> (16 * h)) & 0xffff);
                    ad_score[o] = score;
                    ad_end[o] = end;
                }
            }
        }
    }
}

// Input bytes -> packed matrix codes for the packed kernels: one byte per base, low nibble = adapter-matrix code, high
// nibble = barcode-matrix code.  codes[2r] = read[:W] mapped; codes[2r + 1] = read[-W:] mapped through the complement
// but NOT reversed -- the staging loops of k_adapter_fast / k_context write the bases of odd windows back to front, which
// costs them nothing, and leaves this pass a pure element-wise map: 16 bytes per thread, coalesced in and out.
__global__ void k_map_codes(const uint8_t *__restrict__ win5, const uint8_t *__restrict__ tail3, int stride,
                            const int32_t *__restrict__ wlen, long long n_reads, const uint8_t *__restrict__ comp,
                            const uint8_t *__restrict__ amap, const uint8_t *__restrict__ bmap, uint8_t *__restrict__ codes)
{
    __shared__ uint8_t s_fwd[256], s_rev[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        s_fwd[i] = (uint8_t)(amap[i] | (bmap[i] << 4));
        const uint8_t c = comp[i];
        s_rev[i] = (uint8_t)(amap[c] | (bmap[c] << 4));
    }
    __syncthreads();
    const int vecs = stride >> 4;                                 // 16-byte groups per window slot
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_reads * 2 * vecs) return;
    const long long w = id / vecs;
    const int v = (int)(id % vecs);
    const long long r = w >> 1;
    const int len = min(max(wlen[r], 0), stride);
    const uint8_t *tab = (w & 1) ? s_rev : s_fwd;
    const uint4 in = ((const uint4 *)((w & 1) ? tail3 : win5))[r * vecs + v];
    const uint32_t src[4] = {in.x, in.y, in.z, in.w};
    uint32_t dst[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t o = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = v * 16 + q * 4 + b;
            const uint32_t code = i < len ? tab[(src[q] >> (8 * b)) & 255u] : 0u;
            o |= code << (8 * b);
        }
        dst[q] = o;
    }
    ((uint4 *)codes)[id] = make_uint4(dst[0], dst[1], dst[2], dst[3]);
}

// The same from 4-bit windows (two base classes per byte, qcb_plan_base_classes): tabs = fwd[16] | rev[16], the packed
// code byte of every class as it stands / complemented.  8 input bytes -> 16 code bytes per thread.
__global__ void k_map_codes4(const uint8_t *__restrict__ win5p, const uint8_t *__restrict__ tail3p, int stride4,
                             const int32_t *__restrict__ wlen, long long n_reads, const uint8_t *__restrict__ tabs,
                             uint8_t *__restrict__ codes)
{
    __shared__ uint8_t s_tab[32];
    if (threadIdx.x < 32) s_tab[threadIdx.x] = tabs[threadIdx.x];
    __syncthreads();
    const int stride = 2 * stride4, vecs = stride >> 4;
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_reads * 2 * vecs) return;
    const long long w = id / vecs;
    const int v = (int)(id % vecs);
    const long long r = w >> 1;
    const int len = min(max(wlen[r], 0), stride);
    const uint8_t *tab = s_tab + ((w & 1) ? 16 : 0);
    const uint2 in = ((const uint2 *)(((w & 1) ? tail3p : win5p) + r * stride4))[v];
    const uint32_t src[2] = {in.x, in.y};
    uint32_t dst[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t o = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = v * 16 + q * 4 + b;                          // base position; nibble i of the slot
            const uint32_t cls = (src[q >> 1] >> (4 * ((q & 1) * 4 + b))) & 15u;
            o |= (i < len ? (uint32_t)tab[cls] : 0u) << (8 * b);
        }
        dst[q] = o;
    }
    ((uint4 *)codes)[id] = make_uint4(dst[0], dst[1], dst[2], dst[3]);
}

// 4-bit windows -> ASCII windows of representative bytes (one byte per class), for the generic kernels.
__global__ void k_unpack4(const uint8_t *__restrict__ packed, int stride4, const int32_t *__restrict__ wlen, long long n,
                          const uint8_t *__restrict__ rep, uint8_t *__restrict__ ascii)
{
    const int stride = 2 * stride4;
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n * stride) return;
    const long long r = id / stride;
    const int i = (int)(id % stride);
    const int len = min(max(wlen[r], 0), stride);
    const uint32_t cls = (packed[r * stride4 + (i >> 1)] >> (4 * (i & 1))) & 15u;
    ascii[id] = i < len ? rep[cls] : (uint8_t)0;
}

// ---------------------------------------------------------------------------------------------------
// host side: build the packed-kernel tables
// ---------------------------------------------------------------------------------------------------

inline void fast_plan_free(FastPlan &fp)
{
    if (fp.slab) cudaFree(fp.slab);
    fp.slab = nullptr; fp.slab_bytes = 0;
    if (fp.rowinfo) cudaFree(fp.rowinfo);
    if (fp.taskmeta) cudaFree(fp.taskmeta);
    if (fp.perm) cudaFree(fp.perm);
    fp.rowinfo = fp.taskmeta = fp.perm = nullptr; fp.rowinfo_bytes = fp.taskmeta_bytes = fp.perm_bytes = 0;
    for (auto *sub : fp.subsets) { if (sub->dev) cudaFree(sub->dev); delete sub; }
    fp.subsets.clear();
}

inline void fast_adapter_prepare(FastPlan &fp, const qcb_tables *h)
{
    fp.adapter_ok = false;
    const int g = h->adapter_open, nc = h->amat_size;
    if (h->adapter_open != h->adapter_extend || g <= 0) return;
    if (h->max_align_length > kFastMaxStride) return;
    int smax = 0, mlen = 0;
    for (int i = 0; i < nc * nc; ++i) {
        if (h->amat[i] + 2 * g < 0) return;
        smax = std::max(smax, h->amat[i]);
    }
    fp.a_seq.clear();
    for (int L = 0; L < h->n_layouts; ++L) {
        int len = h->adapter_off[L + 1] - h->adapter_off[L];
        if (len < 1 || len > 104) return;
        mlen = std::max(mlen, len);
        fp.a_seq.emplace_back(h->adapter_seq + h->adapter_off[L], h->adapter_seq + h->adapter_off[L + 1]);
    }
    if (smax * mlen + (kFastMaxStride + mlen) * g + 64 >= 32768) return;
    if (255 - kFastMaxStride < 0 || mlen > 250) return;
    fp.a_gap = g; fp.a_codes = nc;
    fp.a_mat.assign(h->amat, h->amat + nc * nc);
    fp.a_map.assign(h->amap, h->amap + 256);
    fp.adapter_ok = true;
}

inline int fast_plan_build(FastPlan &fp, const qcb_tables *h, int sm_count)
{
    fp.sm_count = sm_count;
    fp.adapter_ok = fp.barcode_ok = false;
    fp.group_pairs.clear();
    if (h->mode == QCB_MODE_SIMPLE) return 0;       // simple mode runs on the generic kernels (needs end_query per barcode)
    fast_adapter_prepare(fp, h);
    fp.barcode_ok = false;
    const int g = h->barcode_open;
    const int nc = h->bmat_size;
    // --- scoring preconditions: linear gap, non-negative shifted scores, values fit 15 bits ---
    if (h->barcode_open != h->barcode_extend || g <= 0) return 0;
    if (nc * kProfWords * 4 > kProfPairBytes || nc > 8) return 0;     // a pair's profile block holds <= 7 code rows
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
        if (smax * tlen + (kFastMaxStride + tlen) * g + 64 >= 2048) return 0;       // F / G are packed into 11 bits
        G.u = u; G.d = d; G.pad = kCore - core; G.tlen = tlen; G.nb = nb;
        fp.max_pairs = std::max(fp.max_pairs, (nb + 1) / 2);
        fp.group_pairs.push_back((nb + 1) / 2);
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
            for (int pr = 0; pr < npairs; ++pr) {
                const size_t base = profile.size();
                profile.resize(base + kProfPairBytes / 4, 0u);
                for (int code = 0; code < nc; ++code)
                    for (int c = G.pad; c < kCore; ++c) {
                        int ba = 2 * pr, bb = std::min(2 * pr + 1, nb - 1);
                        int ca = key[1 + ba * core + (c - G.pad)], cb = key[1 + bb * core + (c - G.pad)];
                        uint32_t lo = (uint32_t)(h->bmat[code * nc + ca] + 2 * g);
                        uint32_t hi = (uint32_t)(h->bmat[code * nc + cb] + 2 * g);
                        profile[base + (size_t)code * kProfWords + c] = lo | (hi << 16);
                    }
            }
        }
        G.prof_off = set_off[found];
    }
    const size_t profile_bytes = profile.size() * 4;
    std::vector<int32_t> sprime(nc * nc);
    for (int i = 0; i < nc * nc; ++i) sprime[i] = h->bmat[i] + 2 * g;

    // k_context tables of shifted scores, right-aligned in ncol columns (0 in the dead columns); shared prefix in the low
    // half, reversed shared suffix in the high half.  Pair form ctx_tab[group][code F][code G][ncol] (one word per cell)
    // while it stays small, else ctx_tab[group][F | G][code][ncol] (two words per cell).
    int max_ctx = 0;
    for (const FastGroup &G : groups) max_ctx = std::max(max_ctx, std::max(G.u, G.d));
    const int ncol = max_ctx <= 12 ? 12 : 16;
#ifndef QCB_CTX_PAIR_LIMIT
#define QCB_CTX_PAIR_LIMIT (64 * 1024)
#endif
    fp.ctx_pair = (size_t)ng * nc * nc * ncol * 4 <= QCB_CTX_PAIR_LIMIT;
    const size_t grp_words = (size_t)(fp.ctx_pair ? nc * nc : 2 * nc) * ncol;
    std::vector<uint32_t> ctx_tab((size_t)ng * grp_words, 0u);
    for (int gi = 0; gi < ng; ++gi) {
        const FastGroup &G = groups[gi];
        for (int code = 0; code < nc; ++code)
            for (int c = 0; c < ncol; ++c) {
                const int jf = c - (ncol - G.u) + 1, jg = c - (ncol - G.d) + 1;
                const uint32_t ef = jf >= 1 ? (uint32_t)sprime[code * nc + ctx[G.up_off + jf - 1]] : 0u;
                const uint32_t eg = jg >= 1 ? (uint32_t)sprime[code * nc + ctx[G.down_off + G.d - jg]] << 16 : 0u;
                if (fp.ctx_pair) {
                    for (int other = 0; other < nc; ++other) {
                        ctx_tab[(size_t)gi * grp_words + ((size_t)code * nc + other) * ncol + c] |= ef;      // code = F's base
                        ctx_tab[(size_t)gi * grp_words + ((size_t)other * nc + code) * ncol + c] |= eg;      // code = G's base
                    }
                } else {
                    ctx_tab[(size_t)gi * grp_words + (size_t)code * ncol + c] = ef;
                    ctx_tab[(size_t)gi * grp_words + (size_t)(nc + code) * ncol + c] = eg;
                }
            }
    }
    fp.ctx_ncol = ncol;
    fp.context_smem = ctx_tab.size() * 4 + (size_t)kCtxWarps * (kCtxEarlyRows * kRowTile * 4 + kRows * kRowTile);
    if (fp.context_smem > 200 * 1024) return 0;

    // one device slab: profile | groups | ctx_tab
    size_t o_prof = 0;
    size_t o_grp = (o_prof + profile_bytes + 255) / 256 * 256;
    size_t o_tab = (o_grp + groups.size() * sizeof(FastGroup) + 255) / 256 * 256;
    size_t total = o_tab + ctx_tab.size() * 4;
    std::vector<uint8_t> slab(total, 0);
    memcpy(slab.data() + o_tab, ctx_tab.data(), ctx_tab.size() * 4);
    memcpy(slab.data() + o_prof, profile.data(), profile_bytes);
    memcpy(slab.data() + o_grp, groups.data(), groups.size() * sizeof(FastGroup));
    if (cudaMalloc(&fp.slab, total) != cudaSuccess) { fp.error = "cudaMalloc failed"; return 1; }
    fp.slab_bytes = total;
    if (cudaMemcpy(fp.slab, slab.data(), total, cudaMemcpyHostToDevice) != cudaSuccess) { fp.error = "cudaMemcpy failed"; return 1; }
    const uint8_t *dptr = (const uint8_t *)fp.slab;
    fp.dev.profile = (const uint32_t *)(dptr + o_prof);
    fp.dev.profile_bytes = (int32_t)profile_bytes;
    fp.dev.groups = (const FastGroup *)(dptr + o_grp);
    fp.ctx_tab = (const uint32_t *)(dptr + o_tab);
    fp.dev.n_codes = nc;
    fp.dev.gap = g;
    fp.dev.n_groups = ng;
    // Rows of a region: the whole window on the epi2me full-window branch, otherwise barcode + 2 x extension + 1
    // (extract_barcode_region) -- except for the odd end_query whose Python slice wraps around.  Dual mode always
    // extracts, so it runs a first launch with a small row tile and a second one (full tile) for the exceptions.
    fp.short_rows = 0;
    {
        int longest = 0;
        for (int i = 0; i < h->n_layouts * 2; ++i) longest = std::max(longest, h->bc_len[i]);
        const int rows = longest + 2 * std::max(0, h->barcode_extension) + 2;
        fp.bucket_rows = std::min(rows, kRows);
        if (h->mode == QCB_MODE_DUAL && rows < kRows / 2) fp.short_rows = rows;
    }
    // All core sets resident while they leave room for three CTAs per SM (explicit kits: one set; dual: two); plans with
    // many sets (`-k auto`: every kit's) keep only the set of the tile at hand in shared memory.
    fp.one_set = profile_bytes > 64 * 1024;
    fp.profile_smem = fp.one_set ? (size_t)fp.max_pairs * kProfPairBytes : profile_bytes;
    fp.barcode_smem = fp.profile_smem + 2 * (size_t)kRows * kRowTile * 4 + kBarcodeSlotBytes;   // one tile per iteration always fits
    if (fp.barcode_smem > 220 * 1024) return 0;
    // Opt every packed kernel into the device's full dynamic shared memory once.  The attribute is a per-device, per-
    // kernel maximum shared by all plans of the process, so it must never be lowered to one plan's own need.
    {
        int dev = 0, optin = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
        if ((size_t)optin < fp.barcode_smem || (size_t)optin < fp.context_smem) return 0;
        fp.smem_optin = optin;
        if (cudaFuncSetAttribute(k_barcode_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) != cudaSuccess ||
            cudaFuncSetAttribute(k_context<12, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) != cudaSuccess ||
            cudaFuncSetAttribute(k_context<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) != cudaSuccess ||
            cudaFuncSetAttribute(k_context<12, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) != cudaSuccess ||
            cudaFuncSetAttribute(k_context<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
    }
    fp.barcode_ok = true;
    return 0;
}

inline int adapter_class_columns(int len) { return len <= 48 ? 48 : (len <= 56 ? 56 : (len <= 64 ? 64 : 104)); }

inline AdapterSubset *adapter_subset(FastPlan &fp, const int32_t *h_subset, int n_subset, cudaStream_t st)
{
    std::vector<int32_t> key(h_subset, h_subset + n_subset);
    for (auto *sub : fp.subsets)
        if (sub->key == key) return sub;
    AdapterSubset *sub = new AdapterSubset();
    sub->key = key;
    const int nc = fp.a_codes, g = fp.a_gap;
    // Templates are grouped by length class (48 / 56 / 64 / 104 register columns) and paired inside their class, shortest
    // first, so that a 39-nt adapter does not pay for the 90-nt one next to it in the subset (`-k auto`: 12 layouts).
    for (int NC : {48, 56, 64, 104}) {
        std::vector<int> slots;
        for (int i = 0; i < n_subset; ++i)
            if (adapter_class_columns((int)fp.a_seq[key[i]].size()) == NC) slots.push_back(i);
        if (slots.empty()) continue;
        std::stable_sort(slots.begin(), slots.end(), [&](int a, int b) { return fp.a_seq[key[a]].size() < fp.a_seq[key[b]].size(); });
        AdapterClass cls;
        cls.nc_cols = NC; cls.row_words = NC + 4; cls.npairs = ((int)slots.size() + 1) / 2;
        cls.profile_bytes = (size_t)cls.npairs * nc * cls.row_words * 4;
        cls.offset = (sub->host.size() + 255) / 256 * 256;
        sub->host.resize(cls.offset + cls.profile_bytes + (size_t)cls.npairs * 16, 0);
        uint32_t *prof = (uint32_t *)(sub->host.data() + cls.offset);
        int32_t *meta = (int32_t *)(sub->host.data() + cls.offset + cls.profile_bytes);
        for (int q = 0; q < cls.npairs; ++q) {
            const int slot_lo = slots[2 * q];
            const bool has_hi = 2 * q + 1 < (int)slots.size();
            const int slot_hi = has_hi ? slots[2 * q + 1] : slot_lo;
            const std::vector<uint8_t> &A = fp.a_seq[key[slot_lo]], &B = fp.a_seq[key[slot_hi]];
            meta[q * 4 + 0] = (int)A.size(); meta[q * 4 + 1] = (int)B.size();
            meta[q * 4 + 2] = has_hi ? 1 : 0; meta[q * 4 + 3] = slot_lo | (slot_hi << 16);
            for (int code = 0; code < nc; ++code)
                for (int c = 0; c < NC; ++c) {
                    int ja = c - (NC - (int)A.size()), jb = c - (NC - (int)B.size());
                    uint32_t lo = ja >= 0 ? (uint32_t)(fp.a_mat[code * nc + fp.a_map[A[ja]]] + 2 * g) : 0u;
                    uint32_t hi = jb >= 0 ? (uint32_t)(fp.a_mat[code * nc + fp.a_map[B[jb]]] + 2 * g) : 0u;
                    prof[((size_t)q * nc + code) * cls.row_words + c] = lo | (hi << 16);
                }
        }
        sub->classes.push_back(cls);
    }
    sub->bytes = sub->host.size();
    if (cudaMalloc(&sub->dev, sub->bytes) != cudaSuccess) { delete sub; return nullptr; }
    if (cudaMemcpyAsync(sub->dev, sub->host.data(), sub->bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) {
        cudaFree(sub->dev); delete sub; return nullptr;
    }
    fp.subsets.push_back(sub);
    return sub;
}

template <int NC>
inline int launch_adapter_fast(FastPlan &fp, AdapterSubset *sub, const AdapterClass &cls, const uint8_t *codes, int stride,
                               const int32_t *wlen, int wshift, long long n_windows, int n_subset, int32_t *ad_score,
                               int32_t *ad_end, cudaStream_t st)
{
    const size_t smem = cls.profile_bytes + (size_t)kAdapterWarps * kRows * kTile;
    // opt into the device's full dynamic shared memory (a per-device, per-kernel maximum shared by all plans: never lower it)
    bool &configured = fp.adapter_smem_configured[NC <= 48 ? 0 : (NC <= 56 ? 1 : (NC <= 64 ? 2 : 3))];
    if (!configured) {
        int dev = 0, optin = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 1;
        if (cudaFuncSetAttribute(k_adapter_fast<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) != cudaSuccess) return 1;
        configured = true;
    }
    const long long n_tiles = (n_windows + kTile - 1) / kTile;
    const long long n_tasks = n_tiles * cls.npairs;
    int per_sm = adapter_blocks(NC);
    while (per_sm > 1 && per_sm * smem > 200 * 1024) --per_sm;
    const int grid = (int)std::min<long long>((n_tasks + kAdapterWarps - 1) / kAdapterWarps, (long long)fp.sm_count * per_sm);
    if (grid <= 0) return 0;
    const uint8_t *base = (const uint8_t *)sub->dev + cls.offset;
    k_adapter_fast<NC><<<grid, kAdapterWarps * 32, smem, st>>>((const uint32_t *)base, (int)(cls.profile_bytes / 4), cls.row_words,
                                                             fp.a_codes, (const int4 *)(base + cls.profile_bytes),
                                                             cls.npairs, codes, stride, wlen, wshift, n_windows, n_subset, fp.a_gap,
                                                             ad_score, ad_end);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

inline int fast_adapter_stage(FastPlan &fp, const DevTables &, const uint8_t *codes, int stride, const int32_t *wlen, int wshift,
                              long long n_windows, const int32_t *h_subset, int n_subset, int32_t *ad_score, int32_t *ad_end,
                              cudaStream_t st, long long *launches)
{
    AdapterSubset *sub = adapter_subset(fp, h_subset, n_subset, st);
    if (!sub) return 1;
    for (const AdapterClass &cls : sub->classes)
        if (cls.profile_bytes + (size_t)kAdapterWarps * kRows * kTile > 200 * 1024) return 2;    // caller falls back
    for (const AdapterClass &cls : sub->classes) {
        int rc;
        if (cls.nc_cols == 48) rc = launch_adapter_fast<48>(fp, sub, cls, codes, stride, wlen, wshift, n_windows, n_subset, ad_score, ad_end, st);
        else if (cls.nc_cols == 56) rc = launch_adapter_fast<56>(fp, sub, cls, codes, stride, wlen, wshift, n_windows, n_subset, ad_score, ad_end, st);
        else if (cls.nc_cols == 64) rc = launch_adapter_fast<64>(fp, sub, cls, codes, stride, wlen, wshift, n_windows, n_subset, ad_score, ad_end, st);
        else rc = launch_adapter_fast<104>(fp, sub, cls, codes, stride, wlen, wshift, n_windows, n_subset, ad_score, ad_end, st);
        if (rc) return rc;
        ++*launches;
    }
    return 0;
}

// Shared-context columns of every (window, set) task -> plan-owned rowinfo / taskmeta buffers.
inline int fast_context_stage(FastPlan &fp, const DevTables &t, const uint8_t *codes, int stride, const int32_t *wlen,
                              long long n_windows, const WindowSel *sel, cudaStream_t st, long long *launches)
{
    const int dual = t.mode == QCB_MODE_DUAL ? 1 : 0;
    const long long n_tasks = dual ? 2 * n_windows : n_windows;
    if (n_tasks <= 0) return 0;
    const long long n_tiles = (n_tasks + kRowTile - 1) / kRowTile;
    const size_t need_rows = (size_t)n_tiles * kRows * kRowTile * 4, need_meta = (size_t)n_tiles * kRowTile * sizeof(int4);
    if (need_rows > fp.rowinfo_bytes) {
        if (fp.rowinfo) cudaFree(fp.rowinfo);
        fp.rowinfo = nullptr; fp.rowinfo_bytes = 0;
        if (cudaMalloc(&fp.rowinfo, need_rows + need_rows / 8) != cudaSuccess) return 1;
        fp.rowinfo_bytes = need_rows + need_rows / 8;
    }
    if (need_meta > fp.taskmeta_bytes) {
        if (fp.taskmeta) cudaFree(fp.taskmeta);
        fp.taskmeta = nullptr; fp.taskmeta_bytes = 0;
        if (cudaMalloc(&fp.taskmeta, need_meta + need_meta / 8) != cudaSuccess) return 1;
        fp.taskmeta_bytes = need_meta + need_meta / 8;
    }
    const size_t need_perm = (size_t)n_tiles * kRowTile * 4 + 16;
    if (need_perm > fp.perm_bytes) {
        if (fp.perm) cudaFree(fp.perm);
        fp.perm = nullptr; fp.perm_bytes = 0;
        if (cudaMalloc(&fp.perm, need_perm + need_perm / 8) != cudaSuccess) return 1;
        fp.perm_bytes = need_perm + need_perm / 8;
    }
    uint32_t *rowinfo = (uint32_t *)fp.rowinfo;
    int4 *taskmeta = (int4 *)fp.taskmeta;
    uint32_t *perm = (uint32_t *)fp.perm;
    unsigned int *counters = (unsigned int *)((uint8_t *)fp.perm + (size_t)n_tiles * kRowTile * 4);
    {
        if (cudaMemsetAsync(counters, 0, 8, st) != cudaSuccess) return 1;
        k_task_order<<<(unsigned)((n_tasks + 255) / 256), 256, 0, st>>>(t, sel, n_windows, dual, fp.bucket_rows, counters, perm);
        ++*launches;
    }
    {
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / fp.context_smem));
        const int cgrid = (int)std::min<long long>((n_tiles + kCtxWarps - 1) / kCtxWarps, (long long)fp.sm_count * per_sm);
#define QCB_LAUNCH_CONTEXT(NCOL, PAIR)                                                                                   \
    k_context<NCOL, PAIR><<<cgrid, kCtxWarps * 32, fp.context_smem, st>>>(fp.dev, t, fp.ctx_tab, codes, stride, wlen, n_windows, sel, \
                                                                          dual, perm, rowinfo, taskmeta)
        if (fp.ctx_ncol == 12) { if (fp.ctx_pair) QCB_LAUNCH_CONTEXT(12, true); else QCB_LAUNCH_CONTEXT(12, false); }
        else { if (fp.ctx_pair) QCB_LAUNCH_CONTEXT(16, true); else QCB_LAUNCH_CONTEXT(16, false); }
#undef QCB_LAUNCH_CONTEXT
        ++*launches;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// Core columns of every barcode against every task (needs fast_context_stage on the same stream first).
inline int fast_barcode_stage(FastPlan &fp, const DevTables &t, long long n_windows, int bmax0, int bslots, int32_t *bc_score,
                              cudaStream_t st, long long *launches)
{
    const int dual = t.mode == QCB_MODE_DUAL ? 1 : 0;
    const long long n_tasks = dual ? 2 * n_windows : n_windows;
    if (n_tasks <= 0) return 0;
    const long long n_tiles = (n_tasks + kRowTile - 1) / kRowTile;
    const uint32_t *rowinfo = (const uint32_t *)fp.rowinfo;
    const int4 *taskmeta = (const int4 *)fp.taskmeta;
    {
        cudaFuncAttributes attr;
        if (cudaFuncGetAttributes(&attr, k_barcode_fast) != cudaSuccess) return 1;
        const size_t profile_bytes = fp.profile_smem;
        const int passes = fp.short_rows > 0 ? 2 : 1;
        for (int pass = 0; pass < passes; ++pass) {
            const int rows_min = pass == 0 ? 1 : fp.short_rows;
            const int rows_cap = (passes == 2 && pass == 0) ? fp.short_rows : kRows;
            // Every warp takes one (tile, barcode pair) item per round: the warp count (<= kBarcodeMaxWarps) that wastes
            // the fewest warp-rounds on the plan's largest set, preferring more warps.  When a smaller set of the plan
            // still leaves warps idle at that count (dual: 12 pairs on 8 warps), a CTA iteration takes two tiles -- if
            // that fills the rounds and costs no CTA per SM (shared memory: profile | double-buffered row tiles | slots).
            // Plans that keep one core set resident stay on one tile (their set switches are CTA-wide).
            int warps = kBarcodeMaxWarps, best_waste = 1 << 30;
            for (int wc = kBarcodeMaxWarps; wc >= 4; --wc) {
                const int waste = (fp.max_pairs + wc - 1) / wc * wc - fp.max_pairs;
                if (waste < best_waste) { best_waste = waste; warps = wc; }
            }
            const size_t regs_per_cta = (size_t)warps * 32 * ((attr.numRegs + 7) / 8 * 8);
            const int by_regs = std::max<int>(1, std::min<int>((int)(65536 / regs_per_cta), std::min(8, 2048 / (warps * 32))));
            auto smem_for = [&](int t) { return profile_bytes + 2 * (size_t)t * rows_cap * kRowTile * 4 + kBarcodeSlotBytes; };
            static const int forced_tpi = getenv("QCB_BC_TPI") ? atoi(getenv("QCB_BC_TPI")) : 0;     // A/B switch (1 or 2)
            int tpi = 1;
            bool idle = false, filled = true;
            for (int np : fp.group_pairs) {
                idle = idle || np % warps != 0;
                filled = filled && (2 * np) % warps == 0;
            }
            if (idle && filled && !fp.one_set && (int)((220 * 1024) / (smem_for(2) + 1024)) >= by_regs) tpi = 2;
            if (forced_tpi >= 1 && forced_tpi <= kMaxTilesPerIter) tpi = forced_tpi;
            const size_t smem = smem_for(tpi);
            const int ctas_per_sm = std::max(1, std::min((int)((220 * 1024) / (smem + 1024)), by_regs));
            const long long n_groups = (n_tiles + tpi - 1) / tpi;
            int grid = (int)std::min<long long>(n_groups, (long long)fp.sm_count * ctas_per_sm);
            const unsigned int *bucket_counts = (const unsigned int *)((const uint8_t *)fp.perm + (size_t)n_tiles * kRowTile * 4);
            k_barcode_fast<<<grid, warps * 32, smem, st>>>(fp.dev, n_windows, dual, bmax0, bslots, rows_min, rows_cap, fp.one_set ? 1 : 0,
                                                           tpi, (int)profile_bytes, rowinfo, taskmeta, bc_score, bucket_counts,
                                                           passes == 2 ? pass : -1);
            ++*launches;
        }
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace qcb
