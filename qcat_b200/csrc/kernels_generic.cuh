// Generic kernels: any gap model (affine), any substitution matrix, any window / template length up to
// kMaxTemplate.  One thread per (window, template) pair, int32 arithmetic, DP column state in local memory.
// They are the always-correct path (custom qcatConfig, qcb_sg_batch primitive) and the stages of the fused
// pipeline that are not DP (orientation, template selection, region geometry, two-end decision).
#pragma once

#include "plan.h"

namespace qcb {

constexpr int kNegInf = INT32_MIN / 2;

// Semi-global alignment, all four ends free, affine gaps (parasail `sg`; reference call sites
// scanner_base.py:111-117 and :214-218).  q = read window (rows), ref = template (columns).
// End-cell rule: R = max of the last row (first column attaining it), C = max of the last column (first
// row attaining it); C > R -> end in the last column at row iC; else end in the last row, except when the
// first row maximum is the corner cell, where end_query = iC.
__device__ __forceinline__ void sg_affine(const uint8_t *__restrict__ q, int n,
                                          const uint8_t *__restrict__ ref, int m,
                                          int open, int extend,
                                          const int32_t *__restrict__ matrix, int msize,
                                          const uint8_t *__restrict__ mapper,
                                          int &score, int &end_query, int &end_ref)
{
    if (n <= 0 || m <= 0) { score = 0; end_query = -1; end_ref = -1; return; }
    int32_t H[kMaxTemplate + 1];
    int32_t F[kMaxTemplate + 1];
    uint8_t rc[kMaxTemplate];
    for (int j = 0; j < m; ++j) rc[j] = mapper[ref[j]];
    for (int j = 0; j <= m; ++j) { H[j] = 0; F[j] = kNegInf; }
    int col_max = INT32_MIN, col_arg = -1;
    for (int i = 1; i <= n; ++i) {
        const int32_t *row = matrix + msize * mapper[q[i - 1]];
        int diag = 0, left = 0, E = kNegInf;
        for (int j = 1; j <= m; ++j) {
            int up = H[j];
            int f = max(F[j] - extend, up - open);
            int e = max(E - extend, left - open);
            int h = max(max(diag + row[rc[j - 1]], e), f);
            F[j] = f; E = e; diag = up; H[j] = h; left = h;
        }
        if (left > col_max) { col_max = left; col_arg = i; }
    }
    int row_max = INT32_MIN, row_arg = -1;
    for (int j = 1; j <= m; ++j)
        if (H[j] > row_max) { row_max = H[j]; row_arg = j; }
    if (col_max > row_max) { score = col_max; end_query = col_arg - 1; end_ref = m - 1; }
    else {
        score = row_max; end_ref = row_arg - 1; end_query = n - 1;
        if (row_arg == m) end_query = col_arg - 1;
    }
}

__device__ __forceinline__ void load_matrix_smem(int32_t *s_mat, uint8_t *s_map, const int32_t *mat, int msize,
                                                 const uint8_t *map)
{
    for (int i = threadIdx.x; i < msize * msize; i += blockDim.x) s_mat[i] = mat[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_map[i] = map[i];
    __syncthreads();
}

// qcb_sg_batch: every query against every reference.
__global__ void k_sg_batch(const uint8_t *__restrict__ queries, const int32_t *__restrict__ qoff, int nq,
                           const uint8_t *__restrict__ refs, const int32_t *__restrict__ roff, int nr,
                           int open, int extend, const int32_t *__restrict__ matrix, int msize,
                           const uint8_t *__restrict__ mapper,
                           int32_t *__restrict__ score, int32_t *__restrict__ end_query, int32_t *__restrict__ end_ref)
{
    __shared__ int32_t s_mat[kMaxMatrix * kMaxMatrix];
    __shared__ uint8_t s_map[256];
    load_matrix_smem(s_mat, s_map, matrix, msize, mapper);
    long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)nq * nr) return;
    int qi = (int)(id / nr), ri = (int)(id % nr);
    int sc, eq, er;
    sg_affine(queries + qoff[qi], qoff[qi + 1] - qoff[qi], refs + roff[ri], roff[ri + 1] - roff[ri],
              open, extend, s_mat, msize, s_map, sc, eq, er);
    score[id] = sc; end_query[id] = eq; end_ref[id] = er;
}

// parasail `sg_stats` (scanner_base.py:20-26; used by the simple scanner through find_highest_scoring_barcode
// (compute_identity=True, :106-123) and by align_adapter_identity :144-188): sg_affine plus the number of exact matches,
// of positive-scoring columns and the length of the alignment, carried along the predecessor the recurrence follows
// (gap opened only when strictly better than extended; diagonal when >= both gap states, else F when F >= E, else E --
// the oracle's qo_sg_stats states the same rule and its "parity unpinned" status).
struct StatCell { int32_t h, m, s, l; };

__device__ __forceinline__ void sg_affine_stats(const uint8_t *__restrict__ q, int n, const uint8_t *__restrict__ ref, int m,
                                                int open, int extend, const int32_t *__restrict__ matrix, int msize,
                                                const uint8_t *__restrict__ mapper,
                                                int &score, int &end_query, int &end_ref, int &matches, int &similar, int &length)
{
    matches = 0; similar = 0; length = 0;
    if (n <= 0 || m <= 0) { score = 0; end_query = -1; end_ref = -1; return; }
    StatCell H[kMaxTemplate + 1];
    StatCell F[kMaxTemplate + 1];
    uint8_t rc[kMaxTemplate];
    for (int j = 0; j < m; ++j) rc[j] = mapper[ref[j]];
    for (int j = 0; j <= m; ++j) { H[j] = StatCell{0, 0, 0, 0}; F[j] = StatCell{kNegInf, 0, 0, 0}; }
    StatCell col_best{INT32_MIN, 0, 0, 0};
    int col_arg = -1;
    for (int i = 1; i <= n; ++i) {
        const int c1 = mapper[q[i - 1]];
        const int32_t *row = matrix + msize * c1;
        StatCell diag = H[0], left{0, 0, 0, 0}, E{kNegInf, 0, 0, 0};
        for (int j = 1; j <= m; ++j) {
            const StatCell up = H[j];
            StatCell f, e, h;
            if (up.h - open > F[j].h - extend) { f = up; f.h = up.h - open; } else { f = F[j]; f.h = F[j].h - extend; }
            f.l += 1;
            if (left.h - open > E.h - extend) { e = left; e.h = left.h - open; } else { e = E; e.h = E.h - extend; }
            e.l += 1;
            const int sub = row[rc[j - 1]];
            const int d = diag.h + sub;
            if (d >= e.h && d >= f.h) h = StatCell{d, diag.m + (c1 == rc[j - 1] ? 1 : 0), diag.s + (sub > 0 ? 1 : 0), diag.l + 1};
            else if (f.h >= e.h) h = f;
            else h = e;
            F[j] = f; E = e; diag = up; H[j] = h; left = h;
        }
        if (left.h > col_best.h) { col_best = left; col_arg = i; }
    }
    int row_max = INT32_MIN, row_arg = -1;
    for (int j = 1; j <= m; ++j)
        if (H[j].h > row_max) { row_max = H[j].h; row_arg = j; }
    StatCell end;
    if (col_best.h > row_max) { score = col_best.h; end_query = col_arg - 1; end_ref = m - 1; end = col_best; }
    else {
        score = row_max; end_ref = row_arg - 1; end_query = n - 1; end = H[row_arg];
        if (row_arg == m) { end_query = col_arg - 1; end = col_best; }
    }
    matches = end.m; similar = end.s; length = end.l;
}

// qcb_sg_stats_batch: every query against every reference, with the alignment statistics.
__global__ void k_sg_stats_batch(const uint8_t *__restrict__ queries, const int32_t *__restrict__ qoff, int nq,
                                 const uint8_t *__restrict__ refs, const int32_t *__restrict__ roff, int nr,
                                 int open, int extend, const int32_t *__restrict__ matrix, int msize,
                                 const uint8_t *__restrict__ mapper, int32_t *__restrict__ out /* [6][nq * nr] */)
{
    __shared__ int32_t s_mat[kMaxMatrix * kMaxMatrix];
    __shared__ uint8_t s_map[256];
    load_matrix_smem(s_mat, s_map, matrix, msize, mapper);
    const long long pairs = (long long)nq * nr;
    long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= pairs) return;
    int qi = (int)(id / nr), ri = (int)(id % nr);
    int sc, eq, er, mt, sm, ln;
    sg_affine_stats(queries + qoff[qi], qoff[qi + 1] - qoff[qi], refs + roff[ri], roff[ri + 1] - roff[ri],
                    open, extend, s_mat, msize, s_map, sc, eq, er, mt, sm, ln);
    out[id] = sc; out[pairs + id] = eq; out[2 * pairs + id] = er; out[3 * pairs + id] = mt; out[4 * pairs + id] = sm;
    out[5 * pairs + id] = ln;
}

// Window orientation: wins[2r] = read[:W], wins[2r+1] = revcomp(read[-W:]) (scanner_base.py:223-244, utils.py:26-27).
__global__ void k_orient(const uint8_t *__restrict__ win5, const uint8_t *__restrict__ tail3, int stride,
                         const int32_t *__restrict__ wlen, long long n_reads, const uint8_t *__restrict__ comp,
                         uint8_t *__restrict__ wins)
{
    long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = n_reads * 2 * stride;
    if (id >= total) return;
    long long w = id / stride;
    int i = (int)(id % stride);
    long long r = w >> 1;
    int len = min(max(wlen[r], 0), stride);
    uint8_t v = 0;
    if (i < len) v = (w & 1) ? comp[tail3[r * stride + (len - 1 - i)]] : win5[r * stride + i];
    wins[id] = v;
}

// Adapter stage, generic: one thread per (window, subset entry) (align_adapter, scanner_base.py:191-220).
__global__ void k_adapter_generic(DevTables t, const uint8_t *__restrict__ wins, int stride,
                                  const int32_t *__restrict__ wlen, int wshift, long long n_windows,
                                  const int32_t *__restrict__ subset, int n_subset,
                                  int32_t *__restrict__ ad_score, int32_t *__restrict__ ad_end)
{
    __shared__ int32_t s_mat[kMaxMatrix * kMaxMatrix];
    __shared__ uint8_t s_map[256];
    load_matrix_smem(s_mat, s_map, t.amat, t.amat_size, t.amap);
    long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_windows * n_subset) return;
    long long w = id / n_subset;
    int s = (int)(id % n_subset);
    int L = subset[s];
    int n = wlen[w >> wshift];
    int sc, eq, er;
    sg_affine(wins + w * stride, n, t.adapter_seq + t.adapter_off[L], t.adapter_off[L + 1] - t.adapter_off[L],
              t.a_open, t.a_extend, s_mat, t.amat_size, s_map, sc, eq, er);
    ad_score[id] = sc; ad_end[id] = eq;
}

// ---- long windows (--detect-middle scans read[W:-W], scanner_base.py:479-519) -----------------------------------------
// One thread per (window, template) walks tens of thousands of rows serially.  Rows are therefore cut into chunks that
// run in parallel: chunk c owns rows (c CH, (c+1) CH] and starts its DP `ov` rows earlier from an all-zero row (the
// state of a window that begins there).  With every matrix entry in [smin, smax] and every gap character costing at
// least gmin > 0, a path that reaches a cell more than ov = m (1 + (smax - smin) / gmin) rows below its start scores less
// than the pure-diagonal path every cell has (>= smin m), so it never decides a maximum: all cells of the owned rows are
// exact, and the per-chunk results below combine to exactly sg_affine's answer.
struct LongPart {
    int32_t col_max, col_arg;        // max of H[i][m] over the chunk's own rows, first row attaining it (1-based, global)
};

__global__ void k_adapter_long(DevTables t, const uint8_t *__restrict__ wins, int stride, const int32_t *__restrict__ wlen,
                               long long n_windows, const int32_t *__restrict__ subset, int n_subset, int CH, int ov, int nch,
                               LongPart *__restrict__ part, int2 *__restrict__ last_row)
{
    __shared__ int32_t s_mat[kMaxMatrix * kMaxMatrix];
    __shared__ uint8_t s_map[256];
    load_matrix_smem(s_mat, s_map, t.amat, t.amat_size, t.amap);
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_windows * n_subset * nch) return;
    const int c = (int)(id % nch);
    const long long ws = id / nch;                   // (window, subset entry)
    const long long w = ws / n_subset;
    const int L = subset[(int)(ws % n_subset)];
    const int n = wlen[w];
    const int own = c * CH;                          // rows (own, end] belong to this chunk
    if (own >= n && !(n <= 0 && c == 0)) { part[id].col_max = INT32_MIN; part[id].col_arg = -1; return; }
    const uint8_t *ref = t.adapter_seq + t.adapter_off[L];
    const int m = t.adapter_off[L + 1] - t.adapter_off[L];
    if (n <= 0 || m <= 0) { part[id].col_max = INT32_MIN; part[id].col_arg = -1; last_row[ws] = make_int2(INT32_MIN, -1); return; }
    const int end = min(n, own + CH);
    const int start = max(0, own - ov);
    const uint8_t *q = wins + w * stride;
    const int open = t.a_open, extend = t.a_extend, msize = t.amat_size;
    int32_t H[kMaxTemplate + 1];
    int32_t F[kMaxTemplate + 1];
    uint8_t rc[kMaxTemplate];
    for (int j = 0; j < m; ++j) rc[j] = s_map[ref[j]];
    for (int j = 0; j <= m; ++j) { H[j] = 0; F[j] = kNegInf; }
    int col_max = INT32_MIN, col_arg = -1;
    for (int i = start + 1; i <= end; ++i) {
        const int32_t *row = s_mat + msize * s_map[q[i - 1]];
        int diag = 0, left = 0, E = kNegInf;
        for (int j = 1; j <= m; ++j) {
            int up = H[j];
            int f = max(F[j] - extend, up - open);
            int e = max(E - extend, left - open);
            int h = max(max(diag + row[rc[j - 1]], e), f);
            F[j] = f; E = e; diag = up; H[j] = h; left = h;
        }
        if (i > own && left > col_max) { col_max = left; col_arg = i; }
    }
    part[id].col_max = col_max; part[id].col_arg = col_arg;
    if (end == n) {                                  // the chunk holding the last row also reports its maximum
        int row_max = INT32_MIN, row_arg = -1;
        for (int j = 1; j <= m; ++j)
            if (H[j] > row_max) { row_max = H[j]; row_arg = j; }
        last_row[ws] = make_int2(row_max, row_arg);
    }
}

// sg_affine's end-cell rule over the chunk results of one (window, template).
__global__ void k_adapter_long_combine(DevTables t, const int32_t *__restrict__ wlen, long long n_windows,
                                       const int32_t *__restrict__ subset, int n_subset, int nch,
                                       const LongPart *__restrict__ part, const int2 *__restrict__ last_row,
                                       int32_t *__restrict__ ad_score, int32_t *__restrict__ ad_end)
{
    const long long ws = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ws >= n_windows * n_subset) return;
    const long long w = ws / n_subset;
    const int L = subset[(int)(ws % n_subset)];
    const int n = wlen[w];
    const int m = t.adapter_off[L + 1] - t.adapter_off[L];
    if (n <= 0 || m <= 0) { ad_score[ws] = 0; ad_end[ws] = -1; return; }
    int col_max = INT32_MIN, col_arg = -1;
    for (int c = 0; c < nch; ++c) {
        const LongPart p = part[ws * nch + c];
        if (p.col_arg >= 0 && p.col_max > col_max) { col_max = p.col_max; col_arg = p.col_arg; }
    }
    const int2 r = last_row[ws];
    int score, end;
    if (col_max > r.x) { score = col_max; end = col_arg - 1; }
    else { score = r.x; end = n - 1; if (r.y == m) end = col_arg - 1; }
    ad_score[ws] = score; ad_end[ws] = end;
}

// Python slice semantics seq[start:stop] -> [lo, hi).
__device__ __forceinline__ void py_slice(int start, int stop, int n, int &lo, int &hi)
{
    if (start < 0) { start += n; if (start < 0) start = 0; } else if (start > n) start = n;
    if (stop < 0) { stop += n; if (stop < 0) stop = 0; } else if (stop > n) stop = n;
    if (stop < start) stop = start;
    lo = start; hi = stop;
}

// extract_barcode_region, scanner_base.py:29-60.
__device__ __forceinline__ void barcode_region(const DevTables &t, int L, int k, int n, int stop, int &lo, int &hi)
{
    int alen = t.adapter_off[L + 1] - t.adapter_off[L];
    int e = stop - (alen - t.bc_end[L * 2 + k]) + 1;
    int s = e - t.bc_len[L * 2 + k];
    s -= min(t.ext, s);
    e += min(t.ext, n - e);
    py_slice(s, e + 1, n, lo, hi);
}

// find_best_adapter_template's arg-max over the subset (scanner_base.py:330-359): strict '<', first wins,
// nothing found -> index -1.  Returns the position inside the subset.  kit >= 0 restricts the scan to the subset
// entries whose layout belongs to that kit (the per-batch kit of the auto-kit flow, scanner_base.py:526-529).
__device__ __forceinline__ void best_template(const DevTables &t, const int32_t *subset, int n_subset, int n,
                                              const int32_t *ad_score, const int32_t *ad_end,
                                              int &best_idx, int &best_end, double &best_score,
                                              const int32_t *kit_of_layout = nullptr, int kit = -1)
{
    best_score = -1.0; best_end = -1; best_idx = -1;
    if (n <= 0) return;
    for (int s = 0; s < n_subset; ++s) {
        int L = subset[s];
        if (kit >= 0 && kit_of_layout[L] != kit) continue;
        if (t.adapter_off[L + 1] - t.adapter_off[L] <= 0) continue;
        double norm = (double)ad_score[s] * 100.0 / t.denom[L];
        if (best_score < norm) { best_score = norm; best_idx = s; best_end = ad_end[s]; }
    }
}

// Template choice + barcode-region geometry per window (scanner_epi2me.py:57-82, scanner_dual.py:57-110).
// Auto-kit flow (batch_kit != nullptr): the layouts considered for read r are those of kit batch_kit[r / batch_size],
// the kit its CLI batch voted for -- the adapter scores of all layouts are already in ad_score / ad_end.
__global__ void k_select(DevTables t, const int32_t *__restrict__ wlen, int wshift, long long n_windows,
                         const int32_t *__restrict__ subset, int n_subset,
                         const int32_t *__restrict__ ad_score, const int32_t *__restrict__ ad_end,
                         WindowSel *__restrict__ sel,
                         const int32_t *__restrict__ kit_of_layout, const int32_t *__restrict__ batch_kit, int batch_size,
                         long long read_offset)
{
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_windows) return;
    int n = wlen[w >> wshift];
    const int kit = batch_kit ? batch_kit[(read_offset + (w >> wshift)) / batch_size] : -1;
    int idx, end; double ascore;
    best_template(t, subset, n_subset, n, ad_score + w * n_subset, ad_end + w * n_subset, idx, end, ascore, kit_of_layout, kit);
    if (idx < 0) {                                      // Python negative index: -1 -> last layout (of the kit)
        idx = n_subset - 1;
        if (kit >= 0) while (idx > 0 && kit_of_layout[subset[idx]] != kit) --idx;
    }
    int L = subset[idx];
    WindowSel s;
    s.layout = L; s.end_query = end; s.lo1 = 0; s.hi1 = 0; s.full = 0; s.pad = 0;
    if (t.mode == QCB_MODE_EPI2ME) {
        if (ascore > 90.0 || t.is_double[L]) barcode_region(t, L, 0, n, end, s.lo0, s.hi0);
        else { s.lo0 = 0; s.hi0 = min(n, t.W); s.full = 1; }
    } else {
        barcode_region(t, L, 0, n, end, s.lo0, s.hi0);
        barcode_region(t, L, 1, n, end, s.lo1, s.hi1);
    }
    sel[w] = s;
}

// Simple mode (scanner_simple.py:47-92): no adapter stage; every barcode is aligned to the whole window.
__global__ void k_select_simple(const int32_t *__restrict__ wlen, int wshift, long long n_windows, WindowSel *__restrict__ sel)
{
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_windows) return;
    WindowSel s;
    s.layout = 0; s.end_query = -1; s.lo0 = 0; s.hi0 = max(wlen[w >> wshift], 0); s.lo1 = 0; s.hi1 = 0; s.full = 1; s.pad = 0;
    sel[w] = s;
}

// Barcode stage, generic: one thread per (window, template slot); slots [0, B0) are set 0 and, in dual mode,
// [bmax0, bmax0 + B1) are set 1 (find_highest_scoring_barcode's inner alignment, scanner_base.py:111-117).
__global__ void k_barcode_generic(DevTables t, const uint8_t *__restrict__ wins, int stride, long long n_windows,
                                  const WindowSel *__restrict__ sel, int bmax0, int bslots,
                                  int32_t *__restrict__ bc_score, int32_t *__restrict__ bc_endq)
{
    __shared__ int32_t s_mat[kMaxMatrix * kMaxMatrix];
    __shared__ uint8_t s_map[256];
    load_matrix_smem(s_mat, s_map, t.bmat, t.bmat_size, t.bmap);
    long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_windows * bslots) return;
    long long w = id / bslots;
    int slot = (int)(id % bslots);
    WindowSel s = sel[w];
    int k = slot >= bmax0 ? 1 : 0;
    int b = k ? slot - bmax0 : slot;
    int g = t.group[s.layout * 2 + k];
    if (g < 0) return;
    int gsize = t.group_off[g + 1] - t.group_off[g];
    if (b >= gsize) return;
    int lo = k ? s.lo1 : s.lo0, hi = k ? s.hi1 : s.hi0;
    int tm = t.group_off[g] + b;
    int sc, eq, er;
    sg_affine(wins + w * stride + lo, hi - lo, t.tmpl_seq + t.tmpl_off[tm], t.tmpl_off[tm + 1] - t.tmpl_off[tm],
              t.b_open, t.b_extend, s_mat, t.bmat_size, s_map, sc, eq, er);
    bc_score[id] = sc;
    if (bc_endq) bc_endq[id] = eq;                       // simple mode keeps the winner's end_query (scanner_base.py:131)
}

struct EndResult {
    int layout;      // -1 = adapter None
    int barcode;     // -1 = None
    int ident;
    double score;
    int adapter_end;
};

__device__ __forceinline__ EndResult empty_end()
{
    EndResult r; r.layout = -1; r.barcode = -1; r.ident = -1; r.score = 0.0; r.adapter_end = 0; return r;
}

// The sequential selection rule of find_highest_scoring_barcode (scanner_base.py:119-134): first barcode
// always taken; afterwards replace when the running best is 0.0 ("not max_score") or strictly smaller.
__device__ __forceinline__ void pick_barcode(const DevTables &t, int g, int rlen, const int32_t *scores,
                                             int &best, double &best_score)
{
    best = -1; best_score = 0.0;
    if (rlen <= 0 || g < 0) return;
    int base = t.group_off[g], cnt = t.group_off[g + 1] - base;
    const int tl_all = t.group_tlen[g];
    if (tl_all > 0) {
        // every template of the group has the same length: score * 100.0 / tlen is strictly increasing in the integer
        // score and is 0.0 exactly for score 0, so the rule can run on the integers and divide once
        // (scores of a window start on a 16-byte boundary: slot counts are multiples of 4 -- one load per four barcodes)
        bool have = false; int mx = 0, arg = -1;
        const int4 *quads = (const int4 *)scores;
        for (int q = 0; q * 4 < cnt; ++q) {
            const int4 v = quads[q];
            const int sc4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int b = q * 4 + e, sc = sc4[e];
                if (b < cnt && (!have || mx == 0 || mx < sc)) { have = true; mx = sc; arg = b; }
            }
        }
        best = arg; best_score = (double)mx * 100.0 / (1.0 * (double)tl_all);
        return;
    }
    bool have = false; double mx = 0.0; int arg = -1;
    for (int b = 0; b < cnt; ++b) {
        int tl = t.tmpl_off[base + b + 1] - t.tmpl_off[base + b];
        double sc = (double)scores[b] * 100.0 / (1.0 * (double)tl);
        if (!have || mx == 0.0 || mx < sc) { have = true; mx = sc; arg = b; }
    }
    best = arg; best_score = mx;
}

__device__ __forceinline__ EndResult scan_result(const DevTables &t, const WindowSel &s, int n,
                                                 const int32_t *scores, int bmax0, const int32_t *endq)
{
    EndResult r;
    int L = s.layout;
    int g0 = t.group[L * 2];
    int b0; double s0;
    pick_barcode(t, g0, s.hi0 - s.lo0, scores, b0, s0);
    if (t.mode == QCB_MODE_SIMPLE) {
        // scanner_simple.py:67-92: what scan() calls `identity` is the third value returned at scanner_base.py:141 --
        // the score -- so the threshold compares the score; the adapter is None, adapter_end is the barcode's end_query
        if (s0 < t.min_quality) return empty_end();
        r.layout = -1; r.barcode = b0; r.score = s0;
        r.ident = b0 >= 0 ? t.tmpl_ident[t.group_off[g0] + b0] : -1;
        r.adapter_end = b0 >= 0 ? endq[b0] : -1;
        return r;
    }
    if (t.mode == QCB_MODE_EPI2ME) {
        r.layout = L; r.barcode = b0; r.score = s0;
        r.ident = b0 >= 0 ? t.tmpl_ident[t.group_off[g0] + b0] : -1;
        r.adapter_end = min(s.end_query + t.trim_offset[L], n);          // scanner_epi2me.py:135-137
        return r;
    }
    int g1 = t.group[L * 2 + 1];
    int b1; double s1;
    pick_barcode(t, g1, s.hi1 - s.lo1, scores + bmax0, b1, s1);
    if (b0 >= 0 && b1 >= 0) {                                            // scanner_dual.py:131-144
        int n1 = t.group_off[g1 + 1] - t.group_off[g1];
        r.layout = L; r.barcode = b0 * n1 + b1;
        r.ident = t.tmpl_ident[t.group_off[g0] + b0] * 65536 + t.tmpl_ident[t.group_off[g1] + b1];
        r.score = s0 < s1 ? s0 : (s1 < s0 ? s1 : s0);
        r.adapter_end = s.end_query;
        return r;
    }
    return empty_end();
}

// BarcodeScanner.detect_barcode's two-end logic (scanner_base.py:542-602), one thread per read.
__global__ void k_finalize(DevTables t, const int32_t *__restrict__ wlen, const int64_t *__restrict__ read_len,
                           long long n_reads, const WindowSel *__restrict__ sel,
                           const int32_t *__restrict__ bc_score, int bmax0, int bslots,
                           qcb_result *__restrict__ out, const int32_t *__restrict__ bc_endq)
{
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    int n = wlen[r];
    EndResult d5 = scan_result(t, sel[2 * r], n, bc_score + (2 * r) * bslots, bmax0, bc_endq ? bc_endq + (2 * r) * bslots : nullptr);
    long long trim5 = 0;
    if (d5.adapter_end > 0) trim5 = d5.adapter_end;
    if (d5.score < t.min_quality) d5 = empty_end();
    EndResult d3 = scan_result(t, sel[2 * r + 1], n, bc_score + (2 * r + 1) * bslots, bmax0,
                               bc_endq ? bc_endq + (2 * r + 1) * bslots : nullptr);
    long long trim3 = read_len[r];
    if (d3.layout >= 0 && d3.adapter_end > 0) trim3 -= d3.adapter_end;
    if (d3.score < t.min_quality) d3 = empty_end();

    int which = -1; double best_score = 0.0;
    if (d5.score > best_score) { best_score = d5.score; which = 0; }
    if (d3.score > best_score) { best_score = d3.score; which = 1; }
    EndResult res = empty_end(); int exit_status = 1;
    if (which >= 0) {
        res = which ? d3 : d5; exit_status = 0;
        if (d5.barcode >= 0 && d3.barcode >= 0 && d5.score >= 60.0 && d3.score >= 60.0 && d5.ident != d3.ident) {
            res = empty_end(); exit_status = 1002;
        }
    }
    if (trim3 < trim5) trim5 = 0;
    qcb_result o;
    o.layout = res.layout; o.barcode = res.barcode; o.barcode_score = res.score; o.adapter_end = res.adapter_end;
    o.trim5p = (int32_t)trim5; o.trim3p = (int32_t)trim3; o.exit_status = exit_status;
    out[r] = o;
}

// BarcodeScanner.scan for stand-alone windows: the per-window record before any two-end logic.
__global__ void k_scan_out(DevTables t, const int32_t *__restrict__ wlen, long long n_windows,
                           const WindowSel *__restrict__ sel, const int32_t *__restrict__ bc_score, int bmax0, int bslots,
                           qcb_result *__restrict__ out, const int32_t *__restrict__ bc_endq)
{
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_windows) return;
    EndResult d = scan_result(t, sel[w], wlen[w], bc_score + w * bslots, bmax0, bc_endq ? bc_endq + w * bslots : nullptr);
    qcb_result o;
    o.layout = d.layout; o.barcode = d.barcode; o.barcode_score = d.score; o.adapter_end = d.adapter_end;
    o.trim5p = 0; o.trim3p = 0; o.exit_status = (d.layout < 0 && d.barcode < 0) ? 1 : 0;
    out[w] = o;
}

// detect_kit's vote (scanner_base.py:618-642): best template over ALL layouts at each end, higher score wins.
__global__ void k_kit_vote(DevTables t, const int32_t *__restrict__ wlen, long long n_reads,
                           const int32_t *__restrict__ subset, int n_subset,
                           const int32_t *__restrict__ ad_score, const int32_t *__restrict__ ad_end,
                           int32_t *__restrict__ vote)
{
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    int n = wlen[r];
    int i5, e5, i3, e3; double s5, s3;
    best_template(t, subset, n_subset, n, ad_score + (2 * r) * n_subset, ad_end + (2 * r) * n_subset, i5, e5, s5);
    best_template(t, subset, n_subset, n, ad_score + (2 * r + 1) * n_subset, ad_end + (2 * r + 1) * n_subset, i3, e3, s3);
    if (i5 < 0) i5 += n_subset;
    if (i3 < 0) i3 += n_subset;
    vote[r] = subset[(s5 > s3) ? i5 : i3];
}

// get_most_abundant_kits per CLI batch (scanner_base.py:657-678, :714-722): one CTA per batch of `batch_size` reads.
// The kit named by most votes wins; among kits with the same count the one seen first in the batch (the reference
// counts in a dict -- insertion order -- and sorts stably by count).
constexpr int kMaxKits = 256;

__global__ void k_batch_kit(const int32_t *__restrict__ vote, long long n_reads, int batch_size,
                            const int32_t *__restrict__ kit_of_layout, int n_kits, int32_t *__restrict__ batch_kit)
{
    __shared__ unsigned cnt[kMaxKits], first[kMaxKits];
    for (int k = threadIdx.x; k < n_kits; k += blockDim.x) { cnt[k] = 0; first[k] = 0xffffffffu; }
    __syncthreads();
    const long long lo = (long long)blockIdx.x * batch_size, hi = min(n_reads, lo + (long long)batch_size);
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const int k = kit_of_layout[vote[i]];
        atomicAdd(&cnt[k], 1u);
        atomicMin(&first[k], (unsigned)(i - lo));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int best = -1;
        for (int k = 0; k < n_kits; ++k)
            if (cnt[k] > 0 && (best < 0 || cnt[k] > cnt[best] || (cnt[k] == cnt[best] && first[k] < first[best]))) best = k;
        batch_kit[blockIdx.x] = best;
    }
}

// Per-barcode counts (the histogram behind cli.py:386-405).  Counts are first gathered per CTA in shared memory (the
// bins are few and hot: 97 for a 96-barcode kit), then added to the global vector with one atomic per non-empty bin.
constexpr int kHistSharedBins = 4096;

__global__ void k_histogram(const qcb_result *__restrict__ res, long long n_reads,
                            const int32_t *__restrict__ layout_bin_base, unsigned long long *__restrict__ counts, int n_bins)
{
    __shared__ unsigned int s_cnt[kHistSharedBins];
    const bool use_shared = n_bins <= kHistSharedBins;
    if (use_shared) {
        for (int i = threadIdx.x; i < n_bins; i += blockDim.x) s_cnt[i] = 0;
        __syncthreads();
    }
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += (long long)gridDim.x * blockDim.x) {
        const qcb_result o = res[r];
        int bin = 0;
        if (o.barcode >= 0) bin = 1 + (o.layout >= 0 ? layout_bin_base[o.layout] : 0) + o.barcode;   // layout -1: simple mode
        if (bin < 0 || bin >= n_bins) continue;
        if (use_shared) atomicAdd(&s_cnt[bin], 1u);
        else atomicAdd(counts + bin, 1ULL);
    }
    if (use_shared) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_bins; i += blockDim.x)
            if (s_cnt[i]) atomicAdd(counts + i, (unsigned long long)s_cnt[i]);
    }
}

}  // namespace qcb
