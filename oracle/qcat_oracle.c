/*
 * qcat_oracle.c -- TEST INFRASTRUCTURE ONLY (see qcat_oracle.h for the parity statement).
 *
 * Plain-C restatement of qcat's EPI2ME / dual detection path.  Every function cites the
 * reference file:line it follows (paths relative to the reference checkout, qcat 1.1.0).
 * Scalar int32 / IEEE double arithmetic, no SIMD, no shortcuts: this is the definition of
 * "correct" for the CUDA path, not a performance target.
 */
#include "qcat_oracle.h"

#include <stdlib.h>
#include <string.h>

#define QO_NEG_INF (INT32_MIN / 2)

static inline int imax(int a, int b) { return a > b ? a : b; }

/*
 * parasail `sg` (semi-global, free end gaps on both sequences at both ends), affine gaps,
 * as invoked by scanner_base.py:111-117 (barcodes) and :214-218 (adapters) through
 * parasail.sg_striped_32.  Published algorithm (Daily 2016, parasail):
 *   H[0][j] = H[i][0] = 0
 *   E[i][j] = max(E[i][j-1] - extend, H[i][j-1] - open)     (gap in the query / along the template)
 *   F[i][j] = max(F[i-1][j] - extend, H[i-1][j] - open)
 *   H[i][j] = max(H[i-1][j-1] + M[s1[i]][s2[j]], E[i][j], F[i][j])
 * a gap of length k costs open + (k-1)*extend.  Characters are mapped through the matrix's
 * 256-entry mapper (case-insensitive, unknown -> wildcard row/column of zeros).
 *
 * End cell (the one rule not pinned by any reference test -- "parity unpinned"):
 *   R = max_j H[n][j] (last query row), first j attaining it; C = max_i H[i][m] (last template
 *   column), first i attaining it.  C > R: end = (iC, m).  Otherwise end = (n, jR), except when the
 *   first row maximum sits in the corner (jR == m): then end_query = iC (parasail's
 *   "*t == score && end_ref == s2Len-1 && temp < end_query" update in its last-column scan).
 * Positions are returned 0-based.
 */
void qo_sg(const uint8_t *s1, int n, const uint8_t *s2, int m, int open, int extend,
           const int32_t *matrix, int msize, const uint8_t *mapper,
           int32_t *score, int32_t *end_query, int32_t *end_ref)
{
    if (n <= 0 || m <= 0) { *score = 0; *end_query = -1; *end_ref = -1; return; }
    int32_t *H = (int32_t *)malloc(sizeof(int32_t) * (size_t)(m + 1) * 2);
    int32_t *F = H + (m + 1);
    int32_t *c2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)m);
    for (int j = 0; j < m; ++j) c2[j] = mapper[s2[j]];
    for (int j = 0; j <= m; ++j) { H[j] = 0; F[j] = QO_NEG_INF; }

    int32_t col_max = INT32_MIN, col_arg = -1;      /* C, iC over the last column */
    for (int i = 1; i <= n; ++i) {
        const int32_t *row = matrix + (size_t)msize * mapper[s1[i - 1]];
        int32_t diag = H[0];            /* H[i-1][0] = 0 */
        int32_t left = 0;               /* H[i][0] */
        int32_t E = QO_NEG_INF;
        H[0] = 0;
        for (int j = 1; j <= m; ++j) {
            int32_t up = H[j];
            int32_t f = imax(F[j] - extend, up - open);
            int32_t e = imax(E - extend, left - open);
            int32_t h = imax(imax(diag + row[c2[j - 1]], e), f);
            F[j] = f; E = e;
            diag = up;
            H[j] = h;
            left = h;
        }
        if (H[m] > col_max) { col_max = H[m]; col_arg = i; }
    }
    int32_t row_max = INT32_MIN, row_arg = -1;      /* R, jR over the last row */
    for (int j = 1; j <= m; ++j)
        if (H[j] > row_max) { row_max = H[j]; row_arg = j; }

    if (col_max > row_max) {
        *score = col_max; *end_query = col_arg - 1; *end_ref = m - 1;
    } else {
        *score = row_max; *end_ref = row_arg - 1; *end_query = n - 1;
        if (row_arg == m) *end_query = col_arg - 1;   /* corner exception: col_max == row_max here */
    }
    free(c2);
    free(H);
}

/*
 * parasail `sg_stats` (parasail.sg_stats_striped_32, bound at scanner_base.py:20-26, used by
 * find_highest_scoring_barcode(compute_identity=True) :106-123 -- the simple scanner -- and by align_adapter_identity
 * :144-188): the same recurrence as qo_sg plus, carried along the path the recurrence takes, the number of exact
 * matches, of positive-scoring ("similar") columns and the alignment length.
 * Score and end cell: identical to qo_sg (same end-cell rule, same "parity unpinned" status).
 * matches / similar / length depend on which of several equal-scoring predecessors the recurrence follows; restated
 * from parasail 2.x's scalar sg_stats from memory, pinned by NO reference test ("parity unpinned"): a gap is opened
 * only when opening is strictly better than extending; H takes the diagonal when it is >= both gap states, else the
 * F (gap in the template direction, from the row above) state when F >= E, else E.  Exact matches compare the
 * mapped character codes.  Nothing on qcat's result path reads these counters: the simple scanner's `identity` is
 * overwritten by the score before it is used (scanner_base.py:141 returns max_score in its place).
 */
void qo_sg_stats(const uint8_t *s1, int n, const uint8_t *s2, int m, int open, int extend,
                 const int32_t *matrix, int msize, const uint8_t *mapper,
                 int32_t *score, int32_t *end_query, int32_t *end_ref,
                 int32_t *matches, int32_t *similar, int32_t *length)
{
    *matches = 0; *similar = 0; *length = 0;
    if (n <= 0 || m <= 0) { *score = 0; *end_query = -1; *end_ref = -1; return; }
    typedef struct { int32_t h, m, s, l; } cell;
    cell *H = (cell *)calloc((size_t)(m + 1), sizeof(cell));       /* H[i-1][*] then H[i][*] */
    cell *F = (cell *)calloc((size_t)(m + 1), sizeof(cell));
    cell *lastcol = (cell *)calloc((size_t)(n + 1), sizeof(cell)); /* H[i][m] for the end-cell rule */
    int32_t *c2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)m);
    for (int j = 0; j < m; ++j) c2[j] = mapper[s2[j]];
    for (int j = 0; j <= m; ++j) F[j].h = QO_NEG_INF;
    int32_t col_max = INT32_MIN, col_arg = -1;
    for (int i = 1; i <= n; ++i) {
        const int c1 = mapper[s1[i - 1]];
        const int32_t *row = matrix + (size_t)msize * c1;
        cell diag = H[0];                          /* H[i-1][0]: zero border */
        cell left; left.h = 0; left.m = 0; left.s = 0; left.l = 0;
        cell E; E.h = QO_NEG_INF; E.m = 0; E.s = 0; E.l = 0;
        for (int j = 1; j <= m; ++j) {
            cell up = H[j];
            cell f, e, h;
            if (up.h - open > F[j].h - extend) { f = up; f.h = up.h - open; } else { f = F[j]; f.h = F[j].h - extend; }
            f.l += 1;
            if (left.h - open > E.h - extend) { e = left; e.h = left.h - open; } else { e = E; e.h = E.h - extend; }
            e.l += 1;
            int32_t sub = row[c2[j - 1]];
            int32_t d = diag.h + sub;
            if (d >= e.h && d >= f.h) {
                h.h = d; h.m = diag.m + (c1 == c2[j - 1]); h.s = diag.s + (sub > 0); h.l = diag.l + 1;
            } else if (f.h >= e.h) h = f;
            else h = e;
            F[j] = f; E = e;
            diag = up;
            H[j] = h;
            left = h;
        }
        lastcol[i] = H[m];
        if (H[m].h > col_max) { col_max = H[m].h; col_arg = i; }
    }
    int32_t row_max = INT32_MIN, row_arg = -1;
    for (int j = 1; j <= m; ++j)
        if (H[j].h > row_max) { row_max = H[j].h; row_arg = j; }
    cell end;
    if (col_max > row_max) {
        *score = col_max; *end_query = col_arg - 1; *end_ref = m - 1; end = lastcol[col_arg];
    } else {
        *score = row_max; *end_ref = row_arg - 1; *end_query = n - 1; end = H[row_arg];
        if (row_arg == m) { *end_query = col_arg - 1; end = lastcol[col_arg]; }
    }
    *matches = end.m; *similar = end.s; *length = end.l;
    free(c2); free(lastcol); free(F); free(H);
}

/* Python slice semantics seq[start:stop] on a sequence of length n -> [lo, hi) (possibly empty). */
static void py_slice(int start, int stop, int n, int *lo, int *hi)
{
    if (start < 0) { start += n; if (start < 0) start = 0; } else if (start > n) start = n;
    if (stop < 0) { stop += n; if (stop < 0) stop = 0; } else if (stop > n) stop = n;
    if (stop < start) stop = start;
    *lo = start; *hi = stop;
}

static inline int imin(int a, int b) { return a < b ? a : b; }

/* extract_barcode_region, scanner_base.py:29-60.  Returns the [lo, hi) range inside the window. */
static void extract_barcode_region(const qo_tables *t, int layout, int k, int n, int alignment_stop_ref,
                                   int *lo, int *hi)
{
    int adapter_length = t->adapter_off[layout + 1] - t->adapter_off[layout];
    int barcode_end = t->bc_end[layout * 2 + k];
    int barcode_length = t->bc_len[layout * 2 + k];
    int barcode_end_ref = alignment_stop_ref - (adapter_length - barcode_end) + 1;      /* :51 */
    int barcode_start_ref = barcode_end_ref - barcode_length;                            /* :52 */
    barcode_start_ref -= imin(t->barcode_extension, barcode_start_ref);                  /* :55 */
    barcode_end_ref += imin(t->barcode_extension, n - barcode_end_ref);                  /* :57 */
    py_slice(barcode_start_ref, barcode_end_ref + 1, n, lo, hi);                         /* :60 */
}

/* find_highest_scoring_barcode, scanner_base.py:63-141.  Returns the template index inside the group
 * (-1 = None) and the normalised score; `not max_score` makes 0.0 behave as "unset" (:125). */
static void find_highest_scoring_barcode(const qo_tables *t, int group, const uint8_t *region, int rlen,
                                         int *best, double *best_score)
{
    *best = -1; *best_score = 0.0;                       /* :93-102: (None, 0, 0.0, -1) */
    if (rlen <= 0) return;
    int have = 0; double max_score = 0.0; int max_idx = -1;
    for (int b = t->group_off[group]; b < t->group_off[group + 1]; ++b) {
        int tl = t->tmpl_off[b + 1] - t->tmpl_off[b];
        int32_t sc, eq, er;
        qo_sg(region, rlen, t->tmpl_seq + t->tmpl_off[b], tl, t->barcode_open, t->barcode_extend,
              t->bmat, t->bmat_size, t->bmap, &sc, &eq, &er);                              /* :111-117 */
        double score = (double)sc * 100.0 / (1.0 * (double)tl);                          /* :119 */
        if (!have || max_score == 0.0 || max_score < score) {                           /* :125 */
            have = 1; max_score = score; max_idx = b - t->group_off[group];
        }
    }
    *best = max_idx; *best_score = max_score;
}

void qo_find_best_adapter_template(const qo_tables *t, const int32_t *subset, int n_subset,
                                   const uint8_t *win, int n,
                                   int32_t *best_index, int32_t *best_end, double *best_score)
{
    *best_score = -1.0; *best_end = -1; *best_index = -1;                /* :330-332 */
    if (n_subset <= 0 || n <= 0) return;                                  /* :334 */
    for (int i = 0; i < n_subset; ++i) {
        int L = subset[i];
        int alen = t->adapter_off[L + 1] - t->adapter_off[L];
        if (alen <= 0) continue;                                          /* :343 */
        int32_t sc, eq, er;
        qo_sg(win, n, t->adapter_seq + t->adapter_off[L], alen, t->adapter_open, t->adapter_extend,
              t->amat, t->amat_size, t->amap, &sc, &eq, &er);             /* :214-218 */
        double norm = (double)sc * 100.0 / t->denom[L];                   /* :310 */
        if (*best_score < norm) { *best_score = norm; *best_index = i; *best_end = eq; }   /* :354-357 */
    }
}

typedef struct {
    int layout;       /* -1 = adapter None */
    int barcode;      /* -1 = barcode None */
    int ident;        /* Barcode.id equality class (dual: pair index) */
    double score;
    int adapter_end;
} scan_result;

static scan_result empty_scan(void)      /* empty_return_dict, scanner_base.py:393-407 */
{
    scan_result r; r.layout = -1; r.barcode = -1; r.ident = -1; r.score = 0.0; r.adapter_end = 0; return r;
}

/* BarcodeScannerSimple.scan (scanner_simple.py:47-92): every barcode of self.barcodes (template group 0: the bare
 * barcode sequences, no context) against the whole window with sg_stats; find_highest_scoring_barcode's first-maximum
 * rule (:119-134) also keeps the end_query of the winner (`max_end`).  What scan() calls `identity` is the third
 * value returned at :141 -- max_score, not the identity -- so the threshold compares the score. */
static scan_result scan_window_simple(const qo_tables *t, const uint8_t *win, int n)
{
    int have = 0; double max_score = 0.0; int max_idx = -1, max_end = -1;
    double identity = 0.0;                                     /* :101: (None, 0, 0.0, -1) for an empty window */
    if (n > 0) {
        for (int b = t->group_off[0]; b < t->group_off[1]; ++b) {
            int tl = t->tmpl_off[b + 1] - t->tmpl_off[b];
            int32_t sc, eq, er, mt, sm, ln;
            qo_sg_stats(win, n, t->tmpl_seq + t->tmpl_off[b], tl, t->barcode_open, t->barcode_extend,
                        t->bmat, t->bmat_size, t->bmap, &sc, &eq, &er, &mt, &sm, &ln);
            double score = (double)sc * 100.0 / (1.0 * (double)tl);
            if (!have || max_score == 0.0 || max_score < score) { have = 1; max_score = score; max_idx = b - t->group_off[0]; max_end = eq; }
        }
        identity = max_score;                                  /* :141 */
    }
    if (identity < t->min_quality) return empty_scan();        /* scanner_simple.py:81-82 */
    scan_result r;
    r.layout = -1;                                             /* best_adapter=None */
    r.barcode = max_idx;
    r.ident = max_idx >= 0 ? t->tmpl_ident[t->group_off[0] + max_idx] : -1;
    r.score = max_score;                                       /* q_score = max_score (:139) */
    r.adapter_end = max_end;                                   /* best_adapter_end=barcode_end */
    return r;
}

/* BarcodeScannerEPI2ME.scan (scanner_epi2me.py:33-144) and BarcodeScannerDual.scan (scanner_dual.py:35-146). */
static scan_result scan_window(const qo_tables *t, const int32_t *subset, int n_subset, const uint8_t *win, int n)
{
    if (t->mode == 2) return scan_window_simple(t, win, n);
    int32_t idx, end; double ascore;
    qo_find_best_adapter_template(t, subset, n_subset, win, n, &idx, &end, &ascore);
    int L = subset[idx < 0 ? n_subset + idx : idx];           /* Python negative index: -1 -> last (:64) */
    int lo, hi, b1; double s1;
    scan_result r;
    if (t->mode == 0) {
        if (ascore > 90.0 || t->is_double[L]) {               /* scanner_epi2me.py:74 */
            extract_barcode_region(t, L, 0, n, end, &lo, &hi);
        } else {
            lo = 0; hi = imin(n, t->max_align_length);        /* :82 */
        }
        find_highest_scoring_barcode(t, t->group[L * 2 + 0], win + lo, hi - lo, &b1, &s1);
        /* :104-131: the second set of a double-barcode layout is scored and discarded -- no effect. */
        r.layout = L; r.barcode = b1; r.score = s1;
        r.ident = b1 >= 0 ? t->tmpl_ident[t->group_off[t->group[L * 2]] + b1] : -1;
        r.adapter_end = imin(end + t->trim_offset[L], n);     /* :135-137 */
        return r;
    }
    /* dual */
    int b2; double s2;
    extract_barcode_region(t, L, 0, n, end, &lo, &hi);                                     /* scanner_dual.py:78 */
    find_highest_scoring_barcode(t, t->group[L * 2 + 0], win + lo, hi - lo, &b1, &s1);    /* :96 */
    extract_barcode_region(t, L, 1, n, end, &lo, &hi);                                     /* :106 */
    find_highest_scoring_barcode(t, t->group[L * 2 + 1], win + lo, hi - lo, &b2, &s2);    /* :124 */
    if (b1 >= 0 && b2 >= 0) {                                                              /* :131 */
        int g2 = t->group[L * 2 + 1];
        int n2 = t->group_off[g2 + 1] - t->group_off[g2];
        r.layout = L; r.barcode = b1 * n2 + b2;
        /* Barcode("barcodeAA/BB", "a/b"): ids compare equal iff both component ids do. */
        int i1 = t->tmpl_ident[t->group_off[t->group[L * 2]] + b1];
        int i2 = t->tmpl_ident[t->group_off[g2] + b2];
        r.ident = i1 * 65536 + i2;
        r.score = s1 < s2 ? s1 : (s2 < s1 ? s2 : s1);        /* min(a, b) (:140) */
        r.adapter_end = end;                                  /* :142, no trim_offset */
        return r;
    }
    return empty_scan();                                      /* :146 */
}

static void detect_one(const qo_tables *t, const uint8_t *w5, const uint8_t *tail, int n, int64_t read_len,
                       const int32_t *subset, int n_subset, uint8_t *w3, qo_result *out)
{
    /* 3' window = revcomp(read[-W:]) (scanner_base.py:242, utils.py:26-27) */
    for (int i = 0; i < n; ++i) w3[i] = t->comp[tail[n - 1 - i]];

    scan_result d5 = scan_window(t, subset, n_subset, w5, n);              /* :536 */
    int64_t trim5 = 0;
    if (d5.adapter_end > 0) trim5 = d5.adapter_end;                        /* :542-544 */
    if (d5.score < t->min_quality) d5 = empty_scan();                      /* :546-548 */

    scan_result d3 = scan_window(t, subset, n_subset, w3, n);              /* :555 */
    int64_t trim3 = read_len;                                              /* :561 */
    if (d3.layout >= 0 && d3.adapter_end > 0) trim3 -= d3.adapter_end;     /* :562-564 */
    if (d3.score < t->min_quality) d3 = empty_scan();                      /* :566-568 */

    const scan_result *best = 0; double best_score = 0.0;                  /* :573-580 */
    if (d5.score > best_score) { best_score = d5.score; best = &d5; }
    if (d3.score > best_score) { best_score = d3.score; best = &d3; }

    scan_result res; int exit_status;
    if (!best) { res = empty_scan(); exit_status = 1; }                    /* :582-583 */
    else {
        res = *best; exit_status = 0;
        if (d5.barcode >= 0 && d3.barcode >= 0 && d5.score >= 60.0 && d3.score >= 60.0 &&
            d5.ident != d3.ident) {                                        /* :585-591 */
            res = empty_scan(); exit_status = 1002;
        }
    }
    if (trim3 < trim5) trim5 = 0;                                          /* :600-602 */
    out->layout = res.layout; out->barcode = res.barcode; out->barcode_score = res.score;
    out->adapter_end = res.adapter_end; out->trim5p = (int32_t)trim5; out->trim3p = (int32_t)trim3;
    out->exit_status = exit_status;
}

void qo_detect(const qo_tables *t, const uint8_t *win5, const uint8_t *tail3, int stride,
               const int32_t *wlen, const int64_t *read_len, int64_t n_reads,
               const int32_t *subset, int n_subset, qo_result *out, int n_threads)
{
    (void)n_threads;
    int32_t *all = 0;
    if (!subset || n_subset <= 0) {            /* convenience: NULL = every layout (self.layouts) */
        all = (int32_t *)malloc(sizeof(int32_t) * (size_t)(t->n_layouts > 0 ? t->n_layouts : 1));
        for (int i = 0; i < t->n_layouts; ++i) all[i] = i;
        subset = all; n_subset = t->n_layouts;
    }
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
    {
        uint8_t *w3 = (uint8_t *)malloc((size_t)(t->max_align_length > 0 ? t->max_align_length : 1) + 16);
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < n_reads; ++r)
            detect_one(t, win5 + (size_t)r * stride, tail3 + (size_t)r * stride, wlen[r], read_len[r],
                       subset, n_subset, w3, out + r);
        free(w3);
    }
    free(all);
}

/* BarcodeScanner.scan on stand-alone, already oriented windows of any length (what scan_middle feeds it,
 * scanner_base.py:479-519): the per-window record before any two-end logic (trims 0, exit_status 0 / 1). */
void qo_scan(const qo_tables *t, const uint8_t *windows, int64_t stride, const int32_t *wlen, int64_t n_windows,
             const int32_t *subset, int n_subset, qo_result *out, int n_threads)
{
    (void)n_threads;
    int32_t *all = 0;
    if (!subset || n_subset <= 0) {
        all = (int32_t *)malloc(sizeof(int32_t) * (size_t)(t->n_layouts > 0 ? t->n_layouts : 1));
        for (int i = 0; i < t->n_layouts; ++i) all[i] = i;
        subset = all; n_subset = t->n_layouts;
    }
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads > 0 ? n_threads : 1)
    for (int64_t w = 0; w < n_windows; ++w) {
        scan_result r = scan_window(t, subset, n_subset, windows + (size_t)w * (size_t)stride, wlen[w]);
        out[w].layout = r.layout; out[w].barcode = r.barcode; out[w].barcode_score = r.score;
        out[w].adapter_end = r.adapter_end; out[w].trim5p = 0; out[w].trim3p = 0;
        out[w].exit_status = (r.layout < 0 && r.barcode < 0) ? 1 : 0;
    }
    free(all);
}

void qo_kit_vote(const qo_tables *t, const uint8_t *win5, const uint8_t *tail3, int stride,
                 const int32_t *wlen, int64_t n_reads, int32_t *vote_layout, int n_threads)
{
    (void)n_threads;
    int32_t *all = (int32_t *)malloc(sizeof(int32_t) * (size_t)(t->n_layouts > 0 ? t->n_layouts : 1));
    for (int i = 0; i < t->n_layouts; ++i) all[i] = i;
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
    {
        uint8_t *w3 = (uint8_t *)malloc((size_t)(t->max_align_length > 0 ? t->max_align_length : 1) + 16);
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < n_reads; ++r) {
            int n = wlen[r];
            const uint8_t *tail = tail3 + (size_t)r * stride;
            for (int i = 0; i < n; ++i) w3[i] = t->comp[tail[n - 1 - i]];
            int32_t i5, e5, i3, e3; double s5, s3;
            qo_find_best_adapter_template(t, all, t->n_layouts, win5 + (size_t)r * stride, n, &i5, &e5, &s5);  /* :624 */
            qo_find_best_adapter_template(t, all, t->n_layouts, w3, n, &i3, &e3, &s3);
            if (i5 < 0) i5 += t->n_layouts;                   /* self.layouts[-1] (:630) */
            if (i3 < 0) i3 += t->n_layouts;
            vote_layout[r] = (s5 > s3) ? i5 : i3;             /* :639-642 */
        }
        free(w3);
    }
    free(all);
}

int64_t qo_count_cells(const qo_tables *t, const uint8_t *win5, const uint8_t *tail3, int stride,
                       const int32_t *wlen, int64_t n_reads, const int32_t *subset, int n_subset,
                       int64_t *full_window_branches)
{
    int64_t cells = 0, full = 0;
    uint8_t *w3 = (uint8_t *)malloc((size_t)(t->max_align_length > 0 ? t->max_align_length : 1) + 16);
    for (int64_t r = 0; r < n_reads; ++r) {
        int n = wlen[r];
        const uint8_t *tail = tail3 + (size_t)r * stride;
        for (int i = 0; i < n; ++i) w3[i] = t->comp[tail[n - 1 - i]];
        for (int end = 0; end < 2; ++end) {
            const uint8_t *w = end ? w3 : win5 + (size_t)r * stride;
            if (n <= 0 || n_subset <= 0) continue;
            for (int i = 0; i < n_subset; ++i)
                cells += (int64_t)n * (t->adapter_off[subset[i] + 1] - t->adapter_off[subset[i]]);
            int32_t idx, e; double sc;
            qo_find_best_adapter_template(t, subset, n_subset, w, n, &idx, &e, &sc);
            int L = subset[idx < 0 ? n_subset + idx : idx];
            for (int k = 0; k < (t->mode == 1 ? 2 : 1); ++k) {
                int lo, hi;
                if (t->mode == 0 && !(sc > 90.0 || t->is_double[L])) { lo = 0; hi = imin(n, t->max_align_length); full++; }
                else extract_barcode_region(t, L, k, n, e, &lo, &hi);
                int g = t->group[L * 2 + k];
                if (g < 0) continue;
                for (int b = t->group_off[g]; b < t->group_off[g + 1]; ++b)
                    cells += (int64_t)(hi - lo) * (t->tmpl_off[b + 1] - t->tmpl_off[b]);
            }
        }
    }
    free(w3);
    if (full_window_branches) *full_window_branches = full;
    return cells;
}
