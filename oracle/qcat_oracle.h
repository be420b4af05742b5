/*
 * qcat_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's EPI2ME / dual barcode-detection hot path
 * (nanoporetech/qcat 1.1.0).  It is the checker the CUDA path is compared against; it is
 * never imported, linked or executed by the product (qcat_b200/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Parity status: the orchestration (everything above the alignment primitive) is validated
 * against the UNMODIFIED reference Python run in the build container (oracle/refshim +
 * tests/golden/make_golden.py).  The alignment primitive itself lives in the third-party
 * parasail C library (PyPI `parasail`, version unpinned by the reference's setup.py:16-21,
 * not vendored, not installable offline); qo_sg() restates its published semi-global
 * algorithm.  The recurrence is pinned by the reference's own known-answer tests
 * (test_barcode.py:291-304 end_query == 101, and every barcode-name KAT), but the
 * END-POSITION TIE RULE among equal-scoring end cells is restated from parasail 2.x's
 * sg_striped source from memory and is pinned by no reference test: "parity unpinned"
 * for that one rule (isolated in qo_sg(), see the comment there).
 */
#ifndef QCAT_ORACLE_H
#define QCAT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Flattened, read-only description of a scanner: layouts of the selected kit(s), their barcode
 * template groups and the two scoring schemes.  Built by qcat_b200/tables.py from the reference's
 * (or the mirror's) AdapterLayout / Barcode / qcatConfig objects. */
typedef struct {
    /* qcatConfig (config.py:12-26) */
    int32_t max_align_length;   /* W, 150 */
    int32_t barcode_extension;  /* extracted_barcode_extension, 11 */
    int32_t adapter_open, adapter_extend;   /* gap_open / gap_extend, 2 / 2 (scanner_base.py:216-217) */
    int32_t barcode_open, barcode_extend;   /* hard-coded 1 / 1 (scanner_base.py:115-116) */
    int32_t amat_size;          /* adapter matrix is amat_size x amat_size (7 for "ATGCNX"+wildcard) */
    const int32_t *amat;
    const uint8_t *amap;        /* 256-entry byte -> matrix index */
    int32_t bmat_size;          /* barcode matrix (6 for "ATGCN"+wildcard) */
    const int32_t *bmat;
    const uint8_t *bmap;
    const uint8_t *comp;        /* 256-entry complement table (utils.py:26-27) */
    /* scanner */
    int32_t mode;               /* 0 = epi2me (scanner_epi2me.py), 1 = dual (scanner_dual.py), 2 = simple (scanner_simple.py:
                                 * template group 0 = the bare barcodes of self.barcodes, one placeholder layout) */
    double min_quality;         /* 58 epi2me / 60 dual unless overridden */
    /* layouts, in self.layouts order */
    int32_t n_layouts;
    const int32_t *adapter_off; /* [n_layouts+1] offsets into adapter_seq */
    const uint8_t *adapter_seq; /* N-masked adapter sequences, upper case */
    const double *denom;        /* [n_layouts] (alen-bclen)*match + bclen*nmatch (scanner_base.py:308-310) */
    const int32_t *bc_end;      /* [n_layouts*2] placeholder end (layout.py:72-96), -1 if absent */
    const int32_t *bc_len;      /* [n_layouts*2] placeholder length, 0 if absent */
    const int32_t *group;       /* [n_layouts*2] template group of barcode set k, -1 if the set is None */
    const int32_t *trim_offset; /* [n_layouts] */
    const int32_t *is_double;   /* [n_layouts] barcode_set_2 is not None (layout.py:240-248) */
    /* template groups: up_context + barcode + down_context, in barcode-set order */
    int32_t n_groups;
    const int32_t *group_off;   /* [n_groups+1] range of template indices */
    const int32_t *tmpl_off;    /* [n_templates+1] offsets into tmpl_seq */
    const uint8_t *tmpl_seq;
    const int32_t *tmpl_ident;  /* [n_templates] equality class of Barcode.id (for the 1002 rule) */
} qo_tables;

/* One record per read; mirrors build_return_dict (scanner_base.py:362-390). */
typedef struct {
    int32_t layout;        /* index into the tables' layouts of result['adapter'], -1 = None */
    int32_t barcode;       /* template index inside its group, -1 = None; dual: idx1 * n2 + idx2 */
    double  barcode_score; /* result['barcode_score'] */
    int32_t adapter_end;
    int32_t trim5p;
    int32_t trim3p;
    int32_t exit_status;
} qo_result;

/* parasail sg (semi-global, all ends free) as called at scanner_base.py:111-117 and :214-218.
 * s1 = read window (query), s2 = template (ref).  Returns score, end_query, end_ref. */
void qo_sg(const uint8_t *s1, int n, const uint8_t *s2, int m, int open, int extend,
           const int32_t *matrix, int msize, const uint8_t *mapper,
           int32_t *score, int32_t *end_query, int32_t *end_ref);

/* parasail sg_stats (scanner_base.py:20-26, :106-123, :168-172): qo_sg plus matches / similar / length of the path the
 * recurrence follows ("parity unpinned" for those three, see qcat_oracle.c). */
void qo_sg_stats(const uint8_t *s1, int n, const uint8_t *s2, int m, int open, int extend,
                 const int32_t *matrix, int msize, const uint8_t *mapper,
                 int32_t *score, int32_t *end_query, int32_t *end_ref,
                 int32_t *matches, int32_t *similar, int32_t *length);

/* find_best_adapter_template (scanner_base.py:313-359) over layouts[subset[0..n_subset)]. */
void qo_find_best_adapter_template(const qo_tables *t, const int32_t *subset, int n_subset,
                                   const uint8_t *win, int n,
                                   int32_t *best_index, int32_t *best_end, double *best_score);

/* BarcodeScanner.detect_barcode (scanner_base.py:521-604) for reads given as their two windows:
 * win5 = read[:W] and tail3 = read[-W:] (NOT yet reverse-complemented), both wlen[i] = min(len, W)
 * bytes in slots of `stride` bytes.  subset = layout indices ("kits" at :526-529). */
void qo_detect(const qo_tables *t, const uint8_t *win5, const uint8_t *tail3, int stride,
               const int32_t *wlen, const int64_t *read_len, int64_t n_reads,
               const int32_t *subset, int n_subset, qo_result *out, int n_threads);

/* BarcodeScanner.scan (scanner_epi2me.py:33-144 / scanner_dual.py:35-146) on n_windows already oriented windows of
 * any length in slots of `stride` bytes: records with trims 0 and exit_status 0, or the empty record (layout -1,
 * exit_status 1) where the reference returns empty_return_dict(). */
void qo_scan(const qo_tables *t, const uint8_t *windows, int64_t stride, const int32_t *wlen, int64_t n_windows,
             const int32_t *subset, int n_subset, qo_result *out, int n_threads);

/* detect_kit's per-read vote (scanner_base.py:618-678): index of the layout of the higher-scoring end
 * over ALL layouts ("adapter_1" at :669). */
void qo_kit_vote(const qo_tables *t, const uint8_t *win5, const uint8_t *tail3, int stride,
                 const int32_t *wlen, int64_t n_reads, int32_t *vote_layout, int n_threads);

/* Algorithmic DP cells of the reference path for these reads; detect_kit duplicates excluded. */
int64_t qo_count_cells(const qo_tables *t, const uint8_t *win5, const uint8_t *tail3, int stride,
                       const int32_t *wlen, int64_t n_reads, const int32_t *subset, int n_subset,
                       int64_t *full_window_branches);

#ifdef __cplusplus
}
#endif
#endif
