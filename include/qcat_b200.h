/*
 * qcat_b200.h -- C ABI of libqcat_b200.so, the B200 (sm_100a) replacement for qcat's per-read
 * adapter / barcode alignment hot path.
 *
 * What it replaces in the reference (nanoporetech/qcat 1.1.0, paths relative to its checkout):
 *   - the in-process FFI the reference itself uses: parasail-python (ctypes) ->
 *     parasail_sg_striped_32(s1, s1Len, s2, s2Len, open, gap, matrix) called at
 *     qcat/scanner_base.py:111-117 (barcodes) and :214-218 (adapters)      -> qcb_sg_batch()
 *     and parasail_sg_stats_striped_32 (:106-123, :168-172)                 -> qcb_sg_stats_batch()
 *   - BarcodeScannerSimple.scan (scanner_simple.py:47-92)                   -> qcb_detect*() / qcb_scan() on a QCB_MODE_SIMPLE plan
 *   - the Python loops around it: find_best_adapter_template (scanner_base.py:313-359),
 *     extract_barcode_region (:29-60), find_highest_scoring_barcode (:63-141),
 *     BarcodeScannerEPI2ME.scan (scanner_epi2me.py:33-144), BarcodeScannerDual.scan
 *     (scanner_dual.py:35-146) and BarcodeScanner.detect_barcode (scanner_base.py:521-604),
 *     batched over reads as in detect_barcode_batch (:714-733)              -> qcb_detect*(), qcb_detect_auto*()
 *   - detect_kit's per-read vote (scanner_base.py:618-678)                  -> qcb_kit_vote*()
 *   - the per-barcode counts behind the CLI histogram (cli.py:386-405)      -> qcb_histogram_device()
 *   - host only, either side of the path: record iteration (cli.py:235-306), window extraction
 *     (scanner_base.py:223-244), trimming / min-length filter and the three output forms
 *     (cli.py:309-358, :408-442, :521-552)       -> qcb_reader_*(), qcb_fastx_index*(), qcb_pack_windows(),
 *                                                   qcb_format_*(), qcb_write_bins()
 *
 * Conventions: plain pointers and sizes only; the caller owns every buffer; functions returning int
 * return 0 on success and non-zero on error, with a thread-local message in qcb_last_error(); nothing
 * throws across the boundary; a plan is immutable after creation and supports one call in flight.
 * Entry points ending in _device take device pointers and a cudaStream_t (as void*) and do not
 * synchronise; the others take host pointers and return when the results are in the caller's buffers.
 * There is no CPU fallback: every compute entry point fails if no CUDA device is usable.
 */
#ifndef QCAT_B200_H
#define QCAT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QCB_MODE_EPI2ME 0
#define QCB_MODE_DUAL 1
#define QCB_MODE_SIMPLE 2   /* scanner_simple.py: template group 0 = the bare barcodes, one placeholder layout, no adapter stage */

/* Flattened scanner description (built by qcat_b200/tables.py).  All pointers are host pointers and are
 * copied by qcb_plan_create(). */
typedef struct {
    /* qcatConfig, config.py:12-26 */
    int32_t max_align_length;              /* W (150) */
    int32_t barcode_extension;             /* extracted_barcode_extension (11) */
    int32_t adapter_open, adapter_extend;  /* gap_open / gap_extend (2 / 2) */
    int32_t barcode_open, barcode_extend;  /* 1 / 1, scanner_base.py:115-116 */
    int32_t amat_size;                     /* adapter substitution matrix, amat_size^2 row-major ints */
    const int32_t *amat;
    const uint8_t *amap;                   /* 256-entry byte -> matrix index */
    int32_t bmat_size;                     /* barcode substitution matrix */
    const int32_t *bmat;
    const uint8_t *bmap;
    const uint8_t *comp;                   /* 256-entry complement table, utils.py:26-27 */
    /* scanner */
    int32_t mode;                          /* QCB_MODE_* */
    double min_quality;
    /* layouts in self.layouts order */
    int32_t n_layouts;
    const int32_t *adapter_off;            /* [n_layouts + 1] */
    const uint8_t *adapter_seq;            /* N-masked adapter templates */
    const double *denom;                   /* [n_layouts] score normaliser, scanner_base.py:308-310 */
    const int32_t *bc_end;                 /* [n_layouts * 2] */
    const int32_t *bc_len;                 /* [n_layouts * 2] */
    const int32_t *group;                  /* [n_layouts * 2] template group of barcode set k, -1 = None */
    const int32_t *trim_offset;            /* [n_layouts] */
    const int32_t *is_double;              /* [n_layouts] */
    /* template groups (up_context + barcode + down_context, in barcode-set order) */
    int32_t n_groups;
    const int32_t *group_off;              /* [n_groups + 1] */
    const int32_t *tmpl_off;               /* [n_templates + 1] */
    const uint8_t *tmpl_seq;
    const int32_t *tmpl_ident;             /* [n_templates] equality class of Barcode.id */
} qcb_tables;

/* One record per read == build_return_dict (scanner_base.py:362-390).  32 bytes. */
typedef struct {
    int32_t layout;         /* index of result['adapter'] in the tables' layouts, -1 = None (always in simple mode) */
    int32_t barcode;        /* index inside the layout's template group, -1 = None; dual: idx1 * n2 + idx2 */
    double  barcode_score;
    int32_t adapter_end;
    int32_t trim5p;
    int32_t trim3p;
    int32_t exit_status;    /* 0 ok, 1 none, 1002 conflicting ends */
} qcb_result;

typedef struct qcb_plan qcb_plan;

/* Which kernels a plan selected, for reporting. */
typedef struct {
    int32_t device;
    int32_t sm_count;
    int32_t fast_adapter;   /* 1 = packed u16x2 linear-gap adapter kernel, 0 = generic int32 affine kernel */
    int32_t fast_barcode;   /* 1 = packed u16x2 shared-context barcode kernel, 0 = generic */
    int32_t max_group_size;
    int32_t n_templates;
    int64_t workspace_bytes;
    int64_t kernel_launches; /* kernels launched by this plan so far */
} qcb_plan_info_t;

int         qcb_device_count(void);
const char *qcb_last_error(void);
const char *qcb_version(void);

qcb_plan *qcb_plan_create(const qcb_tables *tables, int device);
void      qcb_plan_destroy(qcb_plan *plan);
int       qcb_plan_info(qcb_plan *plan, qcb_plan_info_t *out);
/* 0 = automatic (fast kernels when the scoring scheme allows), 1 = force the generic kernels, 2 = generic kernels
 * without the row-chunked adapter stage of long qcb_scan windows (one thread per window x template; cross-check). */
int       qcb_plan_set_force_generic(qcb_plan *plan, int force);

/* Per-stage device timing (CUDA events around each pipeline stage on the launching stream).  Off by default.
 * Stages: 0 orient (window extraction + revcomp), 1 adapter DP, 2 template selection / region geometry,
 * 3 barcode DP (core columns), 4 two-end decision (or kit vote), 5 shared-context columns of the barcode templates
 * (packed path only).  qcb_plan_stage_times synchronises the recorded events and adds the elapsed milliseconds of
 * every launch since the last reset into ms[QCB_N_STAGES] / launches[QCB_N_STAGES]. */
#define QCB_N_STAGES 6
int       qcb_plan_set_profiling(qcb_plan *plan, int enable);
int       qcb_plan_stage_times(qcb_plan *plan, double *ms, int64_t *launches, int reset);

/* Batched semi-global alignment, every query against every reference: out[q * n_refs + r].
 * The semantics of parasail.sg_striped_32 (scanner_base.py:111-117, 214-218) as a stand-alone call -- a parity-test
 * primitive: host buffers, device memory allocated and released per call, default stream.  The production path never
 * aligns one pair at a time; it runs the batched pipeline of qcb_detect*(). */
int qcb_sg_batch(int device,
                 const uint8_t *queries, const int32_t *query_off, int32_t n_queries,
                 const uint8_t *refs, const int32_t *ref_off, int32_t n_refs,
                 int32_t open, int32_t extend,
                 const int32_t *matrix, int32_t msize, const uint8_t *mapper,
                 int32_t *score, int32_t *end_query, int32_t *end_ref);

/* The same with the alignment statistics of parasail.sg_stats_striped_32 (scanner_base.py:20-26; simple scanner
 * :106-123, align_adapter_identity :168-172): matches, similar (positive-scoring columns) and length.  Which of several
 * equal-scoring alignments is counted follows parasail's recurrence as restated in oracle/qcat_oracle.c (unpinned). */
int qcb_sg_stats_batch(int device,
                       const uint8_t *queries, const int32_t *query_off, int32_t n_queries,
                       const uint8_t *refs, const int32_t *ref_off, int32_t n_refs,
                       int32_t open, int32_t extend,
                       const int32_t *matrix, int32_t msize, const uint8_t *mapper,
                       int32_t *score, int32_t *end_query, int32_t *end_ref,
                       int32_t *matches, int32_t *similar, int32_t *length);

/* detect_barcode for n_reads reads given as their two windows (see qcat_b200/tables.py:pack_windows):
 * win5[i] = read[:W], tail3[i] = read[-W:] (not reverse-complemented), slots of `stride` bytes
 * (stride >= W, multiple of 16), wlen[i] = min(len, W), read_len[i] = len(read).
 * subset = the layout indices to consider ("kits", scanner_base.py:526-529); NULL/0 = all layouts. */
int qcb_detect(qcb_plan *plan, const uint8_t *win5, const uint8_t *tail3, int32_t stride,
               const int32_t *wlen, const int64_t *read_len, int64_t n_reads,
               const int32_t *subset, int32_t n_subset, qcb_result *out);

int qcb_detect_device(qcb_plan *plan, const uint8_t *d_win5, const uint8_t *d_tail3, int32_t stride,
                      const int32_t *d_wlen, const int64_t *d_read_len, int64_t n_reads,
                      const int32_t *subset /* host */, int32_t n_subset, qcb_result *d_out, void *stream);

/* ---- 4-bit windows ------------------------------------------------------------------------------------------------------
 * Everything the path does with a base goes through the two matrix mappers and the complement table (config.py:236-253,
 * utils.py:26-27), so bytes that agree on (adapter code, barcode code, codes of the complement) are interchangeable.
 * qcb_plan_base_classes writes that equivalence as a byte -> class table (classes 0 .. n-1; returns n, or 0 when the
 * plan's tables need more than 16 classes) and the packed entry points take windows with two classes per byte (even
 * position in the low nibble; slots of stride4 >= W / 2 bytes, a multiple of 8): half the host -> device traffic of the
 * ASCII form, same records.  qcb_pack_windows4 cuts indexed records straight into that form, qcb_pack_ascii4 converts
 * ASCII windows (host side, no GPU involved). */
int qcb_plan_base_classes(qcb_plan *plan, uint8_t *cls /* [256] */);

int qcb_detect4(qcb_plan *plan, const uint8_t *win5p, const uint8_t *tail3p, int32_t stride4,
                const int32_t *wlen, const int64_t *read_len, int64_t n_reads,
                const int32_t *subset, int32_t n_subset, qcb_result *out);

int qcb_detect4_device(qcb_plan *plan, const uint8_t *d_win5p, const uint8_t *d_tail3p, int32_t stride4,
                       const int32_t *d_wlen, const int64_t *d_read_len, int64_t n_reads,
                       const int32_t *subset /* host */, int32_t n_subset, qcb_result *d_out, void *stream);

int qcb_detect_auto4(qcb_plan *plan, const uint8_t *win5p, const uint8_t *tail3p, int32_t stride4,
                     const int32_t *wlen, const int64_t *read_len, int64_t n_reads,
                     const int32_t *kit_of_layout, int32_t batch_size, qcb_result *out, int32_t *batch_kit);

int qcb_pack_ascii4(const uint8_t *windows, int32_t stride, const int32_t *wlen, int64_t n, const uint8_t *cls,
                    uint8_t *packed, int32_t stride4, int32_t threads);

/* BarcodeScanner.scan (scanner_epi2me.py:33-144 / scanner_dual.py:35-146) for n_windows already-oriented windows
 * of any length (wlen[i] <= stride): one record per window with trim5p = trim3p = 0 and exit_status 0, or the
 * empty record (layout -1, exit_status 1) where the reference returns empty_return_dict().  Host buffers. */
int qcb_scan(qcb_plan *plan, const uint8_t *windows, int32_t stride, const int32_t *wlen, int64_t n_windows,
             const int32_t *subset, int32_t n_subset, qcb_result *out);

/* detect_kit's per-read vote over ALL layouts: layout index of the higher-scoring end (scanner_base.py:632-642). */
int qcb_kit_vote(qcb_plan *plan, const uint8_t *win5, const uint8_t *tail3, int32_t stride,
                 const int32_t *wlen, int64_t n_reads, int32_t *vote_layout);

int qcb_kit_vote_device(qcb_plan *plan, const uint8_t *d_win5, const uint8_t *d_tail3, int32_t stride,
                        const int32_t *d_wlen, int64_t n_reads, int32_t *d_vote_layout, void *stream);

/* detect_barcode_batch with more than one kit among the layouts (scanner_base.py:714-733; `-k auto`, the CLI default) in
 * ONE pass: the adapter stage runs once over all layouts, its scores give every read's vote (scan_ends, :632-642), every
 * batch of `batch_size` consecutive reads (cli.py:500: 4000) elects its kit on the device (most votes; ties: the kit
 * seen first in the batch, :657-660), and detection continues restricted to that kit's layouts on the adapter scores
 * already computed.  kit_of_layout: host array [n_layouts], kit index (0 .. 255) of every layout.  batch_kit (may be
 * NULL): [ceil(n_reads / batch_size)] elected kit per batch -- host memory for qcb_detect_auto, device memory for the
 * _device variant.  Results equal qcb_kit_vote + per-batch qcb_detect with the kit's layouts as subset. */
int qcb_detect_auto(qcb_plan *plan, const uint8_t *win5, const uint8_t *tail3, int32_t stride,
                    const int32_t *wlen, const int64_t *read_len, int64_t n_reads,
                    const int32_t *kit_of_layout, int32_t batch_size, qcb_result *out, int32_t *batch_kit);

int qcb_detect_auto_device(qcb_plan *plan, const uint8_t *d_win5, const uint8_t *d_tail3, int32_t stride,
                           const int32_t *d_wlen, const int64_t *d_read_len, int64_t n_reads,
                           const int32_t *kit_of_layout /* host */, int32_t batch_size, qcb_result *d_out,
                           int32_t *d_batch_kit, void *stream);

/* The same calls spread over several devices inside one process: plans[d] was created on device d with the same tables;
 * reads are dealt to the plans in blocks (>= 64 Ki reads, whole batches in the auto-kit call), round-robin, every plan
 * runs its blocks through its own copy / compute pipeline on its own host thread, and every record is written at its
 * read's own position in `out` (results in input order, scanner_base.py:714-733).  Reads are independent, so there is no
 * device-to-device traffic on this path. */
int qcb_detect_multi(qcb_plan *const *plans, int32_t n_plans, const uint8_t *win5, const uint8_t *tail3, int32_t stride,
                     const int32_t *wlen, const int64_t *read_len, int64_t n_reads,
                     const int32_t *subset, int32_t n_subset, qcb_result *out);

int qcb_detect_auto_multi(qcb_plan *const *plans, int32_t n_plans, const uint8_t *win5, const uint8_t *tail3, int32_t stride,
                          const int32_t *wlen, const int64_t *read_len, int64_t n_reads,
                          const int32_t *kit_of_layout, int32_t batch_size, qcb_result *out, int32_t *batch_kit);

/* The path's only exchange (the counts behind the CLI histogram, cli.py:386-405) for plans of one process: d_counts[d] =
 * device d's int64[n_bins] vector (qcb_histogram_device), d_gathered[d] = device d's int64[n_plans * n_bins] buffer that
 * receives every device's vector, written with peer copies (over NVLink where the devices are peers).  Synchronises.
 * Between processes (one rank per GPU) the same all-gather is one NCCL call: qcat_b200.dist.allgather_counts. */
int qcb_hist_allgather(qcb_plan *const *plans, int32_t n_plans, int64_t *const *d_counts, int32_t n_bins,
                       int64_t *const *d_gathered);

/* counts[bin] += 1 per record, bin = 0 for "none", 1 + layout_bin_base[layout] + barcode otherwise
 * (layout_bin_base: host array [n_layouts]).  d_counts must hold n_bins int64 and is NOT cleared. */
int qcb_histogram_device(qcb_plan *plan, const qcb_result *d_results, int64_t n_reads,
                         const int32_t *layout_bin_base, int64_t *d_counts, int32_t n_bins, void *stream);

/* ---- host-side ingest / egress next to the path (no GPU involved) -------------------------------------------------
 * What qcat/cli.py does in Python around detect_barcode_batch: record iteration (cli.py:235-306, Bio's
 * FastqGeneralIterator / SimpleFastaParser), window extraction (scanner_base.py:223-244) and the per-barcode,
 * optionally trimmed record output (cli.py:309-358, :521-552). */
typedef struct {
    int64_t title_off, title_len;   /* header line without '@' / '>' and without trailing whitespace */
    int64_t seq_off, seq_span;      /* byte range of the sequence lines (may contain line breaks) */
    int64_t seq_len;                /* bases */
    int64_t qual_off, qual_span;    /* FASTQ quality lines; qual_off = -1 for FASTA */
} qcb_fastx_record;

const char *qcb_io_last_error(void);

/* Index the complete records of an in-memory FASTQ / FASTA chunk.  *consumed = bytes covered by the records returned;
 * call again from there with more data.  final_chunk != 0: the buffer ends the file (a last record without a trailing
 * line break is complete).  Errors mirror the reference's ValueErrors (bad '@', missing '+', length mismatch). */
int qcb_fastx_index(const char *buf, int64_t len, int32_t final_chunk, qcb_fastx_record *recs, int64_t max_records,
                    int64_t *n_records, int64_t *consumed, int32_t *is_fastq);

/* win5[i] = read[:W], tail3[i] = read[-W:], wlen, read_len for indexed records (the buffers qcb_detect takes). */
int qcb_pack_windows(const char *buf, const qcb_fastx_record *recs, int64_t n, int32_t W, int32_t stride,
                     uint8_t *win5, uint8_t *tail3, int32_t *wlen, int64_t *read_len, int32_t threads);

/* qcb_pack_windows in the 4-bit form (see qcb_plan_base_classes): win5p / tail3p slots of stride4 bytes. */
int qcb_pack_windows4(const char *buf, const qcb_fastx_record *recs, int64_t n, int32_t W, int32_t stride4, const uint8_t *cls,
                      uint8_t *win5p, uint8_t *tail3p, int32_t *wlen, int64_t *read_len, int32_t threads);

/* Format records into per-bin byte strings ("@name comment\nSEQ\n+\nQUAL\n" / ">name comment\nSEQ\n"), trimmed to
 * [trim5p:trim3p] when trim != 0 and dropped when shorter than min_read_length (cli.py:521-552 with -b).  Call once
 * with out == NULL to obtain bin_bytes / bin_offset, then with a buffer of sum(bin_bytes).  kept[i] = record written. */
int qcb_format_records(const char *buf, const qcb_fastx_record *recs, const qcb_result *results, const int32_t *bin, int64_t n,
                       int32_t n_bins, int32_t fastq, int32_t trim, int64_t min_read_length,
                       int64_t *bin_bytes, uint8_t *out, int64_t out_capacity, int64_t *bin_offset, uint8_t *kept, int32_t threads);

/* Same as qcb_fastx_index with the scan spread over `threads` threads: the buffer is cut at record starts found by a
 * local test ('@' line, '+' two lines later, equal sequence / quality line lengths) that each neighbouring scan then
 * confirms by arriving exactly there; anything else (wrapped FASTQ, malformed input) falls back to the serial scan, so
 * results and errors are those of qcb_fastx_index.  Buffers under 1 MiB are scanned serially unless threads is negative
 * (|threads| pieces whatever the size -- used by the tests to fuzz the cutting logic on small inputs). */
int qcb_fastx_index_mt(const char *buf, int64_t len, int32_t final_chunk, qcb_fastx_record *recs, int64_t max_records,
                       int64_t *n_records, int64_t *consumed, int32_t *is_fastq, int32_t threads);

/* The CLI's default single output stream (cli.py:337-352): every kept record in input order with " barcode=<label>"
 * appended to its comment.  label[i] indexes a caller-provided table of strings (labels + label_off[n_labels + 1]):
 * str(barcode.id) or "none".  *out_bytes = size needed; nothing is written when out == NULL. */
int qcb_format_stream(const char *buf, const qcb_fastx_record *recs, const qcb_result *results, const int32_t *label, int64_t n,
                      const char *labels, const int64_t *label_off, int32_t n_labels, int32_t fastq, int32_t trim,
                      int64_t min_read_length, uint8_t *out, int64_t out_capacity, int64_t *out_bytes, uint8_t *kept, int32_t threads);

/* The --tsv table (cli.py:408-442) without its header line: name, length after trimming, barcode id, repr(score), kit,
 * adapter_end, comment ("None" when the header has none); unclassified records (label[i] < 0) get none / -1 / none / -1. */
int qcb_format_tsv(const char *buf, const qcb_fastx_record *recs, const qcb_result *results, const int32_t *label,
                   const int32_t *kit_label, int64_t n, const char *labels, const int64_t *label_off, int32_t n_labels,
                   int32_t trim, int64_t min_read_length, uint8_t *out, int64_t out_capacity, int64_t *out_bytes,
                   uint8_t *kept, int32_t threads);

/* Append every non-empty bin of a qcb_format_records buffer to its file descriptor (fds[n_bins], -1 = bin has no
 * file), `threads` bins at a time -- the `-b` per-barcode files of cli.py:309-336. */
int qcb_write_bins(const int32_t *fds, const uint8_t *out, const int64_t *bin_offset, const int64_t *bin_bytes, int32_t n_bins,
                   int32_t threads);

/* Chunked reader replacing iter_fastx (cli.py:235-306) for files: qcb_reader_next() returns the next chunk of complete
 * records (bytes read with `threads` concurrent preads, indexed with qcb_fastx_index_mt); every chunk but the last holds
 * a multiple of `multiple_of` records so that CLI batches of 4000 stay aligned.  *chunk = NULL at the end of the file.
 * A chunk stays valid until qcb_chunk_release(); chunks may be released from another thread than the reading one.
 * path "-" = standard input (the CLI's default input, cli.py:256-259); anything that is not a regular file is read
 * sequentially with read(). */
typedef struct qcb_reader qcb_reader;
typedef struct qcb_chunk qcb_chunk;
qcb_reader *qcb_reader_open(const char *path, int64_t chunk_bytes, int32_t threads);
int qcb_reader_next(qcb_reader *reader, int64_t multiple_of, qcb_chunk **chunk);
const char *qcb_chunk_data(const qcb_chunk *chunk, int64_t *len);
const qcb_fastx_record *qcb_chunk_records(const qcb_chunk *chunk, int64_t *n_records, int32_t *is_fastq);
void qcb_chunk_release(qcb_chunk *chunk);
void qcb_reader_close(qcb_reader *reader);

/* Issue-rate micro-benchmark of the DP inner instruction pair on this device (compute-roofline
 * denominator): packed 16-bit DP cell updates per second the SMs can issue. */
int qcb_microbench_cell_rate(int device, double *cells_per_second, double *sm_mhz_effective);

#ifdef __cplusplus
}
#endif
#endif
