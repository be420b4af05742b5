// Host-side ingest / egress next to the hot path: what qcat/cli.py does in Python around
// detect_barcode_batch -- FASTQ / FASTA record iteration (cli.py:235-306 over Bio's FastqGeneralIterator /
// SimpleFastaParser), window extraction (scanner_base.py:223-244) and the trimmed per-barcode record output
// (cli.py:309-358, :521-552) -- on in-memory buffers, multi-threaded where records are independent.
// No CUDA in this file; it is part of libqcat_b200.so so the binding stays one library.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/qcat_b200.h"

namespace {

thread_local std::string g_io_error;

int io_fail(const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_io_error = buf;
    return 1;
}

inline const char *line_end(const char *p, const char *end)
{
    const char *q = (const char *)memchr(p, '\n', (size_t)(end - p));
    return q ? q : end;
}

// length of [p, q) without trailing whitespace (str.rstrip())
inline int64_t rstrip_len(const char *p, const char *q)
{
    while (q > p && (q[-1] == '\r' || q[-1] == ' ' || q[-1] == '\t' || q[-1] == '\n')) --q;
    return q - p;
}

template <typename F>
void parallel_for(int64_t n, int threads, F fn)
{
    threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, n / 256 + 1));
    if (threads == 1) { fn(0, n); return; }
    std::vector<std::thread> pool;
    const int64_t step = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        const int64_t lo = t * step, hi = std::min<int64_t>(n, lo + step);
        if (lo >= hi) break;
        pool.emplace_back([=]() { fn(lo, hi); });
    }
    for (auto &th : pool) th.join();
}

// copy up to `want` sequence characters starting at p (forward), skipping line breaks
inline int copy_forward(const char *p, const char *end, uint8_t *dst, int want)
{
    int k = 0;
    while (k < want && p < end) {
        const char c = *p++;
        if (c == '\n' || c == '\r') continue;
        dst[k++] = (uint8_t)c;
    }
    return k;
}

// the last `want` sequence characters of [begin, end), in read order
inline int copy_tail(const char *begin, const char *end, uint8_t *dst, int want)
{
    int k = 0;
    const char *p = end;
    while (k < want && p > begin) {
        const char c = *--p;
        if (c == '\n' || c == '\r') continue;
        dst[want - 1 - k++] = (uint8_t)c;
    }
    if (k < want) memmove(dst, dst + (want - k), (size_t)k);
    return k;
}

}  // namespace

extern "C" {

const char *qcb_io_last_error(void) { return g_io_error.c_str(); }

int qcb_fastx_index(const char *buf, int64_t len, int32_t final_chunk, qcb_fastx_record *recs, int64_t max_records,
                    int64_t *n_records, int64_t *consumed, int32_t *is_fastq)
{
    if (!buf || !recs || !n_records || !consumed || !is_fastq) return io_fail("NULL argument");
    const char *p = buf, *end = buf + len;
    *n_records = 0; *consumed = 0;
    while (p < end && (*p == '\n' || *p == '\r')) ++p;
    if (p >= end) { *consumed = len; return 0; }
    if (*p != '@' && *p != '>')
        return io_fail("Invalid input file. File must start with '@' or '>'. Current file starts with: %c", *p);
    const bool fastq = *p == '@';
    *is_fastq = fastq ? 1 : 0;
    int64_t n = 0;
    const char *done = p;                 // everything before `done` belongs to complete records
    while (p < end && n < max_records) {
        while (p < end && (*p == '\n' || *p == '\r')) ++p;
        if (p >= end) { done = end; break; }
        const char *rec = p;
        qcb_fastx_record r;
        memset(&r, 0, sizeof r);
        if (*p != (fastq ? '@' : '>')) {
            if (fastq) return io_fail("Records in Fastq files should start with '@' character");
            return io_fail("FASTA record does not start with '>'");
        }
        const char *e = line_end(p, end);
        if (e == end && !final_chunk) break;                       // title line incomplete
        r.title_off = (p + 1) - buf;
        r.title_len = rstrip_len(p + 1, e);
        p = e < end ? e + 1 : end;
        // sequence lines
        r.seq_off = p - buf;
        int64_t bases = 0;
        const char *seq_end = p;
        bool complete = false;
        while (p < end) {
            if (fastq ? *p == '+' : *p == '>') { complete = true; break; }
            e = line_end(p, end);
            if (e == end && !final_chunk) { p = end; break; }
            bases += rstrip_len(p, e);
            seq_end = e;
            p = e < end ? e + 1 : end;
        }
        if (!fastq) {
            if (!complete && !final_chunk) { p = rec; break; }     // the next '>' has not been seen yet
            r.seq_span = seq_end - (buf + r.seq_off);
            r.seq_len = bases;
            r.qual_off = -1;
            recs[n++] = r;
            done = p;
            continue;
        }
        if (!complete) {
            if (!final_chunk) { p = rec; break; }
            return io_fail("Missing '+' line in FASTQ record");
        }
        r.seq_span = seq_end - (buf + r.seq_off);
        r.seq_len = bases;
        e = line_end(p, end);                                      // the '+' line
        if (e == end && !final_chunk) { p = rec; break; }
        p = e < end ? e + 1 : end;
        // quality lines until as many characters as bases
        r.qual_off = p - buf;
        int64_t qlen = 0;
        const char *qual_end = p;
        bool truncated = false;
        while (qlen < bases || (bases == 0 && qual_end == buf + r.qual_off && p < end && *p != '@')) {
            if (p >= end) { truncated = true; break; }
            e = line_end(p, end);
            if (e == end && !final_chunk) { truncated = true; break; }
            qlen += rstrip_len(p, e);
            qual_end = e;
            p = e < end ? e + 1 : end;
            if (bases == 0) break;
        }
        if (truncated) {
            if (!final_chunk) { p = rec; break; }
            return io_fail("Lengths of sequence and quality values differs (%lld and %lld)", (long long)bases, (long long)qlen);
        }
        if (qlen != bases)
            return io_fail("Lengths of sequence and quality values differs (%lld and %lld)", (long long)bases, (long long)qlen);
        r.qual_span = qual_end - (buf + r.qual_off);
        recs[n++] = r;
        done = p;
    }
    *n_records = n;
    *consumed = done - buf;
    return 0;
}

int qcb_pack_windows(const char *buf, const qcb_fastx_record *recs, int64_t n, int32_t W, int32_t stride,
                     uint8_t *win5, uint8_t *tail3, int32_t *wlen, int64_t *read_len, int32_t threads)
{
    if (!buf || !recs || !win5 || !tail3 || !wlen || !read_len) return n == 0 ? 0 : io_fail("NULL argument");
    if (W <= 0 || stride < W) return io_fail("need 0 < W <= stride");
    parallel_for(n, threads, [=](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            const qcb_fastx_record &r = recs[i];
            const char *s = buf + r.seq_off, *e = s + r.seq_span;
            uint8_t *h = win5 + (size_t)i * stride, *t = tail3 + (size_t)i * stride;
            const int k = (int)std::min<int64_t>(r.seq_len, W);
            memset(h, 0, (size_t)stride);
            memset(t, 0, (size_t)stride);
            if (r.seq_span == r.seq_len) {                         // single line: plain copies
                memcpy(h, s, (size_t)k);
                memcpy(t, e - k, (size_t)k);
            } else {
                copy_forward(s, e, h, k);
                copy_tail(s, e, t, k);
            }
            wlen[i] = k;
            read_len[i] = r.seq_len;
        }
    });
    return 0;
}

// Formats the records of one batch the way cli.py:write_to_file does for "-b out/" (per-barcode files):
//   fastq: "@name comment\nSEQ\n+\nQUAL\n", fasta: ">name comment\nSEQ\n", with SEQ / QUAL cut to [trim5p:trim3p] when
// trim != 0, skipping reads shorter than min_read_length after trimming (cli.py:521-530).  Records are appended to
// one growing byte string per output bin (bin[i] chosen by the caller from the record's barcode); the caller owns
// the files.  Two passes: sizes, then parallel formatting into place.
int qcb_format_records(const char *buf, const qcb_fastx_record *recs, const qcb_result *results, const int32_t *bin, int64_t n,
                       int32_t n_bins, int32_t fastq, int32_t trim, int64_t min_read_length,
                       int64_t *bin_bytes /* [n_bins] out */, uint8_t *out, int64_t out_capacity, int64_t *bin_offset /* [n_bins] out */,
                       uint8_t *kept /* [n] out */, int32_t threads)
{
    if (n == 0) { for (int b = 0; b < n_bins; ++b) { bin_bytes[b] = 0; bin_offset[b] = 0; } return 0; }
    if (!buf || !recs || !results || !bin || !bin_bytes || !bin_offset || !kept) return io_fail("NULL argument");
    std::vector<int64_t> size(n), pos(n);
    for (int64_t i = 0; i < n; ++i) {
        const qcb_fastx_record &r = recs[i];
        if (bin[i] < 0 || bin[i] >= n_bins) return io_fail("bin[%lld] out of range", (long long)i);
        if (r.seq_span != r.seq_len || (fastq && r.qual_span != r.seq_len)) return io_fail("multi-line records are not supported by the native writer");
        int64_t a = 0, b = r.seq_len;
        if (trim) {                                                 // Python slice semantics of seq[trim5p:trim3p]
            a = std::min<int64_t>(std::max<int64_t>(results[i].trim5p, 0), r.seq_len);
            b = std::min<int64_t>(std::max<int64_t>(results[i].trim3p, 0), r.seq_len);
            if (b < a) b = a;
        }
        const int64_t m = b - a;
        kept[i] = m >= min_read_length ? 1 : 0;
        // title = name + " " + comment where name / comment come from splitting the header at the first blank
        size[i] = kept[i] ? (1 + r.title_len + (memchr(buf + r.title_off, ' ', (size_t)r.title_len) || memchr(buf + r.title_off, '\t', (size_t)r.title_len) ? 0 : 1) + 1 + m + 1 + (fastq ? 2 + m + 1 : 0)) : 0;
    }
    std::vector<int64_t> cursor(n_bins, 0);
    for (int b = 0; b < n_bins; ++b) bin_bytes[b] = 0;
    for (int64_t i = 0; i < n; ++i) bin_bytes[bin[i]] += size[i];
    int64_t total = 0;
    for (int b = 0; b < n_bins; ++b) { bin_offset[b] = total; cursor[b] = total; total += bin_bytes[b]; }
    if (!out) return 0;                                             // sizing call
    if (total > out_capacity) return io_fail("output buffer too small: need %lld bytes", (long long)total);
    for (int64_t i = 0; i < n; ++i) { pos[i] = cursor[bin[i]]; cursor[bin[i]] += size[i]; }
    parallel_for(n, threads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            if (!kept[i]) continue;
            const qcb_fastx_record &r = recs[i];
            int64_t a = 0, b = r.seq_len;
            if (trim) {
                a = std::min<int64_t>(std::max<int64_t>(results[i].trim5p, 0), r.seq_len);
                b = std::min<int64_t>(std::max<int64_t>(results[i].trim3p, 0), r.seq_len);
                if (b < a) b = a;
            }
            uint8_t *o = out + pos[i];
            *o++ = fastq ? '@' : '>';
            // header.replace("\t", " ") split at blanks, re-joined with single blanks == the header with tabs as blanks;
            // a header without any blank gets "name" + " " + "" (cli.py:319-336)
            bool blank = false;
            for (int64_t j = 0; j < r.title_len; ++j) {
                char c = buf[r.title_off + j];
                if (c == '\t') c = ' ';
                if (c == ' ') blank = true;
                *o++ = (uint8_t)c;
            }
            if (!blank) *o++ = ' ';
            *o++ = '\n';
            memcpy(o, buf + r.seq_off + a, (size_t)(b - a)); o += b - a;
            *o++ = '\n';
            if (fastq) {
                *o++ = '+'; *o++ = '\n';
                memcpy(o, buf + r.qual_off + a, (size_t)(b - a)); o += b - a;
                *o++ = '\n';
            }
        }
    });
    return 0;
}

}  // extern "C"
