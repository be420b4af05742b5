// Host-side ingest / egress next to the hot path: what qcat/cli.py does in Python around
// detect_barcode_batch -- FASTQ / FASTA record iteration (cli.py:235-306 over Bio's FastqGeneralIterator /
// SimpleFastaParser), window extraction (scanner_base.py:223-244) and the trimmed per-barcode record output
// (cli.py:309-358, :521-552) -- on in-memory buffers, multi-threaded where records are independent.
// No CUDA in this file; it is part of libqcat_b200.so so the binding stays one library.
#include <algorithm>
#include <charconv>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <cerrno>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

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

// FASTA sequence lines: Bio's SimpleFastaParser joins line.rstrip() pieces and then removes every ' ' and '\r' that is
// left inside (FastaIO.py: "".join(lines).replace(" ", "").replace("\r", "")), so such characters are not bases.
inline bool fasta_blank(char c) { return c == ' ' || c == '\r'; }

inline int64_t fasta_line_bases(const char *p, const char *q)
{
    int64_t k = rstrip_len(p, q);
    const int64_t full = k;
    for (int64_t i = 0; i < full; ++i) k -= fasta_blank(p[i]) ? 1 : 0;
    return k;
}

// Characters [a, b) of a record's sequence (or quality) whose lines lie in [s, e): every line contributes its characters
// without trailing whitespace, exactly what the indexer counted (Bio's parsers join line.rstrip() pieces); skip_blanks
// (FASTA sequences) also drops blanks inside a line.
inline uint8_t *copy_bases(uint8_t *o, const char *s, const char *e, int64_t a, int64_t b, bool skip_blanks = false)
{
    int64_t pos = 0;
    const char *p = s;
    while (p < e && pos < b) {
        const char *le = line_end(p, e);
        const int64_t k = rstrip_len(p, le);
        if (skip_blanks && (memchr(p, ' ', (size_t)k) || memchr(p, '\r', (size_t)k))) {
            for (int64_t i = 0; i < k && pos < b; ++i) {
                if (fasta_blank(p[i])) continue;
                if (pos >= a) *o++ = (uint8_t)p[i];
                ++pos;
            }
        } else {
            const int64_t lo = std::max<int64_t>(a - pos, 0), hi = std::min<int64_t>(b - pos, k);
            if (hi > lo) { memcpy(o, p + lo, (size_t)(hi - lo)); o += hi - lo; }
            pos += k;
        }
        p = le < e ? le + 1 : e;
    }
    return o;
}

// characters [a, b) of a sequence / quality span: one memcpy for single-line records, line by line for wrapped ones
inline uint8_t *put_span(uint8_t *o, const char *span, int64_t span_bytes, int64_t n_chars, int64_t a, int64_t b,
                         bool skip_blanks = false)
{
    if (span_bytes == n_chars) { memcpy(o, span + a, (size_t)(b - a)); return o + (b - a); }
    return copy_bases(o, span, span + span_bytes, a, b, skip_blanks);
}

// Serial record scan of [start, end); offsets are relative to `base`.  `final_chunk`: the range ends the input (a last
// record without a line break is complete).  Stops after max_records.  *done = end of the last complete record.
int index_range(const char *base, const char *start, const char *end, bool fastq, bool final_chunk,
                qcb_fastx_record *recs, int64_t max_records, int64_t *n_records, const char **done_out)
{
    const char *buf = base;
    const char *p = start;
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
            bases += fastq ? rstrip_len(p, e) : fasta_line_bases(p, e);
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
    *done_out = done;
    return 0;
}

// First line start at or after `from` that begins a record, or NULL.  FASTA: any line starting with '>' (sequence lines
// never do).  FASTQ: a line starting with '@' whose line + 2 starts with '+' and whose lines + 1 and + 3 have equal
// lengths -- in the 4-line layout a quality line starting with '@' is followed two lines later by bases, never by '+'.
// Wrapped FASTQ can defeat this test; index_parallel() therefore verifies every sync point against the scan that
// reaches it from the left and falls back to the serial scan otherwise.
const char *find_sync(const char *buf, const char *from, const char *end, bool fastq)
{
    const char *p = from;
    if (p > buf) {                                                   // advance to a line start
        const char *q = (const char *)memchr(p - 1, '\n', (size_t)(end - (p - 1)));
        if (!q) return nullptr;
        p = q + 1;
    }
    while (p < end) {
        const char *e0 = line_end(p, end);
        if (*p == (fastq ? '@' : '>')) {
            if (!fastq) return p;
            if (e0 < end) {
                const char *l1 = e0 + 1, *e1 = line_end(l1, end);
                if (e1 < end) {
                    const char *l2 = e1 + 1, *e2 = line_end(l2, end);
                    if (*l2 == '+' && e2 < end) {
                        const char *l3 = e2 + 1, *e3 = line_end(l3, end);
                        if (rstrip_len(l1, e1) == rstrip_len(l3, e3) && (e3 + 1 >= end || e3[1] == '@')) return p;
                    }
                }
            }
        }
        if (e0 >= end) break;
        p = e0 + 1;
    }
    return nullptr;
}

// Record array without value-initialisation: capacity is a generous guess (one record per 64 bytes) and only the pages
// that really receive records are ever touched (a std::vector would zero-fill ~0.9 bytes per input byte).
struct RecordArray {
    std::unique_ptr<qcb_fastx_record[]> ptr;
    size_t count = 0;
    void allocate(size_t capacity) { ptr.reset(new qcb_fastx_record[capacity]); count = 0; }
    qcb_fastx_record *data() { return ptr.get(); }
    const qcb_fastx_record *data() const { return ptr.get(); }
    size_t size() const { return count; }
    void clear() { ptr.reset(); count = 0; }
    const qcb_fastx_record &operator[](size_t i) const { return ptr[i]; }
};

struct Segment {
    const char *start = nullptr, *stop = nullptr;
    RecordArray recs;
    const char *done = nullptr;
    int rc = 0;
    std::string err;
};

// Parallel scan: cut [p, end) at verified record starts, scan the pieces concurrently, concatenate.  Returns -1 when
// the pieces do not line up (the caller then scans serially), otherwise index_range's return code.
int index_parallel(const char *buf, const char *p, const char *end, bool fastq, bool final_chunk, int threads,
                   qcb_fastx_record *recs, int64_t max_records, int64_t *n_records, const char **done_out)
{
    const int64_t len = end - p;
    std::vector<Segment> seg;
    {
        seg.emplace_back(); seg.back().start = p;
        for (int t = 1; t < threads; ++t) {
            const char *from = p + len * t / threads;
            if (from <= seg.back().start) continue;
            const char *s = find_sync(buf, from, end, fastq);
            if (!s || s <= seg.back().start) continue;
            seg.emplace_back(); seg.back().start = s;
        }
        for (size_t i = 0; i < seg.size(); ++i) seg[i].stop = i + 1 < seg.size() ? seg[i + 1].start : end;
    }
    if (seg.size() < 2) return -1;
    std::vector<std::thread> pool;
    for (size_t i = 0; i < seg.size(); ++i) {
        pool.emplace_back([&, i]() {
            Segment &s = seg[i];
            const bool last = i + 1 == seg.size();
            const int64_t span = s.stop - s.start;
            int64_t cap = std::max<int64_t>(1024, span / 64);
            for (;;) {                                               // grow-and-retry keeps the scan allocation-light
                s.recs.allocate((size_t)cap);
                int64_t n = 0;
                s.rc = index_range(buf, s.start, s.stop, fastq, last ? final_chunk : true, s.recs.data(), cap, &n, &s.done);
                if (s.rc != 0) { s.err = g_io_error; s.recs.clear(); return; }
                if (n < cap || s.done >= s.stop) { s.recs.count = (size_t)n; return; }
                cap *= 4;
            }
        });
    }
    for (auto &th : pool) th.join();
    int64_t total = 0;
    for (size_t i = 0; i < seg.size(); ++i) {
        const Segment &s = seg[i];
        const bool last = i + 1 == seg.size();
        if (s.rc != 0) {
            // an inner piece that fails may just have been cut at a false sync point: let the serial scan decide
            if (!last || i > 0) return -1;
            g_io_error = s.err;
            return s.rc;
        }
        if (!last && s.done != s.stop) return -1;                    // did not land on the next sync point
        total += (int64_t)s.recs.size();
    }
    int64_t n = 0;
    const char *done = p;
    for (const Segment &s : seg) {
        const int64_t take = std::min<int64_t>((int64_t)s.recs.size(), max_records - n);
        if (take > 0) memcpy(recs + n, s.recs.data(), (size_t)take * sizeof(qcb_fastx_record));
        n += take;
        if (take < (int64_t)s.recs.size()) {                         // max_records reached inside this piece
            done = buf + s.recs[(size_t)take].title_off - 1;
            *n_records = n; *done_out = done;
            return 0;
        }
        done = s.done;
    }
    (void)total;
    *n_records = n;
    *done_out = done;
    return 0;
}

// Scan [p, end) whose first byte starts a record (or blank lines): parallel when worthwhile, serial otherwise.
int index_from(const char *buf, const char *p, const char *end, bool fastq, bool final_chunk, int threads,
               qcb_fastx_record *recs, int64_t max_records, int64_t *n_records, const char **done)
{
    int rc = -1;
    *done = p;
    // threads < 0: cut into |threads| pieces whatever the size (tests); otherwise only buffers worth the thread start-up
    if (threads < -1 || (threads > 1 && end - p >= (int64_t)1 << 20))
        rc = index_parallel(buf, p, end, fastq, final_chunk, threads < 0 ? -threads : threads, recs, max_records, n_records, done);
    if (rc < 0) rc = index_range(buf, p, end, fastq, final_chunk, recs, max_records, n_records, done);
    return rc;
}

int index_buffer(const char *buf, int64_t len, int32_t final_chunk, qcb_fastx_record *recs, int64_t max_records,
                 int64_t *n_records, int64_t *consumed, int32_t *is_fastq, int threads)
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
    const char *done = p;
    const int rc = index_from(buf, p, end, fastq, final_chunk != 0, threads, recs, max_records, n_records, &done);
    if (rc != 0) return rc;
    *consumed = done - buf;
    return 0;
}

}  // namespace

extern "C" {

const char *qcb_io_last_error(void) { return g_io_error.c_str(); }

int qcb_fastx_index(const char *buf, int64_t len, int32_t final_chunk, qcb_fastx_record *recs, int64_t max_records,
                    int64_t *n_records, int64_t *consumed, int32_t *is_fastq)
{
    return index_buffer(buf, len, final_chunk, recs, max_records, n_records, consumed, is_fastq, 1);
}

int qcb_fastx_index_mt(const char *buf, int64_t len, int32_t final_chunk, qcb_fastx_record *recs, int64_t max_records,
                       int64_t *n_records, int64_t *consumed, int32_t *is_fastq, int32_t threads)
{
    return index_buffer(buf, len, final_chunk, recs, max_records, n_records, consumed, is_fastq, threads);
}

// Two bases per byte (even position in the low nibble) through a byte -> class table (qcb_plan_base_classes).
static inline void pack_classes(const uint8_t *src, int k, const uint8_t *cls, uint8_t *dst, int stride4)
{
    memset(dst, 0, (size_t)stride4);
    int i = 0;
    for (; i + 1 < k; i += 2) dst[i >> 1] = (uint8_t)(cls[src[i]] | (cls[src[i + 1]] << 4));
    if (i < k) dst[i >> 1] = cls[src[i]];
}

int qcb_pack_ascii4(const uint8_t *windows, int32_t stride, const int32_t *wlen, int64_t n, const uint8_t *cls,
                    uint8_t *packed, int32_t stride4, int32_t threads)
{
    if (!windows || !wlen || !cls || !packed) return n == 0 ? 0 : io_fail("NULL argument");
    if (stride <= 0 || stride4 * 2 < stride) return io_fail("need stride4 >= stride / 2");
    parallel_for(n, threads, [=](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i)
            pack_classes(windows + (size_t)i * stride, std::min<int>(std::max<int>(wlen[i], 0), stride), cls,
                         packed + (size_t)i * stride4, stride4);
    });
    return 0;
}

int qcb_pack_windows4(const char *buf, const qcb_fastx_record *recs, int64_t n, int32_t W, int32_t stride4, const uint8_t *cls,
                      uint8_t *win5p, uint8_t *tail3p, int32_t *wlen, int64_t *read_len, int32_t threads)
{
    if (!buf || !recs || !cls || !win5p || !tail3p || !wlen || !read_len) return n == 0 ? 0 : io_fail("NULL argument");
    if (W <= 0 || W > 4096 || stride4 * 2 < W) return io_fail("need 0 < W <= 2 * stride4 (W <= 4096)");
    parallel_for(n, threads, [=](int64_t lo, int64_t hi) {
        std::vector<uint8_t> tmp((size_t)2 * W);
        for (int64_t i = lo; i < hi; ++i) {
            const qcb_fastx_record &r = recs[i];
            const char *s = buf + r.seq_off, *e = s + r.seq_span;
            const int k = (int)std::min<int64_t>(r.seq_len, W);
            const uint8_t *h = (const uint8_t *)s, *t = (const uint8_t *)(e - k);
            if (r.seq_span != r.seq_len) {                         // wrapped record: gather the bases line by line first
                copy_bases(tmp.data(), s, e, 0, k, r.qual_off < 0);
                copy_bases(tmp.data() + W, s, e, r.seq_len - k, r.seq_len, r.qual_off < 0);
                h = tmp.data(); t = tmp.data() + W;
            }
            pack_classes(h, k, cls, win5p + (size_t)i * stride4, stride4);
            pack_classes(t, k, cls, tail3p + (size_t)i * stride4, stride4);
            wlen[i] = k;
            read_len[i] = r.seq_len;
        }
    });
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
            } else {                                               // wrapped record: line by line
                copy_bases(h, s, e, 0, k, r.qual_off < 0);
                copy_bases(t, s, e, r.seq_len - k, r.seq_len, r.qual_off < 0);
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
            o = put_span(o, buf + r.seq_off, r.seq_span, r.seq_len, a, b, r.qual_off < 0);
            *o++ = '\n';
            if (fastq) {
                *o++ = '+'; *o++ = '\n';
                o = put_span(o, buf + r.qual_off, r.qual_span, r.seq_len, a, b);
                *o++ = '\n';
            }
        }
    });
    return 0;
}

// ---- one output stream and the TSV table -------------------------------------------------------------------------------

namespace {

struct Cut { int64_t a, b; };

// Python slice semantics of seq[trim5p:trim3p] for 0 <= trims (cli.py:521-526)
inline Cut trimmed_range(const qcb_fastx_record &r, const qcb_result &res, int trim)
{
    Cut c{0, r.seq_len};
    if (trim) {
        c.a = std::min<int64_t>(std::max<int64_t>(res.trim5p, 0), r.seq_len);
        c.b = std::min<int64_t>(std::max<int64_t>(res.trim3p, 0), r.seq_len);
        if (c.b < c.a) c.b = c.a;
    }
    return c;
}

// repr(float) of Python 3 for the magnitudes a score can take (1e-4 <= |x| < 1e16, or 0): shortest round-trip digits
inline char *put_float_repr(char *o, double v)
{
    auto r = std::to_chars(o, o + 40, v, std::chars_format::fixed);
    bool dot = false;
    for (char *q = o; q < r.ptr; ++q) if (*q == '.') dot = true;
    char *e = r.ptr;
    if (!dot) { *e++ = '.'; *e++ = '0'; }
    return e;
}

inline char *put_int(char *o, long long v)
{
    return std::to_chars(o, o + 24, v).ptr;
}

}  // namespace

// The default output of the CLI (no -b): every kept record, input order, on one stream with "barcode=<id>" appended to
// the comment (cli.py:337-352): "@name comment barcode=ID\nSEQ\n+\nQUAL\n" / ">name comment barcode=ID\nSEQ\n".
// label[i] indexes the caller's label table (labels / label_off: concatenated strings); the label of a record is
// str(barcode.id) or "none".  out_bytes = bytes needed; nothing is written when out == NULL or the capacity is short.
int qcb_format_stream(const char *buf, const qcb_fastx_record *recs, const qcb_result *results, const int32_t *label, int64_t n,
                      const char *labels, const int64_t *label_off, int32_t n_labels, int32_t fastq, int32_t trim,
                      int64_t min_read_length, uint8_t *out, int64_t out_capacity, int64_t *out_bytes, uint8_t *kept, int32_t threads)
{
    if (!out_bytes) return io_fail("NULL argument");
    *out_bytes = 0;
    if (n == 0) return 0;
    if (!buf || !recs || !results || !label || !labels || !label_off || !kept) return io_fail("NULL argument");
    std::vector<int64_t> pos((size_t)n + 1);
    for (int64_t i = 0; i < n; ++i)
        if (label[i] < 0 || label[i] >= n_labels) return io_fail("label[%lld] out of range", (long long)i);
    parallel_for(n, threads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            const qcb_fastx_record &r = recs[i];
            const Cut c = trimmed_range(r, results[i], trim);
            const int64_t m = c.b - c.a;
            kept[i] = m >= min_read_length ? 1 : 0;
            const bool blank = memchr(buf + r.title_off, ' ', (size_t)r.title_len) || memchr(buf + r.title_off, '\t', (size_t)r.title_len);
            const int64_t lab = label_off[label[i] + 1] - label_off[label[i]];
            pos[(size_t)i + 1] = kept[i] ? 1 + r.title_len + (blank ? 0 : 1) + 9 + lab + 1 + m + 1 + (fastq ? 2 + m + 1 : 0) : 0;
        }
    });
    pos[0] = 0;
    for (int64_t i = 0; i < n; ++i) pos[(size_t)i + 1] += pos[(size_t)i];
    *out_bytes = pos[(size_t)n];
    if (!out) return 0;
    if (pos[(size_t)n] > out_capacity) return io_fail("output buffer too small: need %lld bytes", (long long)pos[(size_t)n]);
    parallel_for(n, threads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            if (!kept[i]) continue;
            const qcb_fastx_record &r = recs[i];
            const Cut c = trimmed_range(r, results[i], trim);
            uint8_t *o = out + pos[(size_t)i];
            *o++ = fastq ? '@' : '>';
            bool blank = false;
            for (int64_t j = 0; j < r.title_len; ++j) {
                char ch = buf[r.title_off + j];
                if (ch == '\t') ch = ' ';
                if (ch == ' ') blank = true;
                *o++ = (uint8_t)ch;
            }
            if (!blank) *o++ = ' ';
            memcpy(o, " barcode=", 9); o += 9;
            const int64_t l0 = label_off[label[i]], l1 = label_off[label[i] + 1];
            memcpy(o, labels + l0, (size_t)(l1 - l0)); o += l1 - l0;
            *o++ = '\n';
            o = put_span(o, buf + r.seq_off, r.seq_span, r.seq_len, c.a, c.b, r.qual_off < 0);
            *o++ = '\n';
            if (fastq) {
                *o++ = '+'; *o++ = '\n';
                o = put_span(o, buf + r.qual_off, r.qual_span, r.seq_len, c.a, c.b);
                *o++ = '\n';
            }
        }
    });
    return 0;
}

// The --tsv table (cli.py:408-442), one line per kept record, input order, without the header line:
//   name \t len(sequence after trimming) \t barcode.id \t repr(score) \t kit \t adapter_end \t comment
// or  name \t len \t none \t -1 \t none \t -1 \t comment  for unclassified reads; comment = "None" when the header has
// no blank.  label[i] / kit_label[i] index the caller's label table (ignored for unclassified records, label[i] < 0).
int qcb_format_tsv(const char *buf, const qcb_fastx_record *recs, const qcb_result *results, const int32_t *label,
                   const int32_t *kit_label, int64_t n, const char *labels, const int64_t *label_off, int32_t n_labels,
                   int32_t trim, int64_t min_read_length, uint8_t *out, int64_t out_capacity, int64_t *out_bytes,
                   uint8_t *kept, int32_t threads)
{
    if (!out_bytes) return io_fail("NULL argument");
    *out_bytes = 0;
    if (n == 0) return 0;
    if (!buf || !recs || !results || !label || !kit_label || !labels || !label_off || !kept) return io_fail("NULL argument");
    for (int64_t i = 0; i < n; ++i)
        if (label[i] >= n_labels || kit_label[i] >= n_labels || (label[i] >= 0 && kit_label[i] < 0))
            return io_fail("label[%lld] out of range", (long long)i);
    threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, n / 256 + 1));
    std::vector<std::string> part((size_t)threads);
    const int64_t step = (n + threads - 1) / threads;
    auto work = [&](int t) {
        std::string &dst = part[(size_t)t];
        const int64_t lo = t * step, hi = std::min<int64_t>(n, lo + step);
        dst.reserve((size_t)std::max<int64_t>(0, hi - lo) * 96);
        char num[64];
        for (int64_t i = lo; i < hi; ++i) {
            const qcb_fastx_record &r = recs[i];
            const Cut c = trimmed_range(r, results[i], trim);
            const int64_t m = c.b - c.a;
            kept[i] = m >= min_read_length ? 1 : 0;
            if (!kept[i]) continue;
            const char *title = buf + r.title_off;
            int64_t name_len = r.title_len;
            for (int64_t j = 0; j < r.title_len; ++j) if (title[j] == ' ' || title[j] == '\t') { name_len = j; break; }
            dst.append(title, (size_t)name_len);
            dst.push_back('\t');
            dst.append(num, (size_t)(put_int(num, m) - num));
            dst.push_back('\t');
            if (label[i] >= 0) {
                dst.append(labels + label_off[label[i]], (size_t)(label_off[label[i] + 1] - label_off[label[i]]));
                dst.push_back('\t');
                dst.append(num, (size_t)(put_float_repr(num, results[i].barcode_score) - num));
                dst.push_back('\t');
                dst.append(labels + label_off[kit_label[i]], (size_t)(label_off[kit_label[i] + 1] - label_off[kit_label[i]]));
                dst.push_back('\t');
                dst.append(num, (size_t)(put_int(num, results[i].adapter_end) - num));
            } else {
                dst.append("none\t-1\tnone\t-1");
            }
            dst.push_back('\t');
            if (name_len == r.title_len) {
                dst.append("None");
            } else {
                const size_t at = dst.size();
                dst.append(title + name_len + 1, (size_t)(r.title_len - name_len - 1));
                for (size_t j = at; j < dst.size(); ++j) if (dst[j] == '\t') dst[j] = ' ';
            }
            dst.push_back('\n');
        }
    };
    if (threads == 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
        for (auto &th : pool) th.join();
    }
    int64_t total = 0;
    for (const std::string &p : part) total += (int64_t)p.size();
    *out_bytes = total;
    if (!out) return 0;
    if (total > out_capacity) return io_fail("output buffer too small: need %lld bytes", (long long)total);
    int64_t at = 0;
    for (const std::string &p : part) { memcpy(out + at, p.data(), p.size()); at += (int64_t)p.size(); }
    return 0;
}

// Append bin b's bytes (out + bin_offset[b], bin_bytes[b]) to the file descriptor fds[b] for every bin with bytes, using
// `threads` writers that pick bins from a shared counter (one writer per file, so the order inside a file is kept).
int qcb_write_bins(const int32_t *fds, const uint8_t *out, const int64_t *bin_offset, const int64_t *bin_bytes, int32_t n_bins,
                   int32_t threads)
{
    if (n_bins <= 0) return 0;
    if (!fds || !out || !bin_offset || !bin_bytes) return io_fail("NULL argument");
    std::vector<int32_t> todo;
    for (int32_t b = 0; b < n_bins; ++b) {
        if (bin_bytes[b] <= 0) continue;
        if (fds[b] < 0) return io_fail("no file descriptor for bin %d", (int)b);
        todo.push_back(b);
    }
    std::sort(todo.begin(), todo.end(), [&](int32_t a, int32_t b) { return bin_bytes[a] > bin_bytes[b]; });
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    auto work = [&]() {
        for (;;) {
            const size_t k = next.fetch_add(1);
            if (k >= todo.size()) return;
            const int32_t b = todo[k];
            const uint8_t *p = out + bin_offset[b];
            int64_t left = bin_bytes[b];
            while (left > 0) {
                const ssize_t w = write(fds[b], p, (size_t)std::min<int64_t>(left, (int64_t)1 << 30));
                if (w < 0) { if (errno == EINTR) continue; failed.store(errno ? errno : 1); return; }
                p += w; left -= w;
            }
        }
    };
    const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, (int)threads), todo.size()));
    if (T == 1) work();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < T; ++t) pool.emplace_back(work);
        for (auto &th : pool) th.join();
    }
    if (failed.load()) return io_fail("write failed: %s", strerror(failed.load()));
    return 0;
}

// ---- chunked file reader ------------------------------------------------------------------------------------------------

struct qcb_chunk {
    qcb_reader *owner;
    char *data;
    int64_t capacity, len;
    std::vector<qcb_fastx_record> recs;
    int64_t n;
    int32_t fastq;
};

struct qcb_reader {
    int fd = -1;
    int64_t file_size = 0, file_pos = 0;
    int64_t chunk_bytes = 0;
    int threads = 1;
    bool eof = false;
    bool stream = false;                     // not a regular file: read() sequentially until it returns 0
    qcb_chunk *pending = nullptr;            // next chunk's buffer; its first carry_len bytes follow the last record handed out
    int64_t carry_len = 0;
    std::vector<qcb_fastx_record> carry_recs;// records already indexed inside the carried bytes (offsets relative to their start)
    int64_t carry_scanned = 0;               // carried bytes those records cover
    int fastq = -1;                          // -1 until the first byte of the file has been seen
    std::mutex pool_lock;
    std::vector<qcb_chunk *> pool;           // released chunks, reused so their pages stay mapped
};

qcb_reader *qcb_reader_open(const char *path, int64_t chunk_bytes, int32_t threads)
{
    if (!path) { io_fail("NULL argument"); return nullptr; }
    const bool from_stdin = strcmp(path, "-") == 0;                 // the CLI reads stdin when no file is given (cli.py:256-259)
    const int fd = from_stdin ? dup(0) : open(path, O_RDONLY);
    if (fd < 0) { io_fail("cannot open %s", path); return nullptr; }
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); io_fail("cannot stat %s", path); return nullptr; }
    qcb_reader *rd = new qcb_reader();
    rd->fd = fd;
    rd->stream = !S_ISREG(st.st_mode);                               // pipes, FIFOs, terminals: sequential read(), size unknown
    rd->file_size = rd->stream ? INT64_MAX : (int64_t)st.st_size;
    rd->chunk_bytes = std::max<int64_t>(chunk_bytes, 4096);
    rd->threads = std::max(1, (int)threads);
    return rd;
}

void qcb_reader_close(qcb_reader *rd)
{
    if (!rd) return;
    if (rd->fd >= 0) close(rd->fd);
    if (rd->pending) { free(rd->pending->data); delete rd->pending; }
    for (qcb_chunk *c : rd->pool) { free(c->data); delete c; }
    delete rd;
}

void qcb_chunk_release(qcb_chunk *c)
{
    if (!c) return;
    qcb_reader *rd = c->owner;
    std::lock_guard<std::mutex> g(rd->pool_lock);
    rd->pool.push_back(c);
}

const char *qcb_chunk_data(const qcb_chunk *c, int64_t *len)
{
    if (len) *len = c ? c->len : 0;
    return c ? c->data : nullptr;
}

const qcb_fastx_record *qcb_chunk_records(const qcb_chunk *c, int64_t *n, int32_t *is_fastq)
{
    if (n) *n = c ? c->n : 0;
    if (is_fastq) *is_fastq = c ? c->fastq : 1;
    return c ? c->recs.data() : nullptr;
}

namespace {

// a pooled chunk whose buffer holds at least `capacity` bytes
qcb_chunk *acquire_chunk(qcb_reader *rd, int64_t capacity)
{
    qcb_chunk *c = nullptr;
    {
        std::lock_guard<std::mutex> g(rd->pool_lock);
        if (!rd->pool.empty()) { c = rd->pool.back(); rd->pool.pop_back(); }
    }
    if (!c) { c = new qcb_chunk(); c->owner = rd; c->data = nullptr; c->capacity = 0; }
    if (c->capacity < capacity) {
        // sized for a carried batch plus a full read, so pooled buffers (and their mapped pages) fit every later chunk
        free(c->data);
        c->capacity = std::max<int64_t>(capacity + capacity / 8, 2 * rd->chunk_bytes) + 4096;
        c->data = (char *)malloc((size_t)c->capacity);
        if (!c->data) { delete c; io_fail("out of memory"); return nullptr; }
    }
    return c;
}

}  // namespace

// Next chunk of the file: the bytes carried over from the previous chunk followed by up to chunk_bytes new bytes (read
// with `threads` concurrent preads), indexed in parallel.  Every chunk but the last holds a multiple of `multiple_of`
// records (CLI batches of 4000 stay aligned across chunks); the records kept back travel on, already indexed, at the
// front of the next chunk's buffer.  *chunk = NULL at the end of the file.
int qcb_reader_next(qcb_reader *rd, int64_t multiple_of, qcb_chunk **chunk)
{
    if (!rd || !chunk) return io_fail("NULL argument");
    *chunk = nullptr;
    if (multiple_of < 1) multiple_of = 1;
    for (;;) {
        if (rd->eof && rd->carry_len == 0) return 0;
        const int64_t want = std::min<int64_t>(rd->chunk_bytes, rd->file_size - rd->file_pos);
        int64_t total = rd->carry_len + want;                          // (a stream may deliver less: corrected after the read)
        auto T0 = std::chrono::steady_clock::now();
        qcb_chunk *c = rd->pending;
        rd->pending = nullptr;
        if (!c || c->capacity < total + 1) {
            qcb_chunk *bigger = acquire_chunk(rd, total + 1);
            if (!bigger) { if (c) qcb_chunk_release(c); return 1; }
            if (c) { memcpy(bigger->data, c->data, (size_t)rd->carry_len); qcb_chunk_release(c); }
            c = bigger;
        }
        char *dst = c->data + rd->carry_len;
        std::vector<int64_t> got((size_t)rd->threads, 0);
        int64_t stream_got = -1;
        if (rd->stream) {
            stream_got = 0;
            while (stream_got < want) {
                const ssize_t k = read(rd->fd, dst + stream_got, (size_t)(want - stream_got));
                if (k < 0) { if (errno == EINTR) continue; qcb_chunk_release(c); rd->carry_len = 0; return io_fail("read failed: %s", strerror(errno)); }
                if (k == 0) { rd->eof = true; break; }
                stream_got += k;
            }
        } else {
            const int64_t base = rd->file_pos;
            const int fd = rd->fd;
            int64_t *gp = got.data();
            const int T = rd->threads;
            const int64_t step = (want + T - 1) / std::max(T, 1);
            auto reader = [=](int t) {
                int64_t lo = t * step, hi = std::min<int64_t>(want, lo + step);
                while (lo < hi) {
                    const ssize_t k = pread(fd, dst + lo, (size_t)(hi - lo), (off_t)(base + lo));
                    if (k <= 0) break;
                    lo += k;
                    gp[t] += k;
                }
            };
            if (T == 1 || want < (1 << 20)) { for (int t = 0; t < T; ++t) reader(t); }
            else {
                std::vector<std::thread> pool;
                for (int t = 0; t < T; ++t) pool.emplace_back(reader, t);
                for (auto &th : pool) th.join();
            }
        }
        auto T1 = std::chrono::steady_clock::now();
        int64_t nread = 0;
        for (int64_t g : got) nread += g;
        if (rd->stream) nread = stream_got;
        else if (nread != want) { qcb_chunk_release(c); rd->carry_len = 0; return io_fail("short read"); }
        rd->file_pos += nread;
        if (!rd->stream && rd->file_pos >= rd->file_size) rd->eof = true;
        const bool final_chunk = rd->eof;
        total = rd->carry_len + nread;
        if (rd->fastq < 0) {
            const char *p0 = c->data, *e0 = c->data + total;
            while (p0 < e0 && (*p0 == '\n' || *p0 == '\r')) ++p0;
            if (p0 < e0) {
                if (*p0 != '@' && *p0 != '>') {
                    qcb_chunk_release(c);
                    return io_fail("Invalid input file. File must start with '@' or '>'. Current file starts with: %c", *p0);
                }
                rd->fastq = *p0 == '@' ? 1 : 0;
            }
        }
        const int32_t fastq = rd->fastq < 0 ? 1 : rd->fastq;
        // Records of the carried bytes were indexed last time; scan only what follows them, with a record array that
        // grows until the scan is no longer limited by it.
        const int64_t n0 = (int64_t)rd->carry_recs.size();
        int64_t cap = std::max<int64_t>(4096, (total - rd->carry_scanned) / 512), n = 0, consumed = rd->carry_scanned;
        for (;;) {
            // the record array of a pooled chunk only ever grows (in big steps): resizing it to the exact need of every
            // chunk re-allocated and zero-filled ~30 MB per 256 MB chunk, more than the scan itself costs
            if ((int64_t)c->recs.size() < n0 + cap) c->recs.resize((size_t)std::max<int64_t>(n0 + cap, 2 * (int64_t)c->recs.size()));
            const char *done = c->data + rd->carry_scanned;
            int64_t got_n = 0;
            const int rc = index_from(c->data, done, c->data + total, fastq != 0, final_chunk, rd->threads,
                                      c->recs.data() + n0, cap, &got_n, &done);
            if (rc != 0) { qcb_chunk_release(c); rd->carry_len = 0; return rc; }
            n = n0 + got_n;
            consumed = done - c->data;
            if (got_n < cap) break;
            cap *= 4;
        }
        auto T2 = std::chrono::steady_clock::now();
        if (n0) memcpy(c->recs.data(), rd->carry_recs.data(), (size_t)n0 * sizeof(qcb_fastx_record));
        int64_t keep = n;
        if (!final_chunk && n % multiple_of) keep = n - n % multiple_of;
        // bytes handed out end where the first record kept back starts; those records travel on with their bytes
        const int64_t cut = keep < n ? c->recs[(size_t)keep].title_off - 1 : consumed;
        if (final_chunk) {
            for (const char *q = c->data + cut; q < c->data + total; ++q)
                if (!(*q == '\n' || *q == '\r' || *q == ' ' || *q == '\t')) { qcb_chunk_release(c); rd->carry_len = 0; return io_fail("trailing bytes that do not form a record"); }
            rd->carry_recs.clear();
            rd->carry_scanned = 0;
            rd->carry_len = 0;
        } else {
            rd->carry_recs.assign(c->recs.begin() + keep, c->recs.begin() + n);
            for (qcb_fastx_record &r : rd->carry_recs) {
                r.title_off -= cut; r.seq_off -= cut;
                if (r.qual_off >= 0) r.qual_off -= cut;
            }
            rd->carry_scanned = consumed - cut;
            rd->carry_len = total - cut;
            if (rd->carry_len > 64 * rd->chunk_bytes) { qcb_chunk_release(c); rd->carry_len = 0; return io_fail("record larger than 64 chunks"); }
            if (keep == 0) {                                          // nothing to hand out yet: keep reading into this buffer
                if (cut > 0) memmove(c->data, c->data + cut, (size_t)rd->carry_len);
                rd->pending = c;
                continue;
            }
            qcb_chunk *next = acquire_chunk(rd, rd->carry_len + std::min<int64_t>(rd->chunk_bytes, rd->file_size - rd->file_pos) + 1);
            if (!next) { qcb_chunk_release(c); rd->carry_len = 0; return 1; }
            memcpy(next->data, c->data + cut, (size_t)rd->carry_len);
            rd->pending = next;
        }
        auto T3 = std::chrono::steady_clock::now();
        if (getenv("QCB_IO_TRACE")) fprintf(stderr, "read %.1f ms  index %.1f ms  carry %.1f ms  (total %lld carry %lld)\n",
            std::chrono::duration<double, std::milli>(T1 - T0).count(), std::chrono::duration<double, std::milli>(T2 - T1).count(),
            std::chrono::duration<double, std::milli>(T3 - T2).count(), (long long)total, (long long)rd->carry_len);
        c->len = cut;
        c->n = keep;
        c->fastq = fastq;
        if (keep == 0) { qcb_chunk_release(c); return 0; }            // final chunk without records
        *chunk = c;
        return 0;
    }
}

}  // extern "C"
