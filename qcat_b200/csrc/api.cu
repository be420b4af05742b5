// C-ABI layer of libqcat_b200.so: plan (device tables + workspace), pipeline orchestration, host-buffer
// entry points.  See include/qcat_b200.h for the contract and the reference interfaces each call replaces.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "plan.h"
#include "kernels_generic.cuh"
#include "kernels_fast.cuh"
#include "microbench.cuh"

using namespace qcb;

namespace {

thread_local std::string g_error;

int fail(const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return 1;
}

#define QCB_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

inline unsigned grid_for(long long n, int block) { return (unsigned)((n + block - 1) / block); }

struct DeviceBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
    int reserve(size_t need)
    {
        if (need <= bytes) return 0;
        if (ptr) cudaFree(ptr);
        ptr = nullptr; bytes = 0;
        size_t want = need + need / 4;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e != cudaSuccess) return fail("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        bytes = want;
        return 0;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; bytes = 0; }
};

}  // namespace

struct qcb_plan {
    int device = 0;
    int sm_count = 0;
    bool force_generic = false;
    bool no_row_chunks = false;      // force level 2: long windows on one thread per (window, template), as a cross-check
    DevTables t{};
    void *slab = nullptr;            // device tables
    size_t slab_bytes = 0;
    std::vector<int32_t> h_group_off, h_group, h_tmpl_off, h_adapter_off;
    std::vector<int32_t> subset_cached;   // layout subset currently held in subset_dev
    std::vector<int32_t> bin_base_cached; // histogram bin bases currently held in misc
    int bmax0 = 0, bmax1 = 0;        // score slots of barcode set 0 / set 1: the largest set over all layouts, rounded up to 4
    int max_group = 0;               // largest barcode set
    int max_adapter = 0, max_template = 0;
    FastPlan fast;                   // packed-kernel tables (kernels_fast.cuh)
    // workspace, sized per chunk of reads
    DeviceBuffer wins, codes, ad_score, ad_end, sel, bc_score, subset_dev, misc, long_part, long_row;
    DeviceBuffer bc_endq;                           // simple mode: end_query per (window, barcode)
    DeviceBuffer unp;                               // 4-bit windows unpacked to representative ASCII (generic kernels only)
    int n_classes = 0;                              // base classes of the 4-bit window format (0: not available)
    uint8_t h_cls[256] = {0};                       // byte -> class
    void *cls_dev = nullptr;                        // device: fwd[16] | rev[16] | rep[16]
    DeviceBuffer auto_vote, auto_kit, auto_map;   // auto-kit flow: per-read vote, per-batch kit, layout -> kit table
    int long_ov = 0;                 // warm-up rows of k_adapter_long (0 = chunking not provably exact for these tables)
    cudaStream_t stream = nullptr;   // kernels of the host-buffer entry points
    cudaStream_t copy_in = nullptr, copy_out = nullptr;      // H2D / D2H of the host-buffer entry points
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_compute[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
    DeviceBuffer in_stage2[2], out_stage2[2];
    long long host_chunk_reads = 1 << 16;
    long long launches = 0;
    bool profiling = false;
    struct StageEvent { int stage; cudaEvent_t a, b; long long launches; };
    std::vector<StageEvent> stage_events;
    double stage_ms[QCB_N_STAGES] = {0, 0, 0, 0, 0, 0};
    long long stage_launches[QCB_N_STAGES] = {0, 0, 0, 0, 0, 0};
    long long chunk_reads = 1 << 18;
};

namespace {

constexpr int kLongChunkRows = 1024;   // rows per chunk of k_adapter_long

template <typename T>
size_t slab_put(std::vector<uint8_t> &slab, const T *src, size_t count)
{
    size_t off = (slab.size() + 15) / 16 * 16;
    slab.resize(off + sizeof(T) * (count ? count : 1));
    if (count) memcpy(slab.data() + off, src, sizeof(T) * count);
    return off;
}

int upload_tables(qcb_plan *p, const qcb_tables *h)
{
    const int nl = h->n_layouts, ng = h->n_groups;
    const int nt = ng > 0 ? h->group_off[ng] : 0;
    std::vector<uint8_t> slab;
    size_t o_amat = slab_put(slab, h->amat, (size_t)h->amat_size * h->amat_size);
    size_t o_amap = slab_put(slab, h->amap, 256);
    size_t o_bmat = slab_put(slab, h->bmat, (size_t)h->bmat_size * h->bmat_size);
    size_t o_bmap = slab_put(slab, h->bmap, 256);
    size_t o_comp = slab_put(slab, h->comp, 256);
    size_t o_aoff = slab_put(slab, h->adapter_off, nl + 1);
    size_t o_aseq = slab_put(slab, h->adapter_seq, nl ? h->adapter_off[nl] : 0);
    size_t o_den = slab_put(slab, h->denom, nl);
    size_t o_bce = slab_put(slab, h->bc_end, nl * 2);
    size_t o_bcl = slab_put(slab, h->bc_len, nl * 2);
    size_t o_grp = slab_put(slab, h->group, nl * 2);
    size_t o_trm = slab_put(slab, h->trim_offset, nl);
    size_t o_dbl = slab_put(slab, h->is_double, nl);
    size_t o_goff = slab_put(slab, h->group_off, ng + 1);
    size_t o_toff = slab_put(slab, h->tmpl_off, nt + 1);
    size_t o_tseq = slab_put(slab, h->tmpl_seq, nt ? h->tmpl_off[nt] : 0);
    size_t o_tid = slab_put(slab, h->tmpl_ident, nt);
    std::vector<int32_t> group_tlen(ng > 0 ? ng : 1, -1);
    for (int g = 0; g < ng; ++g) {
        const int b0 = h->group_off[g], b1 = h->group_off[g + 1];
        int tl = b1 > b0 ? h->tmpl_off[b0 + 1] - h->tmpl_off[b0] : -1;
        for (int b = b0; b < b1; ++b) if (h->tmpl_off[b + 1] - h->tmpl_off[b] != tl) tl = -1;
        group_tlen[g] = tl;
    }
    size_t o_gtl = slab_put(slab, group_tlen.data(), ng);

    QCB_CUDA(cudaMalloc(&p->slab, slab.size()));
    p->slab_bytes = slab.size();
    QCB_CUDA(cudaMemcpy(p->slab, slab.data(), slab.size(), cudaMemcpyHostToDevice));
    const uint8_t *d = (const uint8_t *)p->slab;
    DevTables &t = p->t;
    t.W = h->max_align_length; t.ext = h->barcode_extension;
    t.a_open = h->adapter_open; t.a_extend = h->adapter_extend;
    t.b_open = h->barcode_open; t.b_extend = h->barcode_extend;
    t.amat_size = h->amat_size; t.bmat_size = h->bmat_size;
    t.mode = h->mode; t.n_layouts = nl; t.n_groups = ng; t.n_templates = nt;
    t.min_quality = h->min_quality;
    t.amat = (const int32_t *)(d + o_amat); t.amap = d + o_amap;
    t.bmat = (const int32_t *)(d + o_bmat); t.bmap = d + o_bmap; t.comp = d + o_comp;
    t.adapter_off = (const int32_t *)(d + o_aoff); t.adapter_seq = d + o_aseq;
    t.denom = (const double *)(d + o_den);
    t.bc_end = (const int32_t *)(d + o_bce); t.bc_len = (const int32_t *)(d + o_bcl);
    t.group = (const int32_t *)(d + o_grp); t.trim_offset = (const int32_t *)(d + o_trm);
    t.is_double = (const int32_t *)(d + o_dbl);
    t.group_off = (const int32_t *)(d + o_goff); t.tmpl_off = (const int32_t *)(d + o_toff);
    t.tmpl_seq = d + o_tseq; t.tmpl_ident = (const int32_t *)(d + o_tid);
    t.group_tlen = (const int32_t *)(d + o_gtl);

    p->h_group_off.assign(h->group_off, h->group_off + ng + 1);
    p->h_group.assign(h->group, h->group + nl * 2);
    p->h_tmpl_off.assign(h->tmpl_off, h->tmpl_off + nt + 1);
    p->h_adapter_off.assign(h->adapter_off, h->adapter_off + nl + 1);
    for (int L = 0; L < nl; ++L) {
        p->max_adapter = std::max(p->max_adapter, h->adapter_off[L + 1] - h->adapter_off[L]);
        for (int k = 0; k < 2; ++k) {
            int g = h->group[L * 2 + k];
            if (g < 0) continue;
            int sz = h->group_off[g + 1] - h->group_off[g];
            if (k == 0) p->bmax0 = std::max(p->bmax0, sz); else p->bmax1 = std::max(p->bmax1, sz);
        }
    }
    p->max_group = std::max(p->bmax0, p->bmax1);
    p->bmax0 = (p->bmax0 + 3) / 4 * 4;       // score slots per window and set: multiples of 4, so that k_finalize reads
    p->bmax1 = (p->bmax1 + 3) / 4 * 4;       // a window's scores with aligned 16-byte loads
    for (int b = 0; b < nt; ++b) p->max_template = std::max(p->max_template, h->tmpl_off[b + 1] - h->tmpl_off[b]);
    {
        // Base classes of the 4-bit window format: bytes with equal (adapter code, barcode code) as they stand and
        // complemented are interchangeable on this path.  Class numbers follow the first byte value of every class.
        std::vector<uint32_t> keys;
        uint8_t tabs[48] = {0};
        p->n_classes = 0;
        bool ok = true;
        for (int b = 0; b < 256 && ok; ++b) {
            const uint8_t c = h->comp[b];
            const uint32_t key = (uint32_t)h->amap[b] | ((uint32_t)h->bmap[b] << 8) | ((uint32_t)h->amap[c] << 16) | ((uint32_t)h->bmap[c] << 24);
            size_t k = 0;
            while (k < keys.size() && keys[k] != key) ++k;
            if (k == keys.size()) {
                if (keys.size() == 16 || h->amap[b] > 15 || h->bmap[b] > 15 || h->amap[c] > 15 || h->bmap[c] > 15) { ok = false; break; }
                keys.push_back(key);
                tabs[k] = (uint8_t)(h->amap[b] | (h->bmap[b] << 4));
                tabs[16 + k] = (uint8_t)(h->amap[c] | (h->bmap[c] << 4));
                tabs[32 + k] = (uint8_t)b;
            }
            p->h_cls[b] = (uint8_t)k;
        }
        if (ok) {
            p->n_classes = (int)keys.size();
            QCB_CUDA(cudaMalloc(&p->cls_dev, sizeof tabs));
            QCB_CUDA(cudaMemcpy(p->cls_dev, tabs, sizeof tabs, cudaMemcpyHostToDevice));
        }
    }
    {
        // k_adapter_long's warm-up length: m (1 + (smax - smin) / gmin) + 1 rows for the longest adapter
        int smax = INT32_MIN, smin = INT32_MAX, max_alen = 0;
        for (int i = 0; i < h->amat_size * h->amat_size; ++i) { smax = std::max(smax, h->amat[i]); smin = std::min(smin, h->amat[i]); }
        for (int L = 0; L < nl; ++L) max_alen = std::max(max_alen, h->adapter_off[L + 1] - h->adapter_off[L]);
        const int gmin = std::min(h->adapter_open, h->adapter_extend);
        p->long_ov = 0;
        if (gmin > 0 && smax >= smin) {
            const long long ov = (long long)max_alen + ((long long)max_alen * (smax - smin) + gmin - 1) / gmin + 1;
            if (ov < (1 << 20)) p->long_ov = (int)ov;
        }
    }
    return 0;
}

int validate_tables(const qcb_tables *h)
{
    if (!h) return fail("tables is NULL");
    if (h->n_layouts <= 0) return fail("no layouts in tables");
    if (h->amat_size <= 0 || h->amat_size > kMaxMatrix || h->bmat_size <= 0 || h->bmat_size > kMaxMatrix)
        return fail("substitution matrix size must be in [1, %d]", kMaxMatrix);
    if (h->max_align_length <= 0) return fail("max_align_length must be positive");
    if (h->mode != QCB_MODE_EPI2ME && h->mode != QCB_MODE_DUAL && h->mode != QCB_MODE_SIMPLE) return fail("unknown mode %d", h->mode);
    if (h->mode == QCB_MODE_SIMPLE && (h->n_layouts != 1 || h->n_groups < 1 || h->group[0] != 0))
        return fail("simple mode needs one placeholder layout whose barcode set is template group 0");
    for (int L = 0; L < h->n_layouts; ++L) {
        int alen = h->adapter_off[L + 1] - h->adapter_off[L];
        if (alen <= 0 || alen > kMaxTemplate) return fail("adapter %d length %d outside [1, %d]", L, alen, kMaxTemplate);
        if (h->group[L * 2] < 0) return fail("layout %d has no barcode set", L);
        if (h->mode == QCB_MODE_DUAL && h->group[L * 2 + 1] < 0) return fail("dual mode needs a second barcode set (layout %d)", L);
    }
    int nt = h->n_groups > 0 ? h->group_off[h->n_groups] : 0;
    for (int b = 0; b < nt; ++b) {
        int tl = h->tmpl_off[b + 1] - h->tmpl_off[b];
        if (tl <= 0 || tl > kMaxTemplate) return fail("barcode template %d length %d outside [1, %d]", b, tl, kMaxTemplate);
    }
    for (int i = 0; i < 256; ++i)
        if (h->amap[i] >= h->amat_size || h->bmap[i] >= h->bmat_size) return fail("mapper entry %d out of range", i);
    return 0;
}

// Run the pipeline over one chunk of reads already resident on the device.
struct StageTimer {
    qcb_plan *p; int stage; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr; long long l0;
    StageTimer(qcb_plan *p_, int stage_, cudaStream_t st_) : p(p_), stage(stage_), st(st_), l0(p_->launches)
    {
        if (!p->profiling) return;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, st);
    }
    ~StageTimer()
    {
        if (!a) return;
        cudaEventRecord(b, st);
        p->stage_events.push_back({stage, a, b, p->launches - l0});
    }
};

// Auto-kit flow of one call (detect_barcode_batch with several kits, scanner_base.py:714-733): the adapter stage runs
// once over all layouts; its scores feed the per-read vote, the per-batch kit and then the kit-restricted selection.
struct AutoKit {
    const int32_t *d_kit_of_layout;   // [n_layouts] kit index of every layout
    int n_kits;
    int batch_size;                   // reads per CLI batch (4000)
    int32_t *d_batch_kit;             // [batches of the call] winning kit
    long long read_offset;            // index of the chunk's first read inside the call
    bool have_kits;                   // d_batch_kit is already filled (batches larger than a chunk: voted in a first pass)
};

// d_tail3 == nullptr selects window mode: d_win5 holds n already-oriented windows (BarcodeScanner.scan).
int run_chunk(qcb_plan *p, const uint8_t *d_win5, const uint8_t *d_tail3, int stride, const int32_t *d_wlen,
              const int64_t *d_read_len, long long n, const int32_t *d_subset, const int32_t *h_subset, int n_subset,
              qcb_result *d_out, int32_t *d_vote, cudaStream_t st, const AutoKit *autokit = nullptr, bool packed4 = false)
{
    const bool window_mode = d_tail3 == nullptr;
    if (packed4) {
        // `stride` counts the bytes of a 4-bit slot; the kernels see windows of 2 * stride bases.  The packed kernels
        // take the classes straight into their code bytes; a generic stage needs ASCII, so representative bytes are
        // written out first (one byte per class: same codes, same complement codes, hence the same records).
        const int cstride = 2 * stride;
        const bool direct = !p->force_generic && !window_mode && p->t.mode != QCB_MODE_SIMPLE && p->fast.adapter_ok &&
                            p->fast.barcode_ok && cstride <= kFastMaxStride && (cstride % 16) == 0 && n_subset <= 64;
        if (!direct) {
            if (window_mode) return fail("4-bit windows are not supported by qcb_scan");
            if (p->unp.reserve((size_t)2 * n * cstride)) return 1;
            uint8_t *u5 = (uint8_t *)p->unp.ptr, *u3 = u5 + (size_t)n * cstride;
            const uint8_t *rep = (const uint8_t *)p->cls_dev + 32;
            k_unpack4<<<grid_for(n * cstride, 256), 256, 0, st>>>(d_win5, stride, d_wlen, n, rep, u5);
            k_unpack4<<<grid_for(n * cstride, 256), 256, 0, st>>>(d_tail3, stride, d_wlen, n, rep, u3);
            p->launches += 2;
            return run_chunk(p, u5, u3, cstride, d_wlen, d_read_len, n, d_subset, h_subset, n_subset, d_out, d_vote, st, autokit, false);
        }
        stride = cstride;
    }
    const int wshift = window_mode ? 0 : 1;
    const long long nw = window_mode ? n : 2 * n;
    const DevTables &t = p->t;
    const int bslots = p->bmax0 + (t.mode == QCB_MODE_DUAL ? p->bmax1 : 0);
    if (t.mode == QCB_MODE_SIMPLE) {
        // scanner_simple.py:47-92: no adapter stage; every bare barcode against the whole window (generic int32 kernel,
        // which also returns the end_query the simple scanner reports as adapter_end)
        if (d_vote || autokit) return fail("the simple scanner has no kit vote");
        const uint8_t *wins = d_win5;
        if (!window_mode) {
            StageTimer timer(p, 0, st);
            if (p->wins.reserve((size_t)nw * stride)) return 1;
            k_orient<<<grid_for(nw * stride, 256), 256, 0, st>>>(d_win5, d_tail3, stride, d_wlen, n, t.comp, (uint8_t *)p->wins.ptr);
            p->launches++;
            wins = (const uint8_t *)p->wins.ptr;
        }
        if (p->sel.reserve((size_t)nw * sizeof(WindowSel)) || p->bc_score.reserve((size_t)nw * bslots * 4) ||
            p->bc_endq.reserve((size_t)nw * bslots * 4)) return 1;
        WindowSel *sel = (WindowSel *)p->sel.ptr;
        int32_t *bc_score = (int32_t *)p->bc_score.ptr, *bc_endq = (int32_t *)p->bc_endq.ptr;
        {
            StageTimer timer(p, 2, st);
            k_select_simple<<<grid_for(nw, 256), 256, 0, st>>>(d_wlen, wshift, nw, sel);
            p->launches++;
        }
        {
            StageTimer timer(p, 3, st);
            k_barcode_generic<<<grid_for(nw * bslots, 128), 128, 0, st>>>(t, wins, stride, nw, sel, p->bmax0, bslots, bc_score, bc_endq);
            p->launches++;
        }
        StageTimer timer(p, 4, st);
        if (window_mode) k_scan_out<<<grid_for(nw, 128), 128, 0, st>>>(t, d_wlen, nw, sel, bc_score, p->bmax0, bslots, d_out, bc_endq);
        else k_finalize<<<grid_for(n, 128), 128, 0, st>>>(t, d_wlen, d_read_len, n, sel, bc_score, p->bmax0, bslots, d_out, bc_endq);
        p->launches++;
        QCB_CUDA(cudaGetLastError());
        return 0;
    }
    // the ASCII copy of the oriented windows is only read by the generic kernels
    const bool packed_only = !p->force_generic && !window_mode && p->fast.adapter_ok && p->fast.barcode_ok &&
                             stride <= kFastMaxStride && (stride % 16) == 0 && n_subset <= 64;
    if (!window_mode && !packed_only && p->wins.reserve((size_t)nw * stride)) return 1;
    if (p->ad_score.reserve((size_t)nw * n_subset * 4)) return 1;
    if (p->ad_end.reserve((size_t)nw * n_subset * 4)) return 1;
    const uint8_t *wins = d_win5;
    int32_t *ad_score = (int32_t *)p->ad_score.ptr, *ad_end = (int32_t *)p->ad_end.ptr;
    const bool fast_ok = !p->force_generic && !window_mode && stride <= kFastMaxStride && (stride % 16) == 0 &&
                         (p->fast.adapter_ok || p->fast.barcode_ok);
    if (!window_mode) {
        StageTimer timer(p, 0, st);
        if (fast_ok) {
            if (p->codes.reserve((size_t)nw * stride)) return 1;
            if (packed4)
                k_map_codes4<<<grid_for(nw * (stride / 16), 256), 256, 0, st>>>(d_win5, d_tail3, stride / 2, d_wlen, n,
                                                                             (const uint8_t *)p->cls_dev, (uint8_t *)p->codes.ptr);
            else
                k_map_codes<<<grid_for(nw * (stride / 16), 256), 256, 0, st>>>(d_win5, d_tail3, stride, d_wlen, n, t.comp, t.amap, t.bmap,
                                                                            (uint8_t *)p->codes.ptr);
        }
        if (!packed_only) {                  // a generic stage will run: it reads the oriented ASCII windows
            k_orient<<<grid_for(nw * stride, 256), 256, 0, st>>>(d_win5, d_tail3, stride, d_wlen, n, t.comp, (uint8_t *)p->wins.ptr);
            if (fast_ok) p->launches++;
        }
        p->launches++;
        wins = (const uint8_t *)p->wins.ptr;
    }

    {
    StageTimer timer(p, 1, st);
    int adapter_rc = 2;
    if (fast_ok && p->fast.adapter_ok) {
        adapter_rc = fast_adapter_stage(p->fast, t, (const uint8_t *)p->codes.ptr, stride, d_wlen, wshift, nw, h_subset, n_subset,
                                        ad_score, ad_end, st, &p->launches);
        if (adapter_rc == 1) return fail("fast adapter stage launch failed");
    }
    if (adapter_rc == 2 && window_mode && !p->no_row_chunks && p->long_ov > 0 && stride >= 2 * kLongChunkRows) {
        // long windows (--detect-middle): rows in parallel chunks, see k_adapter_long
        const int nch = (stride + kLongChunkRows - 1) / kLongChunkRows;
        if (p->long_part.reserve((size_t)nw * n_subset * nch * sizeof(LongPart))) return 1;
        if (p->long_row.reserve((size_t)nw * n_subset * sizeof(int2))) return 1;
        k_adapter_long<<<grid_for(nw * n_subset * nch, 128), 128, 0, st>>>(t, wins, stride, d_wlen, nw, d_subset, n_subset, kLongChunkRows,
                                                                         p->long_ov, nch, (LongPart *)p->long_part.ptr, (int2 *)p->long_row.ptr);
        k_adapter_long_combine<<<grid_for(nw * n_subset, 128), 128, 0, st>>>(t, d_wlen, nw, d_subset, n_subset, nch,
                                                                            (const LongPart *)p->long_part.ptr, (const int2 *)p->long_row.ptr,
                                                                            ad_score, ad_end);
        p->launches += 2;
        adapter_rc = 0;
    }
    if (adapter_rc == 2) {
        if (packed_only) return fail("internal: packed adapter stage unavailable after the ASCII windows were skipped");
        k_adapter_generic<<<grid_for(nw * n_subset, 128), 128, 0, st>>>(t, wins, stride, d_wlen, wshift, nw, d_subset, n_subset,
                                                                       ad_score, ad_end);
        p->launches++;
    }
    }
    if (d_vote) {
        StageTimer timer(p, 4, st);
        k_kit_vote<<<grid_for(n, 256), 256, 0, st>>>(t, d_wlen, n, d_subset, n_subset, ad_score, ad_end, d_vote);
        p->launches++;
        QCB_CUDA(cudaGetLastError());
        return 0;
    }
    if (autokit && !autokit->have_kits) {                // chunk = whole batches: vote and pick the kits right here
        StageTimer timer(p, 4, st);
        if (p->auto_vote.reserve((size_t)n * 4)) return 1;
        int32_t *vote = (int32_t *)p->auto_vote.ptr;
        k_kit_vote<<<grid_for(n, 256), 256, 0, st>>>(t, d_wlen, n, d_subset, n_subset, ad_score, ad_end, vote);
        k_batch_kit<<<grid_for(n, autokit->batch_size), 256, 0, st>>>(vote, n, autokit->batch_size, autokit->d_kit_of_layout, autokit->n_kits,
                                                                      autokit->d_batch_kit + autokit->read_offset / autokit->batch_size);
        p->launches += 2;
    }
    if (p->sel.reserve((size_t)nw * sizeof(WindowSel))) return 1;
    if (p->bc_score.reserve((size_t)nw * bslots * 4)) return 1;
    WindowSel *sel = (WindowSel *)p->sel.ptr;
    int32_t *bc_score = (int32_t *)p->bc_score.ptr;
    {
    StageTimer timer(p, 2, st);
    k_select<<<grid_for(nw, 256), 256, 0, st>>>(t, d_wlen, wshift, nw, d_subset, n_subset, ad_score, ad_end, sel,
                                                autokit ? autokit->d_kit_of_layout : nullptr, autokit ? autokit->d_batch_kit : nullptr,
                                                autokit ? autokit->batch_size : 1, autokit ? autokit->read_offset : 0);
    p->launches++;
    }
    if (fast_ok && p->fast.barcode_ok) {
        StageTimer timer(p, 5, st);
        if (fast_context_stage(p->fast, t, (const uint8_t *)p->codes.ptr, stride, d_wlen, nw, sel, st, &p->launches))
            return fail("shared-context stage launch failed");
    }
    {
    StageTimer timer(p, 3, st);
    if (fast_ok && p->fast.barcode_ok) {
        int rc = fast_barcode_stage(p->fast, t, nw, p->bmax0, bslots, bc_score, st, &p->launches);
        if (rc) return fail("fast barcode stage launch failed");
    } else {
        k_barcode_generic<<<grid_for(nw * bslots, 128), 128, 0, st>>>(t, wins, stride, nw, sel, p->bmax0, bslots, bc_score, nullptr);
        p->launches++;
    }
    }
    StageTimer timer(p, 4, st);
    if (window_mode) k_scan_out<<<grid_for(nw, 128), 128, 0, st>>>(t, d_wlen, nw, sel, bc_score, p->bmax0, bslots, d_out, nullptr);
    else k_finalize<<<grid_for(n, 128), 128, 0, st>>>(t, d_wlen, d_read_len, n, sel, bc_score, p->bmax0, bslots, d_out, nullptr);
    p->launches++;
    QCB_CUDA(cudaGetLastError());
    return 0;
}

int prepare_subset(qcb_plan *p, const int32_t *subset, int n_subset, std::vector<int32_t> &h, cudaStream_t st)
{
    if (!subset || n_subset <= 0) {
        h.resize(p->t.n_layouts);
        for (int i = 0; i < p->t.n_layouts; ++i) h[i] = i;
    } else {
        h.assign(subset, subset + n_subset);
        for (int v : h)
            if (v < 0 || v >= p->t.n_layouts) return fail("layout subset entry %d out of range", v);
    }
    if (h == p->subset_cached && p->subset_dev.ptr) return 0;    // same subset as the previous call: already on the device
    if (p->subset_dev.reserve(h.size() * 4)) return 1;
    // a change of subset is rare (another kit): wait for whatever may still read the old one, on any stream
    QCB_CUDA(cudaDeviceSynchronize());
    QCB_CUDA(cudaMemcpy(p->subset_dev.ptr, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    p->subset_cached = h;
    return 0;
}

// Host description of an auto-kit call; d_batch_kit receives one kit index per batch of the whole call.
struct AutoCall {
    const int32_t *kit_of_layout;     // host, [n_layouts]
    int batch_size;
    int32_t *d_batch_kit;             // device, [ceil(n_reads / batch_size)] or NULL (plan-owned scratch is used)
};

int auto_prepare(qcb_plan *p, const AutoCall *ac, int64_t n_reads, int &n_kits, int32_t *&d_batch_kit, cudaStream_t st)
{
    if (!ac->kit_of_layout) return fail("kit_of_layout is NULL");
    if (ac->batch_size <= 0) return fail("batch_size must be positive");
    n_kits = 0;
    for (int L = 0; L < p->t.n_layouts; ++L) {
        if (ac->kit_of_layout[L] < 0 || ac->kit_of_layout[L] >= kMaxKits) return fail("kit_of_layout[%d] outside [0, %d)", L, kMaxKits);
        n_kits = std::max(n_kits, ac->kit_of_layout[L] + 1);
    }
    if (p->auto_map.reserve((size_t)p->t.n_layouts * 4)) return 1;
    QCB_CUDA(cudaMemcpyAsync(p->auto_map.ptr, ac->kit_of_layout, (size_t)p->t.n_layouts * 4, cudaMemcpyHostToDevice, st));
    d_batch_kit = ac->d_batch_kit;
    if (!d_batch_kit) {
        const size_t nb = (size_t)((n_reads + ac->batch_size - 1) / ac->batch_size);
        if (p->auto_kit.reserve(nb * 4)) return 1;
        d_batch_kit = (int32_t *)p->auto_kit.ptr;
    }
    return 0;
}

int detect_device_impl(qcb_plan *p, const uint8_t *d_win5, const uint8_t *d_tail3, int32_t stride, const int32_t *d_wlen,
                       const int64_t *d_read_len, int64_t n_reads, const int32_t *subset, int32_t n_subset,
                       qcb_result *d_out, int32_t *d_vote, cudaStream_t st, const AutoCall *ac = nullptr, bool packed4 = false)
{
    if (!p) return fail("plan is NULL");
    if (n_reads < 0) return fail("n_reads is negative");
    if (packed4 && (p->n_classes <= 0 || (stride % 8) != 0))
        return fail(p->n_classes <= 0 ? "this plan's tables need more than 16 base classes: no 4-bit window format"
                                      : "4-bit window slots must be a multiple of 8 bytes");
    if (d_tail3 && (packed4 ? 2 * stride : stride) < p->t.W)
        return fail("window slots of %d bases are smaller than max_align_length %d", packed4 ? 2 * stride : stride, p->t.W);
    QCB_CUDA(cudaSetDevice(p->device));
    if (n_reads == 0) return 0;
    std::vector<int32_t> h_subset;
    if (prepare_subset(p, subset, n_subset, h_subset, st)) return 1;
    const int32_t *d_subset = (const int32_t *)p->subset_dev.ptr;
    long long dev_chunk = p->chunk_reads;
    if ((long long)stride * dev_chunk > (1LL << 28)) dev_chunk = std::max<long long>(1, (1LL << 28) / stride);
    AutoKit ak{};
    if (ac) {
        if (auto_prepare(p, ac, n_reads, ak.n_kits, ak.d_batch_kit, st)) return 1;
        ak.d_kit_of_layout = (const int32_t *)p->auto_map.ptr;
        ak.batch_size = ac->batch_size;
        if (ac->batch_size <= dev_chunk) {
            dev_chunk = dev_chunk / ac->batch_size * ac->batch_size;       // every chunk holds whole batches: one pass
        } else {
            // a batch spans several chunks: vote over the whole call first (adapter stage only), then the regular pass
            if (p->auto_vote.reserve((size_t)n_reads * 4)) return 1;
            int32_t *vote = (int32_t *)p->auto_vote.ptr;
            int rc = 0;
            for (long long off = 0; off < n_reads && !rc; off += dev_chunk) {
                long long n = std::min<long long>(dev_chunk, n_reads - off);
                rc = run_chunk(p, d_win5 + off * stride, d_tail3 + off * stride, stride, d_wlen + off, nullptr, n, d_subset,
                               h_subset.data(), (int)h_subset.size(), nullptr, vote + off, st, nullptr, packed4);
            }
            if (rc) return 1;
            k_batch_kit<<<grid_for(n_reads, ac->batch_size), 256, 0, st>>>(vote, n_reads, ac->batch_size, ak.d_kit_of_layout, ak.n_kits,
                                                                          ak.d_batch_kit);
            p->launches++;
            ak.have_kits = true;
        }
    }
    for (long long off = 0; off < n_reads; off += dev_chunk) {
        long long n = std::min<long long>(dev_chunk, n_reads - off);
        ak.read_offset = off;
        if (run_chunk(p, d_win5 + off * stride, d_tail3 ? d_tail3 + off * stride : nullptr, stride, d_wlen + off,
                      d_read_len ? d_read_len + off : nullptr, n, d_subset, h_subset.data(), (int)h_subset.size(),
                      d_out ? d_out + off : nullptr, d_vote ? d_vote + off : nullptr, st, ac ? &ak : nullptr, packed4))
            return 1;
    }
    return 0;
}

// One device's share of a multi-device call: blocks of `block` reads starting at first, first + period, ...
struct Shard { long long first, block, period; };

int validate_wlen(const qcb_plan *p, const int32_t *wlen, int64_t n_reads, int32_t stride, bool window_mode, int64_t first = 0)
{
    const int32_t limit = window_mode ? stride : std::min<int>(stride, p->t.W);
    uint32_t bad = 0;
    for (int64_t i = 0; i < n_reads; ++i) bad |= (uint32_t)(wlen[i] < 0) | (uint32_t)(wlen[i] > limit);   // branch-free: vectorises
    if (!bad) return 0;
    for (int64_t i = 0; i < n_reads; ++i)
        if (wlen[i] < 0 || wlen[i] > limit) return fail("wlen[%lld] = %d outside [0, %d]", (long long)(first + i), wlen[i], limit);
    return 0;
}

// Host-buffer path: H2D -> device pipeline -> D2H, chunk by chunk, software pipelined over three streams with
// double-buffered staging so the copies of chunk k+1 / k-1 overlap the kernels of chunk k (pinned host memory makes
// the copies truly asynchronous; pageable memory still works, just without the overlap).
int detect_host_impl(qcb_plan *p, const uint8_t *win5, const uint8_t *tail3, int32_t stride, const int32_t *wlen,
                     const int64_t *read_len, int64_t n_reads, const int32_t *subset, int32_t n_subset,
                     qcb_result *out, int32_t *vote, const int32_t *kit_of_layout = nullptr, int32_t batch_size = 0,
                     int32_t *batch_kit_out = nullptr, const Shard *shard = nullptr, bool packed4 = false)
{
    if (!p) return fail("plan is NULL");
    if (n_reads < 0) return fail("n_reads is negative");
    if (n_reads == 0) return 0;
    const bool window_mode = tail3 == nullptr;
    if (!win5 || !wlen || (!vote && !out) || (!vote && !window_mode && !read_len)) return fail("NULL input/output buffer");
    // (window lengths are validated chunk by chunk, right before a chunk is copied: off the critical path of a large call)
    QCB_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = p->stream, s_in = p->copy_in, s_out = p->copy_out;
    // Pipeline granularity: regular chunks of a quarter of the call, at least 64 Ki reads (small launches lose time in
    // their last wave of tiles and in launch gaps), never above the device chunk; the first chunks are an eighth, a
    // quarter and a half of that (see the chunk list below), so the kernels start after a short first copy.
    static const long long chunk_div = getenv("QCB_HOST_CHUNK_DIV") ? std::max(1, atoi(getenv("QCB_HOST_CHUNK_DIV"))) : 4;
    long long chunk = std::max<long long>(p->host_chunk_reads, (n_reads / chunk_div + 31) / 32 * 32);
    chunk = std::min<long long>(chunk, p->chunk_reads);
    if ((long long)stride * chunk > (1LL << 27)) chunk = std::max<long long>(1, (1LL << 27) / stride);
    const bool auto_mode = kit_of_layout != nullptr;
    const long long n_batches = auto_mode ? (n_reads + batch_size - 1) / batch_size : 0;
    const long long span = shard ? shard->block : n_reads;       // reads this plan handles in one piece
    if (auto_mode) {
        // pipeline chunks are whole batches; a batch larger than a chunk goes to the device in one piece
        if (batch_size <= 0) return fail("batch_size must be positive");
        chunk = batch_size <= chunk ? chunk / batch_size * batch_size : span;
        if (p->auto_kit.reserve((size_t)n_batches * 4)) return 1;
    }
    // the chunks of this call: the whole input, or this device's blocks of a multi-device call
    // The first three chunks are an eighth, a quarter and a half of the regular size: computing starts sooner.
    std::vector<std::pair<long long, long long>> chunks;
    const long long unit = auto_mode ? batch_size : 32;
    for (long long base = shard ? shard->first : 0; base < n_reads; base += shard ? shard->period : n_reads) {
        const long long end = std::min<long long>(n_reads, base + span);
        for (long long off = base; off < end;) {
            long long size = chunk;
            if (chunks.size() < 3 && chunk >= (1 << 16))
                size = std::max<long long>(unit, (chunk >> (3 - chunks.size())) / unit * unit);
            size = std::min<long long>(size, end - off);
            chunks.emplace_back(off, size);
            off += size;
        }
    }
    const size_t out_item = vote ? 4 : sizeof(qcb_result);
    int rc = 0;
    // errors inside the loop must still reach the stream synchronisations below: async copies on the caller's
    // (pinned) buffers may be in flight
#define LOOP_CUDA(call)                                                                                   \
    {                                                                                                     \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) { rc = fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); break; } \
    }
    for (size_t k = 0; k < chunks.size() && !rc; ++k) {
        const int b = (int)(k & 1);
        const long long off = chunks[k].first, n = chunks[k].second;
        size_t b_win = (size_t)n * stride, b_len = (size_t)n * 4, b_rl = (size_t)n * 8;
        size_t o_tail = (b_win + 255) / 256 * 256, o_len = o_tail + (b_win + 255) / 256 * 256;
        size_t o_rl = o_len + (b_len + 255) / 256 * 256, total = o_rl + b_rl;
        // staging buffer b was last used by chunk k-2: its kernels and its D2H must be done before it is reused
        if (k >= 2) {
            LOOP_CUDA(cudaStreamWaitEvent(s_in, p->ev_compute[b], 0));
            LOOP_CUDA(cudaStreamWaitEvent(st, p->ev_d2h[b], 0));
        }
        if (p->in_stage2[b].bytes < total || p->out_stage2[b].bytes < (size_t)n * out_item) {
            LOOP_CUDA(cudaDeviceSynchronize());             // growing a staging buffer frees memory still in flight
            if (p->in_stage2[b].reserve(total) || p->out_stage2[b].reserve((size_t)n * out_item)) { rc = 1; break; }
        }
        uint8_t *d = (uint8_t *)p->in_stage2[b].ptr;
        void *d_res = p->out_stage2[b].ptr;
        if (validate_wlen(p, wlen + off, n, packed4 ? 2 * stride : stride, window_mode, off)) { rc = 1; break; }
        LOOP_CUDA(cudaMemcpyAsync(d, win5 + off * stride, b_win, cudaMemcpyHostToDevice, s_in));
        if (!window_mode) { LOOP_CUDA(cudaMemcpyAsync(d + o_tail, tail3 + off * stride, b_win, cudaMemcpyHostToDevice, s_in)); }
        LOOP_CUDA(cudaMemcpyAsync(d + o_len, wlen + off, b_len, cudaMemcpyHostToDevice, s_in));
        if (!vote && !window_mode) { LOOP_CUDA(cudaMemcpyAsync(d + o_rl, read_len + off, b_rl, cudaMemcpyHostToDevice, s_in)); }
        LOOP_CUDA(cudaEventRecord(p->ev_h2d[b], s_in));
        LOOP_CUDA(cudaStreamWaitEvent(st, p->ev_h2d[b], 0));
        AutoCall ac{kit_of_layout, batch_size, auto_mode ? (int32_t *)p->auto_kit.ptr + off / batch_size : nullptr};
        rc = detect_device_impl(p, d, window_mode ? nullptr : d + o_tail, stride, (const int32_t *)(d + o_len),
                                (const int64_t *)(d + o_rl), n, subset, n_subset, vote ? nullptr : (qcb_result *)d_res,
                                vote ? (int32_t *)d_res : nullptr, st, auto_mode ? &ac : nullptr, packed4);
        if (rc) break;
        LOOP_CUDA(cudaEventRecord(p->ev_compute[b], st));
        LOOP_CUDA(cudaStreamWaitEvent(s_out, p->ev_compute[b], 0));
        if (vote) { LOOP_CUDA(cudaMemcpyAsync(vote + off, d_res, (size_t)n * 4, cudaMemcpyDeviceToHost, s_out)); }
        else { LOOP_CUDA(cudaMemcpyAsync(out + off, d_res, (size_t)n * sizeof(qcb_result), cudaMemcpyDeviceToHost, s_out)); }
        if (auto_mode && batch_kit_out) {
            const long long b0 = off / batch_size, nb = (n + batch_size - 1) / batch_size;
            LOOP_CUDA(cudaMemcpyAsync(batch_kit_out + b0, (const int32_t *)p->auto_kit.ptr + b0, (size_t)nb * 4, cudaMemcpyDeviceToHost, s_out));
        }
        LOOP_CUDA(cudaEventRecord(p->ev_d2h[b], s_out));
    }
#undef LOOP_CUDA
    cudaError_t e1 = cudaStreamSynchronize(s_in), e2 = cudaStreamSynchronize(st), e3 = cudaStreamSynchronize(s_out);
    if (rc) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        return fail("host path failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
    return 0;
}

}  // namespace

extern "C" {

int qcb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *qcb_last_error(void) { return g_error.c_str(); }

const char *qcb_version(void) { return "qcat_b200 0.1.0 (sm_100a)"; }

qcb_plan *qcb_plan_create(const qcb_tables *tables, int device)
{
    if (validate_tables(tables)) return nullptr;
    int ndev = qcb_device_count();
    if (ndev <= 0) { fail("no CUDA device available (libqcat_b200 has no CPU fallback)"); return nullptr; }
    if (device < 0 || device >= ndev) { fail("device %d out of range (%d devices)", device, ndev); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { fail("cudaSetDevice(%d) failed", device); return nullptr; }
    qcb_plan *p = new qcb_plan();
    p->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { fail("cudaGetDeviceProperties failed"); qcb_plan_destroy(p); return nullptr; }
    p->sm_count = prop.multiProcessorCount;
    if (upload_tables(p, tables)) { qcb_plan_destroy(p); return nullptr; }
    if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&p->copy_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&p->copy_out, cudaStreamNonBlocking) != cudaSuccess) { fail("stream creation failed"); qcb_plan_destroy(p); return nullptr; }
    for (int i = 0; i < 2; ++i)
        if (cudaEventCreateWithFlags(&p->ev_h2d[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->ev_compute[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->ev_d2h[i], cudaEventDisableTiming) != cudaSuccess) { fail("event creation failed"); qcb_plan_destroy(p); return nullptr; }
    if (fast_plan_build(p->fast, tables, p->sm_count)) { fail("fast-plan construction failed: %s", p->fast.error.c_str()); qcb_plan_destroy(p); return nullptr; }
    return p;
}

void qcb_plan_destroy(qcb_plan *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    fast_plan_free(p->fast);
    p->wins.release(); p->codes.release(); p->ad_score.release(); p->ad_end.release(); p->sel.release(); p->bc_score.release();
    p->subset_dev.release(); p->misc.release(); p->long_part.release(); p->long_row.release();
    p->auto_vote.release(); p->auto_kit.release(); p->auto_map.release(); p->bc_endq.release(); p->unp.release();
    if (p->cls_dev) cudaFree(p->cls_dev);
    if (p->slab) cudaFree(p->slab);
    for (int i = 0; i < 2; ++i) {
        p->in_stage2[i].release(); p->out_stage2[i].release();
        if (p->ev_h2d[i]) cudaEventDestroy(p->ev_h2d[i]);
        if (p->ev_compute[i]) cudaEventDestroy(p->ev_compute[i]);
        if (p->ev_d2h[i]) cudaEventDestroy(p->ev_d2h[i]);
    }
    if (p->stream) cudaStreamDestroy(p->stream);
    if (p->copy_in) cudaStreamDestroy(p->copy_in);
    if (p->copy_out) cudaStreamDestroy(p->copy_out);
    delete p;
}

int qcb_plan_info(qcb_plan *p, qcb_plan_info_t *out)
{
    if (!p || !out) return fail("NULL argument");
    out->device = p->device; out->sm_count = p->sm_count;
    out->fast_adapter = (!p->force_generic && p->fast.adapter_ok) ? 1 : 0;
    out->fast_barcode = (!p->force_generic && p->fast.barcode_ok) ? 1 : 0;
    out->max_group_size = p->max_group;
    out->n_templates = p->t.n_templates;
    out->workspace_bytes = (int64_t)(p->wins.bytes + p->ad_score.bytes + p->ad_end.bytes + p->sel.bytes + p->bc_score.bytes +
                                     p->in_stage2[0].bytes + p->in_stage2[1].bytes + p->out_stage2[0].bytes + p->out_stage2[1].bytes + p->fast.workspace_bytes());
    out->kernel_launches = p->launches;
    return 0;
}

int qcb_plan_set_force_generic(qcb_plan *p, int force)
{
    if (!p) return fail("plan is NULL");
    p->force_generic = force != 0;
    p->no_row_chunks = force == 2;
    return 0;
}

int qcb_plan_set_profiling(qcb_plan *p, int enable)
{
    if (!p) return fail("plan is NULL");
    p->profiling = enable != 0;
    return 0;
}

int qcb_plan_stage_times(qcb_plan *p, double *ms, int64_t *launches, int reset)
{
    if (!p) return fail("plan is NULL");
    QCB_CUDA(cudaSetDevice(p->device));
    for (auto &ev : p->stage_events) {
        QCB_CUDA(cudaEventSynchronize(ev.b));
        float t = 0.f;
        QCB_CUDA(cudaEventElapsedTime(&t, ev.a, ev.b));
        p->stage_ms[ev.stage] += t;
        p->stage_launches[ev.stage] += ev.launches;
        cudaEventDestroy(ev.a); cudaEventDestroy(ev.b);
    }
    p->stage_events.clear();
    for (int i = 0; i < QCB_N_STAGES; ++i) {
        if (ms) ms[i] = p->stage_ms[i];
        if (launches) launches[i] = p->stage_launches[i];
        if (reset) { p->stage_ms[i] = 0; p->stage_launches[i] = 0; }
    }
    return 0;
}

// Shared body of qcb_sg_batch / qcb_sg_stats_batch: a parity-test primitive (everything is allocated and released per
// call; the production path never goes through here).  n_out = 3 (score, end_query, end_ref) or 6 (+ statistics).
static int sg_batch_impl(int device, const uint8_t *queries, const int32_t *query_off, int32_t n_queries,
                         const uint8_t *refs, const int32_t *ref_off, int32_t n_refs, int32_t open, int32_t extend,
                         const int32_t *matrix, int32_t msize, const uint8_t *mapper, int32_t *const *outs, int n_out)
{
    if (n_queries < 0 || n_refs < 0) return fail("negative count");
    if (n_queries == 0 || n_refs == 0) return 0;
    if (!queries || !query_off || !refs || !ref_off || !matrix || !mapper) return fail("NULL buffer");
    for (int i = 0; i < n_out; ++i) if (!outs[i]) return fail("NULL buffer");
    if (msize <= 0 || msize > kMaxMatrix) return fail("matrix size must be in [1, %d]", kMaxMatrix);
    for (int i = 0; i < 256; ++i) if (mapper[i] >= msize) return fail("mapper entry %d out of range", i);
    for (int r = 0; r < n_refs; ++r) {
        int m = ref_off[r + 1] - ref_off[r];
        if (m < 0 || m > kMaxTemplate) return fail("reference %d length %d outside [0, %d]", r, m, kMaxTemplate);
    }
    int ndev = qcb_device_count();
    if (ndev <= 0) return fail("no CUDA device available (libqcat_b200 has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail("device %d out of range", device);
    QCB_CUDA(cudaSetDevice(device));
    const size_t qb = (size_t)query_off[n_queries], rb = (size_t)ref_off[n_refs];
    const size_t pairs = (size_t)n_queries * n_refs;
    // one device slab: queries | refs | mapper | query offsets | ref offsets | matrix | outputs
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t o_q = 0, o_r = up(qb + 16), o_map = o_r + up(rb + 16), o_qo = o_map + 256;
    const size_t o_ro = o_qo + up((size_t)(n_queries + 1) * 4), o_mat = o_ro + up((size_t)(n_refs + 1) * 4);
    const size_t o_out = o_mat + up((size_t)msize * msize * 4), total = o_out + pairs * 4 * n_out;
    uint8_t *d = nullptr;
    QCB_CUDA(cudaMalloc(&d, total));
    int rc = 0;
#define SG_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail("%s failed: %s", #call, cudaGetErrorString(e_)); cudaFree(d); return rc; } } while (0)
    SG_CUDA(cudaMemcpy(d + o_q, queries, qb, cudaMemcpyHostToDevice));
    SG_CUDA(cudaMemcpy(d + o_r, refs, rb, cudaMemcpyHostToDevice));
    SG_CUDA(cudaMemcpy(d + o_map, mapper, 256, cudaMemcpyHostToDevice));
    SG_CUDA(cudaMemcpy(d + o_qo, query_off, (size_t)(n_queries + 1) * 4, cudaMemcpyHostToDevice));
    SG_CUDA(cudaMemcpy(d + o_ro, ref_off, (size_t)(n_refs + 1) * 4, cudaMemcpyHostToDevice));
    SG_CUDA(cudaMemcpy(d + o_mat, matrix, (size_t)msize * msize * 4, cudaMemcpyHostToDevice));
    int32_t *d_out = (int32_t *)(d + o_out);
    if (n_out == 3)
        k_sg_batch<<<grid_for((long long)pairs, 128), 128>>>(d + o_q, (const int32_t *)(d + o_qo), n_queries, d + o_r,
                                                            (const int32_t *)(d + o_ro), n_refs, open, extend,
                                                            (const int32_t *)(d + o_mat), msize, d + o_map,
                                                            d_out, d_out + pairs, d_out + 2 * pairs);
    else
        k_sg_stats_batch<<<grid_for((long long)pairs, 64), 64>>>(d + o_q, (const int32_t *)(d + o_qo), n_queries, d + o_r,
                                                                (const int32_t *)(d + o_ro), n_refs, open, extend,
                                                                (const int32_t *)(d + o_mat), msize, d + o_map, d_out);
    SG_CUDA(cudaGetLastError());
    for (int i = 0; i < n_out; ++i) SG_CUDA(cudaMemcpy(outs[i], d_out + (size_t)i * pairs, pairs * 4, cudaMemcpyDeviceToHost));
#undef SG_CUDA
    cudaFree(d);
    return 0;
}

int qcb_sg_batch(int device, const uint8_t *queries, const int32_t *query_off, int32_t n_queries,
                 const uint8_t *refs, const int32_t *ref_off, int32_t n_refs, int32_t open, int32_t extend,
                 const int32_t *matrix, int32_t msize, const uint8_t *mapper,
                 int32_t *score, int32_t *end_query, int32_t *end_ref)
{
    int32_t *outs[3] = {score, end_query, end_ref};
    return sg_batch_impl(device, queries, query_off, n_queries, refs, ref_off, n_refs, open, extend, matrix, msize, mapper, outs, 3);
}

int qcb_sg_stats_batch(int device, const uint8_t *queries, const int32_t *query_off, int32_t n_queries,
                       const uint8_t *refs, const int32_t *ref_off, int32_t n_refs, int32_t open, int32_t extend,
                       const int32_t *matrix, int32_t msize, const uint8_t *mapper,
                       int32_t *score, int32_t *end_query, int32_t *end_ref, int32_t *matches, int32_t *similar, int32_t *length)
{
    int32_t *outs[6] = {score, end_query, end_ref, matches, similar, length};
    return sg_batch_impl(device, queries, query_off, n_queries, refs, ref_off, n_refs, open, extend, matrix, msize, mapper, outs, 6);
}

int qcb_detect(qcb_plan *plan, const uint8_t *win5, const uint8_t *tail3, int32_t stride, const int32_t *wlen,
               const int64_t *read_len, int64_t n_reads, const int32_t *subset, int32_t n_subset, qcb_result *out)
{
    return detect_host_impl(plan, win5, tail3, stride, wlen, read_len, n_reads, subset, n_subset, out, nullptr);
}

int qcb_detect_device(qcb_plan *plan, const uint8_t *d_win5, const uint8_t *d_tail3, int32_t stride, const int32_t *d_wlen,
                      const int64_t *d_read_len, int64_t n_reads, const int32_t *subset, int32_t n_subset,
                      qcb_result *d_out, void *stream)
{
    if (!d_win5 || !d_tail3 || !d_wlen || !d_read_len || !d_out) return n_reads == 0 ? 0 : fail("NULL device buffer");
    return detect_device_impl(plan, d_win5, d_tail3, stride, d_wlen, d_read_len, n_reads, subset, n_subset, d_out, nullptr,
                              (cudaStream_t)stream);
}

int qcb_detect_auto(qcb_plan *plan, const uint8_t *win5, const uint8_t *tail3, int32_t stride, const int32_t *wlen,
                    const int64_t *read_len, int64_t n_reads, const int32_t *kit_of_layout, int32_t batch_size,
                    qcb_result *out, int32_t *batch_kit)
{
    if (!kit_of_layout) return fail("kit_of_layout is NULL");
    if (!tail3) return fail("NULL input/output buffer");
    return detect_host_impl(plan, win5, tail3, stride, wlen, read_len, n_reads, nullptr, 0, out, nullptr, kit_of_layout, batch_size,
                            batch_kit);
}

int qcb_detect_auto_device(qcb_plan *plan, const uint8_t *d_win5, const uint8_t *d_tail3, int32_t stride, const int32_t *d_wlen,
                           const int64_t *d_read_len, int64_t n_reads, const int32_t *kit_of_layout, int32_t batch_size,
                           qcb_result *d_out, int32_t *d_batch_kit, void *stream)
{
    if (!d_win5 || !d_tail3 || !d_wlen || !d_read_len || !d_out) return n_reads == 0 ? 0 : fail("NULL device buffer");
    AutoCall ac{kit_of_layout, batch_size, d_batch_kit};
    return detect_device_impl(plan, d_win5, d_tail3, stride, d_wlen, d_read_len, n_reads, nullptr, 0, d_out, nullptr,
                              (cudaStream_t)stream, &ac);
}

int qcb_plan_base_classes(qcb_plan *plan, uint8_t *cls)
{
    if (!plan || !cls) { fail("NULL argument"); return 0; }
    if (plan->n_classes > 0) memcpy(cls, plan->h_cls, 256);
    return plan->n_classes;
}

int qcb_detect4(qcb_plan *plan, const uint8_t *win5p, const uint8_t *tail3p, int32_t stride4, const int32_t *wlen,
                const int64_t *read_len, int64_t n_reads, const int32_t *subset, int32_t n_subset, qcb_result *out)
{
    if (!tail3p) return fail("NULL input/output buffer");
    return detect_host_impl(plan, win5p, tail3p, stride4, wlen, read_len, n_reads, subset, n_subset, out, nullptr, nullptr, 0, nullptr,
                            nullptr, true);
}

int qcb_detect4_device(qcb_plan *plan, const uint8_t *d_win5p, const uint8_t *d_tail3p, int32_t stride4, const int32_t *d_wlen,
                       const int64_t *d_read_len, int64_t n_reads, const int32_t *subset, int32_t n_subset,
                       qcb_result *d_out, void *stream)
{
    if (!d_win5p || !d_tail3p || !d_wlen || !d_read_len || !d_out) return n_reads == 0 ? 0 : fail("NULL device buffer");
    return detect_device_impl(plan, d_win5p, d_tail3p, stride4, d_wlen, d_read_len, n_reads, subset, n_subset, d_out, nullptr,
                              (cudaStream_t)stream, nullptr, true);
}

int qcb_detect_auto4(qcb_plan *plan, const uint8_t *win5p, const uint8_t *tail3p, int32_t stride4, const int32_t *wlen,
                     const int64_t *read_len, int64_t n_reads, const int32_t *kit_of_layout, int32_t batch_size,
                     qcb_result *out, int32_t *batch_kit)
{
    if (!kit_of_layout) return fail("kit_of_layout is NULL");
    if (!tail3p) return fail("NULL input/output buffer");
    return detect_host_impl(plan, win5p, tail3p, stride4, wlen, read_len, n_reads, nullptr, 0, out, nullptr, kit_of_layout, batch_size,
                            batch_kit, nullptr, true);
}

// One host thread per plan; plan d takes the blocks d, d + D, d + 2 D, ... of `block` reads (round-robin at block
// granularity) through its own three-stream pipeline and writes its records at the reads' own positions.
static int detect_multi_impl(qcb_plan *const *plans, int32_t n_plans, const uint8_t *win5, const uint8_t *tail3, int32_t stride,
                             const int32_t *wlen, const int64_t *read_len, int64_t n_reads, const int32_t *subset, int32_t n_subset,
                             qcb_result *out, const int32_t *kit_of_layout, int32_t batch_size, int32_t *batch_kit)
{
    if (!plans || n_plans <= 0) return fail("no plans");
    for (int d = 0; d < n_plans; ++d) if (!plans[d]) return fail("plan %d is NULL", d);
    if (n_reads < 0) return fail("n_reads is negative");
    if (n_reads == 0) return 0;
    if (!win5 || !tail3 || !wlen || !read_len || !out) return fail("NULL input/output buffer");
    if (kit_of_layout && batch_size <= 0) return fail("batch_size must be positive");
    // block size: ~4 blocks per device for load balance, at least 64 Ki reads, whole CLI batches in auto-kit calls
    long long block = std::max<long long>(1 << 16, (n_reads / (4LL * n_plans) + 31) / 32 * 32);
    if (kit_of_layout) block = std::max<long long>(1, (block + batch_size - 1) / batch_size) * batch_size;
    std::vector<int> rcs(n_plans, 0);
    std::vector<std::string> errors(n_plans);
    std::vector<std::thread> pool;
    for (int d = 0; d < n_plans; ++d) {
        pool.emplace_back([&, d]() {
            const Shard shard{(long long)d * block, block, (long long)n_plans * block};
            rcs[d] = detect_host_impl(plans[d], win5, tail3, stride, wlen, read_len, n_reads, subset, n_subset, out, nullptr,
                                      kit_of_layout, batch_size, batch_kit, &shard);
            if (rcs[d]) errors[d] = g_error;                   // thread-local message of the worker
        });
    }
    for (auto &th : pool) th.join();
    for (int d = 0; d < n_plans; ++d)
        if (rcs[d]) return fail("device shard %d: %s", d, errors[d].c_str());
    return 0;
}

int qcb_detect_multi(qcb_plan *const *plans, int32_t n_plans, const uint8_t *win5, const uint8_t *tail3, int32_t stride,
                     const int32_t *wlen, const int64_t *read_len, int64_t n_reads, const int32_t *subset, int32_t n_subset,
                     qcb_result *out)
{
    return detect_multi_impl(plans, n_plans, win5, tail3, stride, wlen, read_len, n_reads, subset, n_subset, out, nullptr, 0, nullptr);
}

int qcb_detect_auto_multi(qcb_plan *const *plans, int32_t n_plans, const uint8_t *win5, const uint8_t *tail3, int32_t stride,
                          const int32_t *wlen, const int64_t *read_len, int64_t n_reads, const int32_t *kit_of_layout,
                          int32_t batch_size, qcb_result *out, int32_t *batch_kit)
{
    if (!kit_of_layout) return fail("kit_of_layout is NULL");
    return detect_multi_impl(plans, n_plans, win5, tail3, stride, wlen, read_len, n_reads, nullptr, 0, out, kit_of_layout,
                             batch_size, batch_kit);
}

// In-process all-gather of the per-device count vectors over peer copies (NVLink when the devices are peers).
int qcb_hist_allgather(qcb_plan *const *plans, int32_t n_plans, int64_t *const *d_counts, int32_t n_bins, int64_t *const *d_gathered)
{
    if (!plans || n_plans <= 0 || !d_counts || !d_gathered || n_bins <= 0) return fail("bad argument");
    for (int i = 0; i < n_plans; ++i) if (!plans[i] || !d_counts[i] || !d_gathered[i]) return fail("NULL entry %d", i);
    const size_t bytes = (size_t)n_bins * sizeof(int64_t);
    for (int i = 0; i < n_plans; ++i) {
        QCB_CUDA(cudaSetDevice(plans[i]->device));
        for (int j = 0; j < n_plans; ++j) {
            if (plans[i]->device != plans[j]->device) {
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, plans[i]->device, plans[j]->device) == cudaSuccess && can) {
                    cudaError_t e = cudaDeviceEnablePeerAccess(plans[j]->device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                    else if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                }
            }
            QCB_CUDA(cudaMemcpyPeerAsync(d_gathered[i] + (size_t)j * n_bins, plans[i]->device, d_counts[j], plans[j]->device, bytes,
                                         plans[i]->stream));
        }
    }
    for (int i = 0; i < n_plans; ++i) {
        QCB_CUDA(cudaSetDevice(plans[i]->device));
        QCB_CUDA(cudaStreamSynchronize(plans[i]->stream));
    }
    return 0;
}

int qcb_scan(qcb_plan *plan, const uint8_t *windows, int32_t stride, const int32_t *wlen, int64_t n_windows,
             const int32_t *subset, int32_t n_subset, qcb_result *out)
{
    return detect_host_impl(plan, windows, nullptr, stride, wlen, nullptr, n_windows, subset, n_subset, out, nullptr);
}

int qcb_kit_vote(qcb_plan *plan, const uint8_t *win5, const uint8_t *tail3, int32_t stride, const int32_t *wlen,
                 int64_t n_reads, int32_t *vote_layout)
{
    if (!vote_layout) return n_reads == 0 ? 0 : fail("NULL output buffer");
    return detect_host_impl(plan, win5, tail3, stride, wlen, nullptr, n_reads, nullptr, 0, nullptr, vote_layout);
}

int qcb_kit_vote_device(qcb_plan *plan, const uint8_t *d_win5, const uint8_t *d_tail3, int32_t stride, const int32_t *d_wlen,
                        int64_t n_reads, int32_t *d_vote_layout, void *stream)
{
    if (!d_win5 || !d_tail3 || !d_wlen || !d_vote_layout) return n_reads == 0 ? 0 : fail("NULL device buffer");
    return detect_device_impl(plan, d_win5, d_tail3, stride, d_wlen, nullptr, n_reads, nullptr, 0, nullptr, d_vote_layout,
                              (cudaStream_t)stream);
}

int qcb_histogram_device(qcb_plan *p, const qcb_result *d_results, int64_t n_reads, const int32_t *layout_bin_base,
                         int64_t *d_counts, int32_t n_bins, void *stream)
{
    if (!p) return fail("plan is NULL");
    if (n_reads <= 0) return 0;
    if (!d_results || !layout_bin_base || !d_counts) return fail("NULL buffer");
    QCB_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    const std::vector<int32_t> base(layout_bin_base, layout_bin_base + p->t.n_layouts);
    if (base != p->bin_base_cached || !p->misc.ptr) {                 // the bin layout stays on the device between calls
        if (p->misc.reserve(base.size() * 4)) return 1;
        QCB_CUDA(cudaDeviceSynchronize());                            // an earlier call may still read the old table
        QCB_CUDA(cudaMemcpy(p->misc.ptr, base.data(), base.size() * 4, cudaMemcpyHostToDevice));
        p->bin_base_cached = base;
    }
    const unsigned grid = (unsigned)std::min<long long>(grid_for(n_reads, 256), (long long)p->sm_count * 8);
    k_histogram<<<grid, 256, 0, st>>>(d_results, n_reads, (const int32_t *)p->misc.ptr, (unsigned long long *)d_counts, n_bins);
    p->launches++;
    QCB_CUDA(cudaGetLastError());
    return 0;
}

int qcb_microbench_cell_rate(int device, double *cells_per_second, double *sm_mhz_effective)
{
    int ndev = qcb_device_count();
    if (ndev <= 0) return fail("no CUDA device available");
    if (device < 0 || device >= ndev) return fail("device %d out of range", device);
    QCB_CUDA(cudaSetDevice(device));
    std::string err;
    if (microbench_cell_rate(cells_per_second, sm_mhz_effective, err)) return fail("%s", err.c_str());
    return 0;
}

}  // extern "C"
