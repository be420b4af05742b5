// Internal definitions shared by the kernels and the C-ABI layer of libqcat_b200.so.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/qcat_b200.h"

namespace qcb {

constexpr int kMaxTemplate = 256;   // longest adapter / barcode template the generic kernel accepts
constexpr int kMaxMatrix = 16;      // largest substitution matrix dimension
constexpr int kFastMaxStride = 160; // the packed kernels handle windows of up to this many bases

// Device copy of qcb_tables.  Passed to kernels by value (pointers point into one device slab).
struct DevTables {
    int W, ext;
    int a_open, a_extend, b_open, b_extend;
    int amat_size, bmat_size;
    int mode;
    int n_layouts, n_groups, n_templates;
    double min_quality;
    const int32_t *amat;
    const uint8_t *amap;
    const int32_t *bmat;
    const uint8_t *bmap;
    const uint8_t *comp;
    const int32_t *adapter_off;
    const uint8_t *adapter_seq;
    const double *denom;
    const int32_t *bc_end;
    const int32_t *bc_len;
    const int32_t *group;
    const int32_t *trim_offset;
    const int32_t *is_double;
    const int32_t *group_off;
    const int32_t *tmpl_off;
    const uint8_t *tmpl_seq;
    const int32_t *tmpl_ident;
    const int32_t *group_tlen;   // [n_groups] common template length of the group, -1 if its templates differ in length
};

// Per-window decision taken after the adapter stage (scanner_epi2me.py:57-82 / scanner_dual.py:57-110).
struct WindowSel {
    int32_t layout;      // chosen layout (global index)
    int32_t end_query;   // aligned_adapter_end (before trim_offset)
    int32_t lo0, hi0;    // barcode region of set 0 inside the window, [lo, hi)
    int32_t lo1, hi1;    // set 1 (dual mode only)
    int32_t full;        // 1 = full-window branch (scanner_epi2me.py:82)
    int32_t pad;
};

}  // namespace qcb
