// Packed (u16x2) linear-gap kernels -- placeholder until the fast path lands: every plan uses the generic kernels.
#pragma once

#include <string>
#include "plan.h"

namespace qcb {

struct FastPlan {
    bool adapter_ok = false;
    bool barcode_ok = false;
    std::string error;
    size_t workspace_bytes() const { return 0; }
};

inline int fast_plan_build(FastPlan &, const qcb_tables *, int) { return 0; }
inline void fast_plan_free(FastPlan &) {}

inline int fast_adapter_stage(FastPlan &, const DevTables &, const uint8_t *, int, const int32_t *, long long,
                              const int32_t *, const int32_t *, int, int32_t *, int32_t *, cudaStream_t, long long *)
{
    return 1;
}

inline int fast_barcode_stage(FastPlan &, const DevTables &, const uint8_t *, int, long long, const WindowSel *, int, int,
                              int32_t *, cudaStream_t, long long *)
{
    return 1;
}

}  // namespace qcb
