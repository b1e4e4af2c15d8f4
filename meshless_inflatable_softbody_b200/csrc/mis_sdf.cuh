// mis_sdf.cuh -- DeepSDF contact query (extension; deepsdf.py:9-41).  Placeholder state
// until the GEMM chain lands; see DESIGN.md.
#pragma once
namespace mis {
struct SdfState {};
inline void sdf_free(SdfState&) {}
}  // namespace mis
