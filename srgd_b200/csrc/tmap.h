// Host-side TMA descriptor helper shared by conv_igemm.cu, linattn_fused.cu and attention_tc.cu.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace srgd {
// bf16 row-major [rows][inner] tensor (row pitch `row_stride_bytes`), box {box_inner, box_rows}, 128-byte swizzle,
// out-of-range elements read as zero.  Returns SRGD_OK or SRGD_E_CUDA (srgd_last_error() set).
int make_tmap_2d_bf16(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_rows, const char* what);
}  // namespace srgd
