// tmap.cuh — cached cuTensorMapEncodeTiled (driver entry point fetched at run time).
#pragma once
#include <cuda.h>
#include <stdint.h>

// rank <= 4.  dims[0] is the innermost dimension; strides[i] (bytes) is the pitch of dims[i+1].
int mc_make_tmap(CUtensorMap* tm, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle,
                 CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B);

// 2-D bf16 row-major [rows, cols], row pitch ld elements, box = [box_rows, 64 cols], 128 B swizzle.
int mc_make_tmap_2d_bf16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                         uint32_t box_rows);

// same with box = [box_rows, box_cols] for box_cols in {32, 64}: 64 B swizzle for the 32-column (64-byte) box.
int mc_make_tmap_2d_bf16_k(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                           uint32_t box_rows, uint32_t box_cols);
