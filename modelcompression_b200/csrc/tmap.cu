// tmap.cu — TMA descriptor construction + cache (process lifetime, keyed by every encode argument).
#include <mutex>
#include <unordered_map>
#include <string.h>
#include "common.cuh"
#include "tmap.cuh"

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct Key {
  uint64_t v[16];
  bool operator==(const Key& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 16; ++i) { h ^= k.v[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};

}  // namespace

int mc_make_tmap(CUtensorMap* tm, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle,
                 CUtensorMapL2promotion promo) {
  if (rank < 1 || rank > 4) return mc_set_error(MC_ERR_ARG, "mc_make_tmap: rank %d", rank);
  static std::mutex mu;
  static std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  Key key;
  memset(&key, 0, sizeof(key));
  key.v[0] = (uint64_t)(uintptr_t)base;
  key.v[1] = ((uint64_t)dtype << 32) | ((uint64_t)promo << 16) | ((uint64_t)rank << 8) | (uint64_t)swizzle;
  for (int i = 0; i < rank; ++i) {
    key.v[2 + i] = dims[i];
    key.v[6 + i] = (i + 1 < rank) ? strides_bytes[i] : 0;
    key.v[10 + i] = box[i];
  }
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *tm = it->second;
    return 0;
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return mc_set_error(MC_ERR_ARG, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[4];
  cuuint64_t gstr[3];
  cuuint32_t bx[4], estr[4];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUtensorMap out;
  CUresult r = fn(&out, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, promo,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return mc_set_error(MC_ERR_ARG, "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u",
                        (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
                        (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0), bx[0],
                        rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, out);
  *tm = out;
  return 0;
}

int mc_make_tmap_2d_bf16_k(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                           uint32_t box_rows, uint32_t box_cols) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[1] = {ld * 2};
  const uint32_t box[2] = {box_cols, box_rows};
  // the swizzle span equals the box row: 64 bf16 = 128 B, 32 bf16 = 64 B
  return mc_make_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box,
                      box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
}

int mc_make_tmap_2d_bf16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                         uint32_t box_rows) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[1] = {ld * 2};
  const uint32_t box[2] = {64, box_rows};
  return mc_make_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}
