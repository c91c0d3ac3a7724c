// Context + helpers shared by the translation units of liblsps_b200.so
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/lsps_b200.h"

struct TmapKey {
  const void* ptr;
  uint32_t dims[5];
  uint32_t box[5];
  uint32_t rank;
  bool operator==(const TmapKey& o) const {
    if (ptr != o.ptr || rank != o.rank) return false;
    for (int i = 0; i < 5; ++i)
      if (dims[i] != o.dims[i] || box[i] != o.box[i]) return false;
    return true;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull ^ reinterpret_cast<uint64_t>(k.ptr);
    for (int i = 0; i < 5; ++i) {
      h = (h ^ k.dims[i]) * 1099511628211ull;
      h = (h ^ k.box[i]) * 1099511628211ull;
    }
    return static_cast<size_t>(h ^ k.rank);
  }
};

struct lsps_ctx {
  int device = 0;
  int num_sms = 148;
  long long launches = 0;
  std::string err;
  void* encode_fn = nullptr;  // cuTensorMapEncodeTiled
  std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> tmaps;
};

int lsps_set_error(lsps_ctx* ctx, int code, const char* fmt, ...);
// dense bf16 tensor map (rank 2 or 5), 128-byte swizzle, zero OOB fill; cached per (ptr, dims, box)
int lsps_get_tmap(lsps_ctx* ctx, const void* ptr, int rank, const uint32_t* dims, const uint32_t* box,
                  CUtensorMap* out);

#define LSPS_CHECK_LAUNCH(ctx, name)                                                      \
  do {                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) return lsps_set_error(ctx, LSPS_E_CUDA, "%s: %s", name, cudaGetErrorString(e__)); \
    (ctx)->launches++;                                                                    \
  } while (0)
