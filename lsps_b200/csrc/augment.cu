// Input-pipeline kernel: the per-sample depth-crop augmentation of the reference's DataLoader
// workers (data/dataset_hand2.py:34-119 augmentCrop; utils/handdetector.py:682-808; cv2 nearest-neighbour warps) for a
// whole batch in one launch.  The host draws the random parameters in the reference's order and turns them into one
// lsps_aug_sample per crop (lsps_b200/augment.py); the device does everything that touches pixels.  The arithmetic is
// augment_core.h, shared with a host build that is checked bit-for-bit against the numpy oracle on the CPU.
// HBM-/latency-trivial (64 KB in + 64 KB out per crop): the point is to take ~750 us of host time per sample off the
// critical path, not kernel speed.
#include "augment_core.h"
#include "common.h"

namespace {

// premax[i] = max over the de-normalised crop = fl(fl(max(img) * scale) + off)  (scale > 0: rounding is monotone)
__global__ void __launch_bounds__(256) aug_premax_kernel(const float* __restrict__ img, const lsps_aug_sample* __restrict__ ps,
                                                        float* __restrict__ premax) {
  __shared__ float sm[8];
  const int i = blockIdx.x;
  const float* src = img + (size_t)i * LSPS_AUG_SIZE * LSPS_AUG_SIZE;
  float m = -INFINITY;
  for (int k = threadIdx.x; k < LSPS_AUG_SIZE * LSPS_AUG_SIZE; k += 256) m = fmaxf(m, src[k]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, sm[w]);
    premax[i] = __fadd_rn(__fmul_rn(m, ps[i].dn_scale), ps[i].dn_off);
  }
}

__global__ void __launch_bounds__(256) aug_kernel(const float* __restrict__ img, const lsps_aug_sample* __restrict__ ps,
                                                 const float* __restrict__ premax, float* __restrict__ out) {
  const int i = blockIdx.y;
  const int k = blockIdx.x * 256 + threadIdx.x;
  const lsps_aug_sample p = ps[i];
  const float* src = img + (size_t)i * LSPS_AUG_SIZE * LSPS_AUG_SIZE;
  out[(size_t)i * LSPS_AUG_SIZE * LSPS_AUG_SIZE + k] =
      lsps_aug_pixel(src, p, premax[i], k % LSPS_AUG_SIZE, k / LSPS_AUG_SIZE);
}

}  // namespace

extern "C" int lsps_augment_crops(lsps_ctx* ctx, const float* img, const void* params, float* premax, float* out, int n,
                                  lsps_stream st_) {
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  if (!ctx || !img || !params || !premax || !out || n <= 0 || img == out)
    return lsps_set_error(ctx, LSPS_E_ARG, "augment_crops: arg (in-place is not supported: the warp is a gather)");
  const lsps_aug_sample* ps = static_cast<const lsps_aug_sample*>(params);
  aug_premax_kernel<<<n, 256, 0, st>>>(img, ps, premax);
  LSPS_CHECK_LAUNCH(ctx, "aug_premax");
  aug_kernel<<<dim3(LSPS_AUG_SIZE * LSPS_AUG_SIZE / 256, n), 256, 0, st>>>(img, ps, premax, out);
  LSPS_CHECK_LAUNCH(ctx, "augment_crops");
  return LSPS_OK;
}

extern "C" int lsps_aug_sample_bytes(void) { return (int)sizeof(lsps_aug_sample); }
