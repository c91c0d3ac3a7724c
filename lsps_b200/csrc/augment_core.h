// Per-pixel arithmetic of the input-pipeline kernel: de-normalise a 128x128 depth crop, warp it
// (nearest neighbour), threshold to the new cube and normalise again -- what the reference does per sample on the host in
//   data/dataset_hand2.py:34-119 (augmentCrop) -> utils/handdetector.py:682-808 (moveCoM / rotateHand / scaleHand /
//   recropHand) -> cv2.warpPerspective / cv2.warpAffine (INTER_NEAREST, BORDER_CONSTANT 0).
// One function, compiled twice: by nvcc into the kernel in augment.cu, and by g++ into the host harness of
// tests/test_augment_core_cpu.py, which checks it bit-for-bit against the numpy oracle without a GPU.  Every double
// operation is individually rounded (no fused multiply-add): device code uses the *_rn intrinsics, the host build uses
// -ffp-contract=off.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define LSPS_HD __host__ __device__ __forceinline__
#else
#define LSPS_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define LSPS_DMUL(a, b) __dmul_rn((a), (b))
#define LSPS_DADD(a, b) __dadd_rn((a), (b))
#define LSPS_DDIV(a, b) __ddiv_rn((a), (b))
#define LSPS_FMUL(a, b) __fmul_rn((a), (b))
#define LSPS_FADD(a, b) __fadd_rn((a), (b))
#define LSPS_FSUB(a, b) __fsub_rn((a), (b))
#define LSPS_FDIV(a, b) __fdiv_rn((a), (b))
#else
#define LSPS_DMUL(a, b) ((a) * (b))
#define LSPS_DADD(a, b) ((a) + (b))
#define LSPS_DDIV(a, b) ((a) / (b))
#define LSPS_FMUL(a, b) ((a) * (b))
#define LSPS_FADD(a, b) ((a) + (b))
#define LSPS_FSUB(a, b) ((a) - (b))
#define LSPS_FDIV(a, b) ((a) / (b))
#endif

enum { LSPS_AUG_NONE = 0, LSPS_AUG_PERSPECTIVE = 1, LSPS_AUG_AFFINE = 2 };
#define LSPS_AUG_SIZE 128

// Host-computed per-sample parameters (lsps_b200/augment.py).  All thresholds are float32 values computed exactly as
// the reference's numpy expressions compute them.
typedef struct {
  int mode;          // LSPS_AUG_*
  int pad_;
  double m[9];       // PERSPECTIVE: inverse 3x3 (destination -> source), cv::invert closed form, row-major
                     // AFFINE: m[0..5] = inverted 2x3 (cv::invertAffineTransform), row-major
  float dn_scale;    // de-normalise: v = img * dn_scale + dn_off           (old cube_z / 2, old com_z)
  float dn_off;
  float zstart;      // recropHand z-threshold of the perspective modes     (com_z -/+ old cube_z / 2 of the com used there)
  float zend;
  float far_;        // final clamp + normalise with the NEW com / cube      (com_z + cube_z / 2)
  float near_;       //                                                      (com_z - cube_z / 2)
  float out_off;     // new com_z
  float out_scale;   // new cube_z / 2
} lsps_aug_sample;

LSPS_HD long long lsps_aug_round_even(double v) {   // cvRound / saturate_cast<int>(double)
  double r = nearbyint(v);
  if (r < -2147483648.0) r = -2147483648.0;
  if (r > 2147483647.0) r = 2147483647.0;
  return (long long)r;
}

// `img`: the NORMALISED crop (output of normalize(), background +1); premax: max of the de-normalised crop.
LSPS_HD float lsps_aug_pixel(const float* img, const lsps_aug_sample& p, float premax, int x, int y) {
  const int S = LSPS_AUG_SIZE;
  float v;
  if (p.mode == LSPS_AUG_NONE) {
    v = LSPS_FADD(LSPS_FMUL(img[y * S + x], p.dn_scale), p.dn_off);
  } else {
    long long sx, sy;
    bool inside;
    if (p.mode == LSPS_AUG_PERSPECTIVE) {
      const double dx = (double)x, dy = (double)y;
      const double W = LSPS_DADD(LSPS_DADD(LSPS_DMUL(p.m[6], dx), LSPS_DMUL(p.m[7], dy)), p.m[8]);
      const double fX = LSPS_DDIV(LSPS_DADD(LSPS_DADD(LSPS_DMUL(p.m[0], dx), LSPS_DMUL(p.m[1], dy)), p.m[2]), W);
      const double fY = LSPS_DDIV(LSPS_DADD(LSPS_DADD(LSPS_DMUL(p.m[3], dx), LSPS_DMUL(p.m[4], dy)), p.m[5]), W);
      // OpenCV 4.13: border test on the continuous coordinate, then floor(c + 0.5)   (NaN compares false -> border)
      inside = fX >= 0.0 && fX <= (double)(S - 1) && fY >= 0.0 && fY <= (double)(S - 1);
      sx = inside ? (long long)floor(LSPS_DADD(fX, 0.5)) : 0;
      sy = inside ? (long long)floor(LSPS_DADD(fY, 0.5)) : 0;
    } else {
      // cv2.warpAffine nearest: 10-bit fixed point, round-to-nearest-even tables for the x and the y terms
      const double sc = 1024.0;
      const long long ad = lsps_aug_round_even(LSPS_DMUL(LSPS_DMUL(p.m[0], (double)x), sc));
      const long long bd = lsps_aug_round_even(LSPS_DMUL(LSPS_DMUL(p.m[3], (double)x), sc));
      const long long X0 = lsps_aug_round_even(LSPS_DMUL(LSPS_DADD(LSPS_DMUL(p.m[1], (double)y), p.m[2]), sc)) + 512;
      const long long Y0 = lsps_aug_round_even(LSPS_DMUL(LSPS_DADD(LSPS_DMUL(p.m[4], (double)y), p.m[5]), sc)) + 512;
      sx = (X0 + ad) >> 10;
      sy = (Y0 + bd) >> 10;
      inside = sx >= 0 && sx < S && sy >= 0 && sy < S;
    }
    v = inside ? LSPS_FADD(LSPS_FMUL(img[sy * S + sx], p.dn_scale), p.dn_off) : 0.f;
    if (p.mode == LSPS_AUG_PERSPECTIVE) {
      // recropHand: numpy.isclose(v, 32000.) -> background ; clamp to the front plane, cut behind the back plane
      if (fabs((double)v - 32000.0) <= 1e-8 + 1e-5 * 32000.0) v = 0.f;
      if (v < p.zstart && v != 0.f) v = p.zstart;
      else if (v > p.zend && v != 0.f) v = 0.f;
    }
  }
  // augmentCrop tail (normZeroOne False): far plane for the old maximum and for holes, clamp, normalise with the new com / cube
  if (v == premax) v = p.far_;
  if (v == 0.f) v = p.far_;
  if (v >= p.far_) v = p.far_;
  if (v <= p.near_) v = p.near_;
  return LSPS_FDIV(LSPS_FSUB(v, p.out_off), p.out_scale);
}
