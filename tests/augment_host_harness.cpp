// Host build of the augmentation kernel's per-pixel function (lsps_b200/csrc/augment_core.h) for the CPU parity test.
// g++ -O2 -ffp-contract=off -shared -fPIC: every floating-point operation individually rounded, like the *_rn
// intrinsics of the device build.
#include "../lsps_b200/csrc/augment_core.h"

extern "C" int aug_host_sample_bytes() { return (int)sizeof(lsps_aug_sample); }

extern "C" void aug_host(const float* img, const lsps_aug_sample* ps, float* out, int n) {
  const int S = LSPS_AUG_SIZE;
  for (int i = 0; i < n; ++i) {
    const float* src = img + (size_t)i * S * S;
    float m = src[0];
    for (int k = 1; k < S * S; ++k) m = src[k] > m ? src[k] : m;
    const float premax = m * ps[i].dn_scale + ps[i].dn_off;
    for (int k = 0; k < S * S; ++k) out[(size_t)i * S * S + k] = lsps_aug_pixel(src, ps[i], premax, k % S, k / S);
  }
}
