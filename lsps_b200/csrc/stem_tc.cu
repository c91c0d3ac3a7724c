// Cin=1 7x7 stems on the tensor cores.
//
// K = 49 taps is not a TMA/implicit-GEMM shape (one channel = 2 bytes per pixel), so the im2col tile is built by the
// CTA's own threads straight into the 128-byte-swizzled shared-memory image that tcgen05.mma reads: one row per output
// pixel, 64 bf16 "taps" per row (49 real + a constant-one tap for the bias gradient + zero padding).  The fp32 image
// is split into bf16 hi + lo parts (two tiles, two MMA passes) so that the input keeps 16 mantissa bits.
//
//   forward : y[128 px, 64 co]   = [hi | lo][128 px, 64 taps] * W[64 co, 64 taps]^T     (K-major A and B)
//   wgrad   : D[128 tap, 64 co] += [hi | lo][128 px, 64 taps]^T * dy[128 px, 64 co]     (MN-major A and B; dy via TMA)
//
// Replaces the direct fp32 kernels (aux.cu stem_fwd / stem_wgrad: ~3100 FFMA per pixel) for
//   LeakyReLUConv2d(1, 64, 7, {1,2}, 3)  --  /root/reference/src/trainers/lsps_nets.py:104,186
#include "common.h"
#include "ptx.cuh"

using namespace lsps;
typedef __nv_bfloat16 bf16;

namespace {

constexpr int TILE_PX = 128;
constexpr int A_BYTES = TILE_PX * 128;  // one swizzled operand tile: 128 rows x 64 bf16

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// builds rows of the hi / lo tap tiles for pixel `row` of the tile from the fp32 patch
//   patch: [(TR-1)*S+7][PW] floats, pixel (ty, tx) reads patch[ty*S + r][tx*S + c]
template <int S>
__device__ __forceinline__ void build_tap_row(uint8_t* a_hi, uint8_t* a_lo, const float* patch, int pw, int row, int ty,
                                              int tx) {
  const float* p0 = patch + ty * S * pw + tx * S;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int tap = j * 8 + k;
      v[k] = tap < 49 ? p0[(tap / 7) * pw + tap % 7] : (tap == 49 ? 1.0f : 0.0f);  // tap 49 == 1: bias gradient
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
      const float2 hf = __bfloat1622float2(h2);
      hi[k] = *reinterpret_cast<const uint32_t*>(&h2);
      lo[k] = pack_bf16x2(v[2 * k] - hf.x, v[2 * k + 1] - hf.y);
    }
    const int off = row * 128 + ((j ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

template <int S>
__device__ __forceinline__ void load_patch(float* patch, int pw, int prow, const float* im, int h, int wd, int iy0) {
  for (int i = threadIdx.x; i < prow * pw; i += blockDim.x) {
    const int py = i / pw, px = i - py * pw;
    const int iy = iy0 + py, ix = px - 3;
    patch[i] = (iy >= 0 && iy < h && ix >= 0 && ix < wd) ? __ldg(im + (long long)iy * wd + ix) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ forward
// SPLIT: the weights are split into bf16 hi + lo as well (third MMA pass: taps_hi x w_lo) and the output is written as
// [hi | lo] channel halves (128 channels per pixel) -- the "bf16x3" discriminator path.
template <int S, bool SPLIT>
__global__ void __launch_bounds__(128) stem_fwd_tc_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                         const float* __restrict__ bias, bf16* __restrict__ y, int n,
                                                         int h, int wd, float slope) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_hi = base;
  uint8_t* a_lo = base + A_BYTES;
  uint8_t* wt = base + 2 * A_BYTES;                   // [64 co][64 taps] bf16, K-major, swizzled: 8 KB (+ 8 KB lo)
  float* patch = reinterpret_cast<float*>(wt + 16384);
  const int wo = wd / S, ho = h / S;
  const int tr = TILE_PX / wo;                        // output rows per tile (1 or 2)
  const int prow = (tr - 1) * S + 7, pw = wd + 8;
  float* sbias = patch + prow * pw;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sbias + 64);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(tmem_slot, 64); tmem_relinquish(); }
  if (threadIdx.x < 64) {
    const int co = threadIdx.x;
    sbias[co] = bias[co];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t u[4], ul[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int t0 = j * 8 + 2 * k;
        const float w0 = t0 < 49 ? w[co * 49 + t0] : 0.f, w1 = t0 + 1 < 49 ? w[co * 49 + t0 + 1] : 0.f;
        u[k] = pack_bf16x2(w0, w1);
        ul[k] = pack_bf16x2(w0 - bf16lo(u[k]), w1 - bf16hi(u[k]));
      }
      *reinterpret_cast<uint4*>(wt + co * 128 + ((j ^ (co & 7)) << 4)) = make_uint4(u[0], u[1], u[2], u[3]);
      if (SPLIT) *reinterpret_cast<uint4*>(wt + 8192 + co * 128 + ((j ^ (co & 7)) << 4)) = make_uint4(ul[0], ul[1], ul[2], ul[3]);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
  constexpr int OC = SPLIT ? 128 : 64;                // channels per output pixel

  const int tiles_per_img = ho / tr;
  const int total = tiles_per_img * n;
  const int row = threadIdx.x, ty = row / wo, tx = row - ty * wo;
  uint32_t phase = 0;
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    const int im_i = t / tiles_per_img, oy0 = (t - im_i * tiles_per_img) * tr;
    load_patch<S>(patch, pw, prow, img + (long long)im_i * h * wd, h, wd, oy0 * S - 3);
    __syncthreads();
    build_tap_row<S>(a_hi, a_lo, patch, pw, row, ty, tx);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bw = smem_u32(wt);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem, umma_smem_desc(ah + k * 32, 0, 1024), umma_smem_desc(bw + k * 32, 0, 1024), idesc, k > 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem, umma_smem_desc(al + k * 32, 0, 1024), umma_smem_desc(bw + k * 32, 0, 1024), idesc, 1u);
      if (SPLIT) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem, umma_smem_desc(ah + k * 32, 0, 1024), umma_smem_desc(bw + 8192 + k * 32, 0, 1024), idesc, 1u);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    uint4* o = reinterpret_cast<uint4*>(y + (((long long)im_i * ho + oy0 + ty) * wo + tx) * OC);
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(taddr + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float a = __uint_as_float(v[8 * j + k]) + sbias[c0 + 8 * j + k];
          f[k] = a > 0.f ? a : a * slope;
        }
        uint4 u;
        u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
        u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
        o[c0 / 8 + j] = u;
        if (SPLIT) {
          uint4 l;
          l.x = pack_bf16x2(f[0] - bf16lo(u.x), f[1] - bf16hi(u.x)); l.y = pack_bf16x2(f[2] - bf16lo(u.y), f[3] - bf16hi(u.y));
          l.z = pack_bf16x2(f[4] - bf16lo(u.z), f[5] - bf16hi(u.z)); l.w = pack_bf16x2(f[6] - bf16lo(u.w), f[7] - bf16hi(u.w));
          o[8 + c0 / 8 + j] = l;
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // TMEM, operand tiles and patch are reused by the next tile
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

// ------------------------------------------------------------------------------------------------ wgrad
// D[tap rows: 64 hi + 64 lo][co] accumulated in TMEM over all tiles of this CTA; one red.add epilogue at the end.
// SPLIT: dy comes as [hi | lo] channel halves (two TMA boxes); second MMA pass with the lo tile.
template <int S, bool SPLIT>
__global__ void __launch_bounds__(128) stem_wgrad_tc_kernel(const float* __restrict__ img,
                                                           const __grid_constant__ CUtensorMap tmDy,
                                                           float* __restrict__ dw, float* __restrict__ db, int n, int h,
                                                           int wd) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // two buffers of {taps hi, taps lo, dy}: the MMAs of tile i run while the threads build tile i+1.  The split variant
  // carries a fourth tile (dy lo) and keeps ONE buffer: two would be 132 KB and halve the CTAs per SM (measured slower).
  constexpr int BUF = (SPLIT ? 4 : 3) * A_BYTES;
  constexpr int NBUF = SPLIT ? 1 : 2;
  const int wo = wd / S, ho = h / S;
  const int tr = TILE_PX / wo;
  const int prow = (tr - 1) * S + 7, pw = wd + 8;
  float* patch = reinterpret_cast<float*>(base + NBUF * BUF);
  uint64_t* full = reinterpret_cast<uint64_t*>(patch + prow * pw);   // dy tile landed (per buffer)
  uint64_t* done = full + 2;                                          // MMAs of the buffer retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 2);

  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&full[i], 1); mbar_init(&done[i], 1); }
    fence_barrier_init();
    tma_prefetch_desc(&tmDy);
  }
  if (warp == 0) { tmem_alloc(tmem_slot, 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);

  const int tiles_per_img = ho / tr;
  const int total = tiles_per_img * n;
  const int row = threadIdx.x, ty = row / wo, tx = row - ty * wo;
  int it = 0;
  for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
    const int b = it % NBUF;
    const uint32_t use = (it / NBUF) & 1;
    uint8_t* buf = base + b * BUF;
    const int im_i = t / tiles_per_img, oy0 = (t - im_i * tiles_per_img) * tr;
    if (it >= NBUF) mbar_wait(&done[b], use ^ 1);   // MMAs that read this buffer NBUF tiles ago have retired
    if (threadIdx.x == 0) {
      mbar_expect_tx(&full[b], (SPLIT ? 2 : 1) * A_BYTES);
      tma_load_2d(buf + 2 * A_BYTES, &tmDy, &full[b], 0, (im_i * ho + oy0) * wo);
      if (SPLIT) tma_load_2d(buf + 3 * A_BYTES, &tmDy, &full[b], 64, (im_i * ho + oy0) * wo);
    }
    load_patch<S>(patch, pw, prow, img + (long long)im_i * h * wd, h, wd, oy0 * S - 3);
    __syncthreads();
    build_tap_row<S>(buf, buf + A_BYTES, patch, pw, row, ty, tx);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_wait(&full[b], use);
      tc_fence_after();
      const uint32_t at = smem_u32(buf), bd = smem_u32(buf + 2 * A_BYTES);
#pragma unroll
      for (int k = 0; k < 8; ++k)  // 16 pixels per MMA; A = taps (M: hi block, lo block A_BYTES apart), B = dy (N = 64 co)
        umma_bf16(tmem, umma_smem_desc(at + k * 2048, A_BYTES, 1024), umma_smem_desc(bd + k * 2048, A_BYTES, 1024), idesc,
                  (it > 0 || k > 0) ? 1u : 0u);
      if (SPLIT) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem, umma_smem_desc(at + k * 2048, A_BYTES, 1024),
                    umma_smem_desc(bd + A_BYTES + k * 2048, A_BYTES, 1024), idesc, 1u);
      }
      umma_commit(&done[b]);
    }
  }
  // drain: the last commit covers every MMA issued before it
  if (it > 0) {
    const int last = it - 1;
    mbar_wait(&done[last % NBUF], (last / NBUF) & 1);
    tc_fence_after();
    const int tap = (row & 63);
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(taddr + c0, v);
      tmem_ld_wait();
      if (tap < 49) {
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dw + (c0 + j) * 49 + tap, __uint_as_float(v[j]));
      } else if (tap == 49 && db) {
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(db + c0 + j, __uint_as_float(v[j]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

// ------------------------------------------------------------------------------------------------ dgrad
// dimg[y][x] = sum over taps (r,c), channels co of dy[(y+3-r)/S][(x+3-c)/S][co] * W[co][r][c]
//   step 1 (tensor cores): G[px][tap] = dy[px][0:64] . W[0:64][tap] for every dy pixel that touches the 16x8 input tile
//                          (one TMA box with halo, zero-filled outside the image)
//   step 2 (gather)      : each thread owns one input pixel and adds the <= 49 G entries that map onto it.
// SPLIT: dy as [hi | lo] halves (two TMA boxes), weights split hi + lo: dy_hi x w_hi + dy_lo x w_hi + dy_hi x w_lo.
template <int S, bool SPLIT>
__global__ void __launch_bounds__(128) stem_dgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDy,
                                                           const float* __restrict__ w, float* __restrict__ dimg,
                                                           int n, int h, int wd, int accumulate) {
  constexpr int BW = S == 1 ? 22 : 11, BH = S == 1 ? 14 : 7;  // dy patch touching a 16 x 8 input tile
  constexpr int ROWS = BW * BH, MT = (ROWS + 127) / 128;
  constexpr int GP = 49;                                      // G row pitch (odd: conflict-free)
  constexpr int TCOLS = MT == 1 ? 64 : 256;                   // TMEM columns: one 64-column accumulator per M tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NDY = SPLIT ? 2 : 1;
  uint8_t* dyb = base;                                        // MT operand tiles of 128 rows (x2: lo halves behind)
  uint8_t* wt = base + NDY * MT * A_BYTES;                    // [64 tap rows][64 co] bf16, K-major, swizzled (+ lo)
  float* G = reinterpret_cast<float*>(wt + 16384);
  uint64_t* full = reinterpret_cast<uint64_t*>(G + MT * 128 * GP);
  uint64_t* done = full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(full, 1); mbar_init(done, 1); fence_barrier_init(); tma_prefetch_desc(&tmDy); }
  if (warp == 0) { tmem_alloc(tmem_slot, TCOLS); tmem_relinquish(); }
  if (threadIdx.x < 64) {
    const int tap = threadIdx.x;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t u[4], ul[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int co = j * 8 + 2 * k;
        const float w0 = tap < 49 ? w[co * 49 + tap] : 0.f, w1 = tap < 49 ? w[(co + 1) * 49 + tap] : 0.f;
        u[k] = pack_bf16x2(w0, w1);
        ul[k] = pack_bf16x2(w0 - bf16lo(u[k]), w1 - bf16hi(u[k]));
      }
      *reinterpret_cast<uint4*>(wt + tap * 128 + ((j ^ (tap & 7)) << 4)) = make_uint4(u[0], u[1], u[2], u[3]);
      if (SPLIT) *reinterpret_cast<uint4*>(wt + 8192 + tap * 128 + ((j ^ (tap & 7)) << 4)) = make_uint4(ul[0], ul[1], ul[2], ul[3]);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);

  const int ho = h / S, wo = wd / S;
  const int tiles_x = wd / 16, tiles_y = h / 8;
  const int total = tiles_x * tiles_y * n;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  auto issue_load = [&](int t) {
    const int bx = t % tiles_x, by = (t / tiles_x) % tiles_y, im_i = t / (tiles_x * tiles_y);
    const int oxb = S == 1 ? bx * 16 - 3 : bx * 8 - 1, oyb = S == 1 ? by * 8 - 3 : by * 4 - 1;
    mbar_expect_tx(full, NDY * ROWS * 128);
    tma_load_4d(dyb, &tmDy, full, 0, oxb, oyb, im_i);
    if (SPLIT) tma_load_4d(dyb + MT * A_BYTES, &tmDy, full, 64, oxb, oyb, im_i);
  };
  if (threadIdx.x == 0 && (int)blockIdx.x < total) issue_load(blockIdx.x);
  uint32_t phase = 0;
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    const int bx = t % tiles_x, by = (t / tiles_x) % tiles_y, im_i = t / (tiles_x * tiles_y);
    if (threadIdx.x == 0) {
      mbar_wait(full, phase);
      tc_fence_after();
      const uint32_t a0 = smem_u32(dyb), bw = smem_u32(wt);
#pragma unroll
      for (int m = 0; m < MT; ++m) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + m * 64, umma_smem_desc(a0 + m * A_BYTES + k * 32, 0, 1024), umma_smem_desc(bw + k * 32, 0, 1024),
                    idesc, k > 0 ? 1u : 0u);
        if (SPLIT) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + m * 64, umma_smem_desc(a0 + (MT + m) * A_BYTES + k * 32, 0, 1024),
                      umma_smem_desc(bw + k * 32, 0, 1024), idesc, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + m * 64, umma_smem_desc(a0 + m * A_BYTES + k * 32, 0, 1024),
                      umma_smem_desc(bw + 8192 + k * 32, 0, 1024), idesc, 1u);
        }
      }
      umma_commit(done);
    }
    mbar_wait(done, phase);
    tc_fence_after();
    // the dy buffer is free again: fetch the next tile's patch while this one is scattered
    if (threadIdx.x == 0 && t + (int)gridDim.x < total) issue_load(t + gridDim.x);
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const int grow = m * 128 + threadIdx.x;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + m * 64 + c0, v);
        tmem_ld_wait();
        if (grow < ROWS) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < 49) G[grow * GP + c0 + j] = __uint_as_float(v[j]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      const int ny = ty + 3 - r;
      if (S == 2 && (ny & 1)) continue;
      const int prow = S == 1 ? ny + 3 : (ny >> 1) + 1;
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        const int nx = tx + 3 - c;
        if (S == 2 && (nx & 1)) continue;
        const int pcol = S == 1 ? nx + 3 : (nx >> 1) + 1;
        acc += G[(prow * BW + pcol) * GP + r * 7 + c];
      }
    }
    float* o = dimg + ((long long)im_i * h + by * 8 + ty) * wd + bx * 16 + tx;
    *o = accumulate ? *o + acc : acc;
    phase ^= 1;
    __syncthreads();  // G is rewritten by the next tile
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

// one cudaFuncSetAttribute per kernel (keyed by the function pointer: several kernels share one signature)
template <typename K>
int set_smem(lsps_ctx* ctx, K kernel, int bytes) {
  static std::unordered_map<const void*, int> configured;
  const void* key = reinterpret_cast<const void*>(kernel);
  auto it = configured.find(key);
  if (it != configured.end() && it->second >= bytes) return LSPS_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return lsps_set_error(ctx, LSPS_E_CUDA, "stem smem attr: %s", cudaGetErrorString(e));
  configured[key] = bytes;
  return LSPS_OK;
}

}  // namespace

// Called from aux.cu's lsps_stem_fwd / lsps_stem_wgrad / lsps_stem_dgrad when the shape qualifies (output width 64 or
// 128).  split != 0: the bf16x3 variants (y / dy as [hi | lo] channel halves, weights split in the kernel).
int lsps_stem_fwd_tc(lsps_ctx* ctx, const float* img, const float* w, const float* bias, void* y, int n, int h, int wd,
                     int stride, float slope, int split, cudaStream_t st) {
  const int wo = wd / stride, ho = h / stride, tr = TILE_PX / wo;
  const int prow = (tr - 1) * stride + 7, pw = wd + 8;
  const int smem = 1024 + 2 * A_BYTES + 16384 + (prow * pw + 64) * 4 + 64;
  const int total = (ho / tr) * n;
  const int grid = total < 4 * ctx->num_sms ? total : 4 * ctx->num_sms;
  int rc;
#define LSPS_STEM_FWD(S_, SP_)                                                                                      \
  do {                                                                                                              \
    if ((rc = set_smem(ctx, stem_fwd_tc_kernel<S_, SP_>, smem))) return rc;                                          \
    stem_fwd_tc_kernel<S_, SP_><<<grid, 128, smem, st>>>(img, w, bias, static_cast<bf16*>(y), n, h, wd, slope);       \
  } while (0)
  if (stride == 1) { if (split) LSPS_STEM_FWD(1, true); else LSPS_STEM_FWD(1, false); }
  else { if (split) LSPS_STEM_FWD(2, true); else LSPS_STEM_FWD(2, false); }
#undef LSPS_STEM_FWD
  LSPS_CHECK_LAUNCH(ctx, "stem_fwd_tc");
  return LSPS_OK;
}

int lsps_stem_wgrad_tc(lsps_ctx* ctx, const float* img, const void* dy, float* dw, float* db, int n, int h, int wd,
                       int stride, int split, cudaStream_t st) {
  const int wo = wd / stride, ho = h / stride, tr = TILE_PX / wo;
  const int prow = (tr - 1) * stride + 7, pw = wd + 8;
  const int smem = 1024 + (split ? 4 : 6) * A_BYTES + prow * pw * 4 + 128;
  const int total = (ho / tr) * n;
  const int per_sm = split ? 3 : 2;
  const int grid = total < per_sm * ctx->num_sms ? total : per_sm * ctx->num_sms;
  CUtensorMap tm;
  uint32_t dims[2] = {split ? 128u : 64u, (uint32_t)((long long)n * ho * wo)}, box[2] = {64u, (uint32_t)TILE_PX};
  int rc = lsps_get_tmap(ctx, dy, 2, dims, box, &tm);
  if (rc) return rc;
#define LSPS_STEM_WG(S_, SP_)                                                                                       \
  do {                                                                                                              \
    if ((rc = set_smem(ctx, stem_wgrad_tc_kernel<S_, SP_>, smem))) return rc;                                        \
    stem_wgrad_tc_kernel<S_, SP_><<<grid, 128, smem, st>>>(img, tm, dw, db, n, h, wd);                                \
  } while (0)
  if (stride == 1) { if (split) LSPS_STEM_WG(1, true); else LSPS_STEM_WG(1, false); }
  else { if (split) LSPS_STEM_WG(2, true); else LSPS_STEM_WG(2, false); }
#undef LSPS_STEM_WG
  LSPS_CHECK_LAUNCH(ctx, "stem_wgrad_tc");
  return LSPS_OK;
}

int lsps_stem_dgrad_tc(lsps_ctx* ctx, const void* dy, const float* w, float* dimg, int n, int h, int wd, int stride,
                       int accumulate, int split, cudaStream_t st) {
  const int ho = h / stride, wo = wd / stride;
  const int rows = stride == 1 ? 22 * 14 : 11 * 7, mt = (rows + 127) / 128;
  const int smem = 1024 + (split ? 2 : 1) * mt * A_BYTES + 16384 + mt * 128 * 49 * 4 + 64;
  const int total = (wd / 16) * (h / 8) * n;
  const int grid = total < ctx->num_sms ? total : ctx->num_sms * (stride == 1 ? 1 : 2);
  CUtensorMap tm;
  uint32_t dims[4] = {split ? 128u : 64u, (uint32_t)wo, (uint32_t)ho, (uint32_t)n};
  uint32_t box[4] = {64u, stride == 1 ? 22u : 11u, stride == 1 ? 14u : 7u, 1u};
  int rc = lsps_get_tmap(ctx, dy, 4, dims, box, &tm);
  if (rc) return rc;
#define LSPS_STEM_DG(S_, SP_)                                                                                       \
  do {                                                                                                              \
    if ((rc = set_smem(ctx, stem_dgrad_tc_kernel<S_, SP_>, smem))) return rc;                                        \
    stem_dgrad_tc_kernel<S_, SP_><<<grid, 128, smem, st>>>(tm, w, dimg, n, h, wd, accumulate);                        \
  } while (0)
  if (stride == 1) { if (split) LSPS_STEM_DG(1, true); else LSPS_STEM_DG(1, false); }
  else { if (split) LSPS_STEM_DG(2, true); else LSPS_STEM_DG(2, false); }
#undef LSPS_STEM_DG
  LSPS_CHECK_LAUNCH(ctx, "stem_dgrad_tc");
  return LSPS_OK;
}
