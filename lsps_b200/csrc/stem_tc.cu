// Cin=1 7x7 stems on the tensor cores.
//
// K = 49 taps is not a TMA/implicit-GEMM shape (one channel = 2 bytes per pixel), so the im2col tile is built by the
// CTA's own threads straight into the 128-byte-swizzled shared-memory image that tcgen05.mma reads: one row per output
// pixel, 64 bf16 "taps" per row.  The K order is k = 8 r + c (filter row r, column c < 7; k = 8 r + 7 carries a zero
// weight, k = 56 is a constant one for the bias gradient), so one 16-byte chunk of a row is 8 CONSECUTIVE patch
// elements of one patch row: the fp32 patch is split ONCE per tile into bf16 hi + lo planes (the input keeps 16
// mantissa bits: two tiles, two MMA passes) and the rows are assembled from them with word loads + funnel shifts.
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

// tap index k = 8 r + c of the operand rows <-> tap r * 7 + c of w[co][49]
__device__ __forceinline__ int k_of_tap(int tap) { const int r = tap / 7; return 8 * r + (tap - 7 * r); }

// The patch of a tile: two bf16 planes (hi, lo) of [prow][pwb = wd + 16] elements, image column ix at px = ix + 4.  A tile
// spans full image rows, so the halo columns (px < 4, px >= wd + 4) are never inside the image: they are zeroed once per
// CTA and only the [prow][wd] interior is refreshed per tile.  The interior is fetched as float4 into REGISTERS one tile
// ahead (all loads of a thread issued back to back: one global-memory latency per tile, hidden behind the previous
// tile's work), split and written to shared memory once the previous tile's operand rows are built.
// Element e = tid + 128 k of the [prow][wd / 4] float4 interior; wd is a power of two.
template <int S> struct StemPf { static constexpr int N = S == 1 ? 2 : 4; };   // max ceil(prow * wd / 512)

__device__ __forceinline__ void patch_zero(bf16* patch, int halves) {
  for (int i = threadIdx.x; i < halves / 8; i += blockDim.x) reinterpret_cast<uint4*>(patch)[i] = make_uint4(0, 0, 0, 0);
}

template <int NPF>
__device__ __forceinline__ void patch_prefetch(float4 (&pre)[NPF], const float* im, int h, int wd, int wshift, int prow,
                                               int iy0) {
#pragma unroll
  for (int k = 0; k < NPF; ++k) {
    const int e = threadIdx.x + 128 * k;
    const int py = e >> (wshift - 2), ix = (e << 2) & (wd - 1);
    const int iy = iy0 + py;
    pre[k] = (py < prow && (unsigned)iy < (unsigned)h)
                 ? __ldg(reinterpret_cast<const float4*>(im + (long long)iy * wd + ix)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int NPF>
__device__ __forceinline__ void patch_commit(bf16* p_hi, bf16* p_lo, const float4 (&pre)[NPF], int wd, int wshift, int pwb,
                                             int prow) {
#pragma unroll
  for (int k = 0; k < NPF; ++k) {
    const int e = threadIdx.x + 128 * k;
    const int py = e >> (wshift - 2), ix = (e << 2) & (wd - 1);
    if (py < prow) {
      const uint32_t h0 = pack_bf16x2(pre[k].x, pre[k].y), h1 = pack_bf16x2(pre[k].z, pre[k].w);
      const uint32_t l0 = pack_bf16x2(pre[k].x - bf16lo(h0), pre[k].y - bf16hi(h0));
      const uint32_t l1 = pack_bf16x2(pre[k].z - bf16lo(h1), pre[k].w - bf16hi(h1));
      const int off = py * pwb + ix + 4;
      *reinterpret_cast<uint2*>(p_hi + off) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(p_lo + off) = make_uint2(l0, l1);
    }
  }
}

// operand rows (hi and lo tile) of output pixel (ty, tx): chunk r = patch elements [tx*S + 1, tx*S + 9) of patch row
// ty*S + r (five aligned words, funnel-shifted by one element when the start is odd); chunk 7 = {one, 0, ...}
template <int S>
__device__ __forceinline__ void build_tap_row(uint8_t* a_hi, uint8_t* a_lo, const bf16* p_hi, const bf16* p_lo, int pwb,
                                              int row, int ty, int tx) {
  const int x0 = tx * S + 1;
  const uint32_t sh = (x0 & 1) * 16;
  const int w0 = (ty * S * pwb + (x0 & ~1)) >> 1;      // pwb is even
  const uint32_t* wh = reinterpret_cast<const uint32_t*>(p_hi) + w0;
  const uint32_t* wl = reinterpret_cast<const uint32_t*>(p_lo) + w0;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const int off = row * 128 + ((r ^ (row & 7)) << 4);
    uint32_t a[5], b[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) { a[j] = wh[r * (pwb >> 1) + j]; b[j] = wl[r * (pwb >> 1) + j]; }
    *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(__funnelshift_r(a[0], a[1], sh), __funnelshift_r(a[1], a[2], sh),
                                                       __funnelshift_r(a[2], a[3], sh), __funnelshift_r(a[3], a[4], sh));
    *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(__funnelshift_r(b[0], b[1], sh), __funnelshift_r(b[1], b[2], sh),
                                                       __funnelshift_r(b[2], b[3], sh), __funnelshift_r(b[3], b[4], sh));
  }
  const int off7 = row * 128 + ((7 ^ (row & 7)) << 4);
  *reinterpret_cast<uint4*>(a_hi + off7) = make_uint4(0x00003F80u, 0, 0, 0);   // k = 56: bf16 1.0 (bias gradient)
  *reinterpret_cast<uint4*>(a_lo + off7) = make_uint4(0, 0, 0, 0);
}

// [64 rows][64 k] bf16 K-major swizzled weight tile(s) from w[co][49] with coalesced reads: row = co, column = k_of_tap
// (forward) or row = tap, column = co (TRANSPOSED, data gradient); everything else stays zero.
template <bool SPLIT, bool TRANSPOSED>
__device__ __forceinline__ void stage_weights(uint8_t* wt, const float* __restrict__ w) {
  for (int i = threadIdx.x; i < (SPLIT ? 1024 : 512); i += blockDim.x) reinterpret_cast<uint4*>(wt)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 49; i += blockDim.x) {
    const int co = i / 49, tap = i - co * 49;
    const float v = __ldg(w + i);
    const int r = TRANSPOSED ? tap : co, c = TRANSPOSED ? co : k_of_tap(tap);
    const int off = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2;
    const bf16 hi = __float2bfloat16_rn(v);
    *reinterpret_cast<bf16*>(wt + off) = hi;
    if (SPLIT) *reinterpret_cast<bf16*>(wt + 8192 + off) = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

// ------------------------------------------------------------------------------------------------ forward
// SPLIT: the weights are split into bf16 hi + lo as well (third MMA pass: taps_hi x w_lo) and the output is written as
// [hi | lo] channel halves (128 channels per pixel) -- the "bf16x3" discriminator path.
// accumulator row -> bias + LeakyReLU -> bf16 (hi [, lo]) -> this thread's row of the swizzled staging tile(s)
template <bool SPLIT>
__device__ __forceinline__ void stem_fwd_epilogue(uint32_t taddr, const float* sbias, float slope, uint8_t* st_hi,
                                                  uint8_t* st_lo, int row) {
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(taddr + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 b0 = *reinterpret_cast<const float4*>(sbias + c0 + 8 * j);
      const float4 b1 = *reinterpret_cast<const float4*>(sbias + c0 + 8 * j + 4);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float f[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float a = __uint_as_float(v[8 * j + k]) + bb[k];
        f[k] = fmaxf(a, a * slope);                  // 0 <= slope < 1
      }
      uint4 u;
      u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
      u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
      const int off = row * 128 + (((c0 / 8 + j) ^ (row & 7)) << 4);
      *reinterpret_cast<uint4*>(st_hi + off) = u;
      if (SPLIT) {
        uint4 l;
        l.x = pack_bf16x2(f[0] - bf16lo(u.x), f[1] - bf16hi(u.x)); l.y = pack_bf16x2(f[2] - bf16lo(u.y), f[3] - bf16hi(u.y));
        l.z = pack_bf16x2(f[4] - bf16lo(u.z), f[5] - bf16hi(u.z)); l.w = pack_bf16x2(f[6] - bf16lo(u.w), f[7] - bf16hi(u.w));
        *reinterpret_cast<uint4*>(st_lo + off) = l;
      }
    }
  }
}

// Per tile and CTA: [patch of tile t+1 -> registers] | build the operand rows of t | MMAs (async) while the patch of
// t+1 goes to shared memory | accumulator -> bias/LeakyReLU -> bf16 rows staged in the (now free) operand tiles | each
// warp stores its 32 pixel rows with one TMA tensor store (two when split).  The chain of one CTA is serial; four CTAs
// per SM interleave.  Before a warp rebuilds its rows it waits until its own store has read them (warp-local).
template <int S, bool SPLIT>
__global__ void __launch_bounds__(128, 4) stem_fwd_tc_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                            const float* __restrict__ bias,
                                                            const __grid_constant__ CUtensorMap tmY, int n, int h, int wd,
                                                            int wshift, float slope) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer: LDS/STS, not generic LD/ST
  uint8_t* a_hi = base;
  uint8_t* a_lo = base + A_BYTES;
  uint8_t* wt = base + 2 * A_BYTES;                   // [64 co][64 taps] bf16, K-major, swizzled: 8 KB (+ 8 KB lo)
  bf16* p_hi = reinterpret_cast<bf16*>(wt + 16384);
  const int wo = wd / S, ho = h / S;
  const int tr = TILE_PX / wo;                        // output rows per tile (1 or 2)
  const int prow = (tr - 1) * S + 7, pwb = wd + 16;
  bf16* p_lo = p_hi + prow * pwb;
  float* sbias = reinterpret_cast<float*>(p_lo + prow * pwb);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sbias + 64);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  constexpr int NPF = StemPf<S>::N;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); tma_prefetch_desc(&tmY); }
  if (warp == 0) { tmem_alloc(tmem_slot, 64); tmem_relinquish(); }
  if (threadIdx.x < 64) sbias[threadIdx.x] = bias[threadIdx.x];
  patch_zero(p_hi, 2 * prow * pwb);
  stage_weights<SPLIT, false>(wt, w);   // (contains a __syncthreads: the zeroed patch is visible before the first commit)

  const int tiles_per_img = ho / tr;
  const int total = tiles_per_img * n;
  const int row = threadIdx.x, ty = row / wo, tx = row - ty * wo;
  float4 pre[NPF];
  int t = blockIdx.x;
  if (t < total) {
    const int im_i = t / tiles_per_img, oy0 = (t - im_i * tiles_per_img) * tr;
    patch_prefetch<NPF>(pre, img + (long long)im_i * h * wd, h, wd, wshift, prow, oy0 * S - 3);
    patch_commit<NPF>(p_hi, p_lo, pre, wd, wshift, pwb, prow);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
  const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);

  for (int it = 0; t < total; t += gridDim.x, ++it) {
    const int tn = t + gridDim.x;
    if (tn < total) {
      const int im_n = tn / tiles_per_img, oy_n = (tn - im_n * tiles_per_img) * tr;
      patch_prefetch<NPF>(pre, img + (long long)im_n * h * wd, h, wd, wshift, prow, oy_n * S - 3);
    }
    if (it > 0) {                // this warp's rows of the operand tiles still feed its previous tensor store
      if (lane == 0) bulk_wait_read();
      __syncwarp();
    }
    build_tap_row<S>(a_hi, a_lo, p_hi, p_lo, pwb, row, ty, tx);
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
    if (tn < total) patch_commit<NPF>(p_hi, p_lo, pre, wd, wshift, pwb, prow);   // every row of this tile is built: patch is free
    mbar_wait(bar, it & 1);      // MMAs retired: accumulator complete, operand tiles free
    tc_fence_after();
    stem_fwd_epilogue<SPLIT>(taddr, sbias, slope, a_hi, a_lo, row);
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      const int pix = t * TILE_PX + warp * 32;       // tiles are consecutive 128-pixel runs of the NHWC output
      tma_store_2d(&tmY, a_hi + warp * 4096, 0, pix);
      if (SPLIT) tma_store_2d(&tmY, a_lo + warp * 4096, 64, pix);
      bulk_commit();
    }
    tc_fence_before();
    __syncthreads();             // next patch visible; accumulator drained by every warp before the next MMAs
  }
  if (lane == 0) bulk_wait_read();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

// ------------------------------------------------------------------------------------------------ wgrad
// D[tap rows: 64 hi + 64 lo][co] accumulated in TMEM over all tiles of this CTA; one red.add epilogue at the end.
// SPLIT: dy comes as [hi | lo] channel halves (two TMA boxes); second MMA pass with the lo tile.
template <int S, bool SPLIT, int NBUF>
__global__ void __launch_bounds__(128) stem_wgrad_tc_kernel(const float* __restrict__ img,
                                                           const __grid_constant__ CUtensorMap tmDy,
                                                           float* __restrict__ dw, float* __restrict__ db, int n, int h,
                                                           int wd, int wshift) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer: LDS/STS, not generic LD/ST
  // two buffers of {taps hi, taps lo, dy}: the MMAs of tile i run while the threads build tile i+1.  The split variant
  // carries a fourth tile (dy lo) and keeps ONE buffer: two would be 132 KB and halve the CTAs per SM (measured slower).
  constexpr int BUF = (SPLIT ? 4 : 3) * A_BYTES;
  const int wo = wd / S, ho = h / S;
  const int tr = TILE_PX / wo;
  const int prow = (tr - 1) * S + 7, pwb = wd + 16;
  bf16* p_hi = reinterpret_cast<bf16*>(base + NBUF * BUF);
  bf16* p_lo = p_hi + prow * pwb;
  uint64_t* full = reinterpret_cast<uint64_t*>(p_lo + prow * pwb);   // dy tile landed (per buffer)
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
  constexpr int NPF = StemPf<S>::N;
  float4 pre[NPF];
  patch_zero(p_hi, 2 * prow * pwb);
  __syncthreads();
  if ((int)blockIdx.x < total) {
    const int t = blockIdx.x, im_i = t / tiles_per_img, oy0 = (t - im_i * tiles_per_img) * tr;
    patch_prefetch<NPF>(pre, img + (long long)im_i * h * wd, h, wd, wshift, prow, oy0 * S - 3);
    patch_commit<NPF>(p_hi, p_lo, pre, wd, wshift, pwb, prow);
  }
  __syncthreads();
  int it = 0;
  for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
    const int b = it % NBUF;
    const uint32_t use = (it / NBUF) & 1;
    uint8_t* buf = base + b * BUF;
    const int im_i = t / tiles_per_img, oy0 = (t - im_i * tiles_per_img) * tr;
    const int tn = t + gridDim.x;
    if (tn < total) {      // next tile's patch -> registers (one global latency, hidden behind this tile's build)
      const int im_n = tn / tiles_per_img, oy_n = (tn - im_n * tiles_per_img) * tr;
      patch_prefetch<NPF>(pre, img + (long long)im_n * h * wd, h, wd, wshift, prow, oy_n * S - 3);
    }
    if (it >= NBUF) mbar_wait(&done[b], use ^ 1);   // MMAs that read this buffer NBUF tiles ago have retired
    if (threadIdx.x == 0) {
      mbar_expect_tx(&full[b], (SPLIT ? 2 : 1) * A_BYTES);
      tma_load_2d(buf + 2 * A_BYTES, &tmDy, &full[b], 0, (im_i * ho + oy0) * wo);
      if (SPLIT) tma_load_2d(buf + 3 * A_BYTES, &tmDy, &full[b], 64, (im_i * ho + oy0) * wo);
    }
    build_tap_row<S>(buf, buf + A_BYTES, p_hi, p_lo, pwb, row, ty, tx);
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
    if (tn < total) patch_commit<NPF>(p_hi, p_lo, pre, wd, wshift, pwb, prow);
    __syncthreads();       // next tile's patch visible
  }
  // drain: the last commit covers every MMA issued before it
  if (it > 0) {
    const int last = it - 1;
    mbar_wait(&done[last % NBUF], (last / NBUF) & 1);
    tc_fence_after();
    const int k = row & 63, tap = (k >> 3) * 7 + (k & 7);     // operand-row order k = 8 r + c (c < 7); k = 56: ones
    const bool is_tap = k < 56 && (k & 7) < 7;
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(taddr + c0, v);
      tmem_ld_wait();
      if (is_tap) {
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dw + (c0 + j) * 49 + tap, __uint_as_float(v[j]));
      } else if (k == 56 && db) {
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
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer: LDS/STS, not generic LD/ST
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
  stage_weights<SPLIT, true>(wt, w);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);

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
  const int prow = (tr - 1) * stride + 7, pwb = wd + 16;
  const int smem = 1024 + 2 * A_BYTES + 16384 + prow * pwb * 4 + 64 * 4 + 64;
  const int total = (ho / tr) * n;
  const int grid = total < 4 * ctx->num_sms ? total : 4 * ctx->num_sms;
  int wshift = 0;
  while ((1 << wshift) < wd) ++wshift;
  if ((1 << wshift) != wd) return lsps_set_error(ctx, LSPS_E_SHAPE, "stem_fwd_tc: image width must be a power of two");
  CUtensorMap tm;   // output as [pixels][channels], stored in boxes of 32 pixels x 64 channels (one per warp)
  uint32_t dims[2] = {split ? 128u : 64u, (uint32_t)((long long)n * ho * wo)}, box[2] = {64u, 32u};
  int rc = lsps_get_tmap(ctx, y, 2, dims, box, &tm);
  if (rc) return rc;
#define LSPS_STEM_FWD(S_, SP_)                                                                                      \
  do {                                                                                                              \
    if ((rc = set_smem(ctx, stem_fwd_tc_kernel<S_, SP_>, smem))) return rc;                                          \
    stem_fwd_tc_kernel<S_, SP_><<<grid, 128, smem, st>>>(img, w, bias, tm, n, h, wd, wshift, slope);       \
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
  const int prow = (tr - 1) * stride + 7, pwb = wd + 16;
  static int nbuf_env = -1;
  if (nbuf_env < 0) { const char* e = getenv("LSPS_STEM_WG_NBUF"); nbuf_env = e ? atoi(e) : 0; }
  const int nbuf = nbuf_env ? nbuf_env : (split ? 1 : 2);
  const int smem = 1024 + nbuf * (split ? 4 : 3) * A_BYTES + prow * pwb * 4 + 128;
  const int total = (ho / tr) * n;
  const int per_sm = 227 * 1024 / (smem + 1024);
  const int grid = total < per_sm * ctx->num_sms ? total : per_sm * ctx->num_sms;
  int wshift = 0;
  while ((1 << wshift) < wd) ++wshift;
  if ((1 << wshift) != wd) return lsps_set_error(ctx, LSPS_E_SHAPE, "stem_wgrad_tc: image width must be a power of two");
  CUtensorMap tm;
  uint32_t dims[2] = {split ? 128u : 64u, (uint32_t)((long long)n * ho * wo)}, box[2] = {64u, (uint32_t)TILE_PX};
  int rc = lsps_get_tmap(ctx, dy, 2, dims, box, &tm);
  if (rc) return rc;
#define LSPS_STEM_WG(S_, SP_)                                                                                       \
  do {                                                                                                              \
    if (nbuf == 1) {                                                                                                \
      if ((rc = set_smem(ctx, stem_wgrad_tc_kernel<S_, SP_, 1>, smem))) return rc;                                   \
      stem_wgrad_tc_kernel<S_, SP_, 1><<<grid, 128, smem, st>>>(img, tm, dw, db, n, h, wd, wshift);                   \
    } else {                                                                                                        \
      if ((rc = set_smem(ctx, stem_wgrad_tc_kernel<S_, SP_, 2>, smem))) return rc;                                   \
      stem_wgrad_tc_kernel<S_, SP_, 2><<<grid, 128, smem, st>>>(img, tm, dw, db, n, h, wd, wshift);                   \
    }                                                                                                               \
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
