// HBM-/latency-bound kernels around the tcgen05 convolutions: Cin=1 stems, decoder head, InstanceNorm,
// noise/KL, losses, discriminator heads, small dense layers (Post head, pose-VAE), Adam, weight repacking.
// None of these is a tensor-core shape (K=49, N=1, reductions, elementwise); they are written for coalesced
// 16-byte accesses over NHWC bf16 data and fp32 accumulation.
//
// Reference op sites: /root/reference/src/trainers/lsps_nets.py:34-83,102-126,186-229 ;
//                     /root/reference/src/trainers/common_net.py:32-40,160-181 ;
//                     /root/reference/src/trainers/lsps_trainer.py:26-34,42-60,107-121,171-192,241-250
#include <stdlib.h>

#include <curand_kernel.h>

#include "common.h"
#include "ptx.cuh"

using namespace lsps;
typedef __nv_bfloat16 bf16;

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum (blockDim.x <= 1024, multiple of 32); result valid in thread 0
__device__ __forceinline__ float block_sum(float v, float* sm) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sm[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = l < (blockDim.x >> 5) ? sm[l] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
// element i of a logical [pixels][sc] tensor stored split as [pixels][hi sc | lo sc] (sc = 0: plain bf16 tensor)
__device__ __forceinline__ float ld_bf16_maybe_split(const bf16* t, long long i, int sc) {
  if (sc == 0) return __bfloat162float(t[i]);
  const long long px = i / sc;
  const bf16* q = t + px * 2 * sc + (i - px * sc);
  return __bfloat162float(q[0]) + __bfloat162float(q[sc]);
}
inline int grid_for(long long n, int block, int cap) {
  long long g = (n + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// =============================================================================================== stems (7x7, Cin=1)
constexpr int ST = 16;  // output tile edge

template <int S>
__global__ void __launch_bounds__(256) stem_fwd_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                      const float* __restrict__ bias, bf16* __restrict__ y, int h,
                                                      int wd, float slope) {
  constexpr int PE = (ST - 1) * S + 7;
  __shared__ float ws[49][64];
  __shared__ float patch[PE][PE + 1];
  const int ho = h / S, wo = wd / S;
  const int n = blockIdx.z, oy0 = blockIdx.y * ST, ox0 = blockIdx.x * ST;
  for (int i = threadIdx.x; i < 49 * 64; i += 256) ws[i % 49][i / 49] = w[i];  // w is [co][tap]
  const float* im = img + (long long)n * h * wd;
  for (int i = threadIdx.x; i < PE * PE; i += 256) {
    const int py = i / PE, px = i % PE;
    const int iy = oy0 * S - 3 + py, ix = ox0 * S - 3 + px;
    patch[py][px] = (iy >= 0 && iy < h && ix >= 0 && ix < wd) ? im[iy * wd + ix] : 0.f;
  }
  __syncthreads();
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[64];
#pragma unroll
  for (int c = 0; c < 64; ++c) acc[c] = 0.f;
  for (int r = 0; r < 7; ++r)
    for (int c = 0; c < 7; ++c) {
      const float v = patch[ty * S + r][tx * S + c];
      const float4* w4 = reinterpret_cast<const float4*>(ws[r * 7 + c]);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 q = w4[j];
        acc[4 * j] += v * q.x; acc[4 * j + 1] += v * q.y; acc[4 * j + 2] += v * q.z; acc[4 * j + 3] += v * q.w;
      }
    }
  const int oy = oy0 + ty, ox = ox0 + tx;
  if (oy < ho && ox < wo) {
    uint4* o = reinterpret_cast<uint4*>(y + (((long long)n * ho + oy) * wo + ox) * 64);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float f[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float t = acc[8 * j + k] + __ldg(bias + 8 * j + k);
        f[k] = t > 0.f ? t : t * slope;
      }
      o[j] = pack8(f);
    }
  }
}

// dW[co][tap] += sum dy[pix][co] * img[pix @ tap]; db[co] += sum dy.  Persistent over 16x16 output tiles.
// thread = (4 consecutive channels, tap set ts: taps ts, ts+16, ts+32, ts+48): one 8-byte dy load + 4 patch loads
// feed 16 FMAs per pixel.
template <int S>
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ img, const bf16* __restrict__ dy,
                                                        float* __restrict__ dw, float* __restrict__ db, int n, int h,
                                                        int wd) {
  constexpr int PE = (ST - 1) * S + 7;
  __shared__ __align__(16) bf16 dys[ST * ST][64];
  __shared__ float patch[PE][PE + 1];
  const int ho = h / S, wo = wd / S;
  const int tiles_x = wo / ST, tiles_y = ho / ST;
  const int total = tiles_x * tiles_y * n;
  const int cq = threadIdx.x & 15, ts = threadIdx.x >> 4;
  int toff[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int tap = ts + 16 * j;
    toff[j] = tap < 49 ? (tap / 7) * (PE + 1) + tap % 7 : -1;
  }
  float acc[4][4], accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[j][k] = 0.f;
  const float* pflat = &patch[0][0];
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, im_i = t / (tiles_x * tiles_y);
    const int oy0 = ty * ST, ox0 = tx * ST;
    __syncthreads();
    const float* im = img + (long long)im_i * h * wd;
    for (int i = threadIdx.x; i < PE * PE; i += 256) {
      const int py = i / PE, px = i % PE;
      const int iy = oy0 * S - 3 + py, ix = ox0 * S - 3 + px;
      patch[py][px] = (iy >= 0 && iy < h && ix >= 0 && ix < wd) ? im[iy * wd + ix] : 0.f;
    }
    for (int i = threadIdx.x; i < ST * ST * 8; i += 256) {
      const int px = i >> 3, j = i & 7;
      const int oy = oy0 + (px >> 4), ox = ox0 + (px & 15);
      reinterpret_cast<uint4*>(dys[px])[j] =
          __ldg(reinterpret_cast<const uint4*>(dy + (((long long)im_i * ho + oy) * wo + ox) * 64) + j);
    }
    __syncthreads();
#pragma unroll 4
    for (int px = 0; px < ST * ST; ++px) {
      const uint2 u = *reinterpret_cast<const uint2*>(&dys[px][cq * 4]);
      const float d[4] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y)};
      const int pbase = (px >> 4) * S * (PE + 1) + (px & 15) * S;
      if (ts == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) accb[k] += d[k];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (toff[j] >= 0) {
          const float v = pflat[pbase + toff[j]];
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[j][k] += d[k] * v;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int tap = ts + 16 * j;
    if (tap < 49) {
#pragma unroll
      for (int k = 0; k < 4; ++k) atomicAdd(dw + (cq * 4 + k) * 49 + tap, acc[j][k]);
    }
  }
  if (ts == 0 && db) {
#pragma unroll
    for (int k = 0; k < 4; ++k) atomicAdd(db + cq * 4 + k, accb[k]);
  }
}

// dimg[y][x] (+)= sum_{r,c,co} dy[(y+3-r)/S][(x+3-c)/S][co] * w[co][r][c]
template <int S>
__global__ void __launch_bounds__(256) stem_dgrad_kernel(const bf16* __restrict__ dy, const float* __restrict__ w,
                                                        float* __restrict__ dimg, int h, int wd, int accumulate) {
  constexpr int DE = S == 1 ? ST + 6 : (ST + 6) / 2 + 1;  // dy patch edge (output pixels touching a 16x16 input tile)
  extern __shared__ __align__(16) uint8_t dsm[];
  float(*ws)[64] = reinterpret_cast<float(*)[64]>(dsm);                // [49][64]
  uint4* dys = reinterpret_cast<uint4*>(dsm + 49 * 64 * sizeof(float));  // [DE*DE][8] chunks, xor-swizzled
  const int ho = h / S, wo = wd / S;
  const int n = blockIdx.z, y0 = blockIdx.y * ST, x0 = blockIdx.x * ST;
  for (int i = threadIdx.x; i < 49 * 64; i += 256) ws[i % 49][i / 49] = w[i];
  // first output row/col that can touch this tile: oy >= ceil((y0 + 3 - 6)/S)
  const int oyb = S == 1 ? y0 - 3 : (y0 - 3 + 1) >> 1, oxb = S == 1 ? x0 - 3 : (x0 - 3 + 1) >> 1;
  for (int i = threadIdx.x; i < DE * DE * 8; i += 256) {
    const int px = i >> 3, j = i & 7;
    const int oy = oyb + px / DE, ox = oxb + px % DE;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (oy >= 0 && oy < ho && ox >= 0 && ox < wo)
      v = __ldg(reinterpret_cast<const uint4*>(dy + (((long long)n * ho + oy) * wo + ox) * 64) + j);
    dys[px * 8 + (j ^ (px & 7))] = v;
  }
  __syncthreads();
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int y = y0 + ty, x = x0 + tx;
  float acc = 0.f;
  for (int r = 0; r < 7; ++r) {
    const int ny = y + 3 - r;
    if (S == 2 && (ny & 1)) continue;
    const int oy = ny / S - oyb;
    if (ny < 0 || ny / S >= ho) continue;
    for (int c = 0; c < 7; ++c) {
      const int nx = x + 3 - c;
      if (S == 2 && (nx & 1)) continue;
      if (nx < 0 || nx / S >= wo) continue;
      const int ox = nx / S - oxb;
      const int px = oy * DE + ox;
      const float4* w4 = reinterpret_cast<const float4*>(ws[r * 7 + c]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float f[8];
        unpack8(dys[px * 8 + (j ^ (px & 7))], f);
        const float4 a = w4[2 * j], b = w4[2 * j + 1];
        acc += f[0] * a.x + f[1] * a.y + f[2] * a.z + f[3] * a.w + f[4] * b.x + f[5] * b.y + f[6] * b.z + f[7] * b.w;
      }
    }
  }
  if (y < h && x < wd) {
    float* o = dimg + ((long long)n * h + y) * wd + x;
    *o = accumulate ? *o + acc : acc;
  }
}

// =============================================================================================== decoder head
// Optional fused reconstruction loss (nn.L1Loss against the input image, lsps_trainer.py:118-121): for pixels
// [t0, t0 + tn) the kernel also accumulates sum |out - target[p - t0]| into acc and writes dout[p - t0] = scale*sign.
__global__ void __launch_bounds__(256) head_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ out,
                                                      long long npix, const float* __restrict__ target, long long t0,
                                                      long long tn, float scale, float* __restrict__ dout,
                                                      float* __restrict__ acc) {
  __shared__ float sm[8];
  const int oct = threadIdx.x & 7;
  float wr[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) wr[k] = __ldg(w + oct * 8 + k);
  const float b = __ldg(bias);
  float l1 = 0.f;
  for (long long p = (long long)blockIdx.x * 32 + (threadIdx.x >> 3); p < npix; p += (long long)gridDim.x * 32) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + p * 64) + oct), f);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += f[k] * wr[k];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (oct == 0) {
      const float o = tanhf(s + b);
      out[p] = o;
      if (target && p >= t0 && p < t0 + tn) {
        const float d = o - __ldg(target + (p - t0));
        l1 += fabsf(d);
        if (dout) dout[p - t0] = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
      }
    }
  }
  if (target) {           // uniform branch: every thread of every block takes it
    const float r = block_sum(l1, sm);
    if (threadIdx.x == 0) atomicAdd(acc, r);
  }
}

__global__ void __launch_bounds__(256) head_bwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ out, const float* __restrict__ dout,
                                                      bf16* __restrict__ dx, float* __restrict__ dw,
                                                      float* __restrict__ db, long long npix, float slope) {
  __shared__ float red[32][64];
  __shared__ float redb[8];
  const int oct = threadIdx.x & 7, pl = threadIdx.x >> 3;
  float wr[8], aw[8], ab = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { wr[k] = __ldg(w + oct * 8 + k); aw[k] = 0.f; }
  for (long long p = (long long)blockIdx.x * 32 + pl; p < npix; p += (long long)gridDim.x * 32) {
    const float o = __ldg(out + p);
    const float dpre = __ldg(dout + p) * (1.f - o * o);
    float f[8], g[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + p * 64) + oct), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      aw[k] += dpre * f[k];
      g[k] = dpre * wr[k] * (f[k] > 0.f ? 1.f : slope);
    }
    reinterpret_cast<uint4*>(dx + p * 64)[oct] = pack8(g);
    if (oct == 0) ab += dpre;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[pl][oct * 8 + k] = aw[k];
  const float bs = block_sum(ab, redb);  // contains __syncthreads
  if (threadIdx.x < 64) {
    float s = 0.f;
    for (int i = 0; i < 32; ++i) s += red[i][threadIdx.x];
    atomicAdd(dw + threadIdx.x, s);
  }
  if (threadIdx.x == 0) atomicAdd(db, bs);
}

// =============================================================================================== InstanceNorm
// one CTA (512 threads) per (image, 64-channel group); thread = (8-channel octet, one of 64 pixel lanes); 8 independent
// 16-byte loads in flight per thread so that each of the three passes is bandwidth- rather than latency-bound
__device__ __forceinline__ void in_reduce64(const float* v8, float (*red)[64], float* outc) {
  const int oct = threadIdx.x & 7, pl = threadIdx.x >> 3;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) red[pl][oct * 8 + k] = v8[k];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll 8
    for (int i = 0; i < 64; ++i) s += red[i][threadIdx.x];
    outc[threadIdx.x] = s;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512) instnorm_fwd_kernel(const bf16* __restrict__ h, const bf16* __restrict__ res,
                                                          bf16* __restrict__ y, float* __restrict__ stats, int hw,
                                                          int c, int mode, float eps, float slope) {
  __shared__ float red[64][64];
  __shared__ float s_mean[64], s_var[64];
  const int n = blockIdx.y, cg = blockIdx.x;
  const int oct = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const long long base = (long long)n * hw * c + cg * 64 + oct * 8;
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 0.f;
#pragma unroll 8
  for (int p = pl; p < hw; p += 64) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(h + base + (long long)p * c)), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += f[k];
  }
  in_reduce64(a, red, s_mean);
  float mean[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { mean[k] = s_mean[oct * 8 + k] / hw; a[k] = 0.f; }
#pragma unroll 8
  for (int p = pl; p < hw; p += 64) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(h + base + (long long)p * c)), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float d = f[k] - mean[k]; a[k] += d * d; }
  }
  in_reduce64(a, red, s_var);
  float rstd[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) rstd[k] = rsqrtf(s_var[oct * 8 + k] / hw + eps);
  if (pl == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float* s = stats + ((long long)n * c + cg * 64 + oct * 8 + k) * 2;
      s[0] = mean[k]; s[1] = rstd[k];
    }
  }
#pragma unroll 8
  for (int p = pl; p < hw; p += 64) {
    float f[8], r[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(h + base + (long long)p * c)), f);
    if (mode == 1) unpack8(__ldg(reinterpret_cast<const uint4*>(res + base + (long long)p * c)), r);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float v = (f[k] - mean[k]) * rstd[k];
      f[k] = mode == 1 ? r[k] + v : (v > 0.f ? v : v * slope);
    }
    *reinterpret_cast<uint4*>(y + base + (long long)p * c) = pack8(f);
  }
}

__global__ void __launch_bounds__(512) instnorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ h,
                                                          const float* __restrict__ stats, bf16* __restrict__ dh,
                                                          int hw, int c, int mode, float slope,
                                                          float* __restrict__ db_, float* __restrict__ db2,
                                                          int nsplit) {
  __shared__ float red[64][64];
  __shared__ float s_a[64], s_b[64];
  const int n = blockIdx.y, cg = blockIdx.x;
  float* db = (nsplit && n >= nsplit) ? db2 : db_;   // images >= nsplit belong to the second conv (grouped res blocks)
  const int oct = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const long long base = (long long)n * hw * c + cg * 64 + oct * 8;
  float mean[8], rstd[8], sg[8], sgx[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float* s = stats + ((long long)n * c + cg * 64 + oct * 8 + k) * 2;
    mean[k] = s[0]; rstd[k] = s[1]; sg[k] = 0.f; sgx[k] = 0.f;
  }
#pragma unroll 8
  for (int p = pl; p < hw; p += 64) {
    float f[8], g[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(h + base + (long long)p * c)), f);
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy + base + (long long)p * c)), g);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (f[k] - mean[k]) * rstd[k];
      const float gg = (mode == 0 && !(xh > 0.f)) ? g[k] * slope : g[k];
      sg[k] += gg; sgx[k] += gg * xh;
    }
  }
  in_reduce64(sg, red, s_a);
  in_reduce64(sgx, red, s_b);
  float mg[8], mgx[8], sdb[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { mg[k] = s_a[oct * 8 + k] / hw; mgx[k] = s_b[oct * 8 + k] / hw; sdb[k] = 0.f; }
#pragma unroll 8
  for (int p = pl; p < hw; p += 64) {
    float f[8], g[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(h + base + (long long)p * c)), f);
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy + base + (long long)p * c)), g);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (f[k] - mean[k]) * rstd[k];
      const float gg = (mode == 0 && !(xh > 0.f)) ? g[k] * slope : g[k];
      f[k] = rstd[k] * (gg - mg[k] - xh * mgx[k]);
    }
    const uint4 u = pack8(f);
    *reinterpret_cast<uint4*>(dh + base + (long long)p * c) = u;
    if (db) {  // bias gradient of the conv that produced h: column sum of the (bf16-rounded) dh -- analytically zero
      float r[8];
      unpack8(u, r);
#pragma unroll
      for (int k = 0; k < 8; ++k) sdb[k] += r[k];
    }
  }
  if (db) {
    in_reduce64(sdb, red, s_a);
    if (threadIdx.x < 64) atomicAdd(db + cg * 64 + threadIdx.x, s_a[threadIdx.x]);
  }
}

// ----- register-resident variants for hw == 1024 (the 32x32 latent of every LeakyINSResBlock) --------------------------
// One CTA (512 threads) per (image, 32-channel group): the whole 64 KB slab of each input lives in the register file
// (8 x 16 B per thread and tensor), so HBM is touched exactly once per element -- read the inputs, write the output --
// and all 8 loads of a thread are in flight together.  Statistics stay two-pass (mean, then centred sum of squares),
// but the second pass runs over registers.  (Forward only: the backward needs two slabs twice and spills; it keeps the
// three-pass kernel above.)
constexpr int INR_PPT = 8;  // pixels per thread: 1024 pixels / 128 pixel lanes

// sum over the 128 pixel lanes of v8 (8 channels of this thread's octet) -> out[32] (all threads see it after return)
__device__ __forceinline__ void inr_reduce(float* v8, float (*red)[32], float* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, oct = threadIdx.x & 3;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float x = v8[k];
    x += __shfl_xor_sync(0xffffffffu, x, 4);
    x += __shfl_xor_sync(0xffffffffu, x, 8);
    x += __shfl_xor_sync(0xffffffffu, x, 16);
    v8[k] = x;
  }
  __syncthreads();
  if (lane < 4) {
#pragma unroll
    for (int k = 0; k < 8; ++k) red[warp][oct * 8 + k] = v8[k];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) s += red[w][threadIdx.x];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(512) instnorm_fwd_reg_kernel(const bf16* __restrict__ h, const bf16* __restrict__ res,
                                                              bf16* __restrict__ y, float* __restrict__ stats, int c,
                                                              float eps, float slope) {
  __shared__ float red[16][32];
  __shared__ float s_a[32], s_b[32];
  constexpr int hw = 128 * INR_PPT;
  const int n = blockIdx.y, cg = blockIdx.x;
  const int oct = threadIdx.x & 3, pl = threadIdx.x >> 2;
  const long long base = (long long)n * hw * c + cg * 32 + oct * 8;
  uint4 xv[INR_PPT], rv[MODE == 1 ? INR_PPT : 1];
#pragma unroll
  for (int i = 0; i < INR_PPT; ++i) xv[i] = __ldg(reinterpret_cast<const uint4*>(h + base + (long long)(pl + 128 * i) * c));
  if (MODE == 1) {
#pragma unroll
    for (int i = 0; i < INR_PPT; ++i) rv[i] = __ldg(reinterpret_cast<const uint4*>(res + base + (long long)(pl + 128 * i) * c));
  }
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 0.f;
#pragma unroll
  for (int i = 0; i < INR_PPT; ++i) {
    float f[8];
    unpack8(xv[i], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += f[k];
  }
  inr_reduce(a, red, s_a);
  float mean[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { mean[k] = s_a[oct * 8 + k] * (1.f / hw); a[k] = 0.f; }
#pragma unroll
  for (int i = 0; i < INR_PPT; ++i) {
    float f[8];
    unpack8(xv[i], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float d = f[k] - mean[k]; a[k] += d * d; }
  }
  inr_reduce(a, red, s_b);
  float rstd[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) rstd[k] = rsqrtf(s_b[oct * 8 + k] * (1.f / hw) + eps);
  if (pl == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float* st = stats + ((long long)n * c + cg * 32 + oct * 8 + k) * 2;
      st[0] = mean[k]; st[1] = rstd[k];
    }
  }
#pragma unroll
  for (int i = 0; i < INR_PPT; ++i) {
    float f[8], r[8];
    unpack8(xv[i], f);
    if (MODE == 1) unpack8(rv[i], r);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float v = (f[k] - mean[k]) * rstd[k];
      f[k] = MODE == 1 ? r[k] + v : (v > 0.f ? v : v * slope);
    }
    *reinterpret_cast<uint4*>(y + base + (long long)(pl + 128 * i) * c) = pack8(f);
  }
}

// ----- statistics-from-the-conv-epilogue path (round 2) --------------------------------------------------------------
// The conv that produces h already accumulated per-(image, channel) sum / sum of squares of its fp32 result
// (LSPS_EP_STATS, igemm.cu), so the norm itself is ONE streaming pass: no reductions, no block-wide syncs, 8 independent
// 16-byte loads in flight per thread.  `sums` rows are [img][2][c]; BatchNorm passes img_stride = 0 and a batch-reduced
// row.  The blocks of pixel-block 0 write (mean, rstd) to stats_out[img][2][c] for the backward pass.
//   MODE 0: y = lrelu(xhat)   MODE 1: y = res + xhat   MODE 2: y = xhat
template <int MODE>
__global__ void __launch_bounds__(256) norm_apply_fwd_kernel(const bf16* __restrict__ h, const bf16* __restrict__ res,
                                                            bf16* __restrict__ y, const float* __restrict__ sums,
                                                            float* __restrict__ stats_out, int hw, int c, int ppb,
                                                            long long img_stride, float inv_count, float eps,
                                                            float slope, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta) {
  const int octs = c >> 3, lanes = 256 / octs;
  const int oct = threadIdx.x % octs, pl = threadIdx.x / octs;
  const int n = blockIdx.y, p0 = blockIdx.x * ppb;
  const float* sr = sums + n * img_stride + oct * 8;
  float mean[8], rstd[8], ga[8], be[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {   // BatchNorm2d(affine=True) weight / bias, or Bias2d alone (common_net.py:92-103,274,297)
    ga[k] = gamma ? __ldg(gamma + oct * 8 + k) : 1.f;
    be[k] = beta ? __ldg(beta + oct * 8 + k) : 0.f;
  }
  {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(sr)), a1 = __ldg(reinterpret_cast<const float4*>(sr) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(sr + c)), b1 = __ldg(reinterpret_cast<const float4*>(sr + c) + 1);
    const float s1[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float s2[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      mean[k] = s1[k] * inv_count;
      const float var = fmaxf(s2[k] * inv_count - mean[k] * mean[k], 0.f);
      rstd[k] = rsqrtf(var + eps);
    }
  }
  if (stats_out && blockIdx.x == 0 && pl == 0) {
    float* so = stats_out + n * img_stride + oct * 8;      // BatchNorm (img_stride 0): every image writes the same row
    reinterpret_cast<float4*>(so)[0] = make_float4(mean[0], mean[1], mean[2], mean[3]);
    reinterpret_cast<float4*>(so)[1] = make_float4(mean[4], mean[5], mean[6], mean[7]);
    reinterpret_cast<float4*>(so + c)[0] = make_float4(rstd[0], rstd[1], rstd[2], rstd[3]);
    reinterpret_cast<float4*>(so + c)[1] = make_float4(rstd[4], rstd[5], rstd[6], rstd[7]);
  }
  const long long base = (long long)n * hw * c + oct * 8;
  const int pend = min(p0 + ppb, hw);
#pragma unroll 4
  for (int p = p0 + pl; p < pend; p += lanes) {
    float f[8], r[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(h + base + (long long)p * c)), f);
    if (MODE == 1) unpack8(__ldg(reinterpret_cast<const uint4*>(res + base + (long long)p * c)), r);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float v = (f[k] - mean[k]) * rstd[k] * ga[k] + be[k];
      f[k] = MODE == 1 ? r[k] + v : (MODE == 0 ? (v > 0.f ? v : v * slope) : v);
    }
    *reinterpret_cast<uint4*>(y + base + (long long)p * c) = pack8(f);
  }
}

// Backward statistics when the producer of the gradient could not take them (gradient arriving at `res + IN(h)`):
// bsums[img][2][c] += (sum g, sum g*xhat) over this block's pixels; g = dy (MODE 1) or dy * lrelu'(xhat) (MODE 0).
template <int MODE>
__global__ void __launch_bounds__(256) norm_bwd_stats_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ h,
                                                            const float* __restrict__ stats, float* __restrict__ bsums,
                                                            int hw, int c, int ppb, long long stat_stride,
                                                            long long bsum_stride, float slope,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta) {
  extern __shared__ float nred[];   // [lanes][c] x 2
  const int octs = c >> 3, lanes = 256 / octs;
  const int oct = threadIdx.x % octs, pl = threadIdx.x / octs;
  const int n = blockIdx.y, p0 = blockIdx.x * ppb;
  const float* st = stats + n * stat_stride + oct * 8;
  float mean[8], rstd[8], sg[8], sgx[8];
  {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(st)), a1 = __ldg(reinterpret_cast<const float4*>(st) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(st + c)), b1 = __ldg(reinterpret_cast<const float4*>(st + c) + 1);
    mean[0] = a0.x; mean[1] = a0.y; mean[2] = a0.z; mean[3] = a0.w; mean[4] = a1.x; mean[5] = a1.y; mean[6] = a1.z; mean[7] = a1.w;
    rstd[0] = b0.x; rstd[1] = b0.y; rstd[2] = b0.z; rstd[3] = b0.w; rstd[4] = b1.x; rstd[5] = b1.y; rstd[6] = b1.z; rstd[7] = b1.w;
  }
  float ga[8], be[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sg[k] = 0.f; sgx[k] = 0.f;
    ga[k] = gamma ? __ldg(gamma + oct * 8 + k) : 1.f;
    be[k] = beta ? __ldg(beta + oct * 8 + k) : 0.f;
  }
  const long long base = (long long)n * hw * c + oct * 8;
  const int pend = min(p0 + ppb, hw);
#pragma unroll 4
  for (int p = p0 + pl; p < pend; p += lanes) {
    float f[8], g[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(h + base + (long long)p * c)), f);
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy + base + (long long)p * c)), g);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (f[k] - mean[k]) * rstd[k];
      const float gg = (MODE == 0 && !(xh * ga[k] + be[k] > 0.f)) ? g[k] * slope : g[k];   // lrelu'(gamma*xhat + beta)
      sg[k] += gg; sgx[k] += gg * xh;          // = d beta, d gamma of an affine norm
    }
  }
  float* r1 = nred;
  float* r2 = nred + lanes * c;
#pragma unroll
  for (int k = 0; k < 8; ++k) { r1[pl * c + oct * 8 + k] = sg[k]; r2[pl * c + oct * 8 + k] = sgx[k]; }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += 256) {
    float a = 0.f, b = 0.f;
    for (int l = 0; l < lanes; ++l) { a += r1[l * c + ch]; b += r2[l * c + ch]; }
    atomicAdd(bsums + n * bsum_stride + ch, a);
    atomicAdd(bsums + n * bsum_stride + c + ch, b);
  }
}

// dh = rstd * (g - mean(g) - xhat * mean(g*xhat)), the means from bsums (taken by the dgrad epilogue, LSPS_EP_INBWD, or
// by norm_bwd_stats_kernel).  GMODE 0: `g` is the raw gradient w.r.t. lrelu(xhat) (mask applied here); 1: g is used as is
// (the gradient w.r.t. res + xhat); 2: g was masked by the dgrad epilogue and `h` holds a = lrelu(xhat), from which
// xhat is recovered.
template <int GMODE>
__global__ void __launch_bounds__(256) norm_bwd_apply_kernel(const bf16* __restrict__ g_, const bf16* __restrict__ h,
                                                            const float* __restrict__ stats,
                                                            const float* __restrict__ bsums, bf16* __restrict__ dh,
                                                            int hw, int c, int ppb, long long stat_stride,
                                                            long long bsum_stride, float inv_count, float slope,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta) {
  const int octs = c >> 3, lanes = 256 / octs;
  const int oct = threadIdx.x % octs, pl = threadIdx.x / octs;
  const int n = blockIdx.y, p0 = blockIdx.x * ppb;
  const float* st = stats + n * stat_stride + oct * 8;
  const float* bs = bsums + n * bsum_stride + oct * 8;
  const float inv_slope = 1.f / slope;
  float mean[8], rstd[8], mg[8], mgx[8], ga[8], be[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    ga[k] = gamma ? __ldg(gamma + oct * 8 + k) : 1.f;
    be[k] = beta ? __ldg(beta + oct * 8 + k) : 0.f;
  }
  {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(st)), a1 = __ldg(reinterpret_cast<const float4*>(st) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(st + c)), b1 = __ldg(reinterpret_cast<const float4*>(st + c) + 1);
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(bs)), c1 = __ldg(reinterpret_cast<const float4*>(bs) + 1);
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(bs + c)), d1 = __ldg(reinterpret_cast<const float4*>(bs + c) + 1);
    mean[0] = a0.x; mean[1] = a0.y; mean[2] = a0.z; mean[3] = a0.w; mean[4] = a1.x; mean[5] = a1.y; mean[6] = a1.z; mean[7] = a1.w;
    rstd[0] = b0.x; rstd[1] = b0.y; rstd[2] = b0.z; rstd[3] = b0.w; rstd[4] = b1.x; rstd[5] = b1.y; rstd[6] = b1.z; rstd[7] = b1.w;
    mg[0] = c0.x; mg[1] = c0.y; mg[2] = c0.z; mg[3] = c0.w; mg[4] = c1.x; mg[5] = c1.y; mg[6] = c1.z; mg[7] = c1.w;
    mgx[0] = d0.x; mgx[1] = d0.y; mgx[2] = d0.z; mgx[3] = d0.w; mgx[4] = d1.x; mgx[5] = d1.y; mgx[6] = d1.z; mgx[7] = d1.w;
#pragma unroll
    for (int k = 0; k < 8; ++k) { mg[k] *= inv_count; mgx[k] *= inv_count; }
  }
  const long long base = (long long)n * hw * c + oct * 8;
  const int pend = min(p0 + ppb, hw);
#pragma unroll 4
  for (int p = p0 + pl; p < pend; p += lanes) {
    float f[8], g[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(h + base + (long long)p * c)), f);
    unpack8(__ldg(reinterpret_cast<const uint4*>(g_ + base + (long long)p * c)), g);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = GMODE == 2 ? (f[k] > 0.f ? f[k] : f[k] * inv_slope) : (f[k] - mean[k]) * rstd[k];
      const float gg = (GMODE == 0 && !(xh * ga[k] + be[k] > 0.f)) ? g[k] * slope : g[k];
      f[k] = rstd[k] * ga[k] * (gg - mg[k] - xh * mgx[k]);
    }
    *reinterpret_cast<uint4*>(dh + base + (long long)p * c) = pack8(f);
  }
}

// BatchNorm2d running statistics (momentum update with the unbiased variance, as torch) from a batch row of sums,
// and the inverse: a sums row that makes norm_apply_fwd normalise with the running statistics (eval mode).
__global__ void bn_running_update_kernel(const float* __restrict__ sums, float* __restrict__ rmean,
                                         float* __restrict__ rvar, int c, float count, float momentum) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= c) return;
  const float mean = sums[j] / count;
  const float var = fmaxf(sums[c + j] / count - mean * mean, 0.f);
  rmean[j] = (1.f - momentum) * rmean[j] + momentum * mean;
  rvar[j] = (1.f - momentum) * rvar[j] + momentum * var * (count > 1.f ? count / (count - 1.f) : 1.f);
}
__global__ void bn_running_to_sums_kernel(const float* __restrict__ rmean, const float* __restrict__ rvar,
                                          float* __restrict__ sums, int c, float count) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= c) return;
  sums[j] = rmean[j] * count;
  sums[c + j] = (rvar[j] + rmean[j] * rmean[j]) * count;
}

// BatchNorm: per-channel batch statistics = the per-image rows of `sums` added up (tiny: n x 2c floats).
// out[0][c] = sum over images of sums[img][0][c], out[1][c] likewise
__global__ void __launch_bounds__(256) norm_reduce_images_kernel(const float* __restrict__ sums, float* __restrict__ out,
                                                                int n, int c2) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= c2) return;
  float s = 0.f;
  for (int i = 0; i < n; ++i) s += sums[(long long)i * c2 + j];
  out[j] = s;
}

// =============================================================================================== elementwise / losses
__global__ void __launch_bounds__(256) noise_kl_kernel(const bf16* __restrict__ x, const float* __restrict__ noise,
                                                      bf16* __restrict__ z, float* __restrict__ acc, long long n8) {
  __shared__ float sm[8];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n8; i += (long long)gridDim.x * 256) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), f);
    const float4 a = __ldg(reinterpret_cast<const float4*>(noise) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(noise) + 2 * i + 1);
    f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w; f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += f[k] * f[k];
    reinterpret_cast<uint4*>(z)[i] = pack8(f);
  }
  const float t = block_sum(s, sm);
  if (threadIdx.x == 0) atomicAdd(acc, t);
}

// Device-RNG form of the GaussianNoiseLayer (the timed mode): z = x + N(0,1) drawn in the kernel from Philox4x32-10 --
// counter = (seed, subsequence = 16-byte chunk index, offset), two normal4 draws per chunk -- so no noise tensor is ever
// written to or read from HBM (the host-RNG parity mode keeps noise_kl_kernel above, which consumes the reference's draws).
__global__ void __launch_bounds__(256) noise_kl_philox_kernel(const bf16* __restrict__ x, bf16* __restrict__ z,
                                                             float* __restrict__ acc, long long n8,
                                                             unsigned long long seed, unsigned long long offset) {
  __shared__ float sm[8];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n8; i += (long long)gridDim.x * 256) {
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)i, offset, &st);
    const float4 a = curand_normal4(&st), b = curand_normal4(&st);
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), f);
    f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w; f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += f[k] * f[k];
    reinterpret_cast<uint4*>(z)[i] = pack8(f);
  }
  const float t = block_sum(s, sm);
  if (threadIdx.x == 0) atomicAdd(acc, t);
}

__global__ void __launch_bounds__(256) axpy_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b,
                                                       float alpha, bf16* __restrict__ out, long long n8) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n8; i += (long long)gridDim.x * 256) {
    float f[8], g[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(b) + i), g);
    if (a) unpack8(__ldg(reinterpret_cast<const uint4*>(a) + i), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = (a ? f[k] : 0.f) + alpha * g[k];
    reinterpret_cast<uint4*>(out)[i] = pack8(f);
  }
}

// latent matching term of the Mapping net (lsps_trainer.py:52-53,97): acc += sum (a-b)^2 ; g = scale*(a-b)
__global__ void __launch_bounds__(256) l2_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b,
                                                     bf16* __restrict__ g, float scale, float* __restrict__ acc,
                                                     long long n8) {
  __shared__ float sm[8];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n8; i += (long long)gridDim.x * 256) {
    float f[8], h[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(a) + i), f);
    unpack8(__ldg(reinterpret_cast<const uint4*>(b) + i), h);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float d = f[k] - h[k];
      s += d * d;
      f[k] = scale * d;
    }
    reinterpret_cast<uint4*>(g)[i] = pack8(f);
  }
  const float t = block_sum(s, sm);
  if (threadIdx.x == 0) atomicAdd(acc, t);
}

__global__ void __launch_bounds__(256) l1_f32_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                    float* __restrict__ dx, float scale, int accumulate,
                                                    float* __restrict__ acc, long long n) {
  __shared__ float sm[8];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float d = x[i] - t[i];
    s += fabsf(d);
    if (dx) {
      const float g = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
      dx[i] = accumulate ? dx[i] + g : g;
    }
  }
  const float r = block_sum(s, sm);
  if (threadIdx.x == 0) atomicAdd(acc, r);
}

__global__ void __launch_bounds__(256) l1_feat_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b,
                                                     float* __restrict__ da, float* __restrict__ db, float scale,
                                                     float* __restrict__ acc, long long n, int sc) {
  __shared__ float sm[8];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float d = ld_bf16_maybe_split(a, i, sc) - ld_bf16_maybe_split(b, i, sc);
    s += fabsf(d);
    const float g = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
    if (da) da[i] += g;
    if (db) db[i] -= g;
  }
  const float r = block_sum(s, sm);
  if (threadIdx.x == 0) atomicAdd(acc, r);
}

// one warp per row
__global__ void __launch_bounds__(256) dhead_fwd_kernel(const bf16* __restrict__ f, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ logits,
                                                       long long rows, int c, int split) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const long long pitch = split ? 2LL * c : c;
  float s = 0.f;
  for (int j = lane; j < c / 8; j += 32) {
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(f + r * pitch) + j), v);
    if (split) {
      float l[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(f + r * pitch + c) + j), l);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += l[k];
    }
    const float4 a = __ldg(reinterpret_cast<const float4*>(w) + 2 * j), b = __ldg(reinterpret_cast<const float4*>(w) + 2 * j + 1);
    s += v[0] * a.x + v[1] * a.y + v[2] * a.z + v[3] * a.w + v[4] * b.x + v[5] * b.y + v[6] * b.z + v[7] * b.w;
  }
  s = warp_sum(s);
  if (lane == 0) logits[r] = s + __ldg(bias);
}

__global__ void __launch_bounds__(256) bce_logits_kernel(const float* __restrict__ logits, float target, float scale,
                                                        float* __restrict__ dlogits, float* __restrict__ acc,
                                                        long long rows) {
  __shared__ float sm[8];
  float s = 0.f, cnt = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < rows; i += (long long)gridDim.x * 256) {
    const float z = logits[i];
    const float p = 1.f / (1.f + expf(-z));
    const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(1.f - p), -100.f);
    s += -(target * lp + (1.f - target) * lq);
    cnt += target > 0.5f ? (p >= 0.5f ? 1.f : 0.f) : (p <= 0.5f ? 1.f : 0.f);
    if (dlogits) dlogits[i] = scale * (p - target);
  }
  const float r = block_sum(s, sm);
  const float c2 = block_sum(cnt, sm);
  if (threadIdx.x == 0) { atomicAdd(acc, r); atomicAdd(acc + 1, c2); }
}

// D head, sigmoid, binary cross-entropy and the backward of all three in ONE kernel (the GAN BCE fused into the head it
// follows).  One warp per row of f (a 2x2-map pixel of an image): logit = f.w + b, p = sigmoid(logit), BCE vs the
// constant target of the row's group (torch semantics: log clamped at -100), dlogit = scale*(p - t); then
// df[r,:] = dlogit*w (written, so df needs no zero fill), dw += dlogit*f[r,:], db += dlogit.  Rows are grouped by image
// group (real / generated / decoded batches): group g = r / rows_per_group has target tgt[g] (< 0: the group takes no
// part: its df rows are zero) and adds (loss sum, #correct) to acc[slot[g]], acc[slot[g] + 1].
struct DheadGroups { float tgt[8]; int slot[8]; };
__global__ void __launch_bounds__(256) dhead_bce_kernel(const bf16* __restrict__ f, const float* __restrict__ w,
                                                       const float* __restrict__ bias, long long rows, int c, int split,
                                                       long long rows_per_group, DheadGroups G, float scale,
                                                       float* __restrict__ logits, float* __restrict__ df,
                                                       float* __restrict__ dw, float* __restrict__ db,
                                                       float* __restrict__ acc, int rows_per_warp) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long r0 = wid * rows_per_warp;
  const long long pitch = split ? 2LL * c : c;
  const int noct = c / 8;                 // <= 256: at most 8 octets per lane
  float aw[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) aw[i][k] = 0.f;
  float ab = 0.f;
  const float b0 = __ldg(bias);
  for (long long r = r0; r < r0 + rows_per_warp && r < rows; ++r) {
    float fv[8][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = lane + 32 * i;
      if (j < noct) {
        unpack8(__ldg(reinterpret_cast<const uint4*>(f + r * pitch) + j), fv[i]);
        if (split) {
          float l[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(f + r * pitch + c) + j), l);
#pragma unroll
          for (int k = 0; k < 8; ++k) fv[i][k] += l[k];
        }
        const float4 a = __ldg(reinterpret_cast<const float4*>(w) + 2 * j), b = __ldg(reinterpret_cast<const float4*>(w) + 2 * j + 1);
        s += fv[i][0] * a.x + fv[i][1] * a.y + fv[i][2] * a.z + fv[i][3] * a.w + fv[i][4] * b.x + fv[i][5] * b.y + fv[i][6] * b.z + fv[i][7] * b.w;
      }
    }
    s = warp_sum(s);
    const float z = s + b0;
    const int g = (int)(r / rows_per_group);
    const float t = G.tgt[g < 8 ? g : 7];
    float dl = 0.f;
    if (t >= 0.f) {
      const float p = 1.f / (1.f + expf(-z));
      dl = scale * (p - t);
      if (lane == 0) {
        const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(1.f - p), -100.f);
        atomicAdd(acc + G.slot[g], -(t * lp + (1.f - t) * lq));
        atomicAdd(acc + G.slot[g] + 1, t > 0.5f ? (p >= 0.5f ? 1.f : 0.f) : (p <= 0.5f ? 1.f : 0.f));
      }
    }
    if (lane == 0 && logits) logits[r] = z;
    ab += dl;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = lane + 32 * i;
      if (j < noct) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(w) + 2 * j), b = __ldg(reinterpret_cast<const float4*>(w) + 2 * j + 1);
        if (df) {
          float4* o = reinterpret_cast<float4*>(df + r * c) + 2 * j;
          o[0] = make_float4(dl * a.x, dl * a.y, dl * a.z, dl * a.w);
          o[1] = make_float4(dl * b.x, dl * b.y, dl * b.z, dl * b.w);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) aw[i][k] += dl * fv[i][k];
      }
    }
  }
  if (dw) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = lane + 32 * i;
      if (j < noct) {
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(dw + 8 * j + k, aw[i][k]);
      }
    }
  }
  if (db && lane == 0 && r0 < rows) atomicAdd(db, ab);
}

// thread per column, block.y = row chunk
__global__ void __launch_bounds__(256) dhead_bwd_kernel(const bf16* __restrict__ f, const float* __restrict__ w,
                                                       const float* __restrict__ dl, float* __restrict__ df,
                                                       float* __restrict__ dw, float* __restrict__ db, long long rows,
                                                       int c, int rows_per_block, int split) {
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col >= c) return;
  const long long pitch = split ? 2LL * c : c;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  const float wc = __ldg(w + col);
  float aw = 0.f, ab = 0.f;
  for (long long r = r0; r < r1; ++r) {
    const float d = __ldg(dl + r);
    if (df) df[r * c + col] += d * wc;
    aw += d * (__bfloat162float(f[r * pitch + col]) + (split ? __bfloat162float(f[r * pitch + c + col]) : 0.f));
    ab += d;
  }
  if (dw) atomicAdd(dw + col, aw);
  if (db && col == 0) atomicAdd(db, ab);
}

__global__ void __launch_bounds__(256) mask_to_bf16_kernel(const float* __restrict__ df, const bf16* __restrict__ f,
                                                          bf16* __restrict__ out, float slope, long long n, int sc) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    if (sc == 0) {
      const float v = __bfloat162float(f[i]);
      out[i] = __float2bfloat16(df[i] * (v > 0.f ? 1.f : slope));
    } else {   // split tensors: the hi half carries the sign; the result is written as hi | lo
      const long long px = i / sc, o = px * 2 * sc + (i - px * sc);
      const float v = __bfloat162float(f[o]);
      const float g = df[i] * (v > 0.f ? 1.f : slope);
      const bf16 hi = __float2bfloat16(g);
      out[o] = hi;
      out[o + sc] = __float2bfloat16(g - __bfloat162float(hi));
    }
  }
}

__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ dy, long long rows, int c,
                                                    float* __restrict__ db, int row_lanes, int fold) {
  // block = (octets of 8 channels) x (row lanes); grid.x strides over rows, grid.y over 2048-channel slabs
  __shared__ float red[256 * 8];
  const int octs_blk = 256 / row_lanes;
  const int octs = c / 8;
  const int ol = threadIdx.x % octs_blk, rl = threadIdx.x / octs_blk;
  const int oct = blockIdx.y * octs_blk + ol;
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 0.f;
  if (oct < octs) {
    for (long long r = (long long)blockIdx.x * row_lanes + rl; r < rows; r += (long long)gridDim.x * row_lanes) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(dy + r * c) + oct), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] += f[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[(rl * octs_blk + ol) * 8 + k] = a[k];
  __syncthreads();
  const int nch = octs_blk * 8;  // channels handled by this block
  for (int ch = threadIdx.x; ch < nch; ch += 256) {
    float s = 0.f;
    for (int j = 0; j < row_lanes; ++j) s += red[j * nch + ch];
    const int gc = blockIdx.y * nch + ch;
    if (gc < c) atomicAdd(db + ((fold && gc >= fold) ? gc - fold : gc), s);   // fold: columns [fold, 2 fold) are lo halves
  }
}

// =============================================================================================== small dense layers
// XBF16: x is bf16; sc > 0 additionally says x rows are split tensors with sc channels per pixel half
template <bool XBF16>
__global__ void __launch_bounds__(256) linear_fwd_kernel(const void* __restrict__ x_, const float* __restrict__ w,
                                                        const float* __restrict__ b, float* __restrict__ y, int m,
                                                        int n, int k, int act, float slope, int sc) {
  const int lane = threadIdx.x & 31;
  const long long o = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (o >= (long long)m * n) return;
  const int i = (int)(o / n), j = (int)(o % n);
  float s = 0.f;
  for (int q = lane; q < k; q += 32) {
    const float xv = XBF16 ? ld_bf16_maybe_split(static_cast<const bf16*>(x_), (long long)i * k + q, sc)
                           : static_cast<const float*>(x_)[(long long)i * k + q];
    s += xv * __ldg(w + (long long)j * k + q);
  }
  s = warp_sum(s);
  if (lane == 0) {
    s += b ? __ldg(b + j) : 0.f;
    if (act == 1) s = s > 0.f ? s : s * slope;
    else if (act == 2) s = s > 20.f ? s : log1pf(expf(s));
    y[o] = s;
  }
}

template <bool XBF16>
__global__ void __launch_bounds__(256) linear_bwd_dw_kernel(const void* __restrict__ x_, const float* __restrict__ dy,
                                                           float* __restrict__ dw, int m, int n, int k, int sc) {
  const long long o = (long long)blockIdx.x * 256 + threadIdx.x;
  if (o >= (long long)n * k) return;
  const int j = (int)(o / k), q = (int)(o % k);
  float s = 0.f;
  for (int i = 0; i < m; ++i) {
    const float xv = XBF16 ? ld_bf16_maybe_split(static_cast<const bf16*>(x_), (long long)i * k + q, sc)
                           : static_cast<const float*>(x_)[(long long)i * k + q];
    s += __ldg(dy + (long long)i * n + j) * xv;
  }
  dw[o] += s;
}
__global__ void __launch_bounds__(256) linear_bwd_dx_kernel(const float* __restrict__ w, const float* __restrict__ dy,
                                                           float* __restrict__ dx, int accumulate, int m, int n, int k) {
  const long long o = (long long)blockIdx.x * 256 + threadIdx.x;
  if (o >= (long long)m * k) return;
  const int i = (int)(o / k), q = (int)(o % k);
  float s = 0.f;
  for (int j = 0; j < n; ++j) s += __ldg(dy + (long long)i * n + j) * __ldg(w + (long long)j * k + q);
  dx[o] = accumulate ? dx[o] + s : s;
}
__global__ void linear_bwd_db_kernel(const float* __restrict__ dy, float* __restrict__ db, int m, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float s = 0.f;
  for (int i = 0; i < m; ++i) s += dy[(long long)i * n + j];
  db[j] += s;
}

__global__ void act_bwd_kernel(float* __restrict__ dy, const float* __restrict__ y, int act, float slope, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = y[i];
  if (act == 1) dy[i] *= v > 0.f ? 1.f : slope;
  else if (act == 2) dy[i] *= v > 20.f ? 1.f : 1.f - expf(-v);  // softplus' = sigmoid(pre) = 1 - exp(-softplus)
}

__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ p, const float* __restrict__ e,
                                                 float* __restrict__ dp, float scale, float* __restrict__ acc,
                                                 long long n) {
  __shared__ float sm[8];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float d = p[i] - e[i];
    s += d * d;
    if (dp) dp[i] = scale * d;
  }
  const float r = block_sum(s, sm);
  if (threadIdx.x == 0) atomicAdd(acc, r);
}

__global__ void __launch_bounds__(256) vae_reparam_kernel(const float* __restrict__ mu, const float* __restrict__ sd,
                                                         const float* __restrict__ noise, float* __restrict__ z,
                                                         float* __restrict__ acc, long long n) {
  __shared__ float sm[8];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float m = mu[i], d = sd[i];
    z[i] = m + d * noise[i];
    s += m * m + d * d - logf(d * d);
  }
  const float r = block_sum(s, sm);
  if (threadIdx.x == 0 && acc) atomicAdd(acc, r);
}
__global__ void vae_reparam_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ sd,
                                       const float* __restrict__ noise, const float* __restrict__ dz,
                                       float* __restrict__ dmu, float* __restrict__ dsd, float kl_scale, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float d = sd[i];
  dmu[i] = dz[i] + kl_scale * 2.f * mu[i];
  dsd[i] = dz[i] * noise[i] + kl_scale * (2.f * d - 2.f / d);
}

// =============================================================================================== fused pose-VAE step (K11)
// forward + losses + backward of poseVAE (lsps_nets.py:34-83, lsps_trainer.py:59-74) for a block of <= 16 rows per CTA:
// every weight (13.8 K floats at 108/50/20) and every per-row intermediate lives in shared memory; the parameter
// gradients of the block are added to the (zeroed) gradient buffer with atomicAdd.  Replaces 24 launches of the small
// dense kernels by one.
//   y -> h = lrelu(W1 y + b1) -> mu = Wmu h + bmu, sd = softplus(Wsg h + bsg) -> z = mu + sd * noise
//     -> hd = lrelu(Wd1 z + bd1) -> dec = Wd2 hd + bd2 ;  loss = ll_scale * sum|dec - y| + kl_scale * sum(mu^2 + sd^2 - log sd^2)
struct VaeStepParams {
  const float* y; const float* noise;
  const float* w[10];   // W1 b1 Wmu bmu Wsg bsg Wd1 bd1 Wd2 bd2
  float* g[10];         // their gradient accumulators
  float* dec;           // [rows][d]
  float* acc;           // acc[0] += KL sum, acc[1] += L1 sum
  int rows, d, h, z;
  float ll_scale, kl_scale, slope;
};
constexpr int VAE_RPC = 16;   // rows per CTA

__global__ void __launch_bounds__(256) vae_step_kernel(const __grid_constant__ VaeStepParams p) {
  extern __shared__ float vs[];
  const int d = p.d, h = p.h, z = p.z, R = VAE_RPC;
  float* W1 = vs;            float* B1 = W1 + h * d;
  float* Wmu = B1 + h;       float* Bmu = Wmu + z * h;
  float* Wsg = Bmu + z;      float* Bsg = Wsg + z * h;
  float* Wd1 = Bsg + z;      float* Bd1 = Wd1 + h * z;
  float* Wd2 = Bd1 + h;      float* Bd2 = Wd2 + d * h;
  float* Y = Bd2 + d;        float* H = Y + R * d;
  float* MU = H + R * h;     float* SD = MU + R * z;
  float* NZ = SD + R * z;    float* Z = NZ + R * z;
  float* HD = Z + R * z;     float* DD = HD + R * h;
  float* DHD = DD + R * d;   float* DMU = DHD + R * h;
  float* DSD = DMU + R * z;  float* DH = DSD + R * z;
  __shared__ float red[8];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int r0 = blockIdx.x * R;
  const int nr = p.rows - r0 < R ? p.rows - r0 : R;
  {
    float* dst[10] = {W1, B1, Wmu, Bmu, Wsg, Bsg, Wd1, Bd1, Wd2, Bd2};
    const int cnt[10] = {h * d, h, z * h, z, z * h, z, h * z, h, d * h, d};
#pragma unroll
    for (int a = 0; a < 10; ++a)
      for (int i = tid; i < cnt[a]; i += nt) dst[a][i] = __ldg(p.w[a] + i);
  }
  for (int i = tid; i < nr * d; i += nt) Y[i] = p.y[(long long)r0 * d + i];
  for (int i = tid; i < nr * z; i += nt) NZ[i] = p.noise[(long long)r0 * z + i];
  __syncthreads();
  // ---- forward
  for (int o = tid; o < nr * h; o += nt) {
    const int r = o / h, j = o - r * h;
    float s = B1[j];
    for (int i = 0; i < d; ++i) s += W1[j * d + i] * Y[r * d + i];
    H[o] = s > 0.f ? s : s * p.slope;
  }
  __syncthreads();
  float kl = 0.f;
  for (int o = tid; o < nr * z; o += nt) {
    const int r = o / z, k = o - r * z;
    float m = Bmu[k], q = Bsg[k];
    for (int j = 0; j < h; ++j) { const float hv = H[r * h + j]; m += Wmu[k * h + j] * hv; q += Wsg[k * h + j] * hv; }
    const float sd = q > 20.f ? q : log1pf(expf(q));
    MU[o] = m; SD[o] = sd; Z[o] = m + sd * NZ[o];
    kl += m * m + sd * sd - logf(sd * sd);
  }
  __syncthreads();
  for (int o = tid; o < nr * h; o += nt) {
    const int r = o / h, j = o - r * h;
    float s = Bd1[j];
    for (int k = 0; k < z; ++k) s += Wd1[j * z + k] * Z[r * z + k];
    HD[o] = s > 0.f ? s : s * p.slope;
  }
  __syncthreads();
  float l1 = 0.f;
  for (int o = tid; o < nr * d; o += nt) {
    const int r = o / d, i = o - r * d;
    float s = Bd2[i];
    for (int j = 0; j < h; ++j) s += Wd2[i * h + j] * HD[r * h + j];
    p.dec[(long long)r0 * d + o] = s;
    const float df = s - Y[o];
    l1 += fabsf(df);
    DD[o] = df > 0.f ? p.ll_scale : (df < 0.f ? -p.ll_scale : 0.f);
  }
  kl = block_sum(kl, red);
  if (tid == 0) atomicAdd(p.acc, kl);
  l1 = block_sum(l1, red);               // (block_sum ends with the values visible to every thread)
  if (tid == 0) atomicAdd(p.acc + 1, l1);
  __syncthreads();
  // ---- backward: de_fc2
  for (int o = tid; o < d * h; o += nt) {
    const int i = o / h, j = o - i * h;
    float s = 0.f;
    for (int r = 0; r < nr; ++r) s += DD[r * d + i] * HD[r * h + j];
    atomicAdd(p.g[8] + o, s);
  }
  for (int i = tid; i < d; i += nt) {
    float s = 0.f;
    for (int r = 0; r < nr; ++r) s += DD[r * d + i];
    atomicAdd(p.g[9] + i, s);
  }
  for (int o = tid; o < nr * h; o += nt) {
    const int r = o / h, j = o - r * h;
    float s = 0.f;
    for (int i = 0; i < d; ++i) s += Wd2[i * h + j] * DD[r * d + i];
    DHD[o] = HD[o] > 0.f ? s : s * p.slope;
  }
  __syncthreads();
  // ---- de_fc1
  for (int o = tid; o < h * z; o += nt) {
    const int j = o / z, k = o - j * z;
    float s = 0.f;
    for (int r = 0; r < nr; ++r) s += DHD[r * h + j] * Z[r * z + k];
    atomicAdd(p.g[6] + o, s);
  }
  for (int j = tid; j < h; j += nt) {
    float s = 0.f;
    for (int r = 0; r < nr; ++r) s += DHD[r * h + j];
    atomicAdd(p.g[7] + j, s);
  }
  for (int o = tid; o < nr * z; o += nt) {
    const int r = o / z, k = o - r * z;
    float dz = 0.f;
    for (int j = 0; j < h; ++j) dz += Wd1[j * z + k] * DHD[r * h + j];
    const float sd = SD[o];
    DMU[o] = dz + p.kl_scale * 2.f * MU[o];
    const float dsd = dz * NZ[o] + p.kl_scale * (2.f * sd - 2.f / sd);
    DSD[o] = dsd * (sd > 20.f ? 1.f : 1.f - expf(-sd));          // softplus' = 1 - exp(-softplus)
  }
  __syncthreads();
  // ---- en_mu / en_sigma
  for (int o = tid; o < z * h; o += nt) {
    const int k = o / h, j = o - k * h;
    float a = 0.f, b = 0.f;
    for (int r = 0; r < nr; ++r) { const float hv = H[r * h + j]; a += DMU[r * z + k] * hv; b += DSD[r * z + k] * hv; }
    atomicAdd(p.g[2] + o, a);
    atomicAdd(p.g[4] + o, b);
  }
  for (int k = tid; k < z; k += nt) {
    float a = 0.f, b = 0.f;
    for (int r = 0; r < nr; ++r) { a += DMU[r * z + k]; b += DSD[r * z + k]; }
    atomicAdd(p.g[3] + k, a);
    atomicAdd(p.g[5] + k, b);
  }
  for (int o = tid; o < nr * h; o += nt) {
    const int r = o / h, j = o - r * h;
    float s = 0.f;
    for (int k = 0; k < z; ++k) s += Wmu[k * h + j] * DMU[r * z + k] + Wsg[k * h + j] * DSD[r * z + k];
    DH[o] = H[o] > 0.f ? s : s * p.slope;
  }
  __syncthreads();
  // ---- en_fc1
  for (int o = tid; o < h * d; o += nt) {
    const int j = o / d, i = o - j * d;
    float s = 0.f;
    for (int r = 0; r < nr; ++r) s += DH[r * h + j] * Y[r * d + i];
    atomicAdd(p.g[0] + o, s);
  }
  for (int j = tid; j < h; j += nt) {
    float s = 0.f;
    for (int r = 0; r < nr; ++r) s += DH[r * h + j];
    atomicAdd(p.g[1] + j, s);
  }
}

// =============================================================================================== optimiser / packing
__device__ __forceinline__ float adam_one(float pv, float g, float& m, float& v, float step_size, float beta1, float beta2,
                                          float eps, float wd, float inv_sqrt_bc2, float grad_scale) {
  const float gr = g * grad_scale + wd * pv;
  m = beta1 * m + (1.f - beta1) * gr;
  v = beta2 * v + (1.f - beta2) * gr * gr;
  return pv - step_size * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
}
// four parameters per thread and iteration (16-byte loads / stores; the flat stores are 16-byte aligned), scalar tail
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                  float* __restrict__ m, float* __restrict__ v, bf16* __restrict__ w16,
                                                  bf16* __restrict__ w16lo, long long n, float step_size, float beta1,
                                                  float beta2, float eps, float wd, float inv_sqrt_bc2, float grad_scale,
                                                  const float* __restrict__ hyper, int vec) {
  if (hyper) { step_size = hyper[0]; inv_sqrt_bc2 = hyper[1]; }  // CUDA-graph replays: step-dependent factors from memory
  const long long n4 = vec ? n >> 2 : 0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 pv = reinterpret_cast<const float4*>(p)[i], gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<const float4*>(m)[i], vv = reinterpret_cast<const float4*>(v)[i];
    float4 np;
    np.x = adam_one(pv.x, gv.x, mv.x, vv.x, step_size, beta1, beta2, eps, wd, inv_sqrt_bc2, grad_scale);
    np.y = adam_one(pv.y, gv.y, mv.y, vv.y, step_size, beta1, beta2, eps, wd, inv_sqrt_bc2, grad_scale);
    np.z = adam_one(pv.z, gv.z, mv.z, vv.z, step_size, beta1, beta2, eps, wd, inv_sqrt_bc2, grad_scale);
    np.w = adam_one(pv.w, gv.w, mv.w, vv.w, step_size, beta1, beta2, eps, wd, inv_sqrt_bc2, grad_scale);
    reinterpret_cast<float4*>(m)[i] = mv; reinterpret_cast<float4*>(v)[i] = vv; reinterpret_cast<float4*>(p)[i] = np;
    if (w16) {
      const uint32_t h0 = pack_bf16x2(np.x, np.y), h1 = pack_bf16x2(np.z, np.w);
      reinterpret_cast<uint2*>(w16)[i] = make_uint2(h0, h1);
      if (w16lo)   // bf16x3 operands: hi + lo = 16 mantissa bits
        reinterpret_cast<uint2*>(w16lo)[i] = make_uint2(pack_bf16x2(np.x - bf16lo(h0), np.y - bf16hi(h0)),
                                                        pack_bf16x2(np.z - bf16lo(h1), np.w - bf16hi(h1)));
    }
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    float mm = m[i], vv = v[i];
    const float np = adam_one(p[i], g[i], mm, vv, step_size, beta1, beta2, eps, wd, inv_sqrt_bc2, grad_scale);
    m[i] = mm; v[i] = vv; p[i] = np;
    if (w16) {
      const bf16 hi = __float2bfloat16(np);
      w16[i] = hi;
      if (w16lo) w16lo[i] = __float2bfloat16(np - __bfloat162float(hi));
    }
  }
}

__global__ void pack_dgrad_kernel(const float* __restrict__ w, bf16* __restrict__ wt, int cout, int cin) {
  __shared__ float tile[32][33];
  const int tap = blockIdx.z;
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  const float* src = w + (long long)tap * cout * cin;
  bf16* dst = wt + (long long)tap * cout * cin;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int co = co0 + r, ci = ci0 + threadIdx.x;
    tile[r][threadIdx.x] = (co < cout && ci < cin) ? src[(long long)co * cin + ci] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int ci = ci0 + r, co = co0 + threadIdx.x;
    if (ci < cin && co < cout) dst[(long long)ci * cout + co] = __float2bfloat16(tile[threadIdx.x][r]);
  }
}

// every conv weight of a store in ONE launch: desc[e] = {w_off, wt_off, taps, cout, cin, tile0} (int64 each; offsets in
// elements from the bases; tile0 = index of the entry's first 32x32 tile, desc[count].tile0 = total).  wt_lo (may be NULL)
// receives the bf16 remainder w - float(bf16(w)) for the split-bf16 ("bf16x3") operands.
__global__ void pack_dgrad_multi_kernel(const float* __restrict__ w, bf16* __restrict__ wt, bf16* __restrict__ wt_lo,
                                        const long long* __restrict__ desc, int count, int total_tiles) {
  __shared__ float tile[32][33];
  __shared__ int s_e;
  // persistent blocks walk the tile list (42 K tiles of 4 KB: one short-lived block per tile was launch-bound)
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    if (threadIdx.x == 0 && threadIdx.y == 0) {   // last entry whose first tile is <= t
      int lo = 0, hi = count - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(desc + mid * 6 + 5) <= (long long)t) lo = mid; else hi = mid - 1;
      }
      s_e = lo;
    }
    __syncthreads();
    const long long* d = desc + s_e * 6;
    const int cout = (int)d[3], cin = (int)d[4];
    const int tx = (cin + 31) / 32, ty = (cout + 31) / 32;
    int local = t - (int)d[5];
    const int tap = local / (tx * ty);
    local -= tap * tx * ty;
    const int ci0 = (local % tx) * 32, co0 = (local / tx) * 32;
    const float* src = w + d[0] + (long long)tap * cout * cin;
    const long long dbase = d[1] + (long long)tap * cout * cin;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int co = co0 + r, ci = ci0 + threadIdx.x;
      tile[r][threadIdx.x] = (co < cout && ci < cin) ? src[(long long)co * cin + ci] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int ci = ci0 + r, co = co0 + threadIdx.x;
      if (ci < cin && co < cout) {
        const float v = tile[threadIdx.x][r];
        const bf16 hi = __float2bfloat16(v);
        wt[dbase + (long long)ci * cout + co] = hi;
        if (wt_lo) wt_lo[dbase + (long long)ci * cout + co] = __float2bfloat16(v - __bfloat162float(hi));
      }
    }
    __syncthreads();     // tile / s_e are rewritten by the next iteration
  }
}

__global__ void __launch_bounds__(256) f32_split_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ hi,
                                                            bf16* __restrict__ lo, long long n) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float v = x[i];
    const bf16 h = __float2bfloat16(v);
    hi[i] = h;
    lo[i] = __float2bfloat16(v - __bfloat162float(h));
  }
}

// per-frame mean / max joint error in mm (depth_train.py:229-237 + handpose_evaluation.py:92-97,197-203)
__global__ void joint_errors_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                    const int* __restrict__ jidx, int nj, int j3, float sx, float sy, float sz,
                                    float* __restrict__ emean, float* __restrict__ emax, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = pred + (long long)i * j3;
  const float* g = gt + (long long)i * j3;
  float s = 0.f, m = 0.f;
  for (int q = 0; q < nj; ++q) {
    const int j = jidx ? jidx[q] : q;
    const float dx = (g[3 * j] - p[3 * j]) * sx, dy = (g[3 * j + 1] - p[3 * j + 1]) * sy, dz = (g[3 * j + 2] - p[3 * j + 2]) * sz;
    const float e = sqrtf(dx * dx + dy * dy + dz * dz);
    s += e;
    m = fmaxf(m, e);
  }
  emean[i] = s / nj;
  emax[i] = m;
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
    y[i] = __float2bfloat16(x[i]);
}
__global__ void bf16_to_f32_kernel(const bf16* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
    y[i] = __bfloat162float(x[i]);
}

}  // namespace

// tensor-core stems (stem_tc.cu); LSPS_STEM_SIMT=1 keeps the direct fp32 kernels of this file
int lsps_stem_fwd_tc(lsps_ctx* ctx, const float* img, const float* w, const float* bias, void* y, int n, int h, int wd,
                     int stride, float slope, int split, cudaStream_t st);
int lsps_stem_wgrad_tc(lsps_ctx* ctx, const float* img, const void* dy, float* dw, float* db, int n, int h, int wd,
                       int stride, int split, cudaStream_t st);
int lsps_stem_dgrad_tc(lsps_ctx* ctx, const void* dy, const float* w, float* dimg, int n, int h, int wd, int stride,
                       int accumulate, int split, cudaStream_t st);
static bool stem_use_tc(int wd, int stride) {
  static int simt = -1;
  if (simt < 0) { const char* e = getenv("LSPS_STEM_SIMT"); simt = (e && e[0] == '1') ? 1 : 0; }
  const int wo = wd / stride;
  return !simt && (wo == 64 || wo == 128);
}

static bool in_no_reg() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LSPS_IN_NO_REG"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

#define ST_(s) static_cast<cudaStream_t>(s)
#define REQUIRE(ctx, cond, code, msg) \
  do { if (!(cond)) return lsps_set_error(ctx, code, msg); } while (0)

// =============================================================================================== C ABI
extern "C" int lsps_stem_fwd(lsps_ctx* ctx, const float* img, const float* w, const float* bias, void* y, int n, int h,
                             int wd, int stride, float slope, lsps_stream st) {
  REQUIRE(ctx, img && w && bias && y, LSPS_E_ARG, "stem_fwd: null");
  REQUIRE(ctx, (stride == 1 || stride == 2) && h % (ST * stride) == 0 && wd % (ST * stride) == 0 && n > 0, LSPS_E_SHAPE,
          "stem_fwd: h,w must be multiples of 16*stride");
  if (stem_use_tc(wd, stride) && (h / stride) % (128 / (wd / stride)) == 0)
    return lsps_stem_fwd_tc(ctx, img, w, bias, y, n, h, wd, stride, slope, 0, ST_(st));
  dim3 grid(wd / stride / ST, h / stride / ST, n);
  if (stride == 1) stem_fwd_kernel<1><<<grid, 256, 0, ST_(st)>>>(img, w, bias, static_cast<bf16*>(y), h, wd, slope);
  else stem_fwd_kernel<2><<<grid, 256, 0, ST_(st)>>>(img, w, bias, static_cast<bf16*>(y), h, wd, slope);
  LSPS_CHECK_LAUNCH(ctx, "stem_fwd");
  return LSPS_OK;
}

extern "C" int lsps_stem_wgrad(lsps_ctx* ctx, const float* img, const void* dy, float* dw, float* db, int n, int h,
                               int wd, int stride, lsps_stream st) {
  REQUIRE(ctx, img && dy && dw, LSPS_E_ARG, "stem_wgrad: null");
  REQUIRE(ctx, (stride == 1 || stride == 2) && h % (ST * stride) == 0 && wd % (ST * stride) == 0 && n > 0, LSPS_E_SHAPE,
          "stem_wgrad: shape");
  if (stem_use_tc(wd, stride) && (h / stride) % (128 / (wd / stride)) == 0)
    return lsps_stem_wgrad_tc(ctx, img, dy, dw, db, n, h, wd, stride, 0, ST_(st));
  const int total = (wd / stride / ST) * (h / stride / ST) * n;
  const int grid = total < 2 * ctx->num_sms ? total : 2 * ctx->num_sms;
  if (stride == 1) stem_wgrad_kernel<1><<<grid, 256, 0, ST_(st)>>>(img, static_cast<const bf16*>(dy), dw, db, n, h, wd);
  else stem_wgrad_kernel<2><<<grid, 256, 0, ST_(st)>>>(img, static_cast<const bf16*>(dy), dw, db, n, h, wd);
  LSPS_CHECK_LAUNCH(ctx, "stem_wgrad");
  return LSPS_OK;
}

extern "C" int lsps_stem_dgrad(lsps_ctx* ctx, const void* dy, const float* w, float* dimg, int n, int h, int wd,
                               int stride, int accumulate, lsps_stream st) {
  REQUIRE(ctx, dy && w && dimg, LSPS_E_ARG, "stem_dgrad: null");
  REQUIRE(ctx, (stride == 1 || stride == 2) && h % (ST * stride) == 0 && wd % (ST * stride) == 0 && n > 0, LSPS_E_SHAPE,
          "stem_dgrad: shape");
  if (stem_use_tc(wd, stride)) return lsps_stem_dgrad_tc(ctx, dy, w, dimg, n, h, wd, stride, accumulate, 0, ST_(st));
  dim3 grid(wd / ST, h / ST, n);
  const int de = stride == 1 ? ST + 6 : (ST + 6) / 2 + 1;
  const int smem = 49 * 64 * 4 + de * de * 128;
  if (stride == 1) {
    static bool cfg = false;
    if (!cfg) { cudaFuncSetAttribute(stem_dgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); cfg = true; }
    stem_dgrad_kernel<1><<<grid, 256, smem, ST_(st)>>>(static_cast<const bf16*>(dy), w, dimg, h, wd, accumulate);
  } else {
    stem_dgrad_kernel<2><<<grid, 256, smem, ST_(st)>>>(static_cast<const bf16*>(dy), w, dimg, h, wd, accumulate);
  }
  LSPS_CHECK_LAUNCH(ctx, "stem_dgrad");
  return LSPS_OK;
}

// bf16x3 variants (tensor-core kernels only: output width 64 or 128): y / dy are [n,ho,wo,128] = (hi | lo) halves
extern "C" int lsps_stem_fwd_split(lsps_ctx* ctx, const float* img, const float* w, const float* bias, void* y, int n,
                                   int h, int wd, int stride, float slope, lsps_stream st) {
  REQUIRE(ctx, img && w && bias && y, LSPS_E_ARG, "stem_fwd_split: null");
  const int wo = wd / stride;
  REQUIRE(ctx, (stride == 1 || stride == 2) && n > 0 && (wo == 64 || wo == 128) && (h / stride) % (128 / wo) == 0,
          LSPS_E_SHAPE, "stem_fwd_split: output width must be 64 or 128");
  return lsps_stem_fwd_tc(ctx, img, w, bias, y, n, h, wd, stride, slope, 1, ST_(st));
}
extern "C" int lsps_stem_wgrad_split(lsps_ctx* ctx, const float* img, const void* dy, float* dw, float* db, int n, int h,
                                     int wd, int stride, lsps_stream st) {
  REQUIRE(ctx, img && dy && dw, LSPS_E_ARG, "stem_wgrad_split: null");
  const int wo = wd / stride;
  REQUIRE(ctx, (stride == 1 || stride == 2) && n > 0 && (wo == 64 || wo == 128) && (h / stride) % (128 / wo) == 0,
          LSPS_E_SHAPE, "stem_wgrad_split: output width must be 64 or 128");
  return lsps_stem_wgrad_tc(ctx, img, dy, dw, db, n, h, wd, stride, 1, ST_(st));
}
extern "C" int lsps_stem_dgrad_split(lsps_ctx* ctx, const void* dy, const float* w, float* dimg, int n, int h, int wd,
                                     int stride, int accumulate, lsps_stream st) {
  REQUIRE(ctx, dy && w && dimg, LSPS_E_ARG, "stem_dgrad_split: null");
  const int wo = wd / stride;
  REQUIRE(ctx, (stride == 1 || stride == 2) && n > 0 && (wo == 64 || wo == 128) && h % 8 == 0 && wd % 16 == 0, LSPS_E_SHAPE,
          "stem_dgrad_split: output width must be 64 or 128");
  return lsps_stem_dgrad_tc(ctx, dy, w, dimg, n, h, wd, stride, accumulate, 1, ST_(st));
}

extern "C" int lsps_head_fwd(lsps_ctx* ctx, const void* x, const float* w, const float* bias, float* out,
                             long long npix, lsps_stream st) {
  REQUIRE(ctx, x && w && bias && out && npix > 0, LSPS_E_ARG, "head_fwd: null");
  head_fwd_kernel<<<grid_for(npix, 32, 16 * ctx->num_sms), 256, 0, ST_(st)>>>(static_cast<const bf16*>(x), w, bias, out, npix,
                                                                            nullptr, 0, 0, 0.f, nullptr, nullptr);
  LSPS_CHECK_LAUNCH(ctx, "head_fwd");
  return LSPS_OK;
}
extern "C" int lsps_head_fwd_l1(lsps_ctx* ctx, const void* x, const float* w, const float* bias, float* out,
                                long long npix, const float* target, long long t0, long long tn, float scale,
                                float* dout, float* acc, lsps_stream st) {
  REQUIRE(ctx, x && w && bias && out && npix > 0 && target && acc && t0 >= 0 && tn > 0 && t0 + tn <= npix, LSPS_E_ARG,
          "head_fwd_l1: arg");
  head_fwd_kernel<<<grid_for(npix, 32, 16 * ctx->num_sms), 256, 0, ST_(st)>>>(static_cast<const bf16*>(x), w, bias, out, npix,
                                                                            target, t0, tn, scale, dout, acc);
  LSPS_CHECK_LAUNCH(ctx, "head_fwd_l1");
  return LSPS_OK;
}
extern "C" int lsps_head_bwd(lsps_ctx* ctx, const void* x, const float* w, const float* out, const float* dout,
                             void* dx, float* dw, float* db, long long npix, float slope, lsps_stream st) {
  REQUIRE(ctx, x && w && out && dout && dx && dw && db && npix > 0, LSPS_E_ARG, "head_bwd: null");
  head_bwd_kernel<<<grid_for(npix, 32, 8 * ctx->num_sms), 256, 0, ST_(st)>>>(static_cast<const bf16*>(x), w, out, dout,
                                                                           static_cast<bf16*>(dx), dw, db, npix, slope);
  LSPS_CHECK_LAUNCH(ctx, "head_bwd");
  return LSPS_OK;
}

extern "C" int lsps_instnorm_fwd(lsps_ctx* ctx, const void* h, const void* res, void* y, float* stats, int n, int hw,
                                 int c, int mode, float eps, float slope, lsps_stream st) {
  REQUIRE(ctx, h && y && stats && (mode == 0 || res), LSPS_E_ARG, "instnorm_fwd: null");
  REQUIRE(ctx, n > 0 && hw > 0 && c % 64 == 0, LSPS_E_SHAPE, "instnorm_fwd: c must be a multiple of 64");
  if (hw == 128 * INR_PPT && !in_no_reg()) {
    if (mode == 0) instnorm_fwd_reg_kernel<0><<<dim3(c / 32, n), 512, 0, ST_(st)>>>(static_cast<const bf16*>(h), nullptr, static_cast<bf16*>(y), stats, c, eps, slope);
    else instnorm_fwd_reg_kernel<1><<<dim3(c / 32, n), 512, 0, ST_(st)>>>(static_cast<const bf16*>(h), static_cast<const bf16*>(res), static_cast<bf16*>(y), stats, c, eps, slope);
    LSPS_CHECK_LAUNCH(ctx, "instnorm_fwd_reg");
    return LSPS_OK;
  }
  instnorm_fwd_kernel<<<dim3(c / 64, n), 512, 0, ST_(st)>>>(static_cast<const bf16*>(h), static_cast<const bf16*>(res),
                                                          static_cast<bf16*>(y), stats, hw, c, mode, eps, slope);
  LSPS_CHECK_LAUNCH(ctx, "instnorm_fwd");
  return LSPS_OK;
}
extern "C" int lsps_instnorm_bwd(lsps_ctx* ctx, const void* dy, const void* h, const float* stats, void* dh, int n,
                                 int hw, int c, int mode, float slope, float* db, lsps_stream st) {
  REQUIRE(ctx, dy && h && stats && dh, LSPS_E_ARG, "instnorm_bwd: null");
  REQUIRE(ctx, n > 0 && hw > 0 && c % 64 == 0, LSPS_E_SHAPE, "instnorm_bwd: c must be a multiple of 64");
  instnorm_bwd_kernel<<<dim3(c / 64, n), 512, 0, ST_(st)>>>(static_cast<const bf16*>(dy), static_cast<const bf16*>(h),
                                                          stats, static_cast<bf16*>(dh), hw, c, mode, slope, db,
                                                          nullptr, 0);
  LSPS_CHECK_LAUNCH(ctx, "instnorm_bwd");
  return LSPS_OK;
}
extern "C" int lsps_instnorm_bwd_grouped(lsps_ctx* ctx, const void* dy, const void* h, const float* stats, void* dh,
                                         int n, int hw, int c, int mode, float slope, float* db, float* db2, int n_split,
                                         lsps_stream st) {
  REQUIRE(ctx, dy && h && stats && dh && db && db2, LSPS_E_ARG, "instnorm_bwd_grouped: null");
  REQUIRE(ctx, n > 0 && hw > 0 && c % 64 == 0 && n_split > 0 && n_split < n, LSPS_E_SHAPE, "instnorm_bwd_grouped: shape");
  instnorm_bwd_kernel<<<dim3(c / 64, n), 512, 0, ST_(st)>>>(static_cast<const bf16*>(dy), static_cast<const bf16*>(h),
                                                          stats, static_cast<bf16*>(dh), hw, c, mode, slope, db, db2,
                                                          n_split);
  LSPS_CHECK_LAUNCH(ctx, "instnorm_bwd_grouped");
  return LSPS_OK;
}

// ---- norm from epilogue statistics (InstanceNorm: per_image = 1, rows [n][2][c]; BatchNorm: per_image = 0, ONE row)
static bool norm_shape_ok(int n, int hw, int c) { return n > 0 && hw > 0 && c >= 64 && c <= 2048 && (c & (c - 1)) == 0; }
static int norm_ppb(int hw, int c) {
  const int lanes = 256 / (c / 8);
  int ppb = lanes * 8;
  return ppb > hw ? hw : ppb;
}
extern "C" int lsps_norm_apply_fwd(lsps_ctx* ctx, const void* h, const void* res, void* y, const float* sums,
                                   float* stats_out, int n, int hw, int c, int mode, int per_image, float eps,
                                   float slope, const float* gamma, const float* beta, lsps_stream st) {
  REQUIRE(ctx, h && y && sums && (mode != 1 || res), LSPS_E_ARG, "norm_apply_fwd: null");
  REQUIRE(ctx, norm_shape_ok(n, hw, c) && mode >= 0 && mode <= 2, LSPS_E_SHAPE, "norm_apply_fwd: c must be a power of two in [64, 2048]");
  const int ppb = norm_ppb(hw, c);
  const dim3 grid((hw + ppb - 1) / ppb, n);
  const long long stride = per_image ? 2LL * c : 0;
  const float inv = 1.f / (per_image ? (float)hw : (float)hw * (float)n);
  const bf16 *hh = static_cast<const bf16*>(h), *rr = static_cast<const bf16*>(res);
  bf16* yy = static_cast<bf16*>(y);
  float* so = stats_out;
  if (mode == 0) norm_apply_fwd_kernel<0><<<grid, 256, 0, ST_(st)>>>(hh, rr, yy, sums, so, hw, c, ppb, stride, inv, eps, slope, gamma, beta);
  else if (mode == 1) norm_apply_fwd_kernel<1><<<grid, 256, 0, ST_(st)>>>(hh, rr, yy, sums, so, hw, c, ppb, stride, inv, eps, slope, gamma, beta);
  else norm_apply_fwd_kernel<2><<<grid, 256, 0, ST_(st)>>>(hh, rr, yy, sums, so, hw, c, ppb, stride, inv, eps, slope, gamma, beta);
  LSPS_CHECK_LAUNCH(ctx, "norm_apply_fwd");
  return LSPS_OK;
}
extern "C" int lsps_norm_bwd_stats(lsps_ctx* ctx, const void* dy, const void* h, const float* stats, float* bsums, int n,
                                   int hw, int c, int mode, int per_image, float slope, const float* gamma,
                                   const float* beta, lsps_stream st) {
  REQUIRE(ctx, dy && h && stats && bsums, LSPS_E_ARG, "norm_bwd_stats: null");
  REQUIRE(ctx, norm_shape_ok(n, hw, c) && (mode == 0 || mode == 1), LSPS_E_SHAPE, "norm_bwd_stats: shape");
  const int lanes = 256 / (c / 8);
  int ppb = lanes * 16;
  if (ppb > hw) ppb = hw;
  const dim3 grid((hw + ppb - 1) / ppb, n);
  const long long stride = per_image ? 2LL * c : 0;
  if (cudaMemsetAsync(bsums, 0, (size_t)(per_image ? n : 1) * 2 * c * sizeof(float), ST_(st)) != cudaSuccess)
    return lsps_set_error(ctx, LSPS_E_CUDA, "norm_bwd_stats: memset");
  const int smem = 2 * lanes * c * (int)sizeof(float);
  const bf16 *gg = static_cast<const bf16*>(dy), *hh = static_cast<const bf16*>(h);
  if (mode == 0) norm_bwd_stats_kernel<0><<<grid, 256, smem, ST_(st)>>>(gg, hh, stats, bsums, hw, c, ppb, stride, stride, slope, gamma, beta);
  else norm_bwd_stats_kernel<1><<<grid, 256, smem, ST_(st)>>>(gg, hh, stats, bsums, hw, c, ppb, stride, stride, slope, gamma, beta);
  LSPS_CHECK_LAUNCH(ctx, "norm_bwd_stats");
  return LSPS_OK;
}
extern "C" int lsps_norm_bwd_apply(lsps_ctx* ctx, const void* g, const void* h, const float* stats, const float* bsums,
                                   void* dh, int n, int hw, int c, int gmode, int per_image, float slope,
                                   const float* gamma, const float* beta, lsps_stream st) {
  REQUIRE(ctx, g && h && stats && bsums && dh, LSPS_E_ARG, "norm_bwd_apply: null");
  REQUIRE(ctx, norm_shape_ok(n, hw, c) && gmode >= 0 && gmode <= 2 && (gmode != 2 || slope > 0.f), LSPS_E_SHAPE,
          "norm_bwd_apply: shape (gmode 2 recovers xhat from the activation and needs slope > 0)");
  const int ppb = norm_ppb(hw, c);
  const dim3 grid((hw + ppb - 1) / ppb, n);
  const long long stride = per_image ? 2LL * c : 0;
  const float inv = 1.f / (per_image ? (float)hw : (float)hw * (float)n);
  const bf16 *gg = static_cast<const bf16*>(g), *hh = static_cast<const bf16*>(h);
  if (gmode == 0) norm_bwd_apply_kernel<0><<<grid, 256, 0, ST_(st)>>>(gg, hh, stats, bsums, static_cast<bf16*>(dh), hw, c, ppb, stride, stride, inv, slope, gamma, beta);
  else if (gmode == 1) norm_bwd_apply_kernel<1><<<grid, 256, 0, ST_(st)>>>(gg, hh, stats, bsums, static_cast<bf16*>(dh), hw, c, ppb, stride, stride, inv, slope, gamma, beta);
  else norm_bwd_apply_kernel<2><<<grid, 256, 0, ST_(st)>>>(gg, hh, stats, bsums, static_cast<bf16*>(dh), hw, c, ppb, stride, stride, inv, slope, gamma, beta);
  LSPS_CHECK_LAUNCH(ctx, "norm_bwd_apply");
  return LSPS_OK;
}
extern "C" int lsps_bn_running_update(lsps_ctx* ctx, const float* sums, float* running_mean, float* running_var, int c,
                                      float count, float momentum, lsps_stream st) {
  REQUIRE(ctx, sums && running_mean && running_var && c > 0 && count > 0.f, LSPS_E_ARG, "bn_running_update: arg");
  bn_running_update_kernel<<<(c + 127) / 128, 128, 0, ST_(st)>>>(sums, running_mean, running_var, c, count, momentum);
  LSPS_CHECK_LAUNCH(ctx, "bn_running_update");
  return LSPS_OK;
}
extern "C" int lsps_bn_running_to_sums(lsps_ctx* ctx, const float* running_mean, const float* running_var, float* sums,
                                       int c, float count, lsps_stream st) {
  REQUIRE(ctx, sums && running_mean && running_var && c > 0 && count > 0.f, LSPS_E_ARG, "bn_running_to_sums: arg");
  bn_running_to_sums_kernel<<<(c + 127) / 128, 128, 0, ST_(st)>>>(running_mean, running_var, sums, c, count);
  LSPS_CHECK_LAUNCH(ctx, "bn_running_to_sums");
  return LSPS_OK;
}
extern "C" int lsps_norm_reduce_images(lsps_ctx* ctx, const float* sums, float* out, int n, int c, lsps_stream st) {
  REQUIRE(ctx, sums && out && n > 0 && c > 0, LSPS_E_ARG, "norm_reduce_images: arg");
  norm_reduce_images_kernel<<<(2 * c + 255) / 256, 256, 0, ST_(st)>>>(sums, out, n, 2 * c);
  LSPS_CHECK_LAUNCH(ctx, "norm_reduce_images");
  return LSPS_OK;
}

extern "C" int lsps_noise_kl_fwd(lsps_ctx* ctx, const void* x, const float* noise, void* z, float* acc, long long n,
                                 lsps_stream st) {
  REQUIRE(ctx, x && noise && z && acc, LSPS_E_ARG, "noise_kl: null");
  REQUIRE(ctx, n > 0 && n % 8 == 0, LSPS_E_SHAPE, "noise_kl: n % 8");
  noise_kl_kernel<<<grid_for(n / 8, 256, 8 * ctx->num_sms), 256, 0, ST_(st)>>>(static_cast<const bf16*>(x), noise,
                                                                             static_cast<bf16*>(z), acc, n / 8);
  LSPS_CHECK_LAUNCH(ctx, "noise_kl");
  return LSPS_OK;
}
extern "C" int lsps_noise_kl_philox(lsps_ctx* ctx, const void* x, void* z, float* acc, long long n,
                                    unsigned long long seed, unsigned long long offset, lsps_stream st) {
  REQUIRE(ctx, x && z && acc, LSPS_E_ARG, "noise_kl_philox: null");
  REQUIRE(ctx, n > 0 && n % 8 == 0, LSPS_E_SHAPE, "noise_kl_philox: n % 8");
  noise_kl_philox_kernel<<<grid_for(n / 8, 256, 16 * ctx->num_sms), 256, 0, ST_(st)>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(z), acc, n / 8, seed, offset);
  LSPS_CHECK_LAUNCH(ctx, "noise_kl_philox");
  return LSPS_OK;
}
// plain stream-ordered fills / copies (graph-capturable; no kernel of any library involved)
extern "C" int lsps_memset(lsps_ctx* ctx, void* dst, int value, long long bytes, lsps_stream st) {
  REQUIRE(ctx, dst && bytes >= 0, LSPS_E_ARG, "memset: arg");
  if (bytes && cudaMemsetAsync(dst, value, (size_t)bytes, ST_(st)) != cudaSuccess) return lsps_set_error(ctx, LSPS_E_CUDA, "memset failed");
  return LSPS_OK;
}
extern "C" int lsps_memcpy(lsps_ctx* ctx, void* dst, const void* src, long long bytes, lsps_stream st) {
  REQUIRE(ctx, dst && src && bytes >= 0, LSPS_E_ARG, "memcpy: arg");
  if (bytes && cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, ST_(st)) != cudaSuccess)
    return lsps_set_error(ctx, LSPS_E_CUDA, "memcpy failed");
  return LSPS_OK;
}
extern "C" int lsps_axpy_bf16(lsps_ctx* ctx, const void* a, const void* b, float alpha, void* out, long long n,
                              lsps_stream st) {
  REQUIRE(ctx, b && out, LSPS_E_ARG, "axpy: null");
  REQUIRE(ctx, n > 0 && n % 8 == 0, LSPS_E_SHAPE, "axpy: n % 8");
  axpy_bf16_kernel<<<grid_for(n / 8, 256, 8 * ctx->num_sms), 256, 0, ST_(st)>>>(
      static_cast<const bf16*>(a), static_cast<const bf16*>(b), alpha, static_cast<bf16*>(out), n / 8);
  LSPS_CHECK_LAUNCH(ctx, "axpy");
  return LSPS_OK;
}
extern "C" int lsps_l2_bf16(lsps_ctx* ctx, const void* a, const void* b, void* g, float scale, float* acc, long long n,
                            lsps_stream st) {
  REQUIRE(ctx, a && b && g && acc, LSPS_E_ARG, "l2_bf16: null");
  REQUIRE(ctx, n > 0 && n % 8 == 0, LSPS_E_SHAPE, "l2_bf16: n % 8");
  l2_bf16_kernel<<<grid_for(n / 8, 256, 8 * ctx->num_sms), 256, 0, ST_(st)>>>(
      static_cast<const bf16*>(a), static_cast<const bf16*>(b), static_cast<bf16*>(g), scale, acc, n / 8);
  LSPS_CHECK_LAUNCH(ctx, "l2_bf16");
  return LSPS_OK;
}
extern "C" int lsps_l1_f32(lsps_ctx* ctx, const float* x, const float* t, float* dx, float scale, int accumulate,
                           float* acc, long long n, lsps_stream st) {
  REQUIRE(ctx, x && t && acc && n > 0, LSPS_E_ARG, "l1_f32: null");
  l1_f32_kernel<<<grid_for(n, 256, 4 * ctx->num_sms), 256, 0, ST_(st)>>>(x, t, dx, scale, accumulate, acc, n);
  LSPS_CHECK_LAUNCH(ctx, "l1_f32");
  return LSPS_OK;
}
static int run_l1_feat(lsps_ctx* ctx, const void* a, const void* b, float* da, float* db, float scale, float* acc,
                       long long n, int sc, lsps_stream st) {
  REQUIRE(ctx, a && b && acc && n > 0 && (sc == 0 || n % sc == 0), LSPS_E_ARG, "l1_feat: arg");
  l1_feat_kernel<<<grid_for(n, 256, 4 * ctx->num_sms), 256, 0, ST_(st)>>>(static_cast<const bf16*>(a),
                                                                        static_cast<const bf16*>(b), da, db, scale, acc, n, sc);
  LSPS_CHECK_LAUNCH(ctx, "l1_feat");
  return LSPS_OK;
}
extern "C" int lsps_l1_feat(lsps_ctx* ctx, const void* a, const void* b, float* da, float* db, float scale, float* acc,
                            long long n, lsps_stream st) {
  return run_l1_feat(ctx, a, b, da, db, scale, acc, n, 0, st);
}
extern "C" int lsps_l1_feat_split(lsps_ctx* ctx, const void* a, const void* b, float* da, float* db, float scale,
                                  float* acc, long long n, int c, lsps_stream st) {
  REQUIRE(ctx, c > 0, LSPS_E_ARG, "l1_feat_split: c");
  return run_l1_feat(ctx, a, b, da, db, scale, acc, n, c, st);
}
static int run_dhead_fwd(lsps_ctx* ctx, const void* f, const float* w, const float* bias, float* logits, long long rows,
                         int c, int split, lsps_stream st) {
  REQUIRE(ctx, f && w && bias && logits && rows > 0, LSPS_E_ARG, "dhead_fwd: null");
  REQUIRE(ctx, c % 8 == 0, LSPS_E_SHAPE, "dhead_fwd: c % 8");
  dhead_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ST_(st)>>>(static_cast<const bf16*>(f), w, bias, logits, rows, c, split);
  LSPS_CHECK_LAUNCH(ctx, "dhead_fwd");
  return LSPS_OK;
}
extern "C" int lsps_dhead_fwd(lsps_ctx* ctx, const void* f, const float* w, const float* bias, float* logits,
                              long long rows, int c, lsps_stream st) {
  return run_dhead_fwd(ctx, f, w, bias, logits, rows, c, 0, st);
}
extern "C" int lsps_dhead_fwd_split(lsps_ctx* ctx, const void* f, const float* w, const float* bias, float* logits,
                                    long long rows, int c, lsps_stream st) {
  return run_dhead_fwd(ctx, f, w, bias, logits, rows, c, 1, st);
}
extern "C" int lsps_bce_logits(lsps_ctx* ctx, const float* logits, float target, float scale, float* dlogits,
                               float* acc, long long rows, lsps_stream st) {
  REQUIRE(ctx, logits && acc && rows > 0, LSPS_E_ARG, "bce: null");
  bce_logits_kernel<<<grid_for(rows, 256, ctx->num_sms), 256, 0, ST_(st)>>>(logits, target, scale, dlogits, acc, rows);
  LSPS_CHECK_LAUNCH(ctx, "bce_logits");
  return LSPS_OK;
}
static int run_dhead_bwd(lsps_ctx* ctx, const void* f, const float* w, const float* dlogits, float* df, float* dw,
                         float* db, long long rows, int c, int split, lsps_stream st) {
  REQUIRE(ctx, f && w && dlogits && rows > 0, LSPS_E_ARG, "dhead_bwd: null");
  const int rpb = 32;
  dim3 grid((c + 255) / 256, (unsigned)((rows + rpb - 1) / rpb));
  dhead_bwd_kernel<<<grid, 256, 0, ST_(st)>>>(static_cast<const bf16*>(f), w, dlogits, df, dw, db, rows, c, rpb, split);
  LSPS_CHECK_LAUNCH(ctx, "dhead_bwd");
  return LSPS_OK;
}
extern "C" int lsps_dhead_bwd(lsps_ctx* ctx, const void* f, const float* w, const float* dlogits, float* df, float* dw,
                              float* db, long long rows, int c, lsps_stream st) {
  return run_dhead_bwd(ctx, f, w, dlogits, df, dw, db, rows, c, 0, st);
}
extern "C" int lsps_dhead_bwd_split(lsps_ctx* ctx, const void* f, const float* w, const float* dlogits, float* df,
                                    float* dw, float* db, long long rows, int c, lsps_stream st) {
  return run_dhead_bwd(ctx, f, w, dlogits, df, dw, db, rows, c, 1, st);
}
static int run_mask_to_bf16(lsps_ctx* ctx, const float* df, const void* f, void* out, float slope, long long n, int sc,
                            lsps_stream st) {
  REQUIRE(ctx, df && f && out && n > 0 && (sc == 0 || n % sc == 0), LSPS_E_ARG, "mask_to_bf16: arg");
  mask_to_bf16_kernel<<<grid_for(n, 256, 8 * ctx->num_sms), 256, 0, ST_(st)>>>(df, static_cast<const bf16*>(f),
                                                                             static_cast<bf16*>(out), slope, n, sc);
  LSPS_CHECK_LAUNCH(ctx, "mask_to_bf16");
  return LSPS_OK;
}
extern "C" int lsps_dhead_bce(lsps_ctx* ctx, const void* f, const float* w, const float* bias, long long rows, int c,
                              int split, long long rows_per_group, int ngroups, const float* targets, const int* slots,
                              float scale, float* logits, float* df, float* dw, float* db, float* acc, lsps_stream st) {
  REQUIRE(ctx, f && w && bias && acc && targets && slots && rows > 0, LSPS_E_ARG, "dhead_bce: null");
  REQUIRE(ctx, c % 8 == 0 && c <= 2048 && ngroups >= 1 && ngroups <= 8 && rows_per_group > 0 &&
               rows_per_group * ngroups >= rows, LSPS_E_SHAPE, "dhead_bce: c <= 2048, <= 8 row groups covering all rows");
  DheadGroups G;
  for (int i = 0; i < 8; ++i) { G.tgt[i] = i < ngroups ? targets[i] : -1.f; G.slot[i] = i < ngroups ? slots[i] : 0; }
  // few rows per warp keeps the dw atomics count down (rows/rpw x c) while still filling the machine
  int rpw = (int)((rows + 8LL * 2 * ctx->num_sms - 1) / (8LL * 2 * ctx->num_sms));
  if (rpw < 1) rpw = 1;
  const long long warps = (rows + rpw - 1) / rpw;
  dhead_bce_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, ST_(st)>>>(static_cast<const bf16*>(f), w, bias, rows, c, split,
                                                                   rows_per_group, G, scale, logits, df, dw, db, acc, rpw);
  LSPS_CHECK_LAUNCH(ctx, "dhead_bce");
  return LSPS_OK;
}
extern "C" int lsps_mask_to_bf16(lsps_ctx* ctx, const float* df, const void* f, void* out, float slope, long long n,
                                 lsps_stream st) {
  return run_mask_to_bf16(ctx, df, f, out, slope, n, 0, st);
}
extern "C" int lsps_mask_to_bf16_split(lsps_ctx* ctx, const float* df, const void* f, void* out, float slope, long long n,
                                       int c, lsps_stream st) {
  REQUIRE(ctx, c > 0, LSPS_E_ARG, "mask_to_bf16_split: c");
  return run_mask_to_bf16(ctx, df, f, out, slope, n, c, st);
}
static int run_colsum(lsps_ctx* ctx, const void* dy, long long rows, int c, float* db, int fold, lsps_stream st);
extern "C" int lsps_colsum_bf16(lsps_ctx* ctx, const void* dy, long long rows, int c, float* db, lsps_stream st) {
  return run_colsum(ctx, dy, rows, c, db, 0, st);
}
// dy split [rows][hi c | lo c]: db[c] += column sums of hi + lo
extern "C" int lsps_colsum_bf16_split(lsps_ctx* ctx, const void* dy, long long rows, int c, float* db, lsps_stream st) {
  return run_colsum(ctx, dy, rows, 2 * c, db, c, st);
}
static int run_colsum(lsps_ctx* ctx, const void* dy, long long rows, int c, float* db, int fold, lsps_stream st) {
  REQUIRE(ctx, dy && db && rows > 0, LSPS_E_ARG, "colsum: null");
  REQUIRE(ctx, c % 8 == 0 && (c / 8 >= 256 ? (c / 8) % 256 == 0 : 256 % (c / 8) == 0), LSPS_E_SHAPE, "colsum: c");
  const int octs = c / 8;
  const int row_lanes = octs >= 256 ? 1 : 256 / octs;
  const int gy = octs >= 256 ? octs / 256 : 1;
  long long gx = (rows + row_lanes - 1) / row_lanes;
  const long long cap = (2LL * ctx->num_sms + gy - 1) / gy;
  if (gx > cap) gx = cap;
  colsum_kernel<<<dim3((unsigned)gx, gy), 256, 0, ST_(st)>>>(static_cast<const bf16*>(dy), rows, c, db, row_lanes, fold);
  LSPS_CHECK_LAUNCH(ctx, "colsum");
  return LSPS_OK;
}

extern "C" int lsps_linear_fwd(lsps_ctx* ctx, const void* x, int x_bf16, const float* w, const float* b, float* y, int m,
                               int n, int k, int act, float slope, lsps_stream st) {
  REQUIRE(ctx, x && w && y && m > 0 && n > 0 && k > 0, LSPS_E_ARG, "linear_fwd: arg");
  const unsigned grid = (unsigned)(((long long)m * n + 7) / 8);
  // x_bf16: 0 = f32 rows, 1 = bf16 rows, c > 1 = split-bf16 rows ([pixels][hi c | lo c], k a multiple of c)
  const int sc = x_bf16 > 1 ? x_bf16 : 0;
  REQUIRE(ctx, sc == 0 || k % sc == 0, LSPS_E_SHAPE, "linear_fwd: k must be a multiple of the split channel count");
  if (x_bf16) linear_fwd_kernel<true><<<grid, 256, 0, ST_(st)>>>(x, w, b, y, m, n, k, act, slope, sc);
  else linear_fwd_kernel<false><<<grid, 256, 0, ST_(st)>>>(x, w, b, y, m, n, k, act, slope, 0);
  LSPS_CHECK_LAUNCH(ctx, "linear_fwd");
  return LSPS_OK;
}
extern "C" int lsps_linear_bwd(lsps_ctx* ctx, const void* x, int x_bf16, const float* w, const float* dy, float* dx,
                               int dx_accumulate, float* dw, float* db, int m, int n, int k, lsps_stream st) {
  REQUIRE(ctx, dy && m > 0 && n > 0 && k > 0, LSPS_E_ARG, "linear_bwd: arg");
  if (dw) {
    REQUIRE(ctx, x, LSPS_E_ARG, "linear_bwd: dw needs x");
    const unsigned grid = (unsigned)(((long long)n * k + 255) / 256);
    const int sc = x_bf16 > 1 ? x_bf16 : 0;
    if (x_bf16) linear_bwd_dw_kernel<true><<<grid, 256, 0, ST_(st)>>>(x, dy, dw, m, n, k, sc);
    else linear_bwd_dw_kernel<false><<<grid, 256, 0, ST_(st)>>>(x, dy, dw, m, n, k, 0);
    LSPS_CHECK_LAUNCH(ctx, "linear_bwd_dw");
  }
  if (dx) {
    REQUIRE(ctx, w, LSPS_E_ARG, "linear_bwd: dx needs w");
    linear_bwd_dx_kernel<<<(unsigned)(((long long)m * k + 255) / 256), 256, 0, ST_(st)>>>(w, dy, dx, dx_accumulate, m, n, k);
    LSPS_CHECK_LAUNCH(ctx, "linear_bwd_dx");
  }
  if (db) {
    linear_bwd_db_kernel<<<(n + 127) / 128, 128, 0, ST_(st)>>>(dy, db, m, n);
    LSPS_CHECK_LAUNCH(ctx, "linear_bwd_db");
  }
  return LSPS_OK;
}
extern "C" int lsps_act_bwd(lsps_ctx* ctx, float* dy, const float* y, int act, float slope, long long n, lsps_stream st) {
  REQUIRE(ctx, dy && y && n > 0, LSPS_E_ARG, "act_bwd: arg");
  act_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST_(st)>>>(dy, y, act, slope, n);
  LSPS_CHECK_LAUNCH(ctx, "act_bwd");
  return LSPS_OK;
}
extern "C" int lsps_mse(lsps_ctx* ctx, const float* p, const float* e, float* dp, float scale, float* acc, long long n,
                        lsps_stream st) {
  REQUIRE(ctx, p && e && acc && n > 0, LSPS_E_ARG, "mse: arg");
  mse_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, ST_(st)>>>(p, e, dp, scale, acc, n);
  LSPS_CHECK_LAUNCH(ctx, "mse");
  return LSPS_OK;
}
extern "C" int lsps_vae_reparam(lsps_ctx* ctx, const float* mu, const float* sd, const float* noise, float* z,
                                float* acc, long long n, lsps_stream st) {
  REQUIRE(ctx, mu && sd && noise && z && n > 0, LSPS_E_ARG, "vae_reparam: arg");
  vae_reparam_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, ST_(st)>>>(mu, sd, noise, z, acc, n);
  LSPS_CHECK_LAUNCH(ctx, "vae_reparam");
  return LSPS_OK;
}
extern "C" int lsps_vae_reparam_bwd(lsps_ctx* ctx, const float* mu, const float* sd, const float* noise,
                                    const float* dz, float* dmu, float* dsd, float kl_scale, long long n,
                                    lsps_stream st) {
  REQUIRE(ctx, mu && sd && noise && dz && dmu && dsd && n > 0, LSPS_E_ARG, "vae_reparam_bwd: arg");
  vae_reparam_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST_(st)>>>(mu, sd, noise, dz, dmu, dsd, kl_scale, n);
  LSPS_CHECK_LAUNCH(ctx, "vae_reparam_bwd");
  return LSPS_OK;
}

extern "C" int lsps_vae_step(lsps_ctx* ctx, const float* y, const float* noise, const float* const* weights,
                             float* const* grads, float* dec, float* acc, int rows, int d, int h, int z, float ll_scale,
                             float kl_scale, float slope, lsps_stream st) {
  REQUIRE(ctx, y && noise && weights && grads && dec && acc, LSPS_E_ARG, "vae_step: null");
  if (!(rows > 0 && d > 0 && h > 0 && z > 0 && d <= 256 && h <= 128 && z <= 64 && slope > 0.f))
    return lsps_set_error(ctx, LSPS_E_SHAPE, "vae_step: rows %d dims %d/%d/%d (d <= 256, h <= 128, z <= 64)", rows, d, h, z);
  VaeStepParams p{};
  p.y = y; p.noise = noise; p.dec = dec; p.acc = acc; p.rows = rows; p.d = d; p.h = h; p.z = z;
  p.ll_scale = ll_scale; p.kl_scale = kl_scale; p.slope = slope;
  for (int i = 0; i < 10; ++i) {
    if (!weights[i] || !grads[i]) return lsps_set_error(ctx, LSPS_E_ARG, "vae_step: null parameter %d", i);
    p.w[i] = weights[i]; p.g[i] = grads[i];
  }
  const int nw = 2 * h * d + 3 * h * z + 2 * h + 2 * z + d;
  const int nrow = VAE_RPC * (2 * d + 4 * h + 6 * z);
  const int smem = (nw + nrow) * (int)sizeof(float);
  static int configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(vae_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return lsps_set_error(ctx, LSPS_E_CUDA, "vae_step smem attr: %s", cudaGetErrorString(e));
    configured = smem;
  }
  vae_step_kernel<<<(rows + VAE_RPC - 1) / VAE_RPC, 256, smem, ST_(st)>>>(p);
  LSPS_CHECK_LAUNCH(ctx, "vae_step");
  return LSPS_OK;
}

extern "C" int lsps_adam_ex(lsps_ctx* ctx, float* p, const float* g, float* m, float* v, void* w16, void* w16_lo,
                            long long n, float lr, float beta1, float beta2, float eps, float wd, int step,
                            float grad_scale, const float* hyper, lsps_stream st);
extern "C" int lsps_adam(lsps_ctx* ctx, float* p, const float* g, float* m, float* v, void* w16, long long n, float lr,
                         float beta1, float beta2, float eps, float wd, int step, float grad_scale, const float* hyper,
                         lsps_stream st) {
  return lsps_adam_ex(ctx, p, g, m, v, w16, nullptr, n, lr, beta1, beta2, eps, wd, step, grad_scale, hyper, st);
}
extern "C" int lsps_adam_ex(lsps_ctx* ctx, float* p, const float* g, float* m, float* v, void* w16, void* w16_lo,
                            long long n, float lr, float beta1, float beta2, float eps, float wd, int step,
                            float grad_scale, const float* hyper, lsps_stream st) {
  REQUIRE(ctx, p && g && m && v && n > 0 && step > 0, LSPS_E_ARG, "adam: arg");
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const int vec = al16(p) && al16(g) && al16(m) && al16(v) && (reinterpret_cast<uintptr_t>(w16) & 7) == 0 &&
                  (reinterpret_cast<uintptr_t>(w16_lo) & 7) == 0;
  adam_kernel<<<grid_for(vec ? (n + 3) / 4 : n, 256, 8 * ctx->num_sms), 256, 0, ST_(st)>>>(
      p, g, m, v, static_cast<bf16*>(w16), static_cast<bf16*>(w16_lo), n, (float)(lr / bc1), beta1, beta2, eps, wd,
      (float)(1.0 / sqrt(bc2)), grad_scale, hyper, vec);
  LSPS_CHECK_LAUNCH(ctx, "adam");
  return LSPS_OK;
}
extern "C" int lsps_pack_dgrad(lsps_ctx* ctx, const float* w, void* wt, int taps, int cout, int cin, lsps_stream st) {
  REQUIRE(ctx, w && wt && taps > 0 && cout > 0 && cin > 0, LSPS_E_ARG, "pack_dgrad: arg");
  pack_dgrad_kernel<<<dim3((cin + 31) / 32, (cout + 31) / 32, taps), dim3(32, 8), 0, ST_(st)>>>(w, static_cast<bf16*>(wt), cout, cin);
  LSPS_CHECK_LAUNCH(ctx, "pack_dgrad");
  return LSPS_OK;
}
extern "C" int lsps_pack_dgrad_multi(lsps_ctx* ctx, const float* w_base, void* wt_base, void* wt_lo_base,
                                     const long long* desc, int count, int total_tiles, lsps_stream st) {
  REQUIRE(ctx, w_base && wt_base && desc && count > 0 && total_tiles > 0, LSPS_E_ARG, "pack_dgrad_multi: arg");
  const int grid = total_tiles < 16 * ctx->num_sms ? total_tiles : 16 * ctx->num_sms;
  pack_dgrad_multi_kernel<<<grid, dim3(32, 8), 0, ST_(st)>>>(w_base, static_cast<bf16*>(wt_base),
                                                             static_cast<bf16*>(wt_lo_base), desc, count, total_tiles);
  LSPS_CHECK_LAUNCH(ctx, "pack_dgrad_multi");
  return LSPS_OK;
}
extern "C" int lsps_f32_split_bf16(lsps_ctx* ctx, const float* x, void* hi, void* lo, long long n, lsps_stream st) {
  REQUIRE(ctx, x && hi && lo && n > 0, LSPS_E_ARG, "f32_split_bf16: arg");
  f32_split_bf16_kernel<<<grid_for(n, 256, 8 * ctx->num_sms), 256, 0, ST_(st)>>>(x, static_cast<bf16*>(hi),
                                                                               static_cast<bf16*>(lo), n);
  LSPS_CHECK_LAUNCH(ctx, "f32_split_bf16");
  return LSPS_OK;
}
extern "C" int lsps_f32_to_bf16(lsps_ctx* ctx, const float* x, void* y, long long n, lsps_stream st) {
  REQUIRE(ctx, x && y && n > 0, LSPS_E_ARG, "f32_to_bf16: arg");
  f32_to_bf16_kernel<<<grid_for(n, 256, 8 * ctx->num_sms), 256, 0, ST_(st)>>>(x, static_cast<bf16*>(y), n);
  LSPS_CHECK_LAUNCH(ctx, "f32_to_bf16");
  return LSPS_OK;
}
extern "C" int lsps_bf16_to_f32(lsps_ctx* ctx, const void* x, float* y, long long n, lsps_stream st) {
  REQUIRE(ctx, x && y && n > 0, LSPS_E_ARG, "bf16_to_f32: arg");
  bf16_to_f32_kernel<<<grid_for(n, 256, 8 * ctx->num_sms), 256, 0, ST_(st)>>>(static_cast<const bf16*>(x), y, n);
  LSPS_CHECK_LAUNCH(ctx, "bf16_to_f32");
  return LSPS_OK;
}

extern "C" int lsps_joint_errors(lsps_ctx* ctx, const float* pred, const float* gt, const int* joint_idx, int nj, int j3,
                                 float sx, float sy, float sz, float* err_mean, float* err_max, int n, lsps_stream st) {
  REQUIRE(ctx, pred && gt && err_mean && err_max && n > 0 && nj > 0 && j3 >= 3 * nj, LSPS_E_ARG, "joint_errors: arg");
  joint_errors_kernel<<<(n + 127) / 128, 128, 0, ST_(st)>>>(pred, gt, joint_idx, nj, j3, sx, sy, sz, err_mean, err_max, n);
  LSPS_CHECK_LAUNCH(ctx, "joint_errors");
  return LSPS_OK;
}
