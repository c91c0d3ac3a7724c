// Implicit-GEMM convolutions on the 5th-generation tensor cores (tcgen05 + TMEM), operands staged by TMA.
//
//   conv_igemm_kernel : y[pixels, Cout] = sum over taps, Cin of x[pixel shifted by tap, Cin] * W[tap][Cout][Cin]
//                       covers Conv2d 3x3 s1/s2 forward, their data gradients and the stride-2 ConvTranspose2d
//                       (forward = 4 sub-pixel phases, no zero insertion, no col2im atomics).
//   wgrad_kernel      : dW[tap][Cout, Cin] += sum over pixels dy[pixel, Cout]^T * x[pixel shifted by tap, Cin]
//                       (both operands MN-major straight out of NHWC, split-K over pixel tiles, fp32 red.add).
//
// There is no im2col buffer: every (tap, 64-channel chunk) K-step is ONE 5-D TMA box out of the NHWC activation
// tensor; TMA's out-of-bounds zero fill *is* the conv padding.  Stride-2 access never uses elementStrides: the
// tensor is viewed as [N][H/2][2][W/2][2C] ("pair view") so that a stride-2 tap is a plain box at a parity index.
//
// Replaces the cuDNN calls under nn.Conv2d / nn.ConvTranspose2d at
//   /root/reference/src/trainers/common_net.py:160-181 (LeakyINSResBlock), :246-268 (LeakyReLUConv2d / ConvTranspose2d)
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

using namespace lsps;

// ------------------------------------------------------------------------------------------------ ctx
int lsps_set_error(lsps_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" int lsps_abi_version(void) { return 1; }

extern "C" int lsps_ctx_create(lsps_ctx** out, int device) {
  if (!out) return LSPS_E_ARG;
  *out = nullptr;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return LSPS_E_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LSPS_E_CUDA;
  if (prop.major != 10) return LSPS_E_ARCH;  // sm_100a only: no fallback path exists
  lsps_ctx* ctx = new lsps_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    delete ctx;
    return LSPS_E_CUDA;
  }
  ctx->encode_fn = fn;
  *out = ctx;
  return LSPS_OK;
}
extern "C" void lsps_ctx_destroy(lsps_ctx* ctx) { delete ctx; }
extern "C" const char* lsps_last_error(lsps_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
extern "C" long long lsps_launch_count(lsps_ctx* ctx) { return ctx ? ctx->launches : 0; }

int lsps_get_tmap(lsps_ctx* ctx, const void* ptr, int rank, const uint32_t* dims, const uint32_t* box,
                  CUtensorMap* out) {
  TmapKey key{};
  key.ptr = ptr;
  key.rank = rank;
  for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
  auto it = ctx->tmaps.find(key);
  if (it != ctx->tmaps.end()) { *out = it->second; return LSPS_OK; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return lsps_set_error(ctx, LSPS_E_ARG, "tensor not 16-byte aligned");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  uint64_t stride = 2;  // bf16
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    stride *= dims[i];
    if (i < rank - 1) gstr[i] = stride;
  }
  CUtensorMap tm;
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_fn)(
      &tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gdim, gstr, bdim, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return lsps_set_error(ctx, LSPS_E_CUDA, "cuTensorMapEncodeTiled failed (%d) rank %d dims %u %u %u %u %u box %u %u %u %u %u",
                          (int)r, rank, dims[0], dims[1], rank > 2 ? dims[2] : 0, rank > 3 ? dims[3] : 0,
                          rank > 4 ? dims[4] : 0, box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0,
                          rank > 4 ? box[4] : 0);
  if (ctx->tmaps.size() > 4096) ctx->tmaps.clear();
  ctx->tmaps.emplace(key, tm);
  *out = tm;
  return LSPS_OK;
}

// ------------------------------------------------------------------------------------------------ plans
struct Tap {       // coordinate offsets of one filter tap in the A-operand tensor map
  short ac, ax, ap, ay;  // channel offset (pair view: parity of x), x, row-parity index, y
  int brow;              // first row of this tap's [N][K] weight slice
};
struct Phase { short tap0, ntaps, oa, ob; };  // output sub-pixel phase (oa, ob) and its tap list
struct IgemmParams {
  int tiles_x, tiles_y, tiles_i, tiles_n, nphases;
  int twl, thl, nb;   // M tile = 2^twl x 2^thl pixels x nb images = 128 rows
  int txl, tyl;       // log2(tiles_x), log2(tiles_y)
  int nimg, kchunks;  // kchunks = K channels / 64
  int ntaps_all;      // filter taps in the table (9, or 16 for the 4x4 transposed conv)
  // two weight sets in one launch (encoder A | encoder B res blocks on the concatenated batch): images >= nsplit use
  // the rows brow1.. of the weight tensor map and the second bias vector; nsplit = 0 -> one set, brow0 = brow1 = 0
  int nsplit, brow0, brow1;
  const float* bias2;
  long long o_n, o_y, o_x;  // output strides (elements)
  int o_sy, o_sx;           // output pixel = (y*o_sy + oa, x*o_sx + ob)
  __nv_bfloat16* out;
  const float* bias;
  const __nv_bfloat16* mask;
  const __nv_bfloat16* add;
  float slope;
  int flags;
  int dbg;  // profiling experiments only: 1 = skip MMA issue, 2 = skip TMA issue (results are garbage)
  // LSPS_EP_STATS: per-(image, channel) sum / sum of squares of the fp32 result (after bias), accumulated with red.add
  //   into sums[nimg][2][nc_total] -- the InstanceNorm / BatchNorm statistics of the NEXT layer, taken while the values
  //   are still in registers (one streaming apply pass is all that is left of the norm).
  float* sums;
  // LSPS_EP_INBWD (data gradient that lands on a = lrelu(xhat), xhat = IN(h)): `mask` points at a, from which
  //   xhat = a > 0 ? a : a / slope is recovered (no statistics loads in the epilogue); the epilogue stores
  //   g = acc * lrelu'(xhat) and accumulates sum g, sum g*xhat into bsums[nimg][2][nc_total]
  float* bsums;
  float inv_slope;
  int nc_total;     // output channels of the GEMM (row pitch of sums / in_stats / bsums, lo-half offset of split outputs)
  // split-bf16 ("bf16x3") operands: A tensor channels are [hi | lo] (lo half a_lo channels in), the weights come as two
  // tensors (tmB hi, tmBlo lo); K-steps per (tap, chunk): hi*hi, hi*lo, lo*hi; the output is written as [hi | lo] too
  int split, a_lo, kch_eff;
  // split_fused (KCH = 2 kernels): ONE stage per (tap, chunk) carries {A hi, A lo} and {B hi, B lo}; the MMA warp issues
  // hi*hi, hi*lo, lo*hi from it -- one pipeline handshake and four TMA boxes instead of three handshakes and six boxes
  int split_fused;
  int a_group;      // grouped conv: N tile nt reads A channels [nt * a_group, (nt + 1) * a_group); 0 = dense
  int gpb;          // multi-phase launches: M groups per block of the block-major tile order (0 = phase-major)
  Phase ph[4];
  Tap taps[16];
};

constexpr int A_STAGE_BYTES = 128 * 128;  // 128 rows x 64 bf16

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile --
// each CTA stages its own 128 rows of A and HALF of the weight tile, one MMA of M=256 reads both halves, so the
// shared-memory fill + read traffic per MAC drops by a third and the pipeline gets 6 stages instead of 4.
// KCH = 64-channel chunks per pipeline stage: every stage costs one barrier round trip and one tcgen05.commit (a drain
// bubble in the tensor pipe), so the MMA-bound shapes stage 128 channels (8 MMAs) per handshake instead of 64 (4 MMAs).
template <int BN, int CG, int KCH = 1>
struct IgemmCfg {
  static constexpr int B_CHUNK_BYTES = (BN / CG) * 128;
  static constexpr int B_STAGE_BYTES = KCH * B_CHUNK_BYTES;
  static constexpr int A_STAGE = KCH * A_STAGE_BYTES;
  static constexpr int STAGE_BYTES = A_STAGE + B_STAGE_BYTES;
  static constexpr int STAGES = (196608 / STAGE_BYTES) > 8 ? 8 : (196608 / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int BIAS_BYTES = 2048 * 4;  // whole bias vector (Cout <= 2048) staged once per CTA
  static constexpr int KTAB_BYTES = 1024 * 16;     // K-step table (<= 16 taps x 32 chunks, or 9 taps x 3 x 32 split)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + BIAS_BYTES + KTAB_BYTES;
};

// tile index -> (phase, n tile, m tile, pixel origin); the M grid dims are powers of two.
// Multi-phase (2x up-sampling) launches walk the M groups in blocks of gpb groups (a few images): all phases of one
// block before the next block, so the block's input (read by every phase) is fetched from DRAM once and then hit in L2,
// while concurrently running CTAs still sit in the same phase (no interleaved partial-line writes between neighbours).
struct TileCoord { int pi, nt, mt, x0, y0, ti; };
__device__ __forceinline__ void decode_phase(const IgemmParams& p, int t, int per_phase, int groups_m, int& pi, int& r,
                                             int& gcount, int& gbase) {
  pi = 0; r = t; gcount = groups_m; gbase = 0;
  if (p.nphases > 1) {
    if (p.gpb > 0) {
      const int blk_tiles = p.nphases * p.gpb * p.tiles_n;
      const int blk = t / blk_tiles;
      const int rb = t - blk * blk_tiles;
      gbase = blk * p.gpb;
      gcount = groups_m - gbase < p.gpb ? groups_m - gbase : p.gpb;
      pi = rb / (gcount * p.tiles_n);
      r = rb - pi * gcount * p.tiles_n;
    } else {
      pi = t / per_phase; r = t - pi * per_phase;
    }
  }
}
__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int t, int per_phase, int groups_m, int cg, int rank) {
  TileCoord c;
  int r, gcount, gbase;
  decode_phase(p, t, per_phase, groups_m, c.pi, r, gcount, gbase);
  c.nt = 0;
  int mg = r;
  if (p.tiles_n > 1) { c.nt = r / gcount; mg = r - c.nt * gcount; }
  mg += gbase;
  c.mt = mg * cg + rank;
  const int tx = c.mt & (p.tiles_x - 1), ty = (c.mt >> p.txl) & (p.tiles_y - 1);
  c.ti = c.mt >> (p.txl + p.tyl);
  c.x0 = tx << p.twl; c.y0 = ty << p.thl;
  return c;
}

// Sum over the 32 lanes of v[j] for every j, the total for index j landing in lane j: a butterfly that halves the number
// of live values each step (16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 32 x 5).
__device__ __forceinline__ float lane_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

// Epilogue role (4 warps): TMEM -> registers -> (+bias, statistics, LeakyReLU, +residual gradient, *lrelu'(mask) or the
// InstanceNorm-backward front half) -> bf16 (or split hi|lo) -> global.
// EG = epilogue warp groups.  EG == 2 (short-K layers, where one tile's MMAs are done before four warps have drained
// the previous tile): warps 2-5 take the tiles of accumulator buffer 0, warps 6-9 those of buffer 1, so every group has
// two tile periods per epilogue; this variant compiles the plain bias / LeakyReLU / mask / add (+ split-output) epilogue
// only (the statistics and InstanceNorm-backward paths stay in the EG == 1 kernels, which keeps it under 204 registers).
template <int BN, int CG, int EG>
__device__ __forceinline__ void epilogue_role(const IgemmParams& p, const float* sbias, uint32_t tmem_base,
                                              uint64_t* tfull, uint64_t* tempty, int warp, int lane, int rank,
                                              int worker, int nworkers, int groups_m, int per_phase, int total) {
    constexpr bool FULL = EG == 1;
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int egroup = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int tw_mask = (1 << p.twl) - 1, th_mask = (1 << p.thl) - 1;
    const int xl = row & tw_mask, yl = (row >> p.twl) & th_mask, nl = row >> (p.twl + p.thl);
    const bool inbwd = FULL && (p.flags & LSPS_EP_INBWD) != 0;
    int it = 0;
    for (int t = worker; t < total; t += nworkers, ++it) {
      if (EG == 2 && (it & 1) != egroup) continue;
      const TileCoord tc = decode_tile(p, t, per_phase, groups_m, CG, rank);
      const int nt = tc.nt, x0 = tc.x0, y0 = tc.y0, n = tc.ti * p.nb + nl;
      const Phase P = p.ph[tc.pi];
      const bool valid = n < p.nimg;
      const int boff = (p.nsplit && n >= p.nsplit) ? p.tiles_n * BN : 0;   // second bias vector
      const long long off = (long long)n * p.o_n + (long long)((y0 + yl) * p.o_sy + P.oa) * p.o_y +
                            (long long)((x0 + xl) * p.o_sx + P.ob) * p.o_x + nt * BN;
      const int acc = it & 1; const uint32_t accph = (it >> 1) & 1;
      const bool use_add = valid && (p.flags & LSPS_EP_ADD), use_mask = valid && ((p.flags & LSPS_EP_MASK) || inbwd);
      const uint4* a4 = reinterpret_cast<const uint4*>(p.add + off);
      const uint4* m4 = reinterpret_cast<const uint4*>(p.mask + off);
      uint4* o4 = reinterpret_cast<uint4*>(p.out + off);
      // per-(image, channel) rows of the statistics buffers (every row of a warp lies in ONE image: host-checked)
      const long long srow = ((long long)(valid ? n : 0) * 2) * p.nc_total + nt * BN;
      // software pipeline: the global loads (residual / mask) and the TMEM load of chunk c+1 are in flight while
      // chunk c is converted and stored; TMEM is handed back to the MMA warp as soon as its last column is in registers
      uint4 ga[4], gm[4];
      if (use_add) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ga[j] = __ldg(a4 + j);
      }
      if (use_mask) {
#pragma unroll
        for (int j = 0; j < 4; ++j) gm[j] = __ldg(m4 + j);
      }
      mbar_wait(&tfull[acc], accph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      uint32_t v[32];
      tmem_ld32(taddr, v);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float f[32];
        uint4 ca[4], cm[4];
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) { ca[j] = ga[j]; cm[j] = gm[j]; }
        if (c0 + 32 < BN) {
          tmem_ld32(taddr + c0 + 32, v);
          if (use_add) {
#pragma unroll
            for (int j = 0; j < 4; ++j) ga[j] = __ldg(a4 + (c0 + 32) / 8 + j);
          }
          if (use_mask) {
#pragma unroll
            for (int j = 0; j < 4; ++j) gm[j] = __ldg(m4 + (c0 + 32) / 8 + j);
          }
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (CG == 2) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
        }
        if (p.flags & LSPS_EP_BIAS) {
          const float4* b4 = reinterpret_cast<const float4*>(sbias + boff + nt * BN + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = b4[j];
            f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
          }
        }
        if (FULL && (p.flags & LSPS_EP_STATS)) {   // warp-uniform branch; invalid (phantom) rows contribute zeros
          float s[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) s[j] = valid ? f[j] : 0.f;
          const float s1 = lane_transpose_sum(s, lane);
#pragma unroll
          for (int j = 0; j < 32; ++j) s[j] = valid ? f[j] * f[j] : 0.f;
          const float s2 = lane_transpose_sum(s, lane);
          if (valid) {
            red_add_f32(p.sums + srow + c0 + lane, s1);
            red_add_f32(p.sums + srow + p.nc_total + c0 + lane, s2);
          }
        }
        if (p.flags & LSPS_EP_LRELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = f[j] > 0.f ? f[j] : f[j] * p.slope;
        }
        if (p.flags & LSPS_EP_ADD) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t w[4] = {ca[j].x, ca[j].y, ca[j].z, ca[j].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              f[8 * j + 2 * k] += bf16lo(w[k]);
              f[8 * j + 2 * k + 1] += bf16hi(w[k]);
            }
          }
        }
        if (p.flags & LSPS_EP_MASK) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t w[4] = {cm[j].x, cm[j].y, cm[j].z, cm[j].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (!(bf16lo(w[k]) > 0.f)) f[8 * j + 2 * k] *= p.slope;
              if (!(bf16hi(w[k]) > 0.f)) f[8 * j + 2 * k + 1] *= p.slope;
            }
          }
        }
        if (inbwd) {
          // f = dL/da with a = lrelu(xhat) in cm.  g = f * lrelu'(xhat); per-(image, channel) sums of g and g*xhat
          float gx[32];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t w[4] = {cm[j].x, cm[j].y, cm[j].z, cm[j].w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float av = (k & 1) ? bf16hi(w[k >> 1]) : bf16lo(w[k >> 1]);
              const bool pos = av > 0.f;
              const float xh = pos ? av : av * p.inv_slope;
              float g = pos ? f[8 * j + k] : f[8 * j + k] * p.slope;
              g = valid ? g : 0.f;
              f[8 * j + k] = g;
              gx[8 * j + k] = g * xh;
            }
          }
          const float s2 = lane_transpose_sum(gx, lane);
#pragma unroll
          for (int j = 0; j < 32; ++j) gx[j] = f[j];
          const float s1 = lane_transpose_sum(gx, lane);
          if (valid) {
            red_add_f32(p.bsums + srow + c0 + lane, s1);
            red_add_f32(p.bsums + srow + p.nc_total + c0 + lane, s2);
          }
        }
        if (valid) {
          if (p.split) {
            // 16 mantissa bits: hi = bf16(f), lo = bf16(f - hi); the lo half sits nc_total channels further
            uint4* o4l = reinterpret_cast<uint4*>(p.out + off + p.nc_total);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float a = f[8 * j + 2 * k], b = f[8 * j + 2 * k + 1];
                hi[k] = pack_bf16x2(a, b);
                lo[k] = pack_bf16x2(a - bf16lo(hi[k]), b - bf16hi(hi[k]));
              }
              o4[c0 / 8 + j] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              o4l[c0 / 8 + j] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_bf16x2(f[8 * j], f[8 * j + 1]);
              o.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
              o.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
              o.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
              o4[c0 / 8 + j] = o;
            }
          }
        }
      }
    }
}

// warp 0: TMA producer | warp 1: TMEM owner + MMA issuer | warps 2-5: epilogue (TMEM -> regs -> global)
template <int BN, int CG, int KCH, int EG>
__global__ void __launch_bounds__(64 + 128 * EG, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ IgemmParams p) {
  using Cfg = IgemmCfg<BN, CG, KCH>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer: LDS/STS, not generic LD/ST
  uint8_t* sA = base;
  uint8_t* sB = base + STAGES * Cfg::A_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(base + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sbias = reinterpret_cast<float*>(base + STAGES * Cfg::STAGE_BYTES + 256);
  int4* ktab = reinterpret_cast<int4*>(sbias + 2048);   // flattened K-steps: {A channel coord, x | y<<16, parity | kcol<<4, weight row}

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CG == 2 ? (int)cluster_ctarank() : 0;  // position in the CTA pair
  if (p.flags & LSPS_EP_BIAS) {
    const int nc_total = p.tiles_n * BN;
    for (int i = threadIdx.x; i < nc_total; i += blockDim.x) sbias[i] = p.bias[i];
    if (p.nsplit)
      for (int i = threadIdx.x; i < nc_total; i += blockDim.x) sbias[nc_total + i] = p.bias2[i];
  }
  // the producer is ONE thread whose per-stage latency bounds the whole pipeline: no divisions or parameter-space
  // reads in its loop -- every (tap, 64-channel chunk) K-step is a precomputed 16-byte table entry
  // split-bf16: three K-steps per (tap, chunk) -- A hi x B hi, A hi x B lo (bit 3 of .z), A lo (a_lo channels in) x B hi
  for (int i = threadIdx.x; i < p.ntaps_all * p.kch_eff; i += blockDim.x) {
    const int tp = i / p.kch_eff, r = i - tp * p.kch_eff;
    const int per = p.split ? (p.split_fused ? 2 : 3) : 1;       // table entries per (tap, chunk)
    const int kc = r / per, v = r - kc * per;
    const bool a_lo = p.split_fused ? v == 1 : v == 2;           // fused: entry 1 = (A lo, B lo) of the stage
    const Tap T = p.taps[tp];
    ktab[i] = make_int4(kc * 64 + T.ac + (a_lo ? p.a_lo : 0), (T.ax & 0xFFFF) | (T.ay << 16),
                        T.ap | (v == 1 ? 8 : 0) | ((kc * 64) << 4), T.brow);
  }
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4 * CG); }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.split) tma_prefetch_desc(&tmBlo);
  }
  if (warp == 1) {
    if (CG == 2) { tmem_alloc_cg2(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish_cg2(); }
    else { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work items: (phase, n tile, group of CG consecutive m tiles); a phantom m tile (odd count) loads zeros, stores nothing
  const int tiles_m = p.tiles_x * p.tiles_y * p.tiles_i;
  const int groups_m = (tiles_m + CG - 1) / CG;
  const int per_phase = groups_m * p.tiles_n;
  const int total = per_phase * p.nphases;
  const int worker = blockIdx.x / CG, nworkers = gridDim.x / CG;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t ph = 0;
      for (int t = worker; t < total; t += nworkers) {
        const TileCoord tc = decode_tile(p, t, per_phase, groups_m, CG, rank);
        const int nt = tc.nt, x0 = tc.x0, y0 = tc.y0, n0 = tc.ti * p.nb;
        const Phase P = p.ph[tc.pi];
        // K-steps (tap, 64-channel chunk) are flattened; a stage carries up to KCH consecutive ones
        const int nks = P.ntaps * p.kch_eff;
        const int4* kt = ktab + P.tap0 * p.kch_eff;
        const int a_goff = nt * p.a_group;
        const int brow_off = nt * BN + rank * (BN / 2) * (CG - 1) + ((p.nsplit && n0 >= p.nsplit) ? p.brow1 : p.brow0);
        for (int i0 = 0; i0 < nks; i0 += KCH) {
          const int cnt = nks - i0 < KCH ? nks - i0 : KCH;
          mbar_wait(&empty[stage], ph ^ 1);
          if (p.dbg & 2) {
            if (leader) mbar_arrive(&full[stage]);
          } else {
            if (leader) mbar_expect_tx(&full[stage], CG * cnt * (A_STAGE_BYTES + Cfg::B_CHUNK_BYTES));
#pragma unroll
            for (int j = 0; j < KCH; ++j) {
              if (j < cnt) {
                const int4 e = kt[i0 + j];
                const int ax = (short)(e.y & 0xFFFF), ay = e.y >> 16;
                uint8_t* da = sA + stage * Cfg::A_STAGE + j * A_STAGE_BYTES;
                uint8_t* db = sB + stage * Cfg::B_STAGE_BYTES + j * Cfg::B_CHUNK_BYTES;
                const int ap = e.z & 7, kcol = e.z >> 4;
                const CUtensorMap* tb = (e.z & 8) ? &tmBlo : &tmB;
                if (CG == 2) {
                  tma_load_5d_cg2(da, &tmA, &full[stage], e.x + a_goff, x0 + ax, ap, y0 + ay, n0);
                  tma_load_2d_cg2(db, tb, &full[stage], kcol, e.w + brow_off);
                } else {
                  tma_load_5d(da, &tmA, &full[stage], e.x + a_goff, x0 + ax, ap, y0 + ay, n0);
                  tma_load_2d(db, tb, &full[stage], kcol, e.w + brow_off);
                }
              }
            }
          }
          if (++stage == STAGES) { stage = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(128 * CG, BN, 0, 0);
      constexpr uint64_t dbase = umma_desc_base(0, 1024);
      const uint32_t sA_u32 = smem_u32(sA), sB_u32 = smem_u32(sB);
      int stage = 0; uint32_t ph = 0; int it = 0;
      for (int t = worker; t < total; t += nworkers, ++it) {
        int pi, r_, gc_, gb_;
        decode_phase(p, t, per_phase, groups_m, pi, r_, gc_, gb_);
        const int nks = p.ph[pi].ntaps * p.kch_eff;
        const int acc = it & 1; const uint32_t accph = (it >> 1) & 1;
        mbar_wait(&tempty[acc], accph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int i0 = 0; i0 < nks; i0 += KCH) {
          const int cnt = nks - i0 < KCH ? nks - i0 : KCH;
          mbar_wait(&full[stage], ph);
          tc_fence_after();
          const uint64_t a_base = dbase | ((sA_u32 + stage * Cfg::A_STAGE) >> 4);
          const uint64_t b_base = dbase | ((sB_u32 + stage * Cfg::B_STAGE_BYTES) >> 4);
          if (KCH == 2 && p.split_fused) {
#pragma unroll
            for (int v = 0; v < 3; ++v) {          // hi*hi, hi*lo, lo*hi
              const uint64_t av = a_base + (v == 2 ? (A_STAGE_BYTES >> 4) : 0);
              const uint64_t bv = b_base + (v == 1 ? (Cfg::B_CHUNK_BYTES >> 4) : 0);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (CG == 2) umma_bf16_cg2(d_tmem, av + k * 2, bv + k * 2, idesc, (i0 | v | k) != 0 ? 1u : 0u);
                else umma_bf16(d_tmem, av + k * 2, bv + k * 2, idesc, (i0 | v | k) != 0 ? 1u : 0u);
              }
            }
          } else
#pragma unroll
          for (int k = 0; k < 4 * KCH; ++k) {
            if ((p.dbg & 1) || (k >> 2) >= cnt) break;
            // +32 B per K=16 slice (2 in 16-byte units); next 64-channel chunk one operand tile further
            const uint64_t ad = a_base + ((k >> 2) * (A_STAGE_BYTES >> 4) + (k & 3) * 2);
            const uint64_t bd = b_base + ((k >> 2) * (Cfg::B_CHUNK_BYTES >> 4) + (k & 3) * 2);
            if (CG == 2) umma_bf16_cg2(d_tmem, ad, bd, idesc, (i0 | k) != 0 ? 1u : 0u);
            else umma_bf16(d_tmem, ad, bd, idesc, (i0 | k) != 0 ? 1u : 0u);
          }
          if (CG == 2) umma_commit_cg2(&empty[stage]); else umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; ph ^= 1; }
        }
        if (CG == 2) umma_commit_cg2(&tfull[acc]); else umma_commit(&tfull[acc]);
      }
    }
  } else {
    epilogue_role<BN, CG, EG>(p, sbias, tmem_base, tfull, tempty, warp, lane, rank, worker, nworkers, groups_m, per_phase, total);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) { if (CG == 2) tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// ------------------------------------------------------------------------------------------------ 2x up-sampling, N = 64
// ConvTranspose2d(3x3, s2, p1, op1) forward and the data gradient of Conv2d(3x3, s2, p1) write a 2x finer grid: output
// pixel (2i+a, 2j+b) ("phase" (a,b)) is a sum over 1/2/2/4 of the 9 taps of input pixels (i+dy, j+dx), dy,dx in {0,1}.
// The generic kernel runs the four phases as four GEMMs (9 A-tile loads and 9 weight-tile loads per 64-channel chunk,
// one pipeline handshake each, the input re-read from DRAM once per phase).  For the two widest such layers of the
// generator (128 -> 64 channels at 64x64 -> 128x128, N = 64) that is handshake- and L2-bound (ncu: 13-17 % tensor pipe,
// 2 GB of L2->SM traffic for 0.4 GB of algorithmic bytes).  Here ONE CTA computes all four phases of its 128 input
// pixels: the whole weight tensor (9 taps x K <= 128 x 64 = 144 KB) is loaded into shared memory once per CTA, each
// of the 4 shifted A tiles is loaded once per chunk and feeds every tap that uses it (4/2/2/1 MMAs groups), the four
// 128x64 accumulators live side by side in TMEM (256 columns, double-buffered).
struct Up64Params {
  int tiles_x, tiles_y, tiles_i, twl, thl, nb, txl, tyl, nimg, kchunks;
  long long o_n, o_y, o_x;     // OUTPUT strides (elements); output pixel = (2y + a, 2x + b)
  __nv_bfloat16* out;
  const float* bias;
  const __nv_bfloat16* mask;
  float slope;
  int flags;
  int nc;                      // = 64: weight rows per tap
  // fused decoder head (1x1 transposed conv 64 -> 1 + tanh, optional L1 term): see lsps_conv_ext
  const float* head_w; const float* head_b; float* head_out;
  const float* head_target; long long head_t0, head_tn; float head_scale; float* head_dout; float* head_acc;
};
constexpr int UP64_W_TILE = 64 * 128;          // one (tap, chunk) weight tile: 64 rows x 64 bf16
constexpr int UP64_A_SLOTS = 3;
// taps of each A shift s = dy*2 + dx: (tap index r*3+c, phase a*2+b); see t2_axis()
__constant__ int c_up64_ntaps[4] = {4, 2, 2, 1};
__constant__ int c_up64_tap[4][4] = {{4, 5, 7, 8}, {3, 6, 0, 0}, {1, 2, 0, 0}, {0, 0, 0, 0}};
__constant__ int c_up64_phase[4][4] = {{0, 1, 2, 3}, {1, 3, 0, 0}, {2, 3, 0, 0}, {3, 0, 0, 0}};

__global__ void __launch_bounds__(320, 1)
conv_up64_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ Up64Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer: LDS/STS, not generic LD/ST
  const int wtiles = 9 * p.kchunks;
  uint8_t* sW = base;                                   // [tap][chunk] weight tiles, resident
  uint8_t* sA = base + 18 * UP64_W_TILE;                // ring of A tiles
  uint8_t* sS = sA + UP64_A_SLOTS * A_STAGE_BYTES;      // output staging: one 128 x 64 tile per epilogue group
  uint64_t* full = reinterpret_cast<uint64_t*>(sS + 2 * A_STAGE_BYTES);
  uint64_t* empty = full + UP64_A_SLOTS;
  uint64_t* wfull = empty + UP64_A_SLOTS;
  uint64_t* tfull = wfull + 1;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + 128);     // 16-byte aligned (float4 loads)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((p.flags & LSPS_EP_BIAS) && threadIdx.x < 64) sbias[threadIdx.x] = p.bias[threadIdx.x];
  float* shw = sbias + 64;                               // head weights
  if (p.head_out && threadIdx.x >= 64 && threadIdx.x < 128) shw[threadIdx.x - 64] = p.head_w[threadIdx.x - 64];
  if (threadIdx.x == 0) {
    for (int i = 0; i < UP64_A_SLOTS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(wfull, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total = p.tiles_x * p.tiles_y * p.tiles_i;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wfull, wtiles * UP64_W_TILE);
      for (int tap = 0; tap < 9; ++tap)
        for (int kc = 0; kc < p.kchunks; ++kc)
          tma_load_2d(sW + (tap * p.kchunks + kc) * UP64_W_TILE, &tmB, wfull, kc * 64, tap * p.nc);
      int slot = 0; uint32_t ph = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int tx = t & (p.tiles_x - 1), ty = (t >> p.txl) & (p.tiles_y - 1), ti = t >> (p.txl + p.tyl);
        const int x0 = tx << p.twl, y0 = ty << p.thl, n0 = ti * p.nb;
        for (int kc = 0; kc < p.kchunks; ++kc)
          for (int sft = 0; sft < 4; ++sft) {
            mbar_wait(&empty[slot], ph ^ 1);
            mbar_expect_tx(&full[slot], A_STAGE_BYTES);
            tma_load_5d(sA + slot * A_STAGE_BYTES, &tmA, &full[slot], kc * 64, x0 + (sft & 1), 0, y0 + (sft >> 1), n0);
            if (++slot == UP64_A_SLOTS) { slot = 0; ph ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint64_t dbase = umma_desc_base(0, 1024);
      const uint32_t sA_u32 = smem_u32(sA), sW_u32 = smem_u32(sW);
      mbar_wait(wfull, 0);
      tc_fence_after();
      int slot = 0; uint32_t ph = 0; int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int acc = it & 1; const uint32_t accph = (it >> 1) & 1;
        mbar_wait(&tempty[acc], accph ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * 256;
        for (int kc = 0; kc < p.kchunks; ++kc)
          for (int sft = 0; sft < 4; ++sft) {
            mbar_wait(&full[slot], ph);
            tc_fence_after();
            const uint64_t a_base = dbase | ((sA_u32 + slot * A_STAGE_BYTES) >> 4);
            const int nt = c_up64_ntaps[sft];
            for (int j = 0; j < nt; ++j) {
              const uint64_t b_base = dbase | ((sW_u32 + (c_up64_tap[sft][j] * p.kchunks + kc) * UP64_W_TILE) >> 4);
              const uint32_t d = d0 + c_up64_phase[sft][j] * 64;
#pragma unroll
              for (int k = 0; k < 4; ++k)      // the first chunk's shift-0 taps open the four accumulators
                umma_bf16(d, a_base + k * 2, b_base + k * 2, idesc, (kc | sft | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty[slot]);
            if (++slot == UP64_A_SLOTS) { slot = 0; ph ^= 1; }
          }
        umma_commit(&tfull[acc]);
      }
    }
  } else {
    // two epilogue warp groups (warps 2-5: accumulator set 0, warps 6-9: set 1): a tile's 8 TMEM chunks take longer
    // than its 36-72 MMAs, so each group gets two tile periods.  Every phase's 128 x 64 result is staged in shared memory
    // (swizzled like an operand tile) and leaves as ONE TMA tensor store through the pair view of the output -- per-thread
    // 16-byte global stores (32 different lines per warp instruction) kept the LSU queue full (ncu r02: long-scoreboard
    // stalls at the top of the chunk loop).
    const int q = warp & 3;
    const int egroup = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const bool issuer = q == 0 && lane == 0;
    uint8_t* stage = sS + egroup * A_STAGE_BYTES;
    const int tw_mask = (1 << p.twl) - 1, th_mask = (1 << p.thl) - 1;
    const int xl = row & tw_mask, yl = (row >> p.twl) & th_mask, nl = row >> (p.twl + p.thl);
    const float head_b = p.head_out ? __ldg(p.head_b) : 0.f;
    float head_l1 = 0.f;
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      if ((it & 1) != egroup) continue;
      const int tx = t & (p.tiles_x - 1), ty = (t >> p.txl) & (p.tiles_y - 1), ti = t >> (p.txl + p.tyl);
      const int x0 = tx << p.twl, y0 = ty << p.thl, n0 = ti * p.nb, n = n0 + nl;
      const bool valid = n < p.nimg;
      const int acc = it & 1; const uint32_t accph = (it >> 1) & 1;
      mbar_wait(&tfull[acc], accph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
#pragma unroll 1
      for (int phs = 0; phs < 4; ++phs) {
        if (issuer) bulk_wait_read();             // the previous store of this group has read the staging tile
        named_bar_sync(1 + egroup, 128);
        float hs = 0.f;                           // head: dot product of this pixel's 64 stored channels
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int c0 = hf * 32;
          const long long off = (long long)n * p.o_n + (long long)((y0 + yl) * 2 + (phs >> 1)) * p.o_y +
                                (long long)((x0 + xl) * 2 + (phs & 1)) * p.o_x + c0;
          uint4 gm[4];
          if (valid && (p.flags & LSPS_EP_MASK)) {
            const uint4* m4 = reinterpret_cast<const uint4*>(p.mask + off);
#pragma unroll
            for (int j = 0; j < 4; ++j) gm[j] = __ldg(m4 + j);
          }
          uint32_t v[32];
          tmem_ld32(taddr + phs * 64 + c0, v);
          tmem_ld_wait();
          if (phs == 3 && hf == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.flags & LSPS_EP_BIAS) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = *reinterpret_cast<const float4*>(sbias + c0 + 4 * j);
              f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
            }
          }
          if (p.flags & LSPS_EP_LRELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = f[j] > 0.f ? f[j] : f[j] * p.slope;
          }
          if (valid && (p.flags & LSPS_EP_MASK)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t w[4] = {gm[j].x, gm[j].y, gm[j].z, gm[j].w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (!(bf16lo(w[k]) > 0.f)) f[8 * j + 2 * k] *= p.slope;
                if (!(bf16hi(w[k]) > 0.f)) f[8 * j + 2 * k + 1] *= p.slope;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(f[8 * j], f[8 * j + 1]);
            o.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
            o.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
            o.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
            *reinterpret_cast<uint4*>(stage + row * 128 + (((hf * 4 + j) ^ (row & 7)) << 4)) = o;
            if (p.head_out) {
              const float4 w0 = *reinterpret_cast<const float4*>(shw + c0 + 8 * j);
              const float4 w1 = *reinterpret_cast<const float4*>(shw + c0 + 8 * j + 4);
              hs += bf16lo(o.x) * w0.x + bf16hi(o.x) * w0.y + bf16lo(o.y) * w0.z + bf16hi(o.y) * w0.w +
                    bf16lo(o.z) * w1.x + bf16hi(o.z) * w1.y + bf16lo(o.w) * w1.z + bf16hi(o.w) * w1.w;
            }
          }
        }
        if (p.head_out && valid) {
          const int oy = (y0 + yl) * 2 + (phs >> 1), ox = (x0 + xl) * 2 + (phs & 1);
          const long long px = ((long long)n * (p.tiles_y << (p.thl + 1)) + oy) * (p.tiles_x << (p.twl + 1)) + ox;
          const float ov = tanhf(hs + head_b);
          p.head_out[px] = ov;
          if (p.head_target && px >= p.head_t0 && px < p.head_t0 + p.head_tn) {
            const float df = ov - __ldg(p.head_target + (px - p.head_t0));
            head_l1 += fabsf(df);
            if (p.head_dout) p.head_dout[px - p.head_t0] = df > 0.f ? p.head_scale : (df < 0.f ? -p.head_scale : 0.f);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        named_bar_sync(1 + egroup, 128);
        if (issuer) {                             // images past nimg (last group of a partial tile) are clipped by the map
          tma_store_5d(&tmO, stage, (phs & 1) * 64, x0, phs >> 1, y0, n0);
          bulk_commit();
        }
      }
    }
    if (issuer) bulk_wait_read();
    if (p.head_target) {                          // one atomic per warp
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) head_l1 += __shfl_xor_sync(0xffffffffu, head_l1, off);
      if (lane == 0) atomicAdd(p.head_acc, head_l1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}
constexpr int UP64_SMEM = 18 * UP64_W_TILE + (UP64_A_SLOTS + 2) * A_STAGE_BYTES + 1024 + 1024;

// ------------------------------------------------------------------------------------------------ resident weights, N = 128
// 3x3 layers with 64 input channels and 128 output channels (Conv2d(64,128,3,2,1) forward, the data gradient of
// ConvTranspose2d(128,64,3,2,1,1)) write 32 KB and read 9 x 16 KB of activations per 128-pixel tile; in the generic
// kernel every tile also re-fetches the 9 x 16 KB weight tensor, so L2 -> SM traffic is twice what the activations need
// and that, not HBM, bounds it (ncu r02: 0.59 of the HBM bound, 0.77x cuDNN).  Here the whole weight tensor (<= 9
// K-steps x [128 x 64] = 144 KB) is loaded into shared memory once per CTA; the pipeline stages carry A tiles only.
// Same tile order, accumulator hand-off and epilogue (two warp groups) as conv_igemm_kernel<128, 1, 1, 2>.
constexpr int RESB_W_TILE = 128 * 128;         // one K-step's weight tile: 128 rows x 64 bf16
constexpr int RESB_A_SLOTS = 5;
constexpr int RESB_SMEM = 9 * RESB_W_TILE + RESB_A_SLOTS * A_STAGE_BYTES + 1024 + 256 + 512;

__global__ void __launch_bounds__(320, 1)
conv_resb_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer: LDS/STS, not generic LD/ST
  uint8_t* sW = base;                                   // [K-step] weight tiles, resident
  uint8_t* sA = base + 9 * RESB_W_TILE;                 // ring of A tiles
  uint64_t* full = reinterpret_cast<uint64_t*>(sA + RESB_A_SLOTS * A_STAGE_BYTES);
  uint64_t* empty = full + RESB_A_SLOTS;
  uint64_t* wfull = empty + RESB_A_SLOTS;
  uint64_t* tfull = wfull + 1;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((p.flags & LSPS_EP_BIAS) && threadIdx.x < 128) sbias[threadIdx.x] = p.bias[threadIdx.x];
  if (threadIdx.x == 0) {
    for (int i = 0; i < RESB_A_SLOTS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(wfull, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total = p.tiles_x * p.tiles_y * p.tiles_i;  // one phase, one N tile
  const int ntaps = p.ph[0].ntaps;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wfull, ntaps * RESB_W_TILE);
      for (int tp = 0; tp < ntaps; ++tp) tma_load_2d(sW + tp * RESB_W_TILE, &tmB, wfull, 0, p.taps[tp].brow);
      int slot = 0; uint32_t ph = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t, total, total, 1, 0);
        const int n0 = tc.ti * p.nb;
        for (int tp = 0; tp < ntaps; ++tp) {
          const Tap T = p.taps[tp];
          mbar_wait(&empty[slot], ph ^ 1);
          mbar_expect_tx(&full[slot], A_STAGE_BYTES);
          tma_load_5d(sA + slot * A_STAGE_BYTES, &tmA, &full[slot], T.ac, tc.x0 + T.ax, T.ap, tc.y0 + T.ay, n0);
          if (++slot == RESB_A_SLOTS) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint64_t dbase = umma_desc_base(0, 1024);
      const uint32_t sA_u32 = smem_u32(sA), sW_u32 = smem_u32(sW);
      mbar_wait(wfull, 0);
      tc_fence_after();
      int slot = 0; uint32_t ph = 0; int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int acc = it & 1; const uint32_t accph = (it >> 1) & 1;
        mbar_wait(&tempty[acc], accph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 128;
        for (int tp = 0; tp < ntaps; ++tp) {
          mbar_wait(&full[slot], ph);
          tc_fence_after();
          const uint64_t a_base = dbase | ((sA_u32 + slot * A_STAGE_BYTES) >> 4);
          const uint64_t b_base = dbase | ((sW_u32 + tp * RESB_W_TILE) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a_base + k * 2, b_base + k * 2, idesc, (tp | k) != 0 ? 1u : 0u);
          umma_commit(&empty[slot]);
          if (++slot == RESB_A_SLOTS) { slot = 0; ph ^= 1; }
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else {
    epilogue_role<128, 1, 2>(p, sbias, tmem_base, tfull, tempty, warp, lane, 0, blockIdx.x, gridDim.x, total, total, total);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------ wgrad
struct WTap { short mc, mx, mp, my, nc, nx, np, ny; };  // tap offsets in the dy (M side) / x (N side) maps
struct WgradParams {
  int tiles_x, tiles_y, tiles_i;
  int twl, thl, nb;  // K tile = 64 pixels
  int ntaps, co_tiles, ci_tiles, splits;   // co_tiles counts tiles of 128*CG output channels
  int cout, cin;
  int dbg;
  // split-bf16 operands (dy and x stored as [hi | lo] channel halves): three passes per pixel tile into the same
  // accumulator -- dy_hi x x_hi, dy_hi x x_lo (x channels + n_lo), dy_lo (dy channels + m_lo) x x_hi
  int nvar, m_lo, n_lo;
  // grouped conv (gw = group width, 0 = dense): only the diagonal (co tile, ci tile) pairs are computed, the x channels
  // are those of the dy tile, and the epilogue keeps the gw x gw blocks on the diagonal: dw[tap][cout][gw]
  int gw;
  float* dw;
  WTap taps[16];
};
// one TMA box = KPX pixels x 64 channels.  TMA cost is dominated by a fixed per-box overhead, so the CTA-pair kernel
// (2 + 2 boxes per stage) uses 128-pixel boxes; the single-CTA kernel keeps 64 pixels to afford >= 4 stages.
template <int BN, int CG> struct WBox { static constexpr int KPX = (CG == 2 || BN <= 128) ? 128 : 64; static constexpr int BYTES = KPX * 128; };

template <int BN, int CG>
struct WgradCfg {
  static constexpr int NB_BOXES = BN / 64 / CG;                    // x boxes staged by this CTA
  static constexpr int KPX = WBox<BN, CG>::KPX;
  static constexpr int W_BOX_BYTES = WBox<BN, CG>::BYTES;
  static constexpr int STAGE_BYTES = (2 + NB_BOXES) * W_BOX_BYTES;
  static constexpr int STAGES = (196608 / STAGE_BYTES) > 8 ? 8 : (196608 / STAGE_BYTES);
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BN, int CG>
__global__ void __launch_bounds__(192, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmN,
             const __grid_constant__ WgradParams p) {
  using Cfg = WgradCfg<BN, CG>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int W_BOX_BYTES = Cfg::W_BOX_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer: LDS/STS, not generic LD/ST
  uint64_t* full = reinterpret_cast<uint64_t*>(base + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const bool leader = rank == 0;
  int b = blockIdx.x / CG;
  const int split = b % p.splits; b /= p.splits;
  const int cit = b % p.ci_tiles; b /= p.ci_tiles;
  const int cot = (b % p.co_tiles) * CG + rank;   // this CTA's 128-channel tile of dy
  const int nbase = p.gw ? (b % p.co_tiles) * CG * 128 : cit * BN;   // first x channel of the N tile
  const int tap = b / p.co_tiles;
  const int ptiles = p.tiles_x * p.tiles_y * p.tiles_i;
  const int pt0 = (int)((long long)ptiles * split / p.splits);
  const int pt1 = (int)((long long)ptiles * (split + 1) / p.splits);
  if (pt0 >= pt1) return;   // both CTAs of a pair take the same decision
  const int nA = CG == 2 ? 2 : min(2, (p.cout - cot * 128) / 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tfull, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmM);
    tma_prefetch_desc(&tmN);
  }
  if (warp == 1) {
    if (CG == 2) { tmem_alloc_cg2(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish_cg2(); }
    else { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // every box of a stage is issued by its own lane
    const WTap T = p.taps[tap];
    int stage = 0; uint32_t ph = 0;
    const int nboxes = nA + Cfg::NB_BOXES;
    const int niter = (pt1 - pt0) * p.nvar;
    for (int i = 0; i < niter; ++i) {
      const int pt = pt0 + i / p.nvar, var = i - (i / p.nvar) * p.nvar;
      const int m_off = var == 2 ? p.m_lo : 0, n_off = var == 1 ? p.n_lo : 0;
      const int tx = pt % p.tiles_x, ty = (pt / p.tiles_x) % p.tiles_y, ti = pt / (p.tiles_x * p.tiles_y);
      const int x0 = tx << p.twl, y0 = ty << p.thl, n0 = ti * p.nb;
      uint8_t* s = base + stage * Cfg::STAGE_BYTES;
      if (lane == 0) {
        mbar_wait(&empty[stage], ph ^ 1);
        if (p.dbg & 2) { if (leader) mbar_arrive(&full[stage]); }
        else if (CG == 1) mbar_expect_tx(&full[stage], nboxes * W_BOX_BYTES);
        else if (leader) mbar_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
      }
      __syncwarp();
      if (p.dbg & 2) {
      } else if (lane < nA) {
        if (CG == 2) tma_load_5d_cg2(s + lane * W_BOX_BYTES, &tmM, &full[stage], cot * 128 + lane * 64 + T.mc + m_off, x0 + T.mx, T.mp, y0 + T.my, n0);
        else tma_load_5d(s + lane * W_BOX_BYTES, &tmM, &full[stage], cot * 128 + lane * 64 + T.mc + m_off, x0 + T.mx, T.mp, y0 + T.my, n0);
      } else if (lane < nboxes) {
        const int j = lane - nA;
        const int cbase = nbase + (rank * Cfg::NB_BOXES + j) * 64 + T.nc + n_off;
        if (CG == 2) tma_load_5d_cg2(s + (2 + j) * W_BOX_BYTES, &tmN, &full[stage], cbase, x0 + T.nx, T.np, y0 + T.ny, n0);
        else tma_load_5d(s + (2 + j) * W_BOX_BYTES, &tmN, &full[stage], cbase, x0 + T.nx, T.np, y0 + T.ny, n0);
      }
      if (++stage == STAGES) { stage = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(128 * CG, BN, 1, 1);
      int stage = 0; uint32_t ph = 0;
      const int niter = (pt1 - pt0) * p.nvar;
      for (int pt = 0; pt < niter; ++pt) {
        mbar_wait(&full[stage], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(base + stage * Cfg::STAGE_BYTES);
        const uint32_t b_addr = a_addr + 2 * W_BOX_BYTES;
#pragma unroll
        for (int k = 0; k < Cfg::KPX / 16; ++k) {  // 16 pixels per MMA = 16 rows of 128 B
          if (p.dbg & 1) break;
          const uint64_t ad = umma_smem_desc(a_addr + k * 2048, W_BOX_BYTES, 1024);
          const uint64_t bd = umma_smem_desc(b_addr + k * 2048, W_BOX_BYTES, 1024);
          if (CG == 2) umma_bf16_cg2(tmem_base, ad, bd, idesc, (pt > 0 || k > 0) ? 1u : 0u);
          else umma_bf16(tmem_base, ad, bd, idesc, (pt > 0 || k > 0) ? 1u : 0u);
        }
        if (CG == 2) umma_commit_cg2(&empty[stage]); else umma_commit(&empty[stage]);
        if (++stage == STAGES) { stage = 0; ph ^= 1; }
      }
      if (CG == 2) umma_commit_cg2(tfull); else umma_commit(tfull);
    }
  } else {
    const int q = warp & 3;
    const int co = cot * 128 + q * 32 + lane;
    mbar_wait(tfull, 0);
    tc_fence_after();
    float* dst = p.gw ? p.dw + ((long long)tap * p.cout + co) * p.gw - (co / p.gw) * p.gw + nbase
                      : p.dw + ((long long)tap * p.cout + co) * p.cin + cit * BN;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t v[32];
    tmem_ld32(taddr, v);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float f[32];
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      if (c0 + 32 < BN) tmem_ld32(taddr + c0 + 32, v);
      // grouped: x channel nbase + c0 must lie in the group of output channel co (dst then indexes it within the group)
      if (co < p.cout && (p.gw == 0 || (nbase + c0) / p.gw == co / p.gw)) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + 4 * j), "f"(f[4 * j]),
                       "f"(f[4 * j + 1]), "f"(f[4 * j + 2]), "f"(f[4 * j + 3])
                       : "memory");
      }
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) { if (CG == 2) tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// ------------------------------------------------------------------------------------------------ host side
namespace {

struct Geo { int tw, th, nb, twl, thl; };
inline int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }
inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
Geo geo_for(int hg, int wg, int pixels) {
  Geo g;
  g.tw = wg < pixels ? wg : pixels;
  g.th = hg < pixels / g.tw ? hg : pixels / g.tw;
  g.nb = pixels / (g.tw * g.th);
  g.twl = ilog2(g.tw);
  g.thl = ilog2(g.th);
  return g;
}

// activation tensor map: plain view (C, W, 1, H, N) or pair view (2C, W/2, 2, H/2, N); box = (64, tw, 1, th, nb)
int act_tmap(lsps_ctx* ctx, const void* ptr, int n, int h, int w, int c, bool pair, const Geo& g, CUtensorMap* tm) {
  uint32_t dims[5], box[5] = {64, (uint32_t)g.tw, 1, (uint32_t)g.th, (uint32_t)g.nb};
  if (pair) { dims[0] = 2 * c; dims[1] = w / 2; dims[2] = 2; dims[3] = h / 2; dims[4] = n; }
  else      { dims[0] = c;     dims[1] = w;     dims[2] = 1; dims[3] = h;     dims[4] = n; }
  return lsps_get_tmap(ctx, ptr, 5, dims, box, tm);
}

// taps of a stride-2 access along one axis, forward-conv indexing: input index 2*o + r - 1
//   r=0 -> pair o-1, parity 1 ; r=1 -> pair o, parity 0 ; r=2 -> pair o, parity 1
inline void s2_axis(int r, int* d, int* par) { *d = (r == 0) ? -1 : 0; *par = (r == 1) ? 0 : 1; }
// taps of a transposed stride-2 access along one axis for output parity a: list of (r, input offset d)
inline int t2_axis(int a, int* rs, int* ds) {
  if (a == 0) { rs[0] = 1; ds[0] = 0; return 1; }
  rs[0] = 0; ds[0] = 1; rs[1] = 2; ds[1] = 0; return 2;
}
// 4x4 stride-2 pad-1 transposed conv (output index 2*i - 1 + r).  Gather form along one axis: row 2*o - 1 + r of the
// fine tensor = pair o + d, parity par
inline void s4_axis(int r, int* d, int* par) { *d = r == 0 ? -1 : (r == 3 ? 1 : 0); *par = (r == 0 || r == 2) ? 1 : 0; }
// scatter form: output parity a gets taps (r, coarse offset d): a=0 -> (1,0),(3,-1) ; a=1 -> (2,0),(0,+1)
inline int t4_axis(int a, int* rs, int* ds) {
  if (a == 0) { rs[0] = 1; ds[0] = 0; rs[1] = 3; ds[1] = -1; }
  else { rs[0] = 2; ds[0] = 0; rs[1] = 0; ds[1] = 1; }
  return 2;
}

enum Dir { FWD = 0, DGRAD = 1 };

inline int lsps_force_cg() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LSPS_FORCE_CG"); v = e ? atoi(e) : 0; }
  return v;
}
inline bool lsps_no_kch2() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LSPS_NO_KCH2"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
inline int lsps_dbg() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LSPS_DBG"); v = e ? atoi(e) : 0; }
  return v;
}
inline bool lsps_phase_major() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LSPS_PHASE_MAJOR"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
inline bool lsps_one_epi_group() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LSPS_ONE_EPI_GROUP"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
inline bool lsps_no_small_bn() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LSPS_NO_SMALL_BN"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
inline bool lsps_no_up64() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LSPS_NO_UP64"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
inline bool lsps_no_split_fuse() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LSPS_NO_SPLIT_FUSE"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
inline bool lsps_no_resb() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LSPS_NO_RESB"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
inline bool lsps_use_pairs() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LSPS_NO_PAIRS"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

template <typename K, typename... Args>
cudaError_t launch_maybe_cluster(K kernel, int grid, int smem, int cg, cudaStream_t st, int threads, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cg; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <int BN, int CG, int KCH = 1, int EG = 1>
int launch_igemm(lsps_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBlo, const IgemmParams& p,
                 cudaStream_t st) {
  using Cfg = IgemmCfg<BN, CG, KCH>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<BN, CG, KCH, EG>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return lsps_set_error(ctx, LSPS_E_CUDA, "igemm smem attr: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int tiles_m = p.tiles_x * p.tiles_y * p.tiles_i;
  const int total = ((tiles_m + CG - 1) / CG) * p.tiles_n * p.nphases;
  const int workers = total < ctx->num_sms / CG ? total : ctx->num_sms / CG;
  cudaError_t e = launch_maybe_cluster(conv_igemm_kernel<BN, CG, KCH, EG>, workers * CG, Cfg::SMEM_BYTES, CG, st, 64 + 128 * EG,
                                       tmA, tmB, tmBlo, p);
  if (e != cudaSuccess) return lsps_set_error(ctx, LSPS_E_CUDA, "conv_igemm launch: %s", cudaGetErrorString(e));
  LSPS_CHECK_LAUNCH(ctx, "conv_igemm");
  return LSPS_OK;
}

// Builds the plan for forward / data-gradient of any of the three conv kinds and launches it.
//   in : the tensor the GEMM reads (x for FWD, dy for DGRAD); out: what it writes (y / dx)
int run_igemm(lsps_ctx* ctx, const lsps_conv_shape* s, Dir dir, const void* in, const void* wpk, const float* bias,
              void* out, const void* mask, const void* add, int flags, float slope, cudaStream_t st,
              const lsps_conv_ext* ext) {
  if (!ctx || !s || !in || !wpk || !out) return lsps_set_error(ctx, LSPS_E_ARG, "null argument");
  static const lsps_conv_ext no_ext{};
  if (!ext) ext = &no_ext;
  const void* wpk2 = ext->w2;
  const float* bias2 = ext->bias2;
  const int nsplit = ext->n_split;
  const bool split = ext->split != 0;
  if (split && (!ext->w_lo || nsplit > 0 || (flags & (LSPS_EP_ADD | LSPS_EP_STATS | LSPS_EP_INBWD))))
    return lsps_set_error(ctx, LSPS_E_ARG, "split-bf16 conv: needs w_lo; not combinable with grouped / add / stats");
  if ((flags & LSPS_EP_STATS) && !ext->sums) return lsps_set_error(ctx, LSPS_E_ARG, "stats flag without sums");
  if ((flags & LSPS_EP_INBWD) && (!ext->in_a || !ext->bsums || (flags & LSPS_EP_MASK) || !(slope > 0.f)))
    return lsps_set_error(ctx, LSPS_E_ARG, "inbwd flag needs in_a, bsums, a positive slope and excludes mask");
  if (flags & LSPS_EP_INBWD) mask = ext->in_a;
  const int kind = s->kind, n = s->n, h = s->h, w = s->w, cin = s->cin, cout = s->cout;
  if (kind < 0 || kind > 4 || n <= 0 || !is_pow2(h) || !is_pow2(w) || cin % 64 || cout % 64 || cin <= 0 || cout <= 0)
    return lsps_set_error(ctx, LSPS_E_SHAPE, "conv shape kind %d n %d h %d w %d cin %d cout %d", kind, n, h, w, cin, cout);
  const bool k1 = kind == LSPS_CONV1X1;
  if (kind != LSPS_CONV_S1 && !k1 && (h < 2 || w < 2)) return lsps_set_error(ctx, LSPS_E_SHAPE, "stride-2 op needs h,w >= 2");
  // grouped 3x3 stride-1 conv (ResNeXt, common_net.py:118): every N tile is one group and reads only its own K channels
  const int groups = ext->groups > 1 ? ext->groups : 1;
  const int gw = cin / groups;
  if (groups > 1 && (kind != LSPS_CONV_S1 || cin != cout || cin % groups || (gw != 64 && gw != 128) || split || nsplit))
    return lsps_set_error(ctx, LSPS_E_SHAPE, "grouped conv: 3x3 stride 1, cin == cout, group width 64 or 128 (got %d)", gw);
  if (cin > 2048 || cout > 2048) return lsps_set_error(ctx, LSPS_E_SHAPE, "channels > 2048 (bias / K-step tables are sized for 2048)");
  if ((flags & LSPS_EP_BIAS) && !bias) return lsps_set_error(ctx, LSPS_E_ARG, "bias flag without bias");
  if ((flags & LSPS_EP_MASK) && !mask) return lsps_set_error(ctx, LSPS_E_ARG, "mask flag without mask");
  if ((flags & LSPS_EP_ADD) && !add) return lsps_set_error(ctx, LSPS_E_ARG, "add flag without add");

  // forward-op output dims
  const int ho = (kind == LSPS_CONV_S1 || k1) ? h : (kind == LSPS_CONV_S2 ? h / 2 : 2 * h);
  const int wo = (kind == LSPS_CONV_S1 || k1) ? w : (kind == LSPS_CONV_S2 ? w / 2 : 2 * w);
  const bool k4 = kind == LSPS_DECONV4_S2;
  const int ks = k4 ? 4 : (k1 ? 1 : 3);
  // GEMM dims: K channels (of `in`; per group), N channels (of `out`)
  const int kten = dir == FWD ? cin : cout;            // channels of the `in` tensor
  const int kc = groups > 1 ? gw : kten, nc = dir == FWD ? cout : cin;
  // `in` / `out` tensor dims
  const int ih = dir == FWD ? h : ho, iw = dir == FWD ? w : wo;
  const int oh = dir == FWD ? ho : h, ow = dir == FWD ? wo : w;
  // classify the access pattern
  //   plain  : out grid == in grid, 9 shifted taps                       (S1 fwd, S1 dgrad)
  //   down   : out grid = in grid / 2, pair view on `in`                 (S2 fwd, DECONV dgrad)
  //   up     : out grid = 2 * in grid, 4 phases, strided store           (DECONV fwd, S2 dgrad)
  const bool plain = kind == LSPS_CONV_S1 || k1;
  const bool down = (kind == LSPS_CONV_S2 && dir == FWD) || ((kind == LSPS_DECONV_S2 || k4) && dir == DGRAD);
  const int hg = down ? ih / 2 : ih, wg = down ? iw / 2 : iw;  // GEMM pixel grid
  const Geo g = geo_for(hg, wg, 128);

  IgemmParams p{};
  p.tiles_x = wg / g.tw; p.tiles_y = hg / g.th; p.tiles_i = (n + g.nb - 1) / g.nb;
  p.twl = g.twl; p.thl = g.thl; p.nb = g.nb;
  p.txl = ilog2(p.tiles_x); p.tyl = ilog2(p.tiles_y);
  p.nimg = n; p.kchunks = kc / 64; p.ntaps_all = ks * ks;
  // fused split stages pay on the wide, shallow layers (K <= 256 channels: 64->128 data gradient 256 -> 213 us, forward
  // 134 -> 106 us); with more K chunks per tap the 2-3 stage ring of the 64 KB+ stages loses (1024->2048: 195 -> 211 us)
  const bool fuse = split && kc <= 256 && !lsps_no_split_fuse() && !lsps_one_epi_group();
  p.split = split ? 1 : 0; p.split_fused = fuse ? 1 : 0; p.a_lo = kc;
  p.kch_eff = split ? (fuse ? 2 : 3) * p.kchunks : p.kchunks; p.nc_total = nc;
  if (p.ntaps_all * p.kch_eff > 1024) return lsps_set_error(ctx, LSPS_E_SHAPE, "K-step table overflow (%d steps)", p.ntaps_all * p.kch_eff);
  p.sums = ext->sums; p.bsums = ext->bsums; p.inv_slope = slope > 0.f ? 1.f / slope : 0.f;
  const int kct = split ? 2 * kten : kten, nct = split ? 2 * nc : nc;     // channels of the `in` / `out` TENSORS
  p.a_group = groups > 1 ? gw : 0;
  if ((flags & (LSPS_EP_STATS | LSPS_EP_INBWD)) && hg * wg < 32)
    return lsps_set_error(ctx, LSPS_E_SHAPE, "fused statistics need >= 32 GEMM pixels per image");
  p.o_n = (long long)oh * ow * nct; p.o_y = (long long)ow * nct; p.o_x = nct;
  p.out = static_cast<__nv_bfloat16*>(out); p.bias = bias;
  p.mask = static_cast<const __nv_bfloat16*>(mask); p.add = static_cast<const __nv_bfloat16*>(add);
  p.slope = slope; p.flags = flags; p.dbg = lsps_dbg();

  int nt = 0;
  if (k1) {
    p.nphases = 1; p.o_sy = p.o_sx = 1;
    p.ph[0] = Phase{0, 1, 0, 0};
    p.taps[nt++] = Tap{0, 0, 0, 0, 0};
  } else if (plain) {
    p.nphases = 1; p.o_sy = p.o_sx = 1;
    p.ph[0] = Phase{0, 9, 0, 0};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        Tap& T = p.taps[nt++];
        T.ac = 0; T.ap = 0;
        T.ay = dir == FWD ? r - 1 : 1 - r;
        T.ax = dir == FWD ? c - 1 : 1 - c;
        T.brow = (r * 3 + c) * nc;
      }
  } else if (down) {
    p.nphases = 1; p.o_sy = p.o_sx = 1;
    p.ph[0] = Phase{0, (short)(ks * ks), 0, 0};
    for (int r = 0; r < ks; ++r)
      for (int c = 0; c < ks; ++c) {
        Tap& T = p.taps[nt++];
        int dy, py, dx, px;
        if (k4) { s4_axis(r, &dy, &py); s4_axis(c, &dx, &px); }
        else { s2_axis(r, &dy, &py); s2_axis(c, &dx, &px); }
        T.ay = dy; T.ap = py; T.ax = dx; T.ac = px * kct;
        T.brow = (r * ks + c) * nc;
      }
  } else {  // up: heaviest phase first
    p.nphases = 4; p.o_sy = p.o_sx = 2;
    const int order[4][2] = {{1, 1}, {1, 0}, {0, 1}, {0, 0}};
    for (int i = 0; i < 4; ++i) {
      const int a = order[i][0], b = order[i][1];
      int rs[2], dys[2], cs[2], dxs[2];
      const int nr = k4 ? t4_axis(a, rs, dys) : t2_axis(a, rs, dys), ncs = k4 ? t4_axis(b, cs, dxs) : t2_axis(b, cs, dxs);
      p.ph[i] = Phase{(short)nt, (short)(nr * ncs), (short)a, (short)b};
      for (int ri = 0; ri < nr; ++ri)
        for (int ci = 0; ci < ncs; ++ci) {
          Tap& T = p.taps[nt++];
          T.ac = 0; T.ap = 0; T.ay = dys[ri]; T.ax = dxs[ci];
          T.brow = (rs[ri] * ks + cs[ci]) * nc;
        }
    }
  }

  int bn = groups > 1 ? gw : (nc % 256 == 0 ? 256 : (nc % 128 == 0 ? 128 : 64));
  // few tiles (small batches: the estimate-mode generator pass on 8 source images is 64 tiles of 128 x 256): halve the N
  // tile so that twice as many SMs share the launch -- the K loop per CTA stays, its weight tile and epilogue halve
  if (bn == 256 && groups == 1 && !lsps_no_small_bn() &&
      (long long)p.tiles_x * p.tiles_y * p.tiles_i * p.nphases * (nc / 256) * 2 <= ctx->num_sms)
    bn = 128;
  p.tiles_n = nc / bn;
  // grouped launch: one tensor map over both weight sets (they live in one flat buffer)
  const char* wbase = static_cast<const char*>(wpk);
  long long wrows = (long long)ks * ks * nc;
  if (nsplit > 0) {
    const long long delta = static_cast<const char*>(wpk2) - static_cast<const char*>(wpk);   // bytes
    const long long row_bytes = 2LL * kc;
    if (!wpk2 || nsplit >= n || nsplit % g.nb || delta == 0 || delta % row_bytes || 2 * nc > 2048 ||
        ((flags & LSPS_EP_BIAS) && !bias2))
      return lsps_set_error(ctx, LSPS_E_ARG, "grouped conv: n_split %d / weight distance %lld not usable", nsplit, delta);
    const long long drows = (delta < 0 ? -delta : delta) / row_bytes;
    if (drows + wrows > 0x7fffffffLL) return lsps_set_error(ctx, LSPS_E_ARG, "grouped conv: weight sets too far apart");
    p.nsplit = nsplit; p.bias2 = bias2;
    p.brow0 = delta < 0 ? (int)drows : 0;
    p.brow1 = delta < 0 ? 0 : (int)drows;
    if (delta < 0) wbase = static_cast<const char*>(wpk2);
    wrows += drows;
  }
  // CTA pairs (cta_group::2) whenever there are at least two M tiles to pair up
  const int tiles_m = p.tiles_x * p.tiles_y * p.tiles_i;
  //   ... and the tile is MMA-bound (>= 18 K-steps: 3x3 taps x 128 input channels and up) and there are enough tiles to
  //   keep every SM busy in pairs; shorter-K layers are issue-/latency-bound per tile and run faster as 148 independent
  //   CTAs (measured: r01 probes; r02 microbench with LSPS_FORCE_CG=2: 128->256 stride 2 +9 %, 64->128 stride 2 -8 %)
  int nk_min = 1 << 30;
  for (int i = 0; i < p.nphases; ++i) nk_min = p.ph[i].ntaps * p.kch_eff < nk_min ? p.ph[i].ntaps * p.kch_eff : nk_min;
  const int force = lsps_force_cg();
  int cg = (lsps_use_pairs() && tiles_m >= 2 && nk_min >= 18 && (long long)tiles_m * p.tiles_n >= 2 * ctx->num_sms) ? 2 : 1;
  if (force && tiles_m >= 2) cg = force;
  // the two N = 64, K <= 128 up-sampling layers of the generator: all four phases in one pass, resident weights
  if (!plain && !down && !k4 && nc == 64 && p.kchunks <= 2 && !split && nsplit == 0 && tiles_m >= 1 &&
      !(flags & (LSPS_EP_ADD | LSPS_EP_STATS | LSPS_EP_INBWD)) && !lsps_no_up64()) {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(conv_up64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UP64_SMEM);
      if (e != cudaSuccess) return lsps_set_error(ctx, LSPS_E_CUDA, "up64 smem attr: %s", cudaGetErrorString(e));
      configured = true;
    }
    Up64Params u{};
    u.tiles_x = p.tiles_x; u.tiles_y = p.tiles_y; u.tiles_i = p.tiles_i; u.twl = p.twl; u.thl = p.thl; u.nb = p.nb;
    u.txl = p.txl; u.tyl = p.tyl; u.nimg = n; u.kchunks = p.kchunks;
    u.o_n = p.o_n; u.o_y = p.o_y; u.o_x = p.o_x; u.out = p.out; u.bias = bias; u.mask = p.mask; u.slope = slope;
    u.flags = flags; u.nc = nc;
    if (ext->head_out) {
      if (!ext->head_w || !ext->head_b || (ext->head_target && !ext->head_acc))
        return lsps_set_error(ctx, LSPS_E_ARG, "fused head: head_w / head_b (and head_acc with a target) required");
      u.head_w = ext->head_w; u.head_b = ext->head_b; u.head_out = ext->head_out; u.head_target = ext->head_target;
      u.head_t0 = ext->head_t0; u.head_tn = ext->head_tn; u.head_scale = ext->head_scale; u.head_dout = ext->head_dout;
      u.head_acc = ext->head_acc;
    }
    CUtensorMap tA, tB, tO;
    int rc2 = act_tmap(ctx, in, n, ih, iw, kct, false, g, &tA);
    if (rc2) return rc2;
    uint32_t wd2[2] = {(uint32_t)kc, (uint32_t)(9 * nc)}, wb2[2] = {64, 64};
    if ((rc2 = lsps_get_tmap(ctx, wpk, 2, wd2, wb2, &tB))) return rc2;
    if ((rc2 = act_tmap(ctx, out, n, 2 * ih, 2 * iw, 64, true, g, &tO))) return rc2;   // output: pair view, phase = parity
    const int grid = tiles_m < ctx->num_sms ? tiles_m : ctx->num_sms;
    conv_up64_kernel<<<grid, 320, UP64_SMEM, st>>>(tA, tB, tO, u);
    LSPS_CHECK_LAUNCH(ctx, "conv_up64");
    return LSPS_OK;
  }
  if (ext->head_out)
    return lsps_set_error(ctx, LSPS_E_ARG, "fused head: only the fused up-sampling kernel (forward, Cout 64, Cin <= 128) has it");
  // 64 -> 128 channels, one phase, <= 9 K-steps: weights resident in shared memory (conv_resb_kernel)
  if (p.nphases == 1 && bn == 128 && nc == 128 && p.tiles_n == 1 && p.kchunks == 1 && p.ntaps_all <= 9 && !split &&
      nsplit == 0 && p.a_group == 0 && !(flags & (LSPS_EP_STATS | LSPS_EP_INBWD)) && !lsps_no_resb() && !lsps_one_epi_group()) {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(conv_resb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RESB_SMEM);
      if (e != cudaSuccess) return lsps_set_error(ctx, LSPS_E_CUDA, "resb smem attr: %s", cudaGetErrorString(e));
      configured = true;
    }
    CUtensorMap tA, tB;
    int rc2 = act_tmap(ctx, in, n, ih, iw, kct, down, g, &tA);
    if (rc2) return rc2;
    uint32_t wd2[2] = {(uint32_t)kc, (uint32_t)wrows}, wb2[2] = {64, 128};
    if ((rc2 = lsps_get_tmap(ctx, wbase, 2, wd2, wb2, &tB))) return rc2;
    const int grid = tiles_m < ctx->num_sms ? tiles_m : ctx->num_sms;
    conv_resb_kernel<<<grid, 320, RESB_SMEM, st>>>(tA, tB, p);
    LSPS_CHECK_LAUNCH(ctx, "conv_resb");
    return LSPS_OK;
  }
  if (p.nphases > 1 && !lsps_phase_major()) {
    // ~512 tiles per phase and block: enough to fill every SM twice, small enough that the block's input stays in L2
    const int tiles_img = p.tiles_x * p.tiles_y;
    int imgs = (512 + tiles_img - 1) / tiles_img;
    if (imgs < 1) imgs = 1;
    const long long groups = ((long long)imgs * tiles_img) / g.nb / cg;     // tiles_i counts groups of nb images
    p.gpb = groups < 1 ? 1 : (int)groups;
  }
  CUtensorMap tmA, tmB, tmBlo;
  int rc = act_tmap(ctx, in, n, ih, iw, kct, down, g, &tmA);
  if (rc) return rc;
  uint32_t wd[2] = {(uint32_t)kc, (uint32_t)wrows}, wb[2] = {64, (uint32_t)(bn / cg)};
  rc = lsps_get_tmap(ctx, wbase, 2, wd, wb, &tmB);
  if (rc) return rc;
  tmBlo = tmB;
  if (split && (rc = lsps_get_tmap(ctx, ext->w_lo, 2, wd, wb, &tmBlo))) return rc;
  // the statistics accumulators are zeroed here, on the same stream, ahead of the kernel that red.adds into them
  const size_t sbytes = (size_t)n * 2 * nc * sizeof(float);
  if ((flags & LSPS_EP_STATS) && cudaMemsetAsync(ext->sums, 0, sbytes, st) != cudaSuccess)
    return lsps_set_error(ctx, LSPS_E_CUDA, "memset sums");
  if ((flags & LSPS_EP_INBWD) && cudaMemsetAsync(ext->bsums, 0, sbytes, st) != cudaSuccess)
    return lsps_set_error(ctx, LSPS_E_CUDA, "memset bsums");
  if (fuse) {   // split-bf16 with fused stages: the 128-channel-per-stage (KCH = 2) kernels
    if (cg == 2) {
      if (bn == 256) return launch_igemm<256, 2, 2>(ctx, tmA, tmB, tmBlo, p, st);
      if (bn == 128) return launch_igemm<128, 2, 2>(ctx, tmA, tmB, tmBlo, p, st);
      return launch_igemm<64, 2, 2>(ctx, tmA, tmB, tmBlo, p, st);
    }
    if (bn == 256) return launch_igemm<256, 1, 2>(ctx, tmA, tmB, tmBlo, p, st);
    if (bn == 128) return launch_igemm<128, 1, 2, 2>(ctx, tmA, tmB, tmBlo, p, st);
    return launch_igemm<64, 1, 2, 2>(ctx, tmA, tmB, tmBlo, p, st);
  }
  if (cg == 2 && bn == 256 && !lsps_no_kch2()) return launch_igemm<256, 2, 2>(ctx, tmA, tmB, tmBlo, p, st);
  if (cg == 2) {
    if (bn == 256) return launch_igemm<256, 2>(ctx, tmA, tmB, tmBlo, p, st);
    if (bn == 128) return launch_igemm<128, 2>(ctx, tmA, tmB, tmBlo, p, st);
    return launch_igemm<64, 2>(ctx, tmA, tmB, tmBlo, p, st);
  }
  if (bn == 256) return launch_igemm<256, 1>(ctx, tmA, tmB, tmBlo, p, st);
  // single-CTA BN <= 128 kernels stage one 64-channel chunk per pipeline stage (2 and 3 chunks measured slower in r01,
  // 128-channel stages re-measured in r02 with the two epilogue groups: transposed 256->128 forward 0.107 -> 0.105 ms, nil);
  // with a plain epilogue they run two epilogue warp groups (these short-K tiles are epilogue-bound with one)
  const bool light = !(flags & (LSPS_EP_STATS | LSPS_EP_INBWD)) && !lsps_one_epi_group();
  if (bn == 128) return light ? launch_igemm<128, 1, 1, 2>(ctx, tmA, tmB, tmBlo, p, st) : launch_igemm<128, 1>(ctx, tmA, tmB, tmBlo, p, st);
  return light ? launch_igemm<64, 1, 1, 2>(ctx, tmA, tmB, tmBlo, p, st) : launch_igemm<64, 1>(ctx, tmA, tmB, tmBlo, p, st);
}

template <int BN, int CG>
int launch_wgrad(lsps_ctx* ctx, const CUtensorMap& tmM, const CUtensorMap& tmN, const WgradParams& p, cudaStream_t st) {
  using Cfg = WgradCfg<BN, CG>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<BN, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return lsps_set_error(ctx, LSPS_E_CUDA, "wgrad smem attr: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int grid = p.ntaps * p.co_tiles * p.ci_tiles * p.splits * CG;
  cudaError_t e = launch_maybe_cluster(wgrad_kernel<BN, CG>, grid, Cfg::SMEM_BYTES, CG, st, 192, tmM, tmN, p);
  if (e != cudaSuccess) return lsps_set_error(ctx, LSPS_E_CUDA, "wgrad launch: %s", cudaGetErrorString(e));
  LSPS_CHECK_LAUNCH(ctx, "wgrad");
  return LSPS_OK;
}

}  // namespace

constexpr int FWD_FLAGS = LSPS_EP_BIAS | LSPS_EP_LRELU | LSPS_EP_STATS;
constexpr int DGRAD_FLAGS = LSPS_EP_MASK | LSPS_EP_ADD | LSPS_EP_INBWD;

extern "C" int lsps_conv_fwd(lsps_ctx* ctx, const lsps_conv_shape* s, const void* x, const void* w_fwd,
                             const float* bias, void* y, int flags, float slope, lsps_stream st) {
  return run_igemm(ctx, s, FWD, x, w_fwd, bias, y, nullptr, nullptr, flags & (LSPS_EP_BIAS | LSPS_EP_LRELU), slope,
                   static_cast<cudaStream_t>(st), nullptr);
}

extern "C" int lsps_conv_dgrad(lsps_ctx* ctx, const lsps_conv_shape* s, const void* dy, const void* w_dgrad, void* dx,
                               const void* mask, const void* add, int flags, float slope, lsps_stream st) {
  return run_igemm(ctx, s, DGRAD, dy, w_dgrad, nullptr, dx, mask, add, flags & (LSPS_EP_MASK | LSPS_EP_ADD), slope,
                   static_cast<cudaStream_t>(st), nullptr);
}

extern "C" int lsps_conv_fwd_ex(lsps_ctx* ctx, const lsps_conv_shape* s, const void* x, const void* w_fwd,
                                const float* bias, void* y, int flags, float slope, const lsps_conv_ext* ext,
                                lsps_stream st) {
  return run_igemm(ctx, s, FWD, x, w_fwd, bias, y, nullptr, nullptr, flags & FWD_FLAGS, slope,
                   static_cast<cudaStream_t>(st), ext);
}

extern "C" int lsps_conv_dgrad_ex(lsps_ctx* ctx, const lsps_conv_shape* s, const void* dy, const void* w_dgrad, void* dx,
                                  const void* mask, const void* add, int flags, float slope, const lsps_conv_ext* ext,
                                  lsps_stream st) {
  return run_igemm(ctx, s, DGRAD, dy, w_dgrad, nullptr, dx, mask, add, flags & DGRAD_FLAGS, slope,
                   static_cast<cudaStream_t>(st), ext);
}

extern "C" int lsps_conv_fwd_grouped(lsps_ctx* ctx, const lsps_conv_shape* s, const void* x, const void* w_fwd,
                                     const float* bias, const void* w_fwd2, const float* bias2, int n_split, void* y,
                                     int flags, float slope, lsps_stream st) {
  if (n_split <= 0 || !w_fwd2) return lsps_set_error(ctx, LSPS_E_ARG, "conv_fwd_grouped: n_split / second weight set");
  lsps_conv_ext ext{};
  ext.w2 = w_fwd2; ext.bias2 = bias2; ext.n_split = n_split;
  return run_igemm(ctx, s, FWD, x, w_fwd, bias, y, nullptr, nullptr, flags & (LSPS_EP_BIAS | LSPS_EP_LRELU), slope,
                   static_cast<cudaStream_t>(st), &ext);
}

extern "C" int lsps_conv_dgrad_grouped(lsps_ctx* ctx, const lsps_conv_shape* s, const void* dy, const void* w_dgrad,
                                       const void* w_dgrad2, int n_split, void* dx, const void* mask, const void* add,
                                       int flags, float slope, lsps_stream st) {
  if (n_split <= 0 || !w_dgrad2) return lsps_set_error(ctx, LSPS_E_ARG, "conv_dgrad_grouped: n_split / second weight set");
  lsps_conv_ext ext{};
  ext.w2 = w_dgrad2; ext.n_split = n_split;
  return run_igemm(ctx, s, DGRAD, dy, w_dgrad, nullptr, dx, mask, add, flags & (LSPS_EP_MASK | LSPS_EP_ADD), slope,
                   static_cast<cudaStream_t>(st), &ext);
}

static int run_wgrad(lsps_ctx* ctx, const lsps_conv_shape* s, const void* x, const void* dy, float* dw, int split,
                     cudaStream_t st, int groups = 1);
extern "C" int lsps_conv_wgrad(lsps_ctx* ctx, const lsps_conv_shape* s, const void* x, const void* dy, float* dw,
                               lsps_stream st_) {
  return run_wgrad(ctx, s, x, dy, dw, 0, static_cast<cudaStream_t>(st_));
}
extern "C" int lsps_conv_wgrad_split(lsps_ctx* ctx, const lsps_conv_shape* s, const void* x, const void* dy, float* dw,
                                     lsps_stream st_) {
  return run_wgrad(ctx, s, x, dy, dw, 1, static_cast<cudaStream_t>(st_));
}
extern "C" int lsps_conv_wgrad_grouped(lsps_ctx* ctx, const lsps_conv_shape* s, const void* x, const void* dy,
                                       float* dw, int groups, lsps_stream st_) {
  return run_wgrad(ctx, s, x, dy, dw, 0, static_cast<cudaStream_t>(st_), groups);
}
static int run_wgrad(lsps_ctx* ctx, const lsps_conv_shape* s, const void* x, const void* dy, float* dw, int split,
                     cudaStream_t st, int groups) {
  if (!ctx || !s || !x || !dy || !dw) return lsps_set_error(ctx, LSPS_E_ARG, "null argument");
  const int kind = s->kind, n = s->n, h = s->h, w = s->w, cin = s->cin, cout = s->cout;
  if (kind < 0 || kind > 4 || n <= 0 || !is_pow2(h) || !is_pow2(w) || cin % 64 || cout % 64 || cin <= 0 || cout <= 0)
    return lsps_set_error(ctx, LSPS_E_SHAPE, "wgrad shape kind %d n %d h %d w %d cin %d cout %d", kind, n, h, w, cin, cout);
  const bool k1 = kind == LSPS_CONV1X1;
  const int gw = groups > 1 ? cin / groups : 0;
  if (groups > 1 && (kind != LSPS_CONV_S1 || cin != cout || cin % groups || (gw != 64 && gw != 128) || cout % 128 || split))
    return lsps_set_error(ctx, LSPS_E_SHAPE, "grouped wgrad: 3x3 stride 1, cin == cout, group width 64 or 128 (got %d)", gw);
  const int ho = (kind == LSPS_CONV_S1 || k1) ? h : (kind == LSPS_CONV_S2 ? h / 2 : 2 * h);
  const int wo = (kind == LSPS_CONV_S1 || k1) ? w : (kind == LSPS_CONV_S2 ? w / 2 : 2 * w);
  // reduction grid: the coarser of the two spatial grids
  const int hg = kind == LSPS_CONV_S2 ? ho : h, wg = kind == LSPS_CONV_S2 ? wo : w;
  const int bn = gw ? 128 : (cin % 256 == 0 ? 256 : (cin % 128 == 0 ? 128 : 64));
  // CTA pairs: 256 output channels per pair, each CTA stages half of the x tile (needs >= 128 input channels)
  const int cg = (!gw && lsps_use_pairs() && cout % 256 == 0 && bn >= 128) ? 2 : 1;
  const Geo g = geo_for(hg, wg, (cg == 2 || bn <= 128) ? 128 : 64);
  WgradParams p{};
  p.tiles_x = wg / g.tw; p.tiles_y = hg / g.th; p.tiles_i = (n + g.nb - 1) / g.nb;
  p.twl = g.twl; p.thl = g.thl; p.nb = g.nb;
  const bool k4 = kind == LSPS_DECONV4_S2;
  const int ks = k4 ? 4 : (k1 ? 1 : 3);
  p.ntaps = ks * ks; p.cout = cout; p.cin = cin; p.dw = dw; p.dbg = lsps_dbg();
  p.nvar = split ? 3 : 1; p.m_lo = cout; p.n_lo = cin; p.gw = gw;
  const int cout_t = split ? 2 * cout : cout, cin_t = split ? 2 * cin : cin;   // channels of the dy / x TENSORS
  p.co_tiles = (cout + 128 * cg - 1) / (128 * cg);
  p.ci_tiles = gw ? 1 : cin / bn;
  for (int r = 0; r < ks; ++r)
    for (int c = 0; c < ks; ++c) {
      WTap& T = p.taps[r * ks + c];
      T = WTap{0, 0, 0, 0, 0, 0, 0, 0};
      if (k4) {  // dy pair view at parity ((r+1)&1, (c+1)&1); x shifted by +1 (tap 0), 0 (taps 1, 2), -1 (tap 3)
        T.mp = (r + 1) & 1; T.mc = ((c + 1) & 1) * cout_t;
        T.ny = r == 0 ? 1 : (r == 3 ? -1 : 0); T.nx = c == 0 ? 1 : (c == 3 ? -1 : 0);
      } else if (k1) {
      } else if (kind == LSPS_CONV_S1) { T.ny = r - 1; T.nx = c - 1; }
      else if (kind == LSPS_CONV_S2) {
        int dy_, py, dx_, px;
        s2_axis(r, &dy_, &py); s2_axis(c, &dx_, &px);
        T.ny = dy_; T.np = py; T.nx = dx_; T.nc = px * cin_t;
      } else {  // deconv: dy pair view at parity (a,b); x shifted by (di,dj)
        const int a = r == 1 ? 0 : 1, b = c == 1 ? 0 : 1;
        T.mp = a; T.mc = b * cout_t;
        T.ny = r == 0 ? 1 : 0; T.nx = c == 0 ? 1 : 0;
      }
    }
  const int ptiles = p.tiles_x * p.tiles_y * p.tiles_i;
  const int work = p.ntaps * p.co_tiles * p.ci_tiles * cg;
  int splits = ctx->num_sms / work;
  if (splits < 1) splits = 1;
  if (splits > ptiles) splits = ptiles;
  p.splits = splits;
  CUtensorMap tmM, tmN;
  // M side: dy (cout channels) ; N side: x (cin channels)
  int rc = act_tmap(ctx, dy, n, ho, wo, cout_t, kind == LSPS_DECONV_S2 || k4, g, &tmM);
  if (rc) return rc;
  rc = act_tmap(ctx, x, n, h, w, cin_t, kind == LSPS_CONV_S2, g, &tmN);
  if (rc) return rc;
  if (cg == 2) rc = bn == 256 ? launch_wgrad<256, 2>(ctx, tmM, tmN, p, st) : launch_wgrad<128, 2>(ctx, tmM, tmN, p, st);
  else rc = bn == 256 ? launch_wgrad<256, 1>(ctx, tmM, tmN, p, st)
                      : (bn == 128 ? launch_wgrad<128, 1>(ctx, tmM, tmN, p, st) : launch_wgrad<64, 1>(ctx, tmM, tmN, p, st));
  return rc;
}
