// Fused 1-D convolution kernels on CUDA cores (fp32 FMA) for the TemporalUNet.
//
// One kernel computes, for a tile of RB trajectory rows x NCB output channels and ALL output
// positions:   y = [Mish(GroupNorm8(conv(x) + b)) (+ time embedding)] (+ residual)
// i.e. a whole reference Conv1dBlock (blocks.py:13-34) plus the additions of
// ResidualConvolutionBlock.forward (blocks.py:154-166) in one launch.  The convolution is a direct
// conv with the (l_in, tap) -> l_out map resolved at compile time, so zero-padding taps cost
// nothing (only the 61.1 M non-padding MACs per row are executed).
//
// Layout: activations [rows][C][L] float32 (the reference's own [B, C, L]); weights repacked to
// [C_in][tap][C_out] so a K-chunk is contiguous over output channels.
#pragma once
#include "common.cuh"

namespace edmp {

enum ConvOp { OP_CONV5 = 0, OP_DOWN3 = 1, OP_UP4 = 2, OP_PW1 = 3 };

template <int OP> struct OpTraits;
template <> struct OpTraits<OP_CONV5> {   // nn.Conv1d(k=5, padding=2)            blocks.py:23
  static constexpr int NT = 5;
  __host__ __device__ static constexpr int lout(int lin, int tap) { return lin + 2 - tap; }
};
template <> struct OpTraits<OP_DOWN3> {   // nn.Conv1d(k=3, stride=2, padding=1)  blocks.py:211
  static constexpr int NT = 3;
  __host__ __device__ static constexpr int lout(int lin, int tap) {
    return ((lin + 1 - tap) >= 0 && ((lin + 1 - tap) % 2 == 0)) ? (lin + 1 - tap) / 2 : -1;
  }
};
template <> struct OpTraits<OP_UP4> {     // nn.ConvTranspose1d(k=4, stride=2, padding=1) blocks.py:249
  static constexpr int NT = 4;
  __host__ __device__ static constexpr int lout(int lin, int tap) { return 2 * lin - 1 + tap; }
};
template <> struct OpTraits<OP_PW1> {     // nn.Conv1d(k=1)  (residual_conv, blocks.py:151)
  static constexpr int NT = 1;
  __host__ __device__ static constexpr int lout(int lin, int) { return lin; }
};

struct ConvArgs {
  const float* xa; int ca;      // input, channels [0, ca)
  const float* xb; int cb;      // optional second input (skip concat, blocks.py:253): channels [ca, ca+cb)
  const float* w;               // [ca+cb][NT][cout]
  const float* bias;            // [cout]
  const float* gamma;           // GroupNorm affine (GN kernels only)
  const float* beta;
  const float* temb;            // [cout] time-MLP output added after Mish (block 0) or nullptr
  const float* ra; int rca;     // residual source (block 1): channels [0, rca)
  const float* rb; int rcb;     //   second half when the block input was a concat
  const float* wres;            // [rca+rcb][1][cout] residual 1x1 conv, nullptr = identity
  const float* bres;
  float* y;                     // [rows][cout][LOUT]
  int rows, cout;
};

constexpr int kKC = 16;  // input channels per shared-memory K chunk

template <int OP, int LIN, int LOUT, int TR, int TC, int RB, int NCB>
struct ConvTile {
  static constexpr int NT = OpTraits<OP>::NT;
  static constexpr int TX = NCB / TC, TY = RB / TR, THREADS = TX * TY;
  static constexpr int LMAX = LIN > LOUT ? LIN : LOUT;
  static constexpr int XS_STRIDE = kKC * LMAX + (((kKC * LMAX) % 2 == 0) ? 1 : 0);
  static constexpr int XS_FLOATS = RB * XS_STRIDE;
  static constexpr int WS_FLOATS = kKC * NT * NCB;
  static constexpr int OS_FLOATS = RB * NCB * LOUT;
  static constexpr int SMEM_FLOATS =
      (XS_FLOATS + WS_FLOATS) > OS_FLOATS ? (XS_FLOATS + WS_FLOATS) : OS_FLOATS;
  static_assert(THREADS % 32 == 0, "block must be whole warps");
  static_assert(NCB % TC == 0 && RB % TR == 0, "tile divisibility");
};

// acc[TR][TC][LOUTA] += sum over input channels / taps.  P = phase op (the main conv or the 1x1
// residual conv), LI/LO its input/output lengths.
template <int P, int LI, int LO, int TR, int TC, int RB, int NCB, int XS_STRIDE, int THREADS, int LOUTA>
__device__ __forceinline__ void conv_accumulate(float (&acc)[TR][TC][LOUTA], const float* __restrict__ xa,
                                                int ca, const float* __restrict__ xb, int cb,
                                                const float* __restrict__ w, int cout, int rows, int row0,
                                                int co0, int tx, int ty, float* Xs, float* Ws) {
  constexpr int NT = OpTraits<P>::NT;
  const int cin = ca + cb;
  for (int c0 = 0; c0 < cin; c0 += kKC) {
    const int kcn = min(kKC, cin - c0);
    // stage x[row0 .. row0+RB)[c0 .. c0+kcn)[0 .. LI): contiguous kcn*LI floats per row
    const float* src;
    int cs, cl;
    if (c0 < ca) { src = xa; cs = ca; cl = c0; } else { src = xb; cs = cb; cl = c0 - ca; }
    const int per_row = kcn * LI;
    for (int idx = threadIdx.x; idx < RB * per_row; idx += THREADS) {
      const int r = idx / per_row, e = idx - r * per_row;
      const int row = row0 + r;
      Xs[r * XS_STRIDE + e] = row < rows ? __ldg(src + ((size_t)row * cs + cl) * LI + e) : 0.0f;
    }
    for (int idx = threadIdx.x; idx < kcn * NT * NCB; idx += THREADS) {
      const int kt = idx / NCB, co = idx - kt * NCB;
      Ws[idx] = __ldg(w + ((size_t)c0 * NT + kt) * cout + co0 + co);
    }
    __syncthreads();
    for (int kc = 0; kc < kcn; ++kc) {
      float wv[NT][TC];
#pragma unroll
      for (int tap = 0; tap < NT; ++tap)
#pragma unroll
        for (int j = 0; j < TC; ++j) wv[tap][j] = Ws[(kc * NT + tap) * NCB + tx * TC + j];
#pragma unroll
      for (int i = 0; i < TR; ++i) {
        const float* xr = Xs + (ty * TR + i) * XS_STRIDE + kc * LI;
#pragma unroll
        for (int lin = 0; lin < LI; ++lin) {
          const float xv = xr[lin];
#pragma unroll
          for (int tap = 0; tap < NT; ++tap) {
            constexpr int dummy = 0;
            (void)dummy;
            const int lo = OpTraits<P>::lout(lin, tap);
            if (lo >= 0 && lo < LO) {
#pragma unroll
              for (int j = 0; j < TC; ++j) acc[i][j][lo] = fmaf(xv, wv[tap][j], acc[i][j][lo]);
            }
          }
        }
      }
    }
    __syncthreads();
  }
}

// GN_MISH: GroupNorm(8) + Mish (+temb) epilogue.  RES: 0 none, 1 identity residual, 2 1x1-conv residual.
template <int OP, int LIN, int LOUT, int TR, int TC, int RB, int NCB, bool GN_MISH, int RES>
__global__ void __launch_bounds__((NCB / TC) * (RB / TR))
conv_fused_kernel(ConvArgs a) {
  using Tile = ConvTile<OP, LIN, LOUT, TR, TC, RB, NCB>;
  constexpr int TX = Tile::TX, THREADS = Tile::THREADS;
  extern __shared__ float smem[];
  float* Xs = smem;
  float* Ws = smem + Tile::XS_FLOATS;
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int row0 = blockIdx.x * RB, co0 = blockIdx.y * NCB;
  pdl_launch_dependents();
  pdl_wait();

  float acc[TR][TC][LOUT];
#pragma unroll
  for (int i = 0; i < TR; ++i)
#pragma unroll
    for (int j = 0; j < TC; ++j)
#pragma unroll
      for (int l = 0; l < LOUT; ++l) acc[i][j][l] = 0.0f;

  conv_accumulate<OP, LIN, LOUT, TR, TC, RB, NCB, Tile::XS_STRIDE, THREADS, LOUT>(
      acc, a.xa, a.ca, a.xb, a.cb, a.w, a.cout, a.rows, row0, co0, tx, ty, Xs, Ws);

  float bj[TC];
#pragma unroll
  for (int j = 0; j < TC; ++j) bj[j] = __ldg(a.bias + co0 + tx * TC + j);
#pragma unroll
  for (int i = 0; i < TR; ++i)
#pragma unroll
    for (int j = 0; j < TC; ++j)
#pragma unroll
      for (int l = 0; l < LOUT; ++l) acc[i][j][l] += bj[j];

  if (GN_MISH) {
    // nn.GroupNorm(8, C) over (C/8 channels x LOUT) per row, eps 1e-5, biased variance
    // (blocks.py:24-26); the C/8 channels of a group sit in cg/TC adjacent lanes of one warp.
    const int cg = a.cout >> 3;
    const int lanes = cg / TC;  // power of two, <= 32
    const float inv_n = 1.0f / (float)(cg * LOUT);
    float gam[TC], bet[TC], te[TC];
#pragma unroll
    for (int j = 0; j < TC; ++j) {
      const int co = co0 + tx * TC + j;
      gam[j] = __ldg(a.gamma + co);
      bet[j] = __ldg(a.beta + co);
      te[j] = a.temb ? __ldg(a.temb + co) : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < TR; ++i) {
      float s = 0.0f;
#pragma unroll
      for (int j = 0; j < TC; ++j)
#pragma unroll
        for (int l = 0; l < LOUT; ++l) s += acc[i][j][l];
      for (int o = 1; o < lanes; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * inv_n;
      float ss = 0.0f;
#pragma unroll
      for (int j = 0; j < TC; ++j)
#pragma unroll
        for (int l = 0; l < LOUT; ++l) {
          const float d = acc[i][j][l] - mean;
          ss = fmaf(d, d, ss);
        }
      for (int o = 1; o < lanes; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rstd = 1.0f / sqrtf(ss * inv_n + 1e-5f);
#pragma unroll
      for (int j = 0; j < TC; ++j)
#pragma unroll
        for (int l = 0; l < LOUT; ++l) {
          const float v = (acc[i][j][l] - mean) * rstd * gam[j] + bet[j];
          acc[i][j][l] = mish_f(v) + te[j];
        }
    }
  }

  if (RES == 2) {
    // out + residual_conv(x): 1x1 conv of the block input accumulated on top (blocks.py:151,:164)
#pragma unroll
    for (int j = 0; j < TC; ++j) {
      const float br = __ldg(a.bres + co0 + tx * TC + j);
#pragma unroll
      for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int l = 0; l < LOUT; ++l) acc[i][j][l] += br;
    }
    conv_accumulate<OP_PW1, LOUT, LOUT, TR, TC, RB, NCB, Tile::XS_STRIDE, THREADS, LOUT>(
        acc, a.ra, a.rca, a.rb, a.rcb, a.wres, a.cout, a.rows, row0, co0, tx, ty, Xs, Ws);
  }

  // stage the tile in shared memory, then write rows out contiguously (coalesced)
  float* Os = smem;
#pragma unroll
  for (int i = 0; i < TR; ++i)
#pragma unroll
    for (int j = 0; j < TC; ++j)
#pragma unroll
      for (int l = 0; l < LOUT; ++l)
        Os[((ty * TR + i) * NCB + tx * TC + j) * LOUT + l] = acc[i][j][l];
  __syncthreads();
  constexpr int per_row = NCB * LOUT;
  for (int idx = threadIdx.x; idx < RB * per_row; idx += THREADS) {
    const int r = idx / per_row, e = idx - r * per_row;
    const int row = row0 + r;
    if (row < a.rows) {
      const size_t g = ((size_t)row * a.cout + co0) * LOUT + e;
      float v = Os[idx];
      if (RES == 1) v += __ldg(a.ra + g);  // identity residual: same [rows][cout][LOUT] layout
      a.y[g] = v;
    }
  }
}

}  // namespace edmp
