// TemporalUNet engine: weight repacking, time-embedding table, layer plan and launches.
//
// Replaces reference diffusion/models/temporalunet.py:11-76 and blocks.py (Conv1dBlock :13-34,
// SinusoidalPosEmb :38-54, TimeMLP :58-72, TimeEmbedding :76-92, ResidualConvolutionBlock
// :137-166, DownSampler :202-220, MiddleBlock :222-238, UpSampler :240-260).
//
// The time path depends only on the scalar t (the reference feeds a [1] tensor,
// diffusion.py:320), so all 24 TimeMLP outputs are tabulated once per weight load for
// t = 1..255 and the per-step work is a broadcast add inside the conv epilogue.
#include "unet.h"

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "conv_simt.cuh"
#include "conv_tc.cuh"
#include "conv_pm.cuh"
#include "conv_tc2.cuh"
#include "conv_pm2.cuh"
#include "edmp_b200.h"

namespace edmp {

// ---- tile configuration per (output length, output channels) ------------------------------------
template <int LOUT, int COUT> struct TileCfg;
#define EDMP_TILE(L, C, TR_, TC_, RB_, NCB_) \
  template <> struct TileCfg<L, C> { static constexpr int TR = TR_, TC = TC_, RB = RB_, NCB = NCB_; }
EDMP_TILE(50, 32, 1, 1, 4, 32);
EDMP_TILE(25, 32, 1, 1, 8, 32);
EDMP_TILE(25, 64, 1, 2, 4, 64);
EDMP_TILE(13, 64, 1, 2, 4, 64);
EDMP_TILE(13, 128, 1, 2, 4, 64);
EDMP_TILE(7, 128, 4, 4, 16, 64);
EDMP_TILE(7, 256, 4, 4, 16, 64);
EDMP_TILE(4, 256, 4, 4, 32, 64);
EDMP_TILE(4, 512, 4, 4, 32, 64);
EDMP_TILE(2, 512, 8, 4, 64, 64);
#undef EDMP_TILE

typedef int (*ConvLaunchFn)(const ConvArgs&, cudaStream_t);

template <int OP, int LIN, int LOUT, int COUT, bool GN, int RES>
static int launch_conv(const ConvArgs& a, cudaStream_t st) {
  using C = TileCfg<LOUT, COUT>;
  using Tile = ConvTile<OP, LIN, LOUT, C::TR, C::TC, C::RB, C::NCB>;
  dim3 grid((a.rows + C::RB - 1) / C::RB, COUT / C::NCB);
  launch_pdl(conv_fused_kernel<OP, LIN, LOUT, C::TR, C::TC, C::RB, C::NCB, GN, RES>, grid, dim3(Tile::THREADS),
             Tile::SMEM_FLOATS * sizeof(float), st, a);
  return 0;
}

static ConvLaunchFn pick_conv5(int L, int cout, int res) {
#define EDMP_CASE(L_, C_)                                                    \
  if (L == L_ && cout == C_) {                                               \
    if (res == 0) return launch_conv<OP_CONV5, L_, L_, C_, true, 0>;         \
    if (res == 1) return launch_conv<OP_CONV5, L_, L_, C_, true, 1>;         \
    return launch_conv<OP_CONV5, L_, L_, C_, true, 2>;                       \
  }
  EDMP_CASE(50, 32) EDMP_CASE(25, 32) EDMP_CASE(25, 64) EDMP_CASE(13, 64) EDMP_CASE(13, 128)
  EDMP_CASE(7, 128) EDMP_CASE(7, 256) EDMP_CASE(4, 256) EDMP_CASE(4, 512) EDMP_CASE(2, 512)
#undef EDMP_CASE
  return nullptr;
}

static ConvLaunchFn pick_down(int lin, int c) {
  if (lin == 50 && c == 32) return launch_conv<OP_DOWN3, 50, 25, 32, false, 0>;
  if (lin == 25 && c == 64) return launch_conv<OP_DOWN3, 25, 13, 64, false, 0>;
  if (lin == 13 && c == 128) return launch_conv<OP_DOWN3, 13, 7, 128, false, 0>;
  if (lin == 7 && c == 256) return launch_conv<OP_DOWN3, 7, 4, 256, false, 0>;
  if (lin == 4 && c == 512) return launch_conv<OP_DOWN3, 4, 2, 512, false, 0>;
  return nullptr;
}

static ConvLaunchFn pick_up(int lin, int c) {
  // ConvTranspose1d doubles the length; lengths 8/14/26 lose their last column
  // (temporalunet.py:70-71), which here simply is never computed.
  if (lin == 2 && c == 512) return launch_conv<OP_UP4, 2, 4, 512, false, 0>;
  if (lin == 4 && c == 256) return launch_conv<OP_UP4, 4, 7, 256, false, 0>;
  if (lin == 7 && c == 128) return launch_conv<OP_UP4, 7, 13, 128, false, 0>;
  if (lin == 13 && c == 64) return launch_conv<OP_UP4, 13, 25, 64, false, 0>;
  if (lin == 25 && c == 32) return launch_conv<OP_UP4, 25, 50, 32, false, 0>;
  return nullptr;
}

// final nn.Conv1d(32, 7, 1) (temporalunet.py:36): thread per (row, position)
__global__ void final_pw_kernel(const float* __restrict__ h, const float* __restrict__ w /*[7][C]*/,
                                const float* __restrict__ b, int C, int rows, float* __restrict__ eps) {
  extern __shared__ float sw[];
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < 7 * C + 7; i += blockDim.x) sw[i] = i < 7 * C ? w[i] : b[i - 7 * C];
  __syncthreads();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * kHorizon) return;
  const int row = (int)(i / kHorizon), l = (int)(i % kHorizon);
  float acc[7] = {0, 0, 0, 0, 0, 0, 0};
  const float* hr = h + (size_t)row * C * kHorizon + l;
  for (int c = 0; c < C; ++c) {
    const float v = hr[(size_t)c * kHorizon];
#pragma unroll
    for (int j = 0; j < 7; ++j) acc[j] = fmaf(sw[j * C + c], v, acc[j]);
  }
#pragma unroll
  for (int j = 0; j < 7; ++j) eps[((size_t)row * 7 + j) * kHorizon + l] = acc[j] + sw[7 * C + j];
}

// ---- parameter table ------------------------------------------------------------------------------
struct TensorRef {
  size_t off = 0;
  int d0 = 0, d1 = 1, d2 = 1;
  size_t numel() const { return (size_t)d0 * d1 * d2; }
};

struct ParamWalker {
  std::map<std::string, TensorRef> table;
  size_t cursor = 0;
  void add(const std::string& name, int d0, int d1 = 1, int d2 = 1) {
    TensorRef t;
    t.off = cursor; t.d0 = d0; t.d1 = d1; t.d2 = d2;
    table[name] = t;
    cursor += t.numel();
  }
  void res_block(const std::string& p, int cin, int cout) {
    add(p + ".blocks.0.block.0.weight", cout, cin, 5);
    add(p + ".blocks.0.block.0.bias", cout);
    add(p + ".blocks.0.block.2.weight", cout);
    add(p + ".blocks.0.block.2.bias", cout);
    add(p + ".blocks.1.block.0.weight", cout, cout, 5);
    add(p + ".blocks.1.block.0.bias", cout);
    add(p + ".blocks.1.block.2.weight", cout);
    add(p + ".blocks.1.block.2.bias", cout);
    add(p + ".time_mlp.time_mlp.1.weight", cout, 32);
    add(p + ".time_mlp.time_mlp.1.bias", cout);
    if (cin != cout) {
      add(p + ".residual_conv.weight", cout, cin, 1);
      add(p + ".residual_conv.bias", cout);
    }
  }
};

// state_dict order of TemporalUNet(input_dim=7, time_dim=32, dims) -- temporalunet.py:11-36
static void walk_params(const int* dims, int n_dims, ParamWalker& w) {
  std::vector<int> d(1, kDof);
  for (int i = 0; i < n_dims; ++i) d.push_back(dims[i]);
  w.add("time_embedding.time_mlp.1.weight", 128, 32);
  w.add("time_embedding.time_mlp.1.bias", 128);
  w.add("time_embedding.time_mlp.3.weight", 32, 128);
  w.add("time_embedding.time_mlp.3.bias", 32);
  const int n_down = (int)d.size() - 1;
  for (int i = 0; i < n_down; ++i) {
    const std::string p = "down_samplers." + std::to_string(i) + ".down.";
    w.res_block(p + "0", d[i], d[i + 1]);
    w.res_block(p + "1", d[i + 1], d[i + 1]);
    if (i != n_down - 1) {
      w.add(p + "3.weight", d[i + 1], d[i + 1], 3);
      w.add(p + "3.bias", d[i + 1]);
    }
  }
  w.res_block("middle_block.middle.0", d.back(), d.back());
  w.res_block("middle_block.middle.2", d.back(), d.back());
  int n = 0;
  for (int i = (int)d.size() - 1; i > 1; --i, ++n) {
    const std::string p = "up_samplers." + std::to_string(n) + ".up.";
    w.res_block(p + "0", 2 * d[i], d[i - 1]);
    w.res_block(p + "1", d[i - 1], d[i - 1]);
    w.add(p + "3.weight", d[i - 1], d[i - 1], 4);
    w.add(p + "3.bias", d[i - 1]);
  }
  w.add("final_conv.0.block.0.weight", d[1], d[1], 5);
  w.add("final_conv.0.block.0.bias", d[1]);
  w.add("final_conv.0.block.2.weight", d[1]);
  w.add("final_conv.0.block.2.bias", d[1]);
  w.add("final_conv.1.weight", kDof, d[1], 1);
  w.add("final_conv.1.bias", kDof);
}

size_t unet_param_count(const int* dims, int n_dims) {
  ParamWalker w;
  walk_params(dims, n_dims, w);
  return w.cursor;
}

// ---- engine -----------------------------------------------------------------------------------------
struct Act {
  float* p = nullptr;              // plain [rows][C][L] float32 (CUDA-core layers), may be null
  void *thi = nullptr, *tlo = nullptr;   // tiled hi / lo operand blocks (rows-as-M tensor-core layers)
  void *phi = nullptr, *plo = nullptr;   // position-major hi / lo images (conv_pm.cuh)
  int C = 0, L = 0;
  bool ok = true;
};

enum LayerKind { LAYER_SIMT = 0, LAYER_TC = 1, LAYER_PACK = 2, LAYER_PM = 3, LAYER_PM_PACK = 4, LAYER_TC2 = 5 };

struct Layer {
  int kind = LAYER_SIMT;
  ConvLaunchFn fn = nullptr;
  ConvArgs args;
  TcArgs targs;                 // LAYER_TC
  Tc2Args t2;                   // LAYER_TC2 (persistent kernel, conv_tc2.cuh)
  PmArgs pargs;                 // LAYER_PM
  bool pm_final = false;        // LAYER_PM: writes eps (the caller's buffer) through the fused final 1x1 conv
  int tc_tiles = 0;             // column tiles (grid.y) of a tensor-core layer
  int cta_group = 1;            // LAYER_TC2: 2 = CTA pairs (cta_group::2)
  int b_pad = 0;                // LAYER_TC2 pairs: weight stages with shared resident zero slots
  int nsplit = 1;               // LAYER_TC2: CTAs of a cluster that share one GroupNorm group (column split, small batches)
  int t2_lean = 0;              // LAYER_TC2: lean issue path (EDMP_MMA_LEAN=1 at creation; implies one issuing warp)
  int t2_max_slots = 0;         // LAYER_TC2: weight slots per stage, bytes per slot, bytes of the GroupNorm piece area
  size_t t2_slot_bytes = 0, t2_part_bytes = 0;
  size_t tc_smem = 0;
  Act pack_src, pack_dst;       // LAYER_PACK
  int temb_off = -1;  // offset into the per-t time-embedding row, -1 = none
  std::string name;   // reference module path of the op
  double macs_per_row = 0.0;  // non-padding multiply-accumulates per trajectory row
};

// non-padding MACs per row of one op: valid (l_in, tap) pairs x C_in x C_out
template <int OP> static double count_pairs(int lin, int lout) {
  int n = 0;
  for (int l = 0; l < lin; ++l)
    for (int t = 0; t < OpTraits<OP>::NT; ++t) {
      const int lo = OpTraits<OP>::lout(l, t);
      if (lo >= 0 && lo < lout) ++n;
    }
  return (double)n;
}

struct UNet {
  int precision = 0;
  bool tc = false;        // tensor-core path for the L <= 7 levels
  bool tc_split = false;  // three-MMA hi/lo operand split
  int tc_el = TC_EL_TF32; // operand element type of the tensor path (TF32 / BF16 / IEEE half)
  bool tc_16() const { return tc_el != TC_EL_TF32; }
  bool pm = false;        // position-major tensor-core kernels for the horizon 25 / 50 levels (16-bit elements)
  bool final_fused = false;   // final 1x1 conv fused into the last position-major layer
  bool tc2 = false;           // persistent second-generation kernel (conv_tc2.cuh) for the rows-as-M levels
  bool cg2 = false;           // CTA pairs (cta_group::2) for the horizon 2 / 4 levels (large batches)
  bool pm2 = false;           // persistent position-major kernel (conv_pm2.cuh): one CTA per SM, all output channels per CTA
  bool pm_pair = false;       // conv_pm2 with cta_group::2: CTA pairs walk double row blocks (half the MMAs per block)
  bool chain = false;         // row-tile chaining between consecutive conv_tc2 launches (conv_tc2.cuh)
  bool narrow = false;        // small batches: narrower column tiles / column-split GroupNorm groups (more tiles per layer)
  int* tile_done = nullptr;   // [layer][row tile] progress counters, zeroed at the start of every forward
  int tile_stride = 0;        // row tiles per layer in tile_done
  int sm_count = 148;
  int cpc() const { return tc_16() ? 64 : 32; }   // channels per 128-byte K chunk
  long long* dbg = nullptr;  // clock stamps of tensor-core CTAs (debug)
  unsigned* range_flag = nullptr;   // device flag: an activation left the IEEE-half operand range (f16x3 / f16 modes)
  // Packed-weight blob: every device image this engine uploads -- repacked weight tiles,
  // per-channel vectors, the time-embedding table -- in creation order, each item tagged with the tile geometry it was
  // packed for.  `rec`: being recorded while the engine is built from a state_dict; `play`: the engine is built FROM a
  // blob (no parameters, no repacking: items are uploaded as they are, sizes and tags must match the plan).
  std::vector<uint8_t>* rec = nullptr;
  const uint8_t* play = nullptr;
  size_t play_bytes = 0, play_pos = 0;
  bool play_bad = false;
  int max_rows = 0;
  int n_launches = 0;
  // launch units of one forward: a single layer, or a RUN of consecutive conv_tc2 layers in one persistent launch
  struct Launch {
    int first = 0, n = 1;         // layers [first, first + n)
    bool run = false;             // conv_tc2 run (Tc2Run geometry below)
    int a_stages = 0, b_stages = 0, b_stage_stride = 0, b_lo_off = 0, b_real_off = 0, b_total_bytes = 0, slot_bytes = 0;
    size_t smem = 0;
    std::string name;
    double macs_per_row = 0.0;
  };
  std::vector<Launch> launches;
  std::vector<void*> dev_allocs;
  std::vector<Layer> layers;
  std::map<std::string, Act> acts;
  float* temb = nullptr;  // [255][temb_width]
  int temb_width = 0;
  // final 1x1 conv
  float *final_w = nullptr, *final_b = nullptr;
  Act final_in;
  int final_c = 0;
};

constexpr uint32_t kBlobLayoutVersion = 3;        // bump whenever a packed layout, the item order or a tag changes
struct BlobHeader {
  char magic[8];                                  // "EDMPBLOB"
  uint32_t version, precision, n_dims, dims[8], max_rows;
  uint64_t n_params;
};
struct BlobItem { uint64_t bytes, tag; };

static void blob_append(std::vector<uint8_t>& v, const void* data, size_t bytes) {
  const uint8_t* b = static_cast<const uint8_t*>(data);
  v.insert(v.end(), b, b + bytes);
  v.resize((v.size() + 15) & ~(size_t)15, 0);
}
// next item of the blob being played: its payload, or null (and play_bad) when it is not what the plan asks for
static const uint8_t* blob_next(UNet* u, size_t bytes, uint64_t tag) {
  BlobItem it;
  if (u->play_bad || u->play_pos + sizeof(it) > u->play_bytes) { u->play_bad = true; return nullptr; }
  std::memcpy(&it, u->play + u->play_pos, sizeof(it));
  const size_t payload = u->play_pos + sizeof(it);
  if (it.bytes != bytes || it.tag != tag || payload + bytes > u->play_bytes) { u->play_bad = true; return nullptr; }
  u->play_pos = (payload + bytes + 15) & ~(size_t)15;
  return u->play + payload;
}

// device copy of a packed image; recorded into / served from the blob (`tag`: geometry the image was packed for)
static void* upload_bytes(UNet* u, const void* data, size_t bytes, uint64_t tag = 0) {
  if (u->play) {
    data = blob_next(u, bytes, tag);
    if (!data) return nullptr;
  } else if (u->rec) {
    BlobItem it{bytes, tag};
    blob_append(*u->rec, &it, sizeof(it));
    blob_append(*u->rec, data, bytes);
  }
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
  if (cudaMemcpy(p, data, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(p);
    return nullptr;
  }
  u->dev_allocs.push_back(p);
  return p;
}
// a host-side scalar of the plan that depends on the weights (the power-of-two scale of a packed tile)
static void blob_scalar(UNet* u, float* v) {
  if (u->play) {
    const uint8_t* src = blob_next(u, sizeof(float), 0x5CA1E);
    if (src) std::memcpy(v, src, sizeof(float));
  } else if (u->rec) {
    BlobItem it{sizeof(float), 0x5CA1E};
    blob_append(*u->rec, &it, sizeof(it));
    blob_append(*u->rec, v, sizeof(float));
  }
}
static uint64_t geom_tag(uint64_t a, uint64_t b, uint64_t c, uint64_t d, uint64_t e) {
  return ((((a * 1315423911ull + b) * 1315423911ull + c) * 1315423911ull + d) * 1315423911ull + e) | (1ull << 63);
}
static float* upload(UNet* u, const std::vector<float>& v) {
  return static_cast<float*>(upload_bytes(u, v.data(), v.size() * sizeof(float)));
}

static Act new_act(UNet* u, const std::string& name, int C, int L, bool plain = true, bool tiled = false,
                   bool pm = false) {
  Act a;
  a.C = C; a.L = L;
  bool ok = true;
  if (pm) {
    // [row block of 8][L + 4 positions][8 rows][C halves]; zero-filled once: the halo positions and the
    // padding rows of the last block are never written and must stay zero / finite
    const size_t n = (size_t)((u->max_rows + kPmRows - 1) / kPmRows) * (L + 4) * 16 * C;
    ok = cudaMalloc(&a.phi, n) == cudaSuccess;
    if (ok) { u->dev_allocs.push_back(a.phi); cudaMemset(a.phi, 0, n); }
    if (ok && u->tc_split) {
      ok = cudaMalloc(&a.plo, n) == cudaSuccess;
      if (ok) { u->dev_allocs.push_back(a.plo); cudaMemset(a.plo, 0, n); }
    }
  }
  if (plain) {
    ok = cudaMalloc(&a.p, (size_t)u->max_rows * C * L * sizeof(float)) == cudaSuccess;
    if (ok) u->dev_allocs.push_back(a.p);
  }
  if (ok && tiled) {
    // row tiles of 128; zero-filled once so the padding rows of the last tile stay finite
    // (an even number of row tiles: CTA pairs of the cta_group::2 layers own two row tiles each)
    const size_t n = (size_t)(((u->max_rows + kTcRows - 1) / kTcRows + 1) & ~1) * L * (C / u->cpc()) * kTcBlockBytes;
    ok = cudaMalloc(&a.thi, n) == cudaSuccess;
    if (ok) { u->dev_allocs.push_back(a.thi); cudaMemset(a.thi, 0, n); }
    if (ok && u->tc_split) {
      ok = cudaMalloc(&a.tlo, n) == cudaSuccess;
      if (ok) { u->dev_allocs.push_back(a.tlo); cudaMemset(a.tlo, 0, n); }
    }
  }
  if (!ok) { a.p = nullptr; a.thi = nullptr; a.phi = nullptr; a.ok = false; return a; }
  if (!name.empty()) u->acts[name] = a;
  return a;
}

static float host_mish(float x) {
  // nn.Mish: x * tanh(softplus(x))
  double sp = x > 20.0f ? (double)x : std::log1p(std::exp((double)x));
  return (float)((double)x * std::tanh(sp));
}

// y[n] = W[n][k] x[k] + b[n]  (nn.Linear), float32 result
static void host_linear(const float* W, const float* b, const float* x, int n, int k, float* y) {
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int j = 0; j < k; ++j) s += (double)W[(size_t)i * k + j] * (double)x[j];
    y[i] = (float)(s + (double)b[i]);
  }
}

struct Builder {
  UNet* u;
  const float* params;
  ParamWalker* pw;
  std::vector<std::pair<std::string, int>> temb_blocks;  // (res-block prefix, offset)
  int temb_width = 0;
  bool ok = true;
  std::string why;            // which plan check failed (for the error message)

  // (null while the engine is built from a blob: nothing below may read parameters then)
  const float* P(const std::string& name) const { return params ? params + pw->table.at(name).off : nullptr; }
  bool playing() const { return u->play != nullptr; }

  // conv weight [cout][cin][k] -> [cin][k][cout]
  float* pack_conv(const std::string& name, int cout, int cin, int k) {
    const float* w = P(name);
    std::vector<float> v((size_t)cin * k * cout);
    if (!playing())
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci)
        for (int t = 0; t < k; ++t) v[((size_t)ci * k + t) * cout + co] = w[((size_t)co * cin + ci) * k + t];
    float* p = upload(u, v);
    ok = ok && p;
    return p;
  }
  // transposed conv weight [cin][cout][k] -> [cin][k][cout]
  float* pack_convT(const std::string& name, int cin, int cout, int k) {
    const float* w = P(name);
    std::vector<float> v((size_t)cin * k * cout);
    if (!playing())
    for (int ci = 0; ci < cin; ++ci)
      for (int co = 0; co < cout; ++co)
        for (int t = 0; t < k; ++t) v[((size_t)ci * k + t) * cout + co] = w[((size_t)ci * cout + co) * k + t];
    float* p = upload(u, v);
    ok = ok && p;
    return p;
  }
  float* vec(const std::string& name, int n) {
    std::vector<float> v((size_t)n);
    if (!playing()) std::copy(P(name), P(name) + n, v.begin());
    float* p = upload(u, v);
    ok = ok && p;
    return p;
  }

  // Conv1dBlock: conv5 + GN + Mish, optional temb add / residual
  Act conv_block(const std::string& p, const Act& xa, const Act* xb, int cout, const std::string& out_name,
                 int temb_off, int res, const Act* ra, const Act* rb, const std::string& res_prefix) {
    const int cin = xa.C + (xb ? xb->C : 0);
    const int L = xa.L;
    Act y = new_act(u, out_name, cout, L);
    ok = ok && y.p;
    Layer ly;
    std::memset(&ly.args, 0, sizeof(ConvArgs));
    ly.fn = pick_conv5(L, cout, res);
    ok = ok && ly.fn;
    ly.args.xa = xa.p; ly.args.ca = xa.C;
    ly.args.xb = xb ? xb->p : nullptr; ly.args.cb = xb ? xb->C : 0;
    ly.args.w = pack_conv(p + ".block.0.weight", cout, cin, 5);
    ly.args.bias = vec(p + ".block.0.bias", cout);
    ly.args.gamma = vec(p + ".block.2.weight", cout);
    ly.args.beta = vec(p + ".block.2.bias", cout);
    ly.temb_off = temb_off;
    if (res) {
      ly.args.ra = ra->p; ly.args.rca = ra->C;
      ly.args.rb = rb ? rb->p : nullptr; ly.args.rcb = rb ? rb->C : 0;
      if (res == 2) {
        ly.args.wres = pack_conv(res_prefix + ".residual_conv.weight", cout, ra->C + (rb ? rb->C : 0), 1);
        ly.args.bres = vec(res_prefix + ".residual_conv.bias", cout);
      }
    }
    ly.args.y = y.p;
    ly.args.cout = cout;
    ly.name = p;
    ly.macs_per_row = count_pairs<OP_CONV5>(L, L) * cin * cout +
                      (res == 2 ? (double)L * (ra->C + (rb ? rb->C : 0)) * cout : 0.0);
    u->layers.push_back(ly);
    return y;
  }

  // ResidualConvolutionBlock (blocks.py:154-166)
  Act res_block(const std::string& p, const Act& xa, const Act* xb, int cout) {
    const int cin = xa.C + (xb ? xb->C : 0);
    const int off = temb_width;
    temb_blocks.push_back({p, off});
    temb_width += cout;
    Act h = conv_block(p + ".blocks.0", xa, xb, cout, p + ".blocks.0", off, 0, nullptr, nullptr, "");
    return conv_block(p + ".blocks.1", h, nullptr, cout, p, -1, cin != cout ? 2 : 1, &xa, xb, p);
  }


  // ---------------- tensor-core (tcgen05) layers ------------------------------------------------
  static float tf32_round(float x) {   // cvt.rna.tf32.f32 on the host
    uint32_t b;
    std::memcpy(&b, &x, 4);
    b = (b + 0x1000u) & 0xFFFFE000u;
    float r;
    std::memcpy(&r, &b, 4);
    return r;
  }

  static uint16_t bf16_round(float x) {  // round to nearest even
    uint32_t b;
    std::memcpy(&b, &x, 4);
    b += 0x7FFFu + ((b >> 16) & 1u);
    return (uint16_t)(b >> 16);
  }
  static float bf16_to_float(uint16_t h) {
    uint32_t b = (uint32_t)h << 16;
    float r;
    std::memcpy(&r, &b, 4);
    return r;
  }
  static uint16_t f16_round(float x) { return __half_as_ushort(__float2half_rn(x)); }   // RNE incl. subnormals
  static float f16_to_float(uint16_t h) { return __half2float(__ushort_as_half(h)); }

  // Packs weights into UMMA-ready tiles [n_tile][c_chunk][slots*ct rows][128 B], SWIZZLE_128B,
  // K-major; wfn(co, ci, slot) supplies the value.  Produces the hi part and (split modes) the lo
  // remainder in the operand element type (TF32-rounded fp32 or BF16).
  // IEEE-half operands: the weights are multiplied by a power of two `scale` (returned through
  // acc_scale as its exact inverse) so that their lo parts stay in half's normal range; TF32 / BF16
  // have fp32's exponent range and use scale 1.
  template <class F>
  void pack_tc(int cout, int cin, int ct, int slots, F wfn, const void** hi_out, const void** lo_out,
               float* acc_scale, int halves = 1) {
    const int cpc = u->cpc(), ebytes = u->tc_16() ? 2 : 4, epc = 16 / ebytes;
    const int n_tiles = cout / ct, kch = cin / cpc;
    const uint64_t tag = geom_tag(cout, cin, ct, slots, halves);
    if (playing()) {
      const size_t bytes = (size_t)slots * ct * 128 * n_tiles * kch;
      blob_scalar(u, acc_scale);
      *hi_out = upload_bytes(u, nullptr, bytes, tag);
      *lo_out = u->tc_split ? upload_bytes(u, nullptr, bytes, tag) : nullptr;
      ok = ok && *hi_out && (!u->tc_split || *lo_out);
      return;
    }
    float scale = 1.0f;
    if (u->tc_el == TC_EL_F16) {
      float wmax = 0.0f;
      for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci)
          for (int sl = 0; sl < slots; ++sl) wmax = std::max(wmax, std::fabs(wfn(co, ci, sl)));
      int e = 0;
      if (wmax > 0.0f) e = (int)std::floor(std::log2(16384.0f / wmax));   // |scale * w| <= 16384 < 65504
      e = std::max(-8, std::min(24, e));
      scale = std::ldexp(1.0f, e);
    }
    *acc_scale = 1.0f / scale;
    blob_scalar(u, acc_scale);
    const size_t tile = (size_t)slots * ct * 128;   // bytes
    std::vector<uint8_t> hi(tile * n_tiles * kch), lo(u->tc_split ? hi.size() : 0);
    for (int nt = 0; nt < n_tiles; ++nt)
      for (int cc = 0; cc < kch; ++cc) {
        const size_t base = ((size_t)nt * kch + cc) * tile;
        const int cth = ct / halves;   // halves == 2: [half][slot][ct/2 rows], each half is one CTA's share of the tile
        for (int sl = 0; sl < slots; ++sl)
          for (int c = 0; c < ct; ++c) {
            const int hf = c / cth;
            const int r = sl * cth + (c - hf * cth);   // row inside the half
            for (int e = 0; e < cpc; ++e) {
              const float w = scale * wfn(nt * ct + c, cc * cpc + e, sl);
              const size_t off = base + (size_t)hf * (tile / halves) + (size_t)r * 128 +
                                 ((((e / epc) ^ (r & 7)) << 4) | ((e % epc) * ebytes));
              if (u->tc_el == TC_EL_F16) {
                const uint16_t h = f16_round(w);
                std::memcpy(&hi[off], &h, 2);
                if (u->tc_split) {
                  const uint16_t l = f16_round(w - f16_to_float(h));
                  std::memcpy(&lo[off], &l, 2);
                }
              } else if (u->tc_el == TC_EL_BF16) {
                const uint16_t h = bf16_round(w);
                std::memcpy(&hi[off], &h, 2);
                if (u->tc_split) {
                  const uint16_t l = bf16_round(w - bf16_to_float(h));
                  std::memcpy(&lo[off], &l, 2);
                }
              } else {
                const float h = tf32_round(w);
                std::memcpy(&hi[off], &h, 4);
                if (u->tc_split) {
                  const float l = tf32_round(w - h);
                  std::memcpy(&lo[off], &l, 4);
                }
              }
            }
          }
      }
    *hi_out = upload_bytes(u, hi.data(), hi.size(), tag);
    ok = ok && *hi_out;
    *lo_out = nullptr;
    if (u->tc_split) {
      *lo_out = upload_bytes(u, lo.data(), lo.size(), tag);
      ok = ok && *lo_out;
    }
  }

  void finish_tc_layer(Layer& ly, int n_tiles) {
    TcArgs& t = ly.targs;
    t.split = u->tc_split ? 1 : 0;
    const int nparts = t.split ? 2 : 1;
    int max_slots = t.ph[0].slots;
    if (t.n_phases > 1 && t.ph[1].slots > max_slots) max_slots = t.ph[1].slots;
    const size_t a_stage = (size_t)kTcBlockBytes * nparts;
    const size_t b_stage = (size_t)max_slots * t.ct * 128 * nparts;
    const size_t budget = 232448 - 1024 - 1024 - 10240;   // dynamic limit - alignment slack - barriers - static smem
    t.b_stages = (2 * b_stage + 2 * a_stage <= budget) ? 2 : 1;
    size_t rest = budget - t.b_stages * b_stage;
    t.a_stages = (int)std::min<size_t>(4, rest / a_stage);
    ok = ok && t.a_stages >= 2;
    ly.kind = LAYER_TC;
    ly.tc_tiles = n_tiles;
    // epilogue staging reuses the operand stages; plain-output layers need the whole tile at once
    size_t stage_total = t.a_stages * a_stage + t.b_stages * b_stage;
    const int n_units = t.lout * t.ct / 16;
    if (t.out_plain) {
      const size_t need = (size_t)kTcRows * (t.lout * t.ct + 1) * sizeof(float);
      stage_total = std::max(stage_total, need);
      ok = ok && need <= budget;
      t.epi_units = (n_units + 3) & ~3;
    } else {
      // per 16-column unit: hi + lo output staging (128 rows x 64 B tf32 / 32 B bf16 each) + 8 KB residual
      const size_t out_unit = (size_t)128 * (u->tc_16() ? 32 : 64) * 2;
      const size_t per_unit = out_unit + (t.mode == TC_GN_RES_ID ? 8192 : 0);
      stage_total = std::max(stage_total, 4 * per_unit);
      const int cap = (int)(stage_total / per_unit);
      t.epi_units = std::max(4, std::min((n_units + 3) & ~3, cap & ~3));
    }
    t.stage_bytes = (int)((stage_total + 1023) & ~(size_t)1023);
    ly.tc_smem = 1024 + t.stage_bytes + 1024;
    // (bias-only layers may take all 512 accumulator columns, single-buffered: the 13 -> 25 up-sampling has 25 x 16 = 400)
    const int max_cols = (t.mode == TC_BIAS && getenv("EDMP_UP3_V1") == nullptr) ? 512 : 16 * kT2MaxUnits;
    if (u->tc2 && !t.out_plain && t.lout * t.ct <= max_cols) finish_tc2_layer(ly, n_tiles);
  }

  // persistent kernel (conv_tc2.cuh): same operand layouts and weight tiles as conv_tc.cuh
  void finish_tc2_layer(Layer& ly, int n_tiles, int cgrp = 1) {
    const TcArgs& t = ly.targs;
    ly.cta_group = cgrp;
    Tc2Args& v = ly.t2;
    std::memset(&v, 0, sizeof(Tc2Args));
    v.n_phases = t.n_phases;
    for (int p = 0; p < t.n_phases; ++p) {
      const TcPhase& s = t.ph[p];
      Tc2Phase& d = v.ph[p];
      d.a = s.a; d.b = s.b; d.w_hi = s.w_hi; d.w_lo = s.w_lo; d.acc_scale = s.acc_scale;
      d.lin = s.lin; d.slots = s.slots; d.col_step = (cgrp == 2 && p == 0) ? t.ct / 2 : t.ct;
      d.d_col = cgrp == 2 ? p * 256 : s.d_col;
      uint32_t touched = 0;
      for (int li = 0; li < s.lin; ++li) {
        const TcSched& q = s.sched[li];
        d.sched[li].slot_begin = q.slot_begin; d.sched[li].n_slots = q.n_slots; d.sched[li].lo_begin = q.lo_begin;
        int n_acc = 0;
        while (n_acc < q.n_slots && ((touched >> (q.lo_begin + n_acc)) & 1u)) ++n_acc;
        for (int j = n_acc; j < q.n_slots; ++j) ok = ok && !((touched >> (q.lo_begin + j)) & 1u);   // touched positions form a prefix
        d.sched[li].n_acc = (int8_t)n_acc;
        for (int j = 0; j < q.n_slots; ++j) touched |= 1u << (q.lo_begin + j);
      }
    }
    v.lout = t.lout; v.ct = t.ct; v.cout = t.cout; v.cg = t.cg; v.mode = t.mode; v.split = u->tc_split ? 1 : 0;
    const int cols = t.lout * t.ct;
    ok = ok && cols <= (t.mode == TC_BIAS ? 512 : 16 * kT2MaxUnits) && (cols % 16) == 0 && t.ct <= 128;
    v.acc_bufs = ((t.n_phases == 1 && cols <= 256) || (t.n_phases == 2 && cgrp == 1 && t.ph[1].d_col + cols <= 256)) ? 2 : 1;
    v.acc_stride = 256;
    v.half_layout = cgrp == 2 ? 1 : 0;
    v.n_col_tiles = n_tiles;
    auto ilog2 = [](int x) { int l = 0; while ((1 << l) < x) ++l; return l; };
    v.ct_log2 = ilog2(t.ct); v.cg_log2 = ilog2(t.cg); v.nct_log2 = ilog2(n_tiles);
    ok = ok && (1 << v.ct_log2) == t.ct && (1 << v.cg_log2) == t.cg && (1 << v.nct_log2) == n_tiles;
    v.bias = t.bias; v.gamma = t.gamma; v.beta = t.beta; v.bres = t.bres;
    v.res = t.res; v.out_hi = t.out_hi; v.out_lo = t.out_lo; v.out_pm = t.out_pm;
    const int nparts = u->tc_split ? 2 : 1;
    int max_slots = t.ph[0].slots;
    if (t.n_phases > 1 && t.ph[1].slots > max_slots) max_slots = t.ph[1].slots;
    const size_t a_stage = (size_t)kTcBlockBytes * nparts;
    const size_t slot_bytes = (size_t)(t.ct / cgrp) * 128;
    v.b_pad = ly.b_pad;
    {
      // two MMA-issuing warps need every input position used and at least two steps per chunk (EDMP_MMA_WARPS=1: one)
      bool all = true;
      for (int p = 0; p < t.n_phases; ++p) {
        all = all && t.ph[p].lin >= 2;
        for (int li = 0; li < t.ph[p].lin; ++li) all = all && t.ph[p].sched[li].n_slots > 0;
      }
      // (default: the lean issue path -- one thread, tabulated schedule, conv_tc2.cuh; EDMP_MMA_LEAN=0: two alternating issuing
      // warps.  Alone the two forms are within 1-2 % of each other; with the weight stream on its own producer thread
      // (EDMP_PRODUCERS, run_tc2) the lean form is 1 % ahead at 8190 rows and 3 % at 1020, profiles/r2_s2_experiments.md)
      const bool lean = getenv("EDMP_MMA_LEAN") == nullptr || atoi(getenv("EDMP_MMA_LEAN")) != 0;
      static const int want = getenv("EDMP_MMA_WARPS") ? atoi(getenv("EDMP_MMA_WARPS")) : 2;
      v.mma_warps = (!lean && all && want == 2) ? 2 : 1;
      ly.t2_lean = lean ? 1 : 0;
    }
    // bytes of n weight stages (padded layout: per part n * (max_slots + 1) + 1 slots, see conv_tc2.cuh)
    auto b_bytes = [&](int n) {
      return ly.b_pad ? (size_t)nparts * ((size_t)n * (max_slots + 1) + 1) * slot_bytes : (size_t)n * max_slots * slot_bytes * nparts;
    };
    const size_t n_groups = t.cg == 8 ? t.ct / 8 : std::max(1, t.ct / t.cg);
    v.nsplit = ly.nsplit;
    ok = ok && (ly.nsplit == 1 || (cgrp == 1 && ly.nsplit * t.ct == t.cg && ly.nsplit <= 4 && t.mode != TC_BIAS));
    // GroupNorm pieces (mean, M2) per unit and row + group stats (+ the statistics exchange buffer of a column split)
    const size_t part_bytes = ((size_t)(cols / 16) * (t.cg == 8 ? 2 : 1) + n_groups) * 1024 + (ly.nsplit > 1 ? 8192 : 0);
    const size_t budget = 232448 - 1024 - 7168 - part_bytes;   // dynamic limit - alignment slack - static shared memory - pieces
    v.b_stages = (b_bytes(3) + 3 * a_stage <= budget) ? 3 : ((b_bytes(2) + 2 * a_stage <= budget) ? 2 : 1);
    if (getenv("EDMP_MAX_B_STAGES")) v.b_stages = std::min(v.b_stages, std::max(1, atoi(getenv("EDMP_MAX_B_STAGES"))));
    v.a_stages = (int)std::min<size_t>(kT2MaxAStages, (budget - b_bytes(v.b_stages)) / a_stage);
    if (getenv("EDMP_MAX_A_STAGES")) v.a_stages = std::min(v.a_stages, std::max(2, atoi(getenv("EDMP_MAX_A_STAGES"))));
    ok = ok && v.a_stages >= 2;
    ly.kind = LAYER_TC2;
    ly.tc_smem = 1024 + v.a_stages * a_stage + b_bytes(v.b_stages) + part_bytes;
    ly.t2_max_slots = max_slots; ly.t2_slot_bytes = slot_bytes; ly.t2_part_bytes = part_bytes;
  }

  static TcOperand operand(const Act* a) {
    TcOperand o;
    o.hi = a ? a->thi : nullptr;
    o.lo = a ? a->tlo : nullptr;
    o.C = a ? a->C : 0;
    return o;
  }

  // Conv1dBlock on tensor cores; inputs and output are tiled operands.
  Act tc_conv_block(const std::string& p, const Act& xa, const Act* xb, int cout, const std::string& out_name,
                    int temb_off, int res, const Act* ra, const Act* rb, const std::string& res_prefix) {
    const int cin = xa.C + (xb ? xb->C : 0);
    const int L = xa.L;
    const int cg = cout / 8;
    // column tile: one GroupNorm group, or two when groups have 8 channels; the persistent kernel takes 32-channel
    // tiles where 16 would leave the MMAs issue-bound (N = window x ct) and the accumulator still fits 256 columns
    int ct = std::max(16, cg);
    if (u->tc2 && ct < 32 && L * 32 <= 16 * kT2MaxUnits) ct = 32;
    // cta_group::2 (CTA pairs sharing each weight tile) where every window can cover all L positions: L = 2, and
    // L = 4 with one zero tap slot at either end; the accumulator is 256 columns per CTA
    const bool pair = u->cg2 && (L == 2 || L == 4) && cout >= 256;
    if (pair) ct = 256 / L;
    // small batches: halve the column tiles while one wave of tiles still fits the machine (the K loop of a tile is
    // the latency of the layer); below one GroupNorm group the CTAs of a cluster share the group (Tc2Args::nsplit)
    int nsplit = 1;
    if (u->tc2 && !pair && u->narrow) {
      const int rts = (u->max_rows + kTcRows - 1) / kTcRows;
      while (ct > 16 && (ct / 2) * 4 >= cg && rts * (cout / ct) * 2 <= u->sm_count) ct /= 2;
      if (ct < cg) nsplit = cg / ct;
    }
    Act y = new_act(u, out_name, cout, L, /*plain=*/false, /*tiled=*/true);
    ok = ok && y.ok;
    Layer ly;
    std::memset(&ly.args, 0, sizeof(ConvArgs));
    std::memset(&ly.targs, 0, sizeof(TcArgs));
    TcArgs& t = ly.targs;
    t.n_phases = 1;
    t.lout = L; t.ct = ct; t.cout = cout; t.cg = cg;
    t.mode = res == 0 ? TC_GN : (res == 1 ? TC_GN_RES_ID : TC_GN_RES_PW);
    TcPhase& ph = t.ph[0];
    ph.a = operand(&xa);
    ph.b = operand(xb);
    ph.lin = L;
    // slot s holds filter tap 4 - (j_begin + s); pairs pad the tap range so that every window spans all L positions
    // (the two all-zero end slots of a padded pair layer are not stored in the weight tiles: the kernel keeps them
    // resident in shared memory, Tc2Args::b_pad, and slot indices count from the leading zero slot)
    const bool padded = pair && 2 + (L - 1) > 4;
    const int j_begin = std::max(0, 2 - (L - 1)), j_end = std::min(4, 2 + (L - 1));
    const int j_view = padded ? -1 : j_begin;   // tap of view slot 0
    ph.slots = j_end - j_begin + 1;
    ph.d_col = 0;
    ly.b_pad = padded ? 1 : 0;
    ly.nsplit = nsplit;
    for (int li = 0; li < L; ++li) {
      const int lo_min = pair ? 0 : std::max(0, li - 2), lo_max = pair ? L - 1 : std::min(L - 1, li + 2);
      ph.sched[li].slot_begin = (int8_t)((lo_min - li + 2) - j_view);
      ph.sched[li].n_slots = (int8_t)(lo_max - lo_min + 1);
      ph.sched[li].lo_begin = (int8_t)lo_min;
    }
    const float* w = P(p + ".block.0.weight");   // [cout][cin][5]
    pack_tc(cout, cin, ct, ph.slots,
            [&](int co, int ci, int sl) {
              const int tap = 4 - (j_begin + sl);
              return (tap >= 0 && tap <= 4) ? w[((size_t)co * cin + ci) * 5 + tap] : 0.0f;
            },
            &ph.w_hi, &ph.w_lo, &ph.acc_scale, pair ? 2 : 1);
    t.bias = vec(p + ".block.0.bias", cout);
    t.gamma = vec(p + ".block.2.weight", cout);
    t.beta = vec(p + ".block.2.bias", cout);
    ly.temb_off = temb_off;
    double macs = count_pairs<OP_CONV5>(L, L) * cin * cout;
    if (res == 1) t.res = operand(ra);
    if (res == 2) {
      const int rcin = ra->C + (rb ? rb->C : 0);
      t.n_phases = 2;
      TcPhase& pr = t.ph[1];
      pr.a = operand(ra);
      pr.b = operand(rb);
      pr.lin = L;
      pr.slots = 1;
      pr.d_col = (L * ct <= 128) ? 128 : 256;
      for (int li = 0; li < L; ++li) { pr.sched[li].slot_begin = (int8_t)(padded ? 1 : 0); pr.sched[li].n_slots = 1; pr.sched[li].lo_begin = (int8_t)li; }
      const float* wr = P(res_prefix + ".residual_conv.weight");   // [cout][rcin][1]
      pack_tc(cout, rcin, ct, 1, [&](int co, int ci, int) { return wr[(size_t)co * rcin + ci]; }, &pr.w_hi, &pr.w_lo,
              &pr.acc_scale, pair ? 2 : 1);
      t.bres = vec(res_prefix + ".residual_conv.bias", cout);
      macs += (double)L * rcin * cout;
    }
    t.out_hi = y.thi;
    t.out_lo = y.tlo;
    ly.name = p;
    ly.macs_per_row = macs;
    if (pair) finish_tc2_layer(ly, cout / ct, 2);
    else finish_tc_layer(ly, cout / ct);
    u->layers.push_back(ly);
    return y;
  }

  Act tc_res_block(const std::string& p, const Act& xa, const Act* xb, int cout) {
    const int cin = xa.C + (xb ? xb->C : 0);
    const int off = temb_width;
    temb_blocks.push_back({p, off});
    temb_width += cout;
    Act h = tc_conv_block(p + ".blocks.0", xa, xb, cout, p + ".blocks.0", off, 0, nullptr, nullptr, "");
    return tc_conv_block(p + ".blocks.1", h, nullptr, cout, p, -1, cin != cout ? 2 : 1, &xa, xb, p);
  }

  // stride-2 Conv1d / ConvTranspose1d on tensor cores (tiled in; tiled and/or plain out)
  Act tc_resample(const std::string& name, const Act& x, bool up, bool plain_out, bool pm_out = false) {
    const int C = x.C, L = x.L;
    const int lout = up ? ((2 * L == 8 || 2 * L == 14 || 2 * L == 26) ? 2 * L - 1 : 2 * L) : (L + 1) / 2;
    int ct = std::max(16, C / 8);
    if (u->tc2 && ct < 32 && lout * 32 <= 16 * kT2MaxUnits && !plain_out) ct = 32;
    if (u->tc2 && u->narrow && !plain_out) {
      const int rts = (u->max_rows + kTcRows - 1) / kTcRows;
      while (ct > 16 && rts * (C / ct) * 2 <= u->sm_count) ct /= 2;
    }
    Act y = new_act(u, name, C, lout, plain_out, !plain_out && !pm_out, pm_out);
    ok = ok && y.ok;
    Layer ly;
    std::memset(&ly.args, 0, sizeof(ConvArgs));
    std::memset(&ly.targs, 0, sizeof(TcArgs));
    TcArgs& t = ly.targs;
    t.n_phases = 1;
    t.lout = lout; t.ct = ct; t.cout = C; t.cg = ct;
    t.mode = TC_BIAS;
    TcPhase& ph = t.ph[0];
    ph.a = operand(&x);
    ph.b = operand(nullptr);
    ph.lin = L;
    ph.d_col = 0;
    const float* w = P(name + ".weight");
    if (!up) {
      // Conv1d(k=3, s=2, p=1): l_in = 2 l_out + tap - 1.  slots: 0 = tap 2, 1 = tap 0, 2 = tap 1
      ph.slots = 3;
      for (int li = 0; li < L; ++li) {
        if (li & 1) {
          ph.sched[li].slot_begin = 0;
          ph.sched[li].n_slots = (int8_t)(((li + 1) / 2 < lout) ? 2 : 1);
          ph.sched[li].lo_begin = (int8_t)((li - 1) / 2);
        } else {
          ph.sched[li].slot_begin = 2;
          ph.sched[li].n_slots = 1;
          ph.sched[li].lo_begin = (int8_t)(li / 2);
        }
      }
      static const int tap_of_slot[3] = {2, 0, 1};
      pack_tc(C, C, ct, 3, [&](int co, int ci, int sl) { return w[((size_t)co * C + ci) * 3 + tap_of_slot[sl]]; },
              &ph.w_hi, &ph.w_lo, &ph.acc_scale);
    } else {
      // ConvTranspose1d(k=4, s=2, p=1): l_out = 2 l_in - 1 + tap; weight [cin][cout][4]
      ph.slots = 4;
      for (int li = 0; li < L; ++li) {
        const int tmin = li == 0 ? 1 : 0;
        int tmax = 3;
        while (2 * li - 1 + tmax >= lout) --tmax;
        ph.sched[li].slot_begin = (int8_t)tmin;
        ph.sched[li].n_slots = (int8_t)(tmax - tmin + 1);
        ph.sched[li].lo_begin = (int8_t)(2 * li - 1 + tmin);
      }
      pack_tc(C, C, ct, 4, [&](int co, int ci, int sl) { return w[((size_t)ci * C + co) * 4 + sl]; }, &ph.w_hi,
              &ph.w_lo, &ph.acc_scale);
    }
    t.bias = vec(name + ".bias", C);
    t.out_hi = pm_out ? y.phi : y.thi;
    t.out_lo = pm_out ? y.plo : y.tlo;
    t.out_pm = pm_out ? 1 : 0;
    t.out_plain = y.p;
    ly.name = name;
    ly.macs_per_row = (up ? count_pairs<OP_UP4>(L, lout) : count_pairs<OP_DOWN3>(L, lout)) * C * C;
    finish_tc_layer(ly, C / ct);
    u->layers.push_back(ly);
    return y;
  }


  // ---------------- position-major tensor-core layers (conv_pm.cuh), horizon 25 / 50 ----------------
  static PmAct pm_operand(const Act* a) {
    PmAct o;
    o.hi = a ? a->phi : nullptr;
    o.lo = a ? a->plo : nullptr;
    o.C = a ? a->C : 0;
    return o;
  }

  // 16-bit element encode of (scaled) weight w into hi / lo
  void encode16(float w, uint16_t* h, uint16_t* l) const {
    if (u->tc_el == TC_EL_F16) {
      *h = f16_round(w);
      *l = f16_round(w - f16_to_float(*h));
    } else {
      *h = bf16_round(w);
      *l = bf16_round(w - bf16_to_float(*h));
    }
  }

  // weights of one slot group into [32-channel column tile][slot][k chunk][32 rows][2C bytes] swizzled
  // images (hi, lo); total_slots = slots per column tile; wfn(co, ci, slot).  Returns the exact inverse of
  // the power-of-two scale.
  template <class F>
  float pack_pm_slots(std::vector<uint8_t>& hi, std::vector<uint8_t>& lo, int slot0, int n_slots, int total_slots,
                      int cout, int C, int nkc, F wfn) {
    const int rby = 2 * C;
    float scale = 1.0f;
    if (u->tc_el == TC_EL_F16) {
      float wmax = 0.0f;
      for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < C * nkc; ++ci)
          for (int sl = 0; sl < n_slots; ++sl) wmax = std::max(wmax, std::fabs(wfn(co, ci, sl)));
      int e = 0;
      if (wmax > 0.0f) e = (int)std::floor(std::log2(16384.0f / wmax));
      e = std::max(-8, std::min(24, e));
      scale = std::ldexp(1.0f, e);
    }
    for (int sl = 0; sl < n_slots; ++sl)
      for (int kc = 0; kc < nkc; ++kc)
        for (int co = 0; co < cout; ++co)
          for (int c = 0; c < C; ++c) {
            const float w = scale * wfn(co, kc * C + c, sl);
            const int r8 = co & 7, chunk = (c * 2) >> 4;
            const int sw = rby == 128 ? (chunk ^ r8) : (rby == 64 ? (chunk ^ ((r8 >> 1) & 3)) : (chunk ^ ((r8 >> 2) & 1)));
            const int pct = u->pm2 ? cout : kPmCt;   // output channels per weight image (per CTA)
            const int nh = co / pct, cl = co % pct;
            const size_t off = ((size_t)((nh * total_slots + slot0 + sl) * nkc + kc) * pct + cl) * rby + (size_t)(sw << 4) +
                               ((c * 2) & 15);
            uint16_t h, l;
            encode16(w, &h, &l);
            std::memcpy(&hi[off], &h, 2);
            std::memcpy(&lo[off], &l, 2);
          }
    return 1.0f / scale;
  }

  void finish_pm_layer(Layer& ly, int atoms_needed, int n_slots) {
    PmArgs& t = ly.pargs;
    const int C = t.a.C, rby = 2 * C, atom = 8 * rby;
    const int nkc = t.b.C ? 2 : 1, nparts = u->tc_split ? 2 : 1;
    t.split = u->tc_split ? 1 : 0;
    t.a_bytes_img = (t.lin + 4) * atom;
    const int nimg = nkc * nparts;
    t.a_bytes_total = nimg * t.a_bytes_img;
    const int pct = u->pm2 ? t.cout : kPmCt;        // output channels per CTA
    t.w_bytes_part = n_slots * nkc * pct * rby;     // per column tile
    // the last M tile reads (finite garbage) past the last image: that tail may overlap the weights
    const size_t tail_end = (size_t)(nimg - 1) * t.a_bytes_img + (size_t)atoms_needed * atom;
    const int ntiles = (t.n_m + 15) / 16;
    int cols = (t.n_groups + t.aux) * ntiles * pct, p2 = 32;
    while (p2 < cols) p2 *= 2;
    // two-MMA form of the split product (PmArgs::fuse_b): needs twice the accumulator columns inside a 256-column buffer
    t.fuse_b = (u->pm2 && u->tc_split && !u->pm_pair && 2 * cols <= 256 && getenv("EDMP_NO_PM_FUSEB") == nullptr) ? 1 : 0;
    if (t.fuse_b) { cols *= 2; p2 = 32; while (p2 < cols) p2 *= 2; }
    t.tmem_cols = p2;
    if (p2 > (u->pm2 ? 256 : 512)) { ok = false; why += " pm:tmem(" + ly.name + ")"; }   // (persistent kernel: two accumulator buffers of 256 columns)
    if (u->pm2 && ntiles * t.n_terms * nkc * (C / 16) > kPm2MaxIssue) { ok = false; why += " pm:issue-table(" + ly.name + ")"; }
    ly.kind = LAYER_PM;
    ly.tc_smem = 1024 + std::max((((size_t)t.a_bytes_total + 1023) & ~(size_t)1023) + (size_t)nparts * t.w_bytes_part,
                                 tail_end) + 256;
    if (ly.tc_smem > (size_t)(232448 - (u->pm2 ? 13312 : 6144))) { ok = false; why += " pm:smem(" + ly.name + ")"; }
  }

  // Conv1dBlock (conv5 + GroupNorm + Mish [+ time embedding] [+ identity residual]) on the
  // position-major kernel.  aux_prefix != "": also emit the block's 1x1 residual conv of the input.
  Act pm_conv_block(const std::string& p, const Act& xa, const Act* xb, int cout, const std::string& out_name,
                    int temb_off, const Act* res, const std::string& aux_prefix, Act* aux_out, bool final_layer) {
    const int C = xa.C, nkc = xb ? 2 : 1, cin = C * nkc, L = xa.L;
    ok = ok && (!xb || xb->C == C) && (cout == 32 || cout == 64) && (C == 16 || C == 32 || C == 64);
    // (the final block's activation is also kept: it is one of the per-layer parity taps)
    Act y = new_act(u, out_name, cout, L, false, false, true);
    ok = ok && y.ok;
    Layer ly;
    std::memset(&ly.args, 0, sizeof(ConvArgs));
    std::memset(&ly.targs, 0, sizeof(TcArgs));
    std::memset(&ly.pargs, 0, sizeof(PmArgs));
    PmArgs& t = ly.pargs;
    t.a = pm_operand(&xa);
    t.b = pm_operand(xb);
    t.lin = L; t.n_m = L; t.lout = L; t.stride = 1;
    t.n_terms = 5;
    for (int j = 0; j < 5; ++j) { t.terms[j].acc = 0; t.terms[j].slot = (int8_t)j; t.terms[j].off = (int8_t)j; }
    t.n_groups = 1; t.out_step = 1; t.out_off[0] = 0; t.out_off[1] = 0;
    t.cout = cout;
    t.mode = res ? PM_GN_RES : PM_GN;
    const bool aux = !aux_prefix.empty();
    const int n_slots = aux ? 6 : 5;
    const int rby = 2 * C;
    std::vector<uint8_t> hi((size_t)n_slots * nkc * cout * rby), lo(hi.size());
    const float* w = P(p + ".block.0.weight");   // [cout][cin_real][5]; the packed input has C >= cin_real channels
    const int cin_real = pw->table.at(p + ".block.0.weight").d1;
    if (!playing())
      t.acc_scale = pack_pm_slots(hi, lo, 0, 5, n_slots, cout, C, nkc, [&](int co, int ci, int sl) {
        return ci < cin_real ? w[((size_t)co * cin_real + ci) * 5 + sl] : 0.0f;
      });
    blob_scalar(u, &t.acc_scale);
    if (aux) {
      const float* wr = P(aux_prefix + ".residual_conv.weight");   // [cout][cin_real][1]
      t.aux = 1;
      t.terms[5].acc = 1; t.terms[5].slot = 5; t.terms[5].off = 2;
      t.n_terms = 6;
      if (!playing())
        t.aux_scale = pack_pm_slots(hi, lo, 5, 1, n_slots, cout, C, nkc, [&](int co, int ci, int) {
          return ci < cin_real ? wr[(size_t)co * cin_real + ci] : 0.0f;
        });
      blob_scalar(u, &t.aux_scale);
      t.aux_bias = vec(aux_prefix + ".residual_conv.bias", cout);
      *aux_out = new_act(u, aux_prefix + ".residual_conv", cout, L, false, false, true);
      ok = ok && aux_out->ok;
      t.aux_hi = aux_out->phi;
      t.aux_lo = aux_out->plo;
    }
    (void)cin;
    const uint64_t wtag = geom_tag(n_slots, cout, C, nkc, u->pm2 ? cout : kPmCt);
    t.w_hi = upload_bytes(u, hi.data(), hi.size(), wtag);
    t.w_lo = u->tc_split ? upload_bytes(u, lo.data(), lo.size(), wtag) : nullptr;
    ok = ok && t.w_hi && (!u->tc_split || t.w_lo);
    t.bias = vec(p + ".block.0.bias", cout);
    t.gamma = vec(p + ".block.2.weight", cout);
    t.beta = vec(p + ".block.2.bias", cout);
    if (res) t.res = pm_operand(res);
    t.out_hi = y.phi;
    t.out_lo = y.plo;
    if (final_layer) {
      t.fw = vec("final_conv.1.weight", kDof * cout);
      t.fb = vec("final_conv.1.bias", kDof);
      ly.pm_final = true;
    }
    ly.temb_off = temb_off;
    ly.name = p;
    ly.macs_per_row = count_pairs<OP_CONV5>(L, L) * cin_real * cout + (aux ? (double)L * cin_real * cout : 0.0) +
                      (final_layer ? (double)L * cout * kDof : 0.0);
    const int ntiles = (L + 15) / 16;
    finish_pm_layer(ly, 16 * ntiles + 4, n_slots);
    u->layers.push_back(ly);
    return y;
  }

  // ResidualConvolutionBlock (blocks.py:154-166): blocks.0 also emits the 1x1 residual conv of the block
  // input when C_in != C_out, blocks.1 adds it (or the block input itself) as an identity residual
  Act pm_res_block(const std::string& p, const Act& xa, const Act* xb, int cout, int cin_real) {
    const int off = temb_width;
    temb_blocks.push_back({p, off});
    temb_width += cout;
    const bool proj = cin_real != cout;
    Act aux;
    Act h = pm_conv_block(p + ".blocks.0", xa, xb, cout, p + ".blocks.0", off, nullptr, proj ? p : "", &aux, false);
    return pm_conv_block(p + ".blocks.1", h, nullptr, cout, p, -1, proj ? &aux : &xa, "", nullptr, false);
  }

  // stride-2 Conv1d (blocks.py:211) / ConvTranspose1d (:249) on the position-major kernel; tc_out:
  // write the rows-as-M tiles of conv_tc.cuh (the 25 -> 13 hand-over to the deep levels)
  Act pm_resample(const std::string& name, const Act& x, bool up, bool tc_out) {
    const int C = x.C, L = x.L;
    const int lout = up ? 2 * L : (L + 1) / 2;
    Act y = new_act(u, name, C, lout, false, tc_out, !tc_out);
    ok = ok && y.ok && (C == 32 || C == 64);
    Layer ly;
    std::memset(&ly.args, 0, sizeof(ConvArgs));
    std::memset(&ly.targs, 0, sizeof(TcArgs));
    std::memset(&ly.pargs, 0, sizeof(PmArgs));
    PmArgs& t = ly.pargs;
    t.a = pm_operand(&x);
    t.b = pm_operand(nullptr);
    t.lin = L; t.lout = lout; t.cout = C; t.mode = PM_BIAS;
    const int rby = 2 * C;
    const float* w = P(name + ".weight");
    int atoms_needed, n_slots;
    if (!up) {
      // l_in = 2 l_out + tap - 1  ->  padded atom 2 l_out + tap + 1, stride 2
      t.n_m = lout; t.stride = 2; t.n_groups = 1; t.out_step = 1;
      t.n_terms = 3; n_slots = 3;
      for (int j = 0; j < 3; ++j) { t.terms[j].acc = 0; t.terms[j].slot = (int8_t)j; t.terms[j].off = (int8_t)(j + 1); }
      atoms_needed = 32 * ((lout + 15) / 16) + 2;
    } else {
      // l_out = 2 l_in - 1 + tap: even outputs 2m <- taps 1 (l_in = m), 3 (m - 1); odd 2m+1 <- taps 0 (m + 1), 2 (m)
      t.n_m = L; t.stride = 1; t.n_groups = 2; t.out_step = 2; t.out_off[0] = 0; t.out_off[1] = 1;
      t.n_terms = 4; n_slots = 4;
      const int8_t acc[4] = {0, 0, 1, 1}, slot[4] = {1, 3, 0, 2}, off[4] = {2, 1, 3, 2};
      for (int j = 0; j < 4; ++j) { t.terms[j].acc = acc[j]; t.terms[j].slot = slot[j]; t.terms[j].off = off[j]; }
      atoms_needed = 16 * ((L + 15) / 16) + 4;
    }
    std::vector<uint8_t> hi((size_t)n_slots * C * rby), lo(hi.size());
    if (!playing())
      t.acc_scale = pack_pm_slots(hi, lo, 0, n_slots, n_slots, C, C, 1, [&](int co, int ci, int sl) {
        return up ? w[((size_t)ci * C + co) * 4 + sl] : w[((size_t)co * C + ci) * 3 + sl];
      });
    blob_scalar(u, &t.acc_scale);
    const uint64_t wtag = geom_tag(n_slots, C, C, up ? 2 : 1, u->pm2 ? C : kPmCt);
    t.w_hi = upload_bytes(u, hi.data(), hi.size(), wtag);
    t.w_lo = u->tc_split ? upload_bytes(u, lo.data(), lo.size(), wtag) : nullptr;
    ok = ok && t.w_hi && (!u->tc_split || t.w_lo);
    t.bias = vec(name + ".bias", C);
    if (tc_out) { t.tc_hi = y.thi; t.tc_lo = y.tlo; }
    else { t.out_hi = y.phi; t.out_lo = y.plo; }
    ly.name = name;
    ly.macs_per_row = (up ? count_pairs<OP_UP4>(L, lout) : count_pairs<OP_DOWN3>(L, lout)) * C * C;
    finish_pm_layer(ly, atoms_needed, n_slots);
    u->layers.push_back(ly);
    return y;
  }

  // caller's x float32 [rows][7][50] -> position-major image, 16 channels (7 real)
  Act pm_pack_input(const Act& x) {
    Act y = new_act(u, "", 16, x.L, false, false, true);
    ok = ok && y.ok;
    Layer ly;
    std::memset(&ly.args, 0, sizeof(ConvArgs));
    std::memset(&ly.targs, 0, sizeof(TcArgs));
    std::memset(&ly.pargs, 0, sizeof(PmArgs));
    ly.kind = LAYER_PM_PACK;
    ly.pack_src = x;
    ly.pack_dst = y;
    ly.name = "input (position-major pack)";
    u->layers.push_back(ly);
    return y;
  }

  // plain -> tiled operand conversion at the CUDA-core / tensor-core boundary
  Act pack_to_tiled(const Act& x, const std::string& name) {
    Act y = new_act(u, "", x.C, x.L, false, true);
    ok = ok && y.ok;
    Layer ly;
    std::memset(&ly.args, 0, sizeof(ConvArgs));
    std::memset(&ly.targs, 0, sizeof(TcArgs));
    ly.kind = LAYER_PACK;
    ly.pack_src = x;
    ly.pack_dst = y;
    ly.name = name + " (tile)";
    u->layers.push_back(ly);
    return y;
  }

  Act resample(const std::string& name, const Act& x, bool up) {
    const int C = x.C;
    const int lout = up ? ((2 * x.L == 8 || 2 * x.L == 14 || 2 * x.L == 26) ? 2 * x.L - 1 : 2 * x.L)
                        : (x.L + 1) / 2;  // floor((L + 2 - 3) / 2) + 1
    Act y = new_act(u, name, C, lout);
    ok = ok && y.p;
    Layer ly;
    std::memset(&ly.args, 0, sizeof(ConvArgs));
    ly.fn = up ? pick_up(x.L, C) : pick_down(x.L, C);
    ok = ok && ly.fn;
    ly.args.xa = x.p; ly.args.ca = C;
    ly.args.w = up ? pack_convT(name + ".weight", C, C, 4) : pack_conv(name + ".weight", C, C, 3);
    ly.args.bias = vec(name + ".bias", C);
    ly.args.y = y.p;
    ly.args.cout = C;
    ly.name = name;
    ly.macs_per_row = (up ? count_pairs<OP_UP4>(x.L, lout) : count_pairs<OP_DOWN3>(x.L, lout)) * C * C;
    u->layers.push_back(ly);
    return y;
  }
};

static void build_launches(UNet* u, bool& ok);

// params != null: build from a state_dict (record != 0: keep the packed blob, unet_blob_*); params == null: build from
// `blob` (made by an earlier recording with the same dims / precision / layout version and a matching plan)
static int unet_create_impl(const float* params, size_t n_params, const int* dims, int n_dims, int precision,
                            int max_rows, const void* blob, size_t blob_bytes, bool record, UNet** out) {
  EDMP_REQUIRE(n_dims == 6 && dims[0] == 32 && dims[1] == 64 && dims[2] == 128 && dims[3] == 256 &&
                   dims[4] == 512 && dims[5] == 512,
               "only dims=(32,64,128,256,512,512) is compiled in (infer_serial.py:50)");
  EDMP_REQUIRE(precision >= EDMP_PRECISION_FP32 && precision <= EDMP_PRECISION_F16,
               "precision must be one of fp32, tf32x3, tf32, bf16x3, bf16, f16x3, f16");
  EDMP_REQUIRE(max_rows > 0, "max_rows must be positive");
  ParamWalker pw;
  walk_params(dims, n_dims, pw);
  EDMP_REQUIRE(pw.cursor == n_params, "parameter count does not match the state_dict layout");
  EDMP_REQUIRE((params != nullptr) != (blob != nullptr), "either a state_dict or a packed blob");

  UNet* u = new UNet();
  u->precision = precision;
  u->max_rows = max_rows;
  if (blob) {
    u->play = static_cast<const uint8_t*>(blob);
    u->play_bytes = blob_bytes;
    u->play_pos = (sizeof(BlobHeader) + 15) & ~(size_t)15;
  } else if (record) {
    u->rec = new std::vector<uint8_t>();
    BlobHeader h;
    std::memset(&h, 0, sizeof(h));
    std::memcpy(h.magic, "EDMPBLOB", 8);
    h.version = kBlobLayoutVersion; h.precision = (uint32_t)precision; h.n_dims = (uint32_t)n_dims;
    for (int i = 0; i < n_dims; ++i) h.dims[i] = (uint32_t)dims[i];
    h.max_rows = (uint32_t)max_rows; h.n_params = n_params;
    blob_append(*u->rec, &h, sizeof(h));
  }
  u->tc = precision != EDMP_PRECISION_FP32;
  u->tc_split = precision == EDMP_PRECISION_TF32X3 || precision == EDMP_PRECISION_BF16X3 ||
                precision == EDMP_PRECISION_F16X3;
  u->tc_el = (precision == EDMP_PRECISION_BF16X3 || precision == EDMP_PRECISION_BF16) ? TC_EL_BF16
             : (precision == EDMP_PRECISION_F16X3 || precision == EDMP_PRECISION_F16) ? TC_EL_F16 : TC_EL_TF32;
  if (u->tc) {
    static unsigned long long attr_set = 0;   // cudaFuncSetAttribute is per device
    if (once_per_device(&attr_set)) {
      EDMP_CK(cudaFuncSetAttribute(conv_tc_kernel<TC_EL_TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 10240));
      EDMP_CK(cudaFuncSetAttribute(conv_tc_kernel<TC_EL_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 10240));
      EDMP_CK(cudaFuncSetAttribute(conv_tc_kernel<TC_EL_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 10240));
      EDMP_CK(cudaFuncSetAttribute(conv_tc2_kernel<TC_EL_F16, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 7168));
      EDMP_CK(cudaFuncSetAttribute(conv_tc2_kernel<TC_EL_BF16, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 7168));
      EDMP_CK(cudaFuncSetAttribute(conv_tc2_kernel<TC_EL_F16, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 7168));
      EDMP_CK(cudaFuncSetAttribute(conv_tc2_kernel<TC_EL_BF16, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 7168));
      EDMP_CK(cudaFuncSetAttribute(conv_tc2_kernel<TC_EL_F16, 1, kT2MaxRun>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 7168));
      EDMP_CK(cudaFuncSetAttribute(conv_tc2_kernel<TC_EL_BF16, 1, kT2MaxRun>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 7168));
      EDMP_CK(cudaFuncSetAttribute(conv_tc2_kernel<TC_EL_F16, 2, kT2MaxRun>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 7168));
      EDMP_CK(cudaFuncSetAttribute(conv_tc2_kernel<TC_EL_BF16, 2, kT2MaxRun>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 7168));
      EDMP_CK(cudaFuncSetAttribute(conv_pm2_kernel<TC_EL_F16, 4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 13312));
      EDMP_CK(cudaFuncSetAttribute(conv_pm2_kernel<TC_EL_F16, 8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 13312));
      EDMP_CK(cudaFuncSetAttribute(conv_pm2_kernel<TC_EL_BF16, 4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 13312));
      EDMP_CK(cudaFuncSetAttribute(conv_pm2_kernel<TC_EL_BF16, 8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 13312));
      EDMP_CK(cudaFuncSetAttribute(conv_pm2_kernel<TC_EL_F16, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 13312));
      EDMP_CK(cudaFuncSetAttribute(conv_pm2_kernel<TC_EL_F16, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 13312));
      EDMP_CK(cudaFuncSetAttribute(conv_pm2_kernel<TC_EL_BF16, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 13312));
      EDMP_CK(cudaFuncSetAttribute(conv_pm2_kernel<TC_EL_BF16, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 13312));
      EDMP_CK(cudaFuncSetAttribute(conv_pm_kernel<TC_EL_F16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 6144 - 1024));
      EDMP_CK(cudaFuncSetAttribute(conv_pm_kernel<TC_EL_F16, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 6144 - 1024));
      EDMP_CK(cudaFuncSetAttribute(conv_pm_kernel<TC_EL_BF16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 6144 - 1024));
      EDMP_CK(cudaFuncSetAttribute(conv_pm_kernel<TC_EL_BF16, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 6144 - 1024));
    }
  }
  Builder b{u, params, &pw};
  std::vector<int> d(1, kDof);
  for (int i = 0; i < n_dims; ++i) d.push_back(dims[i]);

  Act x = new_act(u, "input", kDof, kHorizon);
  b.ok = b.ok && x.p;
  std::vector<Act> skips;
  const int n_down = (int)d.size() - 1;
  // Levels with horizon <= 13 (94 % of the MACs) run on the rows-as-M tensor-core kernel when
  // precision != fp32; the long-horizon, few-channel levels run on the position-major tensor-core
  // kernel when the operand elements are 16-bit, else on the CUDA-core kernels.
  u->pm = u->tc && u->tc_16() && getenv("EDMP_NO_PM") == nullptr;
  u->tc2 = u->pm && getenv("EDMP_TC_V1") == nullptr;
  u->pm2 = u->pm && getenv("EDMP_PM_V1") == nullptr;
  // (opt-in: measured 5 us per layer SLOWER at 8190 and at 1020 rows -- the position-major layers are bound by their
  // epilogues' instruction issue, not by the number of MMAs; DESIGN.md round log)
  u->pm_pair = u->pm2 && getenv("EDMP_PM_PAIR") != nullptr;
  // pairs halve the number of schedulable tiles: only worth it when the batch still fills the machine
  u->cg2 = u->tc2 && getenv("EDMP_NO_CG2") == nullptr && (max_rows >= 4096 || getenv("EDMP_CG2") != nullptr);
  u->narrow = u->tc2 && getenv("EDMP_NO_NARROW") == nullptr;
  {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      u->sm_count = n;
  }
  auto on_tc = [&](const Act& a) { return u->tc && a.L <= kTcMaxLin; };
  auto on_pm = [&](const Act& a) { return u->pm && a.L > kTcMaxLin; };
  if (u->pm) x = b.pm_pack_input(x);
  for (int i = 0; i < n_down; ++i) {
    const std::string p = "down_samplers." + std::to_string(i) + ".down.";
    if (on_pm(x)) {
      x = b.pm_res_block(p + "0", x, nullptr, d[i + 1], d[i]);
      x = b.pm_res_block(p + "1", x, nullptr, d[i + 1], d[i + 1]);
      skips.push_back(x);
      x = b.pm_resample(p + "3", x, false, /*tc_out=*/(x.L + 1) / 2 <= kTcMaxLin);
      continue;
    }
    if (on_tc(x)) {
      if (!x.thi) x = b.pack_to_tiled(x, p + "0");
      x = b.tc_res_block(p + "0", x, nullptr, d[i + 1]);
      x = b.tc_res_block(p + "1", x, nullptr, d[i + 1]);
    } else {
      x = b.res_block(p + "0", x, nullptr, d[i + 1]);
      x = b.res_block(p + "1", x, nullptr, d[i + 1]);
    }
    skips.push_back(x);
    if (i != n_down - 1) x = on_tc(x) ? b.tc_resample(p + "3", x, false, false) : b.resample(p + "3", x, false);
  }
  if (on_tc(x)) {
    x = b.tc_res_block("middle_block.middle.0", x, nullptr, d.back());
    x = b.tc_res_block("middle_block.middle.2", x, nullptr, d.back());
  } else {
    x = b.res_block("middle_block.middle.0", x, nullptr, d.back());
    x = b.res_block("middle_block.middle.2", x, nullptr, d.back());
  }
  int n = 0;
  for (int i = (int)d.size() - 1; i > 1; --i, ++n) {
    const std::string p = "up_samplers." + std::to_string(n) + ".up.";
    Act h = skips.back();
    skips.pop_back();
    if (on_pm(x)) {
      x = b.pm_res_block(p + "0", x, &h, d[i - 1], 2 * d[i]);
      x = b.pm_res_block(p + "1", x, nullptr, d[i - 1], d[i - 1]);
      x = b.pm_resample(p + "3", x, true, false);
    } else if (on_tc(x)) {
      x = b.tc_res_block(p + "0", x, &h, d[i - 1]);
      x = b.tc_res_block(p + "1", x, nullptr, d[i - 1]);
      // the up-sampled output feeds a position-major / CUDA-core level when it is longer than 13
      const int lup = (2 * x.L == 8 || 2 * x.L == 14 || 2 * x.L == 26) ? 2 * x.L - 1 : 2 * x.L;
      x = b.tc_resample(p + "3", x, true, /*plain_out=*/lup > kTcMaxLin && !u->pm, /*pm_out=*/lup > kTcMaxLin && u->pm);
    } else {
      x = b.res_block(p + "0", x, &h, d[i - 1]);
      x = b.res_block(p + "1", x, nullptr, d[i - 1]);
      x = b.resample(p + "3", x, true);
    }
  }
  if (on_pm(x)) {
    b.pm_conv_block("final_conv.0", x, nullptr, d[1], "final_conv.0", -1, nullptr, "", nullptr, /*final_layer=*/true);
    u->final_fused = true;
  } else {
    x = b.conv_block("final_conv.0", x, nullptr, d[1], "final_conv.0", -1, 0, nullptr, nullptr, "");
    u->final_in = x;
    u->final_c = d[1];
    u->final_w = b.vec("final_conv.1.weight", kDof * d[1]);
    u->final_b = b.vec("final_conv.1.bias", kDof);
  }

  if (u->tc_el == TC_EL_F16) {
    b.ok = b.ok && cudaMalloc(&u->range_flag, sizeof(unsigned)) == cudaSuccess;
    if (b.ok) { u->dev_allocs.push_back(u->range_flag); cudaMemset(u->range_flag, 0, sizeof(unsigned)); }
  }
  // progress counters of the chained conv_tc2 launches
  u->chain = u->tc2 && pdl_enabled() && getenv("EDMP_NO_CHAIN") == nullptr;
  u->tile_stride = ((max_rows + kTcRows - 1) / kTcRows + 1) & ~1;
  if (u->chain) {
    const size_t bytes = u->layers.size() * (size_t)u->tile_stride * sizeof(int);
    b.ok = b.ok && cudaMalloc(&u->tile_done, bytes) == cudaSuccess;
    if (b.ok) { u->dev_allocs.push_back(u->tile_done); cudaMemset(u->tile_done, 0, bytes); }
    for (size_t i = 0; i < u->layers.size() && b.ok; ++i) {
      Layer& ly = u->layers[i];
      if (ly.kind != LAYER_TC2) continue;
      ly.t2.done = u->tile_done + i * (size_t)u->tile_stride;
      if (i > 0 && u->layers[i - 1].kind == LAYER_TC2) {
        ly.t2.dep = u->tile_done + (i - 1) * (size_t)u->tile_stride;
        ly.t2.dep_target = u->layers[i - 1].t2.n_col_tiles;
      }
    }
  }

  build_launches(u, b.ok);

  // time-embedding table: TimeEmbedding (blocks.py:76-92) then each block's TimeMLP (:58-72)
  u->temb_width = b.temb_width;
  std::vector<float> table((size_t)kTSteps * b.temb_width);
  const double freq_scale = std::log(10000.0) / (16 - 1);
  for (int t = 1; t <= kTSteps && !b.playing(); ++t) {
    float emb[32], h1[128], te[32], mte[32];
    for (int i = 0; i < 16; ++i) {
      const float f = expf((float)i * (float)(-freq_scale));
      const float ang = (float)t * f;
      emb[i] = sinf(ang);
      emb[16 + i] = cosf(ang);
    }
    host_linear(b.P("time_embedding.time_mlp.1.weight"), b.P("time_embedding.time_mlp.1.bias"), emb, 128, 32, h1);
    for (int i = 0; i < 128; ++i) h1[i] = host_mish(h1[i]);
    host_linear(b.P("time_embedding.time_mlp.3.weight"), b.P("time_embedding.time_mlp.3.bias"), h1, 32, 128, te);
    for (int i = 0; i < 32; ++i) mte[i] = host_mish(te[i]);
    for (const auto& blk : b.temb_blocks) {
      const TensorRef& wr = pw.table.at(blk.first + ".time_mlp.time_mlp.1.weight");
      host_linear(params + wr.off, b.P(blk.first + ".time_mlp.time_mlp.1.bias"), mte, wr.d0, 32,
                  table.data() + (size_t)(t - 1) * b.temb_width + blk.second);
    }
  }
  u->temb = upload(u, table);
  b.ok = b.ok && u->temb;
  if (u->play && (u->play_bad || u->play_pos != ((u->play_bytes + 15) & ~(size_t)15))) {
    set_error("unet_create_from_blob: the blob does not match this engine's plan (packed for another batch-size class, "
              "kernel generation or layout): repack from the state_dict");
    unet_destroy(u);
    return 4;
  }
  if (!b.ok) {
    set_error("unet_create: allocation failed or an unsupported layer shape was requested:" + b.why);
    unet_destroy(u);
    return 1;
  }
  u->play = nullptr;
  u->n_launches = (int)u->launches.size() + (u->final_fused ? 0 : 1);
  *out = u;
  return 0;
}

void unet_destroy(UNet* u) {
  if (!u) return;
  for (void* p : u->dev_allocs) cudaFree(p);
  delete u->rec;
  delete u;
}


// Shared-memory geometry of a run of conv_tc2 layers [first, first + n): one carve-up for all of them (the most demanding
// layer decides).  Returns false when the layers cannot share a launch.
static bool plan_run(const UNet* u, int first, int n, UNet::Launch* out) {
  const Layer& l0 = u->layers[first];
  const int nparts = u->tc_split ? 2 : 1;
  const size_t a_stage = (size_t)kTcBlockBytes * nparts;
  int max_slots = 0;
  size_t part_bytes = 0, part_max = 0;   // weight bytes of one part of a stage (un-padded layout), GroupNorm piece area
  for (int i = first; i < first + n; ++i) {
    const Layer& ly = u->layers[i];
    if (ly.kind != LAYER_TC2 || ly.cta_group != l0.cta_group || ly.b_pad != l0.b_pad || ly.t2.mma_warps != l0.t2.mma_warps)
      return false;
    if (n > 1 && ly.nsplit != 1) return false;                                   // column-split layers launch alone (their cluster)
    if (l0.b_pad && (ly.t2_slot_bytes != l0.t2_slot_bytes || ly.t2_max_slots != l0.t2_max_slots)) return false;
    max_slots = std::max(max_slots, ly.t2_max_slots);
    part_bytes = std::max(part_bytes, (size_t)ly.t2_max_slots * ly.t2_slot_bytes);
    part_max = std::max(part_max, ly.t2_part_bytes);
  }
  const size_t slot_bytes = l0.t2_slot_bytes;
  auto b_bytes = [&](int k) {
    return l0.b_pad ? (size_t)nparts * ((size_t)k * (max_slots + 1) + 1) * slot_bytes : (size_t)k * part_bytes * nparts;
  };
  const size_t budget = 232448 - 1024 - 7168 - part_max;   // dynamic limit - alignment slack - static shared memory - pieces
  // (EDMP_MAX_B_STAGES / EDMP_MAX_A_STAGES: A/B knobs for the operand-feed experiments, profiles/r2_stage_sweep.txt)
  const int cap_b = getenv("EDMP_MAX_B_STAGES") ? std::max(1, atoi(getenv("EDMP_MAX_B_STAGES"))) : 3;
  const int cap_a = getenv("EDMP_MAX_A_STAGES") ? std::max(2, atoi(getenv("EDMP_MAX_A_STAGES"))) : kT2MaxAStages;
  int bst = 0;
  for (int k = std::min(3, cap_b); k >= 1 && !bst; --k)
    if (b_bytes(k) + (size_t)std::max(2, std::min(k, 3)) * a_stage <= budget) bst = k;
  if (!bst) return false;
  const int ast = (int)std::min<size_t>(std::min(kT2MaxAStages, cap_a), (budget - b_bytes(bst)) / a_stage);
  if (ast < 2) return false;
  out->first = first; out->n = n; out->run = true;
  out->a_stages = ast; out->b_stages = bst;
  if (l0.b_pad) {
    out->b_stage_stride = (int)((max_slots + 1) * slot_bytes);
    out->b_lo_off = (int)(((size_t)bst * (max_slots + 1) + 1) * slot_bytes);
    out->b_real_off = (int)slot_bytes;
    out->b_total_bytes = (int)(nparts * (size_t)out->b_lo_off);
  } else {
    out->b_stage_stride = (int)(part_bytes * nparts);
    out->b_lo_off = (int)part_bytes;
    out->b_real_off = 0;
    out->b_total_bytes = (int)(bst * part_bytes * nparts);
  }
  out->slot_bytes = (int)slot_bytes;
  out->smem = 1024 + (size_t)ast * a_stage + (size_t)out->b_total_bytes + part_max;
  return true;
}

// Groups the layers of a forward into launch units: maximal runs of consecutive conv_tc2 layers that can share a
// persistent launch (same cta_group, zero-slot layout and issuing warps; EDMP_NO_RUNS=1: one layer per launch).
static void build_launches(UNet* u, bool& ok) {
  // Runs need the row-tile chaining (the layers of a run synchronise through the progress counters only) and a batch
  // that fills the machine: with fewer tiles than SMs, separate launches put consecutive layers on DIFFERENT idle SMs
  // and overlap them completely, which one set of resident CTAs cannot (measured at 1020 rows: 2858 -> 2685 traj/s).
  const bool runs = getenv("EDMP_NO_RUNS") == nullptr && u->chain;
  const int rts = (u->max_rows + kTcRows - 1) / kTcRows;
  auto fills = [&](const Layer& ly) { return (rts / ly.cta_group) * ly.t2.n_col_tiles >= u->sm_count / ly.cta_group; };
  const int max_run = getenv("EDMP_MAX_RUN") ? std::max(1, std::min(kT2MaxRun, atoi(getenv("EDMP_MAX_RUN")))) : kT2MaxRun;
  const int nl = (int)u->layers.size();
  for (int i = 0; i < nl;) {
    UNet::Launch L;
    L.first = i; L.n = 1;
    if (u->layers[i].kind == LAYER_TC2) {
      int n = 1;
      UNet::Launch best;
      if (!plan_run(u, i, 1, &best)) { ok = false; return; }
      // (a run is only extended while its layers stay chained: a conv_tc2 layer always follows a conv_tc2 layer here)
      int own_a = u->layers[i].t2.a_stages, own_b = u->layers[i].t2.b_stages;   // fewest stages any layer of the run plans alone
      while (runs && n < max_run && i + n < nl && u->layers[i + n].kind == LAYER_TC2 && fills(u->layers[i]) &&
             fills(u->layers[i + n])) {
        UNet::Launch cand;
        if (!plan_run(u, i, n + 1, &cand)) break;
        // a longer run must not cost stages: the shared carve-up has to give what the neediest layer gets alone
        const int na = std::min(own_a, u->layers[i + n].t2.a_stages), nb = std::min(own_b, u->layers[i + n].t2.b_stages);
        if (cand.a_stages < na || cand.b_stages < nb) break;
        own_a = na; own_b = nb;
        best = cand;
        ++n;
      }
      L = best;
    }
    L.name = u->layers[L.first].name;
    if (L.n > 1) L.name += " .. " + u->layers[L.first + L.n - 1].name + " (" + std::to_string(L.n) + " layers)";
    for (int k = L.first; k < L.first + L.n; ++k) L.macs_per_row += u->layers[k].macs_per_row;
    u->launches.push_back(L);
    i += L.n;
  }
}

int unet_create(const float* params, size_t n_params, const int* dims, int n_dims, int precision,
                int max_rows, UNet** out) {
  return unet_create_impl(params, n_params, dims, n_dims, precision, max_rows, nullptr, 0, false, out);
}

// checkpoint -> engine + packed blob (reference temporalunet.py:78-92 is the checkpoint side)
int unet_pack(const float* params, size_t n_params, const int* dims, int n_dims, int precision, int max_rows,
              UNet** out) {
  return unet_create_impl(params, n_params, dims, n_dims, precision, max_rows, nullptr, 0, true, out);
}

int unet_blob_layout_version() { return (int)kBlobLayoutVersion; }
size_t unet_blob_bytes(const UNet* u) { return u->rec ? u->rec->size() : 0; }

// copies the recorded blob out and releases the host copy
int unet_blob_read(UNet* u, void* dst, size_t cap) {
  EDMP_REQUIRE(u->rec != nullptr, "this engine holds no packed blob (create it with edmp_unet_pack)");
  EDMP_REQUIRE(cap >= u->rec->size(), "destination too small for the packed blob");
  std::memcpy(dst, u->rec->data(), u->rec->size());
  delete u->rec;
  u->rec = nullptr;
  return 0;
}

int unet_create_from_blob(const void* blob, size_t bytes, int max_rows, UNet** out) {
  EDMP_REQUIRE(blob != nullptr && bytes >= sizeof(BlobHeader), "not a packed-weight blob");
  BlobHeader h;
  std::memcpy(&h, blob, sizeof(h));
  EDMP_REQUIRE(std::memcmp(h.magic, "EDMPBLOB", 8) == 0, "not a packed-weight blob (bad magic)");
  if (h.version != kBlobLayoutVersion) {
    set_error("unet_create_from_blob: stale blob (layout version " + std::to_string(h.version) + ", this library packs version " +
              std::to_string(kBlobLayoutVersion) + "): repack from the state_dict");
    return 4;
  }
  EDMP_REQUIRE(h.n_dims <= 8, "corrupt blob header");
  int dims[8];
  for (uint32_t i = 0; i < h.n_dims; ++i) dims[i] = (int)h.dims[i];
  return unet_create_impl(nullptr, (size_t)h.n_params, dims, (int)h.n_dims, (int)h.precision, max_rows, blob, bytes, false, out);
}

int unet_precision(const UNet* u) { return u->precision; }

// Reads and clears the operand-range flag (synchronises the stream): 1 = some activation of a forward since the last
// call exceeded the IEEE-half range and the results are not to be trusted (use tf32x3 / bf16x3 / fp32 for this model).
int unet_range_status(UNet* u, int* overflow, cudaStream_t st) {
  *overflow = 0;
  if (!u->range_flag) return 0;
  unsigned v = 0;
  EDMP_CK(cudaMemcpyAsync(&v, u->range_flag, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  EDMP_CK(cudaStreamSynchronize(st));
  if (v) EDMP_CK(cudaMemsetAsync(u->range_flag, 0, sizeof(unsigned), st));
  *overflow = v ? 1 : 0;
  return 0;
}
int unet_launches(const UNet* u) { return u->n_launches; }

template <int NR>
static void launch_tc2(UNet* u, int cg, int cl, dim3 grid, size_t smem, cudaStream_t st, const Tc2RunT<NR>& r) {
  if (cl > 1) {
    if (u->tc_el == TC_EL_F16) {
      if (cg == 2) launch_cluster(conv_tc2_kernel<TC_EL_F16, 2, NR>, grid, dim3(kT2Threads), smem, st, cl, r);
      else launch_cluster(conv_tc2_kernel<TC_EL_F16, 1, NR>, grid, dim3(kT2Threads), smem, st, cl, r);
    } else {
      if (cg == 2) launch_cluster(conv_tc2_kernel<TC_EL_BF16, 2, NR>, grid, dim3(kT2Threads), smem, st, cl, r);
      else launch_cluster(conv_tc2_kernel<TC_EL_BF16, 1, NR>, grid, dim3(kT2Threads), smem, st, cl, r);
    }
    return;
  }
  if (u->tc_el == TC_EL_F16) launch_pdl(conv_tc2_kernel<TC_EL_F16, 1, NR>, grid, dim3(kT2Threads), smem, st, r);
  else launch_pdl(conv_tc2_kernel<TC_EL_BF16, 1, NR>, grid, dim3(kT2Threads), smem, st, r);
}

// one persistent launch over the conv_tc2 layers [L.first, L.first + L.n)
static void run_tc2(UNet* u, const UNet::Launch& L, const float* temb_row, int rows, cudaStream_t st) {
  static Tc2Run r;   // (7 KB: filled per launch; the library is single-threaded per handle)
  const Layer& l0 = u->layers[L.first];
  const int cg = l0.cta_group;
  r.n = L.n;
  r.a_stages = L.a_stages; r.b_stages = L.b_stages;
  r.b_stage_stride = L.b_stage_stride; r.b_lo_off = L.b_lo_off; r.b_real_off = L.b_real_off; r.b_total_bytes = L.b_total_bytes;
  r.b_pad = l0.b_pad; r.slot_bytes = L.slot_bytes;
  r.mma_warps = l0.t2.mma_warps; r.split = u->tc_split ? 1 : 0;
  r.lean = l0.t2_lean;
  {
    // operand producer threads (conv_tc2.cuh): 1 = warp 0 issues both streams, 2 = the weight stream moves to warp 19,
    // 3 = (lean issuer only: warp 2 is free) the lo parts of the activation stream move to warp 2 as well
    const int want = getenv("EDMP_PRODUCERS") ? std::max(1, std::min(3, atoi(getenv("EDMP_PRODUCERS")))) : 2;
    r.producers = (want == 3 && !l0.t2_lean) ? 2 : want;
  }
  int max_tiles = 0;
  for (int k = 0; k < L.n; ++k) {
    const Layer& ly = u->layers[L.first + k];
    Tc2Args& a = r.l[k];
    a = ly.t2;
    a.rows = rows;
    a.temb = ly.temb_off >= 0 ? temb_row + ly.temb_off : nullptr;
    a.dbg = u->dbg;
    a.range_flag = u->range_flag;
    a.n_row_tiles = (rows + kTcRows - 1) / kTcRows;
    if (cg == 2) a.n_row_tiles = (a.n_row_tiles + 1) & ~1;
    max_tiles = std::max(max_tiles, (a.n_row_tiles / cg) * a.n_col_tiles);
  }
  const int cl = cg == 2 ? 2 : l0.nsplit;                              // cluster size
  const int walkers = std::min(max_tiles, u->sm_count / cl * cl / cg);   // CTA pairs (cg = 2) or CTAs that walk tiles
  // the tiles of all layers of the run form one round-robin sequence over the walkers
  long long g = 0;
  for (int k = 0; k < L.n; ++k) {
    r.rot[k] = (int)(g % walkers);
    g += (r.l[k].n_row_tiles / cg) * r.l[k].n_col_tiles;
  }
  dim3 grid(walkers * cg);
  if (L.n == 1) {
    // single layer: the small kernel parameter
    Tc2RunT<1> r1;
    std::memcpy(&r1, &r, offsetof(Tc2Run, rot));
    r1.rot[0] = 0;
    r1.l[0] = r.l[0];
    launch_tc2<1>(u, cg, cl, grid, L.smem, st, r1);
    return;
  }
  launch_tc2<kT2MaxRun>(u, cg, cl, grid, L.smem, st, r);
}

static void run_launch(UNet* u, const UNet::Launch& L, const float* x, const float* temb_row, int rows, float* eps, cudaStream_t st);

static void run_layer(UNet* u, Layer& ly, const float* x, const float* temb_row, int rows, float* eps, cudaStream_t st) {
  const float* input_act = u->acts.at("input").p;
  if (ly.kind == LAYER_PM) {
    PmArgs a = ly.pargs;
    a.rows = rows;
    a.temb = ly.temb_off >= 0 ? temb_row + ly.temb_off : nullptr;
    if (ly.pm_final) a.eps = eps;
    a.dbg = u->dbg;
    a.range_flag = u->range_flag;
    if (u->pm2) {
      const int n_blocks = (rows + kPmRows - 1) / kPmRows;
      if (u->pm_pair) {
        dim3 gridp(2 * std::min((n_blocks + 1) / 2, u->sm_count / 2));
        auto kp = u->tc_el == TC_EL_F16 ? (a.cout == 32 ? conv_pm2_kernel<TC_EL_F16, 4, 1> : conv_pm2_kernel<TC_EL_F16, 8, 1>)
                                        : (a.cout == 32 ? conv_pm2_kernel<TC_EL_BF16, 4, 1> : conv_pm2_kernel<TC_EL_BF16, 8, 1>);
        launch_cluster(kp, gridp, dim3(kPm2Threads), ly.tc_smem, st, 2, a);
        return;
      }
      dim3 grid2(std::min(n_blocks, u->sm_count));
      auto k2 = u->tc_el == TC_EL_F16 ? (a.cout == 32 ? conv_pm2_kernel<TC_EL_F16, 4, 0> : conv_pm2_kernel<TC_EL_F16, 8, 0>)
                                      : (a.cout == 32 ? conv_pm2_kernel<TC_EL_BF16, 4, 0> : conv_pm2_kernel<TC_EL_BF16, 8, 0>);
      launch_pdl(k2, grid2, dim3(kPm2Threads), ly.tc_smem, st, a);
      return;
    }
    dim3 grid((rows + kPmRows - 1) / kPmRows, a.cout / kPmCt);
    auto k = u->tc_el == TC_EL_F16 ? (a.cout == 32 ? conv_pm_kernel<TC_EL_F16, 4> : conv_pm_kernel<TC_EL_F16, 8>)
                                   : (a.cout == 32 ? conv_pm_kernel<TC_EL_BF16, 4> : conv_pm_kernel<TC_EL_BF16, 8>);
    launch_pdl(k, grid, dim3(kPmThreads), ly.tc_smem, st, a);
  } else if (ly.kind == LAYER_PM_PACK) {
    const int total = rows * ly.pack_src.L;
    auto k = u->tc_el == TC_EL_F16 ? pm_pack_input_kernel<TC_EL_F16> : pm_pack_input_kernel<TC_EL_BF16>;
    launch_pdl(k, dim3((total + 255) / 256), dim3(256), 0, st, x, rows, ly.pack_src.L, ly.pack_dst.phi, ly.pack_dst.plo, u->range_flag);
  } else if (ly.kind == LAYER_SIMT) {
    ConvArgs a = ly.args;
    a.rows = rows;
    if (a.xa == input_act) a.xa = x;   // the first block reads (and its residual re-reads) the caller's x
    if (a.ra == input_act) a.ra = x;
    a.temb = ly.temb_off >= 0 ? temb_row + ly.temb_off : nullptr;
    ly.fn(a, st);
  } else if (ly.kind == LAYER_TC) {
    TcArgs a = ly.targs;
    a.rows = rows;
    a.temb = ly.temb_off >= 0 ? temb_row + ly.temb_off : nullptr;
    a.dbg = u->dbg;
    dim3 grid((rows + kTcRows - 1) / kTcRows, ly.tc_tiles);
    if (u->tc_el == TC_EL_F16) launch_pdl(conv_tc_kernel<TC_EL_F16>, grid, dim3(kTcThreads), ly.tc_smem, st, a);
    else if (u->tc_el == TC_EL_BF16) launch_pdl(conv_tc_kernel<TC_EL_BF16>, grid, dim3(kTcThreads), ly.tc_smem, st, a);
    else launch_pdl(conv_tc_kernel<TC_EL_TF32>, grid, dim3(kTcThreads), ly.tc_smem, st, a);
  } else {
    const size_t total = (size_t)rows * ly.pack_src.L * (ly.pack_src.C / (u->tc_16() ? 8 : 4));
    const unsigned blocks = (unsigned)((total + 255) / 256);
    auto pack = u->tc_el == TC_EL_F16 ? tc_pack_kernel<TC_EL_F16>
                : u->tc_el == TC_EL_BF16 ? tc_pack_kernel<TC_EL_BF16> : tc_pack_kernel<TC_EL_TF32>;
    launch_pdl(pack, dim3(blocks), dim3(256), 0, st, (const float*)ly.pack_src.p, rows, ly.pack_src.C, ly.pack_src.L,
               ly.pack_dst.thi, ly.pack_dst.tlo);
  }
}

static void run_launch(UNet* u, const UNet::Launch& L, const float* x, const float* temb_row, int rows, float* eps, cudaStream_t st) {
  if (L.run) run_tc2(u, L, temb_row, rows, st);
  else run_layer(u, u->layers[L.first], x, temb_row, rows, eps, st);
}

bool unet_input_image(UNet* u, void** hi, void** lo, int* el, unsigned** range_flag) {
  if (u->launches.empty() || u->layers[0].kind != LAYER_PM_PACK || getenv("EDMP_NO_PACK_FOLD") != nullptr) return false;
  *hi = u->layers[0].pack_dst.phi;
  *lo = u->layers[0].pack_dst.plo;
  *el = u->tc_el;
  *range_flag = u->range_flag;
  return true;
}

static int unet_forward_impl(UNet* u, const float* x, int t, int rows, float* eps, bool packed, cudaStream_t st) {
  EDMP_REQUIRE(rows > 0 && rows <= u->max_rows, "rows exceeds the workspace (max_rows)");
  EDMP_REQUIRE(t >= 1 && t <= kTSteps, "t must be in 1..255");
  const float* temb_row = u->temb + (size_t)(t - 1) * u->temb_width;
  NvtxRange range("edmp_unet_forward");
  if (u->tile_done) EDMP_CK(cudaMemsetAsync(u->tile_done, 0, u->layers.size() * (size_t)u->tile_stride * sizeof(int), st));
  for (const UNet::Launch& L : u->launches) {
    if (packed && u->layers[L.first].kind == LAYER_PM_PACK) continue;   // the input image is already in place
    run_launch(u, L, x, temb_row, rows, eps, st);
  }
  if (!u->final_fused) {
    const int threads = 128;
    const size_t n = (size_t)rows * kHorizon;
    launch_pdl(final_pw_kernel, dim3((unsigned)((n + threads - 1) / threads)), dim3(threads),
               (7 * u->final_c + 7) * sizeof(float), st, (const float*)u->final_in.p, (const float*)u->final_w,
               (const float*)u->final_b, u->final_c, rows, eps);
  }
  EDMP_CK(cudaGetLastError());
  return 0;
}

int unet_forward(UNet* u, const float* x, int t, int rows, float* eps, cudaStream_t st) {
  return unet_forward_impl(u, x, t, rows, eps, false, st);
}
int unet_forward_packed(UNet* u, int t, int rows, float* eps, cudaStream_t st) {
  EDMP_REQUIRE(!u->launches.empty() && u->layers[0].kind == LAYER_PM_PACK, "this engine has no packed input image");
  return unet_forward_impl(u, nullptr, t, rows, eps, true, st);
}

// Per-op device time (CUDA events around every launch on the launching stream), averaged over
// `iters` forwards.  ms/macs have n_ops = unet_launches(u) entries (last = final 1x1 conv).
int unet_profile(UNet* u, const float* x, int t, int rows, int iters, float* ms, double* macs, float* eps,
                 cudaStream_t st) {
  EDMP_REQUIRE(rows > 0 && rows <= u->max_rows, "rows exceeds the workspace (max_rows)");
  EDMP_REQUIRE(iters > 0, "iters must be positive");
  const int nl = (int)u->launches.size();
  const int n = nl + (u->final_fused ? 0 : 1);
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) EDMP_CK(cudaEventCreate(&e));
  std::vector<std::vector<float>> samples(n);   // per launch: one duration per profiled forward
  const float* temb_row = u->temb + (size_t)(t - 1) * u->temb_width;
  for (int it = 0; it < iters; ++it) {
    if (u->tile_done) EDMP_CK(cudaMemsetAsync(u->tile_done, 0, u->layers.size() * (size_t)u->tile_stride * sizeof(int), st));
    EDMP_CK(cudaEventRecord(ev[0], st));
    for (int i = 0; i < nl; ++i) {
      run_launch(u, u->launches[i], x, temb_row, rows, eps, st);
      EDMP_CK(cudaEventRecord(ev[i + 1], st));
    }
    if (!u->final_fused) {
      const size_t ne = (size_t)rows * kHorizon;
      final_pw_kernel<<<(unsigned)((ne + 127) / 128), 128, (7 * u->final_c + 7) * sizeof(float), st>>>(
          u->final_in.p, u->final_w, u->final_b, u->final_c, rows, eps);
      EDMP_CK(cudaEventRecord(ev[n], st));
    }
    EDMP_CK(cudaStreamSynchronize(st));
    for (int i = 0; i < n; ++i) {
      float m = 0.f;
      EDMP_CK(cudaEventElapsedTime(&m, ev[i], ev[i + 1]));
      samples[i].push_back(m);
    }
  }
  for (int i = 0; i < n; ++i) {
    // the MEDIAN of the profiled forwards: one forward that catches a clock transition of the power-capped part (seen once:
    // a 600 us launch read as 837 us in a 10-forward mean) must not move the per-kernel figure
    std::sort(samples[i].begin(), samples[i].end());
    ms[i] = (iters & 1) ? samples[i][iters / 2] : 0.5f * (samples[i][iters / 2 - 1] + samples[i][iters / 2]);
    macs[i] = i < nl ? u->launches[i].macs_per_row * rows : (double)kHorizon * u->final_c * kDof * rows;
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return 0;
}

// debug: run op `op` alone `iters` times and return the clock64 stamps of its CTAs ([ctas][8])
int unet_tc_trace(UNet* u, int op, int rows, long long* out_h, int max_ctas, int* n_ctas, cudaStream_t st) {
  EDMP_REQUIRE(op >= 0 && op < (int)u->launches.size(), "no such launch");
  const UNet::Launch& L = u->launches[op];
  Layer& ly = u->layers[L.first];
  EDMP_REQUIRE(ly.kind == LAYER_TC || ly.kind == LAYER_PM || ly.kind == LAYER_TC2, "op is not a tensor-core layer");
  int ctas;
  if (ly.kind == LAYER_PM) {
    const int nb = (rows + kPmRows - 1) / kPmRows;
    ctas = u->pm_pair ? 2 * std::min((nb + 1) / 2, u->sm_count / 2) : u->pm2 ? std::min(nb, u->sm_count) : nb * (ly.pargs.cout / kPmCt);
  } else if (ly.kind == LAYER_TC2) {
    const int cg = ly.cta_group, cl = cg == 2 ? 2 : ly.nsplit;
    int rts = (rows + kTcRows - 1) / kTcRows, max_tiles = 0;
    if (cg == 2) rts = (rts + 1) & ~1;
    for (int k = 0; k < L.n; ++k) max_tiles = std::max(max_tiles, (rts / cg) * u->layers[L.first + k].t2.n_col_tiles);
    ctas = std::min(max_tiles, u->sm_count / cl * cl / cg) * cg;
  } else {
    ctas = ((rows + kTcRows - 1) / kTcRows) * ly.tc_tiles;
  }
  EDMP_REQUIRE(ctas <= max_ctas, "trace buffer too small");
  long long* d = nullptr;
  EDMP_CK(cudaMalloc(&d, (size_t)ctas * 16 * sizeof(long long)));
  EDMP_CK(cudaMemsetAsync(d, 0, (size_t)ctas * 16 * sizeof(long long), st));
  u->dbg = d;
  {
    const int ablate = getenv("EDMP_ABLATE") ? atoi(getenv("EDMP_ABLATE")) : 0;
    EDMP_CK(cudaMemcpyToSymbolAsync(g_edmp_ablate, &ablate, sizeof(int), 0, cudaMemcpyHostToDevice, st));
  }
  const float* temb_row = u->temb;
  float* eps_tmp = nullptr;
  if (ly.pm_final) EDMP_CK(cudaMalloc(&eps_tmp, (size_t)rows * kRowElems * sizeof(float)));
  for (int it = 0; it < 3; ++it) {
    // the run's own progress counters start from zero like in a forward (its first layer's producer counters keep the
    // full counts of the last forward: no wait on a kernel that is not running)
    if (u->tile_done && ly.kind == LAYER_TC2)
      EDMP_CK(cudaMemsetAsync(u->tile_done + (size_t)L.first * u->tile_stride, 0, (size_t)L.n * u->tile_stride * sizeof(int), st));
    run_launch(u, L, u->acts.at("input").p, temb_row, rows, eps_tmp, st);
  }
  u->dbg = nullptr;
  EDMP_CK(cudaMemcpyAsync(out_h, d, (size_t)ctas * 16 * sizeof(long long), cudaMemcpyDeviceToHost, st));
  EDMP_CK(cudaStreamSynchronize(st));
  cudaFree(d);
  if (eps_tmp) cudaFree(eps_tmp);
  *n_ctas = ctas;
  return 0;
}

const char* unet_op_name(const UNet* u, int i) {
  if (i < 0 || i > (int)u->launches.size()) return nullptr;
  return i == (int)u->launches.size() ? "final_conv.1" : u->launches[i].name.c_str();
}

const char* unet_op_kernel(const UNet* u, int i) {
  if (i < 0 || i > (int)u->launches.size()) return nullptr;
  if (i == (int)u->launches.size()) return "final_pw";
  const Layer& ly = u->layers[u->launches[i].first];
  switch (ly.kind) {
    case LAYER_TC2: return ly.cta_group == 2 ? "conv_tc2_pair" : "conv_tc2";
    case LAYER_TC: return "conv_tc";
    case LAYER_PM: return u->pm2 ? "conv_pm2" : "conv_pm";
    case LAYER_SIMT: return "conv_simt";
    default: return "pack";
  }
}

int unet_read_activation(UNet* u, const char* name, int rows, float* out, int* C, int* L, cudaStream_t st) {
  auto it = u->acts.find(name);
  EDMP_REQUIRE(it != u->acts.end(), std::string("unknown activation name: ") + name);
  EDMP_REQUIRE(rows > 0 && rows <= u->max_rows, "rows exceeds the workspace");
  if (C) *C = it->second.C;
  if (L) *L = it->second.L;
  if (out) {
    const Act& a = it->second;
    const size_t n = (size_t)rows * a.C * a.L;
    if (a.p) {
      EDMP_CK(cudaMemcpyAsync(out, a.p, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else if (a.phi) {
      auto unpack = u->tc_el == TC_EL_F16 ? pm_unpack_kernel<TC_EL_F16> : pm_unpack_kernel<TC_EL_BF16>;
      unpack<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a.phi, a.plo, rows, a.C, a.L, out);
      EDMP_CK(cudaGetLastError());
    } else {
      auto unpack = u->tc_el == TC_EL_F16 ? tc_unpack_kernel<TC_EL_F16>
                    : u->tc_el == TC_EL_BF16 ? tc_unpack_kernel<TC_EL_BF16> : tc_unpack_kernel<TC_EL_TF32>;
      unpack<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a.thi, a.tlo, rows, a.C, a.L, out);
      EDMP_CK(cudaGetLastError());
    }
  }
  return 0;
}

}  // namespace edmp
