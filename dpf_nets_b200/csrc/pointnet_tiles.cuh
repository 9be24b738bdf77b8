// Operand tile builders shared by the train-mode PointNet kernels (pointnet_layers.cu, pointnet_train.cu): a tile of
// C channel rows x 64 points of a (B, C, N) fp32 tensor -> bf16 hi | lo images with 128-byte swizzled rows (row = channel,
// the 64 bf16 of a row = the tile's 64 points), with a per-channel function applied on the way.
#pragma once
#include "umma.cuh"

namespace pnt {

constexpr int T = 256;                          // threads per CTA
constexpr int NT = 64;                          // points per tile
constexpr int TAB = 8;                          // floats per channel in a loader table

enum { LD_X3 = 0, LD_AFFINE = 1, LD_BNBWD = 2, LD_RAW = 3 };
// loader tables, TAB floats per channel:
//   LD_X3     {a0, a1, a2, c, sub}:       v = relu(a0 x0 + a1 x1 + a2 x2 + c) - sub          (layer 0 with its BatchNorm folded in)
//   LD_AFFINE {sc, sh, sub}:              v = relu(sc z + sh) - sub                          (sub: centring for the Gram form)
//   LD_BNBWD  {sc, sh, g, gm1, k2, mu}:   v = (sc z + sh > 0 ? g dA : 0) - gm1 - k2 (z - mu) (BatchNorm + ReLU backward)
//   LD_RAW    no table:                   v = z
template <int LOADER> struct LoaderInputs { static constexpr int n = LOADER == LD_X3 ? 3 : LOADER == LD_BNBWD ? 2 : 1; };

template <int LOADER>
__device__ __forceinline__ float loader_value(const float t[6], float i0, float i1, float i2) {
  if (LOADER == LD_X3) return fmaxf(fmaf(t[0], i0, fmaf(t[1], i1, fmaf(t[2], i2, t[3]))), 0.f) - t[4];
  if (LOADER == LD_AFFINE) return fmaxf(fmaf(t[0], i0, t[1]), 0.f) - t[2];
  if (LOADER == LD_RAW) return i0;
  const float y = fmaf(t[0], i1, t[1]);      // LD_BNBWD: i0 = dA, i1 = z
  return (y > 0.f ? t[2] * i0 : 0.f) - t[3] - t[4] * (i1 - t[5]);
}

__device__ __forceinline__ void store_split_chunk(unsigned char* hi_img, unsigned char* lo_img, uint32_t off, const float v[8]) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    hi[e] = umma::pack_bf16(v[2 * e], v[2 * e + 1]);
    lo[e] = umma::pack_bf16(v[2 * e] - __uint_as_float(hi[e] << 16), v[2 * e + 1] - __uint_as_float(hi[e] & 0xffff0000u));
  }
  *reinterpret_cast<uint4*>(hi_img + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(lo_img + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor rows: COALESCED loads.  A warp-wide 16-byte load covers two channel rows x 256 contiguous bytes (lanes 0-15 the 64
// points of one row, lanes 16-31 of the next) - a thread-per-row mapping touches 32 cache lines per instruction and stalls
// in the L1 tag stage (ncu: lg_throttle).  A 16-byte shared-memory chunk needs 8 consecutive points of ONE row, i.e. the
// float4s of two neighbouring lanes: rows are processed in pairs of load iterations (i, i + 1) and neighbouring lanes swap
// halves with one shuffle each, so that the even lane owns 8 points of the iteration-i row and the odd lane 8 points of the
// iteration-(i + 1) row.  The row of a (pair, lane) slot is the same for every tile.
// ---------------------------------------------------------------------------------------------------------------
template <int C, int LOADER, int NW = T / 32>
struct TileLoader {
  static constexpr int RPI = 2 * NW;              // rows per load iteration (2 per warp)
  static constexpr int NIT = C / RPI;             // load iterations per tile
  static constexpr int NIN = LoaderInputs<LOADER>::n;
  static_assert(C == 64 || C == 128 || C == 256, "channel rows per tile");
  static_assert(LOADER != LD_X3, "layer 0 has its own builder");
  static_assert(NIT >= 2 && NIT % 2 == 0, "rows are processed in pairs of load iterations");
  float4 buf[NIN][NIT];
  const float* in[2];
  const float* tab;
  int N, n_tiles, warp, lane;
  bool vec_ok;

  __device__ __forceinline__ void init(const float* in0, const float* in1, const float* tab_, int Nn, int tid) {
    in[0] = in0;
    in[1] = in1;
    tab = tab_;
    N = Nn;
    n_tiles = (Nn + NT - 1) / NT;
    warp = tid >> 5;
    lane = tid & 31;
    vec_ok = (N % 4) == 0;
  }
  __device__ __forceinline__ int slot_row(int pair) const { return (2 * pair + (lane & 1)) * RPI + warp * 2 + (lane >> 4); }
  // flat tile index over all shapes: tile -> (shape tile / n_tiles, first point (tile % n_tiles) * 64)
  __device__ __forceinline__ void load(int tile) {
    const int b = tile / n_tiles;
    const int n = (tile - b * n_tiles) * NT + 4 * (lane & 15);
#pragma unroll
    for (int j = 0; j < NIN; ++j) {
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        const float* src = in[j] + ((size_t)b * C + i * RPI + warp * 2 + (lane >> 4)) * N + n;
        if (vec_ok && n + 4 <= N) {
          buf[j][i] = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = (n + e < N) ? __ldg(src + e) : 0.f;
          buf[j][i] = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
  }
  // values of this thread's 8 points of slot `pair` (row slot_row(pair), points chunk * 8 ..): swap halves with the neighbour
  __device__ __forceinline__ void slot_values(int pair, float t[6], float v[8]) const {
    const bool odd = lane & 1;
    float4 a[2][NIN];           // [first / second 4 points][input]
#pragma unroll
    for (int j = 0; j < NIN; ++j) {
      const float4 mine0 = buf[j][2 * pair], mine1 = buf[j][2 * pair + 1];
      const float4 send = odd ? mine0 : mine1;
      float4 recv;
      recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
      recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
      recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1);
      recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1);
      a[0][j] = odd ? recv : mine0;
      a[1][j] = odd ? mine1 : recv;
    }
    if (LOADER != LD_RAW) {
      const float4* tr = reinterpret_cast<const float4*>(tab + (size_t)slot_row(pair) * TAB);
      const float4 t0 = __ldg(tr), t1 = __ldg(tr + 1);
      t[0] = t0.x; t[1] = t0.y; t[2] = t0.z; t[3] = t0.w; t[4] = t1.x; t[5] = t1.y;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 i0 = a[h][0], i1 = a[h][NIN > 1 ? 1 : 0];
      v[4 * h + 0] = loader_value<LOADER>(t, i0.x, i1.x, 0.f);
      v[4 * h + 1] = loader_value<LOADER>(t, i0.y, i1.y, 0.f);
      v[4 * h + 2] = loader_value<LOADER>(t, i0.z, i1.z, 0.f);
      v[4 * h + 3] = loader_value<LOADER>(t, i0.w, i1.w, 0.f);
    }
  }
  // bf16 hi | lo images of the tile; points beyond N are ZERO operands when `zero_tail` (they would otherwise enter sums
  // over points).  acc (optional, NIT / 2 floats): per-slot running sums of the operand values over the valid points.
  template <bool ZERO_TAIL, bool SUMS>
  __device__ __forceinline__ void convert(unsigned char* hi_img, unsigned char* lo_img, int tile, float* acc) const {
    const int chunk = (lane & 15) >> 1;
    const int base = (tile % n_tiles) * NT + chunk * 8;
#pragma unroll
    for (int pair = 0; pair < NIT / 2; ++pair) {
      float t[6], v[8];
      slot_values(pair, t, v);
      if ((ZERO_TAIL || SUMS) && base + 8 > N) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (base + e >= N) v[e] = 0.f;
      }
      if (SUMS) acc[pair] += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
      store_split_chunk(hi_img, lo_img, umma::sw128_offset(slot_row(pair), chunk), v);
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Layer 0 recomputed from the three coordinates: 64 channel rows, thread = (channel, 16 consecutive points); the
// coordinate loads are warp-wide broadcasts
// ---------------------------------------------------------------------------------------------------------------
struct X3TileLoader {
  static constexpr int C = 64, PPT = 16, NV = 4;
  float4 buf[3][NV];
  float t[6];
  int ch, seg, N, n_tiles;
  const float* x;
  bool vec_ok;

  __device__ __forceinline__ void init(const float* in0, const float*, const float* tab, int Nn, int tid) {
    ch = tid % C;
    seg = tid / C;
    N = Nn;
    n_tiles = (Nn + NT - 1) / NT;
#pragma unroll
    for (int j = 0; j < 6; ++j) t[j] = tab[ch * TAB + j];
    x = in0;
    vec_ok = (N % 4) == 0;
  }
  __device__ __forceinline__ void load(int tile) {
    const int b = tile / n_tiles;
    const int base = (tile - b * n_tiles) * NT + seg * PPT;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float* src = x + ((size_t)b * 3 + j) * N;
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        const int n = base + q * 4;
        if (vec_ok && n + 4 <= N) {
          buf[j][q] = __ldg(reinterpret_cast<const float4*>(src + n));
        } else {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = (n + e < N) ? __ldg(src + n + e) : 0.f;
          buf[j][q] = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
  }
  template <bool ZERO_TAIL, bool SUMS>
  __device__ __forceinline__ void convert(unsigned char* hi_img, unsigned char* lo_img, int tile, float*) const {
    const int base = (tile % n_tiles) * NT + seg * PPT;
#pragma unroll
    for (int q = 0; q < NV / 2; ++q) {
      float v[8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 a0 = buf[0][2 * q + h], a1 = buf[1][2 * q + h], a2 = buf[2][2 * q + h];
        v[4 * h + 0] = loader_value<LD_X3>(t, a0.x, a1.x, a2.x);
        v[4 * h + 1] = loader_value<LD_X3>(t, a0.y, a1.y, a2.y);
        v[4 * h + 2] = loader_value<LD_X3>(t, a0.z, a1.z, a2.z);
        v[4 * h + 3] = loader_value<LD_X3>(t, a0.w, a1.w, a2.w);
      }
      if (ZERO_TAIL && base + q * 8 + 8 > N) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (base + q * 8 + e >= N) v[e] = 0.f;
      }
      store_split_chunk(hi_img, lo_img, umma::sw128_offset(ch, seg * (PPT / 8) + q), v);
    }
  }
};

template <int C, int LOADER> struct LoaderFor { using type = TileLoader<C, LOADER>; };
template <int C> struct LoaderFor<C, LD_X3> { using type = X3TileLoader; };

}  // namespace pnt
