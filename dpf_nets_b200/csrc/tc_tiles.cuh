// Tile constants and the forward UMMA chain shared by the tensor-core coupling kernels
// (coupling_tc.cu: per-layer train / backward kernels; coupling_eval.cu: fused eval decoder).
#pragma once
#include "coupling.cuh"
#include "umma.cuh"

namespace {

constexpr int F = DPF_F;
constexpr uint32_t IMG_W = F * F * 2;           // 8 KB  : one 64x64 bf16 weight image
constexpr uint32_t IMG_H = DPF_TILE * F * 2;    // 16 KB : one 128x64 bf16 activation tile
constexpr int N_IMG = 3;                         // weight images per (layer, branch): W1 hi, W1 lo, W1^T hi
constexpr uint64_t DESC_K = umma::make_desc_template(16, 1024, umma::LAYOUT_SW128);        // K-major SW128
constexpr uint64_t DESC_MN = umma::make_desc_template(IMG_H, 1024, umma::LAYOUT_SW128);    // MN-major, 64-blocks IMG_H apart
constexpr uint32_t IDESC_GEMM = umma::make_idesc_bf16(128, 64, 0, 0);
constexpr uint32_t IDESC_WGRAD = umma::make_idesc_bf16(128, 128, 1, 1);

__device__ __forceinline__ float pick3(const float v[3], int i) { return i == 0 ? v[0] : (i == 1 ? v[1] : v[2]); }

// packed fp32x2 arithmetic of sm_100a (SASS FMUL2 / FFMA2): two IEEE fp32 lanes per instruction, each rounded like the
// scalar instruction - used for the per-channel-pair epilogue math
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// forward UMMA chain of ONE branch into TMEM columns [tcol, tcol+64): K = 64 in 4 steps,
// SPLIT: hi*hi + lo*hi + hi*lo
template <bool SPLIT>
__device__ __forceinline__ void issue_gemm1(uint32_t tcol, const unsigned char* Hhi, const unsigned char* Hlo,
                                            const unsigned char* Whi, const unsigned char* Wlo) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    umma::mma_bf16(tcol, umma::desc_at(DESC_K, umma::smem_u32(Hhi) + 32 * k), umma::desc_at(DESC_K, umma::smem_u32(Whi) + 32 * k),
                   IDESC_GEMM, k > 0);
  if (SPLIT) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma::mma_bf16(tcol, umma::desc_at(DESC_K, umma::smem_u32(Hlo) + 32 * k), umma::desc_at(DESC_K, umma::smem_u32(Whi) + 32 * k),
                     IDESC_GEMM, 1u);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma::mma_bf16(tcol, umma::desc_at(DESC_K, umma::smem_u32(Hhi) + 32 * k), umma::desc_at(DESC_K, umma::smem_u32(Wlo) + 32 * k),
                     IDESC_GEMM, 1u);
  }
}

}  // namespace
