// Self-test of the tcgen05 plumbing: runs D[128 x N] = sum_k A_k * B_k with caller-supplied raw
// shared-memory operand images and descriptor fields, so that the operand layouts / descriptor
// encodings used by the coupling kernels are validated on the GPU against a host matmul
// (tests/test_umma_gpu.py) independently of the big kernels.
#include "common.cuh"
#include "umma.cuh"

namespace {

__global__ void __launch_bounds__(128)
umma_selftest_kernel(const uint4* __restrict__ a_img, int a_bytes, const uint4* __restrict__ b_img, int b_bytes,
                     unsigned long long a_templ, unsigned long long b_templ, unsigned int idesc, int num_k,
                     int a_kstep, int b_kstep, int ncols, int use_bulk, float* __restrict__ d_out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar_mma, bar_load;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t base = (umma::smem_u32(smem) + 1023u) & ~1023u;
  unsigned char* A = smem + (base - umma::smem_u32(smem));
  unsigned char* B = A + ((a_bytes + 1023) & ~1023);

  if (tid == 0) {
    umma::mbar_init(&bar_mma, 1);
    umma::mbar_init(&bar_load, 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 256);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (use_bulk) {
    if (tid == 0) {
      umma::mbar_expect_tx(&bar_load, (uint32_t)(a_bytes + b_bytes));
      umma::bulk_g2s(A, a_img, (uint32_t)a_bytes, &bar_load);
      umma::bulk_g2s(B, b_img, (uint32_t)b_bytes, &bar_load);
    }
    umma::mbar_wait(&bar_load, 0);
  } else {
    for (int i = tid; i < a_bytes / 16; i += 128) reinterpret_cast<uint4*>(A)[i] = a_img[i];
    for (int i = tid; i < b_bytes / 16; i += 128) reinterpret_cast<uint4*>(B)[i] = b_img[i];
    umma::fence_async_smem();
  }
  __syncthreads();
  if (tid == 0) {
    umma::fence_after_sync();
    for (int k = 0; k < num_k; ++k) {
      const uint64_t da = umma::desc_at(a_templ, umma::smem_u32(A) + k * a_kstep);
      const uint64_t db = umma::desc_at(b_templ, umma::smem_u32(B) + k * b_kstep);
      umma::mma_bf16(tmem, da, db, idesc, k > 0 ? 1u : 0u);
    }
    umma::mma_commit(&bar_mma);
  }
  umma::mbar_wait(&bar_mma, 0);
  umma::fence_after_sync();
  for (int c0 = 0; c0 < ncols; c0 += 32) {
    float v[32];
    umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) d_out[(size_t)tid * ncols + c0 + i] = v[i];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

}  // namespace

// a_img / b_img: device buffers holding the exact shared-memory bytes of the operand tiles.
// a_templ / b_templ: 64-bit descriptor without the start address; *_kstep: bytes added per K step.
DPF_API int dpf_umma_selftest(const void* a_img, int a_bytes, const void* b_img, int b_bytes,
                              unsigned long long a_templ, unsigned long long b_templ, unsigned int idesc,
                              int num_k, int a_kstep, int b_kstep, int ncols, int use_bulk, float* d_out,
                              void* stream) {
  DPF_REQUIRE(a_img && b_img && d_out, DPF_ERR_NULL_PTR, "dpf_umma_selftest: null pointer");
  DPF_REQUIRE(a_bytes > 0 && b_bytes > 0 && a_bytes % 16 == 0 && b_bytes % 16 == 0 && a_bytes + b_bytes <= 160 * 1024,
              DPF_ERR_BAD_ARG, "dpf_umma_selftest: bad image sizes");
  DPF_REQUIRE(ncols > 0 && ncols <= 256 && ncols % 32 == 0 && num_k > 0, DPF_ERR_BAD_ARG, "dpf_umma_selftest: bad shape");
  const size_t smem = (size_t)((a_bytes + 1023) & ~1023) + (size_t)((b_bytes + 1023) & ~1023) + 1024;
  cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const uint4*)a_img, a_bytes, (const uint4*)b_img, b_bytes,
                                                             a_templ, b_templ, idesc, num_k, a_kstep, b_kstep, ncols,
                                                             use_bulk, d_out);
  return dpf_check_launch("umma_selftest_kernel");
}

int tc_debug_occupancy(int which, int smem);
// Debug probe: occupancy the runtime reports for the merged (which = 0) / plain (1) forward kernel.
DPF_API int dpf_debug_occupancy(int which, int smem, int* out) {
  DPF_REQUIRE(out, DPF_ERR_NULL_PTR, "dpf_debug_occupancy: null pointer");
  *out = tc_debug_occupancy(which, smem);
  return DPF_OK;
}
