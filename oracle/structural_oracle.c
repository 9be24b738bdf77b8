/*
 * structural_oracle.c — TEST INFRASTRUCTURE ONLY (parity checker; never linked into the product).
 *
 * Plain-C restatement of the reference's structural-loss CUDA kernels
 * (lib/metrics/pytorch_structural_losses/src/nndistance.cu, approxmatch.cu), following the
 * reference's loop order and chunking so that fp32 results are reproducible:
 *   - nvcc contracts x*x+y*y+z*z to fma(z,z,fma(x,x,y*y)) (FMUL y*y; FFMA x,x; FFMA z,z in the
 *     reference kernel's sm_100a SASS); fmaf() reproduces that bit-exactly.
 *   - approxmatch sums run in ascending index order like the reference's sequential inner loops;
 *     __expf (GPU fast exp) is restated with expf, so EMD parity is to a tolerance, not bit-exact.
 * Pinned by: the reference kernels themselves built into oracle/_ref (see build_ref.py) and a
 * numpy brute force, both compared in tests/.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline float sq3(float x2, float y2, float z2) {
  return fmaf(z2, z2, fmaf(x2, x2, y2 * y2));
}

/* NmDistanceKernel (nndistance.cu:2-124): chunks of 512 targets, first element of every chunk
 * taken unconditionally, strict '<' afterwards; cross-chunk merge `k2==0 || result > best`. */
static void nm_one_direction(int b, int n, const float* xyz, int m, const float* xyz2, float* result,
                             int* result_i) {
  const int batch = 512;
  for (int i = 0; i < b; ++i) {
    for (int j = 0; j < n; ++j) {
      const float x1 = xyz[(i * (size_t)n + j) * 3 + 0];
      const float y1 = xyz[(i * (size_t)n + j) * 3 + 1];
      const float z1 = xyz[(i * (size_t)n + j) * 3 + 2];
      for (int k2 = 0; k2 < m; k2 += batch) {
        const int end_k = (m < k2 + batch ? m : k2 + batch) - k2;
        const float* buf = xyz2 + (i * (size_t)m + k2) * 3;
        int best_i = 0;
        float best = 0;
        for (int k = 0; k < end_k; ++k) {
          const float d = sq3(buf[k * 3 + 0] - x1, buf[k * 3 + 1] - y1, buf[k * 3 + 2] - z1);
          if (k == 0 || d < best) {
            best = d;
            best_i = k + k2;
          }
        }
        if (k2 == 0 || result[i * (size_t)n + j] > best) {
          result[i * (size_t)n + j] = best;
          result_i[i * (size_t)n + j] = best_i;
        }
      }
    }
  }
}

/* nndistance (nndistance.cu:125-128): two launches with the roles swapped. */
void oracle_nndistance(int b, int n, const float* xyz, int m, const float* xyz2, float* result,
                       int* result_i, float* result2, int* result2_i) {
  nm_one_direction(b, n, xyz, m, xyz2, result, result_i);
  nm_one_direction(b, m, xyz2, n, xyz, result2, result2_i);
}

/* NmDistanceGradKernel + nndistancegrad (nndistance.cu:129-154). */
static void nm_grad_one(int b, int n, const float* xyz1, int m, const float* xyz2, const float* grad_dist1,
                        const int* idx1, float* grad_xyz1, float* grad_xyz2) {
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < n; ++j) {
      const size_t e = i * (size_t)n + j;
      const int j2 = idx1[e];
      const float* p1 = xyz1 + e * 3;
      const float* p2 = xyz2 + (i * (size_t)m + j2) * 3;
      const float g = grad_dist1[e] * 2;
      for (int c = 0; c < 3; ++c) {
        grad_xyz1[e * 3 + c] += g * (p1[c] - p2[c]);
        grad_xyz2[(i * (size_t)m + j2) * 3 + c] += -(g * (p1[c] - p2[c]));
      }
    }
}

void oracle_nndistance_grad(int b, int n, const float* xyz1, int m, const float* xyz2,
                            const float* grad_dist1, const int* idx1, const float* grad_dist2,
                            const int* idx2, float* grad_xyz1, float* grad_xyz2) {
  memset(grad_xyz1, 0, sizeof(float) * (size_t)b * n * 3);
  memset(grad_xyz2, 0, sizeof(float) * (size_t)b * m * 3);
  nm_grad_one(b, n, xyz1, m, xyz2, grad_dist1, idx1, grad_xyz1, grad_xyz2);
  nm_grad_one(b, m, xyz2, n, xyz1, grad_dist2, idx2, grad_xyz2, grad_xyz1);
}

/* pairwise_CD (lib/networks/utils.py:90-117) for one (i,j): dl.mean + dr.mean with fp32 sums. */
void oracle_pairwise_cd(int S1, int S2, int n, int m, const float* A, const float* B, float* out) {
  float* d1 = (float*)malloc(sizeof(float) * n);
  float* d2 = (float*)malloc(sizeof(float) * m);
  int* i1 = (int*)malloc(sizeof(int) * n);
  int* i2 = (int*)malloc(sizeof(int) * m);
  for (int i = 0; i < S1; ++i)
    for (int j = 0; j < S2; ++j) {
      oracle_nndistance(1, n, A + (size_t)i * n * 3, m, B + (size_t)j * m * 3, d1, i1, d2, i2);
      double s1 = 0, s2 = 0; /* double accumulation: the mean is compared to a tolerance */
      for (int k = 0; k < n; ++k) s1 += d1[k];
      for (int k = 0; k < m; ++k) s2 += d2[k];
      out[(size_t)i * S2 + j] = (float)(s1 / n) + (float)(s2 / m);
    }
  free(d1); free(d2); free(i1); free(i2);
}

/* approxmatchkernel (approxmatch.cu:3-182), one batch element at a time (blockIdx.x strides b;
 * the per-block scratch lives in temp + blk*(n+m)*2 — we take blk = i % 32 like <<<32,512>>>). */
void oracle_approxmatch(int b, int n, int m, const float* xyz1, const float* xyz2, float* match,
                        float* temp) {
  float multiL, multiR;
  if (n >= m) { multiL = 1; multiR = (float)(n / m); }
  else        { multiL = (float)(m / n); multiR = 1; }
  for (int i = 0; i < b; ++i) {
    const int blk = i % 32;
    float* remainL = temp + (size_t)blk * (n + m) * 2;
    float* remainR = remainL + n;
    float* ratioL = remainL + n + m;
    float* ratioR = remainL + n + m + n;
    float* mt = match + (size_t)i * n * m;
    const float* p1 = xyz1 + (size_t)i * n * 3;
    const float* p2 = xyz2 + (size_t)i * m * 3;
    for (size_t j = 0; j < (size_t)n * m; ++j) mt[j] = 0;
    for (int j = 0; j < n; ++j) remainL[j] = multiL;
    for (int j = 0; j < m; ++j) remainR[j] = multiR;
    for (int j = 7; j > -2; --j) {
      const float level = -powf(4.0f, (float)j);
      for (int k = 0; k < n; ++k) {
        const float x1 = p1[k * 3], y1 = p1[k * 3 + 1], z1 = p1[k * 3 + 2];
        float suml = 1e-9f;
        for (int l = 0; l < m; ++l) {
          const float x2 = p2[l * 3], y2 = p2[l * 3 + 1], z2 = p2[l * 3 + 2];
          const float d = level * ((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1));
          suml += expf(d) * remainR[l];
        }
        ratioL[k] = remainL[k] / suml;
      }
      for (int l = 0; l < m; ++l) {
        const float x2 = p2[l * 3], y2 = p2[l * 3 + 1], z2 = p2[l * 3 + 2];
        float sumr = 0;
        for (int k = 0; k < n; ++k) {
          const float x1 = p1[k * 3], y1 = p1[k * 3 + 1], z1 = p1[k * 3 + 2];
          sumr += expf(level * ((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1))) * ratioL[k];
        }
        sumr *= remainR[l];
        const float consumption = fminf(remainR[l] / (sumr + 1e-9f), 1.0f);
        ratioR[l] = consumption * remainR[l];
        remainR[l] = fmaxf(0.0f, remainR[l] - sumr);
      }
      for (int k = 0; k < n; ++k) {
        const float x1 = p1[k * 3], y1 = p1[k * 3 + 1], z1 = p1[k * 3 + 2];
        const float rl = ratioL[k];
        float suml = 0;
        for (int l = 0; l < m; ++l) {
          const float x2 = p2[l * 3], y2 = p2[l * 3 + 1], z2 = p2[l * 3 + 2];
          const float w = expf(level * ((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1))) * rl * ratioR[l];
          mt[(size_t)l * n + k] += w;
          suml += w;
        }
        remainL[k] = fmaxf(0.0f, remainL[k] - suml);
      }
    }
  }
}

/* matchcostkernel (approxmatch.cu:184-224): sum_l sum_k match[l,k] * |x1_k - x2_l|. */
void oracle_matchcost(int b, int n, int m, const float* xyz1, const float* xyz2, const float* match,
                      float* out) {
  for (int i = 0; i < b; ++i) {
    double s = 0;
    const float* p1 = xyz1 + (size_t)i * n * 3;
    const float* p2 = xyz2 + (size_t)i * m * 3;
    const float* mt = match + (size_t)i * n * m;
    for (int l = 0; l < m; ++l)
      for (int k = 0; k < n; ++k) {
        const float x2 = p2[l * 3] - p1[k * 3], y2 = p2[l * 3 + 1] - p1[k * 3 + 1], z2 = p2[l * 3 + 2] - p1[k * 3 + 2];
        s += (double)(mt[(size_t)l * n + k] * sqrtf(x2 * x2 + y2 * y2 + z2 * z2));
      }
    out[i] = (float)s;
  }
}

/* matchcostgrad1kernel / matchcostgrad2kernel (approxmatch.cu:229-291). */
void oracle_matchcost_grad(int b, int n, int m, const float* xyz1, const float* xyz2, const float* match,
                           float* grad1, float* grad2) {
  for (int i = 0; i < b; ++i) {
    const float* p1 = xyz1 + (size_t)i * n * 3;
    const float* p2 = xyz2 + (size_t)i * m * 3;
    const float* mt = match + (size_t)i * n * m;
    for (int l = 0; l < n; ++l) {
      double dx = 0, dy = 0, dz = 0;
      const float x1 = p1[l * 3], y1 = p1[l * 3 + 1], z1 = p1[l * 3 + 2];
      for (int k = 0; k < m; ++k) {
        const float x2 = p2[k * 3], y2 = p2[k * 3 + 1], z2 = p2[k * 3 + 2];
        const float d = mt[(size_t)k * n + l] / sqrtf(fmaxf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2), 1e-20f));
        dx += (x1 - x2) * d; dy += (y1 - y2) * d; dz += (z1 - z2) * d;
      }
      grad1[((size_t)i * n + l) * 3 + 0] = (float)dx;
      grad1[((size_t)i * n + l) * 3 + 1] = (float)dy;
      grad1[((size_t)i * n + l) * 3 + 2] = (float)dz;
    }
    for (int k = 0; k < m; ++k) {
      double sx = 0, sy = 0, sz = 0;
      const float x2 = p2[k * 3], y2 = p2[k * 3 + 1], z2 = p2[k * 3 + 2];
      for (int j = 0; j < n; ++j) {
        const float x1 = x2 - p1[j * 3], y1 = y2 - p1[j * 3 + 1], z1 = z2 - p1[j * 3 + 2];
        const float d = mt[(size_t)k * n + j] / sqrtf(fmaxf(x1 * x1 + y1 * y1 + z1 * z1, 1e-20f));
        sx += x1 * d; sy += y1 * d; sz += z1 * d;
      }
      grad2[((size_t)i * m + k) * 3 + 0] = (float)sx;
      grad2[((size_t)i * m + k) * 3 + 1] = (float)sy;
      grad2[((size_t)i * m + k) * 3 + 2] = (float)sz;
    }
  }
}
