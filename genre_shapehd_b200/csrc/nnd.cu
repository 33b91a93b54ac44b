// nnd.cu — Chamfer nearest-neighbour distance (both directions) and its backward.
//
// Reference: toolbox/nndistance/src/nnd_cuda.cu
//   NmDistanceKernel :6-128  fixed <<<(32,16),512>>> grid regardless of problem size, one thread per
//                            query scanning all candidates, legacy default stream (:130-131)
//   NmDistanceGradKernel :143-162, launcher :163-177 (two cudaMemset + two launches, default stream)
// Here one launch covers both directions.  A CTA owns 64 queries; each query is scanned by 4 lanes
// that take interleaved candidates from a shared-memory tile (16-byte broadcast loads), and the 4
// partial (distance, index) pairs meet in a 2-step shuffle reduction that keeps the lowest index on
// ties — the reference's semantics (strict '<' on an ascending scan, nnd_cuda.cu:33,120).
// The squared distance is rounded exactly as nvcc contracts the reference expression
// x2*x2+y2*y2+z2*z2 (SASS of oracle/_ref/libref_nnd_cuda.so):  fma(dz,dz, fma(dx,dx, dy*dy)),
// with d* = candidate - query, so distances AND indices are bit-identical to the reference kernel.
#include "common.cuh"
#include <math_constants.h>

namespace gb {

constexpr int NND_THREADS = 256;
constexpr int NND_SLICES = 4;                        // lanes per query
constexpr int NND_QUERIES = NND_THREADS / NND_SLICES;  // 64 queries per CTA
constexpr int NND_TILE = 1024;                       // candidates per shared-memory tile (16 KiB)

__global__ void __launch_bounds__(NND_THREADS)
nnd_forward_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2, int N, int M,
                   float *__restrict__ dist1, float *__restrict__ dist2, int *__restrict__ idx1,
                   int *__restrict__ idx2) {
  __shared__ float4 s_c[NND_TILE];
  const int b = blockIdx.y;
  const bool fwd = blockIdx.z == 0;  // z = 0: queries from xyz1 against xyz2;  z = 1: the other way round
  const int nq = fwd ? N : M, nc = fwd ? M : N;
  if (blockIdx.x * NND_QUERIES >= nq) return;
  const float *q = (fwd ? xyz1 : xyz2) + (size_t)b * nq * 3;
  const float *cand = (fwd ? xyz2 : xyz1) + (size_t)b * nc * 3;
  float *dist = (fwd ? dist1 : dist2) + (size_t)b * nq;
  int *idx = (fwd ? idx1 : idx2) + (size_t)b * nq;

  const int slice = threadIdx.x & (NND_SLICES - 1);
  const int qi = blockIdx.x * NND_QUERIES + (threadIdx.x >> 2);
  const bool live = qi < nq;
  const float qx = live ? q[qi * 3 + 0] : 0.f, qy = live ? q[qi * 3 + 1] : 0.f, qz = live ? q[qi * 3 + 2] : 0.f;

  float best = CUDART_INF_F;
  int besti = 0;
  for (int k0 = 0; k0 < nc; k0 += NND_TILE) {
    const int tn = min(NND_TILE, nc - k0);
    __syncthreads();  // previous tile fully consumed
    for (int j = threadIdx.x; j < tn; j += NND_THREADS) {
      const float *c = cand + (size_t)(k0 + j) * 3;
      s_c[j] = make_float4(c[0], c[1], c[2], 0.f);
    }
    __syncthreads();
#pragma unroll 4
    for (int j = slice; j < tn; j += NND_SLICES) {
      const float4 c = s_c[j];
      const float dx = __fadd_rn(c.x, -qx), dy = __fadd_rn(c.y, -qy), dz = __fadd_rn(c.z, -qz);
      const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
      if (d < best) {
        best = d;
        besti = k0 + j;
      }
    }
  }
  // combine the 4 slices of a query: smaller distance wins, equal distance -> smaller index
#pragma unroll
  for (int o = 1; o < NND_SLICES; o <<= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (od < best || (od == best && oi < besti)) {
      best = od;
      besti = oi;
    }
  }
  if (live && slice == 0) {
    dist[qi] = best;
    idx[qi] = besti;
  }
}

// grad_xyz1[j] += 2 g1[j] (x1[j] - x2[idx1[j]]);  grad_xyz2[idx1[j]] -= the same; and symmetrically for g2
__global__ void __launch_bounds__(256)
nnd_backward_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2, int N, int M,
                    const float *__restrict__ g1, const float *__restrict__ g2, const int *__restrict__ idx1,
                    const int *__restrict__ idx2, float *__restrict__ grad1, float *__restrict__ grad2) {
  const int b = blockIdx.y;
  const bool fwd = blockIdx.z == 0;
  const int nq = fwd ? N : M, nc = fwd ? M : N;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nq) return;
  const float *a = (fwd ? xyz1 : xyz2) + (size_t)b * nq * 3;
  const float *o = (fwd ? xyz2 : xyz1) + (size_t)b * nc * 3;
  float *ga = (fwd ? grad1 : grad2) + (size_t)b * nq * 3;
  float *go = (fwd ? grad2 : grad1) + (size_t)b * nc * 3;
  const int j2 = (fwd ? idx1 : idx2)[(size_t)b * nq + j];
  const float g = (fwd ? g1 : g2)[(size_t)b * nq + j] * 2.0f;
  const float vx = g * (a[j * 3 + 0] - o[j2 * 3 + 0]);
  const float vy = g * (a[j * 3 + 1] - o[j2 * 3 + 1]);
  const float vz = g * (a[j * 3 + 2] - o[j2 * 3 + 2]);
  atomicAdd(ga + j * 3 + 0, vx);
  atomicAdd(ga + j * 3 + 1, vy);
  atomicAdd(ga + j * 3 + 2, vz);
  atomicAdd(go + j2 * 3 + 0, -vx);
  atomicAdd(go + j2 * 3 + 1, -vy);
  atomicAdd(go + j2 * 3 + 2, -vz);
}

static int nnd_check(int64_t B, int64_t N, int64_t M) {
  GB_REQUIRE(B > 0 && B <= 65535, GENRE_B200_EINVAL, "nnd: batch %lld must be in [1, 65535]", (long long)B);
  GB_REQUIRE(N > 0 && M > 0 && N < (1ll << 29) && M < (1ll << 29), GENRE_B200_EINVAL,
             "nnd: point counts (%lld, %lld) must be in [1, 2^29)", (long long)N, (long long)M);
  return 0;
}

}  // namespace gb

using namespace gb;

extern "C" int genre_b200_nnd_forward(const float *xyz1, const float *xyz2, int64_t B, int64_t N, int64_t M,
                                      float *dist1, float *dist2, int32_t *idx1, int32_t *idx2, void *stream) {
  GB_REQUIRE(xyz1 && xyz2 && dist1 && dist2 && idx1 && idx2, GENRE_B200_EINVAL, "nnd forward: null pointer");
  if (int rc = nnd_check(B, N, M)) return rc;
  const int64_t mx = N > M ? N : M;
  dim3 grid((unsigned)((mx + NND_QUERIES - 1) / NND_QUERIES), (unsigned)B, 2);
  nnd_forward_kernel<<<grid, NND_THREADS, 0, as_stream(stream)>>>(xyz1, xyz2, (int)N, (int)M, dist1, dist2, idx1, idx2);
  return check_launch("nnd forward kernel");
}

extern "C" int genre_b200_nnd_backward(const float *xyz1, const float *xyz2, int64_t B, int64_t N, int64_t M,
                                       const float *grad_dist1, const float *grad_dist2, const int32_t *idx1,
                                       const int32_t *idx2, float *grad_xyz1, float *grad_xyz2, void *stream) {
  GB_REQUIRE(xyz1 && xyz2 && grad_dist1 && grad_dist2 && idx1 && idx2 && grad_xyz1 && grad_xyz2, GENRE_B200_EINVAL,
             "nnd backward: null pointer");
  if (int rc = nnd_check(B, N, M)) return rc;
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(grad_xyz1, 0, (size_t)B * N * 3 * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(grad_xyz2, 0, (size_t)B * M * 3 * sizeof(float), st);
  if (e != cudaSuccess) {
    set_error("nnd backward: memset: %s", cudaGetErrorString(e));
    return (int)e;
  }
  const int64_t mx = N > M ? N : M;
  dim3 grid((unsigned)((mx + 255) / 256), (unsigned)B, 2);
  nnd_backward_kernel<<<grid, 256, 0, st>>>(xyz1, xyz2, (int)N, (int)M, grad_dist1, grad_dist2, idx1, idx2, grad_xyz1,
                                            grad_xyz2);
  return check_launch("nnd backward kernel");
}
