// skinny_gemm.cu — the two degenerate 3D convolutions at the bottom of the U-Net as weight-streaming FP32 products.
//
// Reference layers: networks/networks.py:157 (Unet_3D.enc6 = Conv3d(16 nf -> 32 nf, k 4, s 1, p 0) on a 4^3 input -> 1^3:
//     out[b, co] = sum_{ci,k} x[b, ci, k] W[co, ci, k]            =  x[B, Cin*64] @ W[Cout, Cin*64]^T          ("NT")
// and :162 (Unet_3D.dec1 = ConvTranspose3d(64 nf -> 16 nf, k 4, s 1, p 0) on a 1^3 input -> 4^3, also the first layer of
// VoxelDecoder / VoxelGenerator :40,:79):
//     out[b, co, k] = sum_ci x[b, ci] W[ci, co, k]                 =  x[B, Cin] @ W[Cin, Cout*64]               ("NN")
// With B = 16 rows these are bound by reading the weights once (52 MB / 105 MB at nf = 20); there is nothing for the tensor
// cores to win, and FP32 FMAs keep the reference's fp32 semantics exactly.  Each CTA streams a (rows x K-chunk) or
// (K-chunk x columns) block of W with 16-byte coalesced loads against the activation rows staged in shared memory, and
// writes a partial product; a second pass sums the K-chunks in a fixed order (deterministic) and applies
// y = lrelu(scale[c] * acc + shift[c]) (bias, folded eval BatchNorm, activation).
#include "common.cuh"

namespace gb {

constexpr int SK_MB = 16;          // activation rows per block
constexpr int SK_NT_ROWS = 8;      // W rows per CTA (NT)
constexpr int SK_NT_KCHUNK = 1024; // = 256 threads x float4
constexpr int SK_NN_COLS = 512;    // = 128 threads x float4
constexpr int SK_NN_KCHUNK = 128;

// sum 32 per-lane values across the warp: afterwards r[0] of lane L is the total of value L (31 shuffles instead of 160)
__device__ __forceinline__ void warp_transpose_reduce32(float (&r)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = upper ? r[i + s] : r[i];
      const float send = upper ? r[i] : r[i + s];
      r[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}

// partial[kc][m][n] = sum_{k in chunk kc} x[m0 + m][k] * W[n][k];  grid (ceil(N/8), ceil(K/1024), ceil(M/16))
__global__ void __launch_bounds__(256)
skinny_nt_kernel(const float *__restrict__ x, const float *__restrict__ W, int M, int N, int K, float *__restrict__ partial) {
  extern __shared__ __align__(16) float xs[];   // [16][1024]
  __shared__ float red[8][SK_NT_ROWS * SK_MB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * SK_NT_ROWS, k0 = blockIdx.y * SK_NT_KCHUNK, m0 = blockIdx.z * SK_MB;
  for (int i = tid; i < SK_MB * (SK_NT_KCHUNK / 4); i += 256) {
    const int r = i / (SK_NT_KCHUNK / 4), c = (i % (SK_NT_KCHUNK / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + r < M && k0 + c < K) v = __ldg(reinterpret_cast<const float4 *>(x + (size_t)(m0 + r) * K + k0 + c));
    *reinterpret_cast<float4 *>(xs + r * SK_NT_KCHUNK + c) = v;
  }
  const int k = k0 + tid * 4;
  float4 w[SK_NT_ROWS];
#pragma unroll
  for (int r = 0; r < SK_NT_ROWS; ++r)
    w[r] = (n0 + r < N && k < K) ? __ldg(reinterpret_cast<const float4 *>(W + (size_t)(n0 + r) * K + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  float acc[SK_NT_ROWS * SK_MB];   // [row][m]
#pragma unroll
  for (int m = 0; m < SK_MB; ++m) {
    const float4 xv = *reinterpret_cast<const float4 *>(xs + m * SK_NT_KCHUNK + tid * 4);
#pragma unroll
    for (int r = 0; r < SK_NT_ROWS; ++r)
      acc[r * SK_MB + m] = fmaf(w[r].x, xv.x, fmaf(w[r].y, xv.y, fmaf(w[r].z, xv.z, w[r].w * xv.w)));
  }
#pragma unroll
  for (int g = 0; g < SK_NT_ROWS * SK_MB / 32; ++g) {
    float r32[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r32[i] = acc[g * 32 + i];
    warp_transpose_reduce32(r32, lane);
    red[warp][g * 32 + lane] = r32[0];
  }
  __syncthreads();
  if (tid < SK_NT_ROWS * SK_MB) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += red[wv][tid];
    const int r = tid / SK_MB, m = tid % SK_MB;
    if (n0 + r < N && m0 + m < M) partial[((size_t)blockIdx.y * M + m0 + m) * N + n0 + r] = s;
  }
}

// partial[kc][m][n] = sum_{k in chunk kc} x[m0 + m][k] * W[k][n];  grid (ceil(N/512), ceil(K/128), ceil(M/16)); N % 4 == 0
__global__ void __launch_bounds__(128)
skinny_nn_kernel(const float *__restrict__ x, const float *__restrict__ W, int M, int N, int K, float *__restrict__ partial) {
  __shared__ __align__(16) float xs[SK_NN_KCHUNK][SK_MB];   // [k][m]
  const int tid = threadIdx.x;
  const int n = blockIdx.x * SK_NN_COLS + tid * 4, k0 = blockIdx.y * SK_NN_KCHUNK, m0 = blockIdx.z * SK_MB;
  for (int i = tid; i < SK_NN_KCHUNK * SK_MB; i += 128) {
    const int m = i / SK_NN_KCHUNK, kk = i % SK_NN_KCHUNK;    // coalesced along k
    xs[kk][m] = (m0 + m < M && k0 + kk < K) ? __ldg(x + (size_t)(m0 + m) * K + k0 + kk) : 0.f;
  }
  __syncthreads();
  float acc[SK_MB][4];
#pragma unroll
  for (int m = 0; m < SK_MB; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
  const int kn = min(SK_NN_KCHUNK, K - k0);
  if (n < N) {
    const float *wp = W + (size_t)k0 * N + n;
#pragma unroll 8
    for (int kk = 0; kk < kn; ++kk) {
      const float4 w = __ldg(reinterpret_cast<const float4 *>(wp + (size_t)kk * N));
#pragma unroll
      for (int m4 = 0; m4 < SK_MB / 4; ++m4) {
        const float4 xv = *reinterpret_cast<const float4 *>(&xs[kk][m4 * 4]);
        const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[m4 * 4 + j][0] = fmaf(xa[j], w.x, acc[m4 * 4 + j][0]);
          acc[m4 * 4 + j][1] = fmaf(xa[j], w.y, acc[m4 * 4 + j][1]);
          acc[m4 * 4 + j][2] = fmaf(xa[j], w.z, acc[m4 * 4 + j][2]);
          acc[m4 * 4 + j][3] = fmaf(xa[j], w.w, acc[m4 * 4 + j][3]);
        }
      }
    }
#pragma unroll
    for (int m = 0; m < SK_MB; ++m)
      if (m0 + m < M)
        *reinterpret_cast<float4 *>(partial + ((size_t)blockIdx.y * M + m0 + m) * N + n) = make_float4(acc[m][0], acc[m][1], acc[m][2], acc[m][3]);
  }
}

// out[m][n] = lrelu(scale[n / div] * sum_kc partial[kc][m][n] + shift[n / div])
__global__ void __launch_bounds__(256)
skinny_finish_kernel(const float *__restrict__ partial, int chunks, long long MN, int N, int div, const float *__restrict__ scale,
                     const float *__restrict__ shift, float slope, float *__restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= MN) return;
  float s = 0.f;
  for (int c = 0; c < chunks; ++c) s += __ldg(partial + (size_t)c * MN + i);
  const int ch = (int)(i % N) / div;
  const float t = fmaf(s, __ldg(scale + ch), __ldg(shift + ch));
  out[i] = t > 0.f ? t : t * slope;
}

}  // namespace gb

using namespace gb;

extern "C" size_t genre_b200_skinny_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int w_is_nk) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int64_t chunk = w_is_nk ? SK_NT_KCHUNK : SK_NN_KCHUNK;
  return (size_t)((K + chunk - 1) / chunk) * M * N * sizeof(float);
}

// out[M][N] = lrelu(scale[n / chan_div] * (x[M][K] @ Wm) + shift[n / chan_div]), all fp32, FP32 FMAs.
//   w_is_nk = 1: W is [N][K] (Conv3d whose kernel covers its whole input: Unet_3D.enc6), K % 4 == 0;
//   w_is_nk = 0: W is [K][N] (ConvTranspose3d on a 1^3 input: Unet_3D.dec1, VoxelDecoder / VoxelGenerator main.0), N % 4 == 0;
//   scale, shift: [ceil(N / chan_div)] (chan_div = k^3 for the transposed case, 1 otherwise); workspace: see above.
extern "C" int genre_b200_skinny_gemm(const float *x, const float *W, int64_t M, int64_t N, int64_t K, int w_is_nk, int chan_div,
                                      const float *scale, const float *shift, float slope, float *out, void *workspace,
                                      size_t workspace_bytes, void *stream) {
  GB_REQUIRE(x && W && scale && shift && out && workspace, GENRE_B200_EINVAL, "skinny_gemm: null pointer");
  GB_REQUIRE(M > 0 && N > 0 && K > 0 && chan_div > 0 && M <= 65535 * SK_MB && N < (1ll << 30) && K < (1ll << 30), GENRE_B200_EINVAL,
             "skinny_gemm: M=%lld N=%lld K=%lld out of range", (long long)M, (long long)N, (long long)K);
  GB_REQUIRE(aligned16(x) && aligned16(W) && aligned16(workspace), GENRE_B200_EALIGN, "skinny_gemm: operands must be 16-byte aligned");
  GB_REQUIRE(w_is_nk ? K % 4 == 0 : N % 4 == 0, GENRE_B200_EINVAL, "skinny_gemm: the contiguous extent of W must be a multiple of 4");
  GB_REQUIRE(workspace_bytes >= genre_b200_skinny_gemm_workspace_bytes(M, N, K, w_is_nk), GENRE_B200_EWORKSPACE,
             "skinny_gemm: workspace too small");
  cudaStream_t st = as_stream(stream);
  float *partial = (float *)workspace;
  const unsigned mblocks = (unsigned)((M + SK_MB - 1) / SK_MB);
  int chunks;
  if (w_is_nk) {
    chunks = (int)((K + SK_NT_KCHUNK - 1) / SK_NT_KCHUNK);
    GB_REQUIRE(chunks <= 65535, GENRE_B200_EINVAL, "skinny_gemm: K too large");
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem = (size_t)SK_MB * SK_NT_KCHUNK * sizeof(float);
    if (!configured[dev & 63]) {
      if (cudaFuncSetAttribute(skinny_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return check_launch("skinny_gemm: cudaFuncSetAttribute");
      configured[dev & 63] = true;
    }
    dim3 grid((unsigned)((N + SK_NT_ROWS - 1) / SK_NT_ROWS), (unsigned)chunks, mblocks);
    skinny_nt_kernel<<<grid, 256, smem, st>>>(x, W, (int)M, (int)N, (int)K, partial);
  } else {
    chunks = (int)((K + SK_NN_KCHUNK - 1) / SK_NN_KCHUNK);
    GB_REQUIRE(chunks <= 65535, GENRE_B200_EINVAL, "skinny_gemm: K too large");
    dim3 grid((unsigned)((N + SK_NN_COLS - 1) / SK_NN_COLS), (unsigned)chunks, mblocks);
    skinny_nn_kernel<<<grid, 128, 0, st>>>(x, W, (int)M, (int)N, (int)K, partial);
  }
  if (int rc = check_launch("skinny_gemm product kernel")) return rc;
  const long long MN = (long long)M * N;
  skinny_finish_kernel<<<(unsigned)((MN + 255) / 256), 256, 0, st>>>(partial, chunks, MN, (int)N, chan_div, scale, shift, slope, out);
  return check_launch("skinny_gemm finish kernel");
}
