// voxel_surface.cu — ground-truth surface-voxel extraction on the GPU.
//
// Reference: models/genre_full_model.py:86-96 (Model.preprocess, run per sample on the CPU inside DataLoader workers):
//     val   = flip(transpose(voxel[0], (0, 2, 1)), 2)
//     surf  = clip(val - binary_erosion(val, structure=ones((3,3,3)), iterations=2).astype(float), 0, 1)
// scipy's binary_erosion works on (val != 0) with border_value = 0; two passes with the 3^3 cube are one erosion with the 5^3 cube
// (Minkowski sum), and voxels within `iterations` of the border are always eroded.  The cube is symmetric under the axis swap and
// the flip, so the erosion is done in SOURCE coordinates on a bit-packed copy of the volume (1 bit per voxel, 256 KB per 128^3
// shape) and the transpose + flip happens in the output pass through a shared-memory tile:
//   pack   : a warp per z row: four coalesced 128-byte loads -> four ballots = the row's bits -> z erosion by shifts/ANDs ->
//            R/32 words per row of the workspace;
//   surface: a 32 x 32 (y', z') tile of one x slab: each row ANDs the z-eroded words of its (2r+1)^2 (x, y') neighbours (zeros
//            outside the volume), surf = clip(val - eroded, 0, 1), written transposed + flipped, coalesced on both sides.
// HBM traffic: read the volume twice, write it once (24 MiB per 128^3 shape).
#include "common.cuh"

namespace gb {

constexpr int VS_MAX_WORDS = 8;  // R <= 256

// bits of one z row (R = 32 * W words), every lane returns all words
template <int W>
__device__ __forceinline__ void vs_row_bits(const float *__restrict__ row, unsigned (&w)[W]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < W; ++j) w[j] = __ballot_sync(0xffffffffu, __ldg(row + 32 * j + lane) != 0.0f);
}

// 1-D erosion of a W-word bit row by radius r (zeros beyond both ends)
template <int W>
__device__ __forceinline__ void vs_erode_z(unsigned (&w)[W], int r) {
  unsigned e[W];
#pragma unroll
  for (int j = 0; j < W; ++j) e[j] = w[j];
  for (int s = 1; s <= r; ++s) {
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const unsigned up = (w[j] >> s) | (j + 1 < W ? w[j + 1] << (32 - s) : 0u);    // bit z <- bit z + s
      const unsigned dn = (w[j] << s) | (j > 0 ? w[j - 1] >> (32 - s) : 0u);        // bit z <- bit z - s
      e[j] &= up & dn;
    }
  }
#pragma unroll
  for (int j = 0; j < W; ++j) w[j] = e[j];
}

template <int W>
__global__ void __launch_bounds__(256)
vs_pack_kernel(const float *__restrict__ vox, unsigned *__restrict__ zer, int R, int iters, long long rows_total) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);   // n * R * R + x * R + y'
  if (row >= rows_total) return;
  unsigned w[W];
  vs_row_bits<W>(vox + (size_t)row * R, w);
  vs_erode_z<W>(w, iters);
  const int lane = threadIdx.x & 31;
  if (lane < W) {
    unsigned v = 0;
#pragma unroll
    for (int j = 0; j < W; ++j)
      if (lane == j) v = w[j];
    zer[(size_t)row * W + lane] = v;
  }
}

// blockIdx.x = (y' tile, z' tile), blockIdx.y = x, blockIdx.z = n; 32 x 8 threads, 4 rows per thread
template <int W, bool XFORM>
__global__ void __launch_bounds__(256)
vs_surface_kernel(const float *__restrict__ vox, const unsigned *__restrict__ zer, float *__restrict__ out, int R, int iters) {
  __shared__ float tile[32][33];
  __shared__ unsigned s_er[32];      // eroded bits of the tile's 32 rows (the word of this z' tile)
  const int tiles = R / 32;
  const int ty = blockIdx.x / tiles, tz = blockIdx.x % tiles;
  const int x = blockIdx.y, n = blockIdx.z;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;   // 8 warps
  const size_t vol = (size_t)n * R * R * R;
  const unsigned *zv = zer + (size_t)n * R * R * W;
  if (threadIdx.x < 32) {   // row r = y' of the tile: AND over the (2 iters + 1)^2 neighbours in (x, y')
    const int yp = ty * 32 + threadIdx.x;
    unsigned acc = 0xffffffffu;
    for (int dx = -iters; dx <= iters; ++dx)
      for (int dy = -iters; dy <= iters; ++dy) {
        const int xx = x + dx, yy = yp + dy;
        acc &= ((unsigned)xx < (unsigned)R && (unsigned)yy < (unsigned)R) ? zv[((size_t)xx * R + yy) * W + tz] : 0u;
      }
    s_er[threadIdx.x] = acc;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = ly * 4 + k;                 // y' within the tile
    const int yp = ty * 32 + r, zp = tz * 32 + lx;
    const float v = __ldg(vox + vol + ((size_t)x * R + yp) * R + zp);
    const float er = (s_er[r] >> lx) & 1u ? 1.0f : 0.0f;
    const float s = fminf(fmaxf(v - er, 0.0f), 1.0f);
    if (XFORM) tile[r][lx] = s;
    else out[vol + ((size_t)x * R + yp) * R + zp] = s;
  }
  if (XFORM) {
    __syncthreads();
    // out[x][y][z] = surf_src[x][y' = R-1-z][z' = y]: the tile's z' range becomes y, its reversed y' range becomes z
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yo = tz * 32 + ly * 4 + k;              // y = z'
      const int zo = (R - 1 - ty * 32) - 31 + lx;       // z = R-1-y', ascending with the lane
      out[vol + ((size_t)x * R + yo) * R + zo] = tile[31 - lx][ly * 4 + k];
    }
  }
}

}  // namespace gb

using namespace gb;

extern "C" size_t genre_b200_voxel_surface_workspace_bytes(int64_t N, int res) {
  if (N <= 0 || res <= 0 || res % 32) return 0;
  return (size_t)N * res * res * (res / 32) * sizeof(unsigned);
}

// out = clip(val - binary_erosion(val != 0, ones(3,3,3), iterations), 0, 1), val = the volume itself (transpose_flip = 0) or
// flip(transpose(vol, (0,2,1)), 2) (transpose_flip = 1: the layout change of Model.preprocess, genre_full_model.py:89-90).
//   vox, out [N, R, R, R] dense fp32 (must not alias), R a multiple of 32 and <= 256, 1 <= iterations <= 15
//   workspace: genre_b200_voxel_surface_workspace_bytes(N, R) bytes
extern "C" int genre_b200_voxel_surface(const float *vox, int64_t N, int res, int iterations, int transpose_flip, float *out,
                                        void *workspace, size_t workspace_bytes, void *stream) {
  GB_REQUIRE(vox && out && workspace, GENRE_B200_EINVAL, "voxel_surface: null pointer");
  GB_REQUIRE(vox != out, GENRE_B200_EINVAL, "voxel_surface: in-place operation is not supported");
  GB_REQUIRE(N > 0 && N <= 65535 && res >= 32 && res % 32 == 0 && res <= 32 * VS_MAX_WORDS, GENRE_B200_EINVAL,
             "voxel_surface: N=%lld, R=%d unsupported (R a multiple of 32, at most %d)", (long long)N, res, 32 * VS_MAX_WORDS);
  GB_REQUIRE(iterations >= 1 && iterations <= 15, GENRE_B200_EINVAL, "voxel_surface: iterations %d out of range", iterations);
  GB_REQUIRE(workspace_bytes >= genre_b200_voxel_surface_workspace_bytes(N, res) && ((uintptr_t)workspace & 3) == 0,
             GENRE_B200_EWORKSPACE, "voxel_surface: workspace too small or misaligned");
  cudaStream_t st = as_stream(stream);
  const long long rows = (long long)N * res * res;
  const unsigned pgrid = (unsigned)((rows + 7) / 8);
  const int W = res / 32, tiles = W;
  dim3 sgrid((unsigned)(tiles * tiles), (unsigned)res, (unsigned)N);
  unsigned *zer = (unsigned *)workspace;
#define GB_VS(WW)                                                                                               \
  case WW:                                                                                                      \
    vs_pack_kernel<WW><<<pgrid, 256, 0, st>>>(vox, zer, res, iterations, rows);                                  \
    if (int rc = check_launch("voxel_surface pack kernel")) return rc;                                           \
    if (transpose_flip) vs_surface_kernel<WW, true><<<sgrid, 256, 0, st>>>(vox, zer, out, res, iterations);      \
    else vs_surface_kernel<WW, false><<<sgrid, 256, 0, st>>>(vox, zer, out, res, iterations);                    \
    break;
  switch (W) {
    GB_VS(1) GB_VS(2) GB_VS(3) GB_VS(4) GB_VS(5) GB_VS(6) GB_VS(7) GB_VS(8)
    default: return fail_arg(GENRE_B200_EINVAL, "voxel_surface: R=%d unsupported", res);
  }
#undef GB_VS
  return check_launch("voxel_surface kernel");
}
