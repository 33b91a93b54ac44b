// sph_bp.cu — spherical back-projection (spherical depth map -> voxel TDF + count) and its backward.
//
// Reference: toolbox/cam_bp/cam_bp/src/back_projection_kernel.cu
//   forward  kernel :474-542 + inplace_safe_divide(bias 0) :281-306 + wrap :629-703
//   backward kernel :544-627 + wrap :704-757
// Forward shares the bin + splat stages with the camera back-projection (voxelize.cuh); only the
// per-pixel projection differs:  p = grid[h, w, :] * r.
#include "voxelize.cuh"

namespace gb {

constexpr int SPH_THREADS = 256;

struct SphPoint {
  float gx, gy, gz;
  int ix, iy, iz;
  bool in_bounds;
};

__device__ __forceinline__ SphPoint sph_unproject(float r, float dxg, float dyg, float dzg, int R, float Rf) {
  SphPoint p;
  p.gx = __fmul_rn(dxg, r);
  p.gy = __fmul_rn(dyg, r);
  p.gz = __fmul_rn(dzg, r);
  p.ix = floor_i_ref(__fmul_rn(__fadd_rn(p.gx, 0.5f), Rf));
  p.iy = floor_i_ref(__fmul_rn(__fadd_rn(p.gy, 0.5f), Rf));
  p.iz = floor_i_ref(__fmul_rn(__fadd_rn(p.gz, 0.5f), Rf));
  p.in_bounds = (p.ix >= 0) & (p.ix < R) & (p.iy >= 0) & (p.iy < R) & (p.iz >= 0) & (p.iz < R);
  return p;
}

constexpr int SPH_PIX = 4;  // pixels per thread in the project stage

// AFFINE: the radius is in_bias + in_scale * map value (GenRe feeds 1 - cropped map, genre_full_model.py:139)
template <bool W_FAST, bool AFFINE>
__global__ void __launch_bounds__(SPH_THREADS)
sph_project_kernel(const float *__restrict__ sph, int C, int H, int W, long long sN, long long sC, long long sH,
                   long long sW, const float *__restrict__ grid, long long gN, long long gC, long long gH,
                   long long gW, long long gD, int R, float qscale, VoxWorkspace ws, float in_scale, float in_bias) {
  extern __shared__ unsigned s_hist[];  // [ntiles] CTA-local tile histogram
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // let the splat grid start launching behind us
  const int map = blockIdx.y;
  const int n = map / C, c = map - n * C;
  const int P = H * W;
  const float *smap = sph + n * sN + c * sC;
  const float *gmap = grid + n * gN + c * gC;
  const VoxGrid vg = make_grid(R);
  const int fast = W_FAST ? W : H;
  const int p0 = blockIdx.x * (SPH_THREADS * SPH_PIX) + threadIdx.x;
  float r[SPH_PIX], dxg[SPH_PIX], dyg[SPH_PIX], dzg[SPH_PIX];
#pragma unroll
  for (int k = 0; k < SPH_PIX; ++k) {
    const int p = p0 + k * SPH_THREADS;
    const int slow = p / fast, fst = p - slow * fast;
    const int h = W_FAST ? slow : fst, w = W_FAST ? fst : slow;
    r[k] = -1.0f;
    dxg[k] = dyg[k] = dzg[k] = 0.0f;
    if (p < P) {
      r[k] = smap[h * sH + w * sW];
      if (AFFINE) r[k] = __fadd_rn(in_bias, __fmul_rn(in_scale, r[k]));
      const float *g = gmap + h * gH + w * gW;
      dxg[k] = g[0];
      dyg[k] = g[gD];
      dzg[k] = g[2 * gD];
    }
  }
  unsigned gv[SPH_PIX], q[SPH_PIX];
#pragma unroll
  for (int k = 0; k < SPH_PIX; ++k) {
    gv[k] = VOX_INVALID;
    q[k] = 0;
    if (!(r[k] < 0.0f)) {
      const SphPoint pt = sph_unproject(r[k], dxg[k], dyg[k], dzg[k], R, vg.Rf);
      if (pt.in_bounds) {
        const float dx = __fadd_rn(pt.gx, -vox_centre(pt.ix, vg));
        const float dy = __fadd_rn(pt.gy, -vox_centre(pt.iy, vg));
        const float dz = __fadd_rn(pt.gz, -vox_centre(pt.iz, vg));
        const float dist = __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))));
        gv[k] = (unsigned)((pt.ix * R + pt.iy) * R + pt.iz);
        q[k] = vox_quantise(dist, qscale);
      }
    }
  }
  vox_emit<SPH_PIX, SPH_THREADS>(gv, q, map, ws, P, s_hist);
}

__global__ void __launch_bounds__(SPH_THREADS)
sph_bp_backward_kernel(const float *__restrict__ sph, int C, int H, int W, long long sN, long long sC, long long sH,
                       long long sW, const float *__restrict__ grid, long long gN, long long gC, long long gH,
                       long long gW, long long gD, const float *__restrict__ cnt, const float *__restrict__ grad_tdf,
                       int R, float *__restrict__ grad_sph) {
  const int map = blockIdx.y;
  const int n = map / C, c = map - n * C;
  const int P = H * W;
  const int p = blockIdx.x * SPH_THREADS + threadIdx.x;
  if (p >= P) return;
  const int h = p / W, w = p - h * W;
  float out = 0.0f;
  const float r = sph[n * sN + c * sC + h * sH + w * sW];
  if (!(r < 0.0f)) {
    const float *g = grid + n * gN + c * gC + h * gH + w * gW;
    const float Rf = (float)R;
    const SphPoint pt = sph_unproject(r, g[0], g[gD], g[2 * gD], R, Rf);
    if (pt.in_bounds) {
      const float cx = (float)((((double)(float)pt.ix + 0.5) / (double)Rf) - 0.5);
      const float cy = (float)((((double)(float)pt.iy + 0.5) / (double)Rf) - 0.5);
      const float cz = (float)((((double)(float)pt.iz + 0.5) / (double)Rf) - 0.5);
      float len = sqrtf(pt.gx * pt.gx + pt.gy * pt.gy + pt.gz * pt.gz);
      if ((double)len < 1e-5) len = 1e-5f;
      const float ux = pt.gx / len, uy = pt.gy / len, uz = pt.gz / len;
      const float cos_cc = ux * cx + uy * cy + uz * cz;
      const float ex = pt.gx - cx, ey = pt.gy - cy, ez = pt.gz - cz;
      float dist = sqrtf(ex * ex + ey * ey + ez * ez);
      const size_t v = (size_t)map * R * R * R + ((size_t)pt.ix * R + pt.iy) * R + pt.iz;
      float ptnum = cnt[v];
      if (ptnum < 1.0f) ptnum = 1.0f;
      if ((double)dist < 1e-5) dist = 1e-5f;
      const float gd = grad_tdf[v];
      out = gd * (r - cos_cc) / (ptnum * dist);
    }
  }
  grad_sph[(size_t)map * P + p] = out;
}

static int sph_check(const float *sph, int64_t N, int64_t C, int64_t H, int64_t W, const float *grid, int res) {
  GB_REQUIRE(sph && grid, GENRE_B200_EINVAL, "sph_bp: null input pointer");
  GB_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, GENRE_B200_EINVAL, "sph_bp: empty spherical map [%lld,%lld,%lld,%lld]",
             (long long)N, (long long)C, (long long)H, (long long)W);
  GB_REQUIRE(H < (1 << 20) && W < (1 << 20), GENRE_B200_EINVAL, "sph_bp: map too large");
  return vox_check_common(N * C, H * W, res);
}

}  // namespace gb

using namespace gb;

static int sph_forward_impl(const float *sph, int64_t N, int64_t C, int64_t H, int64_t W, int64_t sN, int64_t sC,
                            int64_t sH, int64_t sW, const float *grid, int64_t gN, int64_t gC, int64_t gH, int64_t gW,
                            int64_t gD, float *tdf, float *cnt, int res, void *workspace, size_t workspace_bytes,
                            void *stream, bool affine, float in_scale, float in_bias, float alpha, float beta,
                            int64_t out_stride) {
  if (int rc = sph_check(sph, N, C, H, W, grid, res)) return rc;
  GB_REQUIRE(tdf != nullptr, GENRE_B200_EINVAL, "sph_bp: tdf is null");
  VoxWorkspace w;
  GB_REQUIRE(vox_carve(workspace, workspace_bytes, N * C, H * W, res, &w), GENRE_B200_EWORKSPACE,
             "sph_bp: workspace too small or misaligned (need %zu bytes)", vox_workspace_bytes(N * C, H * W, res));
  cudaStream_t st = as_stream(stream);
  const int64_t P = H * W;
  const float qscale = (float)res * 16777216.0f;
  if (int rc = vox_clear_counts(w, N * C, st)) return rc;
  dim3 grd((unsigned)((P + SPH_THREADS * SPH_PIX - 1) / (SPH_THREADS * SPH_PIX)), (unsigned)(N * C));
  const size_t smem = (size_t)w.ntiles * 4;
  const bool wf = llabs(sW) <= llabs(sH);
#define GB_SPH(WF, AF)                                                                                              \
  sph_project_kernel<WF, AF><<<grd, SPH_THREADS, smem, st>>>(sph, (int)C, (int)H, (int)W, sN, sC, sH, sW, grid, gN, gC, \
                                                            gH, gW, gD, res, qscale, w, in_scale, in_bias)
  if (wf && affine) GB_SPH(true, true);
  else if (wf) GB_SPH(true, false);
  else if (affine) GB_SPH(false, true);
  else GB_SPH(false, false);
#undef GB_SPH
  if (int rc = check_launch("sph_bp project kernel")) return rc;
  return vox_splat(w, N * C, P, res, tdf, cnt, alpha, beta, 0.0f, st, true, out_stride);
}

extern "C" int genre_b200_sph_bp_forward(const float *sph, int64_t N, int64_t C, int64_t H, int64_t W, int64_t sN,
                                         int64_t sC, int64_t sH, int64_t sW, const float *grid, int64_t gN, int64_t gC,
                                         int64_t gH, int64_t gW, int64_t gD, float *tdf, float *cnt, int res,
                                         void *workspace, size_t workspace_bytes, void *stream) {
  // tdf = mean distance on hit voxels, 0 elsewhere (sperical_to_tdf.py:26-27 zero init, kernel bias 0 at :695)
  const float beta = (float)((1.0 / 16777216.0) / (double)res);
  return sph_forward_impl(sph, N, C, H, W, sN, sC, sH, sW, grid, gN, gC, gH, gW, gD, tdf, cnt, res, workspace,
                          workspace_bytes, stream, false, 1.0f, 0.0f, 0.0f, beta, 0);
}

// Fused form of GenRe's spherical back-projection glue (genre_full_model.py:134-143):
//     radius = in_bias + in_scale * sph          (the caller's `1 - crop_sph`; the crop is just strides)
//     out    = (-tdf + 1/R) * R * clamp(cnt,0,1) = 1 - R * mean distance on hit voxels, 0 elsewhere
// written with a per-map stride of out_map_stride floats (>= R^3), i.e. straight into a channel of the refiner's
// [B,2,R,R,R] input.  No count volume, no elementwise passes.  Inference only (no backward through this form).
extern "C" int genre_b200_sph_bp_forward_fused(const float *sph, int64_t N, int64_t C, int64_t H, int64_t W, int64_t sN,
                                               int64_t sC, int64_t sH, int64_t sW, const float *grid, int64_t gN,
                                               int64_t gC, int64_t gH, int64_t gW, int64_t gD, float in_scale,
                                               float in_bias, float *out, int64_t out_map_stride, int res,
                                               void *workspace, size_t workspace_bytes, void *stream) {
  GB_REQUIRE(out_map_stride >= (int64_t)res * res * res, GENRE_B200_EINVAL, "sph_bp fused: out_map_stride too small");
  return sph_forward_impl(sph, N, C, H, W, sN, sC, sH, sW, grid, gN, gC, gH, gW, gD, out, nullptr, res, workspace,
                          workspace_bytes, stream, true, in_scale, in_bias, 1.0f, -(1.0f / 16777216.0f), out_map_stride);
}

extern "C" int genre_b200_sph_bp_backward(const float *sph, int64_t N, int64_t C, int64_t H, int64_t W, int64_t sN,
                                          int64_t sC, int64_t sH, int64_t sW, const float *grid, int64_t gN, int64_t gC,
                                          int64_t gH, int64_t gW, int64_t gD, const float *cnt, const float *grad_tdf,
                                          int res, float *grad_sph, void *stream) {
  if (int rc = sph_check(sph, N, C, H, W, grid, res)) return rc;
  GB_REQUIRE(cnt && grad_tdf && grad_sph, GENRE_B200_EINVAL, "sph_bp backward: null pointer");
  const int64_t P = H * W;
  dim3 grd((unsigned)((P + SPH_THREADS - 1) / SPH_THREADS), (unsigned)(N * C));
  sph_bp_backward_kernel<<<grd, SPH_THREADS, 0, as_stream(stream)>>>(sph, (int)C, (int)H, (int)W, sN, sC, sH, sW, grid,
                                                                     gN, gC, gH, gW, gD, cnt, grad_tdf, res, grad_sph);
  return check_launch("sph_bp backward kernel");
}
