/*
 * genre_oracle.c — CPU restatement of the reference's toolbox ops.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product (libgenre_b200.so and the Python packages above it) never does: it has
 * no CPU path at all.
 *
 * Every function follows the cited lines of the reference (paths relative to the reference root,
 * /root/reference) operation by operation, single-threaded, in the reference's loop order.
 * fp32 arithmetic is done with one rounding per operation (build with -ffp-contract=off, see
 * oracle/Makefile); where nvcc contracts the reference's CUDA source into an FMA (read from the
 * sm_100a SASS of oracle/_ref/libref_cam_bp.so / libref_nnd_cuda.so) an explicit fmaf() is used, so
 * the voxel / neighbour indices are bit-identical to the reference kernels, not merely close.
 *
 * Parity pinning: the reference ships no golden vectors.  This restatement is pinned
 * (a) against the reference's own kernels compiled unmodified into oracle/_ref/ and run on the GPU box
 *     (tests/test_gpu_oracle_pin.py), and
 * (b) for nndistance against the reference's CPU code toolbox/nndistance/src/my_lib.c compiled
 *     unmodified into oracle/_ref/libref_nnd_cpu.so (tests/test_oracle_cpu.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* back_projection_kernel.cu:36-43 */
static inline int floor_i(float a) {
  /* CUDA's float->int conversion saturates; keep the C conversion defined for wild inputs */
  if (!(a > -2.0e9f && a < 2.0e9f)) return a < 0 ? INT32_MIN + 1 : INT32_MAX;
  return a < 0 ? (int)a - 1 : (int)a;
}
static inline int floor_i_d(double a) {
  if (!(a > -2.0e9 && a < 2.0e9)) return a < 0 ? INT32_MIN + 1 : INT32_MAX;
  return a < 0 ? (int)a - 1 : (int)a;
}
static inline int round_i_d(double a) { /* ROUND_I :42-43 with a double argument */
  int fi = floor_i_d(a);
  float ff = (float)fi;
  return (a - ff > ff + 1.0 - a) ? fi + 1 : fi;
}

typedef struct {
  float gx, gy, gz, imh, imw, norm;
  int ix, iy, iz, in_bounds;
} cam_point;

/* back_projection_kernel.cu:229-252 (forward) == :403-420 (backward) */
static cam_point cam_unproject(float d, float fl, float cam_dist, int h, int w, int H, int W, int R) {
  cam_point p;
  float Rf = (float)R;
  p.imh = (float)h - ((float)H - 1.0f) / 2.0f;
  p.imw = (float)w - ((float)W - 1.0f) / 2.0f;
  /* vec3d_norm(imind_h, imind_w, fl): nvcc -> FMUL h*h; FFMA w*w+; FFMA fl*fl+; IEEE sqrt */
  p.norm = sqrtf(fmaf(fl, fl, fmaf(p.imw, p.imw, p.imh * p.imh)));
  float cos_theta = fl / p.norm;
  float z = d * cos_theta;
  p.gy = (-z * p.imw) / fl;
  p.gz = (-z * p.imh) / fl;
  p.gx = z - cam_dist;
  p.ix = floor_i((p.gx + 0.5f) * Rf);
  p.iy = floor_i((p.gy + 0.5f) * Rf);
  p.iz = floor_i((p.gz + 0.5f) * Rf);
  p.in_bounds = p.ix >= 0 && p.ix < R && p.iy >= 0 && p.iy < R && p.iz >= 0 && p.iz < R;
  return p;
}

/* distance to the voxel centre, forward flavour (:256-263): fp32 centre; nvcc -> FMUL dy*dy; FFMA dx; FFMA dz */
static float centre_dist_f32(float gx, float gy, float gz, int ix, int iy, int iz, int R) {
  float Rf = (float)R;
  float cx = (((float)ix + 0.5f) / Rf) - 0.5f;
  float cy = (((float)iy + 0.5f) / Rf) - 0.5f;
  float cz = (((float)iz + 0.5f) / Rf) - 0.5f;
  float dx = gx - cx, dy = gy - cy, dz = gz - cz;
  return sqrtf(fmaf(dz, dz, fmaf(dx, dx, dy * dy)));
}

/* inplace_safe_divide :281-306 */
static void safe_divide(float *tdf, const float *cnt, size_t n, float bias, int R) {
  for (size_t i = 0; i < n; ++i) {
    float pt = cnt[i];
    if ((double)pt < 1e-5) continue;
    tdf[i] = (tdf[i] - bias / (float)R) / pt;
  }
}

/*
 * CameraBackProjection.forward: cam_back_projection.py:22-30 + back_projection_kernel.cu:199-306,760-838.
 * tdf/cnt [N,C,R,R,R] dense.  shift != 0 additionally applies camera_backprojection_module.py:26-28.
 * Accumulation order: ascending pixel index (the reference's float atomics have no defined order).
 */
ORACLE_API void oracle_cam_bp_forward(const float *depth, int N, int C, int H, int W, long sN, long sC, long sH,
                                      long sW, const float *fl, long fN, long fC, const float *cd, long dN, long dC,
                                      float *tdf, float *cnt, int R, int shift) {
  size_t nvox = (size_t)R * R * R;
  float init = (float)(1.0 / (double)R); /* python: zero_() + 1 / res */
  for (size_t i = 0; i < (size_t)N * C * nvox; ++i) {
    tdf[i] = 0.0f + init;
    cnt[i] = 0.0f;
  }
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c) {
      float *t = tdf + ((size_t)n * C + c) * nvox, *k = cnt + ((size_t)n * C + c) * nvox;
      float f = fl[n * fN + c * fC], cam = cd[n * dN + c * dC];
      for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
          float d = depth[n * sN + c * sC + h * sH + w * sW];
          if (d < 0.0f) continue; /* :225 */
          cam_point p = cam_unproject(d, f, cam, h, w, H, W, R);
          if (!p.in_bounds) continue; /* :252 */
          size_t v = ((size_t)p.ix * R + p.iy) * R + p.iz;
          t[v] += centre_dist_f32(p.gx, p.gy, p.gz, p.ix, p.iy, p.iz, R); /* :273 */
          k[v] += 1.0f;                                                     /* :274 */
        }
    }
  safe_divide(tdf, cnt, (size_t)N * C * nvox, 1.0f, R); /* :814-830, bias 1 */
  if (shift)
    for (size_t i = 0; i < (size_t)N * C * nvox; ++i) tdf[i] = 1.0f - (float)R * tdf[i];
}

/* the voxel index (or -1) of every pixel, [N,C,H,W] int32: the bit-exact part of the contract */
ORACLE_API void oracle_cam_bp_voxel_index(const float *depth, int N, int C, int H, int W, long sN, long sC, long sH,
                                          long sW, const float *fl, long fN, long fC, const float *cd, long dN,
                                          long dC, int32_t *vidx, int R) {
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c)
      for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
          float d = depth[n * sN + c * sC + h * sH + w * sW];
          int32_t out = -1;
          if (!(d < 0.0f)) {
            cam_point p = cam_unproject(d, fl[n * fN + c * fC], cd[n * dN + c * dC], h, w, H, W, R);
            if (p.in_bounds) out = (p.ix * R + p.iy) * R + p.iz;
          }
          vidx[(((size_t)n * C + c) * H + h) * W + w] = out;
        }
}

/*
 * CameraBackProjection.backward: cam_back_projection.py:34-46 + back_projection_kernel.cu:365-471,897-963,
 * with cam_dist indexed by its own strides (the reference's :401 uses cnt's strides: out of bounds for n >= 1).
 */
ORACLE_API void oracle_cam_bp_backward(const float *depth, int N, int C, int H, int W, long sN, long sC, long sH,
                                       long sW, const float *fl, long fN, long fC, const float *cd, long dN, long dC,
                                       const float *cnt, const float *grad_tdf, int R, float *grad_depth,
                                       float *grad_fl, float *grad_cd) {
  size_t nvox = (size_t)R * R * R;
  memset(grad_depth, 0, sizeof(float) * (size_t)N * C * H * W);
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c) {
      size_t map = (size_t)n * C + c;
      float f = fl[n * fN + c * fC], cam = cd[n * dN + c * dC];
      double acc_fl = 0.0, acc_cd = 0.0; /* reference: fp32 atomics in arbitrary order; summed in double here */
      for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
          float d = depth[n * sN + c * sC + h * sH + w * sW];
          if (d < 0.0f) continue;
          cam_point p = cam_unproject(d, f, cam, h, w, H, W, R);
          if (!p.in_bounds) continue;
          float Rf = (float)R;
          float cx = (float)((((double)(float)p.ix + 0.5) / (double)Rf) - 0.5); /* :428-430, double literals */
          float cy = (float)((((double)(float)p.iy + 0.5) / (double)Rf) - 0.5);
          float cz = (float)((((double)(float)p.iz + 0.5) / (double)Rf) - 0.5);
          float len = p.norm;
          if ((double)len < 1e-5) len = (float)1e-5;
          float dirx = -f / len, diry = p.imw / len, dirz = p.imh / len;
          float ex = p.gx - cx, ey = p.gy - cy, ez = p.gz - cz;
          float vlen = sqrtf(ex * ex + ey * ey + ez * ez);
          if ((double)vlen < 1e-5) vlen = (float)1e-5;
          float ux = ex / vlen, uy = ey / vlen, uz = ez / vlen;
          float cos_cc = dirx * ux + diry * uy + dirz * uz;
          size_t v = map * nvox + ((size_t)p.ix * R + p.iy) * R + p.iz;
          float ptnum = cnt[v];
          if (ptnum < 1) ptnum = 1;
          float gd = grad_tdf[v];
          grad_depth[(map * H + h) * W + w] = -gd * cos_cc / ptnum; /* :455 */
          float len3 = len * len * len;
          float gfx = ux * (p.imw * p.imw + p.imh * p.imh) / len3;
          float gfy = uy * (p.imw * f) / len3;
          float gfz = uz * (p.imh * f) / len3;
          acc_fl += (double)((gfx + gfy + gfz) * gd * d / ptnum); /* :459-464 */
          acc_cd += (double)(-ux * gd / ptnum);                    /* :469 */
        }
      grad_fl[map] = (float)acc_fl;
      grad_cd[map] = (float)acc_cd;
    }
}

/* get_surface_mask: functions/get_surface_mask.py:25-40 (mask part) + back_projection_kernel.cu:309-358,840-891 */
ORACLE_API void oracle_surface_mask(const float *depth, int N, int C, int H, int W, long sN, long sC, long sH, long sW,
                                    const float *fl, long fN, long fC, const float *cd, long dN, long dC,
                                    const float *cnt, float *mask, int R) {
  size_t nvox = (size_t)R * R * R;
  float Rf = (float)R;
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c) {
      size_t map = (size_t)n * C + c;
      float f = fl[n * fN + c * fC], cam = cd[n * dN + c * dC];
      for (int ix = 0; ix < R; ++ix)
        for (int iy = 0; iy < R; ++iy)
          for (int iz = 0; iz < R; ++iz) {
            size_t v = map * nvox + ((size_t)ix * R + iy) * R + iz;
            mask[v] = 1.0f; /* THCudaTensor_fill(mask, 1) :855 */
            if ((double)cnt[v] > 1e-5) continue;
            float cx = (float)((((double)(float)ix + 0.5) / (double)Rf) - 0.5);
            float cy = (float)((((double)(float)iy + 0.5) / (double)Rf) - 0.5);
            float cz = (float)((((double)(float)iz + 0.5) / (double)Rf) - 0.5);
            float den = cx + cam;
            float im_h = (-cz * f) / den;
            float im_w = (-cy * f) / den;
            int idh = round_i_d(0.5 * ((double)(float)H - 1.0) + (double)im_h);
            int idw = round_i_d(0.5 * ((double)(float)W - 1.0) + (double)im_w);
            if (idh < 0 || idh >= H || idw < 0 || idw >= W) continue;
            float d = depth[n * sN + c * sC + idh * sH + idw * sW];
            if (d < 0) continue;
            /* vec3d_norm(x+cam, y, z): nvcc -> FMUL den*den; FFMA cy; FFMA cz */
            float ray = sqrtf(fmaf(cz, cz, fmaf(cy, cy, den * den)));
            if (d < ray) mask[v] = 0.0f;
          }
    }
}

/* SphericalBackProjection.forward: sperical_to_tdf.py:24-31 + back_projection_kernel.cu:474-542,629-703 */
ORACLE_API void oracle_sph_bp_forward(const float *sph, int N, int C, int H, int W, long sN, long sC, long sH, long sW,
                                      const float *grid, long gN, long gC, long gH, long gW, long gD, float *tdf,
                                      float *cnt, int R) {
  size_t nvox = (size_t)R * R * R;
  float Rf = (float)R;
  memset(tdf, 0, sizeof(float) * (size_t)N * C * nvox);
  memset(cnt, 0, sizeof(float) * (size_t)N * C * nvox);
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c) {
      float *t = tdf + ((size_t)n * C + c) * nvox, *k = cnt + ((size_t)n * C + c) * nvox;
      for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
          float r = sph[n * sN + c * sC + h * sH + w * sW];
          const float *g = grid + n * gN + c * gC + h * gH + w * gW;
          if (r < 0.0f) continue;
          float gx = g[0] * r, gy = g[gD] * r, gz = g[2 * gD] * r;
          int ix = floor_i((gx + 0.5f) * Rf), iy = floor_i((gy + 0.5f) * Rf), iz = floor_i((gz + 0.5f) * Rf);
          if (!(ix >= 0 && ix < R && iy >= 0 && iy < R && iz >= 0 && iz < R)) continue;
          size_t v = ((size_t)ix * R + iy) * R + iz;
          t[v] += centre_dist_f32(gx, gy, gz, ix, iy, iz, R);
          k[v] += 1.0f;
        }
    }
  safe_divide(tdf, cnt, (size_t)N * C * nvox, 0.0f, R); /* :680-696, bias 0 */
}

/* SphericalBackProjection.backward: sperical_to_tdf.py:35-47 + back_projection_kernel.cu:544-627,704-757 */
ORACLE_API void oracle_sph_bp_backward(const float *sph, int N, int C, int H, int W, long sN, long sC, long sH,
                                       long sW, const float *grid, long gN, long gC, long gH, long gW, long gD,
                                       const float *cnt, const float *grad_tdf, int R, float *grad_sph) {
  size_t nvox = (size_t)R * R * R;
  float Rf = (float)R;
  memset(grad_sph, 0, sizeof(float) * (size_t)N * C * H * W);
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c) {
      size_t map = (size_t)n * C + c;
      for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
          float r = sph[n * sN + c * sC + h * sH + w * sW];
          const float *g = grid + n * gN + c * gC + h * gH + w * gW;
          if (r < 0.0f) continue;
          float gx = g[0] * r, gy = g[gD] * r, gz = g[2 * gD] * r;
          int ix = floor_i((gx + 0.5f) * Rf), iy = floor_i((gy + 0.5f) * Rf), iz = floor_i((gz + 0.5f) * Rf);
          if (!(ix >= 0 && ix < R && iy >= 0 && iy < R && iz >= 0 && iz < R)) continue;
          float cx = (float)((((double)(float)ix + 0.5) / (double)Rf) - 0.5);
          float cy = (float)((((double)(float)iy + 0.5) / (double)Rf) - 0.5);
          float cz = (float)((((double)(float)iz + 0.5) / (double)Rf) - 0.5);
          float len = sqrtf(gx * gx + gy * gy + gz * gz);
          if ((double)len < 1e-5) len = (float)1e-5;
          float ux = gx / len, uy = gy / len, uz = gz / len;
          float cos_cc = ux * cx + uy * cy + uz * cz;
          float ex = gx - cx, ey = gy - cy, ez = gz - cz;
          float dist = sqrtf(ex * ex + ey * ey + ez * ez);
          size_t v = map * nvox + ((size_t)ix * R + iy) * R + iz;
          float ptnum = cnt[v];
          if (ptnum < 1) ptnum = 1;
          if ((double)dist < 1e-5) dist = (float)1e-5;
          grad_sph[(map * H + h) * W + w] = grad_tdf[v] * (r - cos_cc) / (ptnum * dist); /* :621 */
        }
    }
}

/* calc_stop_forward_kernel, calc_prob_kernel.cu:120-142: fp64 step arithmetic (1.0 literals), fp32 storage */
ORACLE_API void oracle_calc_prob_forward(const float *prob, float *stop, long n_rays, int Z) {
  for (long r = 0; r < n_rays; ++r) {
    const float *p = prob + r * Z;
    float *s = stop + r * Z;
    for (int z = 0; z < Z; ++z) {
      if (z == 0) s[0] = p[0];
      else s[z] = (float)((double)s[z - 1] * ((1.0 / (double)p[z - 1]) - 1.0) * (double)p[z]);
    }
  }
}

/* calc_stop_backward_kernel, calc_prob_kernel.cu:155-188 */
ORACLE_API void oracle_calc_prob_backward(const float *prob, const float *wgt, float *grad, long n_rays, int Z) {
  for (long r = 0; r < n_rays; ++r) {
    const float *p = prob + r * Z, *sw = wgt + r * Z;
    float *g = grad + r * Z;
    float head = 0, delay_sum = 0;
    for (int z = Z - 1; z >= 0; --z) {
      if (z == Z - 1) {
        head = sw[z] / p[z];
        g[z] = head;
      } else {
        float cur = p[z], prev = p[z + 1];
        float v1 = sw[z] / cur;
        float v2 = (float)((double)(head * prev) / (1.0 - (double)cur));
        float v3 = (float)((double)delay_sum * (1.0 - (double)prev) / (double)(1 - cur));
        delay_sum = v2 + v3;
        head = v1;
        g[z] = v1 - v2 - v3;
      }
    }
  }
}

/*
 * render_spherical.forward, toolbox/spherical_proj.py:62-72, with the registered buffers passed in:
 *   grid [S,S,Z,3] fp32 (gen_grid :39-60), depth_weight [Z].  grid_sample = trilinear, zero padding,
 *   align_corners=True (torch 0.4.1 semantics, environment.yml:14) on vox.permute(0,1,4,3,2), i.e.
 *   grid (x,y,z) addresses vox dims (2,3,4).  Optionally returns the clamped samples prob [N,S,S,Z].
 */
static float trilinear(const float *vol, int R, float gx, float gy, float gz) {
  float Rm1 = (float)(R - 1);
  float fx = ((gx + 1.0f) * 0.5f) * Rm1, fy = ((gy + 1.0f) * 0.5f) * Rm1, fz = ((gz + 1.0f) * 0.5f) * Rm1;
  float x0f = floorf(fx), y0f = floorf(fy), z0f = floorf(fz);
  int x0 = (int)x0f, y0 = (int)y0f, z0 = (int)z0f;
  float wx[2] = {(x0f + 1.0f) - fx, fx - x0f}, wy[2] = {(y0f + 1.0f) - fy, fy - y0f},
        wz[2] = {(z0f + 1.0f) - fz, fz - z0f};
  float acc = 0.0f;
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b)
      for (int c = 0; c < 2; ++c) {
        int x = x0 + a, y = y0 + b, z = z0 + c;
        if (x < 0 || x >= R || y < 0 || y >= R || z < 0 || z >= R) continue;
        acc = fmaf(vol[((size_t)x * R + y) * R + z], wx[a] * wy[b] * wz[c], acc);
      }
  return acc;
}

ORACLE_API void oracle_render_spherical(const float *vox, int N, int R, const float *grid, int S, int Z,
                                        const float *depth_weight, float *out, float *prob_out) {
  const float pmin = 1e-5f, pmax = (float)(1.0 - 1e-5);
  float *p = (float *)malloc(sizeof(float) * Z), *s = (float *)malloc(sizeof(float) * Z);
  for (int n = 0; n < N; ++n)
    for (int i = 0; i < S * S; ++i) {
      const float *vol = vox + (size_t)n * R * R * R;
      for (int k = 0; k < Z; ++k) {
        const float *g = grid + ((size_t)i * Z + k) * 3;
        float v = trilinear(vol, R, g[0], g[1], g[2]);
        p[k] = v < pmin ? pmin : (v > pmax ? pmax : v);
      }
      oracle_calc_prob_forward(p, s, 1, Z);
      float e = 0.0f, bg = 1.0f;
      for (int k = 0; k < Z; ++k) {
        e += s[k] * depth_weight[k];
        bg *= 1.0f - p[k];
      }
      out[(size_t)n * S * S + i] = e + bg * 1.0f;
      if (prob_out) memcpy(prob_out + ((size_t)n * S * S + i) * Z, p, sizeof(float) * Z);
    }
  free(p);
  free(s);
}

/*
 * nndistance: restates the GPU kernel's arithmetic (nnd_cuda.cu:6-128): fp32 distance with the FMA
 * contraction nvcc applies, strict '<' over ascending candidates (lowest index on ties).
 * fused != 0 -> GPU rounding fma(dz,dz,fma(dx,dx,dy*dy)); fused == 0 -> the CPU code's rounding
 * (my_lib.c:15-19: three rounded products summed left to right, compared as double).
 */
ORACLE_API void oracle_nnsearch(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int32_t *idx,
                                int fused) {
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < n; ++j) {
      float x1 = xyz1[((size_t)i * n + j) * 3 + 0], y1 = xyz1[((size_t)i * n + j) * 3 + 1],
            z1 = xyz1[((size_t)i * n + j) * 3 + 2];
      float best = 0;
      int besti = 0;
      for (int k = 0; k < m; ++k) {
        float x2 = xyz2[((size_t)i * m + k) * 3 + 0] - x1, y2 = xyz2[((size_t)i * m + k) * 3 + 1] - y1,
              z2 = xyz2[((size_t)i * m + k) * 3 + 2] - z1;
        float d = fused ? fmaf(z2, z2, fmaf(x2, x2, y2 * y2)) : (x2 * x2 + y2 * y2) + z2 * z2;
        if (k == 0 || d < best) {
          best = d;
          besti = k;
        }
      }
      dist[(size_t)i * n + j] = best;
      idx[(size_t)i * n + j] = besti;
    }
}

/* nnd_backward, my_lib.c:50-118 (same arithmetic as NmDistanceGradKernel nnd_cuda.cu:143-162) */
ORACLE_API void oracle_nnd_backward(int b, int n, int m, const float *xyz1, const float *xyz2, const float *g1,
                                    const float *g2, const int32_t *idx1, const int32_t *idx2, float *grad1,
                                    float *grad2) {
  memset(grad1, 0, sizeof(float) * (size_t)b * n * 3);
  memset(grad2, 0, sizeof(float) * (size_t)b * m * 3);
  for (int i = 0; i < b; ++i) {
    for (int j = 0; j < n; ++j) {
      int j2 = idx1[(size_t)i * n + j];
      float g = g1[(size_t)i * n + j] * 2;
      for (int a = 0; a < 3; ++a) {
        float v = g * (xyz1[((size_t)i * n + j) * 3 + a] - xyz2[((size_t)i * m + j2) * 3 + a]);
        grad1[((size_t)i * n + j) * 3 + a] += v;
        grad2[((size_t)i * m + j2) * 3 + a] -= v;
      }
    }
    for (int j = 0; j < m; ++j) {
      int j2 = idx2[(size_t)i * m + j];
      float g = g2[(size_t)i * m + j] * 2;
      for (int a = 0; a < 3; ++a) {
        float v = g * (xyz2[((size_t)i * m + j) * 3 + a] - xyz1[((size_t)i * n + j2) * 3 + a]);
        grad2[((size_t)i * m + j) * 3 + a] += v;
        grad1[((size_t)i * n + j2) * 3 + a] -= v;
      }
    }
  }
}
