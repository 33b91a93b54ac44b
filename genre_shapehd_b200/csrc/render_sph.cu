// render_sph.cu — fused spherical renderer (voxel occupancy -> spherical depth map), forward + backward.
//
// Reference: toolbox/spherical_proj.py:31-72 (render_spherical):
//     grid   = dirs * 2 * (1 - linspace(0,1,Z))            [S,S,Z,3], fp64 numpy -> fp32      (:39-60)
//     prob   = grid_sample(vox.permute(0,1,4,3,2), grid)   trilinear, zero padding, torch-0.4.1
//              semantics == align_corners=True                                               (:63-65)
//     prob   = clamp(prob, 1e-5, 1 - 1e-5)                                                    (:66)
//     stop   = CalcStopProb(prob)                          calc_prob_kernel.cu:112-143        (:67)
//     out    = matmul(stop, linspace(0,1,Z)) + prod(1 - prob, dim=4)                          (:68-71)
// The reference materialises four [N,1,S,S,Z] tensors (256 MiB each at N=16).  Here one warp owns one
// ray: lanes take consecutive samples along the ray (neighbouring samples are ~0.5 voxel apart, so the
// 8-tap gathers of a warp land in a handful of cache lines), the transmittance crosses lanes with a
// shuffle product-scan and crosses 32-sample chunks through a register carry.  Only the [N,S,S] map
// is written.  Sample positions are formed in fp64 from the fp64 direction table exactly like the
// numpy code, then rounded to fp32, so they equal the reference's registered `grid` buffer.
#include "common.cuh"

namespace gb {

constexpr int RS_THREADS = 256;  // 8 rays per CTA

struct Taps {
  int base;        // linear index of corner (x0, y0, z0); may be out of range, see masks
  float w[8];      // weights in order (x,y,z) = 000, 001, 010, 011, 100, 101, 110, 111
  unsigned valid;  // bit i set if tap i is inside the volume
};

__device__ __forceinline__ void make_taps(float gx, float gy, float gz, int R, Taps &t) {
  const float Rm1 = (float)(R - 1);
  // grid_sampler unnormalise, align_corners=True: ((coord + 1) / 2) * (size - 1)
  const float fx = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), Rm1);
  const float fy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), Rm1);
  const float fz = __fmul_rn(__fmul_rn(__fadd_rn(gz, 1.0f), 0.5f), Rm1);
  const float x0f = floorf(fx), y0f = floorf(fy), z0f = floorf(fz);
  // clamp before the int conversion so far-away samples (|coord| up to 2) stay well defined
  const int x0 = (int)fminf(fmaxf(x0f, -2.0f), (float)R), y0 = (int)fminf(fmaxf(y0f, -2.0f), (float)R),
            z0 = (int)fminf(fmaxf(z0f, -2.0f), (float)R);
  const float wx1 = fx - x0f, wx0 = (x0f + 1.0f) - fx;
  const float wy1 = fy - y0f, wy0 = (y0f + 1.0f) - fy;
  const float wz1 = fz - z0f, wz0 = (z0f + 1.0f) - fz;
  const bool vx0 = (unsigned)x0 < (unsigned)R, vx1 = (unsigned)(x0 + 1) < (unsigned)R;
  const bool vy0 = (unsigned)y0 < (unsigned)R, vy1 = (unsigned)(y0 + 1) < (unsigned)R;
  const bool vz0 = (unsigned)z0 < (unsigned)R, vz1 = (unsigned)(z0 + 1) < (unsigned)R;
  t.base = (x0 * R + y0) * R + z0;
  t.w[0] = wx0 * wy0 * wz0; t.w[1] = wx0 * wy0 * wz1; t.w[2] = wx0 * wy1 * wz0; t.w[3] = wx0 * wy1 * wz1;
  t.w[4] = wx1 * wy0 * wz0; t.w[5] = wx1 * wy0 * wz1; t.w[6] = wx1 * wy1 * wz0; t.w[7] = wx1 * wy1 * wz1;
  t.valid = (unsigned)(vx0 & vy0 & vz0) | ((unsigned)(vx0 & vy0 & vz1) << 1) | ((unsigned)(vx0 & vy1 & vz0) << 2) |
            ((unsigned)(vx0 & vy1 & vz1) << 3) | ((unsigned)(vx1 & vy0 & vz0) << 4) |
            ((unsigned)(vx1 & vy0 & vz1) << 5) | ((unsigned)(vx1 & vy1 & vz0) << 6) |
            ((unsigned)(vx1 & vy1 & vz1) << 7);
}

__device__ __forceinline__ int tap_offset(int i, int R) {
  return ((i >> 2) & 1) * R * R + ((i >> 1) & 1) * R + (i & 1);
}

// Optional transform of every voxel value as it is fetched: v -> clamp(v * scale, lo, hi).  GenRe renders
// clamp(proj * 50, 1e-5, 1 - 1e-5) (depth_pred_with_sph_inpaint.py:124): applying the two elementwise ops here (same
// fp32 operations, same order) removes two dense passes over the volume.  Out-of-volume taps stay 0 (zero padding).
struct VoxPre {
  float scale, lo, hi;
};
template <bool PRE>
__device__ __forceinline__ float fetch_vox(const float *__restrict__ p, const VoxPre &pre) {
  const float v = __ldg(p);
  return PRE ? fminf(fmaxf(__fmul_rn(v, pre.scale), pre.lo), pre.hi) : v;
}

template <bool PRE>
__device__ __forceinline__ float sample_trilinear(const float *__restrict__ vol, const Taps &t, int R, const VoxPre &pre) {
  float acc = 0.0f;
  if (t.valid == 0xFFu) {  // interior: 4 pairs of z-adjacent taps
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(fetch_vox<PRE>(vol + t.base + tap_offset(i, R), pre), t.w[i], acc);
  } else if (t.valid) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (t.valid & (1u << i)) acc = fmaf(fetch_vox<PRE>(vol + t.base + tap_offset(i, R), pre), t.w[i], acc);
  }
  return acc;
}

// sample position k on the ray with fp64 direction (dx,dy,dz): (float)(dir*2*(1-alpha_k)), alpha = np.linspace(0,1,Z)
__device__ __forceinline__ void ray_point(double dx2, double dy2, double dz2, int k, int Z, double step, float &gx,
                                          float &gy, float &gz) {
  const double alpha = (k == Z - 1 && Z > 1) ? 1.0 : (double)k * step;
  const double f = 1.0 - alpha;
  gx = (float)(dx2 * f);
  gy = (float)(dy2 * f);
  gz = (float)(dz2 * f);
}

__device__ __forceinline__ float warp_excl_prod32(float v, float &total) {
  float incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float u = __shfl_up_sync(0xffffffffu, incl, d);
    if ((threadIdx.x & 31) >= d) incl *= u;
  }
  total = __shfl_sync(0xffffffffu, incl, 31);
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if ((threadIdx.x & 31) == 0) excl = 1.0f;
  return excl;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr float RS_PMIN = 1e-5f;
constexpr float RS_PMAX = (float)(1.0 - 1e-5);

// forward pass over one ray; returns (sum_k s_k w_k, prod_k (1 - p_k)) to every lane
template <bool PRE>
__device__ __forceinline__ void render_ray(const float *__restrict__ vol, int R, double dx2, double dy2, double dz2,
                                           int Z, const float *__restrict__ depth_weight, const VoxPre &pre,
                                           float &exp_depth, float &trans) {
  const int lane = threadIdx.x & 31;
  const double step = Z > 1 ? 1.0 / (double)(Z - 1) : 0.0;
  float carry = 1.0f, acc = 0.0f;
  for (int k0 = 0; k0 < Z; k0 += 32) {
    const int k = k0 + lane;
    float p = 0.0f;  // lanes past the end behave like p = 0 (factor 1, no contribution)
    if (k < Z) {
      float gx, gy, gz;
      ray_point(dx2, dy2, dz2, k, Z, step, gx, gy, gz);
      Taps t;
      make_taps(gx, gy, gz, R, t);
      p = fminf(fmaxf(sample_trilinear<PRE>(vol, t, R, pre), RS_PMIN), RS_PMAX);
    }
    float total;
    const float before = carry * warp_excl_prod32(1.0f - p, total);
    if (k < Z) acc = fmaf(p * before, __ldg(depth_weight + k), acc);
    carry *= total;
    if (carry == 0.0f) break;  // transmittance underflowed: every later term is exactly 0 (warp-uniform)
  }
  exp_depth = warp_sum(acc);
  trans = carry;
}

template <bool PRE>
__global__ void __launch_bounds__(RS_THREADS)
render_spherical_forward_kernel(const float *__restrict__ vox, int R, const double *__restrict__ dirs, int S, int Z,
                                const float *__restrict__ depth_weight, float *__restrict__ out, long long n_rays,
                                const VoxPre pre) {
  const long long ray = (long long)blockIdx.x * (RS_THREADS / 32) + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const int pix = (int)(ray % ((long long)S * S));
  const long long n = ray / ((long long)S * S);
  const float *vol = vox + (size_t)n * R * R * R;
  const double dx2 = dirs[pix * 3 + 0] * 2.0, dy2 = dirs[pix * 3 + 1] * 2.0, dz2 = dirs[pix * 3 + 2] * 2.0;
  float e, t;
  render_ray<PRE>(vol, R, dx2, dy2, dz2, Z, depth_weight, pre, e, t);
  if ((threadIdx.x & 31) == 0) out[ray] = e + t;
}

// ---- empty-space skipping (forward) ---------------------------------------------------------------------------------
// GenRe renders a thin shell: clamp(proj * 50, 1e-5, 1 - 1e-5) is 1e-5 everywhere except on the ~1 % of voxels the depth
// map hit, and a third of every ray lies outside the volume, yet the plain kernel pays 8 gathers + fp64 positions for each
// of its 4.19 M samples per shape.  A sample whose 8 taps are all <= 1e-5 (or outside: zero padding) clamps to p = 1e-5
// (up to a 1e-12 rounding of the interpolation): a run of such samples has a closed form.
//   pre-pass  (render_occupancy_kernel): one read of the volume -> a bit per 8^3 brick, set when any voxel within the brick
//             DILATED by one voxel exceeds 1e-5 (the dilation covers the +1 taps and the fp32 position estimate below);
//   render    per 32-sample chunk each lane locates its sample's brick from an fp32 estimate of the position (3 FMAs);
//             if the whole chunk is empty (warp vote):  acc += 1e-5 * T * CW[c],  T *= Q[c]
//             with CW[c] = sum_j q^j w_{32c+j}, Q[c] = q^(samples in chunk), q = 1 - 1e-5 (tables built per CTA);
//             otherwise the chunk takes the exact path (fp64 positions, 8 taps), lanes on empty samples skipping their gathers.
// Error of the closed form against the sample-by-sample product: a few 1e-7 relative on T (bounded by 1e-6 over a ray).
constexpr int RS_BRICK = 8;
constexpr int RS_RAYS_PER_WARP = 4;
constexpr int RS_MAX_Z_CHUNKS = 32;  // Z <= 1024 on the skipping path
constexpr float RS_LOG2_Q = -1.4427022e-05f;  // log2(1 - 1e-5)

__host__ __device__ inline int rs_bricks(int R) { return (R + RS_BRICK - 1) / RS_BRICK; }
__host__ __device__ inline int rs_occ_words(int R) {
  const int nb = rs_bricks(R);
  return (nb * nb * nb + 31) / 32;
}

template <bool PRE>
__global__ void __launch_bounds__(256)
render_occupancy_kernel(const float *__restrict__ vox, int R, long long vox_per_vol4, unsigned *__restrict__ occ,
                        const VoxPre pre) {
  // one thread per 4 z-adjacent voxels (R % 4 == 0); blockIdx.y = volume
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= vox_per_vol4) return;
  const int n = blockIdx.y;
  const float4 v = *reinterpret_cast<const float4 *>(vox + ((size_t)n * vox_per_vol4 + i4) * 4);
  float a[4] = {v.x, v.y, v.z, v.w};
  int zlo = 4, zhi = -1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float t = PRE ? fminf(fmaxf(__fmul_rn(a[i], pre.scale), pre.lo), pre.hi) : a[i];
    if (!(t <= RS_PMIN)) {  // NaN counts as occupied
      zlo = min(zlo, i);
      zhi = max(zhi, i);
    }
  }
  if (zhi < 0) return;
  const long long lin = i4 * 4;
  const int z = (int)(lin % R), y = (int)((lin / R) % R), x = (int)(lin / ((long long)R * R));
  const int nb = rs_bricks(R);
  const int bx0 = max(x - 1, 0) / RS_BRICK, bx1 = min(x + 1, R - 1) / RS_BRICK;
  const int by0 = max(y - 1, 0) / RS_BRICK, by1 = min(y + 1, R - 1) / RS_BRICK;
  const int bz0 = max(z + zlo - 1, 0) / RS_BRICK, bz1 = min(z + zhi + 1, R - 1) / RS_BRICK;
  unsigned *o = occ + (size_t)n * rs_occ_words(R);
  for (int bx = bx0; bx <= bx1; ++bx)
    for (int by = by0; by <= by1; ++by)
      for (int bz = bz0; bz <= bz1; ++bz) {
        const int bit = (bx * nb + by) * nb + bz;
        const unsigned m = 1u << (bit & 31);
        if (!(o[bit >> 5] & m)) atomicOr(&o[bit >> 5], m);
      }
}

template <bool PRE>
__global__ void __launch_bounds__(RS_THREADS)
render_spherical_forward_skip_kernel(const float *__restrict__ vox, int R, const double *__restrict__ dirs, int S, int Z,
                                     const float *__restrict__ depth_weight, const unsigned *__restrict__ occ,
                                     float *__restrict__ out, const VoxPre pre) {
  extern __shared__ unsigned rs_smem[];
  const int words = rs_occ_words(R), nchunk = (Z + 31) / 32;
  unsigned *s_occ = rs_smem;
  float *s_cw = reinterpret_cast<float *>(rs_smem + words);  // [nchunk]
  float *s_qn = s_cw + RS_MAX_Z_CHUNKS;                       // [nchunk]
  const int n = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < words; i += RS_THREADS) s_occ[i] = occ[(size_t)n * words + i];
  const float qpow = exp2f((float)lane * RS_LOG2_Q);
  for (int c = warp; c < nchunk; c += RS_THREADS / 32) {
    const int k = c * 32 + lane;
    const float cw = warp_sum(k < Z ? qpow * __ldg(depth_weight + k) : 0.0f);
    if (lane == 0) {
      s_cw[c] = cw;
      s_qn[c] = exp2f((float)min(32, Z - c * 32) * RS_LOG2_Q);
    }
  }
  __syncthreads();
  const int nb = rs_bricks(R);
  const float h = 0.5f * (float)(R - 1), stepf = Z > 1 ? 1.0f / (float)(Z - 1) : 0.0f;
  const double step = Z > 1 ? 1.0 / (double)(Z - 1) : 0.0;
  const float *vol = vox + (size_t)n * R * R * R;
  const int pix0 = (blockIdx.x * (RS_THREADS / 32) + warp) * RS_RAYS_PER_WARP;
  for (int rr = 0; rr < RS_RAYS_PER_WARP; ++rr) {
    const int pix = pix0 + rr;
    if (pix >= S * S) break;  // warp-uniform
    const double dx = dirs[pix * 3 + 0], dy = dirs[pix * 3 + 1], dz = dirs[pix * 3 + 2];
    const double dx2 = dx * 2.0, dy2 = dy * 2.0, dz2 = dz * 2.0;
    const float dxh = (float)dx * h, dyh = (float)dy * h, dzh = (float)dz * h;
    float carry = 1.0f, acc = 0.0f, acc_u = 0.0f;
    for (int c = 0; c < nchunk; ++c) {
      const int k = c * 32 + lane;
      bool empty = true;
      if (k < Z) {
        const float r = 2.0f * (1.0f - (float)k * stepf);  // |position| in normalised units (0 at the centre)
        const float fx = fmaf(dxh, r, h), fy = fmaf(dyh, r, h), fz = fmaf(dzh, r, h);
        const float lo = -1.01f, hi = (float)R + 0.01f;      // taps floor(f), floor(f)+1: all invalid outside (-1, R)
        if (fx > lo && fx < hi && fy > lo && fy < hi && fz > lo && fz < hi) {
          const int bx = min(max((int)floorf(fx), 0), R - 1) / RS_BRICK, by = min(max((int)floorf(fy), 0), R - 1) / RS_BRICK,
                    bz = min(max((int)floorf(fz), 0), R - 1) / RS_BRICK;
          const int bit = (bx * nb + by) * nb + bz;
          empty = !((s_occ[bit >> 5] >> (bit & 31)) & 1u);
        }
      }
      if (__all_sync(0xffffffffu, empty)) {
        acc_u = fmaf(RS_PMIN * carry, s_cw[c], acc_u);
        carry *= s_qn[c];
        continue;
      }
      float p = 0.0f;  // lanes past the end behave like p = 0 (factor 1, no contribution)
      if (k < Z) {
        p = RS_PMIN;
        if (!empty) {
          float gx, gy, gz;
          ray_point(dx2, dy2, dz2, k, Z, step, gx, gy, gz);
          Taps t;
          make_taps(gx, gy, gz, R, t);
          p = fminf(fmaxf(sample_trilinear<PRE>(vol, t, R, pre), RS_PMIN), RS_PMAX);
        }
      }
      float total;
      const float before = carry * warp_excl_prod32(1.0f - p, total);
      if (k < Z) acc = fmaf(p * before, __ldg(depth_weight + k), acc);
      carry *= total;
      if (carry == 0.0f) break;  // transmittance underflowed: every later term is exactly 0 (warp-uniform)
    }
    const float e = warp_sum(acc) + acc_u;
    if (lane == 0) out[(size_t)n * S * S + pix] = e + carry;
  }
}

// backward: d out / d p_k = T_k w_k - (A_k + T_Z) / (1 - p_k),  A_k = sum_{m>k} w_m s_m, scattered through the
// trilinear weights (zero where the clamp was active, as torch.clamp's backward).  Pass 1 walks the ray
// forward keeping the raw samples and prefix transmittances in registers; pass 2 walks it backward so
// A_k is a plain running sum of non-negative terms (no cancellation against 1/(1-p) ~ 1e5).
constexpr int RS_MAX_CHUNKS = 8;  // backward keeps Z/32 samples per lane in registers -> Z <= 256

__global__ void __launch_bounds__(RS_THREADS)
render_spherical_backward_kernel(const float *__restrict__ vox, int R, const double *__restrict__ dirs, int S, int Z,
                                 const float *__restrict__ depth_weight, const float *__restrict__ grad_out,
                                 float *__restrict__ grad_vox, long long n_rays) {
  const long long ray = (long long)blockIdx.x * (RS_THREADS / 32) + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const int lane = threadIdx.x & 31;
  const int pix = (int)(ray % ((long long)S * S));
  const long long n = ray / ((long long)S * S);
  const float *vol = vox + (size_t)n * R * R * R;
  float *gvol = grad_vox + (size_t)n * R * R * R;
  const float g = grad_out[ray];
  if (g == 0.0f) return;
  const double dx2 = dirs[pix * 3 + 0] * 2.0, dy2 = dirs[pix * 3 + 1] * 2.0, dz2 = dirs[pix * 3 + 2] * 2.0;
  const double step = Z > 1 ? 1.0 / (double)(Z - 1) : 0.0;

  float raw[RS_MAX_CHUNKS], Tk[RS_MAX_CHUNKS];
  float carry = 1.0f;
  int last_chunk = 0;
#pragma unroll
  for (int c = 0; c < RS_MAX_CHUNKS; ++c) {
    raw[c] = 0.0f;
    Tk[c] = 0.0f;
    const int k = c * 32 + lane;
    if (c * 32 < Z && carry != 0.0f) {  // warp-uniform
      float p = 0.0f;
      if (k < Z) {
        float gx, gy, gz;
        ray_point(dx2, dy2, dz2, k, Z, step, gx, gy, gz);
        Taps t;
        make_taps(gx, gy, gz, R, t);
        raw[c] = sample_trilinear<false>(vol, t, R, VoxPre{});
        p = fminf(fmaxf(raw[c], RS_PMIN), RS_PMAX);
      }
      float total;
      Tk[c] = carry * warp_excl_prod32(1.0f - p, total);
      carry *= total;
      last_chunk = c;
    }
  }
  const float t_all = carry;  // prod over the whole ray (0 if it underflowed; later samples then have zero gradient)

  float suffix = 0.0f;  // sum_{m in later chunks} w_m s_m
#pragma unroll
  for (int c = RS_MAX_CHUNKS - 1; c >= 0; --c) {
    if (c > last_chunk) continue;  // warp-uniform
    const int k = c * 32 + lane;
    const bool live = k < Z;
    const float p = live ? fminf(fmaxf(raw[c], RS_PMIN), RS_PMAX) : 0.0f;
    const float wk = live ? __ldg(depth_weight + k) : 0.0f;
    const float ws = wk * p * Tk[c];
    // exclusive suffix sum over lanes
    float incl = ws;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float u = __shfl_down_sync(0xffffffffu, incl, d);
      if (lane + d < 32) incl += u;
    }
    float excl = __shfl_down_sync(0xffffffffu, incl, 1);
    if (lane == 31) excl = 0.0f;
    const float Ak = suffix + excl;
    suffix += __shfl_sync(0xffffffffu, incl, 0);
    if (live && raw[c] >= RS_PMIN && raw[c] <= RS_PMAX) {
      const float dp = g * (Tk[c] * wk - (Ak + t_all) / (1.0f - p));
      if (dp != 0.0f) {
        float gx, gy, gz;
        ray_point(dx2, dy2, dz2, k, Z, step, gx, gy, gz);
        Taps t;
        make_taps(gx, gy, gz, R, t);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (t.valid & (1u << i)) atomicAdd(gvol + t.base + tap_offset(i, R), dp * t.w[i]);
      }
    }
  }
}

static int rs_check(const float *vox, int64_t N, int res, const double *dirs, int S, int Z, const float *dw) {
  GB_REQUIRE(vox && dirs && dw, GENRE_B200_EINVAL, "render_spherical: null pointer");
  GB_REQUIRE(N > 0 && res >= 2 && (int64_t)res * res * res < (1ll << 31) && S > 0 && Z > 0, GENRE_B200_EINVAL,
             "render_spherical: bad shape (N=%lld, R=%d, S=%d, Z=%d)", (long long)N, res, S, Z);
  GB_REQUIRE(((int64_t)N * S * S + 7) / 8 < (1ll << 31), GENRE_B200_EINVAL, "render_spherical: too many rays");
  return 0;
}

}  // namespace gb

using namespace gb;

extern "C" int genre_b200_render_spherical_forward(const float *vox, int64_t N, int res, const double *dirs,
                                                   int sph_res, int z_res, const float *depth_weight, float *out,
                                                   void *stream) {
  if (int rc = rs_check(vox, N, res, dirs, sph_res, z_res, depth_weight)) return rc;
  GB_REQUIRE(out != nullptr, GENRE_B200_EINVAL, "render_spherical: out is null");
  const long long n_rays = (long long)N * sph_res * sph_res;
  const unsigned grid = (unsigned)((n_rays + RS_THREADS / 32 - 1) / (RS_THREADS / 32));
  render_spherical_forward_kernel<false><<<grid, RS_THREADS, 0, as_stream(stream)>>>(vox, res, dirs, sph_res, z_res,
                                                                                     depth_weight, out, n_rays, VoxPre{});
  return check_launch("render_spherical forward kernel");
}

// Same renderer over clamp(vox * pre_scale, pre_lo, pre_hi) without materialising that volume (forward only: the
// fused GenRe inference path, depth_pred_with_sph_inpaint.py:124 `render_spherical(clamp(proj * 50, 1e-5, 1 - 1e-5))`).
extern "C" int genre_b200_render_spherical_forward_pre(const float *vox, int64_t N, int res, const double *dirs,
                                                       int sph_res, int z_res, const float *depth_weight,
                                                       float pre_scale, float pre_lo, float pre_hi, float *out,
                                                       void *stream) {
  if (int rc = rs_check(vox, N, res, dirs, sph_res, z_res, depth_weight)) return rc;
  GB_REQUIRE(out != nullptr, GENRE_B200_EINVAL, "render_spherical: out is null");
  GB_REQUIRE(pre_lo <= pre_hi, GENRE_B200_EINVAL, "render_spherical: empty clamp range");
  const long long n_rays = (long long)N * sph_res * sph_res;
  const unsigned grid = (unsigned)((n_rays + RS_THREADS / 32 - 1) / (RS_THREADS / 32));
  render_spherical_forward_kernel<true><<<grid, RS_THREADS, 0, as_stream(stream)>>>(
      vox, res, dirs, sph_res, z_res, depth_weight, out, n_rays, VoxPre{pre_scale, pre_lo, pre_hi});
  return check_launch("render_spherical forward kernel (pre-transform)");
}

extern "C" int genre_b200_render_spherical_backward(const float *vox, int64_t N, int res, const double *dirs,
                                                    int sph_res, int z_res, const float *depth_weight,
                                                    const float *grad_out, float *grad_vox, void *stream) {
  if (int rc = rs_check(vox, N, res, dirs, sph_res, z_res, depth_weight)) return rc;
  GB_REQUIRE(grad_out && grad_vox, GENRE_B200_EINVAL, "render_spherical backward: null pointer");
  GB_REQUIRE(z_res <= 32 * RS_MAX_CHUNKS, GENRE_B200_EINVAL, "render_spherical backward: z_res %d > %d unsupported",
             z_res, 32 * RS_MAX_CHUNKS);
  const long long n_rays = (long long)N * sph_res * sph_res;
  const unsigned grid = (unsigned)((n_rays + RS_THREADS / 32 - 1) / (RS_THREADS / 32));
  render_spherical_backward_kernel<<<grid, RS_THREADS, 0, as_stream(stream)>>>(vox, res, dirs, sph_res, z_res,
                                                                               depth_weight, grad_out, grad_vox, n_rays);
  return check_launch("render_spherical backward kernel");
}

extern "C" size_t genre_b200_render_spherical_workspace_bytes(int64_t N, int res) {
  if (N <= 0 || res < 2) return 0;
  return (size_t)N * rs_occ_words(res) * sizeof(unsigned);
}

// The same renderer with empty-space skipping (see the header of the skipping section): identical results up to ~1e-6.
//   use_pre != 0: render clamp(vox * pre_scale, pre_lo, pre_hi) without materialising it
//   workspace: genre_b200_render_spherical_workspace_bytes(N, res) bytes, caller-owned, zeroed here (memset node)
// Supported: res % 4 == 0, z_res <= 1024, 16-byte aligned vox; otherwise the call is forwarded to the plain kernel.
extern "C" int genre_b200_render_spherical_forward_skip(const float *vox, int64_t N, int res, const double *dirs,
                                                        int sph_res, int z_res, const float *depth_weight, int use_pre,
                                                        float pre_scale, float pre_lo, float pre_hi, float *out,
                                                        void *workspace, size_t workspace_bytes, void *stream) {
  if (int rc = rs_check(vox, N, res, dirs, sph_res, z_res, depth_weight)) return rc;
  GB_REQUIRE(out != nullptr, GENRE_B200_EINVAL, "render_spherical: out is null");
  GB_REQUIRE(!use_pre || pre_lo <= pre_hi, GENRE_B200_EINVAL, "render_spherical: empty clamp range");
  const size_t need = genre_b200_render_spherical_workspace_bytes(N, res);
  const bool can_skip = res % 4 == 0 && z_res <= 32 * RS_MAX_Z_CHUNKS && aligned16(vox) && N < 65536;
  if (!can_skip) {
    if (use_pre) return genre_b200_render_spherical_forward_pre(vox, N, res, dirs, sph_res, z_res, depth_weight, pre_scale, pre_lo, pre_hi, out, stream);
    return genre_b200_render_spherical_forward(vox, N, res, dirs, sph_res, z_res, depth_weight, out, stream);
  }
  GB_REQUIRE(workspace && workspace_bytes >= need && ((uintptr_t)workspace & 3) == 0, GENRE_B200_EINVAL,
             "render_spherical: workspace of %zu bytes needed (got %zu)", need, workspace_bytes);
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(workspace, 0, need, st);
  if (e != cudaSuccess) return fail_arg((int)e, "render_spherical: cudaMemsetAsync: %s", cudaGetErrorString(e));
  const VoxPre pre = use_pre ? VoxPre{pre_scale, pre_lo, pre_hi} : VoxPre{};
  const long long v4 = (long long)res * res * res / 4;
  dim3 og((unsigned)((v4 + 255) / 256), (unsigned)N);
  if (use_pre) render_occupancy_kernel<true><<<og, 256, 0, st>>>(vox, res, v4, (unsigned *)workspace, pre);
  else render_occupancy_kernel<false><<<og, 256, 0, st>>>(vox, res, v4, (unsigned *)workspace, pre);
  if (int rc = check_launch("render_spherical occupancy kernel")) return rc;
  const int rays_per_cta = (RS_THREADS / 32) * RS_RAYS_PER_WARP;
  dim3 rg((unsigned)((sph_res * sph_res + rays_per_cta - 1) / rays_per_cta), (unsigned)N);
  const size_t smem = (size_t)rs_occ_words(res) * 4 + 2 * RS_MAX_Z_CHUNKS * sizeof(float);
  if (use_pre)
    render_spherical_forward_skip_kernel<true><<<rg, RS_THREADS, smem, st>>>(vox, res, dirs, sph_res, z_res, depth_weight,
                                                                             (const unsigned *)workspace, out, pre);
  else
    render_spherical_forward_skip_kernel<false><<<rg, RS_THREADS, smem, st>>>(vox, res, dirs, sph_res, z_res, depth_weight,
                                                                              (const unsigned *)workspace, out, pre);
  return check_launch("render_spherical forward kernel (empty-space skipping)");
}
