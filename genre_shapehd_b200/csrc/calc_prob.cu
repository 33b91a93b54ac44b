// calc_prob.cu — stop probability along rays, forward and backward.
//
// Reference: toolbox/calc_prob/calc_prob/src/calc_prob_kernel.cu
//   forward  :112-143  one THREAD per ray, 256 dependent global round trips through its own output:
//            s[0] = p[0];  s[z] = s[z-1] * (1/p[z-1] - 1) * p[z]        ( == p[z] * prod_{k<z} (1 - p[k]) )
//   backward :145-189  reverse recurrence producing  w[j]/p[j] - (sum_{z>j} w[z]) / (1 - p[j]),
//            w = stop_prob * grad_stop (functions/calc_prob.py:27)
// Here: one WARP per ray.  Each lane owns 4 consecutive samples of a 128-sample chunk (one 16-byte
// load), the running product / suffix sum crosses lanes with shuffles and crosses chunks through a
// register carry, so every element is read once and written once, fully coalesced.
// The closed forms are evaluated in fp32; the reference evaluates each step in fp64 and rounds to
// fp32 per step — the two agree to ~1e-6 relative (both are ~256 roundings away from the exact value).
#include "common.cuh"

namespace gb {

constexpr int CP_THREADS = 256;          // 8 rays per CTA
constexpr int CP_CHUNK = 128;            // samples per warp iteration

__device__ __forceinline__ float warp_excl_prod(float v, float &total) {
  // inclusive product scan over lanes, then shift
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

__device__ __forceinline__ float warp_excl_suffix_sum(float v, float &total) {
  // inclusive suffix sum over lanes (lane 31 first), then shift
  float incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float u = __shfl_down_sync(0xffffffffu, incl, d);
    if ((threadIdx.x & 31) + d < 32) incl += u;
  }
  total = __shfl_sync(0xffffffffu, incl, 0);
  float excl = __shfl_down_sync(0xffffffffu, incl, 1);
  if ((threadIdx.x & 31) == 31) excl = 0.0f;
  return excl;
}

template <bool VEC>
__global__ void __launch_bounds__(CP_THREADS)
calc_prob_forward_kernel(const float *__restrict__ prob, float *__restrict__ stop, long long n_rays, int Z) {
  const int lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * (CP_THREADS / 32) + (threadIdx.x >> 5);
  if (ray >= n_rays) return;  // whole warp exits together
  const float *p = prob + ray * Z;
  float *s = stop + ray * Z;
  float carry = 1.0f;  // prod_{k < chunk start} (1 - p[k])
  for (int z0 = 0; z0 < Z; z0 += CP_CHUNK) {
    const int z = z0 + lane * 4;
    float v[4];
    if (VEC) {
      if (z < Z) {
        const float4 t = *reinterpret_cast<const float4 *>(p + z);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
        v[0] = v[1] = v[2] = v[3] = 0.0f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = (z + i < Z) ? p[z + i] : 0.0f;
    }
    const float t0 = 1.0f - v[0], t1 = 1.0f - v[1], t2 = 1.0f - v[2], t3 = 1.0f - v[3];
    const float e1 = t0, e2 = t0 * t1, e3 = e2 * t2;
    float total;
    const float before = carry * warp_excl_prod(e3 * t3, total);
    float o[4] = {v[0] * before, v[1] * (before * e1), v[2] * (before * e2), v[3] * (before * e3)};
    if (VEC) {
      if (z < Z) *reinterpret_cast<float4 *>(s + z) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (z + i < Z) s[z + i] = o[i];
    }
    carry *= total;
  }
}

template <bool VEC>
__global__ void __launch_bounds__(CP_THREADS)
calc_prob_backward_kernel(const float *__restrict__ prob, const float *__restrict__ wgt, float *__restrict__ grad,
                          long long n_rays, int Z) {
  const int lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * (CP_THREADS / 32) + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  const float *p = prob + ray * Z;
  const float *wv = wgt + ray * Z;
  float *g = grad + ray * Z;
  float carry = 0.0f;  // sum_{z >= chunk end} w[z]
  const int nchunks = (Z + CP_CHUNK - 1) / CP_CHUNK;
  for (int ch = nchunks - 1; ch >= 0; --ch) {
    const int z = ch * CP_CHUNK + lane * 4;
    float pv[4], ww[4];
    if (VEC) {
      if (z < Z) {
        const float4 a = *reinterpret_cast<const float4 *>(p + z);
        const float4 b = *reinterpret_cast<const float4 *>(wv + z);
        pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w;
        ww[0] = b.x; ww[1] = b.y; ww[2] = b.z; ww[3] = b.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) { pv[i] = 0.5f; ww[i] = 0.0f; }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = z + i < Z;
        pv[i] = ok ? p[z + i] : 0.5f;
        ww[i] = ok ? wv[z + i] : 0.0f;
      }
    }
    // suffix sums inside the lane: a_i = sum_{j > i, same lane} w_j
    const float a3 = 0.0f, a2 = ww[3], a1 = ww[3] + ww[2], a0 = a1 + ww[1];
    float total;
    const float after = carry + warp_excl_suffix_sum(a0 + ww[0], total);
    float o[4];
    // the second term vanishes when nothing follows (the ray's last sample: calc_prob_kernel.cu:169-172 takes head = w/p there);
    // skipping it when the suffix sum is exactly 0 keeps p == 1 at such a sample finite (0/0 otherwise)
    const float sfx[4] = {after + a0, after + a1, after + a2, after + a3};
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = ww[i] / pv[i] - (sfx[i] != 0.0f ? sfx[i] / (1.0f - pv[i]) : 0.0f);
    if (VEC) {
      if (z < Z) *reinterpret_cast<float4 *>(g + z) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (z + i < Z) g[z + i] = o[i];
    }
    carry += total;
  }
}

static int cp_check(const void *a, const void *b, int64_t n_rays, int64_t Z) {
  GB_REQUIRE(a && b, GENRE_B200_EINVAL, "calc_prob: null pointer");
  GB_REQUIRE(n_rays > 0 && Z > 0 && Z < (1ll << 30), GENRE_B200_EINVAL, "calc_prob: bad shape [%lld rays, %lld samples]",
             (long long)n_rays, (long long)Z);
  GB_REQUIRE((n_rays + 7) / 8 < (1ll << 31), GENRE_B200_EINVAL, "calc_prob: too many rays");
  return 0;
}

}  // namespace gb

using namespace gb;

extern "C" int genre_b200_calc_prob_forward(const float *prob_in, float *stop_prob, int64_t n_rays, int64_t Z,
                                            void *stream) {
  if (int rc = cp_check(prob_in, stop_prob, n_rays, Z)) return rc;
  const unsigned grid = (unsigned)((n_rays + CP_THREADS / 32 - 1) / (CP_THREADS / 32));
  const bool vec = (Z % 4 == 0) && aligned16(prob_in) && aligned16(stop_prob);
  if (vec)
    calc_prob_forward_kernel<true><<<grid, CP_THREADS, 0, as_stream(stream)>>>(prob_in, stop_prob, n_rays, (int)Z);
  else
    calc_prob_forward_kernel<false><<<grid, CP_THREADS, 0, as_stream(stream)>>>(prob_in, stop_prob, n_rays, (int)Z);
  return check_launch("calc_prob forward kernel");
}

extern "C" int genre_b200_calc_prob_backward(const float *prob_in, const float *stop_prob_weighted, float *grad_prob,
                                             int64_t n_rays, int64_t Z, void *stream) {
  if (int rc = cp_check(prob_in, grad_prob, n_rays, Z)) return rc;
  GB_REQUIRE(stop_prob_weighted != nullptr, GENRE_B200_EINVAL, "calc_prob backward: null pointer");
  const unsigned grid = (unsigned)((n_rays + CP_THREADS / 32 - 1) / (CP_THREADS / 32));
  const bool vec = (Z % 4 == 0) && aligned16(prob_in) && aligned16(stop_prob_weighted) && aligned16(grad_prob);
  if (vec)
    calc_prob_backward_kernel<true><<<grid, CP_THREADS, 0, as_stream(stream)>>>(prob_in, stop_prob_weighted, grad_prob,
                                                                                n_rays, (int)Z);
  else
    calc_prob_backward_kernel<false><<<grid, CP_THREADS, 0, as_stream(stream)>>>(prob_in, stop_prob_weighted,
                                                                                 grad_prob, n_rays, (int)Z);
  return check_launch("calc_prob backward kernel");
}
