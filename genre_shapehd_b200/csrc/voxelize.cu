// voxelize.cu — workspace + splat stage of the voxelisation pipeline (see voxelize.cuh for the design).
#include "voxelize.cuh"

namespace gb {

// ------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static size_t counters_bytes(int64_t n_maps, size_t nt) { return align_up(((size_t)n_maps * nt + (size_t)n_maps) * 4, 256); }

size_t vox_workspace_bytes(int64_t n_maps, int64_t P, int res) {
  const size_t nt = (size_t)vox_ntiles(res);
  size_t b = counters_bytes(n_maps, nt);                         // counts + ovf_count (contiguous: one memset)
  b += align_up((size_t)n_maps * nt * VOX_BUCKET * 8, 256);      // buckets
  b += align_up((size_t)n_maps * P * 8, 256);                    // overflow lists
  return b;
}

bool vox_carve(void *ws, size_t ws_bytes, int64_t n_maps, int64_t P, int res, VoxWorkspace *out) {
  if (!ws || !aligned16(ws) || ws_bytes < vox_workspace_bytes(n_maps, P, res)) return false;
  const size_t nt = (size_t)vox_ntiles(res);
  char *p = (char *)ws;
  out->counts = (unsigned *)p;
  out->ovf_count = out->counts + (size_t)n_maps * nt;
  p += counters_bytes(n_maps, nt);
  out->buckets = (uint2 *)p;
  p += align_up((size_t)n_maps * nt * VOX_BUCKET * 8, 256);
  out->ovf = (uint2 *)p;
  out->ntiles = (int)nt;
  return true;
}

int vox_clear_counts(const VoxWorkspace &w, int64_t n_maps, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(w.counts, 0, ((size_t)n_maps * w.ntiles + (size_t)n_maps) * 4, st);
  if (e != cudaSuccess) {
    set_error("voxelize: clearing tile counters: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// splat: one CTA per output tile
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float vox_finalize(unsigned lo, unsigned hi, float alpha, float beta, float bg,
                                              float &count_out) {
  const unsigned c = hi >> 12;
  count_out = (float)c;
  if (c == 0) return bg;
  const unsigned long long sum = ((unsigned long long)(hi & 0xFFFu) << 32) | lo;
  return fmaf(beta, __ull2float_rn(sum) / (float)c, alpha);
}

constexpr int SPLAT_KEEP = VOX_BUCKET / VOX_SPLAT_THREADS;  // the whole bucket fits in registers (4 records/thread)

template <bool VEC, bool WRITE_CNT>
__global__ void __launch_bounds__(VOX_SPLAT_THREADS)
vox_splat_kernel(const uint2 *__restrict__ buckets, const uint2 *__restrict__ ovf,
                 const unsigned *__restrict__ counts, const unsigned *__restrict__ ovf_count,
                 float *__restrict__ tdf, float *__restrict__ cnt, long long P, long long nvox, int ntiles,
                 float alpha, float beta, float bg, long long out_stride) {
  __shared__ __align__(16) unsigned s_lo[VOX_TILE];
  __shared__ __align__(16) unsigned s_hi[VOX_TILE];
  const int tile = blockIdx.x, map = blockIdx.y, tid = threadIdx.x;
  const size_t tix = (size_t)map * ntiles + tile;
  // Programmatic dependent launch: this grid may become resident while the project kernel drains; nothing the
  // project kernel wrote is read before this point.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const unsigned n = counts[tix];
  const long long start = (long long)tile * VOX_TILE;
  const int nv = (int)min((long long)VOX_TILE, nvox - start);
  float *out = tdf + (size_t)map * out_stride + start;  // out_stride > nvox: a channel of a wider tensor
  float *cout = WRITE_CNT ? cnt + (size_t)map * nvox + start : nullptr;

  if (n == 0) {  // background-only tile: pure streaming fill, no shared memory touched
    if (VEC) {
      const float4 b4 = make_float4(bg, bg, bg, bg), z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = tid * 4; j < VOX_TILE; j += VOX_SPLAT_THREADS * 4) {
        if (j < nv) {
          st_stream_f4(out + j, b4);
          if (WRITE_CNT) st_stream_f4(cout + j, z4);
        }
      }
    } else {
      for (int j = tid; j < nv; j += VOX_SPLAT_THREADS) {
        st_stream_f1(out + j, bg);
        if (WRITE_CNT) st_stream_f1(cout + j, 0.f);
      }
    }
    return;
  }

  // the bucket is fetched before the accumulators are cleared so the load latency overlaps the clearing
  const uint2 *seg = buckets + tix * VOX_BUCKET;
  const unsigned nb = min(n, (unsigned)VOX_BUCKET);
  uint2 r[SPLAT_KEEP];
#pragma unroll
  for (int k = 0; k < SPLAT_KEEP; ++k) {
    const unsigned i = tid + k * VOX_SPLAT_THREADS;
    r[k] = i < nb ? seg[i] : make_uint2(0, 0);
  }
#pragma unroll
  for (int j = tid * 4; j < VOX_TILE; j += VOX_SPLAT_THREADS * 4) {
    *reinterpret_cast<uint4 *>(s_lo + j) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4 *>(s_hi + j) = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  auto add = [&](unsigned v, unsigned q) {
    const unsigned old = atomicAdd(&s_lo[v], q);
    const unsigned carry = (old + q < old) ? 1u : 0u;
    atomicAdd(&s_hi[v], (1u << 12) + carry);
  };
#pragma unroll
  for (int k = 0; k < SPLAT_KEEP; ++k)
    if (tid + k * VOX_SPLAT_THREADS < nb) add(r[k].x, r[k].y);
  if (n > (unsigned)VOX_BUCKET) {  // CTA-uniform, rare: this tile spilled; pick its records out of the map's list
    const unsigned novf = ovf_count[map];
    const uint2 *list = ovf + (size_t)map * P;
    for (unsigned i = tid; i < novf; i += VOX_SPLAT_THREADS) {
      const uint2 x = list[i];
      if (x.x / VOX_TILE == (unsigned)tile) add(x.x - tile * VOX_TILE, x.y);
    }
  }
  __syncthreads();

  if (VEC) {
#pragma unroll
    for (int j = tid * 4; j < VOX_TILE; j += VOX_SPLAT_THREADS * 4) {
      if (j < nv) {
        const uint4 lo = *reinterpret_cast<const uint4 *>(s_lo + j);
        const uint4 hi = *reinterpret_cast<const uint4 *>(s_hi + j);
        float4 o, c;
        o.x = vox_finalize(lo.x, hi.x, alpha, beta, bg, c.x);
        o.y = vox_finalize(lo.y, hi.y, alpha, beta, bg, c.y);
        o.z = vox_finalize(lo.z, hi.z, alpha, beta, bg, c.z);
        o.w = vox_finalize(lo.w, hi.w, alpha, beta, bg, c.w);
        st_stream_f4(out + j, o);
        if (WRITE_CNT) st_stream_f4(cout + j, c);
      }
    }
  } else {
    for (int j = tid; j < nv; j += VOX_SPLAT_THREADS) {
      float c;
      const float o = vox_finalize(s_lo[j], s_hi[j], alpha, beta, bg, c);
      st_stream_f1(out + j, o);
      if (WRITE_CNT) st_stream_f1(cout + j, c);
    }
  }
}

template <bool VEC, bool WRITE_CNT>
static int launch_splat(const VoxWorkspace &w, int64_t n_maps, int64_t P, long long nvox, float *tdf, float *cnt,
                        float alpha, float beta, float bg, cudaStream_t st, bool pdl, long long out_stride) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)w.ntiles, (unsigned)n_maps);
  cfg.blockDim = dim3(VOX_SPLAT_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // overlap this launch with the project kernel's tail
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;  // stand-alone launches (stage entry point, used for timing) keep plain stream ordering
  const uint2 *buckets = w.buckets, *ovf = w.ovf;
  const unsigned *counts = w.counts, *ovf_count = w.ovf_count;
  const long long Pll = (long long)P;
  const int ntiles = w.ntiles;
  cudaError_t e = cudaLaunchKernelEx(&cfg, vox_splat_kernel<VEC, WRITE_CNT>, buckets, ovf, counts, ovf_count, tdf, cnt,
                                     Pll, nvox, ntiles, alpha, beta, bg, out_stride);
  if (e != cudaSuccess) {
    set_error("voxelize splat kernel: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return (int)e;
  }
  return check_launch("voxelize splat kernel");
}

int vox_splat(const VoxWorkspace &w, int64_t n_maps, int64_t P, int res, float *tdf, float *cnt, float alpha,
              float beta, float bg, cudaStream_t st, bool pdl, long long out_stride) {
  const long long nvox = (long long)res * res * res;
  if (out_stride <= 0) out_stride = nvox;
  const bool vec = (nvox % 4 == 0) && (out_stride % 4 == 0) && aligned16(tdf) && (!cnt || aligned16(cnt));
  if (vec) {
    return cnt ? launch_splat<true, true>(w, n_maps, P, nvox, tdf, cnt, alpha, beta, bg, st, pdl, out_stride)
               : launch_splat<true, false>(w, n_maps, P, nvox, tdf, cnt, alpha, beta, bg, st, pdl, out_stride);
  }
  return cnt ? launch_splat<false, true>(w, n_maps, P, nvox, tdf, cnt, alpha, beta, bg, st, pdl, out_stride)
             : launch_splat<false, false>(w, n_maps, P, nvox, tdf, cnt, alpha, beta, bg, st, pdl, out_stride);
}

// arguments shared by both back-projections
int vox_check_common(int64_t n_maps, int64_t P, int res) {
  GB_REQUIRE(n_maps > 0 && n_maps <= 65535, GENRE_B200_EINVAL, "N*C = %lld must be in [1, 65535]", (long long)n_maps);
  GB_REQUIRE(P > 0 && P <= VOX_MAX_PIXELS, GENRE_B200_EINVAL, "H*W = %lld must be in [1, 2^20)", (long long)P);
  GB_REQUIRE(res > 0 && (int64_t)res * res * res < (1ll << 31), GENRE_B200_EINVAL, "voxel resolution %d unsupported", res);
  GB_REQUIRE(vox_ntiles(res) <= VOX_MAX_TILES, GENRE_B200_EINVAL, "voxel resolution %d unsupported (too many tiles)", res);
  return 0;
}

}  // namespace gb

// ------------------------------------------------------------------------------------------------
// C ABI: workspace size + stage entry point
// ------------------------------------------------------------------------------------------------
extern "C" size_t genre_b200_voxelize_workspace_bytes(int64_t n_maps, int64_t pixels_per_map, int res) {
  if (n_maps <= 0 || pixels_per_map <= 0 || res <= 0) return 0;
  return gb::vox_workspace_bytes(n_maps, pixels_per_map, res);
}

extern "C" int genre_b200_voxelize_stage_splat(int64_t n_maps, int64_t P, int res, float *tdf, float *cnt,
                                               float hit_alpha, float hit_beta, float background, void *workspace,
                                               size_t workspace_bytes, void *stream) {
  if (int rc = gb::vox_check_common(n_maps, P, res)) return rc;
  GB_REQUIRE(tdf != nullptr, GENRE_B200_EINVAL, "tdf is null");
  gb::VoxWorkspace w;
  GB_REQUIRE(gb::vox_carve(workspace, workspace_bytes, n_maps, P, res, &w), GENRE_B200_EWORKSPACE,
             "workspace too small or misaligned (need %zu bytes)", gb::vox_workspace_bytes(n_maps, P, res));
  return gb::vox_splat(w, n_maps, P, res, tdf, cnt, hit_alpha, hit_beta, background, gb::as_stream(stream), false);
}
