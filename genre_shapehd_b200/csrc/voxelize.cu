// voxelize.cu — bin + splat stages of the voxelisation pipeline (see voxelize.cuh for the design).
#include "voxelize.cuh"

namespace gb {

// ------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

size_t vox_workspace_bytes(int64_t n_maps, int64_t P, int res) {
  const size_t nt = (size_t)vox_ntiles(res);
  size_t b = 0;
  b += align_up((size_t)n_maps * nt * 4, 256) * 2;  // counts, offsets
  b += align_up((size_t)n_maps * P * 4, 256) * 3;   // pix_gv, pix_q, pix_rank
  b += align_up((size_t)n_maps * P * 8, 256);       // sorted
  return b;
}

bool vox_carve(void *ws, size_t ws_bytes, int64_t n_maps, int64_t P, int res, VoxWorkspace *out) {
  if (!ws || !aligned16(ws) || ws_bytes < vox_workspace_bytes(n_maps, P, res)) return false;
  const size_t nt = (size_t)vox_ntiles(res);
  char *p = (char *)ws;
  out->counts = (unsigned *)p;   p += align_up((size_t)n_maps * nt * 4, 256);
  out->offsets = (unsigned *)p;  p += align_up((size_t)n_maps * nt * 4, 256);
  out->pix_gv = (unsigned *)p;   p += align_up((size_t)n_maps * P * 4, 256);
  out->pix_q = (unsigned *)p;    p += align_up((size_t)n_maps * P * 4, 256);
  out->pix_rank = (unsigned *)p; p += align_up((size_t)n_maps * P * 4, 256);
  out->sorted = (uint2 *)p;
  out->ntiles = (int)nt;
  return true;
}

int vox_clear_counts(const VoxWorkspace &w, int64_t n_maps, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(w.counts, 0, (size_t)n_maps * w.ntiles * 4, st);
  if (e != cudaSuccess) {
    set_error("voxelize: clearing tile counters: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// bin: per-map exclusive scan of the tile counters + scatter of the pixel records into tile order
// ------------------------------------------------------------------------------------------------
constexpr int BIN_THREADS = 256;
constexpr int BIN_PIX_PER_THREAD = 4;

__global__ void __launch_bounds__(BIN_THREADS)
vox_bin_kernel(const unsigned *__restrict__ counts, unsigned *__restrict__ offsets,
               const unsigned *__restrict__ pix_gv, const unsigned *__restrict__ pix_q,
               const unsigned *__restrict__ pix_rank, uint2 *__restrict__ sorted, int P, int ntiles) {
  extern __shared__ unsigned s_off[];  // [ntiles] exclusive offsets of this map
  __shared__ unsigned s_warp[BIN_THREADS / 32];
  const int map = blockIdx.y;
  const int tid = threadIdx.x;
  const unsigned *cmap = counts + (size_t)map * ntiles;

  // each thread owns a contiguous chunk of tiles; block-scan the chunk sums
  const int chunk = (ntiles + BIN_THREADS - 1) / BIN_THREADS;
  const int t0 = min(tid * chunk, ntiles), t1 = min(t0 + chunk, ntiles);
  unsigned local = 0;
  for (int t = t0; t < t1; ++t) local += cmap[t];
  unsigned incl = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
    if ((tid & 31) >= d) incl += v;
  }
  if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
  __syncthreads();
  unsigned warp_base = 0;
  for (int w = 0; w < (tid >> 5); ++w) warp_base += s_warp[w];
  unsigned run = warp_base + incl - local;
  for (int t = t0; t < t1; ++t) {
    s_off[t] = run;
    run += cmap[t];
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    unsigned *omap = offsets + (size_t)map * ntiles;
    for (int t = tid; t < ntiles; t += BIN_THREADS) omap[t] = s_off[t];
  }

  const size_t mbase = (size_t)map * P;
  const int p0 = blockIdx.x * (BIN_THREADS * BIN_PIX_PER_THREAD) + tid;
  unsigned gv[BIN_PIX_PER_THREAD];
#pragma unroll
  for (int k = 0; k < BIN_PIX_PER_THREAD; ++k) {
    const int p = p0 + k * BIN_THREADS;
    gv[k] = (p < P) ? pix_gv[mbase + p] : VOX_INVALID;
  }
#pragma unroll
  for (int k = 0; k < BIN_PIX_PER_THREAD; ++k) {
    if (gv[k] == VOX_INVALID) continue;
    const int p = p0 + k * BIN_THREADS;
    const unsigned tile = gv[k] / VOX_TILE;
    const unsigned dst = s_off[tile] + pix_rank[mbase + p];
    sorted[mbase + dst] = make_uint2(gv[k] - tile * VOX_TILE, pix_q[mbase + p]);
  }
}

int vox_bin(const VoxWorkspace &w, int64_t n_maps, int64_t P, cudaStream_t st) {
  dim3 grid((unsigned)((P + BIN_THREADS * BIN_PIX_PER_THREAD - 1) / (BIN_THREADS * BIN_PIX_PER_THREAD)),
            (unsigned)n_maps);
  vox_bin_kernel<<<grid, BIN_THREADS, (size_t)w.ntiles * 4, st>>>(w.counts, w.offsets, w.pix_gv, w.pix_q,
                                                                 w.pix_rank, w.sorted, (int)P, w.ntiles);
  return check_launch("voxelize bin kernel");
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

template <bool VEC, bool WRITE_CNT>
__global__ void __launch_bounds__(VOX_SPLAT_THREADS)
vox_splat_kernel(const uint2 *__restrict__ sorted, const unsigned *__restrict__ counts,
                 const unsigned *__restrict__ offsets, float *__restrict__ tdf, float *__restrict__ cnt,
                 int P, long long nvox, int ntiles, float alpha, float beta, float bg) {
  extern __shared__ __align__(16) unsigned s_acc[];  // lo[VOX_TILE] then hi[VOX_TILE]
  unsigned *s_lo = s_acc;
  unsigned *s_hi = s_acc + VOX_TILE;
  const int tile = blockIdx.x, map = blockIdx.y, tid = threadIdx.x;
  const unsigned n = counts[(size_t)map * ntiles + tile];
  const long long start = (long long)tile * VOX_TILE;
  const int nv = (int)min((long long)VOX_TILE, nvox - start);
  float *out = tdf + (size_t)map * nvox + start;
  float *cout = WRITE_CNT ? cnt + (size_t)map * nvox + start : nullptr;

  if (n == 0) {  // background-only tile: pure streaming fill, no shared memory touched
    if (VEC) {
      const float4 b4 = make_float4(bg, bg, bg, bg), z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = tid * 4; j < nv; j += VOX_SPLAT_THREADS * 4) {
        st_stream_f4(out + j, b4);
        if (WRITE_CNT) st_stream_f4(cout + j, z4);
      }
    } else {
      for (int j = tid; j < nv; j += VOX_SPLAT_THREADS) {
        st_stream_f1(out + j, bg);
        if (WRITE_CNT) st_stream_f1(cout + j, 0.f);
      }
    }
    return;
  }

  // zero the accumulators (both halves are contiguous)
  {
    uint4 *z = reinterpret_cast<uint4 *>(s_acc);
    for (int j = tid; j < 2 * VOX_TILE / 4; j += VOX_SPLAT_THREADS) z[j] = make_uint4(0, 0, 0, 0);
  }
  const uint2 *seg = sorted + (size_t)map * P + offsets[(size_t)map * ntiles + tile];
  __syncthreads();
  for (unsigned i = tid; i < n; i += VOX_SPLAT_THREADS) {
    const uint2 r = seg[i];
    const unsigned old = atomicAdd(&s_lo[r.x], r.y);
    const unsigned carry = (old + r.y < old) ? 1u : 0u;
    atomicAdd(&s_hi[r.x], (1u << 12) + carry);
  }
  __syncthreads();

  if (VEC) {
    for (int j = tid * 4; j < nv; j += VOX_SPLAT_THREADS * 4) {
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
                        float alpha, float beta, float bg, cudaStream_t st) {
  auto kern = vox_splat_kernel<VEC, WRITE_CNT>;
  const size_t smem = 2 * VOX_TILE * sizeof(unsigned);
  static bool configured_dev[64] = {};  // per template instantiation, per device
  int dev = 0;
  cudaGetDevice(&dev);
  bool &configured = configured_dev[dev & 63];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("voxelize splat: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  dim3 grid((unsigned)w.ntiles, (unsigned)n_maps);
  kern<<<grid, VOX_SPLAT_THREADS, smem, st>>>(w.sorted, w.counts, w.offsets, tdf, cnt, (int)P, nvox, w.ntiles,
                                              alpha, beta, bg);
  return check_launch("voxelize splat kernel");
}

int vox_splat(const VoxWorkspace &w, int64_t n_maps, int64_t P, int res, float *tdf, float *cnt, float alpha,
              float beta, float bg, cudaStream_t st) {
  const long long nvox = (long long)res * res * res;
  const bool vec = (nvox % 4 == 0) && aligned16(tdf) && (!cnt || aligned16(cnt));
  if (vec) {
    return cnt ? launch_splat<true, true>(w, n_maps, P, nvox, tdf, cnt, alpha, beta, bg, st)
               : launch_splat<true, false>(w, n_maps, P, nvox, tdf, cnt, alpha, beta, bg, st);
  }
  return cnt ? launch_splat<false, true>(w, n_maps, P, nvox, tdf, cnt, alpha, beta, bg, st)
             : launch_splat<false, false>(w, n_maps, P, nvox, tdf, cnt, alpha, beta, bg, st);
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
// C ABI: workspace size + stage entry points
// ------------------------------------------------------------------------------------------------
extern "C" size_t genre_b200_voxelize_workspace_bytes(int64_t n_maps, int64_t pixels_per_map, int res) {
  if (n_maps <= 0 || pixels_per_map <= 0 || res <= 0) return 0;
  return gb::vox_workspace_bytes(n_maps, pixels_per_map, res);
}

extern "C" int genre_b200_voxelize_stage_bin(int64_t n_maps, int64_t P, int res, void *workspace,
                                             size_t workspace_bytes, void *stream) {
  if (int rc = gb::vox_check_common(n_maps, P, res)) return rc;
  gb::VoxWorkspace w;
  GB_REQUIRE(gb::vox_carve(workspace, workspace_bytes, n_maps, P, res, &w), GENRE_B200_EWORKSPACE,
             "workspace too small or misaligned (need %zu bytes)", gb::vox_workspace_bytes(n_maps, P, res));
  return gb::vox_bin(w, n_maps, P, gb::as_stream(stream));
}

extern "C" int genre_b200_voxelize_stage_splat(int64_t n_maps, int64_t P, int res, float *tdf, float *cnt,
                                               float hit_alpha, float hit_beta, float background, void *workspace,
                                               size_t workspace_bytes, void *stream) {
  if (int rc = gb::vox_check_common(n_maps, P, res)) return rc;
  GB_REQUIRE(tdf != nullptr, GENRE_B200_EINVAL, "tdf is null");
  gb::VoxWorkspace w;
  GB_REQUIRE(gb::vox_carve(workspace, workspace_bytes, n_maps, P, res, &w), GENRE_B200_EWORKSPACE,
             "workspace too small or misaligned (need %zu bytes)", gb::vox_workspace_bytes(n_maps, P, res));
  return gb::vox_splat(w, n_maps, P, res, tdf, cnt, hit_alpha, hit_beta, background, gb::as_stream(stream));
}
