// convt_c1_col2im.cu — ConvTranspose3d(Cin -> 1, k 4, s 2, p 1) on 64-wide volumes as a tcgen05 GEMM over the 64 kernel taps
// followed by a shared-memory col2im: the last layer of every decoder at full size.
//
// Reference layers: networks/networks.py:167-168 (Unet_3D.dec6 = cat(dec5, enc1) -> ConvTranspose3d(2 nf -> 1) on 64^3),
// :57 (VoxelDecoder main.17: 32 -> 1 on 64^3), :98 (VoxelGenerator: 64 -> 1).
//
// With ONE output channel the usual implicit GEMM has N = 8 parity classes: 3% of an MMA's width (convt3d.cu MODE 4), and the
// exact FP32 stencil (convt_c1.cu) is bound by shared-memory / L1 operand traffic at 22% of the FP32 peak.  Transposing the
// roles of taps and outputs gives a real GEMM:
//     P[j, k] = sum_c x[c, j] * W[c, k]          M = input positions j,  N = 64 taps k = (kz,ky,kx),  K = Cin
//     out[2j - 1 + k] += P[j, k]                 col2im: every output voxel is the sum of 8 entries of P
// The A operand of the GEMM is the channel-blocked activation itself (no halo, no im2col): 128 consecutive positions of one
// (plane, channel group) are 128 contiguous 16-byte units = the K-major no-swizzle operand, fetched with plain bulk copies.
//
// CTA = one (sample, band of 8 input rows), sweeping z.  Per z step: 512 positions (4 M-tiles of 2 rows x 64) x 64 taps are
// accumulated in TMEM (fp16 operands; OP 2 = the fp32-accurate hi/lo split of convt3d.cu: columns [0,64) hi*hi, [64,128) the
// 2^11-scaled cross terms), then 16 epilogue warps (one thread per position) add them into a ring of four output planes in
// shared memory.  Along x the two contributions of an output are summed in registers (warp shuffles; the two lanes at a warp
// seam exchange through a scratch), and the taps are processed in 4 phases (kz>>1, ky>>1): within a phase a position adds one
// float2 (outputs 2x, 2x+1) to each of 4 (kz, ky) rows and different positions hit different cells, so the adds are race-free
// without atomics; a barrier separates phases (shared-memory traffic: 1/4 of a per-tap scatter, and conflict-free).  Output
// planes 2z-1 and 2z are complete after step z (their other contributions came from step z-1, carried in the ring) and leave
// with 16-byte stores.  Only the two output rows on each side of a band are shared with the neighbouring band's CTA: those are
// added with red.global.add.v4.f32 onto zeroed rows (two commutative contributions: deterministic), zeroed by one strided
// memset before the launch.  The bias is added by exactly one contributor of every output.
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace gb {

constexpr int CI_W = 64;             // input width (positions per row)
constexpr int CI_ROWS = 8;           // input rows per band
constexpr int CI_POS = CI_W * CI_ROWS;   // positions per z step = 4 M-tiles
constexpr int CI_MT = 4;
constexpr int CI_EPI_WARPS = 16;          // one warp per (M-tile, TMEM lane quarter): a thread owns one input position per z step
constexpr int CI_THREADS = (CI_EPI_WARPS + 2) * 32;   // + MMA warp + producer warp
constexpr int CI_MAX_STAGES = 6;
constexpr int CI_UY = 2 * CI_ROWS + 2;   // output rows touched by a band
constexpr int CI_PITCH = 136;            // floats per ring row: column ux + 3 holds output x = ux - 1 (x = 0 is 16-byte aligned)
constexpr int CI_PLANE = CI_UY * CI_PITCH;

struct ColParams {
  const __half *src0, *src1;   // blocked [B*D][parts*cg][H][64][8 fp16]
  int cg0, cg1;                // channel groups of 8 per part
  int B, D, H;
  const __half *wpack;         // [ksteps][2 kcore][NACC/8][8][8]
  int ksteps, stages;
  const float *bias;           // 1 value
  float *out;                  // [B][2D][2H][128]
};

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// 16 accumulator columns of this thread's row, no wait (tmem_ld_wait() before the registers are read)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void red_add_v4(float *p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int OP>
__global__ void __launch_bounds__(CI_THREADS, 1) convt_c1_col2im_kernel(const ColParams p) {
  constexpr bool X2 = OP == 2;
  constexpr int PARTS = X2 ? 2 : 1;
  constexpr int NACC = PARTS * 64;
  constexpr int CG_BYTES = CI_POS * 16;                  // one channel group of one part: 8 rows x 64 positions x 16 B
  constexpr int STAGE_BYTES = PARTS * 2 * CG_BYTES;      // one K step (16 channels)
  constexpr int W_KSTEP_BYTES = 2 * (NACC / 8) * 128;
  constexpr int TMEM_COLS = CI_MT * NACC;                // 256 or 512
  constexpr float LO_SCALE = 1.0f / 2048.0f;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t *stages = smem;
  uint8_t *sw = smem + (size_t)p.stages * STAGE_BYTES;
  float *ring = reinterpret_cast<float *>(sw + (size_t)p.ksteps * W_KSTEP_BYTES);
  uint64_t *full = reinterpret_cast<uint64_t *>(ring + 4 * CI_PLANE);
  uint64_t *empty = full + CI_MAX_STAGES;
  uint64_t *wbar = empty + CI_MAX_STAGES;
  uint64_t *tmem_full = wbar + 1;
  uint64_t *tmem_empty = tmem_full + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int band = blockIdx.x, nbands = gridDim.x, b = blockIdx.y;
  const int y0 = band * CI_ROWS;

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(wbar, 1);
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, CI_EPI_WARPS * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == CI_EPI_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == CI_EPI_WARPS + 1) {
    // ===================== producer =============================================================================
    if (lane == 0) {
      mbar_arrive_expect_tx(wbar, (uint32_t)(p.ksteps * W_KSTEP_BYTES));
      bulk_g2s(sw, p.wpack, (uint32_t)(p.ksteps * W_KSTEP_BYTES), wbar);
      int it = 0;
      for (int z = 0; z < p.D; ++z) {
        const size_t bd = (size_t)b * p.D + z;
        for (int ks = 0; ks < p.ksteps; ++ks, ++it) {
          const int s = it % p.stages, use = it / p.stages;
          if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
          uint8_t *sa = stages + (size_t)s * STAGE_BYTES;
          mbar_arrive_expect_tx(&full[s], (uint32_t)STAGE_BYTES);
#pragma unroll
          for (int part = 0; part < PARTS; ++part) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              int cg = ks * 2 + c;
              const __half *src = p.src0;
              int ncg = p.cg0;
              if (cg >= p.cg0) { cg -= p.cg0; src = p.src1; ncg = p.cg1; }
              const __half *g = src + ((((bd * PARTS + part) * ncg + cg) * p.H + y0) * (size_t)CI_W) * 8;
              bulk_g2s(sa + (part * 2 + c) * CG_BYTES, g, (uint32_t)CG_BYTES, &full[s]);
            }
          }
        }
      }
    }
  } else if (warp == CI_EPI_WARPS) {
    // ===================== MMA issuer ===========================================================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(128, NACC);
      constexpr uint32_t idesc_lo = umma_idesc_f16(128, 64);
      mbar_wait(wbar, 0);
      int it = 0;
      for (int z = 0; z < p.D; ++z) {
        if (z > 0) mbar_wait(tmem_empty, (z - 1) & 1);   // the epilogue has read step z-1's accumulators
        tc_fence_after();
        for (int ks = 0; ks < p.ksteps; ++ks, ++it) {
          const int s = it % p.stages, use = it / p.stages;
          mbar_wait(&full[s], use & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(stages + (size_t)s * STAGE_BYTES);
          const uint64_t bdesc = umma_desc(smem_u32(sw) + ks * W_KSTEP_BYTES, (NACC / 8) * 128, 128);
#pragma unroll
          for (int mt = 0; mt < CI_MT; ++mt) {
            const uint32_t a0 = sa + mt * 128 * 16;
            umma_f16(tmem_base + mt * NACC, umma_desc(a0, CG_BYTES, 128), bdesc, idesc, ks > 0);
            if constexpr (X2) umma_f16(tmem_base + mt * NACC + 64, umma_desc(a0 + 2 * CG_BYTES, CG_BYTES, 128), bdesc, idesc_lo, true);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(tmem_full);
      }
    }
  } else {
    // ===================== epilogue: TMEM -> scatter-add into the plane ring -> global ================================
    const int etid = tid;                                 // 0..511
    const int quarter = warp & 3, mt = warp >> 2;         // TMEM lane quarter of this warp; its M-tile
    const int row = quarter * 32 + lane;                  // accumulator row within the M-tile: input row (row >> 6), x = row & 63
    const int x = row & 63;
    const int yl = 2 * mt + (row >> 6);                   // input row of this thread within the band
    for (int i = etid; i < 4 * CI_PLANE; i += CI_EPI_WARPS * 32) ring[i] = 0.0f;
    named_bar_sync(1, CI_EPI_WARPS * 32);
    const float bias = __ldg(p.bias);
    const int Ho = 2 * p.H, Do = 2 * p.D;
    constexpr int Wo = 2 * CI_W;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    // ring cell of output column ox = 2x of this thread's two input rows (column ox + 4: even, so an (ox, ox + 1) pair is one float2)
    float *base = ring + (2 * yl) * CI_PITCH + 2 * x + 4;
    const int xhalf = quarter & 1, ysub = quarter >> 1;
    // taps kx = 3 of x = 31 and kx = 0 of x = 32 belong to cells of the neighbouring warp: exchanged through this scratch
    float *scr = reinterpret_cast<float *>(tmem_slot + 2);   // [2 buffers][4 M-tiles][2 rows][2 directions][4 (rz,ry)]

    auto flush_plane = [&](int slot, int oz) {   // complete plane -> global, ring slot back to zero
      float *pl = ring + slot * CI_PLANE;
      for (int i = etid; i < CI_UY * (Wo / 4); i += CI_EPI_WARPS * 32) {
        const int uy = i / (Wo / 4), q = i - uy * (Wo / 4);
        float4 *sp = reinterpret_cast<float4 *>(pl + uy * CI_PITCH + 4 + 4 * q);
        float4 v = *sp;
        *sp = make_float4(0.f, 0.f, 0.f, 0.f);
        const int oy = 2 * y0 - 1 + uy;
        if (oy < 0 || oy >= Ho) continue;
        if (uy >= 2 || band == 0) { v.x += bias; v.y += bias; v.z += bias; v.w += bias; }
        float *dst = p.out + (((size_t)b * Do + oz) * Ho + oy) * (size_t)Wo + 4 * q;
        const bool shared_row = (uy <= 1 && band > 0) || (uy >= CI_UY - 2 && band < nbands - 1);
        if (shared_row) red_add_v4(dst, v);
        else st_stream_f4(dst, v);
      }
    };

    for (int z = 0; z < p.D; ++z) {
      mbar_wait(tmem_full, z & 1);
      tc_fence_after();
      const int zs = (2 * z) & 3;   // ring slot of tap kz = 0 (output plane 2z - 1); tap kz uses (zs + kz) & 3
      // 4 phases (tz, ty): the 16 taps (rz, ry, kx = 0..3) of a position.  Along x the two contributions of an output are
      // summed in registers first (ox = 2x: kx = 1 of x and kx = 3 of x - 1; ox = 2x + 1: kx = 2 of x and kx = 0 of x + 1, via
      // warp shuffles), so a position adds ONE float2 per (kz, ky) to the ring: distinct cells for distinct positions.
      // The TMEM read of phase ph + 1 (hi half) is issued before the adds of phase ph: its latency hides behind the
      // shared-memory work; the 4 ring cells are loaded before any of them is touched (no load-add-store chains).
      float *pb[4];   // this position's cell row in the ring plane of tap kz
#pragma unroll
      for (int k = 0; k < 4; ++k) pb[k] = base + ((zs + k) & 3) * CI_PLANE;
      uint32_t rh[16], rl[16];
      tmem_ld16_nowait(trow + (uint32_t)(mt * NACC), rh);
#pragma unroll
      for (int ph = 0; ph < 4; ++ph) {
        const int tz = ph >> 1, ty = ph & 1;
        if constexpr (X2) tmem_ld16_nowait(trow + (uint32_t)(mt * NACC + 64 + ph * 16), rl);
        tmem_ld_wait();
        float v[16];   // index tx * 8 + rz * 4 + ry * 2 + rx, kx = 2 tx + rx
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = X2 ? fmaf(__uint_as_float(rl[i]), LO_SCALE, __uint_as_float(rh[i])) : __uint_as_float(rh[i]);
        if (ph < 3) {
          tmem_ld16_nowait(trow + (uint32_t)(mt * NACC + (ph + 1) * 16), rh);
        } else {        // the last TMEM read of this step has landed: the MMA warp may overwrite the accumulators
          tc_fence_before();
          mbar_arrive(tmem_empty);
        }
        float *sb = scr + (ph & 1) * 64 + (mt * 2 + ysub) * 8;
        if (lane == 31 && xhalf == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) sb[j] = v[8 + 2 * j + 1];       // kx = 3 of x = 31
        }
        if (lane == 0 && xhalf == 1) {
#pragma unroll
          for (int j = 0; j < 4; ++j) sb[4 + j] = v[2 * j];           // kx = 0 of x = 32
        }
        named_bar_sync(1, CI_EPI_WARPS * 32);   // scratch visible; the previous phase's adds are done
        float2 a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {           // j = rz * 2 + ry
          const int kz = 2 * tz + (j >> 1), ky = 2 * ty + (j & 1);
          a[j] = *reinterpret_cast<const float2 *>(pb[kz] + ky * CI_PITCH);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float up = __shfl_up_sync(0xffffffffu, v[8 + 2 * j + 1], 1), dn = __shfl_down_sync(0xffffffffu, v[2 * j], 1);
          if (lane == 0) up = xhalf ? sb[j] : 0.0f;          // x = 0: output -1 does not exist
          if (lane == 31) dn = xhalf ? 0.0f : sb[4 + j];     // x = 63: output 128 does not exist
          a[j].x += v[2 * j + 1] + up;      // ox = 2x:     kx = 1 here + kx = 3 of x - 1
          a[j].y += v[8 + 2 * j] + dn;      // ox = 2x + 1: kx = 2 here + kx = 0 of x + 1
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kz = 2 * tz + (j >> 1), ky = 2 * ty + (j & 1);
          *reinterpret_cast<float2 *>(pb[kz] + ky * CI_PITCH) = a[j];
        }
      }
      named_bar_sync(1, CI_EPI_WARPS * 32);     // all adds of this step are in the ring
      // planes 2z - 1 (tap kz = 0) and 2z (kz = 1) have received everything
      if (z > 0) flush_plane(zs, 2 * z - 1);
      else {      // output plane -1 does not exist: drop what tap kz = 0 of the first step scattered
        float *pl = ring + zs * CI_PLANE;
        for (int i = etid; i < CI_PLANE; i += CI_EPI_WARPS * 32) pl[i] = 0.0f;
      }
      flush_plane((zs + 1) & 3, 2 * z);
      named_bar_sync(1, CI_EPI_WARPS * 32);
    }
    // the last step's tap kz = 2 plane (output 2D - 1) has no later contributor; its kz = 3 plane (output 2D) is dropped
    flush_plane((2 * (p.D - 1) + 2) & 3, Do - 1);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == CI_EPI_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

template <int OP>
static int launch_col2im(ColParams &p, cudaStream_t st) {
  constexpr int PARTS = OP == 2 ? 2 : 1, NACC = PARTS * 64;
  constexpr int STAGE_BYTES = PARTS * 2 * CI_POS * 16;
  const int fixed = p.ksteps * 2 * (NACC / 8) * 128 + 4 * CI_PLANE * 4 + 1024;   // weights + plane ring + barriers, TMEM slot, exchange scratch
  int stages = (226 * 1024 - fixed) / STAGE_BYTES;
  if (stages > CI_MAX_STAGES) stages = CI_MAX_STAGES;
  GB_REQUIRE(stages >= 2, GENRE_B200_EINVAL, "convt_c1_col2im: %d K steps of weights leave no room for the pipeline", p.ksteps);
  p.stages = stages;
  const size_t smem = (size_t)stages * STAGE_BYTES + fixed;
  auto kern = convt_c1_col2im_kernel<OP>;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return check_launch("convt_c1_col2im: cudaFuncSetAttribute");
    configured[dev & 63] = true;
  }
  dim3 grid((unsigned)(p.H / CI_ROWS), (unsigned)p.B);
  kern<<<grid, CI_THREADS, smem, st>>>(p);
  return check_launch("convt_c1_col2im kernel");
}

}  // namespace gb

using namespace gb;

// ConvTranspose3d(8*(cg0+cg1) -> 1, kernel 4, stride 2, padding 1) on channel-blocked fp16 operands of a [B, C, D, H, 64] volume.
//   src0 [B*D][parts*cg0][H][64][8 fp16], src1 likewise with cg1 groups or NULL (the second half of a skip concatenation);
//   parts = 1 (op 1: fp16 operands) or 2 (op 2: hi | lo' = (a - hi) * 2^11 parts: fp32-accurate); cg0 + cg1 even; H % 8 == 0;
//   wpack [(cg0+cg1)/2][2][parts*8][8][8] fp16: row n = t*8 + r holds tap k = 2t + r per dimension (ops_conv.pack_convt_c1_col2im_weights);
//   bias: 1 float on the device;  out [B][2D][2H][128] fp32, fully overwritten.
extern "C" int genre_b200_convt_c1_col2im_forward(const void *src0, int cg0, const void *src1, int cg1, int64_t B, int64_t D, int64_t H,
                                                  int64_t W, const void *wpack, int op, const float *bias, float *out, void *stream) {
  GB_REQUIRE(src0 && wpack && bias && out, GENRE_B200_EINVAL, "convt_c1_col2im: null pointer");
  GB_REQUIRE(cg0 > 0 && cg1 >= 0 && (cg1 == 0 || src1) && (cg0 + cg1) % 2 == 0, GENRE_B200_EINVAL, "convt_c1_col2im: bad channel groups");
  GB_REQUIRE(B > 0 && B <= 65535 && D > 0 && H > 0 && H % CI_ROWS == 0 && W == CI_W, GENRE_B200_EINVAL,
             "convt_c1_col2im: extent %lldx%lldx%lld unsupported (W = 64, H %% 8 == 0)", (long long)D, (long long)H, (long long)W);
  GB_REQUIRE(op == 1 || op == 2, GENRE_B200_EINVAL, "convt_c1_col2im: op %d (1 = fp16, 2 = fp16 hi/lo)", op);
  GB_REQUIRE(aligned16(src0) && aligned16(wpack) && aligned16(out) && (!src1 || aligned16(src1)), GENRE_B200_EALIGN, "convt_c1_col2im: alignment");
  ColParams p{};
  p.src0 = (const __half *)src0; p.src1 = (const __half *)src1;
  p.cg0 = cg0; p.cg1 = cg1;
  p.B = (int)B; p.D = (int)D; p.H = (int)H;
  p.wpack = (const __half *)wpack;
  p.ksteps = (cg0 + cg1) / 2;
  p.bias = bias;
  p.out = out;
  cudaStream_t st = as_stream(stream);
  // zero the output rows two bands share (rows 16 j - 1, 16 j of every plane: pairs of adjacent rows at a pitch of 16 rows)
  const int64_t Wo = 2 * W, Ho = 2 * H, Do = 2 * D, nbands = H / CI_ROWS;
  const int64_t pairs = B * Do * nbands - 1;
  if (nbands > 1 && pairs > 0) {
    if (cudaMemset2DAsync(out + 15 * Wo, (size_t)16 * Wo * sizeof(float), 0, (size_t)2 * Wo * sizeof(float), (size_t)pairs, st) != cudaSuccess)
      return check_launch("convt_c1_col2im: memset");
  }
  (void)Ho;
  return op == 2 ? launch_col2im<2>(p, st) : launch_col2im<1>(p, st);
}
