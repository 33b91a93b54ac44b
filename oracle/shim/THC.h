/*
 * THC.h stand-in (TEST INFRASTRUCTURE, not product code).
 *
 * The reference's CUDA sources (toolbox/cam_bp/cam_bp/src/back_projection_kernel.cu,
 * toolbox/calc_prob/calc_prob/src/calc_prob_kernel.cu) include <THC.h> from PyTorch 0.4,
 * which no longer exists.  This header declares the handful of THC symbols those two files
 * touch so that they compile UNMODIFIED, from where they lie under /root/reference, into
 * oracle/_ref/ (see oracle/Makefile).  The tensor is a plain {data, ndim, size[], stride[]}
 * view that tests fill in through ctypes; nothing here allocates device memory.
 */
#ifndef GENRE_B200_ORACLE_THC_SHIM_H
#define GENRE_B200_ORACLE_THC_SHIM_H

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>

#define REF_MAX_DIM 8

typedef struct THCState {
  cudaStream_t stream;
} THCState;

typedef struct THCudaTensor {
  float *data;
  int ndim;
  long size[REF_MAX_DIM];
  long stride[REF_MAX_DIM];
} THCudaTensor;

typedef struct THCDescBuff {
  char str[64];
} THCDescBuff;

#ifdef __cplusplus
#define REF_INLINE static inline
#else
#define REF_INLINE static inline
#endif

REF_INLINE int THCudaTensor_nDimension(THCState *s, const THCudaTensor *t) { (void)s; return t->ndim; }
REF_INLINE long THCudaTensor_size(THCState *s, const THCudaTensor *t, int d) { (void)s; return t->size[d]; }
REF_INLINE long THCudaTensor_stride(THCState *s, const THCudaTensor *t, int d) { (void)s; return t->stride[d]; }
REF_INLINE float *THCudaTensor_data(THCState *s, const THCudaTensor *t) { (void)s; return t->data; }
REF_INLINE cudaStream_t THCState_getCurrentStream(THCState *s) { return s->stream; }

REF_INLINE long ref_numel_contig_check(const THCudaTensor *t) {
  long n = 1;
  for (int i = 0; i < t->ndim; ++i) n *= t->size[i];
  return n;
}

/* resizeNd in the reference is always a no-op resize to the shape the caller already allocated;
 * the shim verifies that and aborts otherwise (it never reallocates). */
REF_INLINE void ref_resize_check(THCudaTensor *t, int nd, const long *sz) {
  if (t->ndim != nd) { fprintf(stderr, "THC shim: resize would change ndim\n"); abort(); }
  for (int i = 0; i < nd; ++i)
    if (t->size[i] != sz[i]) { fprintf(stderr, "THC shim: resize would change shape\n"); abort(); }
}
REF_INLINE void THCudaTensor_resize2d(THCState *s, THCudaTensor *t, long a, long b) {
  (void)s; long z[2] = {a, b}; ref_resize_check(t, 2, z);
}
REF_INLINE void THCudaTensor_resize4d(THCState *s, THCudaTensor *t, long a, long b, long c, long d) {
  (void)s; long z[4] = {a, b, c, d}; ref_resize_check(t, 4, z);
}
REF_INLINE void THCudaTensor_resize5d(THCState *s, THCudaTensor *t, long a, long b, long c, long d, long e) {
  (void)s; long z[5] = {a, b, c, d, e}; ref_resize_check(t, 5, z);
}

/* zero / fill: only ever applied to dense contiguous outputs by the reference wrappers. */
void ref_shim_fill(THCState *s, THCudaTensor *t, float v);
REF_INLINE void THCudaTensor_zero(THCState *s, THCudaTensor *t) {
  cudaMemsetAsync(t->data, 0, sizeof(float) * (size_t)ref_numel_contig_check(t), s->stream);
}
REF_INLINE void THCudaTensor_fill(THCState *s, THCudaTensor *t, float v) { ref_shim_fill(s, t, v); }

REF_INLINE int THCudaTensor_checkGPU(THCState *s, unsigned int n, ...) { (void)s; (void)n; return 1; }

REF_INLINE THCDescBuff THCudaTensor_sizeDesc(THCState *s, const THCudaTensor *t) {
  (void)s;
  THCDescBuff b; b.str[0] = 0; char *p = b.str;
  for (int i = 0; i < t->ndim && (p - b.str) < 50; ++i) p += sprintf(p, i ? "x%ld" : "%ld", t->size[i]);
  return b;
}

REF_INLINE void THError(const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); abort();
}
REF_INLINE void THArgCheck(int cond, int argn, const char *fmt, ...) {
  if (cond) return;
  va_list ap; va_start(ap, fmt); fprintf(stderr, "arg %d: ", argn); vfprintf(stderr, fmt, ap); va_end(ap);
  fprintf(stderr, "\n"); abort();
}
#define THAssertMsg(cond, ...) do { if (!(cond)) { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); abort(); } } while (0)

#endif
