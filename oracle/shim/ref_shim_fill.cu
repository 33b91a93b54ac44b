/* Test infrastructure: dense fill used by the THC shim (reference get_surface_mask_wrap fills mask with 1). */
#include "THC.h"
__global__ void ref_shim_fill_kernel(float *p, float v, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)blockDim.x * gridDim.x) p[i] = v;
}
void ref_shim_fill(THCState *s, THCudaTensor *t, float v) {
  long n = ref_numel_contig_check(t);
  ref_shim_fill_kernel<<<1024, 256, 0, s->stream>>>(t->data, v, n);
}
