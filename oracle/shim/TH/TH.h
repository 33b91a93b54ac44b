/* TH.h stand-in (TEST INFRASTRUCTURE): just enough for toolbox/nndistance/src/my_lib.c to compile unmodified. */
#ifndef GENRE_B200_ORACLE_TH_SHIM_H
#define GENRE_B200_ORACLE_TH_SHIM_H
#include "../THC.h"
typedef THCudaTensor THFloatTensor; /* my_lib.c passes THFloatTensor* to THCudaTensor_size */
typedef struct THIntTensor {
  int *data;
  int ndim;
  long size[REF_MAX_DIM];
  long stride[REF_MAX_DIM];
} THIntTensor;
static inline float *THFloatTensor_data(const THFloatTensor *t) { return t->data; }
static inline int *THIntTensor_data(const THIntTensor *t) { return t->data; }
#endif
