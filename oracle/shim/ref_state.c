/* TEST INFRASTRUCTURE: the global the reference's C shims expect (`extern THCState *state;`, my_lib.c:4). */
#include "THC.h"
static THCState ref_state_storage = {0};
THCState *state = &ref_state_storage;
