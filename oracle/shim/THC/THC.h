/* <THC/THC.h> stand-in (TEST INFRASTRUCTURE): forwards to the flat shim. */
#include "../THC.h"
