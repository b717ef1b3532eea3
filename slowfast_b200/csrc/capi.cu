// Small C-ABI entry points that do not belong to a kernel file.
#include "../../include/slowfast_b200.h"
#include "tmap.h"

#ifndef SFB_BUILD_ARCH
#define SFB_BUILD_ARCH "unknown"
#endif

extern "C" const char* sfb_last_error(void) { return sfb::last_error(); }
extern "C" int sfb_abi_version(void) { return 1; }
extern "C" const char* sfb_build_arch(void) { return SFB_BUILD_ARCH; }
