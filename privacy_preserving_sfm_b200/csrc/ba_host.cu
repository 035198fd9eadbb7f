// ba_host.cu — line-reprojection bundle adjustment (host side). Filled in below.
#include "common.h"

extern "C" void ppsfm_ba_state_free(ppsfm_ctx* ctx) { (void)ctx; }
