#include "engine.h"
extern "C" {
int mb200_consistency_iter(mb200_ctx *ctx, uint32_t, uint32_t) { return mb_fail(ctx, MB200_EINVAL, "mb200_consistency_iter: not implemented in this build"); }
}
