#include "engine.h"
extern "C" {
int mb200_align_pairs(mb200_ctx *ctx, uint32_t, const uint32_t *, char *, const uint64_t *, float *) { return mb_fail(ctx, MB200_EINVAL, "mb200_align_pairs: not implemented in this build"); }
int mb200_align_groups(mb200_ctx *ctx, uint32_t, const uint32_t *, const uint32_t *, uint32_t, uint32_t, const uint32_t *, const uint32_t *, uint32_t, char *, float *, float *) { return mb_fail(ctx, MB200_EINVAL, "mb200_align_groups: not implemented in this build"); }
}
