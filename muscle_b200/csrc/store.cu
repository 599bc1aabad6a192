// store.cu -- packed images of the sparse store (export, multi-GPU exchange).  (stubs for now)
#include "engine.h"
extern "C" {
#define NOTYET(name) return mb_fail(ctx, MB200_EINVAL, name ": not implemented in this build")
int mb200_export_all(mb200_ctx *ctx, uint32_t *, mb200_entry *) { NOTYET("mb200_export_all"); }
int mb200_store_pack(mb200_ctx *ctx, const uint32_t **, uint64_t *, const mb200_entry **, uint64_t *) { NOTYET("mb200_store_pack"); }
int mb200_store_load_allpairs(mb200_ctx *ctx, uint32_t, uint32_t, const uint32_t *, uint64_t, const mb200_entry *, uint64_t) { NOTYET("mb200_store_load_allpairs"); }
int mb200_store_values(mb200_ctx *ctx, float *, uint64_t) { NOTYET("mb200_store_values"); }
int mb200_store_set_values(mb200_ctx *ctx, const float *, uint64_t, uint64_t) { NOTYET("mb200_store_set_values"); }
}
