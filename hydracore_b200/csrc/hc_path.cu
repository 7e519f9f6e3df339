// hc_path.cu — path tracing entry points (placeholder until the shading kernels land)
#include "hc_context.h"
void hc_path_free(hc_ctx*) {}
extern "C" {
int hc_resize(hc_ctx* ctx, int width, int height) { if (!ctx || width <= 0 || height <= 0) return HC_E_ARG; ctx->width = width; ctx->height = height; return HC_OK; }
int hc_pt_init(hc_ctx*, int) { hc_set_error("not implemented"); return HC_E_STATE; }
int hc_pt_set_tiles(hc_ctx*, int, int, int) { hc_set_error("not implemented"); return HC_E_STATE; }
int hc_pt_pass(hc_ctx*, int, int) { hc_set_error("not implemented"); return HC_E_STATE; }
int hc_fb_clear(hc_ctx*) { hc_set_error("not implemented"); return HC_E_STATE; }
int hc_fb_device_ptr(hc_ctx*, float**, int64_t*) { hc_set_error("not implemented"); return HC_E_STATE; }
int hc_fb_read_hdr(hc_ctx*, float*, int, int) { hc_set_error("not implemented"); return HC_E_STATE; }
int hc_fb_read_ldr(hc_ctx*, uint32_t*, int, int) { hc_set_error("not implemented"); return HC_E_STATE; }
int hc_get_spp(hc_ctx*, float*) { hc_set_error("not implemented"); return HC_E_STATE; }
}
