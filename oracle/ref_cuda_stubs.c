/* oracle/ref_cuda_stubs.c -- TEST INFRASTRUCTURE ONLY.
 * Aborting stand-ins for the nine libgpuscale symbols that the reference's
 * libswscale.so leaves undefined (swscale_internal.h:704,973-1009,
 * swscale_unscaled.c:1970-1991, rgb2rgb.h:175), so that the reference's CPU
 * libswscale can be loaded on its own for the CPU timing baseline.  The
 * product library exports the real ones (include/gmat_b200_sws.h). */
#include <stdio.h>
#include <stdlib.h>
#define STUB(name) void name(void) { fprintf(stderr, "oracle stub " #name " called\n"); abort(); }
STUB(ff_sws_init_swscale_cuda)
STUB(ff_swscale_cuda)
STUB(ff_sws_free_swscale_cuda)
STUB(ff_yuv2rgb_init_tables_cuda)
STUB(yuv2rgb_cuda)
STUB(rgb2yuv_cuda)
STUB(yuv2yuv_cuda)
STUB(rgb24tobgr24_cuda)
void rgb2rgb_init_cuda(void) {}
