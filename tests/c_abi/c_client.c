/* A plain C99 client of include/sdnq_b200.h: proves the header is valid C (not only C++), that a program with no torch / no
 * CUDA headers links against libsdnq_b200.so, and that the library answers argument errors with status codes.  Built and run by
 * tests/test_c_abi.py::test_plain_c_client_links_and_runs (no GPU needed: every call below returns before touching CUDA). */
#include <stdio.h>
#include <string.h>

#include "sdnq_b200.h"

int main(void) {
    int failures = 0;
    sdnq_weight_format wide = {SDNQ_W_INT, 12, 0, 0, 0, 1};
    sdnq_conv2d_geometry geo;
    char dummy[64];
    int rc;

    if (sdnq_b200_abi_version() != SDNQ_B200_ABI_VERSION) { printf("abi version mismatch\n"); ++failures; }
    if (sdnq_b200_linear_w8a8_workspace_bytes(0, 640) != 0) { printf("empty workspace not 0\n"); ++failures; }
    if (sdnq_b200_linear_w8a8_workspace_bytes(4096, 640) < (size_t)4096 * 640) { printf("workspace too small\n"); ++failures; }

    rc = sdnq_b200_unpack(dummy, &wide, dummy, SDNQ_I8, 8, NULL);
    if (rc != SDNQ_EUNSUPPORTED || strstr(sdnq_b200_last_error(), "8 bits") == NULL) { printf("unpack: rc=%d msg=%s\n", rc, sdnq_b200_last_error()); ++failures; }

    rc = sdnq_b200_scaled_mm(dummy, dummy, SDNQ_I8, NULL, NULL, NULL, SDNQ_F32, 0, NULL, NULL, NULL, NULL, dummy, SDNQ_BF16, 64, 64, 24, NULL);
    if (rc != SDNQ_EUNSUPPORTED) { printf("scaled_mm: rc=%d msg=%s\n", rc, sdnq_b200_last_error()); ++failures; }

    memset(&geo, 0, sizeof geo);
    geo.batch = 1; geo.channels = 64; geo.height = 8; geo.width = 8;
    geo.x_stride_b = 4096; geo.x_stride_c = 64; geo.x_stride_h = 8; geo.x_stride_w = 1;
    geo.kernel_h = 3; geo.kernel_w = 3; geo.stride_h = 0; geo.stride_w = 1; geo.pad_h = 1; geo.pad_w = 1; geo.dilation_h = 1; geo.dilation_w = 1;
    rc = sdnq_b200_conv_act_quant(dummy, SDNQ_BF16, &geo, 0, SDNQ_I8, dummy, (float*)dummy, NULL, NULL, NULL, NULL);
    if (rc != SDNQ_EINVAL) { printf("conv_act_quant: rc=%d msg=%s\n", rc, sdnq_b200_last_error()); ++failures; }

    printf("c_client: %d failure(s)\n", failures);
    return failures;
}
