/* C consumer of the qsb200 C ABI (include/qsb200.h): what a cgo / JNI / ctypes binding relies on, checked from plain
 * C without a GPU -- the library links, every descriptor is a POD, invalid descriptors come back as status codes
 * (never a crash or an exception), the flop and workspace queries are pure host code. */
#include <stdio.h>
#include <string.h>
#include "qsb200.h"

#define CHECK(cond) do { if (!(cond)) { fprintf(stderr, "abi_check failed: %s (line %d)\n", #cond, __LINE__); return 1; } } while (0)

int main(void) {
    CHECK(qsb_version() >= 200);
    /* the structs this translation unit sees are the structs the library was built with */
    CHECK(qsb_sizeof_heff_desc() == sizeof(qsb_heff_desc));
    CHECK(qsb_sizeof_env_desc() == sizeof(qsb_env_desc));
    CHECK(qsb_sizeof_gemm_desc() == sizeof(qsb_gemm_desc));
    CHECK(strcmp(qsb_status_string(QSB_OK), "ok") == 0);
    CHECK(strlen(qsb_status_string(QSB_ERR_WORKSPACE)) > 0);

    qsb_heff_desc d;
    memset(&d, 0, sizeof d);
    size_t bytes = 0;
    CHECK(qsb_heff_workspace_bytes(&d, &bytes) == QSB_ERR_SHAPE);       /* zero extents */
    CHECK(qsb_heff_workspace_bytes(NULL, &bytes) == QSB_ERR_ARG);
    d.dtype = QSB_F64;
    d.chi_l_out = d.chi_l_in = d.chi_r_out = d.chi_r_in = 2048;
    d.d1_out = d.d1_in = d.d2_out = d.d2_in = 2;
    d.W_l = d.W_m = d.W_r = 5;
    d.L_stride[0] = 2048LL * 2048; d.L_stride[1] = 2048; d.L_stride[2] = 1;
    d.R_stride[0] = 2048LL * 2048; d.R_stride[1] = 2048; d.R_stride[2] = 1;
    CHECK(qsb_heff_workspace_bytes(&d, &bytes) == QSB_OK);
    CHECK(bytes >= 2ull * 5 * 4 * 2048 * 2048 * 8);                     /* T1 + T3 at BASELINE config 3 */
    double flops = qsb_heff_flops(&d);
    CHECK(flops > 6.90e11 && flops < 6.92e11);                          /* 6.906e11 */
    CHECK(qsb_heff_flops_executed(&d) == flops);                        /* no identity channel declared */
    d.L_identity = 5; d.R_identity = 1;                                 /* 1 + channel: L[4], R[0] (FSM layout) */
    double ex = qsb_heff_flops_executed(&d);
    CHECK(flops - ex == 2.0 * 2.0 * 4.0 * 2048.0 * 2048.0 * 2048.0);    /* one chi^3 d^2 GEMM slice per side */
    d.L_identity = d.R_identity = 0;
    CHECK(qsb_heff_apply(&d, NULL) == QSB_ERR_ARG);                     /* null data pointers, no launch */
    CHECK(qsb_env_identity_deviation(QSB_F64, 5, 8, 8, NULL, NULL, 0, NULL, NULL) == QSB_ERR_ARG);
    d.dtype = QSB_F32;
    CHECK(qsb_heff_apply(&d, NULL) == QSB_ERR_DTYPE);

    qsb_env_desc e;
    memset(&e, 0, sizeof e);
    CHECK(qsb_env_workspace_bytes(&e, &bytes) == QSB_ERR_SHAPE);
    e.dtype = QSB_C128; e.W_in = e.W_out = 5; e.d = 2; e.chi_in = e.chi_out = 64;
    e.env_stride[0] = 64 * 64; e.env_stride[1] = 64; e.env_stride[2] = 1;
    e.site_stride[0] = 128; e.site_stride[1] = 64; e.site_stride[2] = 1;
    CHECK(qsb_env_workspace_bytes(&e, &bytes) == QSB_OK && bytes > 0);
    CHECK(qsb_env_flops(&e) == 8.0 * (5.0 * 2 * 64 * 64 * 64 + 25.0 * 4 * 64 * 64 + 5.0 * 2 * 64 * 64 * 64));
    CHECK(qsb_env_left(&e, NULL) == QSB_ERR_ARG);

    CHECK(qsb_lz_scratch_bytes(8) >= 1024);
    CHECK(qsb_lz_dot(99, 16, 1, &d, 0, &d, (double*)&d, 0, 0, &d, NULL) == QSB_ERR_DTYPE);
    CHECK(qsb_lz_scale(QSB_F64, -1, &d, &d, (const double*)&d, 0, NULL) == QSB_ERR_ARG);
    CHECK(qsb_push_shard(NULL, 16, NULL, NULL, 2, 1, NULL, NULL) == QSB_ERR_ARG);
    uint64_t counters[3] = {7, 7, 7};
    qsb_launch_counters(counters, 1);
    CHECK(counters[0] + counters[1] + counters[2] == 0);                /* nothing was launched */
    printf("abi_check ok\n");
    return 0;
}
