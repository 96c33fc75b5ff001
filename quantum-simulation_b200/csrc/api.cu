// Library info, status strings and launch counters of the qsb200 C ABI (include/qsb200.h).
#include "common.cuh"

namespace qsb {
uint64_t g_count_tma = 0, g_count_cpasync = 0, g_count_other = 0;
int g_last_cuda_error = 0;
}  // namespace qsb

extern "C" int qsb_version(void) { return 200; }

extern "C" const char* qsb_status_string(int s) {
    switch (s) {
        case QSB_OK: return "ok";
        case QSB_ERR_SHAPE: return "inconsistent tensor extents";
        case QSB_ERR_DTYPE: return "unsupported dtype";
        case QSB_ERR_STRIDE: return "unsupported strides or alignment";
        case QSB_ERR_DEVICE: return "no usable sm_100 CUDA device";
        case QSB_ERR_WORKSPACE: return "workspace missing or too small";
        case QSB_ERR_LAUNCH: return "CUDA launch/runtime failure";
        case QSB_ERR_ARG: return "invalid argument";
        default: return "unknown status";
    }
}

extern "C" int qsb_last_cuda_error(void) { return qsb::g_last_cuda_error; }

extern "C" void qsb_launch_counters(uint64_t out[3], int reset) {
    if (out) { out[0] = qsb::g_count_tma; out[1] = qsb::g_count_cpasync; out[2] = qsb::g_count_other; }
    if (reset) { qsb::g_count_tma = qsb::g_count_cpasync = qsb::g_count_other = 0; }
}

extern "C" size_t qsb_sizeof_heff_desc(void) { return sizeof(qsb_heff_desc); }
extern "C" size_t qsb_sizeof_env_desc(void) { return sizeof(qsb_env_desc); }
extern "C" size_t qsb_sizeof_gemm_desc(void) { return sizeof(qsb_gemm_desc); }
