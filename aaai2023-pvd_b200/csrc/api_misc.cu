// api_misc.cu -- library identification and diagnostics of the C ABI (include/pvd_b200.h).
#include "common.cuh"

extern "C" {

int pvd_abi_version(void) { return 1; }

const char* pvd_error_string(int code) {
    if (code == PVD_OK) return "success";
    if (code == PVD_EINVAL) return "pvd: invalid argument (null pointer, bad size or misaligned buffer)";
    if (code == PVD_EUNSUPPORTED) return "pvd: unsupported configuration (e.g. level_dim not in {1,2,4,8}, input_dim not in {2,3})";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "pvd: unknown error";
}

int pvd_device_sm_count(int* out_sms) {
    if (!out_sms) return PVD_EINVAL;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(out_sms, cudaDevAttrMultiProcessorCount, dev);
    return (int)e;
}

}  // extern "C"
