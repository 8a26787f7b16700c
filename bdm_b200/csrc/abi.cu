// abi.cu -- version and error text of the C-ABI.
#include "common.cuh"

extern "C" int bdm_abi_version(void) { return BDM_ABI_VERSION; }

extern "C" const char *bdm_error_string(int code) {
  switch (code) {
    case BDM_OK: return "success";
    case BDM_ERR_NULL_POINTER: return "bdm_b200: required pointer argument is NULL";
    case BDM_ERR_BAD_SIZE: return "bdm_b200: size argument out of range";
    case BDM_ERR_WORKSPACE_TOO_SMALL: return "bdm_b200: workspace smaller than *_workspace_bytes()";
    case BDM_ERR_MISALIGNED: return "bdm_b200: workspace must be 16-byte aligned";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "bdm_b200: unknown error code";
}
