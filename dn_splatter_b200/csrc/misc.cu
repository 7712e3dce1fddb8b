// Version / error strings of the C ABI.
#include "common.cuh"

extern "C" int dnr_version(void) { return DNR_VERSION; }

extern "C" const char* dnr_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case DNR_E_NULL: return "DNR_E_NULL: a required pointer is NULL";
    case DNR_E_SIZE: return "DNR_E_SIZE: non-positive or inconsistent sizes";
    case DNR_E_OPTION: return "DNR_E_OPTION: unsupported option";
    case DNR_E_OVERFLOW: return "DNR_E_OVERFLOW: more than 2^31-1 tile intersections";
    case DNR_E_WORKSPACE: return "DNR_E_WORKSPACE: workspace too small";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown dnr error";
}
