// Version / error-string entry points of libwssdl_b200.so (include/wssdl_b200.h).
#include "common.cuh"

extern "C" int wssdl_version(void) { return 100; }  // 1.00

extern "C" const char* wssdl_error_string(int code) {
  switch (code) {
    case WSSDL_OK: return "ok";
    case WSSDL_EINVAL: return "invalid argument";
    case WSSDL_EWORKSPACE: return "workspace too small";
    case WSSDL_EALIGN: return "pointer alignment";
    case WSSDL_ELIMIT: return "size beyond kernel limits";
    case WSSDL_EZERODIV: return "float division (a visited box pair has zero union)";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown error";
}
