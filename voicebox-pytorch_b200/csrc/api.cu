// Library-level entry points of libvbx_sm100a.so (version, error strings).
#include "common.cuh"

extern "C" int vbx_version(void) { return VBX_VERSION; }

extern "C" const char* vbx_strerror(int rc) {
  switch (rc) {
    case VBX_OK: return "ok";
    case VBX_E_NULL: return "vbx: required pointer is NULL";
    case VBX_E_SHAPE: return "vbx: invalid size (non-positive or violates a divisibility rule)";
    case VBX_E_ALIGN: return "vbx: pointer or stride is not 16-byte aligned";
    case VBX_E_UNSUPPORTED: return "vbx: configuration not supported by this kernel";
    case VBX_E_DRIVER: return "vbx: CUDA driver entry point cuTensorMapEncodeTiled unavailable or failed";
    default: break;
  }
  if (rc > 0) return cudaGetErrorString((cudaError_t)rc);
  return "vbx: unknown error";
}
