// C-ABI odds and ends of libcetpick_sm100a.so (include/cetpick.h).
#include "common.cuh"

namespace cetpick {
thread_local int64_t g_launches = 0;
thread_local std::string g_cuda_err;
}  // namespace cetpick

extern "C" int cetpick_version(void) { return CETPICK_ABI_VERSION; }

extern "C" const char* cetpick_strerror(int code) {
  switch (code) {
    case CETPICK_OK: return "ok";
    case CETPICK_ERR_BAD_ARG: return "bad argument";
    case CETPICK_ERR_UNSUPPORTED: return "unsupported configuration";
    case CETPICK_ERR_WORKSPACE: return "workspace missing, too small or misaligned";
    case CETPICK_ERR_CUDA: return "CUDA error";
    case CETPICK_ERR_STATE: return "plan not ready (missing parameter or not finalized)";
    case CETPICK_ERR_SHAPE: return "parameter has the wrong number of elements";
  }
  return "unknown error";
}

extern "C" const char* cetpick_last_cuda_error(void) { return cetpick::g_cuda_err.c_str(); }

extern "C" int64_t cetpick_last_launch_count(void) { return cetpick::g_launches; }
