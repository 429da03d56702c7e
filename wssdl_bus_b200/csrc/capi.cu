// Version / error-string entry points of libwssdl_b200.so (include/wssdl_b200.h).
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

// ---- process-wide tuning switches: defaults read from the environment ONCE, at the first call
// that needs one; wssdl_set_tuning changes them afterwards.  No entry point calls getenv() per
// launch.
namespace {
int g_tuning[WSSDL_TUNE_COUNT];
std::once_flag g_tuning_once;

int env_choice(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

void init_tuning() {
  for (int i = 0; i < WSSDL_TUNE_COUNT; ++i) g_tuning[i] = 0;
  const char* k = getenv("WSSDL_ROI_FWD_KERNEL");     // direct | tiled | band | sorted
  if (k && *k) {
    g_tuning[WSSDL_TUNE_ROI_FWD_KERNEL] =
        !strcmp(k, "direct") ? 1 : !strcmp(k, "tiled") ? 2 : !strcmp(k, "band") ? 3
        : !strcmp(k, "sorted") ? 4 : 0;
  }
  g_tuning[WSSDL_TUNE_ROI_FWD_SLICES] = env_choice("WSSDL_ROI_FWD_SLICES", 0);
  g_tuning[WSSDL_TUNE_ROI_FWD_CHUNKS] = env_choice("WSSDL_ROI_FWD_CHUNKS", 0);
  g_tuning[WSSDL_TUNE_NMS_SWEEP_CLUSTER] = env_choice("WSSDL_NMS_SWEEP_CLUSTER", -1);
  g_tuning[WSSDL_TUNE_PROPOSALS_CLUSTER] = env_choice("WSSDL_PROPOSALS_CLUSTER", -1);
  g_tuning[WSSDL_TUNE_ROI_FWD_THREADS] = env_choice("WSSDL_ROI_FWD_THREADS", 0);
  g_tuning[WSSDL_TUNE_PDL] = env_choice("WSSDL_PDL", 1);
  g_tuning[WSSDL_TUNE_ROI_FWD_BALANCED] = env_choice("WSSDL_ROI_FWD_BALANCED", -1);
}
}  // namespace

int wssdl_tuning(int key) {
  std::call_once(g_tuning_once, init_tuning);
  return (key >= 0 && key < WSSDL_TUNE_COUNT) ? g_tuning[key] : 0;
}

extern "C" int wssdl_get_tuning(int key) { return wssdl_tuning(key); }

extern "C" int wssdl_set_tuning(int key, int value) {
  std::call_once(g_tuning_once, init_tuning);
  if (key < 0 || key >= WSSDL_TUNE_COUNT) return WSSDL_EINVAL;
  __atomic_store_n(&g_tuning[key], value, __ATOMIC_RELAXED);
  return WSSDL_OK;
}

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
