// Losses, metrics and optimiser kernels (placeholder bodies are filled in below as they land).
#include "../../include/rsb200.h"
#include "rsb_host.h"
using namespace rsb;
extern "C" int rsb_cross_entropy(const float*, const int64_t*, const float*, float*, float*, double*, int32_t, int32_t, int32_t, void*) { return set_error(RSB_E_INVALID, "rsb_cross_entropy: not built yet"); }
extern "C" int64_t rsb_lovasz_workspace_bytes(int32_t, int32_t, int32_t) { return 0; }
extern "C" int rsb_lovasz(const float*, const int64_t*, float*, float*, void*, int64_t, int32_t, int32_t, int32_t, void*) { return set_error(RSB_E_INVALID, "rsb_lovasz: not built yet"); }
extern "C" int rsb_metrics_count(const float*, const int64_t*, int64_t*, int32_t, int32_t, int32_t, void*) { return set_error(RSB_E_INVALID, "rsb_metrics_count: not built yet"); }
extern "C" int rsb_adam_step(float*, const float*, float*, float*, int64_t, float, float, float, float, int32_t, void*) { return set_error(RSB_E_INVALID, "rsb_adam_step: not built yet"); }
