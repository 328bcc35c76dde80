// large.cu — single large systems (placeholder until the sparse path lands in this round).
#include "device.h"

namespace ezs {

void release_large(DeviceCopy* d) { (void)d; }

int32_t solve_large(ezpz_context* ctx, const ezpz_structure* s, const ezpz_config_t* config, const ezpz_one_io_t* io,
                    ezpz_error_detail_t* detail) {
    (void)ctx; (void)s; (void)config; (void)io;
    if (detail) std::snprintf(detail->message, sizeof detail->message, "large-system path not built yet");
    return EZPZ_ERR_TOO_LARGE;
}

}  // namespace ezs
