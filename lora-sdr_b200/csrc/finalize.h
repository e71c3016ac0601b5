// finalize.h -- host-side completion of raw detector outputs (see finalize.cpp).
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/lorab200.h"

namespace lorab200 {

// Host mirror of the device RawResult record (common.cuh), 32 bytes.
struct alignas(16) RawResultHost {
    int32_t index;
    float   max_mag2;
    double  total;
    float   left_re, left_im;
    float   right_re, right_im;
};
static_assert(sizeof(RawResultHost) == 32, "raw record layout");

void finalize_host(const RawResultHost *raw, size_t n, float power_scale, lorab200_result *out);

} // namespace lorab200
