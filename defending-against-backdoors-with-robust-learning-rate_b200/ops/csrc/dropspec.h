// Parameters of a fused dropout (plain struct: shared by host-only bindings and device code; the mask function lives in common.cuh).
#pragma once
#include <stdint.h>

namespace rlr {

struct DropSpec {
    uint32_t thr;               // p * 65536, 0 = dropout off
    float scale;                // 1 / (1 - p)
    uint64_t seed, stream;      // stream = node id of the dropout layer
    const long long* step;      // device step counter (advanced once per local step by advance_cursor)
};

}  // namespace rlr
