// Tensor-core (tcgen05, 3xTF32) variant of the blend-shape contraction -- placeholder
// until the UMMA kernel lands; the FFMA kernel in bf_skin.cuh is the active path.
#pragma once
#include "bf_common.cuh"
static inline bool bf_tc_enabled() { return false; }
static inline int bf_skin_forward_tc(const BfModel*, const BfVSet*, const BfFrames*, cudaStream_t) { return 1; }
