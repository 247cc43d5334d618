// Chip-synchronous fast path for B1C wide-band tracking (placeholder until bds_track_fast is filled in).
#pragma once
#include "bds_track.cuh"
namespace bds {
struct FastSmem { int dummy; };
inline bool fast_wb_supported(int, int, double, double, int, double) { return false; }
__device__ inline void correlate_fast_wb(const TrkDev&, const EpochParams&, const uint32_t*, const uint32_t*, int, int,
                                         float*, FastSmem*) {}
}  // namespace bds
