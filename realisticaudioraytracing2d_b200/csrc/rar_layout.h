// rar_layout.h -- host-side re-layout of the reference's 40-byte AoS walls into the planes the ray
// kernel reads.  Plain C++ (no CUDA), shared by the C-ABI implementation and tests/host_emulation.cpp.
//
//   geo [w] = { a.x, a.y, b.x - a.x, b.y - a.y }     16 B, the only data the inner loops touch
//   mat0[w] = { normal.x, normal.y, absorption, scattering }
//   mat1[w] = { transmission, ior }
// b - a is the `v2` of Common.hlsl:15; computing it once per wall instead of once per test performs
// the same binary32 subtraction, so results are unchanged.
#pragma once

#include "../../include/rar2d.h"
#include "rar_ray.cuh"

namespace rar {

inline void split_walls(const rar_segment *in, int n, f4 *geo, f4 *mat0, f2 *mat1) {
    for (int w = 0; w < n; w++) {
        const rar_segment &s = in[w];
        volatile float ex = s.end[0] - s.start[0];  // volatile: keep the rounded binary32 difference
        volatile float ey = s.end[1] - s.start[1];
        geo[w] = f4{s.start[0], s.start[1], ex, ey};
        mat0[w] = f4{s.normal[0], s.normal[1], s.absorption, s.scattering};
        mat1[w] = f2{s.transmission, s.ior};
    }
}

inline RayConsts ray_consts(const rar_trace_params &p) {
    RayConsts c;
    c.source_x = p.source_pos[0];
    c.source_y = p.source_pos[1];
    c.listener_x = p.listener_pos[0];
    c.listener_y = p.listener_pos[1];
    c.listener_radius = p.listener_radius;
    c.speed_of_sound = p.speed_of_sound;
    c.input_gain = p.input_gain;
    c.max_bounce_count = p.max_bounce_count;
    c.rng_state_offset = p.rng_state_offset;
    c.ray_count = p.ray_count;
    c.sample_rate = p.sample_rate;
    c.impulse_length = p.impulse_length;
    c.time_divisor = p.time_divisor;
    c.count_executed = (p.flags & RAR_FLAG_COUNT_EXECUTED) ? 1 : 0;
    return c;
}

// Thread-id range a trace call covers: the reference dispatches ceil(rayCount/64) groups of 64 threads
// with no bounds guard (Raytrace2D.compute:49-52, Helpers/ComputeHelper.cs:27-31).
inline void ray_range(const rar_trace_params &p, long long &lo, long long &hi) {
    lo = p.ray_begin;
    hi = p.ray_end;
    if (lo == 0 && hi == 0) {
        hi = (p.flags & RAR_FLAG_EXACT_RAY_COUNT) ? (long long)p.ray_count : ((long long)p.ray_count + 63) / 64 * 64;
    }
}

}  // namespace rar
