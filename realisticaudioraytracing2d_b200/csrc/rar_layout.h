// rar_layout.h -- host-side re-layout of the reference's 40-byte AoS walls into the planes the ray
// kernel reads.  Plain C++ (no CUDA), shared by the C-ABI implementation and tests/host_emulation.cpp.
//
//   geo [w] = { a.x, a.y, b.x - a.x, b.y - a.y }     16 B, the only data the inner loops touch
//   mat0[w] = { normal.x, normal.y, absorption, scattering }
//   mat1[w] = { transmission, ior }
// b - a is the `v2` of Common.hlsl:15; computing it once per wall instead of once per test performs
// the same binary32 subtraction, so results are unchanged.
#pragma once

#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>

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

// The endpoint plane once more, two walls per record, for the packed-FP32 wall tests (trace_kernel.cu, PACKED variants):
// pair_a[p] = {x(2p), x(2p+1), y(2p), y(2p+1)}, pair_b[p] = {ex(2p), ex(2p+1), ey(2p), ey(2p+1)}; n_pairs = (n+1)/2.
// A missing second wall gets NaN coordinates: every comparison of the filter is false for it, so it is never a survivor.
inline void pair_planes(const f4 *geo, int n, f4 *pair_a, f4 *pair_b) {
    const float nan = std::numeric_limits<float>::quiet_NaN();
    for (int p = 0; 2 * p < n; p++) {
        const f4 g0 = geo[2 * p];
        const f4 g1 = 2 * p + 1 < n ? geo[2 * p + 1] : f4{nan, nan, nan, nan};
        pair_a[p] = f4{g0.x, g1.x, g0.y, g1.y};
        pair_b[p] = f4{g0.z, g1.z, g0.w, g1.w};
    }
}

// True when no wall can transmit: `rngVal < transmission` (Raytrace2D.compute:131) is then false for every draw.
inline bool walls_opaque(const rar_segment *in, int n) {
    for (int w = 0; w < n; w++)
        if (in[w].transmission > 0.0f) return false;  // NaN compares false, exactly like `rngVal < NaN`
    return true;
}

// Preconditions of the SPEC kernels (rar_math.cuh "*_inrange", rar_ray.cuh): with every wall coordinate finite and
// at most 2^30 in magnitude, the source inside the same bound, the speed of sound within [2^-20, 2^20] and at most
// 2^15 bounces, a ray's position stays below 2^43 (each step is shorter than 1e8), so for every filter survivor
// eps <= |dotP| <= 2^32 and |num1| <= 2^75, and closest / c lies within 2^+-47.
inline bool walls_bounded(const rar_segment *in, int n) {
    const float lim = 1073741824.0f;  // 2^30
    for (int w = 0; w < n; w++) {
        const float v[4] = {in[w].start[0], in[w].start[1], in[w].end[0], in[w].end[1]};
        for (float x : v)
            if (!(std::fabs(x) <= lim)) return false;  // also rejects NaN
    }
    return true;
}
inline bool spec_ranges_ok(const rar_trace_params &p, bool walls_are_bounded, bool walls_are_opaque) {
    const float lim = 1073741824.0f;
    if (!walls_are_bounded || !walls_are_opaque) return false;
    if (!(std::fabs(p.source_pos[0]) <= lim && std::fabs(p.source_pos[1]) <= lim)) return false;
    if (!(std::fabs(p.listener_pos[0]) <= lim && std::fabs(p.listener_pos[1]) <= lim)) return false;
    if (!(p.speed_of_sound >= 9.5367431640625e-07f && p.speed_of_sound <= 1048576.0f)) return false;
    return p.max_bounce_count <= 32768;
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
    c.sample_rate_f = (float)p.sample_rate;
    c.impulse_length_f = (float)p.impulse_length;
    c.air_on = 0;
    for (int b = 0; b < 8; b++) c.air[b] = 0.0f;
    return c;
}

// Uniform grid over the walls for RAR_FLAG_USE_GRID (see rar_ray.cuh GridView).  Conservative by construction:
// a wall is registered in every cell whose box, grown by the margin m, meets the wall's supporting strip
// (separating-axis test on the wall's normal, inside the wall's bounding box grown by m).
//
// The frame (extent, cell size, margin) is computed on the host in one pass over the walls; which cells a wall is
// registered in is decided by wall_cell_box / wall_in_cell, plain IEEE double arithmetic shared by the host builder
// below (tests, host emulation) and the device builder (grid_kernel.cu, compiled without FMA contraction), so both
// produce the same lists.
struct GridFrame {
    float x0 = 0, y0 = 0, cw = 1, ch = 1;  // the float-rounded frame the ray kernels walk
    int nx = 0, ny = 0;
    double m = 0;                          // registration margin
};

// Cell index range of the wall's bounding box grown by the margin, clipped to the grid.
RAR_HD void wall_cell_box(const GridFrame &g, double ax, double ay, double bx, double by, int &ix0, int &ix1, int &iy0, int &iy1) {
    const double fx0 = g.x0, fy0 = g.y0, fcw = g.cw, fch = g.ch;
    const double lox = ax < bx ? ax : bx, hix = ax < bx ? bx : ax, loy = ay < by ? ay : by, hiy = ay < by ? by : ay;
    ix0 = (int)floor((lox - g.m - fx0) / fcw);
    ix1 = (int)floor((hix + g.m - fx0) / fcw);
    iy0 = (int)floor((loy - g.m - fy0) / fch);
    iy1 = (int)floor((hiy + g.m - fy0) / fch);
    ix0 = ix0 > 0 ? ix0 : 0;
    iy0 = iy0 > 0 ? iy0 : 0;
    ix1 = ix1 < g.nx - 1 ? ix1 : g.nx - 1;
    iy1 = iy1 < g.ny - 1 ? iy1 : g.ny - 1;
}

// Does the wall's supporting strip meet cell (ix, iy) grown by the margin?
RAR_HD bool wall_in_cell(const GridFrame &g, double ax, double ay, double bx, double by, int ix, int iy) {
    const double fx0 = g.x0, fy0 = g.y0, fcw = g.cw, fch = g.ch;
    const double vx = bx - ax, vy = by - ay;
    const double hx = 0.5 * fcw + g.m, hy = 0.5 * fch + g.m;
    const double reach = fabs(vy) * hx + fabs(vx) * hy;
    const double cx = fx0 + (ix + 0.5) * fcw, cy = fy0 + (iy + 0.5) * fch;
    const double dist = fabs(vx * (cy - ay) - vy * (cx - ax));
    return !(dist > reach * (1.0 + 1e-9) + 1e-300);
}

struct GridHost {
    float x0 = 0, y0 = 0, cw = 1, ch = 1;
    int nx = 0, ny = 0;
    std::vector<uint32_t> cell_start, items;
    std::vector<f4> item_geo;  // endpoint record of items[i], so a cell's walls are one contiguous read
};

// The frame for a wall set; nx == 0 when no grid can be built (no walls, non-finite coordinates: brute force handles
// those).  *bound receives an upper bound of the number of (cell, wall) registrations: the cells of every wall's box.
inline GridFrame grid_frame(const rar_segment *walls, int n, long long *bound = nullptr) {
    GridFrame g;
    if (bound) *bound = 0;
    if (n <= 0) return g;
    double minx = 1e300, miny = 1e300, maxx = -1e300, maxy = -1e300, maxabs = 0;
    for (int w = 0; w < n; w++) {
        for (int e = 0; e < 2; e++) {
            const double x = e ? walls[w].end[0] : walls[w].start[0], y = e ? walls[w].end[1] : walls[w].start[1];
            if (!(std::isfinite(x) && std::isfinite(y))) return g;
            minx = std::min(minx, x); maxx = std::max(maxx, x);
            miny = std::min(miny, y); maxy = std::max(maxy, y);
            maxabs = std::max(maxabs, std::max(std::fabs(x), std::fabs(y)));
        }
    }
    const double pad = 1e-3 * std::max(maxx - minx, maxy - miny) + 1e-4;
    minx -= pad; miny -= pad; maxx += pad; maxy += pad;
    const double ex = maxx - minx, ey = maxy - miny;
    // 0.25 ... 2 cells per wall measured within 5 % of each other on the 10k-wall maze; 0.5 keeps the lists short
    const double target_cells = std::max(1.0, n * 0.5);
    const double cell = std::sqrt(ex * ey / target_cells);
    const int nx = (int)std::min(2048.0, std::max(1.0, std::ceil(ex / cell)));
    const int ny = (int)std::min(2048.0, std::max(1.0, std::ceil(ey / cell)));
    const double cw = ex / nx, ch = ey / ny;
    const double ulp = std::ldexp(std::max(maxabs, 1e-30), -23);
    g.m = std::max(0.02 * std::min(cw, ch), 128.0 * ulp);
    g.x0 = (float)minx; g.y0 = (float)miny; g.cw = (float)cw; g.ch = (float)ch; g.nx = nx; g.ny = ny;
    if (bound) {
        for (int w = 0; w < n; w++) {
            int ix0, ix1, iy0, iy1;
            wall_cell_box(g, walls[w].start[0], walls[w].start[1], walls[w].end[0], walls[w].end[1], ix0, ix1, iy0, iy1);
            if (ix1 >= ix0 && iy1 >= iy0) *bound += (long long)(ix1 - ix0 + 1) * (iy1 - iy0 + 1);
        }
    }
    return g;
}

inline void build_grid(const rar_segment *walls, int n, GridHost &g) {
    g = GridHost();
    const GridFrame fr = grid_frame(walls, n);
    if (fr.nx <= 0) return;
    const int nx = fr.nx, ny = fr.ny;
    g.x0 = fr.x0; g.y0 = fr.y0; g.cw = fr.cw; g.ch = fr.ch; g.nx = nx; g.ny = ny;
    std::vector<uint32_t> count((size_t)nx * ny + 1, 0);
    for (int pass = 0; pass < 2; pass++) {
        for (int w = 0; w < n; w++) {
            const double ax = walls[w].start[0], ay = walls[w].start[1], bx = walls[w].end[0], by = walls[w].end[1];
            int ix0, ix1, iy0, iy1;
            wall_cell_box(fr, ax, ay, bx, by, ix0, ix1, iy0, iy1);
            for (int iy = iy0; iy <= iy1; iy++) {
                for (int ix = ix0; ix <= ix1; ix++) {
                    if (!wall_in_cell(fr, ax, ay, bx, by, ix, iy)) continue;
                    const size_t cell_id = (size_t)iy * nx + ix;
                    if (pass == 0) count[cell_id + 1]++;
                    else g.items[g.cell_start[cell_id] + count[cell_id]++] = (uint32_t)w;
                }
            }
        }
        if (pass == 1) {
            g.item_geo.resize(g.items.size());
            for (size_t i = 0; i < g.items.size(); i++) {
                const rar_segment &sgm = walls[g.items[i]];
                volatile float ex = sgm.end[0] - sgm.start[0];  // the same rounded difference as split_walls()
                volatile float ey = sgm.end[1] - sgm.start[1];
                g.item_geo[i] = f4{sgm.start[0], sgm.start[1], ex, ey};
            }
        }
        if (pass == 0) {
            g.cell_start.assign((size_t)nx * ny + 1, 0);
            for (size_t c = 0; c < (size_t)nx * ny; c++) g.cell_start[c + 1] = g.cell_start[c] + count[c + 1];
            g.items.assign(g.cell_start.back(), 0);
            std::fill(count.begin(), count.end(), 0);
        }
    }
}

// FNV-style digest of a grid's lists (tests: the device-built grid equals the host-built one).
inline uint64_t grid_digest(const uint32_t *cell_start, size_t n_cells_plus_1, const uint32_t *items, size_t n_items) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n_cells_plus_1; i++) h = (h ^ cell_start[i]) * 1099511628211ull;
    for (size_t i = 0; i < n_items; i++) h = (h ^ items[i]) * 1099511628211ull;
    return h;
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
