// rar_ray.cuh -- one ray of Raytrace2D.compute `Trace` (:49-156) as a per-bounce state machine.
//
// ray_init() and the bounce pieces (bounce_nearest, listener_direct, bounce_advance, listener_nee, the shadow
// phase, nee_arrival, bounce_scatter) are written once and used by the CUDA kernels (trace_kernel.cu); they also
// compile as host C++ for tests/host_emulation.cpp.  A bounce yields up to two pending arrivals (direct
// listener crossing, next-event estimate); the caller deposits them, which lets the kernel do that with
// warp-convergent aggregation.
//
// The inner loops over walls do NOT evaluate Common.hlsl:14-21 `intersect` literally.  They run a
// division-free conservative filter and evaluate the exact acceptance test (one correctly rounded division for
// t1; t2 in [0,1] decided by sign and magnitude comparisons equivalent to comparing the rounded quotient) only
// for the survivors, so the accepted hits and their distances are bit-identical to the literal evaluation:
//   literal: dotP = dot(v2,v3); t1 = cross(v2,v1)/dotP; t2 = dot(v1,v3)/dotP;
//            hit iff |dotP| >= eps, t1 >= eps, 0 <= t2 <= 1; nearest loop also needs t1 < closest.
//   filter : with num1 = cross(v2,v1), num2 = dot(v1,v3), r = closest*(1+2^-20)*dotP
//            pass iff |2*num2 - dotP| <= |dotP|   (<=> num2/dotP in [0,1], incl. num2 = -0)
//                 and |2*num1 - r|    <= |r|      (<=> num1/dotP in [0, closest*(1+2^-20)])
//   For binary32 a, b>0: a > b implies fl(a/b) > 1 (a/b > 1+2^-24), and rounding is monotone, so every
//   hit the literal test accepts passes the filter.  The one literal-accepting case the filter does
//   not see is t2 = num2/dotP underflowing to -0 for opposite signs, which needs |num2| < 2^-100 for
//   any scene with coordinates below 2^50; HLSL flushes such values anyway.  ("filter".)
//
// Walls are scanned by brute force (the reference's algorithm and the measured mode) or, on request, looked up
// through a uniform grid that visits a superset of the walls a query could accept (GridView below).
#pragma once

#include <cmath>

#include "rar_math.cuh"

namespace rar {

struct alignas(16) f4 { float x, y, z, w; };
struct alignas(8) f2 { float x, y; };

// The fields of rar_trace_params the ray logic reads (same meaning, see include/rar2d.h).
struct RayConsts {
    float source_x, source_y, listener_x, listener_y;
    float listener_radius, speed_of_sound, input_gain;
    int max_bounce_count;
    uint32_t rng_state_offset;
    int ray_count;
    int sample_rate, impulse_length;
    float time_divisor;
    int count_executed;  // with COUNT: resolve (and count) only the shadow rays the production kernel resolves
    float sample_rate_f, impulse_length_f;  // (float) of the two integers above, converted once on the host
    // banded model: air absorption of the (up to 8) bands this launch carries, 1/m; band arrivals are scaled by
    // exp(-air[b] * path length) when air_on (rar_set_air_absorption)
    int air_on;
    float air[8];
};

// SPEC kernels (see Scene::kSpec): the refined reciprocal of the launch-invariant speed of sound, formed once per
// thread from the MUFU seed, so that x / c is three FMAs (div_with_rcp) instead of the guarded 12-instruction
// sequence -- bit-identical to __fdiv_rn for the operands the SPEC preconditions admit.
struct SpecConsts {
    float c, inv_c;
};
RAR_HD SpecConsts spec_consts(const RayConsts &p) {
    SpecConsts s;
    s.c = p.speed_of_sound;
    s.inv_c = rcp_refine(p.speed_of_sound, mufu_rcp(p.speed_of_sound));
    return s;
}

struct RayCounters {
    unsigned long long ray_bounces, nearest_tests, shadow_tests, direct_hits, nee_hits;
};

template <int BANDS>
struct RayState {
    float px, py, dx, dy;
    float energy, time, dist, speed;
    int wall_depth;
    uint32_t rng;
    float band_e[BANDS > 1 ? BANDS : 1];
};

template <int BANDS>
struct Arrival {
    int has;
    float t, e, hx, hy;
    float band_e[BANDS > 1 ? BANDS : 1];
};

// Raytrace2D.compute:51-61
template <int BANDS>
RAR_HD void ray_init(RayState<BANDS> &r, uint32_t id, const RayConsts &p, uint32_t frame = 0) {
    r.rng = id + (p.rng_state_offset + frame) * 719393u;  // frame: extra frames batched into one launch
    float u = pcg_random(r.rng);
    float angle = rar_div((float)id + u, (float)p.ray_count) * 2.0f * kPi;
    sincos_poly(angle, r.dy, r.dx);
    r.px = p.source_x;
    r.py = p.source_y;
    r.energy = p.input_gain;
    r.time = 0.0f;
    r.dist = 0.0f;
    r.speed = p.speed_of_sound;
    r.wall_depth = 0;
    if (BANDS > 1) {
#pragma unroll
        for (int b = 0; b < BANDS; b++) r.band_e[b] = p.input_gain;
    }
}

// Common.hlsl:14-21 on the precomputed edge e = b - a; used on filter survivors only.  t1 is the
// correctly rounded quotient the literal formula returns.  The test on t2 = fl(num2/dotP) needs no
// division: for binary32 values fl(num2/dotP) <= 1 <=> |num2| <= |dotP| (a > b > 0 implies a/b > 1 + 2^-24,
// which cannot round down to 1), and fl(num2/dotP) >= 0 <=> num2 is zero or has the sign of dotP (the only
// other way is a quotient of opposite signs underflowing to -0, |num2/dotP| <= 2^-150, which the filter
// has already excluded -- see the header comment).
// SPEC (scene and source coordinates bounded by 2^30, checked by the host): eps <= |dotP| <= 2^40 and
// |num1| <= 2^63, so the division needs no range guard; a |num1| below 2^-100 yields some quotient below 2^-80,
// rejected by t1 >= eps exactly like the true quotient.
template <bool SPEC = false>
RAR_HD float intersect_exact(float num1, float num2, float dotP) {
    if (fabsf(dotP) < kEps) return kInf;
    const float t1 = SPEC ? div_inrange(num1, dotP) : rar_div(num1, dotP);
    const bool same_sign = (num2 >= 0.0f) == (dotP >= 0.0f);
    const bool t2_ok = (num2 == 0.0f || same_sign) && fabsf(num2) <= fabsf(dotP);
    return (t1 >= kEps && t2_ok) ? t1 : kInf;
}

constexpr float kSlack = 1.00000095367431640625f;  // 1 + 2^-20

// The per-wall quantities of the filter (see the header comment).
struct WallTest {
    float dotP, num1, num2;
};
RAR_HD WallTest wall_test(const f4 s, float ox, float oy, float dx, float ndy) {
    WallTest t;
    const float v1x = ox - s.x, v1y = oy - s.y;
    t.dotP = rar_fma(s.z, ndy, s.w * dx);
    t.num2 = rar_fma(v1x, ndy, v1y * dx);
    t.num1 = rar_fma(s.z, v1y, -(s.w * v1x));
    return t;
}
// bound_m = (closest or shadow limit) * (1 + 2^-20).  A stale (larger) bound only loosens the filter.
RAR_HD bool wall_pass(const WallTest &t, float bound_m) {
    const float r = bound_m * t.dotP;
    return (fabsf(rar_fma(2.0f, t.num2, -t.dotP)) <= fabsf(t.dotP)) & (fabsf(rar_fma(2.0f, t.num1, -r)) <= fabsf(r));
}

// The filter while no hit has been found yet (bound = inf).  With an infinite bound the second clause of
// wall_pass accepts every wall, including those the ray's line crosses BEHIND the origin, which the literal
// formula then rejects at the price of a division.  Until a bound exists the clause is replaced by the sign
// test it degenerates from: t1 = num1/dotP >= eps > 0 needs num1 and dotP of the same sign.
RAR_HD bool wall_pass_unbounded(const WallTest &t) {
#ifdef __CUDA_ARCH__
    const bool forward = (__float_as_int(t.num1) ^ __float_as_int(t.dotP)) >= 0;
#else
    const bool forward = std::signbit(t.num1) == std::signbit(t.dotP);
#endif
    return (fabsf(rar_fma(2.0f, t.num2, -t.dotP)) <= fabsf(t.dotP)) & forward;
}

// ---- packed wall tests (device only) ------------------------------------------------------------------------
//
// sm_100 has packed binary32 arithmetic (PTX fma/mul/add/sub .rn.f32x2 -> FFMA2 / FMUL2 / FADD2): one instruction, two
// IEEE results, each rounded exactly like its scalar form.  Measured (tools/microbench.cu): a packed instruction
// issues at half the rate of a scalar one, i.e. the same FP32 results per second for half the issue slots -- and the
// wall scans are bound by issue slots (profiles/r02_phase_tables.txt).  The filter of two walls is evaluated in one
// pass over a pair record (rar_layout.h pair_planes); the ray's constants enter as scalar operands, which the
// hardware broadcasts to both halves.  Bit-identity with wall_test / wall_pass: the quantities are the same
// expressions except for three sign conventions chosen so that no negation instruction is needed, each exact --
//   nv1x = sx - ox = -(ox - sx),     ndotP = fma(ex, dy, ey * (-dx)) = -dotP,     nr = bound * ndotP = -r
// (a product or an FMA of negated operands is the negated result in round-to-nearest) -- and the comparisons use
// magnitudes only.  Where a zero's sign differs (x - x is +0 either way round) the result of the FILTER can only
// differ for num1 = 0 or dotP = 0, which the exact evaluation rejects in both cases (t1 = 0 < eps, |dotP| < eps).
#ifdef __CUDACC__
struct WallTest2 {
    uint64_t ndotP, num1, num2;  // (wall 2p, wall 2p+1)
};
__device__ __forceinline__ uint64_t pair_bc(float v) { return (uint64_t)__float_as_uint(v) | ((uint64_t)__float_as_uint(v) << 32); }
__device__ __forceinline__ float pair_lo(uint64_t v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float pair_hi(uint64_t v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
template <class Scene>
__device__ __forceinline__ WallTest2 wall_test2(const Scene &sc, int p, float ox, float oy, float dx, float ndy) {
    uint64_t sx, sy, ex, ey;
    sc.pair(p, sx, sy, ex, ey);
    const uint64_t dy2 = pair_bc(-ndy), dx2 = pair_bc(dx);
    const uint64_t nv1x = sub2(sx, pair_bc(ox));
    const uint64_t v1y = sub2(pair_bc(oy), sy);
    WallTest2 t;
    t.ndotP = fma2(ex, dy2, mul2(ey, pair_bc(-dx)));
    t.num2 = fma2(nv1x, dy2, mul2(v1y, dx2));
    t.num1 = fma2(ex, v1y, mul2(ey, nv1x));
    return t;
}
__device__ __forceinline__ void wall_pass2(const WallTest2 &t, float bound_m, bool &p0, bool &p1) {
    const uint64_t two = pair_bc(2.0f);
    const uint64_t nr = mul2(pair_bc(bound_m), t.ndotP);
    const uint64_t c1 = fma2(two, t.num2, t.ndotP);  // 2 num2 - dotP
    const uint64_t c2 = fma2(two, t.num1, nr);       // 2 num1 - r
    p0 = (fabsf(pair_lo(c1)) <= fabsf(pair_lo(t.ndotP))) & (fabsf(pair_lo(c2)) <= fabsf(pair_lo(nr)));
    p1 = (fabsf(pair_hi(c1)) <= fabsf(pair_hi(t.ndotP))) & (fabsf(pair_hi(c2)) <= fabsf(pair_hi(nr)));
}
// wall_pass_unbounded for a pair: "num1 and dotP of the same sign" reads "num1 and ndotP of opposite signs"
__device__ __forceinline__ void wall_pass_unbounded2(const WallTest2 &t, bool &p0, bool &p1) {
    const uint64_t c1 = fma2(pair_bc(2.0f), t.num2, t.ndotP);
    const uint64_t x = t.num1 ^ t.ndotP;  // sign bits: bit 31 (wall 2p), bit 63 (wall 2p+1)
    p0 = (fabsf(pair_lo(c1)) <= fabsf(pair_lo(t.ndotP))) & ((int)(unsigned)x < 0);
    p1 = (fabsf(pair_hi(c1)) <= fabsf(pair_hi(t.ndotP))) & ((long long)x < 0);
}
#endif

// ---- optional uniform grid over the walls (RAR_FLAG_USE_GRID) -------------------------------------------
//
// SURVEY.md 8(f)-2.  Brute force stays the default and the measured mode; the grid changes which walls are
// LOOKED AT, never the outcome: every wall a query could accept is still evaluated with the same filter and
// the same literal formula, and ties are broken by wall index exactly as the ascending brute-force scan does.
// Why nothing is missed: a wall is registered in every cell within a margin m of any of its points
// (m = 2 % of a cell, at least 128 ulp of the largest coordinate; rar_layout.h build_grid), and the cell walk
// follows the ray to within float rounding (<< m).  So a point of the ray that is farther than m from every
// visited cell does not exist, and a wall whose hit point lies on the visited stretch of the ray is registered
// in a visited cell.  The nearest-hit walk stops once the best hit lies before the exit of the current cell.
struct GridView {
    float x0, y0;          // lower corner of cell (0, 0)
    float cw, ch;          // cell size
    float inv_cw, inv_ch;
    int nx, ny;
    const uint32_t *cell_start;  // [nx * ny + 1]
    const uint32_t *items;       // wall indices, ascending within a cell
    const f4 *item_geo;          // endpoint record {a.x, a.y, e.x, e.y} of items[i]
};

struct GridWalk {
    int ix, iy, stepx, stepy;
    float tmx, tmy, tdx, tdy;  // ray parameter of the next x / y cell boundary, and per-cell increments
};

// Clips the half-line o + t d, t >= 0, to the grid and positions the walk on its first cell.
RAR_HD bool grid_walk_init(const GridView &g, float ox, float oy, float dx, float dy, GridWalk &w) {
    const float bx1 = g.x0 + (float)g.nx * g.cw, by1 = g.y0 + (float)g.ny * g.ch;
    const float big = 3.0e38f;
    float t0 = 0.0f, t1 = big;
    const float idx = dx != 0.0f ? 1.0f / dx : 0.0f, idy = dy != 0.0f ? 1.0f / dy : 0.0f;
    if (dx != 0.0f) {
        const float ta = (g.x0 - ox) * idx, tb = (bx1 - ox) * idx;
        t0 = fmaxf(t0, fminf(ta, tb));
        t1 = fminf(t1, fmaxf(ta, tb));
    } else if (ox < g.x0 || ox > bx1) {
        return false;
    }
    if (dy != 0.0f) {
        const float ta = (g.y0 - oy) * idy, tb = (by1 - oy) * idy;
        t0 = fmaxf(t0, fminf(ta, tb));
        t1 = fminf(t1, fmaxf(ta, tb));
    } else if (oy < g.y0 || oy > by1) {
        return false;
    }
    if (!(t0 <= t1)) return false;
    const float px = ox + dx * t0, py = oy + dy * t0;
    int ix = (int)floorf((px - g.x0) * g.inv_cw), iy = (int)floorf((py - g.y0) * g.inv_ch);
    ix = ix < 0 ? 0 : (ix >= g.nx ? g.nx - 1 : ix);
    iy = iy < 0 ? 0 : (iy >= g.ny ? g.ny - 1 : iy);
    w.ix = ix;
    w.iy = iy;
    w.stepx = dx > 0.0f ? 1 : -1;
    w.stepy = dy > 0.0f ? 1 : -1;
    w.tmx = dx != 0.0f ? (g.x0 + (float)(ix + (dx > 0.0f ? 1 : 0)) * g.cw - ox) * idx : big;
    w.tmy = dy != 0.0f ? (g.y0 + (float)(iy + (dy > 0.0f ? 1 : 0)) * g.ch - oy) * idy : big;
    w.tdx = dx != 0.0f ? g.cw * fabsf(idx) : big;
    w.tdy = dy != 0.0f ? g.ch * fabsf(idy) : big;
    return true;
}

// Moves to the next cell along the ray; false when the walk leaves the grid.
RAR_HD bool grid_walk_step(const GridView &g, GridWalk &w) {
    if (w.tmx < w.tmy) {
        w.ix += w.stepx;
        w.tmx += w.tdx;
        return w.ix >= 0 && w.ix < g.nx;
    }
    w.iy += w.stepy;
    w.tmy += w.tdy;
    return w.iy >= 0 && w.iy < g.ny;
}

template <class Scene>
RAR_HD void nearest_hit_grid(const Scene &sc, float ox, float oy, float dx, float dy, float &closest_out, int &hit_out,
                             int *tests) {
    float closest = kInf, closest_m = kInf * kSlack;
    int hit = -1, n_tests = 0;
    const float ndy = -dy;
    const GridView &g = sc.grid();
    GridWalk w;
    if (g.nx > 0 && grid_walk_init(g, ox, oy, dx, dy, w)) {
        for (int guard = g.nx + g.ny + 2; guard > 0; guard--) {
            const int cell = w.iy * g.nx + w.ix;
            const uint32_t i0 = g.cell_start[cell], i1 = g.cell_start[cell + 1];
            for (uint32_t i = i0; i < i1; i++) {
                const WallTest t = wall_test(sc.grid_geo(i), ox, oy, dx, ndy);
                n_tests++;
                if (wall_pass(t, closest_m)) {
                    const int k = (int)g.items[i];
                    const float d = intersect_exact(t.num1, t.num2, t.dotP);
                    // cells are not visited in wall order: break ties by index like the ascending scan does
                    if (d < closest || (d == closest && d < kInf && k < hit)) {
                        closest = d;
                        closest_m = d * kSlack;
                        hit = k;
                    }
                }
            }
            if (closest <= fminf(w.tmx, w.tmy)) break;  // the best hit lies before this cell's exit
            if (!grid_walk_step(g, w)) break;
        }
    }
    if (tests) *tests = n_tests;
    closest_out = closest;
    hit_out = hit;
}

template <class Scene>
RAR_HD bool check_vis_grid(const Scene &sc, float sx, float sy, float dx, float ndy, float lim, int *tests) {
    const GridView &g = sc.grid();
    int n_tests = 0;
    bool blocked = false;
    if (sc.n_walls() > 0 && kInf < lim) {
        blocked = true;  // every intersect() result (<= inf) is < lim
        n_tests = 1;
    } else {
        const float lim_m = lim * kSlack;
        GridWalk w;
        if (g.nx > 0 && grid_walk_init(g, sx, sy, dx, -ndy, w)) {
            for (int guard = g.nx + g.ny + 2; guard > 0 && !blocked; guard--) {
                const int cell = w.iy * g.nx + w.ix;
                const uint32_t i0 = g.cell_start[cell], i1 = g.cell_start[cell + 1];
                for (uint32_t i = i0; i < i1; i++) {
                    const WallTest t = wall_test(sc.grid_geo(i), sx, sy, dx, ndy);
                    n_tests++;
                    if (wall_pass(t, lim_m) && intersect_exact(t.num1, t.num2, t.dotP) < lim) {
                        blocked = true;
                        break;
                    }
                }
                if (fminf(w.tmx, w.tmy) >= lim) break;  // walls beyond the limit cannot block
                if (!grid_walk_step(g, w)) break;
            }
        }
    }
    if (tests) *tests = n_tests;
    return !blocked;
}

// Raytrace2D.compute:69-72: nearest hit over all walls, lowest index wins ties.  Walls are filtered four
// at a time against the same bound so that the common case is four independent loads, four filters and
// a single branch; survivors are then evaluated literally, in wall order.
template <class Scene>
RAR_HD void nearest_hit(const Scene &sc, float ox, float oy, float dx, float dy, float &closest_out, int &hit_out,
                        int *grid_tests = nullptr) {
    if constexpr (Scene::kGrid) {
        nearest_hit_grid(sc, ox, oy, dx, dy, closest_out, hit_out, grid_tests);
        return;
    }
    float closest = kInf, closest_m = kInf * kSlack;
    int hit = -1;
    const float ndy = -dy;
    const int n = sc.n_walls();
    int w = 0;
#define RAR_NEAREST_EXACT(T, W)                                      \
    {                                                                \
        const float d = intersect_exact<Scene::kSpec>((T).num1, (T).num2, (T).dotP); \
        if (d < closest) {                                           \
            closest = d;                                             \
            closest_m = d * kSlack;                                  \
            hit = (W);                                               \
        }                                                            \
    }
#define RAR_NEAREST_EXACT2(T, H, W)                                  \
    {                                                                \
        const float d = intersect_exact<Scene::kSpec>(pair_##H((T).num1), pair_##H((T).num2), -pair_##H((T).ndotP)); \
        if (d < closest) {                                           \
            closest = d;                                             \
            closest_m = d * kSlack;                                  \
            hit = (W);                                               \
        }                                                            \
    }
#ifdef __CUDACC__
    if constexpr (Scene::kPacked && Scene::kFixed4) {  // exactly four walls = two pair records, no bound yet
        const WallTest2 ta = wall_test2(sc, 0, ox, oy, dx, ndy), tb = wall_test2(sc, 1, ox, oy, dx, ndy);
        bool p0, p1, p2, p3;
        wall_pass_unbounded2(ta, p0, p1);
        wall_pass_unbounded2(tb, p2, p3);
        if (p0 | p1 | p2 | p3) {
            if (p0) RAR_NEAREST_EXACT2(ta, lo, 0)
            if (p1) RAR_NEAREST_EXACT2(ta, hi, 1)
            if (p2) RAR_NEAREST_EXACT2(tb, lo, 2)
            if (p3) RAR_NEAREST_EXACT2(tb, hi, 3)
        }
        closest_out = closest;
        hit_out = hit;
        return;
    } else if constexpr (Scene::kPacked) {
        // two pair records = four walls per iteration, in wall order (ties go to the lower index, as in the scalar scan)
        const int np = (n + 1) >> 1;
        int p = 0;
        for (; p + 2 <= np; p += 2) {
            const WallTest2 ta = wall_test2(sc, p, ox, oy, dx, ndy), tb = wall_test2(sc, p + 1, ox, oy, dx, ndy);
            bool p0, p1, p2, p3;
            wall_pass2(ta, closest_m, p0, p1);
            wall_pass2(tb, closest_m, p2, p3);
            if (p0 | p1 | p2 | p3) {
                if (p0) RAR_NEAREST_EXACT2(ta, lo, 2 * p)
                if (p1) RAR_NEAREST_EXACT2(ta, hi, 2 * p + 1)
                if (p2) RAR_NEAREST_EXACT2(tb, lo, 2 * p + 2)
                if (p3) RAR_NEAREST_EXACT2(tb, hi, 2 * p + 3)
            }
        }
        if (p < np) {
            const WallTest2 ta = wall_test2(sc, p, ox, oy, dx, ndy);
            bool p0, p1;
            wall_pass2(ta, closest_m, p0, p1);
            if (p0) RAR_NEAREST_EXACT2(ta, lo, 2 * p)
            if (p1) RAR_NEAREST_EXACT2(ta, hi, 2 * p + 1)
        }
        closest_out = closest;
        hit_out = hit;
        return;
    }
#endif
    if (Scene::kFixed4 || (Scene::kPeelFirstBatch && n >= 4)) {  // first batch: no bound yet
        const WallTest t0 = wall_test(sc.geo(0), ox, oy, dx, ndy);
        const WallTest t1 = wall_test(sc.geo(1), ox, oy, dx, ndy);
        const WallTest t2 = wall_test(sc.geo(2), ox, oy, dx, ndy);
        const WallTest t3 = wall_test(sc.geo(3), ox, oy, dx, ndy);
        const bool p0 = wall_pass_unbounded(t0), p1 = wall_pass_unbounded(t1);
        const bool p2 = wall_pass_unbounded(t2), p3 = wall_pass_unbounded(t3);
        if (p0 | p1 | p2 | p3) {
            if (p0) RAR_NEAREST_EXACT(t0, 0)
            if (p1) RAR_NEAREST_EXACT(t1, 1)
            if (p2) RAR_NEAREST_EXACT(t2, 2)
            if (p3) RAR_NEAREST_EXACT(t3, 3)
        }
        w = 4;
    }
    if constexpr (!Scene::kFixed4) {  // kFixed4: the scene has exactly four walls, the batch above was all of them
        for (; w + 4 <= n; w += 4) {
            const WallTest t0 = wall_test(sc.geo(w), ox, oy, dx, ndy);
            const WallTest t1 = wall_test(sc.geo(w + 1), ox, oy, dx, ndy);
            const WallTest t2 = wall_test(sc.geo(w + 2), ox, oy, dx, ndy);
            const WallTest t3 = wall_test(sc.geo(w + 3), ox, oy, dx, ndy);
            const bool p0 = wall_pass(t0, closest_m), p1 = wall_pass(t1, closest_m);
            const bool p2 = wall_pass(t2, closest_m), p3 = wall_pass(t3, closest_m);
            if (p0 | p1 | p2 | p3) {
                if (p0) RAR_NEAREST_EXACT(t0, w)
                if (p1) RAR_NEAREST_EXACT(t1, w + 1)
                if (p2) RAR_NEAREST_EXACT(t2, w + 2)
                if (p3) RAR_NEAREST_EXACT(t3, w + 3)
            }
        }
        for (; w < n; w++) {
            const WallTest t = wall_test(sc.geo(w), ox, oy, dx, ndy);
            if (wall_pass(t, closest_m)) RAR_NEAREST_EXACT(t, w)
        }
    }
#undef RAR_NEAREST_EXACT
#undef RAR_NEAREST_EXACT2
    closest_out = closest;
    hit_out = hit;
}

// ---- shadow ray (Raytrace2D.compute:40-47 checkVis) -------------------------------------------------

// The shadow ray of one next-event estimate, ready to be tested against the walls.
struct ShadowRay {
    float sx, sy;   // start = hit point + stored normal * eps (:105)
    float dx, ndy;  // direction (end - start) * (1/dist): not unit length, as in the reference; ndy = -dy
    float lim;      // dist - 0.1: a wall hit closer than this blocks (:44)
};

RAR_HD ShadowRay make_shadow_ray(float sx, float sy, float ex, float ey, float dist) {
    ShadowRay q;
    const float inv_dist = rar_rcp(dist);  // vector / scalar := vector * (1 / scalar)
    q.sx = sx;
    q.sy = sy;
    q.dx = (ex - sx) * inv_dist;
    q.ndy = -((ey - sy) * inv_dist);
    q.lim = dist - 0.1f;
    return q;
}

// Does wall record s block the shadow ray?  (filter, then the literal formula for survivors)
template <bool SPEC = false>
RAR_HD bool shadow_blocked_by(const f4 s, const ShadowRay &q, float lim_m) {
    const WallTest t = wall_test(s, q.sx, q.sy, q.dx, q.ndy);
    return wall_pass(t, lim_m) && intersect_exact<SPEC>(t.num1, t.num2, t.dotP) < q.lim;
}

// One thread walks all walls with the reference's early exit.  Returns true when visible; *tests receives
// the number of intersect() evaluations the reference's loop performs (first blocking wall + 1, else n).
template <class Scene>
RAR_HD bool check_vis(const Scene &sc, const ShadowRay &q, int *tests) {
    if constexpr (Scene::kGrid) return check_vis_grid(sc, q.sx, q.sy, q.dx, q.ndy, q.lim, tests);
    const int n = sc.n_walls();
    if (n > 0 && kInf < q.lim) {  // every intersect() result (<= inf) is < lim: blocked by wall 0
        if (tests) *tests = 1;
        return false;
    }
    const float lim_m = q.lim * kSlack;
    constexpr bool SP = Scene::kSpec;
#ifdef __CUDACC__
    if constexpr (Scene::kFixed4 && Scene::kPacked) {
        const WallTest2 ta = wall_test2(sc, 0, q.sx, q.sy, q.dx, q.ndy), tb = wall_test2(sc, 1, q.sx, q.sy, q.dx, q.ndy);
        bool p0, p1, p2, p3;
        wall_pass2(ta, lim_m, p0, p1);
        wall_pass2(tb, lim_m, p2, p3);
        int first4 = -1;
        if (p0 | p1 | p2 | p3) {
            if (p3 && intersect_exact<SP>(pair_hi(tb.num1), pair_hi(tb.num2), -pair_hi(tb.ndotP)) < q.lim) first4 = 3;
            if (p2 && intersect_exact<SP>(pair_lo(tb.num1), pair_lo(tb.num2), -pair_lo(tb.ndotP)) < q.lim) first4 = 2;
            if (p1 && intersect_exact<SP>(pair_hi(ta.num1), pair_hi(ta.num2), -pair_hi(ta.ndotP)) < q.lim) first4 = 1;
            if (p0 && intersect_exact<SP>(pair_lo(ta.num1), pair_lo(ta.num2), -pair_lo(ta.ndotP)) < q.lim) first4 = 0;
        }
        if (tests) *tests = first4 >= 0 ? first4 + 1 : 4;
        return first4 < 0;
    }
#endif
    if constexpr (Scene::kFixed4) {  // exactly four walls: one batch, no loop
        const WallTest t0 = wall_test(sc.geo(0), q.sx, q.sy, q.dx, q.ndy);
        const WallTest t1 = wall_test(sc.geo(1), q.sx, q.sy, q.dx, q.ndy);
        const WallTest t2 = wall_test(sc.geo(2), q.sx, q.sy, q.dx, q.ndy);
        const WallTest t3 = wall_test(sc.geo(3), q.sx, q.sy, q.dx, q.ndy);
        const bool p0 = wall_pass(t0, lim_m), p1 = wall_pass(t1, lim_m);
        const bool p2 = wall_pass(t2, lim_m), p3 = wall_pass(t3, lim_m);
        int first4 = -1;
        if (p0 | p1 | p2 | p3) {
            if (p3 && intersect_exact<SP>(t3.num1, t3.num2, t3.dotP) < q.lim) first4 = 3;
            if (p2 && intersect_exact<SP>(t2.num1, t2.num2, t2.dotP) < q.lim) first4 = 2;
            if (p1 && intersect_exact<SP>(t1.num1, t1.num2, t1.dotP) < q.lim) first4 = 1;
            if (p0 && intersect_exact<SP>(t0.num1, t0.num2, t0.dotP) < q.lim) first4 = 0;
        }
        if (tests) *tests = first4 >= 0 ? first4 + 1 : 4;
        return first4 < 0;
    }
    int w = 0;
    int first = -1;  // first blocking wall
    for (; w + 4 <= n; w += 4) {
        const WallTest t0 = wall_test(sc.geo(w), q.sx, q.sy, q.dx, q.ndy);
        const WallTest t1 = wall_test(sc.geo(w + 1), q.sx, q.sy, q.dx, q.ndy);
        const WallTest t2 = wall_test(sc.geo(w + 2), q.sx, q.sy, q.dx, q.ndy);
        const WallTest t3 = wall_test(sc.geo(w + 3), q.sx, q.sy, q.dx, q.ndy);
        const bool p0 = wall_pass(t0, lim_m), p1 = wall_pass(t1, lim_m);
        const bool p2 = wall_pass(t2, lim_m), p3 = wall_pass(t3, lim_m);
        if (p0 | p1 | p2 | p3) {
            if (p0 && intersect_exact<SP>(t0.num1, t0.num2, t0.dotP) < q.lim) { first = w; break; }
            if (p1 && intersect_exact<SP>(t1.num1, t1.num2, t1.dotP) < q.lim) { first = w + 1; break; }
            if (p2 && intersect_exact<SP>(t2.num1, t2.num2, t2.dotP) < q.lim) { first = w + 2; break; }
            if (p3 && intersect_exact<SP>(t3.num1, t3.num2, t3.dotP) < q.lim) { first = w + 3; break; }
        }
    }
    if (first < 0) {
        for (; w < n; w++) {
            if (shadow_blocked_by<SP>(sc.geo(w), q, lim_m)) { first = w; break; }
        }
    }
    if (tests) *tests = first >= 0 ? first + 1 : n;
    return first < 0;
}

// ---- one iteration of the bounce loop (Raytrace2D.compute:66-155) in three phases --------------------
//
//   bounce_begin : nearest hit, direct listener crossing, advance, and everything of the next-event
//                  estimate that does not need the shadow ray (:69-104, :106-112 reordered);
//   shadow phase : visibility of the listener from the hit point (:105) -- check_vis() per thread, or the
//                  warp-cooperative version in trace_kernel.cu;
//   bounce_finish: emit the estimate if visible (:113-118), absorb (:121-122), transmit or reflect (:124-154).
//
// The reference evaluates checkVis before it knows whether the estimate clears the 1e-5 threshold (:111).
// checkVis has no side effect, so the phases compute the contribution first and resolve the shadow ray
// only when the result can matter (WITH COUNT the shadow ray is always resolved, so that the test counters
// are the reference's).

template <int BANDS>
struct BounceCtx {
    int hit;            // wall index, -1: ray left the scene
    float closest;      // distance to it
    int want_shadow;    // the shadow ray must be resolved
    int nee_candidate;  // contribution clears the threshold (deposit iff visible)
    ShadowRay shadow;
    f4 m0;              // nx, ny, absorption, scattering
    f2 m1;              // transmission, ior
    float keep, dir_dot_n;
    float nee_t, nee_e, geo, inv, total;  // total: path length source -> hit point -> listener of the estimate
    float band_keep[BANDS > 1 ? BANDS : 1];
};

// The pieces below are independent of the listener (nearest, advance, scatter) or take one listener
// position (direct crossing, next-event estimate).  The ray's path never depends on the listener -- it is
// tested but never hit (Raytrace2D.compute:74-84,101-119) -- which is what lets the batched-listener kernel
// (BASELINE config 4) trace each ray once and run only the per-listener pieces for every listener.

// :69-72
template <int BANDS, bool COUNT, class Scene>
RAR_HD void bounce_nearest(const Scene &sc, const RayState<BANDS> &r, BounceCtx<BANDS> &c, RayCounters *ctr) {
    int tests = sc.n_walls();  // brute force: the reference's count; grid: the tests actually evaluated
    nearest_hit(sc, r.px, r.py, r.dx, r.dy, c.closest, c.hit, (COUNT && Scene::kGrid) ? &tests : nullptr);
    if (COUNT) {
        ctr->ray_bounces += 1;
        ctr->nearest_tests += (unsigned long long)tests;
    }
}

// :74-84: does the ray cross the listener circle at (lx, ly) before the wall?  Uses the state BEFORE the advance.
// SPEC: see bounce_advance.
template <int BANDS, bool COUNT, bool SPEC = false>
RAR_HD void listener_direct(const RayConsts &p, float lx, float ly, const RayState<BANDS> &r, float closest,
                            Arrival<BANDS> &direct, RayCounters *ctr, const SpecConsts *sp = nullptr) {
    direct.has = 0;
    if (r.wall_depth != 0) return;
    float dl = intersect_circle(r.px, r.py, r.dx, r.dy, lx, ly, p.listener_radius);
    if (dl < closest && dl < kInf) {
        direct.has = 1;
        direct.hx = rar_fma(r.dx, dl, r.px);
        direct.hy = rar_fma(r.dy, dl, r.py);
        direct.t = r.time + (SPEC ? div_with_rcp(dl, sp->c, sp->inv_c) : rar_div(dl, r.speed));  // eps < dl < 1e8
        float total = r.dist + dl;
        float denom = fmaxf(1.0f, total * total);
        direct.e = rar_div(r.energy, denom);
        if (BANDS > 1) {
#pragma unroll
            for (int b = 0; b < BANDS; b++) {
                direct.band_e[b] = rar_div(r.band_e[b], denom);
                if (p.air_on) direct.band_e[b] *= exp_neg_poly(p.air[b], total);
            }
        }
        if (COUNT) ctr->direct_hits += 1;
    }
}

// :86-99: end the ray if nothing was hit, else move it to the wall and fetch the wall's material.
// dbg: where to record this bounce's vertex for the debugRays buffer (dereferenced only when dbg_flags asks for
// it); dbg_flags bit 0: record
// wall hits (:96-97, thread id < 100), bit 1: record the escape vertex (:87-88, thread id < debugRayCount).
// SPEC (opaque scene, so the ray's speed is the launch-invariant speed of sound, itself within [2^-20, 2^20]):
// closest lies in [eps, 1e8), so closest / speed needs no range guard and uses the per-thread reciprocal.
template <int BANDS, class Scene>
RAR_HD bool bounce_advance(const Scene &sc, RayState<BANDS> &r, BounceCtx<BANDS> &c, f4 *dbg, int dbg_flags,
                           const SpecConsts *sp = nullptr) {
    if (c.hit < 0) {  // :86-90
        if (dbg_flags & 2) *dbg = f4{rar_fma(r.dx, 20.0f, r.px), rar_fma(r.dy, 20.0f, r.py), 0.0f, 0.0f};
        return false;
    }
    r.px = rar_fma(r.dx, c.closest, r.px);  // :92-94
    r.py = rar_fma(r.dy, c.closest, r.py);
    r.time += Scene::kSpec ? div_with_rcp(c.closest, sp->c, sp->inv_c) : rar_div(c.closest, r.speed);
    r.dist += c.closest;
    if (dbg_flags & 1) *dbg = f4{r.px, r.py, r.energy, 0.0f};  // :96-97

    c.m0 = sc.mat0(c.hit);  // :99
    c.m1 = sc.mat1(c.hit);
    c.keep = 1.0f - c.m0.z;
    if (BANDS > 1) {
        const float *ba = sc.band_abs(c.hit);
#pragma unroll
        for (int b = 0; b < BANDS; b++) c.band_keep[b] = 1.0f - ba[b];
    }
    c.dir_dot_n = dot2(r.dx, r.dy, c.m0.x, c.m0.y);
    return true;
}

// :101-112 for the listener at (lx, ly): the estimate's energy and arrival time, and the shadow ray that
// decides it.  The reference evaluates checkVis before it knows whether the estimate clears the 1e-5
// threshold (:111); checkVis has no side effect, so the contribution is computed first and the shadow ray
// is requested only when the outcome can matter (with COUNT in reference mode it is always requested, so
// that the test counters are the reference's).
template <int BANDS, bool COUNT, bool SPEC = false>
RAR_HD void listener_nee(const RayConsts &p, float lx, float ly, const RayState<BANDS> &r, BounceCtx<BANDS> &c,
                         const SpecConsts *sp = nullptr) {
    static_assert(!(SPEC && COUNT), "the SPEC pieces exist for the production kernels only");
    c.want_shadow = 0;
    c.nee_candidate = 0;
    if (r.wall_depth != 0) return;
    const float wnx = c.m0.x, wny = c.m0.y;
    if (!(COUNT && !p.count_executed)) {
        // Cheap certificate that the estimate cannot clear the threshold, before the square root and the two
        // reciprocals: cosT <= |normal| (1 + 1e-6) and totalD >= dist, so
        //   contrib <= (E keep) 0.5 |n| / dist^2,
        // compared in squared form with a 2 % margin, far above the few ulp of the literal evaluation below.
        const float ek = (r.energy * c.keep) * 0.5f;
        const float d2 = r.dist * r.dist;
        if (ek <= 0.0f || (ek * ek) * dot2(wnx, wny, wnx, wny) < (0.96e-10f * d2) * d2) return;
    }
    const float tlx = lx - r.px, tly = ly - r.py;
    const float dl2 = dot2(tlx, tly, tlx, tly);
    const bool flip = c.dir_dot_n > 0.0f;
    const float enx = flip ? -wnx : wnx, eny = flip ? -wny : wny;
    if (SPEC && __builtin_expect(in_safe_range(dl2), 1)) {
        // One range test for the whole estimate: 2^-100 <= dl^2 <= 2^100 puts dl and 1/dl within 2^+-50, and with
        // 0 <= dist < 2^34 (a sum of at most 2^15 distances below 1e8) total^2 within [2^-100, 2^69]; every
        // operation below is then the fast path of its correctly rounded intrinsic, bit for bit.
        const float dl = sqrt_inrange(dl2);
        const float inv_dl = rcp_inrange(dl);
        const float cos_t = fmaxf(0.0f, dot2(enx, eny, tlx * inv_dl, tly * inv_dl));
        const float total = r.dist + dl;
        c.total = total;
        c.geo = cos_t * 0.5f;
        c.inv = rcp_inrange(total * total);
        c.nee_e = ((r.energy * c.keep) * c.geo) * c.inv;
        c.nee_candidate = c.nee_e > 1e-5f;
        c.want_shadow = c.nee_candidate;
        if (c.want_shadow) {
            const float sx = rar_fma(wnx, kEps, r.px), sy = rar_fma(wny, kEps, r.py);
            c.shadow.sx = sx;  // make_shadow_ray with the reciprocal already at hand
            c.shadow.sy = sy;
            c.shadow.dx = (lx - sx) * inv_dl;
            c.shadow.ndy = -((ly - sy) * inv_dl);
            c.shadow.lim = dl - 0.1f;
            c.nee_t = r.time + div_with_rcp(dl, sp->c, sp->inv_c);
        }
        return;
    }
    const float dl = rar_sqrt(dl2);
    const float inv_dl = rar_rcp(dl);  // toList / distList := toList * (1 / distList)
    const float cos_t = fmaxf(0.0f, dot2(enx, eny, tlx * inv_dl, tly * inv_dl));
    const float total = r.dist + dl;
    c.total = total;
    c.geo = cos_t * 0.5f;
    c.inv = rar_rcp(total * total);
    c.nee_e = ((r.energy * c.keep) * c.geo) * c.inv;
    c.nee_candidate = c.nee_e > 1e-5f;
    c.want_shadow = (COUNT && !p.count_executed) ? 1 : c.nee_candidate;
    if (c.want_shadow) {
        c.shadow = make_shadow_ray(rar_fma(wnx, kEps, r.px), rar_fma(wny, kEps, r.py), lx, ly, dl);
        c.nee_t = r.time + rar_div(dl, p.speed_of_sound);
    }
}

// :113-118: the estimate arrives iff it cleared the threshold and the listener is visible.
template <int BANDS, bool COUNT>
RAR_HD void nee_arrival(const RayState<BANDS> &r, const BounceCtx<BANDS> &c, bool visible, Arrival<BANDS> &nee,
                        RayCounters *ctr, const RayConsts *p = nullptr) {
    nee.has = 0;
    if (c.nee_candidate && visible) {
        nee.has = 1;
        nee.hx = r.px;
        nee.hy = r.py;
        nee.t = c.nee_t;
        nee.e = c.nee_e;
        if (BANDS > 1) {
#pragma unroll
            for (int b = 0; b < BANDS; b++) {
                nee.band_e[b] = ((r.band_e[b] * c.band_keep[b]) * c.geo) * c.inv;
                if (p && p->air_on) nee.band_e[b] *= exp_neg_poly(p->air[b], c.total);
            }
        }
        if (COUNT) ctr->nee_hits += 1;
    }
}

// :121-154: absorb, then transmit or reflect.  Returns false when the ray ended (energy below 1e-3, :122).
// OPAQUE: the caller guarantees that no wall of the scene has transmission > 0.  `rngVal < transmission` (:131)
// is then false for every draw (random() >= 0), so the whole transmit/refract branch is compiled out -- the draw
// itself still happens.  Scenes of opaque walls are the common case and the kernel is a third smaller without
// that branch (config 2: 0.607 -> 0.56 ms).
// SPEC: one range test on |m|^2 covers the square root and the reciprocal of the final normalisation.
template <int BANDS, bool OPAQUE = false, bool SPEC = false>
RAR_HD bool bounce_scatter(const RayConsts &p, RayState<BANDS> &r, const BounceCtx<BANDS> &c) {
    static_assert(!SPEC || OPAQUE, "SPEC needs the ray's speed to be launch-invariant: opaque scenes only");
    r.energy *= c.keep;  // :121-122
    if (BANDS > 1) {
#pragma unroll
        for (int b = 0; b < BANDS; b++) r.band_e[b] *= c.band_keep[b];
    }
    if (r.energy < 1e-3f) return false;

    const float wnx = c.m0.x, wny = c.m0.y;
    const bool entering = c.dir_dot_n < 0.0f;  // :124-128
    const float nx = entering ? wnx : -wnx, ny = entering ? wny : -wny;
    const float rng_val = pcg_random(r.rng);  // :129 (always drawn)

    if (!OPAQUE && rng_val < c.m1.x) {  // :131-147
        // wallSpeed / nextSpeed / eta (:126-128) are only consumed here; evaluating them lazily gives the
        // same values.
        const float wall_speed = rar_div(p.speed_of_sound, c.m1.y);
        const float next_speed = entering ? wall_speed : ((r.wall_depth <= 1) ? p.speed_of_sound : wall_speed);
        const float eta = rar_div(next_speed, r.speed);
        float rx, ry;
        refract2(r.dx, r.dy, nx, ny, eta, rx, ry);
        if (rar_sqrt(dot2(rx, ry, rx, ry)) > 0.0f) {
            if (c.m0.w > 0.0f) {
                float jitter = (pcg_random(r.rng) - 0.5f) * 2.0f * c.m0.w;
                float sn, cs;
                sincos_poly(jitter, sn, cs);
                float jx = rar_fma(rx, cs, -(ry * sn));
                float jy = rar_fma(rx, sn, ry * cs);
                rx = jx;
                ry = jy;
            }
            float inv = rar_rcp(rar_sqrt(dot2(rx, ry, rx, ry)));
            r.dx = rx * inv;
            r.dy = ry * inv;
            r.speed = next_speed;
            if (entering) r.wall_depth++;
            else r.wall_depth = r.wall_depth - 1 > 0 ? r.wall_depth - 1 : 0;
            r.px = rar_fma(r.dx, kEps, r.px);
            r.py = rar_fma(r.dy, kEps, r.py);
            return true;
        }
    }

    // :149-154.  dot(dir, n) with the flipped normal is exactly +-dot(dir, wall.normal) (negation is exact).
    const float k2 = 2.0f * (entering ? c.dir_dot_n : -c.dir_dot_n);
    const float spx = rar_fma(-k2, nx, r.dx), spy = rar_fma(-k2, ny, r.dy);
    const float u = rar_fma(2.0f, pcg_random(r.rng), -1.0f);  // always drawn (:150)
    float mx, my;
    if (c.m0.w == 0.0f && spx != 0.0f && spy != 0.0f) {
        // lerp(spec, diff, 0) = fma(0, diff - spec, spec) = spec exactly when spec != 0 (0 * finite = +-0 and
        // x + +-0 = x); only a zero component could pick up the sign of diff - spec, so that case takes
        // the general path below.
        mx = spx;
        my = spy;
    } else {
        const float ang = asin_poly(u);
        float sn, cs;
        sincos_poly(ang, sn, cs);
        const float dfx = rar_fma(nx, cs, -(ny * sn));
        const float dfy = rar_fma(nx, sn, ny * cs);
        mx = rar_fma(c.m0.w, dfx - spx, spx);
        my = rar_fma(c.m0.w, dfy - spy, spy);
    }
    const float m2 = dot2(mx, my, mx, my);
    const float inv = (SPEC && __builtin_expect(in_safe_range(m2), 1)) ? rcp_inrange(sqrt_inrange(m2)) : rar_rcp(rar_sqrt(m2));
    r.dx = mx * inv;
    r.dy = my * inv;
    r.px = rar_fma(nx, kEps, r.px);
    r.py = rar_fma(ny, kEps, r.py);
    return true;
}

// Single-listener composition of the pieces, split around the shadow phase.
// Returns false when the ray ended without hitting a wall (:86-90).
template <int BANDS, bool COUNT, class Scene>
RAR_HD bool bounce_begin(const Scene &sc, const RayConsts &p, RayState<BANDS> &r, Arrival<BANDS> &direct,
                         BounceCtx<BANDS> &c, RayCounters *ctr, f4 *dbg = nullptr, int dbg_flags = 0,
                         const SpecConsts *sp = nullptr) {
    c.want_shadow = 0;
    c.nee_candidate = 0;
    bounce_nearest<BANDS, COUNT>(sc, r, c, ctr);
    listener_direct<BANDS, COUNT, Scene::kSpec>(p, p.listener_x, p.listener_y, r, c.closest, direct, ctr, sp);
    if (!bounce_advance(sc, r, c, dbg, dbg_flags, sp)) return false;
    listener_nee<BANDS, COUNT, Scene::kSpec>(p, p.listener_x, p.listener_y, r, c, sp);
    return true;
}

// `visible` is the outcome of the shadow phase (ignored unless c.want_shadow).  Returns false when the ray
// ended (energy below 1e-3, :122).
template <int BANDS, bool COUNT, bool OPAQUE = false, class Scene>
RAR_HD bool bounce_finish(const Scene &sc, const RayConsts &p, RayState<BANDS> &r, Arrival<BANDS> &nee,
                          const BounceCtx<BANDS> &c, bool visible, RayCounters *ctr) {
    (void)sc;
    nee_arrival<BANDS, COUNT>(r, c, visible, nee, ctr, &p);
    return bounce_scatter<BANDS, OPAQUE, Scene::kSpec>(p, r, c);
}

// The three phases with a per-thread shadow walk: what a single thread of the reference does.
template <int BANDS, bool COUNT, bool OPAQUE = false, class Scene>
RAR_HD bool ray_bounce(const Scene &sc, const RayConsts &p, RayState<BANDS> &r, Arrival<BANDS> &direct,
                       Arrival<BANDS> &nee, RayCounters *ctr, f4 *dbg = nullptr, int dbg_flags = 0,
                       const SpecConsts *sp = nullptr) {
    BounceCtx<BANDS> c;
    nee.has = 0;
    if (!bounce_begin<BANDS, COUNT>(sc, p, r, direct, c, ctr, dbg, dbg_flags, sp)) return false;
    bool visible = true;
    if (c.want_shadow) {
        int tests = 0;
        visible = check_vis(sc, c.shadow, COUNT ? &tests : nullptr);
        if (COUNT) ctr->shadow_tests += (unsigned long long)tests;
    }
    return bounce_finish<BANDS, COUNT, OPAQUE>(sc, p, r, nee, c, visible, ctr);
}

}  // namespace rar
