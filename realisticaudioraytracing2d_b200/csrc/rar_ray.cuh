// rar_ray.cuh -- one ray of Raytrace2D.compute `Trace` (:49-156) as a per-bounce state machine.
//
// ray_init() + ray_bounce() are written once and used by the CUDA kernel (trace_kernel.cu); they
// also compile as host C++ for tests/host_emulation.cpp.  A bounce returns up to two pending
// arrivals (direct listener crossing, next-event estimate); the caller deposits them, which lets the
// kernel do that with warp-convergent aggregation.
//
// The inner loops over walls do NOT evaluate Common.hlsl:14-21 `intersect` literally.  They run a
// division-free conservative filter and evaluate the literal formula (two correctly rounded
// divisions and the original comparisons) only for the survivors, so the accepted hits and their
// distances are bit-identical to the literal evaluation:
//   literal: dotP = dot(v2,v3); t1 = cross(v2,v1)/dotP; t2 = dot(v1,v3)/dotP;
//            hit iff |dotP| >= eps, t1 >= eps, 0 <= t2 <= 1; nearest loop also needs t1 < closest.
//   filter : with num1 = cross(v2,v1), num2 = dot(v1,v3), r = closest*(1+2^-20)*dotP
//            pass iff |2*num2 - dotP| <= |dotP|   (<=> num2/dotP in [0,1], incl. num2 = -0)
//                 and |2*num1 - r|    <= |r|      (<=> num1/dotP in [0, closest*(1+2^-20)])
//   For binary32 a, b>0: a > b implies fl(a/b) > 1 (a/b > 1+2^-24), and rounding is monotone, so every
//   hit the literal test accepts passes the filter.  The one literal-accepting case the filter does
//   not see is t2 = num2/dotP underflowing to -0 for opposite signs, which needs |num2| < 2^-100 for
//   any scene with coordinates below 2^50; HLSL flushes such values anyway.  ("filter".)
#pragma once

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
};

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
RAR_HD void ray_init(RayState<BANDS> &r, uint32_t id, const RayConsts &p) {
    r.rng = id + p.rng_state_offset * 719393u;
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

// Literal Common.hlsl:14-21 on the precomputed edge e = b - a; used on filter survivors only.
RAR_HD float intersect_exact(float num1, float num2, float dotP) {
    if (fabsf(dotP) < kEps) return kInf;
    float t1 = rar_div(num1, dotP);
    float t2 = rar_div(num2, dotP);
    return (t1 >= kEps && t2 >= 0.0f && t2 <= 1.0f) ? t1 : kInf;
}

constexpr float kSlack = 1.00000095367431640625f;  // 1 + 2^-20

// Raytrace2D.compute:69-72: nearest hit over all walls, lowest index wins ties.
template <class Scene>
RAR_HD void nearest_hit(const Scene &sc, float ox, float oy, float dx, float dy, float &closest_out, int &hit_out) {
    float closest = kInf, closest_m = kInf * kSlack;
    int hit = -1;
    const float ndy = -dy;
    const int n = sc.n_walls();
#pragma unroll 4
    for (int w = 0; w < n; w++) {
        const f4 s = sc.geo(w);
        float v1x = ox - s.x, v1y = oy - s.y;
        float dotP = rar_fma(s.z, ndy, s.w * dx);
        float num2 = rar_fma(v1x, ndy, v1y * dx);
        float num1 = rar_fma(s.z, v1y, -(s.w * v1x));
        float r = closest_m * dotP;
        bool pass = (fabsf(rar_fma(2.0f, num2, -dotP)) <= fabsf(dotP)) & (fabsf(rar_fma(2.0f, num1, -r)) <= fabsf(r));
        if (pass) {
            float d = intersect_exact(num1, num2, dotP);
            if (d < closest) {
                closest = d;
                closest_m = d * kSlack;
                hit = w;
            }
        }
    }
    closest_out = closest;
    hit_out = hit;
}

// Raytrace2D.compute:40-47 checkVis.  Returns true when visible; *tests receives the number of
// intersect() evaluations the reference's early-exit loop performs.
template <class Scene>
RAR_HD bool check_vis(const Scene &sc, float sx, float sy, float ex, float ey, float dist, int *tests) {
    float dx = rar_div(ex - sx, dist), dy = rar_div(ey - sy, dist);
    const float lim = dist - 0.1f;
    const int n = sc.n_walls();
    if (n > 0 && kInf < lim) {  // every intersect() result (<= inf) is < lim: blocked by wall 0
        if (tests) *tests = 1;
        return false;
    }
    const float lim_m = lim * kSlack;
    const float ndy = -dy;
    int w = 0;
    bool blocked = false;
#pragma unroll 4
    for (; w < n; w++) {
        const f4 s = sc.geo(w);
        float v1x = sx - s.x, v1y = sy - s.y;
        float dotP = rar_fma(s.z, ndy, s.w * dx);
        float num2 = rar_fma(v1x, ndy, v1y * dx);
        float num1 = rar_fma(s.z, v1y, -(s.w * v1x));
        float r = lim_m * dotP;
        bool pass = (fabsf(rar_fma(2.0f, num2, -dotP)) <= fabsf(dotP)) & (fabsf(rar_fma(2.0f, num1, -r)) <= fabsf(r));
        if (pass) {
            float d = intersect_exact(num1, num2, dotP);
            if (d < lim) { blocked = true; break; }
        }
    }
    if (tests) *tests = blocked ? w + 1 : n;
    return !blocked;
}

// One iteration of the bounce loop, Raytrace2D.compute:66-155.  Returns false when the ray ended.
// dbg_hit / dbg_miss: where to record this bounce's vertex for the debugRays buffer (:87-88, :96-97), or
// nullptr.
template <int BANDS, bool COUNT, class Scene>
RAR_HD bool ray_bounce(const Scene &sc, const RayConsts &p, RayState<BANDS> &r, Arrival<BANDS> &direct,
                       Arrival<BANDS> &nee, RayCounters *ctr, f4 *dbg_hit = nullptr, f4 *dbg_miss = nullptr) {
    direct.has = 0;
    nee.has = 0;
    float closest;
    int hit;
    nearest_hit(sc, r.px, r.py, r.dx, r.dy, closest, hit);  // :69-72
    if (COUNT) {
        ctr->ray_bounces += 1;
        ctr->nearest_tests += (unsigned long long)sc.n_walls();
    }

    if (r.wall_depth == 0) {  // :74-84
        float dl = intersect_circle(r.px, r.py, r.dx, r.dy, p.listener_x, p.listener_y, p.listener_radius);
        if (dl < closest && dl < kInf) {
            direct.has = 1;
            direct.hx = rar_fma(r.dx, dl, r.px);
            direct.hy = rar_fma(r.dy, dl, r.py);
            direct.t = r.time + rar_div(dl, r.speed);
            float total = r.dist + dl;
            float denom = fmaxf(1.0f, total * total);
            direct.e = rar_div(r.energy, denom);
            if (BANDS > 1) {
#pragma unroll
                for (int b = 0; b < BANDS; b++) direct.band_e[b] = rar_div(r.band_e[b], denom);
            }
            if (COUNT) ctr->direct_hits += 1;
        }
    }
    if (hit < 0) {  // :86-90
        if (dbg_miss) *dbg_miss = f4{rar_fma(r.dx, 20.0f, r.px), rar_fma(r.dy, 20.0f, r.py), 0.0f, 0.0f};
        return false;
    }

    r.px = rar_fma(r.dx, closest, r.px);  // :92-94
    r.py = rar_fma(r.dy, closest, r.py);
    r.time += rar_div(closest, r.speed);
    r.dist += closest;
    if (dbg_hit) *dbg_hit = f4{r.px, r.py, r.energy, 0.0f};  // :96-97

    const f4 m0 = sc.mat0(hit);  // nx, ny, absorption, scattering   (:99)
    const f2 m1 = sc.mat1(hit);  // transmission, ior
    const float wnx = m0.x, wny = m0.y;
    const float keep = 1.0f - m0.z;
    float band_keep[BANDS > 1 ? BANDS : 1];
    if (BANDS > 1) {
        const float *ba = sc.band_abs(hit);
#pragma unroll
        for (int b = 0; b < BANDS; b++) band_keep[b] = 1.0f - ba[b];
    }
    const float dir_dot_n = dot2(r.dx, r.dy, wnx, wny);

    if (r.wall_depth == 0) {  // :101-119
        float tlx = p.listener_x - r.px, tly = p.listener_y - r.py;
        float dl = rar_sqrt(dot2(tlx, tly, tlx, tly));
        float sx = rar_fma(wnx, kEps, r.px), sy = rar_fma(wny, kEps, r.py);
        int tests = 0;
        bool vis = check_vis(sc, sx, sy, p.listener_x, p.listener_y, dl, COUNT ? &tests : nullptr);
        if (COUNT) ctr->shadow_tests += (unsigned long long)tests;
        if (vis) {
            bool flip = dir_dot_n > 0.0f;
            float enx = flip ? -wnx : wnx, eny = flip ? -wny : wny;
            float cos_t = fmaxf(0.0f, dot2(enx, eny, rar_div(tlx, dl), rar_div(tly, dl)));
            float total = r.dist + dl;
            float geo = cos_t * 0.5f;
            float inv = rar_div(1.0f, total * total);
            float contrib = ((r.energy * keep) * geo) * inv;
            if (contrib > 1e-5f) {
                nee.has = 1;
                nee.hx = r.px;
                nee.hy = r.py;
                nee.t = r.time + rar_div(dl, p.speed_of_sound);
                nee.e = contrib;
                if (BANDS > 1) {
#pragma unroll
                    for (int b = 0; b < BANDS; b++) nee.band_e[b] = ((r.band_e[b] * band_keep[b]) * geo) * inv;
                }
                if (COUNT) ctr->nee_hits += 1;
            }
        }
    }

    r.energy *= keep;  // :121-122
    if (BANDS > 1) {
#pragma unroll
        for (int b = 0; b < BANDS; b++) r.band_e[b] *= band_keep[b];
    }
    if (r.energy < 1e-3f) return false;

    const bool entering = dir_dot_n < 0.0f;  // :124-128
    const float nx = entering ? wnx : -wnx, ny = entering ? wny : -wny;
    const float wall_speed = rar_div(p.speed_of_sound, m1.y);
    const float next_speed = entering ? wall_speed : ((r.wall_depth <= 1) ? p.speed_of_sound : wall_speed);
    const float eta = rar_div(next_speed, r.speed);
    const float rng_val = pcg_random(r.rng);  // :129

    if (rng_val < m1.x) {  // :131-147
        float rx, ry;
        refract2(r.dx, r.dy, nx, ny, eta, rx, ry);
        if (rar_sqrt(dot2(rx, ry, rx, ry)) > 0.0f) {
            if (m0.w > 0.0f) {
                float jitter = (pcg_random(r.rng) - 0.5f) * 2.0f * m0.w;
                float s, c;
                sincos_poly(jitter, s, c);
                float jx = rar_fma(rx, c, -(ry * s));
                float jy = rar_fma(rx, s, ry * c);
                rx = jx;
                ry = jy;
            }
            float inv = rar_div(1.0f, rar_sqrt(dot2(rx, ry, rx, ry)));
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

    // :149-154.  dir_dot_n with the flipped normal is exactly +-dir_dot_n (negation is exact).
    const float k2 = 2.0f * (entering ? dir_dot_n : -dir_dot_n);
    const float spx = rar_fma(-k2, nx, r.dx), spy = rar_fma(-k2, ny, r.dy);
    const float u = rar_fma(2.0f, pcg_random(r.rng), -1.0f);
    const float ang = asin_poly(u);
    float s, c;
    sincos_poly(ang, s, c);
    const float dfx = rar_fma(nx, c, -(ny * s));
    const float dfy = rar_fma(nx, s, ny * c);
    const float mx = rar_fma(m0.w, dfx - spx, spx);
    const float my = rar_fma(m0.w, dfy - spy, spy);
    const float inv = rar_div(1.0f, rar_sqrt(dot2(mx, my, mx, my)));
    r.dx = mx * inv;
    r.dy = my * inv;
    r.px = rar_fma(nx, kEps, r.px);
    r.py = rar_fma(ny, kEps, r.py);
    return true;
}

}  // namespace rar
