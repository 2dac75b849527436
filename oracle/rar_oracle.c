/*
 * rar_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the reference's hot path, written from the reference's
 * shader/C# text (file:line cited per function, paths relative to
 * /root/reference/Assets/Script/).  It exists only so that tests/, the smoke test
 * and bench.py's cpu_baseline leg can CHECK and TIME-BESIDE the CUDA path.  Nothing
 * in the product package may import, link or call it.
 *
 * PARITY STATUS: **parity unpinned**.  The reference ships no tests, no golden
 * vectors and no CPU implementation (it is HLSL compute run by Unity, which cannot
 * execute in this image: no Unity, dotnet, mono or HLSL compiler).  The only
 * known-answer values available are the PRNG values derived by hand from
 * Common.hlsl:8-12, the scene fixtures of Appendix B.1-B.2
 * and analytic properties (image-source arrival times, 1/d^2 energies).  This
 * oracle is checked against those; beyond that it DEFINES the semantics.
 *
 * ARITHMETIC CONTRACT (what "bit-exact" means; the CUDA kernels obey the same one).
 *   - every operation is IEEE-754 binary32, round-to-nearest-even, no flush-to-zero;
 *   - no implicit contraction (build with -ffp-contract=off); a fused multiply-add
 *     happens exactly where fmaf() is written below and nowhere else;
 *   - dot(a,b)   := fmaf(a.x, b.x, a.y*b.y)
 *     cross(a,b) := fmaf(a.x, b.y, -(a.y*b.x))
 *     p + d*t    := fmaf(d, t, p)
 *     length(v)  := sqrtf(dot(v,v));  normalize(v) := v * (1.0f / length(v))
 *     v / s      := v * (1.0f / s) for a 2-vector v and a scalar s (the reading of HLSL's
 *                   vector/scalar division that GPU compilers emit: one reciprocal, two multiplies)
 *     reflect(i,n) := i - (2*dot(i,n))*n  evaluated as fmaf(-(2*dot), n, i)
 *     lerp(a,b,s)  := fmaf(s, b-a, a)
 *   - sin/cos/asin are the fixed polynomial kernels orc_sincosf/orc_asinf below
 *     (HLSL's own intrinsics are implementation-approximate, so the reference does
 *     not define these bits; this file does);
 *   - uint -> float is round-to-nearest; float -> int truncates toward zero;
 *   - energies are deposited as signed Q23.40 fixed point:
 *       q(e) = (int64) (clamp(e, -2^22, 2^22) * 2^40)   (NaN deposits nothing).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fno-fast-math -mfma -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ---- the contract-sensitivity variant (tests only) --------------------------------------------------------------
 * Built with -DORC_VARIANT_NAIVE (oracle/Makefile target `naive`) this same file becomes a second oracle that makes
 * the OTHER defensible reading of every point where the arithmetic contract had to choose, i.e. what a different
 * HLSL compiler might legitimately have emitted:
 *   - no fused multiply-add anywhere: fmaf(a,b,c) := (a*b) + c with two roundings;
 *   - sin / cos / asin from libm instead of the fixed polynomial kernels;
 *   - vector / scalar as true divisions instead of one reciprocal and multiplies (normalize, checkVis direction,
 *     toList / distList, the 1 / totalD^2 factor).
 * tests/test_oracle_sensitivity.py traces the bundled rooms and the shoebox with both and shows that the impulse
 * responses differ by far less than the frame-to-frame Monte-Carlo spread of either: the unpinned choices of the
 * contract do not move the result a listener hears. */
#ifdef ORC_VARIANT_NAIVE
#undef fmaf
#define fmaf(a, b, c) orc_naive_fma((a), (b), (c))
static inline float orc_naive_fma(float a, float b, float c) {
    volatile float p = a * b; /* rounded product, then a rounded sum */
    return p + c;
}
/* x / s for a component x of a vector divided by the scalar s (inv = 1/s is the contract's reading) */
#define ORC_VDIV(x, s, inv) ((x) / (s))
#else
#define ORC_VDIV(x, s, inv) ((x) * (inv))
#endif

/* Common.hlsl:4-6 */
static const float ORC_EPS = 1e-4f;
static const float ORC_INF = 1e8f;
static const float ORC_PI = 3.14159265f;

/* ---- wall / hit / parameter layouts -------------------------------------------------- */

/* Helpers/SceneHelper.cs:8-22 (Segment, LayoutKind.Sequential) == Raytrace2D.compute:12-22 (Wall). 40 bytes. */
typedef struct {
    float ax, ay, bx, by, nx, ny;
    float absorption, scattering, transmission, ior;
} orc_segment;

/* Raytrace2D.compute:24-28 (RayInfo), 16 bytes, plus a key so unordered GPU hits can be matched. */
typedef struct {
    float time_delay, energy, hit_x, hit_y;
    uint32_t ray;
    uint16_t bounce;
    uint16_t kind; /* 0 = direct listener crossing (:74-84), 1 = next-event estimation (:101-119) */
} orc_hit;

/* Uniforms of Trace / ProcessHits: Raytrace2D.compute:5-10,36 and RayTraceManager.cs:191-201,227-228. */
typedef struct {
    float source_x, source_y, listener_x, listener_y;
    float listener_radius, speed_of_sound, input_gain;
    int32_t max_bounce_count;
    uint32_t rng_state_offset;
    int32_t ray_count;       /* `rayCount` uniform: the angle denominator */
    int32_t debug_ray_count; /* unused by the oracle's numeric path */
    int32_t sample_rate;
    int32_t impulse_length; /* number of TIME bins */
    int32_t bands;          /* 1 = broadband (live variant); >1 = banded layout IR[bin*bands+band] */
    float time_divisor;     /* 1 => bin=(int)(t*SR); W => (int)(t*SR/W) (RaytraceOcclusion2D.compute:241) */
    uint32_t flags;         /* bit0: exact ray range (no round-up of the dispatch to 64 threads) */
    int64_t ray_begin, ray_end; /* thread-id range traced by this call; (0,0) => whole dispatch */
} orc_trace_params;

typedef struct {
    uint64_t ray_bounces;   /* bounce-loop iterations executed (Raytrace2D.compute:66) */
    uint64_t nearest_tests; /* intersect() calls from the nearest-hit loop (:69-72) = ray_bounces*numWalls */
    uint64_t shadow_tests;  /* intersect() calls from checkVis (:40-47), early-exit aware */
    uint64_t direct_hits, nee_hits;
} orc_counters;

/* ---- Common.hlsl ------------------------------------------------------------------------ */

/* Common.hlsl:8-12. `4294967295.0` is not representable in binary32 and rounds to 2^32, so the
 * division is an exact scaling by 2^-32; (float)uint rounds to nearest, so 1.0f is reachable. */
ORC_API float orc_random(uint32_t *state) {
    uint32_t s = *state * 747796405u + 2891336453u;
    *state = s;
    uint32_t res = ((s >> ((s >> 28) + 4u)) ^ s) * 277803737u;
    uint32_t v = (res >> 22) ^ res;
    return (float)v / 4294967296.0f;
}

static inline float dot2(float ax, float ay, float bx, float by) { return fmaf(ax, bx, ay * by); }

/* Common.hlsl:14-21 */
ORC_API float orc_intersect(float ox, float oy, float dx, float dy, float ax, float ay, float bx, float by) {
    float v1x = ox - ax, v1y = oy - ay;
    float v2x = bx - ax, v2y = by - ay;
    float v3x = -dy, v3y = dx;
    float dotP = dot2(v2x, v2y, v3x, v3y);
    if (fabsf(dotP) < ORC_EPS) return ORC_INF;
    float t1 = fmaf(v2x, v1y, -(v2y * v1x)) / dotP;
    float t2 = dot2(v1x, v1y, v3x, v3y) / dotP;
    return (t1 >= ORC_EPS && t2 >= 0.0f && t2 <= 1.0f) ? t1 : ORC_INF;
}

/* Common.hlsl:23-36 */
ORC_API float orc_intersect_circle(float px, float py, float dx, float dy, float cx, float cy, float radius) {
    float Lx = cx - px, Ly = cy - py;
    float tca = dot2(Lx, Ly, dx, dy);
    if (tca < 0.0f) return ORC_INF;
    float d2 = fmaf(-tca, tca, dot2(Lx, Ly, Lx, Ly));
    float r2 = radius * radius;
    if (d2 > r2) return ORC_INF;
    float thc = sqrtf(r2 - d2);
    float t0 = tca - thc;
    float t1 = tca + thc;
    if (t0 > ORC_EPS) return t0;
    if (t1 > ORC_EPS) return t1;
    return ORC_INF;
}

/* Common.hlsl:38-43 (2-D: the z components are zero throughout). Returns 0 on total internal reflection. */
ORC_API int orc_refract(float ix, float iy, float nx, float ny, float eta, float *tx, float *ty) {
    float cosi = dot2(-ix, -iy, nx, ny);
    float cost2 = 1.0f - (eta * eta) * (1.0f - cosi * cosi);
    float k = eta * cosi - sqrtf(fabsf(cost2));
    float rx = fmaf(k, nx, eta * ix);
    float ry = fmaf(k, ny, eta * iy);
    if (cost2 > 0.0f) { *tx = rx; *ty = ry; return 1; }
    *tx = 0.0f; *ty = 0.0f;
    return 0;
}

/* ---- fixed transcendental kernels (part of the arithmetic contract) ---------------------- */

/* sin and cos of x for |x| < ~1e4: quadrant by the 1.5*2^23 rounding trick (pure adds, so identical on
 * any IEEE machine), 3-term Cody-Waite reduction by pi/2, then the classic single-precision minimax
 * polynomials on [-pi/4, pi/4]. */
ORC_API void orc_sincosf(float x, float *sn, float *cs) {
#ifdef ORC_VARIANT_NAIVE
    *sn = sinf(x);
    *cs = cosf(x);
    return;
#endif
    const float TWO_OVER_PI = 0.636619772f;
    const float MAGIC = 12582912.0f; /* 1.5 * 2^23 */
    float kf = fmaf(x, TWO_OVER_PI, MAGIC) - MAGIC;
    int q = (int)kf;
    float r = fmaf(-kf, 1.5703125f, x);              /* pi/2 high   */
    r = fmaf(-kf, 4.837512969970703125e-4f, r);      /* pi/2 middle */
    r = fmaf(-kf, 7.54978995489188e-8f, r);          /* pi/2 low    */
    float z = r * r;
    float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(z, ps, -1.6666654611e-1f);
    float s = fmaf(r * z, ps, r);
    float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(z, pc, 4.166664568298827e-2f);
    float c = fmaf(z * z, pc, fmaf(z, -0.5f, 1.0f));
    switch (q & 3) {
        case 0: *sn = s;  *cs = c;  break;
        case 1: *sn = c;  *cs = -s; break;
        case 2: *sn = -s; *cs = -c; break;
        default: *sn = -c; *cs = s; break;
    }
}

/* asin on [-1,1] (inputs outside are clamped): polynomial for |x|<=0.5, half-angle identity above. */
ORC_API float orc_asinf(float x) {
#ifdef ORC_VARIANT_NAIVE
    return asinf(x > 1.0f ? 1.0f : (x < -1.0f ? -1.0f : x));
#endif
    float a = fabsf(x);
    if (a > 1.0f) a = 1.0f;
    int big = a > 0.5f;
    float z, w;
    if (big) { z = 0.5f * (1.0f - a); w = sqrtf(z); }
    else     { w = a; z = a * a; }
    float p = fmaf(z, 4.2163199048e-2f, 2.4181311049e-2f);
    p = fmaf(z, p, 4.5470025998e-2f);
    p = fmaf(z, p, 7.4953002686e-2f);
    p = fmaf(z, p, 1.6666752422e-1f);
    float r = fmaf(w * z, p, w);
    if (big) r = 1.5707963267948966f - (r + r);
    return (x < 0.0f) ? -r : r;
}

/* exp(-alpha * d), alpha >= 0 in 1/m, d a path length in m: the air attenuation of the banded model (this build's
 * extension; the reference's placeholder is `muffle`, RaytraceOcclusion2D.compute:125-126,247-248).  Fixed kernel:
 * y = (alpha*d) * (-log2 e); k = round(y) by the 1.5*2^23 trick; 2^(y-k) by a degree-6 polynomial; times 2^k. */
ORC_API float orc_exp_neg(float alpha, float d) {
#ifdef ORC_VARIANT_NAIVE
    return expf(-(alpha * d));
#endif
    float y = (alpha * d) * -1.4426950408889634f;
    if (!(y > -126.0f)) return 0.0f;
    if (!(y < 0.0f)) return 1.0f;
    const float MAGIC = 12582912.0f;
    float kf = (y + MAGIC) - MAGIC;
    float f = y - kf;
    float p = fmaf(f, 1.5403530e-4f, 1.3333558e-3f);
    p = fmaf(f, p, 9.6181291e-3f);
    p = fmaf(f, p, 5.5504109e-2f);
    p = fmaf(f, p, 2.4022651e-1f);
    p = fmaf(f, p, 6.9314718e-1f);
    p = fmaf(f, p, 1.0f);
    uint32_t bits = (uint32_t)((int)kf + 127) << 23;
    float scale;
    memcpy(&scale, &bits, sizeof scale);
    return p * scale;
}

/* ---- deposit (Raytrace2D.compute:157-165 ProcessHits, fixed-point instead of the racy float +=) ------- */

ORC_API int64_t orc_quantize(float e) {
    if (!(e == e)) return 0;
    if (e > 4194304.0f) e = 4194304.0f;
    if (e < -4194304.0f) e = -4194304.0f;
    return (int64_t)(e * 1099511627776.0f); /* 2^40: exact scaling, then truncation toward zero */
}

/* time -> bin. Returns -1 when the hit falls outside [0, impulse_length). (int) of a value in (-1,0)
 * truncates to 0 and is accepted, exactly as `(int)(t*SampleRate)` followed by `index >= 0` does. */
ORC_API int32_t orc_time_bin(float t, int32_t sample_rate, float time_divisor, int32_t impulse_length) {
    float ts = t * (float)sample_rate;
    if (time_divisor != 1.0f) ts = ts / time_divisor;
    if (!(ts > -1.0f && ts < (float)impulse_length)) return -1;
    int32_t idx = (int32_t)ts;
    return (idx >= 0 && idx < impulse_length) ? idx : -1;
}

typedef struct {
    const orc_segment *walls;
    int n_walls;
    const float *band_abs; /* [n_walls][bands] or NULL */
    const float *air;      /* [bands] air absorption in 1/m, or NULL */
    const orc_trace_params *p;
    int64_t *hist;         /* [impulse_length][bands], accumulated atomically */
    orc_hit *hits;
    int64_t hit_cap;
    int64_t *hit_count;
} trace_env;

static void emit(const trace_env *env, uint32_t ray, int bounce, int kind, float t, float e, float hx, float hy,
                 const float *band_e, orc_counters *ctr) {
    const orc_trace_params *p = env->p;
    if (kind == 0) ctr->direct_hits++; else ctr->nee_hits++;
    if (env->hits) {
        int64_t slot = __atomic_fetch_add(env->hit_count, 1, __ATOMIC_RELAXED);
        if (slot < env->hit_cap) {
            orc_hit h = {t, e, hx, hy, ray, (uint16_t)bounce, (uint16_t)kind};
            env->hits[slot] = h;
        }
    }
    if (!env->hist) return;
    int32_t bin = orc_time_bin(t, p->sample_rate, p->time_divisor, p->impulse_length);
    if (bin < 0) return;
    if (p->bands <= 1) {
        __atomic_fetch_add(&env->hist[bin], orc_quantize(e), __ATOMIC_RELAXED);
    } else {
        for (int b = 0; b < p->bands; b++)
            __atomic_fetch_add(&env->hist[(int64_t)bin * p->bands + b], orc_quantize(band_e[b]), __ATOMIC_RELAXED);
    }
}

/* Raytrace2D.compute:40-47 */
static int check_vis(const trace_env *env, float sx, float sy, float ex, float ey, float dist, orc_counters *ctr) {
    float inv_dist = 1.0f / dist; /* (end - start) / dist as vector * (1/scalar) */
    float dx = ORC_VDIV(ex - sx, dist, inv_dist), dy = ORC_VDIV(ey - sy, dist, inv_dist);
    float lim = dist - 0.1f;
    for (int w = 0; w < env->n_walls; w++) {
        const orc_segment *s = &env->walls[w];
        ctr->shadow_tests++;
        float d = orc_intersect(sx, sy, dx, dy, s->ax, s->ay, s->bx, s->by);
        if (d < lim) return 0;
    }
    return 1;
}

#define ORC_MAX_BANDS 128

/* Raytrace2D.compute:49-156 -- one thread of Trace. Band energies are the build's extension: band b
 * carries its own energy attenuated by band_abs[wall][b]; ray life, thresholds and branching follow
 * the broadband energy exactly as the reference. */
static void trace_one(const trace_env *env, uint32_t id, orc_counters *ctr) {
    const orc_trace_params *p = env->p;
    const int nb = p->bands > 1 ? p->bands : 0;
    uint32_t rng = id + p->rng_state_offset * 719393u;                       /* :51 */
    float angle = (((float)id + orc_random(&rng)) / (float)p->ray_count) * 2.0f * ORC_PI; /* :52 */
    float dirx, diry;
    orc_sincosf(angle, &diry, &dirx);                                          /* :54 */
    float posx = p->source_x, posy = p->source_y;
    float energy = p->input_gain, time = 0.0f, dist = 0.0f;
    float speed = p->speed_of_sound;
    int wall_depth = 0;
    float band_e[ORC_MAX_BANDS], band_out[ORC_MAX_BANDS];
    for (int b = 0; b < nb; b++) band_e[b] = p->input_gain;

    for (int i = 0; i < p->max_bounce_count; i++) {                          /* :66 */
        ctr->ray_bounces++;
        float closest = ORC_INF; int hit = -1;
        for (int w = 0; w < env->n_walls; w++) {                             /* :69-72 */
            const orc_segment *s = &env->walls[w];
            float d = orc_intersect(posx, posy, dirx, diry, s->ax, s->ay, s->bx, s->by);
            if (d < closest) { closest = d; hit = w; }
        }
        ctr->nearest_tests += (uint64_t)env->n_walls;

        if (wall_depth == 0) {                                               /* :74-84 */
            float dl = orc_intersect_circle(posx, posy, dirx, diry, p->listener_x, p->listener_y, p->listener_radius);
            if (dl < closest && dl < ORC_INF) {
                float hx = fmaf(dirx, dl, posx), hy = fmaf(diry, dl, posy);
                float t = time + dl / speed;
                float total = dist + dl;
                float denom = fmaxf(1.0f, total * total);
                float e = energy / denom;
                for (int b = 0; b < nb; b++) {
                    band_out[b] = band_e[b] / denom;
                    if (env->air) band_out[b] *= orc_exp_neg(env->air[b], total);
                }
                emit(env, id, i, 0, t, e, hx, hy, band_out, ctr);
            }
        }
        if (hit < 0) break;                                                  /* :86-90 */

        posx = fmaf(dirx, closest, posx);                                    /* :92-94 */
        posy = fmaf(diry, closest, posy);
        time += closest / speed;
        dist += closest;

        const orc_segment *wall = &env->walls[hit];                          /* :99 */
        const float *wabs = nb ? env->band_abs + (size_t)hit * nb : NULL;
        float keep = 1.0f - wall->absorption;

        if (wall_depth == 0) {                                               /* :101-119 */
            float tlx = p->listener_x - posx, tly = p->listener_y - posy;
            float dl = sqrtf(dot2(tlx, tly, tlx, tly));
            float sx = fmaf(wall->nx, ORC_EPS, posx), sy = fmaf(wall->ny, ORC_EPS, posy);
            if (check_vis(env, sx, sy, p->listener_x, p->listener_y, dl, ctr)) {
                int flip = dot2(dirx, diry, wall->nx, wall->ny) > 0.0f;
                float enx = flip ? -wall->nx : wall->nx, eny = flip ? -wall->ny : wall->ny;
                float inv_dl = 1.0f / dl; /* toList / distList as vector * (1/scalar) */
                float cos_t = fmaxf(0.0f, dot2(enx, eny, ORC_VDIV(tlx, dl, inv_dl), ORC_VDIV(tly, dl, inv_dl)));
                float total = dist + dl;
                float geo = (cos_t * 0.5f);
                float inv = 1.0f / (total * total);
                float contrib = ORC_VDIV((energy * keep) * geo, total * total, inv);
                if (contrib > 1e-5f) {
                    float t = time + dl / p->speed_of_sound;
                    for (int b = 0; b < nb; b++) {
                        band_out[b] = ORC_VDIV((band_e[b] * (1.0f - wabs[b])) * geo, total * total, inv);
                        if (env->air) band_out[b] *= orc_exp_neg(env->air[b], total);
                    }
                    emit(env, id, i, 1, t, contrib, posx, posy, band_out, ctr);
                }
            }
        }

        energy *= keep;                                                      /* :121-122 */
        for (int b = 0; b < nb; b++) band_e[b] *= (1.0f - wabs[b]);
        if (energy < 1e-3f) break;

        int entering = dot2(dirx, diry, wall->nx, wall->ny) < 0.0f;          /* :124-128 */
        float nx = entering ? wall->nx : -wall->nx, ny = entering ? wall->ny : -wall->ny;
        float wall_speed = p->speed_of_sound / wall->ior;
        float next_speed = entering ? wall_speed : ((wall_depth <= 1) ? p->speed_of_sound : wall_speed);
        float eta = next_speed / speed;
        float rng_val = orc_random(&rng);                                    /* :129 */

        if (rng_val < wall->transmission) {                                  /* :131-147 */
            float rx, ry;
            orc_refract(dirx, diry, nx, ny, eta, &rx, &ry);
            if (sqrtf(dot2(rx, ry, rx, ry)) > 0.0f) {
                if (wall->scattering > 0.0f) {
                    float jitter = (orc_random(&rng) - 0.5f) * 2.0f * wall->scattering;
                    float s, c;
                    orc_sincosf(jitter, &s, &c);
                    float jx = fmaf(rx, c, -(ry * s));
                    float jy = fmaf(rx, s, ry * c);
                    rx = jx; ry = jy;
                }
                float len = sqrtf(dot2(rx, ry, rx, ry));
                float inv = 1.0f / len;
                dirx = ORC_VDIV(rx, len, inv); diry = ORC_VDIV(ry, len, inv);
                speed = next_speed;
                if (entering) wall_depth++; else wall_depth = wall_depth - 1 > 0 ? wall_depth - 1 : 0;
                posx = fmaf(dirx, ORC_EPS, posx);
                posy = fmaf(diry, ORC_EPS, posy);
                continue;
            }
        }

        float k2 = 2.0f * dot2(dirx, diry, nx, ny);                          /* :149-154 */
        float spx = fmaf(-k2, nx, dirx), spy = fmaf(-k2, ny, diry);
        float u = fmaf(2.0f, orc_random(&rng), -1.0f);
        float ang = orc_asinf(u);
        float s, c;
        orc_sincosf(ang, &s, &c);
        float dfx = fmaf(nx, c, -(ny * s));
        float dfy = fmaf(nx, s, ny * c);
        float mx = fmaf(wall->scattering, dfx - spx, spx);
        float my = fmaf(wall->scattering, dfy - spy, spy);
        float len = sqrtf(dot2(mx, my, mx, my));
        float inv = 1.0f / len;
        dirx = ORC_VDIV(mx, len, inv); diry = ORC_VDIV(my, len, inv);
        posx = fmaf(nx, ORC_EPS, posx);
        posy = fmaf(ny, ORC_EPS, posy);
    }
}

/* Dispatch of Trace (RayTraceManager.cs:205 + Helpers/ComputeHelper.cs:27-31): ceil(rayCount/64) groups of
 * 64 threads and no `id < rayCount` guard, so ceil(rayCount/64)*64 rays are traced (flag bit0 disables). */
ORC_API int64_t orc_dispatch_threads(const orc_trace_params *p) {
    if (p->flags & 1u) return p->ray_count;
    return ((int64_t)p->ray_count + 63) / 64 * 64;
}

ORC_API int orc_trace_air(const orc_segment *walls, int n_walls, const float *band_abs, const float *air, const orc_trace_params *p,
                          int64_t *hist, orc_hit *hits, int64_t hit_cap, int64_t *hit_count, orc_counters *out_ctr, int n_threads);

ORC_API int orc_trace(const orc_segment *walls, int n_walls, const float *band_abs, const orc_trace_params *p,
                      int64_t *hist, orc_hit *hits, int64_t hit_cap, int64_t *hit_count, orc_counters *out_ctr,
                      int n_threads) {
    return orc_trace_air(walls, n_walls, band_abs, NULL, p, hist, hits, hit_cap, hit_count, out_ctr, n_threads);
}

/* The same with per-band air absorption (air: [bands] in 1/m, or NULL). */
ORC_API int orc_trace_air(const orc_segment *walls, int n_walls, const float *band_abs, const float *air, const orc_trace_params *p,
                          int64_t *hist, orc_hit *hits, int64_t hit_cap, int64_t *hit_count, orc_counters *out_ctr, int n_threads) {
    if (p->bands > ORC_MAX_BANDS) return -1;
    if (p->bands > 1 && !band_abs) return -2;
    int64_t lo = p->ray_begin, hi = p->ray_end;
    if (lo == 0 && hi == 0) hi = orc_dispatch_threads(p);
    int64_t hc = 0;
    trace_env env = {walls, n_walls, band_abs, p->bands > 1 ? air : NULL, p, hist, hits, hit_cap, &hc};
    orc_counters total;
    memset(&total, 0, sizeof total);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    (void)n_threads;
#pragma omp parallel
    {
        orc_counters c;
        memset(&c, 0, sizeof c);
#pragma omp for schedule(dynamic, 256) nowait
        for (int64_t id = lo; id < hi; id++) trace_one(&env, (uint32_t)id, &c);
#pragma omp critical
        {
            total.ray_bounces += c.ray_bounces;
            total.nearest_tests += c.nearest_tests;
            total.shadow_tests += c.shadow_tests;
            total.direct_hits += c.direct_hits;
            total.nee_hits += c.nee_hits;
        }
    }
    if (hit_count) *hit_count = hc;
    if (out_ctr) *out_ctr = total;
    return 0;
}

/* Float view of a fixed-point IR slot: the un-normalised SUM over accumulated frames (the division by
 * accumCount happens in the convolution, AudioConvolve.compute:30). */
ORC_API void orc_ir_to_float(const int64_t *hist, int64_t n, float *out) {
    for (int64_t i = 0; i < n; i++) out[i] = (float)hist[i] * 9.094947017729282e-13f; /* 2^-40 */
}

/* ---- AudioConvolve.compute:13-31 --------------------------------------------------------- */

ORC_API void orc_convolve(const float *in, int32_t input_length, const float *ir, int32_t ir_length,
                          int32_t accum_count, float *out, int n_threads) {
    int32_t output_length = input_length + ir_length;                        /* :15 (one more than N+M-1) */
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    (void)n_threads;
#pragma omp parallel for schedule(dynamic, 1024)
    for (int32_t n = 0; n < output_length; n++) {
        float sum = 0.0f;
        int32_t start_k = n - ir_length + 1 > 0 ? n - ir_length + 1 : 0;     /* :19 */
        int32_t end_k = n < input_length - 1 ? n : input_length - 1;         /* :20 */
        for (int32_t k = start_k; k <= end_k; k++) {
            float val = in[k];
            if (fabsf(val) > ORC_EPS) sum += val * ir[n - k];                /* :24-27 */
        }
        out[n] = (accum_count > 0) ? (sum / (float)accum_count) : 0.0f;      /* :30 */
    }
}

/* ---- banded model: filter-bank synthesis (direct form) --------------------------------------------------------
 *
 * The reference's banded variant stops at the banded histogram IR[bin*WindowSize + band]
 * (RaytraceOcclusion2D.compute:241-248); its FFT/IFFT kernels (:352-425) are never dispatched and no C# consumes the
 * bands.  The synthesis below is therefore THIS BUILD's definition (include/rar2d.h, "banded model"), restated here
 * in direct form as the checker of the GPU's frequency-domain overlap-add:
 *   g_b[n] = w[n] (hi sinc(hi m) - lo sinc(lo m)),  m = n - 127,  n in [0, 255),  w = Hann over the 255 taps,
 *            lo/hi = the band's edges as fractions of Nyquist (taps rounded to binary32, as the product stores them);
 *   h_b[n] = IR[(n / stride) * bands + b] * 2^-40 when stride divides n, else 0;
 *   out[n] = sum_b sum_k g_b[k] h_b[n + 127 - k],  n in [0, bins * stride), accumulated in double. */
ORC_API void orc_band_filter_taps(double lo, double hi, float *g) {
    const double pi = 3.14159265358979323846;
    for (int n = 0; n < 255; n++) {
        int m = n - 127;
        double ideal = m == 0 ? hi - lo : (sin(pi * hi * m) - sin(pi * lo * m)) / (pi * m);
        double w = 0.5 - 0.5 * cos(2.0 * pi * (n + 1) / 256.0);
        g[n] = (float)(ideal * w);
    }
}

ORC_API void orc_synthesize_ir(const int64_t *hist, int32_t bins, int32_t bands, int32_t stride, const float *edges, float *out) {
    int64_t n_out = (int64_t)bins * stride;
    double *acc = (double *)calloc((size_t)(n_out > 0 ? n_out : 1), sizeof(double));
    float g[255];
    for (int32_t b = 0; b < bands; b++) {
        double lo = edges ? edges[b] : (double)b / bands, hi = edges ? edges[b + 1] : (double)(b + 1) / bands;
        orc_band_filter_taps(lo, hi, g);
        for (int32_t bin = 0; bin < bins; bin++) {
            int64_t q = hist[(int64_t)bin * bands + b];
            if (q == 0) continue;
            double h = (double)((float)q * 9.094947017729282e-13f);          /* the float view of the slot, orc_ir_to_float */
            int64_t j = (int64_t)bin * stride;                                 /* sample the bin stands on */
            for (int k = 0; k < 255; k++) {
                int64_t n = j - 127 + k;                                       /* h_b[j] meets g_b[k] at out[j + k - 127] */
                if (n >= 0 && n < n_out) acc[n] += (double)g[k] * h;
            }
        }
    }
    for (int64_t n = 0; n < n_out; n++) out[n] = (float)acc[n];
    free(acc);
}

/* Same sum in double precision, for judging which of two float results is closer to the truth. */
ORC_API void orc_convolve_f64(const float *in, int32_t input_length, const float *ir, int32_t ir_length,
                              int32_t accum_count, double *out, int n_threads) {
    int32_t output_length = input_length + ir_length;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    (void)n_threads;
#pragma omp parallel for schedule(dynamic, 1024)
    for (int32_t n = 0; n < output_length; n++) {
        double sum = 0.0;
        int32_t start_k = n - ir_length + 1 > 0 ? n - ir_length + 1 : 0;
        int32_t end_k = n < input_length - 1 ? n : input_length - 1;
        for (int32_t k = start_k; k <= end_k; k++) {
            float val = in[k];
            if (fabsf(val) > ORC_EPS) sum += (double)val * (double)ir[n - k];
        }
        out[n] = (accum_count > 0) ? (sum / (double)accum_count) : 0.0;
    }
}

/* ---- Helpers/SceneHelper.cs:78-98 AddLoopToSegments ----------------------------------------
 * Transform = position + Rz(quaternion (0,0,qz,qw)) * (scale (.) local), all binary32.
 * normal = normalized(end-start) rotated to (dir.y, -dir.x), times sign(sx*sy). */
ORC_API int orc_add_loop(const float *local_xy, int n_points, float pos_x, float pos_y, float qz, float qw,
                         float scale_x, float scale_y, float absorption, float scattering, float transmission,
                         float ior, orc_segment *out) {
    float r00 = 1.0f - 2.0f * (qz * qz), r01 = -(2.0f * (qz * qw));
    float r10 = 2.0f * (qz * qw), r11 = r00;
    float prod = scale_x * scale_y;
    float winding = prod >= 0.0f ? 1.0f : -1.0f; /* Mathf.Sign: 1 for >= 0 */
    for (int i = 0; i < n_points; i++) {
        int j = (i + 1) % n_points;
        float lx1 = local_xy[2 * i] * scale_x, ly1 = local_xy[2 * i + 1] * scale_y;
        float lx2 = local_xy[2 * j] * scale_x, ly2 = local_xy[2 * j + 1] * scale_y;
        orc_segment s;
        s.ax = (r00 * lx1 + r01 * ly1) + pos_x;
        s.ay = (r10 * lx1 + r11 * ly1) + pos_y;
        s.bx = (r00 * lx2 + r01 * ly2) + pos_x;
        s.by = (r10 * lx2 + r11 * ly2) + pos_y;
        float dx = s.bx - s.ax, dy = s.by - s.ay;
        float len = sqrtf(dx * dx + dy * dy);
        if (len > 1e-5f) { dx /= len; dy /= len; } else { dx = 0.0f; dy = 0.0f; } /* Vector2.normalized */
        s.nx = dy * winding;
        s.ny = -dx * winding;
        s.absorption = absorption; s.scattering = scattering; s.transmission = transmission; s.ior = ior;
        out[i] = s;
    }
    return n_points;
}

/* ---- RayTraceManager.cs:135-167 LoadSample -------------------------------------------------------
 * C# on the CPU: IEEE binary32, no contraction.  Mathf.RoundToInt rounds half to even; Mathf.Lerp clamps t to
 * [0,1].  Returns the prepared length; writes it to `out` when out != NULL. */
ORC_API int64_t orc_load_sample(const float *raw, int64_t samples, int32_t channels, int32_t clip_frequency,
                                int32_t sample_rate, float *out) {
    if (samples <= 0) return 0;
    float ratio = (float)clip_frequency / (float)sample_rate;                 /* :152 */
    int resample = clip_frequency != sample_rate;                            /* :150 */
    volatile float q = (float)samples / ratio;                               /* :153 */
    int64_t new_len = resample ? (int64_t)nearbyintf(q) : samples;
    if (!out) return new_len;
    float *mono = (float *)malloc((size_t)samples * sizeof(float));
    for (int64_t i = 0; i < samples; i++) {                                  /* :141-147 */
        float sum = 0.0f;
        for (int c = 0; c < channels; c++) sum += raw[i * channels + c];
        mono[i] = sum / (float)channels;
    }
    if (!resample) {
        memcpy(out, mono, (size_t)samples * sizeof(float));
    } else {
        for (int64_t i = 0; i < new_len; i++) {                              /* :156-163 */
            float src = (float)(int32_t)i * ratio;
            int64_t idx0 = (int64_t)floorf(src);
            if (idx0 > samples - 1) idx0 = samples - 1; /* the C# would throw here; unreachable for sane ratios */
            int64_t idx1 = idx0 + 1 < samples - 1 ? idx0 + 1 : samples - 1;
            float t = src - (float)idx0;
            t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
            out[i] = mono[idx0] + (mono[idx1] - mono[idx0]) * t;
        }
    }
    free(mono);
    return new_len;
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
