// host_emulation.cpp -- TEST TOOLING.  Compiles the product's per-ray logic (csrc/rar_ray.cuh: the
// conservative intersection filter, the bounce state machine, the fixed-point deposit) as host C++ so
// that it can be compared bit-for-bit with the CPU oracle WITHOUT a GPU.  It is not shipped, not
// imported by the package and not a fallback: the product's only execution path is the CUDA kernel
// that includes the same header.
#include <cstring>
#include <vector>

#include "../realisticaudioraytracing2d_b200/csrc/rar_layout.h"

namespace {
struct HostScene {
    const rar::f4 *g, *m0;
    const rar::f2 *m1;
    const float *ba;
    int n, nb;
    int n_walls() const { return n; }
    rar::f4 geo(int w) const { return g[w]; }
    rar::f4 mat0(int w) const { return m0[w]; }
    rar::f2 mat1(int w) const { return m1[w]; }
    const float *band_abs(int w) const { return ba + (size_t)w * nb; }
};

struct Hit { float t, e, x, y; uint32_t ray; uint16_t bounce, kind; };

template <int BANDS>
void run(const HostScene &sc, const rar_trace_params &p, long long *hist, Hit *hits, long long cap, long long *count,
         rar::RayCounters &ctr) {
    rar::RayConsts c = rar::ray_consts(p);
    long long lo, hi;
    rar::ray_range(p, lo, hi);
    for (long long id = lo; id < hi; id++) {
        rar::RayState<BANDS> r;
        rar::ray_init(r, (uint32_t)id, c);
        for (int i = 0; i < c.max_bounce_count; i++) {
            rar::Arrival<BANDS> a[2];
            bool alive = rar::ray_bounce<BANDS, true>(sc, c, r, a[0], a[1], &ctr);
            for (int k = 0; k < 2; k++) {
                if (!a[k].has) continue;
                if (hits) {
                    if (*count < cap) hits[*count] = Hit{a[k].t, a[k].e, a[k].hx, a[k].hy, (uint32_t)id, (uint16_t)i, (uint16_t)k};
                    ++*count;
                }
                if (hist) {
                    int bin = rar::time_bin(a[k].t, c.sample_rate, c.time_divisor, c.impulse_length);
                    if (bin < 0) continue;
                    if (BANDS == 1) hist[bin] += rar::quantize_energy(a[k].e);
                    else for (int b = 0; b < BANDS; b++) hist[(long long)bin * BANDS + b] += rar::quantize_energy(a[k].band_e[b]);
                }
            }
            if (!alive) break;
        }
    }
}
}  // namespace

extern "C" __attribute__((visibility("default")))
int emu_trace(const rar_segment *walls, int n, const float *band_abs, const rar_trace_params *p, long long *hist,
              void *hits, long long cap, long long *count, rar_counters *out) {
    std::vector<rar::f4> g(n + 1), m0(n + 1);
    std::vector<rar::f2> m1(n + 1);
    rar::split_walls(walls, n, g.data(), m0.data(), m1.data());
    HostScene sc{g.data(), m0.data(), m1.data(), band_abs, n, p->bands};
    rar::RayCounters ctr;
    std::memset(&ctr, 0, sizeof ctr);
    long long cnt = 0;
    if (p->bands <= 1) run<1>(sc, *p, hist, (Hit *)hits, cap, &cnt, ctr);
    else if (p->bands == 8) run<8>(sc, *p, hist, (Hit *)hits, cap, &cnt, ctr);
    else return -5;
    if (count) *count = cnt;
    if (out) {
        out->ray_bounces = ctr.ray_bounces; out->nearest_tests = ctr.nearest_tests; out->shadow_tests = ctr.shadow_tests;
        out->direct_hits = ctr.direct_hits; out->nee_hits = ctr.nee_hits;
    }
    return 0;
}
