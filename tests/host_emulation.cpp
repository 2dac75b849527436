// host_emulation.cpp -- TEST TOOLING.  Compiles the product's per-ray logic (csrc/rar_ray.cuh: the
// conservative intersection filter, the bounce state machine, the fixed-point deposit) as host C++ so
// that it can be compared bit-for-bit with the CPU oracle WITHOUT a GPU.  It is not shipped, not
// imported by the package and not a fallback: the product's only execution path is the CUDA kernel
// that includes the same header.
#include <cstring>
#include <vector>

#include "../realisticaudioraytracing2d_b200/csrc/rar_layout.h"

namespace {
template <bool GRID, int FAST = 0>
struct HostSceneT {
    static constexpr bool kGrid = GRID;
    static constexpr bool kPeelFirstBatch = !GRID;  // exercise the small-scene filter of the nearest-hit scan
    static constexpr bool kSpec = (FAST & 1) != 0;   // the range-checked-once pieces (on the host: plain IEEE ops)
    static constexpr bool kFixed4 = (FAST & 2) != 0; // exactly four walls, no loops
    static constexpr bool kPacked = false;           // (the packed-FP32 wall scans are device code)
    rar::GridView gv;
    const rar::GridView &grid() const { return gv; }
    rar::f4 grid_geo(uint32_t i) const { return gv.item_geo[i]; }
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
using HostScene = HostSceneT<false>;

struct Hit { float t, e, x, y; uint32_t ray; uint16_t bounce, kind; };

std::vector<float> g_air;  // per-band air absorption for the next emu_trace calls (empty: none)

template <int BANDS, bool COUNT, bool OPAQUE, class SceneT>
void run(const SceneT &sc, const rar_trace_params &p, long long *hist, Hit *hits, long long cap, long long *count,
         rar::RayCounters &ctr) {
    rar::RayConsts c = rar::ray_consts(p);
    if (BANDS > 1 && (int)g_air.size() == BANDS) {
        c.air_on = 1;
        for (int b = 0; b < BANDS; b++) c.air[b] = g_air[b];
    }
    const rar::SpecConsts spc = rar::spec_consts(c);
    long long lo, hi;
    rar::ray_range(p, lo, hi);
    for (long long id = lo; id < hi; id++) {
        rar::RayState<BANDS> r;
        rar::ray_init(r, (uint32_t)id, c);
        for (int i = 0; i < c.max_bounce_count; i++) {
            rar::Arrival<BANDS> a[2];
            bool alive = rar::ray_bounce<BANDS, COUNT, OPAQUE>(sc, c, r, a[0], a[1], &ctr, nullptr, 0, SceneT::kSpec ? &spc : nullptr);
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

static int emu_trace_impl(bool counting, const rar_segment *walls, int n, const float *band_abs, const rar_trace_params *p,
                          long long *hist, void *hits, long long cap, long long *count, rar_counters *out) {
    std::vector<rar::f4> g(n + 1), m0(n + 1);
    std::vector<rar::f2> m1(n + 1);
    rar::split_walls(walls, n, g.data(), m0.data(), m1.data());
    rar::RayCounters ctr;
    std::memset(&ctr, 0, sizeof ctr);
    long long cnt = 0;
    const bool opaque = rar::walls_opaque(walls, n);  // the production (non-counting) path then uses the OPAQUE instantiation
    if (p->flags & RAR_FLAG_USE_GRID) {  // the uniform-grid instantiation (broadband only in this harness)
        rar::GridHost gh;
        rar::build_grid(walls, n, gh);
        if (gh.nx <= 0 || p->bands > 1) return -6;
        HostSceneT<true> sg;
        sg.gv = rar::GridView{gh.x0, gh.y0, gh.cw, gh.ch, 1.0f / gh.cw, 1.0f / gh.ch, gh.nx, gh.ny, gh.cell_start.data(), gh.items.data(), gh.item_geo.data()};
        sg.g = g.data(); sg.m0 = m0.data(); sg.m1 = m1.data(); sg.ba = band_abs; sg.n = n; sg.nb = p->bands;
        if (counting) run<1, true, false>(sg, *p, hist, (Hit *)hits, cap, &cnt, ctr);
        else if (opaque) run<1, false, true>(sg, *p, hist, (Hit *)hits, cap, &cnt, ctr);
        else run<1, false, false>(sg, *p, hist, (Hit *)hits, cap, &cnt, ctr);
        if (count) *count = cnt;
        if (out) { out->ray_bounces = ctr.ray_bounces; out->nearest_tests = ctr.nearest_tests; out->shadow_tests = ctr.shadow_tests;
                   out->direct_hits = ctr.direct_hits; out->nee_hits = ctr.nee_hits; }
        return 0;
    }
    HostScene sc;
    sc.g = g.data(); sc.m0 = m0.data(); sc.m1 = m1.data(); sc.ba = band_abs; sc.n = n; sc.nb = p->bands;
    // the launch code's choice of the range-checked-once (SPEC) and four-wall (FIXED4) production instantiations
    if (!counting && (p->bands <= 1 || p->bands == 8) && rar::spec_ranges_ok(*p, rar::walls_bounded(walls, n), opaque)) {
        HostSceneT<false, 1> s1;
        s1.g = sc.g; s1.m0 = sc.m0; s1.m1 = sc.m1; s1.ba = sc.ba; s1.n = n; s1.nb = sc.nb;
        HostSceneT<false, 3> s3;
        s3.g = sc.g; s3.m0 = sc.m0; s3.m1 = sc.m1; s3.ba = sc.ba; s3.n = n; s3.nb = sc.nb;
        if (p->bands <= 1) {
            if (n == 4) run<1, false, true>(s3, *p, hist, (Hit *)hits, cap, &cnt, ctr);
            else run<1, false, true>(s1, *p, hist, (Hit *)hits, cap, &cnt, ctr);
        } else {
            if (n == 4) run<8, false, true>(s3, *p, hist, (Hit *)hits, cap, &cnt, ctr);
            else run<8, false, true>(s1, *p, hist, (Hit *)hits, cap, &cnt, ctr);
        }
        if (count) *count = cnt;
        return 0;
    }
    if (p->bands <= 1) {
        if (counting) run<1, true, false>(sc, *p, hist, (Hit *)hits, cap, &cnt, ctr);
        else if (opaque) run<1, false, true>(sc, *p, hist, (Hit *)hits, cap, &cnt, ctr);
        else run<1, false, false>(sc, *p, hist, (Hit *)hits, cap, &cnt, ctr);
    } else if (p->bands == 8) {
        if (counting) run<8, true, false>(sc, *p, hist, (Hit *)hits, cap, &cnt, ctr);
        else if (opaque) run<8, false, true>(sc, *p, hist, (Hit *)hits, cap, &cnt, ctr);
        else run<8, false, false>(sc, *p, hist, (Hit *)hits, cap, &cnt, ctr);
    } else {
        return -5;
    }
    if (count) *count = cnt;
    if (out) {
        out->ray_bounces = ctr.ray_bounces; out->nearest_tests = ctr.nearest_tests; out->shadow_tests = ctr.shadow_tests;
        out->direct_hits = ctr.direct_hits; out->nee_hits = ctr.nee_hits;
    }
    return 0;
}

extern "C" __attribute__((visibility("default")))
void emu_set_air(const float *alpha, int n) { g_air.assign(alpha, alpha + (alpha ? n : 0)); }

// counting != 0: the instantiation that keeps the reference's test counters (always resolves the shadow ray);
// counting == 0: the production instantiation (skips shadow rays whose estimate cannot clear the threshold).
extern "C" __attribute__((visibility("default")))
int emu_trace(const rar_segment *walls, int n, const float *band_abs, const rar_trace_params *p, long long *hist,
              void *hits, long long cap, long long *count, rar_counters *out) {
    return emu_trace_impl(true, walls, n, band_abs, p, hist, hits, cap, count, out);
}
extern "C" __attribute__((visibility("default")))
int emu_trace_nocount(const rar_segment *walls, int n, const float *band_abs, const rar_trace_params *p, long long *hist,
                      void *hits, long long cap, long long *count, rar_counters *out) {
    return emu_trace_impl(false, walls, n, band_abs, p, hist, hits, cap, count, out);
}

// Digest of the host-built uniform grid (the device builder must produce the same lists).
extern "C" __attribute__((visibility("default")))
int emu_grid_digest(const rar_segment *walls, int n, int *nx, int *ny, long long *n_items, unsigned long long *digest) {
    rar::GridHost g;
    rar::build_grid(walls, n, g);
    *nx = g.nx; *ny = g.ny; *n_items = (long long)g.items.size();
    *digest = g.nx > 0 ? rar::grid_digest(g.cell_start.data(), g.cell_start.size(), g.items.data(), g.items.size()) : 0;
    return 0;
}

// the pair-record planes of the packed wall scans (rar_layout.h pair_planes)
extern "C" __attribute__((visibility("default")))
void emu_pair_planes(const rar_segment *walls, int n, float *pair_a, float *pair_b) {
    std::vector<rar::f4> geo(n), m0(n);
    std::vector<rar::f2> m1(n);
    rar::split_walls(walls, n, geo.data(), m0.data(), m1.data());
    rar::pair_planes(geo.data(), n, (rar::f4 *)pair_a, (rar::f4 *)pair_b);
}

// ---- FFT / partitioned overlap-save convolution: the same index logic as conv_kernels.cu ----------------
#include <cmath>
#include "../realisticaudioraytracing2d_b200/csrc/rar_fft.cuh"

namespace {
struct Tables {
    rar::f2 tw[rar::kFftM], tw2[rar::kFftM / 2 + 1];
    Tables() {
        const double two_pi = 6.283185307179586476925286766559;
        for (int k = 0; k < rar::kFftM; k++) tw[k] = rar::f2{(float)std::cos(two_pi * k / rar::kFftM), (float)-std::sin(two_pi * k / rar::kFftM)};
        for (int k = 0; k <= rar::kFftM / 2; k++)
            tw2[k] = rar::f2{(float)std::cos(two_pi * k / (2 * rar::kFftM)), (float)-std::sin(two_pi * k / (2 * rar::kFftM))};
    }
};
const Tables &tables() { static Tables t; return t; }

void fft256(rar::f2 *a, rar::f2 *b, bool inverse) {
    const Tables &t = tables();
    const int ps[4] = {1, 4, 16, 64};
    rar::f2 *src = a, *dst = b;
    for (int s = 0; s < 4; s++) {
        for (int i = 0; i < 64; i++) rar::fft_pass_r4(src, dst, i, ps[s], t.tw, inverse);
        rar::f2 *tmp = src; src = dst; dst = tmp;
    }  // four passes: result back in a
}
void rfft512(const float *w, rar::f2 *P) {
    rar::f2 a[256], b[256];
    for (int n = 0; n < 256; n++) a[n] = rar::f2{w[2 * n], w[2 * n + 1]};
    fft256(a, b, false);
    for (int k = 0; k <= 128; k++) rar::rfft_split(a, P, k, tables().tw2);
}
void irfft512(const rar::f2 *P, float *w) {  // w = 256 * true inverse
    rar::f2 a[256], b[256];
    for (int k = 0; k <= 128; k++) rar::irfft_merge(P, a, k, tables().tw2);
    fft256(a, b, true);
    for (int n = 0; n < 256; n++) { w[2 * n] = a[n].x; w[2 * n + 1] = a[n].y; }
}
}  // namespace

extern "C" __attribute__((visibility("default"))) void emu_rfft512(const float *w, float *P) { rfft512(w, (rar::f2 *)P); }
extern "C" __attribute__((visibility("default"))) void emu_irfft512(const float *P, float *w) { irfft512((const rar::f2 *)P, w); }

// ---- filter-bank synthesis with the 16 x 16 register transforms of rar_synth16.cuh (band_synth.cu's arithmetic) ----
#include "../realisticaudioraytracing2d_b200/csrc/rar_synth16.cuh"

extern "C" __attribute__((visibility("default")))
void emu_fft16(float *x /*16 complex*/, int inverse, int half) {
    rar::f2 v[16];
    std::memcpy(v, x, sizeof v);
    if (inverse) { if (half) rar::fft16<true, true>(v); else rar::fft16<true, false>(v); }
    else { if (half) rar::fft16<false, true>(v); else rar::fft16<false, false>(v); }
    std::memcpy(x, v, sizeof v);
}

// out[n] += sum_b (g_b * h_b)[n + 127], h_b[n] = hist[n*bands + b] * 2^-40 * scale; taps [bands][256]
extern "C" __attribute__((visibility("default")))
void emu_band_synth16(const long long *hist, int bins, int bands, float scale, const float *taps, float *out, int out_len) {
    using rar::f2;
    std::vector<f2> T(rar::synth_table_len(bands));
    rar::synth_tables(taps, bands, T.data());
    const f2 *tw = T.data(), *w2 = T.data() + 256, *wt = T.data() + 512;
    const int n_seg = (bins + 255) / 256;
    for (int p = 0; p < n_seg; p++) {
        const long long first = (long long)p * 256;
        f2 U[16][16], W[16][16], buf[16][16];
        std::memset(U, 0, sizeof U);
        std::memset(W, 0, sizeof W);
        for (int b = 0; b < bands; b++) {
            for (int t = 0; t < 16; t++) {  // stage 1: thread t holds z[t + 16 r], r < 8
                f2 y[16];
                for (int r = 0; r < 8; r++) {
                    const long long g = first + 2 * (t + 16 * r);
                    const float a = g < bins ? ((float)hist[g * bands + b] * 9.094947017729282e-13f) * scale : 0.0f;
                    const float c = g + 1 < bins ? ((float)hist[(g + 1) * bands + b] * 9.094947017729282e-13f) * scale : 0.0f;
                    y[r] = f2{a, c};
                }
                rar::fft16<false, true>(y);
                for (int k1 = 0; k1 < 16; k1++) buf[k1][t] = k1 ? rar::cmul(y[k1], tw[rar::synth_tab(t, k1)]) : y[k1];
            }
            for (int k1 = 0; k1 < 16; k1++) {  // stage 2: thread k1 holds Z[k1 + 16 k2]
                f2 v[16], ab[16];
                for (int t = 0; t < 16; t++) v[t] = buf[k1][t];
                rar::fft16<false, false>(v);
                for (int k2 = 0; k2 < 16; k2++) ab[k2] = wt[(size_t)b * 256 + rar::synth_tab(k1, k2)];
                rar::synth_accumulate(U[k1], W[k1], v, ab);
            }
        }
        for (int t = 0; t < 16; t++) {  // partner exchange, merge, inverse stage 1
            f2 Zp[16];
            const int src = (16 - t) & 15;
            for (int k2 = 0; k2 < 16; k2++) {
                const int reg = t == 0 ? (16 - k2) & 15 : 15 - k2;
                Zp[k2] = rar::synth_merge1(t == 0 && k2 == 0, U[t][k2], W[t][k2], rar::csub(W[src][reg], U[src][reg]), w2[rar::synth_tab(t, k2)]);
            }
            rar::fft16<true, false>(Zp);
            for (int n2 = 0; n2 < 16; n2++) buf[n2][t] = n2 ? rar::cmul(Zp[n2], rar::conj2(tw[rar::synth_tab(t, n2)])) : Zp[n2];
        }
        for (int n2 = 0; n2 < 16; n2++) {  // inverse stage 2: thread n2 holds z'[n2 + 16 n1]
            f2 v[16];
            for (int k1 = 0; k1 < 16; k1++) v[k1] = buf[n2][k1];
            rar::fft16<true, false>(v);
            for (int n1 = 0; n1 < 16; n1++) {
                const int m = 2 * (n2 + 16 * n1);
                const long long o = first + (m < 384 ? m : m - 512);
                if (o >= 0 && o < out_len) out[o] += v[n1].x * (1.0f / 256.0f);
                if (o + 1 >= 0 && o + 1 < out_len) out[o + 1] += v[n1].y * (1.0f / 256.0f);
            }
        }
    }
}

// The one-shot pipeline of rar_convolve_begin: input windows, per-block CMAC over partitions, output blocks.
extern "C" __attribute__((visibility("default")))
void emu_convolve(const float *x, int x_len, const float *ir, int ir_len, int accum, float *out) {
    const int B = 256, out_len = x_len + ir_len;
    for (int i = 0; i < out_len; i++) out[i] = 0.f;
    if (accum <= 0 || x_len == 0 || ir_len == 0) return;
    const int n_part = (ir_len + B - 1) / B, n_xwin = (x_len + B - 1) / B + 1, n_out = (out_len + B - 1) / B;
    std::vector<rar::f2> H((size_t)n_part * B), X((size_t)n_xwin * B), Y(B);
    float w[512];
    for (int p = 0; p < n_part; p++) {
        for (int t = 0; t < 512; t++) { long long g = (long long)p * B + t; w[t] = (t < B && g < ir_len) ? ir[g] : 0.f; }
        rfft512(w, &H[(size_t)p * B]);
    }
    for (int j = 0; j < n_xwin; j++) {
        for (int t = 0; t < 512; t++) {
            long long g = ((long long)j - 1) * B + t;
            float v = (g >= 0 && g < x_len) ? x[g] : 0.f;
            w[t] = std::fabs(v) > 1e-4f ? v : 0.f;
        }
        rfft512(w, &X[(size_t)j * B]);
    }
    const float scale = (1.0f / (float)accum) / (float)B;
    for (int j = 0; j < n_out; j++) {
        int p_lo = j - (n_xwin - 1); if (p_lo < 0) p_lo = 0;
        int p_hi = j < n_part - 1 ? j : n_part - 1;
        for (int k = 0; k < B; k++) {
            float re = 0, im = 0, ac = 0, bd = 0;
            for (int p = p_lo; p <= p_hi; p++) {
                rar::f2 a = X[(size_t)(j - p) * B + k], h = H[(size_t)p * B + k];
                ac += a.x * h.x; bd += a.y * h.y; im += a.x * h.y + a.y * h.x;
            }
            re = ac - bd;
            Y[k] = k == 0 ? rar::f2{ac, bd} : rar::f2{re, im};
        }
        irfft512(Y.data(), w);
        for (int i = 0; i < B; i++) { long long o = (long long)j * B + i; if (o < out_len) out[o] = w[B + i] * scale; }
    }
}

// Single ray-segment test through the product's filter + exact path (bound = closest), for comparison
// with the oracle's literal Common.hlsl:14-21.
extern "C" __attribute__((visibility("default")))
void emu_intersect_many(const float *rays /*ox,oy,dx,dy*/, const float *segs /*ax,ay,bx,by*/, const float *closest,
                        int n, float *out) {
    for (int i = 0; i < n; i++) {
        const float *r = rays + 4 * i, *s = segs + 4 * i;
        volatile float ex = s[2] - s[0], ey = s[3] - s[1];
        rar::f4 g{s[0], s[1], ex, ey};
        rar::WallTest t = rar::wall_test(g, r[0], r[1], r[2], -r[3]);
        float d = rar::kInf;
        if (rar::wall_pass(t, closest[i] * rar::kSlack)) d = rar::intersect_exact(t.num1, t.num2, t.dotP);
        out[i] = d < closest[i] ? d : rar::kInf;
    }
}
