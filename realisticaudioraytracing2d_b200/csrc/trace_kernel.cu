// trace_kernel.cu -- the fused Trace + ProcessHits kernel for sm_100a.
//
// Replaces Raytrace2D.compute `Trace` (:49-156) and `ProcessHits` (:157-165): one thread per ray,
// looping over bounces; every arrival is added straight into the time-binned (optionally banded)
// 64-bit fixed-point impulse-response histogram, so there is no hit buffer and no hit-count
// round trip through the host (RayTraceManager.cs:208-209).
//
//  * The wall planes are staged into shared memory once per CTA by 1-D TMA bulk copies
//    (cp.async.bulk.shared::cluster.global + mbarrier complete_tx; SASS: UBLKCP).  All lanes of a
//    warp read the same wall record, so the inner loop is one broadcast LDS.128 per test.
//  * CTAs are persistent: the grid is sized to the resident capacity of the device and each CTA
//    strides over tiles of rays, so a 10k-wall scene (160 KB) is staged 148 times, not 65 536 times.
//  * Deposits are warp-aggregated: lanes that hit the same bin are found with __match_any_sync, their
//    Q23.40 values are summed with three __reduce_add_sync limb reductions, and one lane issues a single
//    64-bit integer atomic.  Integer addition commutes, so the histogram is bit-identical for any
//    scheduling, any grid and any number of GPUs.
//
// Compiled with --fmad=false: the arithmetic contract of rar_math.cuh fixes every rounding.
#include <cuda_runtime.h>

#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "rar_internal.h"

namespace rar {
namespace {

constexpr unsigned kFull = 0xffffffffu;

#ifndef RAR_GRID_MIN_BLOCKS
#define RAR_GRID_MIN_BLOCKS 5
#endif

// ---- TMA / mbarrier primitives (PTX ISA 8.x, sm_90+) ------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// ---- scene views ------------------------------------------------------------------------------------

// STAGE 0: endpoint and material planes in shared memory; 1: endpoints in shared, materials in global;
// 2: everything read through the read-only global path (scenes too large for shared memory).
// FAST bit 0 (SPEC): the range-checked-once arithmetic of rar_math.cuh (host-validated operand ranges, opaque scene);
// FAST bit 1 (FIXED4): the scene has exactly four walls, so both wall scans are one batch with no loop.
//
// Shared-memory planes are addressed through their 32-bit shared-space address with explicit ld.shared: reading
// them through generic pointers made the compiler re-derive the CTA's shared-window base (S2UR SR_CgaCtaId + ULEA)
// at every use, three times per bounce.
__device__ __forceinline__ f4 lds_f4(uint32_t addr) {
    f4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ f2 lds_f2(uint32_t addr) {
    f2 v;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

template <int STAGE, bool GRID = false, int FAST = 0>
struct SceneView {
    static constexpr bool kGrid = GRID;
    static constexpr bool kSpec = (FAST & 1) != 0;
    static constexpr bool kFixed4 = (FAST & 2) != 0;
    // the staged endpoint plane holds two walls per record (rar_layout.h pair_planes) and the wall scans run on packed
    // FP32 instructions, two walls each (rar_ray.cuh "packed wall tests")
    static constexpr bool kPacked = (FAST & 4) != 0;
    // small scenes (everything in shared memory, < kCoopMinWalls walls): the first filter batch of the
    // nearest-hit scan is peeled and uses the sign-aware filter (rar_ray.cuh wall_pass_unbounded); measured
    // -5.5 % on the 4-wall config 2, +6 % on the 10 000-wall maze when applied there too (code layout), so it
    // is confined to this variant
    static constexpr bool kPeelFirstBatch = STAGE == 0 && !GRID;
    const f4 *g;
    const f4 *m0;
    const f2 *m1;
    const f4 *pa, *pb;      // kPacked, STAGE 2: the pair planes in global memory
    uint32_t gs, m0s, m1s;  // shared-space addresses of the staged planes (STAGE 0: all three, STAGE 1: gs)
    uint32_t pas, pbs;      // kPacked: the pair planes (STAGE 1: pair_a sits where the endpoint plane would, pas == gs)
    const float *ba;
    int n, nb, boff;
    GridView gv;
    __device__ __forceinline__ const GridView &grid() const { return gv; }
    __device__ __forceinline__ f4 grid_geo(uint32_t i) const {
        float4 v = __ldg(reinterpret_cast<const float4 *>(gv.item_geo) + i);
        return f4{v.x, v.y, v.z, v.w};
    }
    __device__ __forceinline__ int n_walls() const { return n; }
    // pair record p: (x0 x1), (y0 y1) from pair_a; (ex0 ex1), (ey0 ey1) from pair_b -- one LDS.128 each
    __device__ __forceinline__ void pair(int p, uint64_t &sx, uint64_t &sy, uint64_t &sz, uint64_t &sw) const {
        if (STAGE == 2) {  // read-only global path: warp-broadcast in the nearest-hit scan, coalesced in the cooperative one
            const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2 *>(pa) + p), b = __ldg(reinterpret_cast<const ulonglong2 *>(pb) + p);
            sx = a.x; sy = a.y; sz = b.x; sw = b.y;
            return;
        }
        asm("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(sx), "=l"(sy) : "r"(pas + (uint32_t)p * 16u));
        asm("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(sz), "=l"(sw) : "r"(pbs + (uint32_t)p * 16u));
    }
    __device__ __forceinline__ f4 geo(int w) const {
        if (kPacked && STAGE == 1) {  // (only code that was not converted to pairs asks for a single wall: none at present)
            const uint32_t o = (uint32_t)(w >> 1) * 16u + (uint32_t)(w & 1) * 4u;
            float x, y, z, ww;
            asm("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(pas + o));
            asm("ld.shared.f32 %0, [%1];" : "=f"(y) : "r"(pas + o + 8u));
            asm("ld.shared.f32 %0, [%1];" : "=f"(z) : "r"(pbs + o));
            asm("ld.shared.f32 %0, [%1];" : "=f"(ww) : "r"(pbs + o + 8u));
            return f4{x, y, z, ww};
        }
        if (STAGE == 2) {
            float4 v = __ldg(reinterpret_cast<const float4 *>(g) + w);
            return f4{v.x, v.y, v.z, v.w};
        }
        return lds_f4(gs + (uint32_t)w * 16u);
    }
    __device__ __forceinline__ f4 mat0(int w) const {
        if (STAGE != 0) {
            float4 v = __ldg(reinterpret_cast<const float4 *>(m0) + w);
            return f4{v.x, v.y, v.z, v.w};
        }
        return lds_f4(m0s + (uint32_t)w * 16u);
    }
    __device__ __forceinline__ f2 mat1(int w) const {
        if (STAGE != 0) {
            float2 v = __ldg(reinterpret_cast<const float2 *>(m1) + w);
            return f2{v.x, v.y};
        }
        return lds_f2(m1s + (uint32_t)w * 8u);
    }
    // row w of the [n][band_total] table, starting at this chunk's first band (the table is padded by 8 floats,
    // so a short last chunk may read, and then ignore, a few values past its row)
    __device__ __forceinline__ const float *band_abs(int w) const { return ba + (size_t)w * nb + boff; }
};

// ---- warp-aggregated fixed-point deposit ------------------------------------------------------------

// Sum of q over the lanes in `peers` (all of which call this with the same mask).  64-bit values are
// reduced as three limbs with the 32-bit REDUX instruction: 24 + 24 + 16 (signed) bits.
__device__ __forceinline__ long long group_sum_q(unsigned peers, long long q) {
    unsigned lo = (unsigned)(q & 0xffffff);
    unsigned mid = (unsigned)((q >> 24) & 0xffffff);
    int hi = (int)(q >> 48);
    lo = __reduce_add_sync(peers, lo);
    mid = __reduce_add_sync(peers, mid);
    hi = __reduce_add_sync(peers, hi);
    return (long long)lo + ((long long)mid << 24) + ((long long)hi << 48);
}

// The same for values known to lie in [0, 2^54): two limbs of 27 bits (32 lanes x 2^27 < 2^32).
__device__ __forceinline__ long long group_sum_q2(unsigned peers, long long q) {
    unsigned lo = (unsigned)q & 0x7ffffffu;
    unsigned hi = (unsigned)(q >> 27);
    lo = __reduce_add_sync(peers, lo);
    hi = __reduce_add_sync(peers, hi);
    return (long long)lo + ((long long)hi << 27);
}
// An energy whose fixed-point value fits the two-limb reduction: 0 <= e < 2^14 (false for NaN).
__device__ __forceinline__ bool fits_two_limbs(float e) { return e >= 0.0f && e < 16384.0f; }

// SPARSE: up to kSparseArrivals arrivals per warp are deposited without aggregation.  Measured: -2.8 % time on
// the 10 000-wall maze, +1.2 % on the 4-wall config 2 (whose kernel is bound by instruction fetch/issue and pays
// for the extra code), so the small-scene variants keep the single aggregated path.
constexpr int kSparseArrivals = 6;

// One lane's arrival added with plain 64-bit atomics, no warp cooperation: callable from divergent code.
template <int BANDS>
__device__ __forceinline__ void deposit_lane(const TraceLaunch &a, unsigned long long *hist, const Arrival<BANDS> &h, int bin) {
    if (BANDS == 1) {
        const long long q = quantize_energy(h.e);
        if (q != 0) atomicAdd(hist + bin, (unsigned long long)q);
    } else {
        unsigned long long *row = hist + (size_t)bin * a.band_total + a.band_offset;
#pragma unroll
        for (int b = 0; b < BANDS; b++) {
            if (b >= a.band_valid) break;  // short last chunk of a banded slot
            const long long q = quantize_energy(h.band_e[b]);
            if (q != 0) atomicAdd(row + b, (unsigned long long)q);
        }
    }
}

// The direct listener crossing of a bounce (Raytrace2D.compute:74-84) is rare -- one warp-bounce in ten holds one on
// config 2 -- so it is deposited where it is found, inside the live-ray region: no full-mask vote and no reconvergence
// point on the other nine (the convergent path cost 14 issue slots per warp-bounce just to find out that nobody had
// one).  But when it happens it usually happens to the whole warp (adjacent rays follow the same mirror path and cross
// the listener together, into the same bin), and 32 atomics on one address serialise in one L2 slice: the rank whose
// ray sector faces the listener ran 13 % longer than the others.  So the lanes that meet here aggregate
// opportunistically: whoever is converged at this point (__activemask) matches bins and sums shared ones before the
// atomic.  Any grouping gives the same integer total, so the histogram stays deterministic.
template <int BANDS>
__device__ __forceinline__ void deposit_direct(const TraceLaunch &a, unsigned long long *hist, const Arrival<BANDS> &h) {
    if (!h.has) return;
    const int bin = time_bin(h.t, a.p.sample_rate_f, a.p.time_divisor, a.p.impulse_length_f, a.p.impulse_length);
    if (bin < 0) return;
    const unsigned here = __activemask();
    const unsigned peers = __match_any_sync(here, bin);
    if ((peers & (peers - 1)) == 0) {  // alone in its bin
        deposit_lane<BANDS>(a, hist, h, bin);
        return;
    }
    const bool leader = (threadIdx.x & 31u) == (unsigned)(__ffs(peers) - 1);
    if (BANDS == 1) {
        long long q = quantize_energy(h.e);
        if (!fits_two_limbs(h.e)) {
            if (q != 0) atomicAdd(hist + bin, (unsigned long long)q);
            q = 0;
        }
        q = group_sum_q2(peers, q);
        if (leader && q != 0) atomicAdd(hist + bin, (unsigned long long)q);
    } else {
        unsigned long long *row = hist + (size_t)bin * a.band_total + a.band_offset;
#pragma unroll
        for (int b = 0; b < BANDS; b++) {
            if (b >= a.band_valid) break;  // short last chunk of a banded slot
            long long q = quantize_energy(h.band_e[b]);
            if (!fits_two_limbs(h.band_e[b])) {
                if (q != 0) atomicAdd(row + b, (unsigned long long)q);
                q = 0;
            }
            q = group_sum_q2(peers, q);
            if (leader && q != 0) atomicAdd(row + b, (unsigned long long)q);
        }
    }
}

// Warp-convergent aggregated deposit (the next-event arrivals: up to 32 per warp-bounce, adjacent rays share bins).
// The lanes that hold an arrival find their bin peers with __match_any_sync over exactly those lanes; a bin held by
// one lane goes out as it is, a shared bin is summed by limb reductions first.
template <int BANDS, bool SPARSE>
__device__ __forceinline__ void deposit_hist(const TraceLaunch &a, unsigned long long *hist, const Arrival<BANDS> &h,
                                             unsigned lane) {
    int bin = -1;
    if (h.has) bin = time_bin(h.t, a.p.sample_rate_f, a.p.time_divisor, a.p.impulse_length_f, a.p.impulse_length);
    const bool valid = bin >= 0;
    const unsigned have = __ballot_sync(kFull, valid);
    if (have == 0) return;
    if (SPARSE && __popc(have) <= kSparseArrivals) {
        // a few arrivals in the warp: plain atomics cost less than finding out whether two of them share a bin
        if (valid) deposit_lane<BANDS>(a, hist, h, bin);
        return;
    }
    if (!valid) return;
    const unsigned peers = __match_any_sync(have, bin);
    const bool shared_bin = (peers & (peers - 1)) != 0;  // same for every lane of a peer group
    const bool leader = lane == (unsigned)(__ffs(peers) - 1);
    // Shared bins are summed with two 27-bit limb reductions, which covers energies in [0, 2^14); a lane holding
    // anything else (a huge or negative gain) adds its own value and enters the reduction with zero.
    if (BANDS == 1) {
        long long q = quantize_energy(h.e);
        if (shared_bin) {
            if (!fits_two_limbs(h.e)) {
                if (q != 0) atomicAdd(hist + bin, (unsigned long long)q);
                q = 0;
            }
            q = group_sum_q2(peers, q);
        }
        if (leader && q != 0) atomicAdd(hist + bin, (unsigned long long)q);
    } else {
        unsigned long long *row = hist + (size_t)bin * a.band_total + a.band_offset;
#pragma unroll
        for (int b = 0; b < BANDS; b++) {
            if (b >= a.band_valid) break;  // short last chunk of a banded slot
            long long q = quantize_energy(h.band_e[b]);
            if (shared_bin) {
                if (!fits_two_limbs(h.band_e[b])) {
                    if (q != 0) atomicAdd(row + b, (unsigned long long)q);
                    q = 0;
                }
                q = group_sum_q2(peers, q);
            }
            if (leader && q != 0) atomicAdd(row + b, (unsigned long long)q);
        }
    }
}

template <int BANDS>
__device__ __forceinline__ void emit_hit(const TraceLaunch &a, const Arrival<BANDS> &h, uint32_t id, int bounce, int kind) {
    if (!h.has) return;
    unsigned long long slot = atomicAdd(a.hit_count, 1ull);
    if ((long long)slot < a.hit_cap) {
        a.hits[slot] = rar_ray_info{h.t, h.e, {h.hx, h.hy}};
        if (a.keys) a.keys[slot] = rar_hit_key{id, (uint16_t)bounce, (uint16_t)kind};
    }
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// ---- warp-cooperative shadow rays -------------------------------------------------------------------
//
// In a large scene a per-thread checkVis walk wastes the warp: a lane whose shadow ray is blocked early
// idles while its neighbours scan all n walls.  Here the warp takes the pending shadow rays one at a time;
// the 32 lanes test 64 consecutive walls per iteration (conflict-free 16-byte shared-memory reads) and a
// ballot finds the first blocking wall, which is exactly where the reference's loop would have stopped.
// Work per shadow ray is proportional to the tests the reference performs, not to n.
template <bool COUNT, class Scene>
__device__ __forceinline__ bool coop_shadow(const Scene &sc, bool pending, const ShadowRay &q, unsigned lane,
                                            RayCounters &ctr) {
    unsigned need = __ballot_sync(kFull, pending);
    bool my_vis = true;
    const int n = sc.n_walls();
    while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        ShadowRay b;
        b.sx = __shfl_sync(kFull, q.sx, src);
        b.sy = __shfl_sync(kFull, q.sy, src);
        b.dx = __shfl_sync(kFull, q.dx, src);
        b.ndy = __shfl_sync(kFull, q.ndy, src);
        b.lim = __shfl_sync(kFull, q.lim, src);
        int first = -1;
        if (n > 0 && kInf < b.lim) {
            first = 0;  // every intersect() result (<= inf) is < lim: wall 0 blocks
        } else {
            const float lim_m = b.lim * kSlack;
            if constexpr (Scene::kPacked) {
                // 32 lanes x one pair record = the same 64 consecutive walls per iteration, walls 2p and 2p+1 per lane
                const int np = (n + 1) >> 1;
                for (int pbase = 0; pbase < np; pbase += 32) {
                    const int pi = pbase + (int)lane;
                    bool k0 = false, k1 = false;
                    if (pi < np) {
                        const WallTest2 t = wall_test2(sc, pi, b.sx, b.sy, b.dx, b.ndy);
                        bool p0, p1;
                        wall_pass2(t, lim_m, p0, p1);
                        k0 = p0 && intersect_exact(pair_lo(t.num1), pair_lo(t.num2), -pair_lo(t.ndotP)) < b.lim;
                        k1 = p1 && intersect_exact(pair_hi(t.num1), pair_hi(t.num2), -pair_hi(t.ndotP)) < b.lim;
                    }
                    const unsigned m0 = __ballot_sync(kFull, k0), m1 = __ballot_sync(kFull, k1);
                    if (m0 | m1) {
                        const int f0 = m0 ? 2 * (__ffs(m0) - 1) : (1 << 30), f1 = m1 ? 2 * (__ffs(m1) - 1) + 1 : (1 << 30);
                        first = 2 * pbase + min(f0, f1);
                        break;
                    }
                }
            } else
            for (int base = 0; base < n; base += 64) {
                const int w0 = base + (int)lane, w1 = w0 + 32;
                const bool k0 = w0 < n && shadow_blocked_by(sc.geo(w0), b, lim_m);
                const bool k1 = w1 < n && shadow_blocked_by(sc.geo(w1), b, lim_m);
                const unsigned m0 = __ballot_sync(kFull, k0), m1 = __ballot_sync(kFull, k1);
                if (m0 | m1) {
                    first = m0 ? base + __ffs(m0) - 1 : base + 32 + __ffs(m1) - 1;
                    break;
                }
            }
        }
        if (lane == (unsigned)src) {
            my_vis = first < 0;
            if (COUNT) ctr.shadow_tests += (unsigned long long)(first >= 0 ? first + 1 : n);
        }
    }
    return my_vis;
}

// Stages the wall planes into shared memory with 1-D TMA bulk copies (STAGE 0: all three planes, STAGE 1: the
// endpoint plane) and returns the view the ray code reads.  All threads of the CTA call this once.
template <int STAGE, bool GRID, int FAST>
__device__ __forceinline__ SceneView<STAGE, GRID, FAST> stage_scene(const TraceLaunch &a, unsigned char *smem_raw) {
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    f4 *s_geo = reinterpret_cast<f4 *>(smem_raw + 16);
    f4 *s_mat0 = s_geo + a.n_walls;
    f2 *s_mat1 = reinterpret_cast<f2 *>(s_mat0 + a.n_walls);

    if (STAGE < 2 && a.n_walls > 0) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            fence_barrier_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t geo_bytes = (uint32_t)a.n_walls * 16u;
            const uint32_t m1_bytes = ((uint32_t)a.n_walls * 8u + 15u) & ~15u;  // planes are padded to 16 B
            const uint32_t pair_bytes = (FAST & 4) ? (uint32_t)((a.n_walls + 1) >> 1) * 16u : 0u;
            // PACKED: stage 1 holds the two pair planes INSTEAD of the endpoint plane, stage 0 holds them behind the others
            const uint32_t total = STAGE == 0 ? geo_bytes * 2u + m1_bytes + 2u * pair_bytes : ((FAST & 4) ? 2u * pair_bytes : geo_bytes);
            mbar_arrive_expect_tx(bar, total);
            unsigned char *s_pairs = STAGE == 0 ? reinterpret_cast<unsigned char *>(s_mat1) + m1_bytes : reinterpret_cast<unsigned char *>(s_geo);
            if (FAST & 4) {
                bulk_copy_g2s(s_pairs, a.pair_a, pair_bytes, bar);
                bulk_copy_g2s(s_pairs + pair_bytes, a.pair_b, pair_bytes, bar);
            }
            if (STAGE == 0 || !(FAST & 4)) bulk_copy_g2s(s_geo, a.geo, geo_bytes, bar);
            if (STAGE == 0) {
                bulk_copy_g2s(s_mat0, a.mat0, geo_bytes, bar);
                bulk_copy_g2s(s_mat1, a.mat1, m1_bytes, bar);
            }
        }
        while (!mbar_try_wait(bar, 0)) {
        }
    }

    SceneView<STAGE, GRID, FAST> sc;
    sc.gv = a.grid;
    sc.g = STAGE < 2 ? s_geo : a.geo;
    sc.pa = a.pair_a;
    sc.pb = a.pair_b;
    sc.m0 = STAGE == 0 ? s_mat0 : a.mat0;
    sc.m1 = STAGE == 0 ? s_mat1 : a.mat1;
    // The plain-asm shared loads carry no memory dependence of their own.  Their addresses derive from a value that
    // passes through a volatile asm placed after the barrier wait (volatile asms keep their program order), which
    // keeps every load behind the wait -- and keeps the base in a register instead of being re-derived (S2UR + ULEA)
    // at each use.
    uint32_t base = smem_u32(smem_raw) + 16u;
    asm volatile("" : "+r"(base) : : "memory");
    sc.gs = base;
    sc.m0s = base + (uint32_t)a.n_walls * 16u;
    sc.m1s = base + (uint32_t)a.n_walls * 32u;
    sc.pas = STAGE == 0 ? base + (uint32_t)a.n_walls * 32u + (((uint32_t)a.n_walls * 8u + 15u) & ~15u) : base;
    sc.pbs = sc.pas + (uint32_t)((a.n_walls + 1) >> 1) * 16u;
    // (opaque to the compiler from here on: with 64 registers it preferred to RE-DERIVE pbs from the launch
    // arguments inside the wall loop -- seven instructions per iteration, 13 % of the maze kernel -- to holding it)
    if (FAST & 4) asm volatile("" : "+r"(sc.pas), "+r"(sc.pbs));
    sc.ba = a.band_abs;
    sc.n = a.n_walls;
    sc.nb = a.band_total;
    sc.boff = a.band_offset;
    return sc;
}

// ---- the kernel -------------------------------------------------------------------------------------

// Resident CTAs per SM the register allocation is held to (0 = ptxas' own choice).  Measured on B200:
//  * the small-scene variant is pinned at 64 registers (4 CTAs/SM): 0.607 ms on config 2, against 0.645 ms at 76
//    registers (3 CTAs) and 0.625 ms at 51 (5 CTAs, spills);
//  * the 8-band variants otherwise take ~105 registers (2 CTAs); at 64 with a few spills they run 20-30 % faster;
//  * grid walks are latency-bound (dependent loads, divergent lanes) and like more warps still.
constexpr int trace_min_blocks(int maxt, int bands, int stage, bool grid, int fast = 0) {
    if (maxt != 256) return 0;
    if (grid) return RAR_GRID_MIN_BLOCKS;
    if (bands == 8) return 4;
    return (bands == 1 && stage == 0) ? 4 : 0;
}

template <int BANDS, bool COUNT, bool HITS, int STAGE, int MAXT, bool COOP, bool GRID = false, bool OPAQUE = false, int FAST = 0>
__global__ void __launch_bounds__(MAXT, trace_min_blocks(MAXT, BANDS, STAGE, GRID, FAST)) trace_deposit_kernel(const __grid_constant__ TraceLaunch a) {
    static_assert(FAST == 0 || FAST == 4 || (OPAQUE && !COUNT && !HITS && !GRID && !COOP && STAGE == 0), "FAST 1/3/7: production small-scene kernels");
    static_assert(FAST != 4 || (COOP && !GRID && !COUNT && !HITS), "PACKED: production kernels with cooperative shadow rays (256 walls and more)");
    static_assert(FAST <= 4 || FAST == 7, "PACKED small-scene kernel: the four-wall variant only");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const SceneView<STAGE, GRID, FAST> sc = stage_scene<STAGE, GRID, FAST>(a, smem_raw);
    SpecConsts spec_c = {0.f, 0.f};
    if (FAST & 1) spec_c = spec_consts(a.p);
    const SpecConsts *sp = (FAST & 1) ? &spec_c : nullptr;

    const unsigned lane = threadIdx.x & 31u;
    const long long rays_per_frame = a.ray_end - a.ray_begin;
    const long long n_rays = rays_per_frame * a.n_frames;  // frames are laid out one after the other
    const int max_b = a.p.max_bounce_count;
    RayCounters ctr = {0, 0, 0, 0, 0};

    // Tiles of 32 consecutive rays, one warp each.  A warp's first tile is its own global index; further tiles are
    // claimed from a counter in global memory (a.tile_counter, zeroed before the launch) instead of a fixed stride: rays
    // differ in cost (a maze ray's shadow walks end anywhere between wall 1 and wall 10 000), so a fixed assignment
    // leaves the slowest warp of the launch behind the mean.  Claiming per WARP keeps the warps of a CTA independent
    // of each other (a per-CTA claim needs a block-wide barrier per tile, which cost the 1024-thread kernels 6 %).
    const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long tile = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile * 32 < n_rays;) {
        const long long idx = tile * 32 + lane;
        long long next_tile = tile + warps_total;
        if (a.tile_counter != nullptr && lane == 0)  // claimed now, used after this tile: the atomic's latency is hidden
            next_tile = warps_total + (long long)atomicAdd(a.tile_counter, 1ull);
        bool alive = idx < n_rays;
        const uint32_t frame = a.n_frames > 1 ? (uint32_t)(idx / rays_per_frame) : 0u;
        uint32_t id = (uint32_t)(a.ray_begin + (a.n_frames > 1 ? idx - (long long)frame * rays_per_frame : idx));
        if (a.cyc_world > 1) {
            // block-cyclic shard (rar_trace_interleaved): chunk k of this launch is chunk k * world + rank of the dispatch
            const long long g = (((idx >> a.cyc_shift) * a.cyc_world + a.cyc_rank) << a.cyc_shift) + (idx & ((1LL << a.cyc_shift) - 1));
            alive = alive && g < a.cyc_total;
            id = (uint32_t)g;
        }
        RayState<BANDS> r;
        if (alive) ray_init(r, id, a.p, frame);

        // debugRays (Raytrace2D.compute:63,87,96): thread ids < 100 record wall hits (hard-coded in the shader),
        // ids < debugRayCount record the escape vertex.  dbg_flags stays 0 for every other thread, and the vertex
        // address is formed only where a vertex is recorded, so the bounce loop pays one predicated test for it.
        f4 *dbg = a.debug_rays;
        int dbg_flags = 0;
        if (a.debug_rays != nullptr && alive && frame == 0) {
            const long long row = (long long)id * (max_b + 1);
            dbg_flags = (id < 100u ? 1 : 0) | (id < (uint32_t)a.debug_ray_count ? 2 : 0);
            // every row of the buffer is re-initialised by the thread of that ray (rows exist for ids below
            // max(100, debugRayCount)): vertices this frame does not reach read as zero, and no separate clear of the
            // buffer is enqueued before the launch
            if (row + max_b < a.debug_capacity) {
                f4 *rowp = a.debug_rays + row;
                for (int k = 0; k <= max_b; k++) rowp[k] = f4{0.f, 0.f, 0.f, 0.f};
            } else {
                dbg_flags = 0;
            }
            if (dbg_flags) {
                dbg = a.debug_rays + row;
                if (dbg_flags & 1) dbg[0] = f4{r.px, r.py, r.energy, 0.0f};
            }
        }

        for (int i = 0; i < max_b; i++) {
            if (!__any_sync(kFull, alive)) break;
            Arrival<BANDS> direct, nee;
            BounceCtx<BANDS> c;
            direct.has = 0;
            nee.has = 0;
            c.want_shadow = 0;
            c.shadow = ShadowRay{0.f, 0.f, 0.f, 0.f, 0.f};
            if constexpr (COOP || GRID) {
                // the warp meets between the phases: for the cooperative shadow rays, and in grid mode because
                // reconverging before the (long, divergent) cell walks measured 3 % faster than one merged region
                bool hit_wall = false;
                if (alive) {
                    hit_wall = bounce_begin<BANDS, COUNT>(sc, a.p, r, direct, c, &ctr, dbg + (i + 1), dbg_flags, sp);
                }
                const bool pending = hit_wall && c.want_shadow;
                bool visible = true;
                if (COOP) {
                    __syncwarp();
                    visible = coop_shadow<COUNT>(sc, pending, c.shadow, lane, ctr);
                } else if (pending) {
                    int tests = 0;
                    visible = check_vis(sc, c.shadow, COUNT ? &tests : nullptr);
                    if (COUNT) ctr.shadow_tests += (unsigned long long)tests;
                }
                if (alive) alive = hit_wall && bounce_finish<BANDS, COUNT, OPAQUE>(sc, a.p, r, nee, c, visible, &ctr);
            } else if (alive) {
                // per-thread shadow rays: the three phases of a live ray run inside one divergent region
                alive = bounce_begin<BANDS, COUNT>(sc, a.p, r, direct, c, &ctr, dbg + (i + 1), dbg_flags, sp);
                if (!HITS) deposit_direct<BANDS>(a, a.hist, direct);
                if (alive) {
                    bool visible = true;
                    if (c.want_shadow) {
                        int tests = 0;
                        visible = check_vis(sc, c.shadow, COUNT ? &tests : nullptr);
                        if (COUNT) ctr.shadow_tests += (unsigned long long)tests;
                    }
                    alive = bounce_finish<BANDS, COUNT, OPAQUE>(sc, a.p, r, nee, c, visible, &ctr);
                }
            }
            if (HITS) __syncwarp();  // the deposit starts with a full-mask vote, which reconverges the warp by itself
            if (HITS) {
                emit_hit(a, direct, id, i, 0);
                emit_hit(a, nee, id, i, 1);
            } else {
                if (COOP || GRID) deposit_hist<BANDS, true>(a, a.hist, direct, lane);
                deposit_hist<BANDS, (STAGE != 0 || GRID)>(a, a.hist, nee, lane);
            }
        }
        tile = a.tile_counter != nullptr ? __shfl_sync(kFull, next_tile, 0) : next_tile;
    }

    if (COUNT && a.counters != nullptr) {
        unsigned long long v[5] = {ctr.ray_bounces, ctr.nearest_tests, ctr.shadow_tests, ctr.direct_hits, ctr.nee_hits};
#pragma unroll
        for (int k = 0; k < 5; k++) {
            unsigned long long s = warp_sum_u64(v[k]);
            if (lane == 0 && s != 0) atomicAdd(a.counters + k, s);
        }
    }
}

// ---- batched listeners (BASELINE config 4) ---------------------------------------------------------------
//
// The ray's path does not depend on the listener, so each ray is traced ONCE; per bounce the kernel runs the
// nearest-hit loop once and then, for every listener, only the listener pieces: the direct crossing test
// before the advance, and the next-event estimate with its shadow ray after it.  Listener l deposits into its
// own histogram.  Result per listener: identical to a single-listener trace (same pieces, same order).
template <bool COUNT, int STAGE, int MAXT, bool COOP, bool GRID = false, bool OPAQUE = false, int FAST = 0>
__global__ void __launch_bounds__(MAXT) trace_listeners_kernel(const __grid_constant__ TraceLaunch a) {
    static_assert(FAST == 0 || (FAST == 4 && STAGE == 1 && COOP && !GRID && !COUNT), "PACKED: production staged-endpoint kernels");
    constexpr int BANDS = 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const SceneView<STAGE, GRID, FAST> sc = stage_scene<STAGE, GRID, FAST>(a, smem_raw);

    const unsigned lane = threadIdx.x & 31u;
    const long long rays_per_frame = a.ray_end - a.ray_begin;
    const long long n_rays = rays_per_frame * a.n_frames;
    const int max_b = a.p.max_bounce_count;
    const int n_l = a.n_listeners;
    RayCounters ctr = {0, 0, 0, 0, 0};

    for (long long base = (long long)blockIdx.x * blockDim.x; base < n_rays; base += (long long)gridDim.x * blockDim.x) {
        const long long idx = base + threadIdx.x;
        bool alive = idx < n_rays;
        const uint32_t frame = a.n_frames > 1 ? (uint32_t)(idx / rays_per_frame) : 0u;
        const uint32_t id = (uint32_t)(a.ray_begin + (a.n_frames > 1 ? idx - (long long)frame * rays_per_frame : idx));
        RayState<BANDS> r;
        if (alive) ray_init(r, id, a.p, frame);

        for (int i = 0; i < max_b; i++) {
            if (!__any_sync(kFull, alive)) break;
            BounceCtx<BANDS> c;
            c.hit = -1;
            c.closest = kInf;
            if (alive) bounce_nearest<BANDS, COUNT>(sc, r, c, &ctr);
            __syncwarp();
            for (int l = 0; l < n_l; l++) {  // direct crossings, state before the advance
                const float2 lp = __ldg(reinterpret_cast<const float2 *>(a.listeners) + l);
                Arrival<BANDS> direct;
                direct.has = 0;
                if (alive) listener_direct<BANDS, COUNT>(a.p, lp.x, lp.y, r, c.closest, direct, &ctr);
                deposit_hist<BANDS, true>(a, a.listener_hists[l], direct, lane);
            }
            const bool hit_wall = alive && bounce_advance(sc, r, c, nullptr, 0);
            for (int l = 0; l < n_l; l++) {  // next-event estimates from the hit point
                const float2 lp = __ldg(reinterpret_cast<const float2 *>(a.listeners) + l);
                c.want_shadow = 0;
                c.nee_candidate = 0;
                c.shadow = ShadowRay{0.f, 0.f, 0.f, 0.f, 0.f};
                if (hit_wall) listener_nee<BANDS, COUNT>(a.p, lp.x, lp.y, r, c);
                const bool pending = hit_wall && c.want_shadow;
                bool visible = true;
                if (COOP) {
                    __syncwarp();
                    visible = coop_shadow<COUNT>(sc, pending, c.shadow, lane, ctr);
                } else if (pending) {
                    int tests = 0;
                    visible = check_vis(sc, c.shadow, COUNT ? &tests : nullptr);
                    if (COUNT) ctr.shadow_tests += (unsigned long long)tests;
                }
                Arrival<BANDS> nee;
                nee.has = 0;
                if (hit_wall) nee_arrival<BANDS, COUNT>(r, c, visible, nee, &ctr, &a.p);
                __syncwarp();
                deposit_hist<BANDS, true>(a, a.listener_hists[l], nee, lane);
            }
            if (alive) alive = hit_wall && bounce_scatter<BANDS, OPAQUE>(a.p, r, c);
        }
    }

    if (COUNT && a.counters != nullptr) {
        unsigned long long v[5] = {ctr.ray_bounces, ctr.nearest_tests, ctr.shadow_tests, ctr.direct_hits, ctr.nee_hits};
#pragma unroll
        for (int k = 0; k < 5; k++) {
            unsigned long long s = warp_sum_u64(v[k]);
            if (lane == 0 && s != 0) atomicAdd(a.counters + k, s);
        }
    }
}

// Register-only FFMA loop: the FP32 issue peak the ray stage is measured against.
__global__ void __launch_bounds__(256) fp32_peak_kernel(float *sink, int iters) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.0f, x2 = x0 + 2.0f, x3 = x0 + 3.0f;
    float x4 = x0 + 4.0f, x5 = x0 + 5.0f, x6 = x0 + 6.0f, x7 = x0 + 7.0f;
    const float m = 0.999f, c = 1e-4f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            x0 = __fmaf_rn(x0, m, c); x1 = __fmaf_rn(x1, m, c); x2 = __fmaf_rn(x2, m, c); x3 = __fmaf_rn(x3, m, c);
            x4 = __fmaf_rn(x4, m, c); x5 = __fmaf_rn(x5, m, c); x6 = __fmaf_rn(x6, m, c); x7 = __fmaf_rn(x7, m, c);
        }
    }
    float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 12345.678f) sink[0] = s;
}

// Bit-for-bit comparison of the range-checked-once operations (rar_math.cuh) with the correctly rounded intrinsics
// on operands drawn over the whole admitted range: mism[0] reciprocal, [1] square root, [2] division,
// [3] division by a launch-invariant divisor through its refined reciprocal, [4] in_safe_range at its edges.
__device__ __forceinline__ float selftest_operand(uint32_t &st, int e_lo, int e_hi, bool allow_negative) {
    st = st * 747796405u + 2891336453u;
    const uint32_t h = ((st >> ((st >> 28) + 4u)) ^ st) * 277803737u;
    const uint32_t v = (h >> 22) ^ h;
    const uint32_t e = (uint32_t)(e_lo + 127) + (v >> 9) % (uint32_t)(e_hi - e_lo);  // biased exponent in [e_lo, e_hi)
    st = st * 747796405u + 2891336453u;
    const uint32_t m = ((st >> 7) ^ st) & 0x7fffffu;
    const uint32_t sign = (allow_negative && (v & 1u)) ? 0x80000000u : 0u;
    return __uint_as_float(sign | (e << 23) | ((v & 6u) == 6u ? 0u : m));  // one operand in four is a power of two
}
__global__ void __launch_bounds__(256) arithmetic_selftest_kernel(long long n, uint32_t seed, unsigned long long *mism) {
    unsigned long long bad[5] = {0, 0, 0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        uint32_t st = seed + (uint32_t)i * 2654435761u + (uint32_t)(i >> 32);
        const float x = selftest_operand(st, -100, 100, true);
        if (__float_as_uint(rcp_inrange(x)) != __float_as_uint(__frcp_rn(x))) bad[0]++;
        const float y = fabsf(selftest_operand(st, -100, 100, false));
        if (__float_as_uint(sqrt_inrange(y)) != __float_as_uint(__fsqrt_rn(y))) bad[1]++;
        // divisions: quotient magnitude kept within 2^+-120
        const float b = selftest_operand(st, -100, 100, true);
        float a = selftest_operand(st, -100, 100, true);
        const int eb = (int)((__float_as_uint(b) >> 23) & 0xffu) - 127, ea = (int)((__float_as_uint(a) >> 23) & 0xffu) - 127;
        if (ea - eb > 119 || ea - eb < -119) a = __uint_as_float((__float_as_uint(a) & 0x807fffffu) | ((uint32_t)(eb + 127) << 23));
        if ((i & 1023) == 0) a = 0.0f;
        // (a zero dividend gives a zero whose sign may differ from IEEE's: +0 / -b is -0, the FMA chain returns +0;
        //  the ray code only ever compares such a quotient with eps)
        const uint32_t qa = __float_as_uint(div_inrange(a, b)), qb = __float_as_uint(__fdiv_rn(a, b));
        if (qa != qb && !(a == 0.0f && ((qa | qb) & 0x7fffffffu) == 0u)) bad[2]++;
        const float c = fabsf(selftest_operand(st, -20, 20, false));
        const float t = fabsf(selftest_operand(st, -14, 27, false));  // 1e-4 .. 1e8, the distances divided by the speed
        const float inv_c = rcp_refine(c, mufu_rcp(c));
        if (__float_as_uint(div_with_rcp(t, c, inv_c)) != __float_as_uint(__fdiv_rn(t, c))) bad[3]++;
        // in_safe_range against its definition, around both edges and on specials
        const uint32_t edge[8] = {0x0d7fffffu, 0x0d800000u, 0x71800000u, 0x71800001u, 0x00000000u, 0x7f800000u, 0x7fc00000u, 0x8d800000u};
        const float z = __uint_as_float(edge[i & 7]);
        const bool want = z >= 7.888609052210118e-31f && z <= 1.2676506002282294e30f;
        if (in_safe_range(z) != want || !in_safe_range(y)) bad[4]++;
    }
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const unsigned long long s = warp_sum_u64(bad[k]);
        if ((threadIdx.x & 31u) == 0 && s != 0) atomicAdd(mism + k, s);
    }
}

// ---- launch selection -------------------------------------------------------------------------------

constexpr int kCoopMinWalls = 256;  // from this many walls shadow rays are resolved warp-cooperatively

struct KernelChoice {
    const void *fn;
    int max_threads;
};

// cudaFuncSetAttribute + cudaOccupancyMaxActiveBlocksPerMultiprocessor cost microseconds each; a 15 000-ray
// frame (config 1) runs for ~25 us, so the answers are remembered per (device, kernel, block size, smem).
cudaError_t resident_blocks(const void *fn, int threads, size_t smem, int *blocks) {
    static std::mutex mu;
    static std::map<std::tuple<int, const void *, int, size_t>, int> cache;
    static std::map<std::pair<int, const void *>, size_t> opted_in;  // largest dynamic smem enabled so far
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    size_t &have = opted_in[std::make_pair(dev, fn)];
    if (smem > have) {  // the attribute is a per-kernel maximum: only ever raise it
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        have = smem;
    }
    const auto key = std::make_tuple(dev, fn, threads, smem);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *blocks = it->second;
        return cudaSuccess;
    }
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, fn, threads, smem);
    if (e == cudaSuccess) cache[key] = *blocks;
    return e;
}

// Small scenes (all planes in shared memory, several CTAs per SM): per-thread shadow walk, 256 threads.
// Large scenes: warp-cooperative shadow rays; one CTA of 1024 threads per SM once the endpoint plane
// takes more than half of shared memory.
// stage 1 production kernels with the packed-FP32 wall scans
template <int BANDS>
KernelChoice pick_packed(int stage, bool big_block, bool opaque) {
    if (stage == 0) {  // everything staged, 256 walls and more (cooperative shadow rays)
        if (opaque) return {(const void *)trace_deposit_kernel<BANDS, false, false, 0, 256, true, false, true, 4>, 256};
        return {(const void *)trace_deposit_kernel<BANDS, false, false, 0, 256, true, false, false, 4>, 256};
    }
    if (stage == 2) return {(const void *)trace_deposit_kernel<BANDS, false, false, 2, 256, true, false, false, 4>, 256};
    if (big_block) return {(const void *)trace_deposit_kernel<BANDS, false, false, 1, 1024, true, false, false, 4>, 1024};
    return {(const void *)trace_deposit_kernel<BANDS, false, false, 1, 256, true, false, false, 4>, 256};
}
template <int BANDS, bool COUNT, bool HITS>
KernelChoice pick_kernel(int stage, bool big_block, bool coop) {
    if (stage == 0) {
        if (coop) return {(const void *)trace_deposit_kernel<BANDS, COUNT, HITS, 0, 256, true>, 256};
        return {(const void *)trace_deposit_kernel<BANDS, COUNT, HITS, 0, 256, false>, 256};
    }
    if (stage == 1) {
        if (big_block) return {(const void *)trace_deposit_kernel<BANDS, COUNT, HITS, 1, 1024, true>, 1024};
        return {(const void *)trace_deposit_kernel<BANDS, COUNT, HITS, 1, 256, true>, 256};
    }
    return {(const void *)trace_deposit_kernel<BANDS, COUNT, HITS, 2, 256, true>, 256};
}
// production small-scene kernels without the transmit/refract branch; fast: 0, 1 (SPEC: range-checked-once
// arithmetic, when the host has validated the operand ranges) or 3 (SPEC + exactly four walls)
template <int BANDS>
KernelChoice pick_kernel_opaque(bool coop, int fast) {
    if (coop) return {(const void *)trace_deposit_kernel<BANDS, false, false, 0, 256, true, false, true>, 256};
    if (fast == 7) return {(const void *)trace_deposit_kernel<BANDS, false, false, 0, 256, false, false, true, 7>, 256};
    if (fast == 3) return {(const void *)trace_deposit_kernel<BANDS, false, false, 0, 256, false, false, true, 3>, 256};
    if (fast == 1) return {(const void *)trace_deposit_kernel<BANDS, false, false, 0, 256, false, false, true, 1>, 256};
    return {(const void *)trace_deposit_kernel<BANDS, false, false, 0, 256, false, false, true>, 256};
}
template <bool COUNT>
KernelChoice pick_listeners(int stage, bool big_block, bool coop) {
    if (stage == 0) {
        if (coop) return {(const void *)trace_listeners_kernel<COUNT, 0, 256, true>, 256};
        return {(const void *)trace_listeners_kernel<COUNT, 0, 256, false>, 256};
    }
    if (stage == 1) {
        if (big_block) return {(const void *)trace_listeners_kernel<COUNT, 1, 1024, true>, 1024};
        return {(const void *)trace_listeners_kernel<COUNT, 1, 256, true>, 256};
    }
    return {(const void *)trace_listeners_kernel<COUNT, 2, 256, true>, 256};
}
// The OPAQUE instantiations exist for the production mode only (no test counters, no hit list).
template <int BANDS>
KernelChoice pick_mode(bool count, bool hits, bool opaque, int stage, bool big, bool coop, int fast, bool packed) {
    if (packed && coop && !count && !hits) return pick_packed<BANDS>(stage, big, opaque);
    if (hits) return count ? pick_kernel<BANDS, true, true>(stage, big, coop) : pick_kernel<BANDS, false, true>(stage, big, coop);
    if (count) return pick_kernel<BANDS, true, false>(stage, big, coop);
    // Large scenes spend their time in the wall loops; there the smaller scatter code changes nothing measurable (the
    // 10k-wall maze was 2 % slower with it), so only the small-scene (stage 0) and grid kernels have the variant.
    if (opaque && stage == 0) return pick_kernel_opaque<BANDS>(coop, fast);
    return pick_kernel<BANDS, false, false>(stage, big, coop);
}
KernelChoice pick_listeners_mode(bool count, bool opaque, int stage, bool big, bool coop, bool packed) {
    if (packed && stage == 1 && !count) {
        if (big) return {(const void *)trace_listeners_kernel<false, 1, 1024, true, false, false, 4>, 1024};
        return {(const void *)trace_listeners_kernel<false, 1, 256, true, false, false, 4>, 256};
    }
    if (count) return pick_listeners<true>(stage, big, coop);
    (void)opaque;  // the listener kernels spend their time in per-listener work, not in the scatter step
    return pick_listeners<false>(stage, big, coop);
}
template <int BANDS>
KernelChoice pick_grid(bool count, bool opaque) {
    if (count) return {(const void *)trace_deposit_kernel<BANDS, true, false, 2, 256, false, true, false>, 256};
    if (opaque) return {(const void *)trace_deposit_kernel<BANDS, false, false, 2, 256, false, true, true>, 256};
    return {(const void *)trace_deposit_kernel<BANDS, false, false, 2, 256, false, true, false>, 256};
}
KernelChoice pick_grid_listeners(bool count, bool opaque) {
    if (count) return {(const void *)trace_listeners_kernel<true, 2, 256, false, true, false>, 256};
    return {(const void *)trace_listeners_kernel<false, 2, 256, false, true, false>, 256};
}

}  // namespace

cudaError_t launch_trace(const TraceLaunch &a, bool count_tests, const DeviceFacts &dev, cudaStream_t stream,
                         int *launches) {
    const long long n_rays = (a.ray_end - a.ray_begin) * (a.n_frames > 1 ? a.n_frames : 1);
    if (n_rays <= 0 || a.p.max_bounce_count <= 0) return cudaSuccess;
    if (a.bands != 1 && a.bands != 8) return cudaErrorInvalidValue;
    if (a.n_listeners > 0 && (a.bands != 1 || a.hits != nullptr)) return cudaErrorInvalidValue;

    // Candidate configurations, by shared-memory footprint (16-byte barrier slot + planes):
    //   stage 0 / 256 threads : all three planes on chip;
    //   stage 1 / 256 threads : endpoint plane on chip, materials read through L1 once per bounce;
    //   stage 1 / 1024 threads: one CTA per SM when the endpoint plane needs most of shared memory;
    //   stage 2 / 256 threads : planes too large for shared memory, broadcast loads served by L1/L2
    //                           (measured as fast per test as the staged path: the loads are broadcast
    //                           and 32 resident warps hide their latency).
    // The one that keeps the most threads resident per SM wins (ties: lower stage).
    const size_t geo_bytes = (size_t)a.n_walls * 16;
    const size_t m1_bytes = ((size_t)a.n_walls * 8 + 15) & ~(size_t)15;
    const size_t budget = (size_t)dev.smem_optin - 1024;
    const bool hits = a.hits != nullptr;
    // a.spec_ok: the host has validated the operand ranges the SPEC arithmetic assumes (rar2d_api.cu spec_ranges_ok);
    // RAR_NO_FAST=1 in the environment keeps the guarded kernels (A/B measurements, bisecting a parity failure)
    const char *nf = getenv("RAR_NO_FAST");
    const bool no_fast = nf != nullptr && nf[0] == '1';
    const int fast = (a.spec_ok && !no_fast) ? (a.n_walls == 4 ? 3 : 1) : 0;
    // PACKED: the staged-endpoint production kernels scan two walls per packed FP32 instruction (RAR_NO_PACKED=1: scalar)
    const char *np_env = getenv("RAR_NO_PACKED");
    const bool packed = !(np_env != nullptr && np_env[0] == '1') && !count_tests && !hits && a.pair_a != nullptr;
    const size_t stage1_bytes = packed ? (size_t)((a.n_walls + 1) / 2) * 32 : geo_bytes;
    const int fast0 = (fast == 3 && packed) ? 7 : fast;  // the four-wall kernel with packed wall tests (pair planes staged too)
    // stage 0 stages the pair planes behind the others when its kernel is a packed one: four walls, or cooperative (256+ walls)
    const bool packed0 = fast0 == 7 || (packed && a.n_walls >= kCoopMinWalls && a.n_listeners == 0);
    struct Cand { int stage; bool big; size_t smem; };
    const Cand cands[4] = {{0, false, 16 + 2 * geo_bytes + m1_bytes + (packed0 ? (size_t)((a.n_walls + 1) / 2) * 32 : 0)}, {1, false, 16 + stage1_bytes}, {1, true, 16 + stage1_bytes}, {2, false, 16}};
    KernelChoice k{nullptr, 0};
    size_t smem = 0;
    bool big_block = false;
    int best_resident = -1, per_sm = 0;
    if (a.use_grid) {
        // Grid walks touch a few walls per query at data-dependent addresses: no staging, per-thread shadow walk.
        if (a.hits != nullptr) return cudaErrorInvalidValue;
        const bool opq = a.opaque != 0;
        if (a.n_listeners > 0) k = pick_grid_listeners(count_tests, opq);
        else k = a.bands == 8 ? pick_grid<8>(count_tests, opq) : pick_grid<1>(count_tests, opq);
        smem = 16;
        cudaError_t e = resident_blocks(k.fn, k.max_threads, smem, &per_sm);
        if (e != cudaSuccess) return e;
        best_resident = per_sm * k.max_threads;
    }
    for (const Cand &c : cands) {
        if (a.use_grid) break;
        if (c.smem > budget) continue;
        // The stage 1/2 kernels resolve shadow rays warp-cooperatively (one ray at a time), which is 3x slower than
        // the per-thread walk on a handful of walls (measured: per-thread wins at 64 and 128 walls, a tie at 256-1000,
        // cooperative wins beyond): small scenes always take stage 0, whatever the occupancy says.
        if (a.n_walls < kCoopMinWalls && c.stage != 0) continue;
        const bool coop = c.stage != 0 || a.n_walls >= kCoopMinWalls;
        KernelChoice kc;
        const bool opq = a.opaque != 0;
        if (a.n_listeners > 0) kc = pick_listeners_mode(count_tests, opq, c.stage, c.big, coop, packed);
        else kc = a.bands == 8 ? pick_mode<8>(count_tests, hits, opq, c.stage, c.big, coop, fast0, packed)
                               : pick_mode<1>(count_tests, hits, opq, c.stage, c.big, coop, fast0, packed);
        int blocks = 0;
        cudaError_t e = resident_blocks(kc.fn, kc.max_threads, c.smem, &blocks);
        if (e != cudaSuccess) return e;
        const int resident = blocks * kc.max_threads;
        if (resident > best_resident) {
            best_resident = resident;
            k = kc;
            smem = c.smem;
            big_block = c.big;
            per_sm = blocks;
        }
    }
    if (best_resident <= 0) return cudaErrorLaunchOutOfResources;

    // Block size: large enough to amortise the staging, small enough that small dispatches still
    // cover the machine (config 1 traces only 15 040 rays).
    int threads = k.max_threads;
    if (!big_block) {
        while (threads > 64 && (n_rays + threads - 1) / threads < 2LL * dev.sm_count) threads >>= 1;
        if (threads != k.max_threads) {
            cudaError_t e = resident_blocks(k.fn, threads, smem, &per_sm);
            if (e != cudaSuccess) return e;
            if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        }
    }
    long long want = (n_rays + threads - 1) / threads;
    long long cap = (long long)per_sm * dev.sm_count;
    int grid = (int)(want < cap ? want : cap);
    cudaError_t e;

    TraceLaunch arg = a;
    if (arg.n_frames < 1) arg.n_frames = 1;
    if (arg.band_total < 1) {  // callers that do not chunk: the slot has exactly `bands` bands
        arg.band_total = arg.bands;
        arg.band_offset = 0;
        arg.band_valid = arg.bands;
    }
    if (arg.n_listeners > 0 || want <= cap) arg.tile_counter = nullptr;  // one tile per warp (or the listener kernel): nothing to claim
    if (const char *nd = getenv("RAR_NO_DYNAMIC_TILES"))                  // A/B measurements: fixed stride over the tiles
        if (nd[0] == '1') arg.tile_counter = nullptr;
    if (arg.tile_counter != nullptr) {
        e = cudaMemsetAsync(arg.tile_counter, 0, sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return e;
    }
    void *params[] = {&arg};
    e = cudaLaunchKernel(k.fn, dim3(grid), dim3(threads), params, smem, stream);
    if (e == cudaSuccess && launches) ++*launches;
    return e;
}

cudaError_t launch_arithmetic_selftest(long long n, uint32_t seed, unsigned long long *d_mism, int blocks, cudaStream_t stream) {
    arithmetic_selftest_kernel<<<blocks, 256, 0, stream>>>(n, seed, d_mism);
    return cudaGetLastError();
}

cudaError_t launch_fp32_peak(float *d_sink, int blocks, int threads, int iters, cudaStream_t stream) {
    fp32_peak_kernel<<<blocks, threads, 0, stream>>>(d_sink, iters);
    return cudaGetLastError();
}

}  // namespace rar
