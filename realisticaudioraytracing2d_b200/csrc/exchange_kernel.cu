// exchange_kernel.cu -- the histogram all-reduce of the ray-range sharding as ONE kernel over NVLink peer memory,
// for hosts that run one process per GPU (rar_exchange_* in include/rar2d.h).
//
// Every rank owns an "exchange region" (one cudaMalloc, mapped into the other processes through CUDA IPC):
//
//     flags   [2 phases][kExMaxBlocks][kExMaxRanks] u32   written by the peers, spun on by the owner
//     status  u32                                          set when a barrier timed out
//     in [2]  capacity words each, double-buffered by call parity: this rank's histogram, staged for its peers
//     out[2]  capacity words each: totals written by the slice owners (two-shot mode)
//
// One call = one launch per rank, no host synchronisation and no collective library:
//   phase 0  copy the local histogram into in[parity]                       (local HBM)
//   barrier  block b of every rank signals block b of every peer and waits  (st.release.sys / ld.acquire.sys)
//   one-shot: hist[v] = sum over ranks of in_r[v]                           (world-1 remote reads of everything)
//   two-shot: for the slice this rank owns, total = sum over ranks of in_r[v]; total -> every peer's out[parity]
//             and the local hist; second barrier; copy the other slices' totals from the local out[parity]
// Work is mapped to blocks so that block b only ever reads what block b of a peer wrote, so the barriers are
// per block (no grid-wide synchronisation, no co-residency requirement).  The double buffering makes a trailing
// barrier unnecessary: a rank can only be overwriting in[parity] of call e+2 after every peer has entered call
// e+1, i.e. finished reading call e.  Sums are 64-bit integers, so the result is bit-identical to any other
// reduction order (NCCL, the single-process peer_reduce_kernel, a one-GPU trace).
#include <cuda_runtime.h>

#include "rar_internal.h"

namespace rar {

namespace {

constexpr int kExThreads = 256;

struct alignas(16) Words2 { unsigned long long a, b; };

__device__ __forceinline__ Words2 ld_sys(const long long *p) {
    Words2 v;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.a), "=l"(v.b) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(long long *p, Words2 v) {
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.a), "l"(v.b) : "memory");
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Block b of this rank tells block b of every peer that it reached `phase` of call `epoch`, then waits for the
// same word from every peer.  Returns false when a peer did not arrive within the time limit.
__device__ bool peer_barrier(const ExchangeLaunch &x, int phase, int *s_failed) {
    __syncthreads();
    if ((int)threadIdx.x < x.world) {
        const int p = threadIdx.x;
        const size_t row = ((size_t)phase * kExMaxBlocks + blockIdx.x) * kExMaxRanks;
        // release at system scope, cumulative over the block's writes ordered before it by the barrier above
        st_release_sys(x.flags[p] + row + x.rank, x.epoch);
        const unsigned *mine = x.flags[x.rank] + row + p;
        const unsigned long long t0 = global_ns();
        unsigned spins = 0;
        while ((int)(ld_acquire_sys(mine) - x.epoch) < 0) {
            if ((++spins & 255u) == 0 && global_ns() - t0 > x.timeout_ns) {
                atomicExch(x.status, 1u);
                *s_failed = 1;
                break;
            }
        }
    }
    __syncthreads();
    return *s_failed == 0;
}

// Sum of vector v over every rank's staged histogram: the loads of eight ranks are issued back to back before
// the first add so that the NVLink round trips overlap.
__device__ __forceinline__ Words2 sum_over_ranks(const ExchangeLaunch &x, long long v) {
    Words2 acc{0, 0};
    for (int r0 = 0; r0 < x.world; r0 += 8) {
        Words2 t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            t[k] = Words2{0, 0};
            if (r0 + k < x.world) t[k] = ld_sys(x.stage_in[r0 + k] + 2 * v);
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            acc.a += t[k].a;
            acc.b += t[k].b;
        }
    }
    return acc;
}

__global__ void __launch_bounds__(kExThreads) exchange_allreduce_kernel(const __grid_constant__ ExchangeLaunch x) {
    __shared__ int s_failed;
    if (threadIdx.x == 0) s_failed = 0;
    const long long nvs = x.slice_vecs;  // 16-byte vectors per rank slice; world * nvs >= n_vecs
    const long long step = (long long)gridDim.x * kExThreads;
    const long long first = (long long)blockIdx.x * kExThreads + threadIdx.x;
    long long *hist = x.hist;

    long long *my_in = x.stage_in[x.rank];
    if (!x.two_shot) {
        // one-shot: flat mapping, block b touches the same vectors on every rank
        for (long long v = first; v < x.n_vecs; v += step)
            *reinterpret_cast<Words2 *>(my_in + 2 * v) = *reinterpret_cast<const Words2 *>(hist + 2 * v);
        if (!peer_barrier(x, 0, &s_failed)) return;
        for (long long v = first; v < x.n_vecs; v += step) {
            const Words2 acc = sum_over_ranks(x, v);
            *reinterpret_cast<Words2 *>(hist + 2 * v) = acc;
        }
        return;
    }

    // two-shot, phase 0: stage the local histogram (the unit set of block b is the same in every slice and on
    // every rank, so block b of the slice owner reads only what block b of each peer staged)
    for (int s = 0; s < x.world; s++)
        for (long long u = first; u < nvs; u += step) {
            const long long v = (long long)s * nvs + u;
            if (v < x.n_vecs) *reinterpret_cast<Words2 *>(my_in + 2 * v) = *reinterpret_cast<const Words2 *>(hist + 2 * v);
        }
    if (!peer_barrier(x, 0, &s_failed)) return;

    // reduce the slice this rank owns and scatter the totals
    for (long long u = first; u < nvs; u += step) {
        const long long v = (long long)x.rank * nvs + u;
        if (v >= x.n_vecs) continue;
        const Words2 acc = sum_over_ranks(x, v);
        for (int r = 0; r < x.world; r++)
            if (r != x.rank) st_sys(x.stage_out[r] + 2 * v, acc);
        *reinterpret_cast<Words2 *>(hist + 2 * v) = acc;
    }
    if (!peer_barrier(x, 1, &s_failed)) return;
    const long long *my_out = x.stage_out[x.rank];
    for (int s = 0; s < x.world; s++) {
        if (s == x.rank) continue;
        for (long long u = first; u < nvs; u += step) {
            const long long v = (long long)s * nvs + u;
            if (v < x.n_vecs) *reinterpret_cast<Words2 *>(hist + 2 * v) = ld_sys(my_out + 2 * v);
        }
    }
}

}  // namespace

cudaError_t launch_exchange_allreduce(const ExchangeLaunch &x, cudaStream_t stream) {
    if (x.n_vecs <= 0 || x.world <= 1) return cudaSuccess;
    const long long per_block = x.two_shot ? x.slice_vecs : x.n_vecs;
    long long blocks = (per_block + kExThreads - 1) / kExThreads;
    if (blocks < 1) blocks = 1;
    if (blocks > kExMaxBlocks) blocks = kExMaxBlocks;
    exchange_allreduce_kernel<<<(int)blocks, kExThreads, 0, stream>>>(x);
    return cudaGetLastError();
}

}  // namespace rar
