// grid_kernel.cu -- the uniform grid of RAR_FLAG_USE_GRID built on the device.
//
// The reference re-uploads the whole wall list every FixedUpdate when obstacles move (RayTraceManager.cs:67,
// 246-250; Helpers/SceneHelper.cs:29-98), i.e. at 50 Hz.  Building the grid on the host meant a host loop over every
// (wall, cell) pair, three pageable uploads and a stream synchronisation on that path.  Here the host only computes
// the frame (one pass over the walls, rar_layout.h grid_frame); the lists are built by four kernels enqueued on the
// context's stream, with no synchronisation:
//   count  : one thread per wall walks the cells of the wall's box and counts its registrations per cell
//   scan   : exclusive prefix sum of the counts -> cell_start (one CTA; the grid has about n/2 cells)
//   fill   : the same walk again, claiming list positions with an atomic cursor per cell
//   finish : one thread per cell sorts its (short) list by wall index -- the order the brute-force scan and the host
//            builder use, and what makes test counters in grid mode reproducible -- and gathers the endpoint records
// Registration is decided by the double-precision predicates shared with the host builder (wall_cell_box /
// wall_in_cell); this unit is compiled with --fmad=false so that they round exactly as on the host, and the lists
// come out identical (tests compare digests).
#include <cuda_runtime.h>

#include "rar_internal.h"
#include "rar_layout.h"

namespace rar {
namespace {

__global__ void __launch_bounds__(128) grid_count_kernel(const f4 *__restrict__ geo, const f2 *__restrict__ end, int n, const GridFrame fr,
                                                          unsigned *__restrict__ cnt) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    const f4 g = geo[w];
    const f2 b = end[w];
    int ix0, ix1, iy0, iy1;
    wall_cell_box(fr, g.x, g.y, b.x, b.y, ix0, ix1, iy0, iy1);
    for (int iy = iy0; iy <= iy1; iy++)
        for (int ix = ix0; ix <= ix1; ix++)
            if (wall_in_cell(fr, g.x, g.y, b.x, b.y, ix, iy)) atomicAdd(cnt + (size_t)iy * fr.nx + ix, 1u);
}

// cell_start[c] = sum of cnt[0..c), cell_start[n_cells] = total; cnt is zeroed for its second life as the fill cursor.
__global__ void __launch_bounds__(1024) grid_scan_kernel(unsigned *__restrict__ cnt, unsigned *__restrict__ cell_start, int n_cells) {
    __shared__ unsigned part[1024];
    const int t = threadIdx.x;
    const int per = (n_cells + 1023) / 1024;
    const int lo = t * per, hi = min(lo + per, n_cells);
    unsigned s = 0;
    for (int c = lo; c < hi; c++) s += cnt[c];
    part[t] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {  // inclusive scan of the per-thread sums
        const unsigned v = t >= d ? part[t - d] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned run = t > 0 ? part[t - 1] : 0u;
    for (int c = lo; c < hi; c++) {
        const unsigned k = cnt[c];
        cell_start[c] = run;
        run += k;
        cnt[c] = 0u;
    }
    if (t == 1023) cell_start[n_cells] = part[1023];
}

__global__ void __launch_bounds__(128) grid_fill_kernel(const f4 *__restrict__ geo, const f2 *__restrict__ end, int n, const GridFrame fr,
                                                         const unsigned *__restrict__ cell_start, unsigned *__restrict__ cursor,
                                                         unsigned *__restrict__ items) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    const f4 g = geo[w];
    const f2 b = end[w];
    int ix0, ix1, iy0, iy1;
    wall_cell_box(fr, g.x, g.y, b.x, b.y, ix0, ix1, iy0, iy1);
    for (int iy = iy0; iy <= iy1; iy++)
        for (int ix = ix0; ix <= ix1; ix++)
            if (wall_in_cell(fr, g.x, g.y, b.x, b.y, ix, iy)) {
                const size_t cell = (size_t)iy * fr.nx + ix;
                items[cell_start[cell] + atomicAdd(cursor + cell, 1u)] = (unsigned)w;
            }
}

__global__ void __launch_bounds__(128) grid_finish_kernel(const unsigned *__restrict__ cell_start, unsigned *__restrict__ items,
                                                           f4 *__restrict__ item_geo, const f4 *__restrict__ geo, int n_cells) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const unsigned i0 = cell_start[c], i1 = cell_start[c + 1];
    for (unsigned i = i0 + 1; i < i1; i++) {  // insertion sort: the lists hold a handful of walls
        const unsigned v = items[i];
        unsigned j = i;
        while (j > i0 && items[j - 1] > v) {
            items[j] = items[j - 1];
            j--;
        }
        items[j] = v;
    }
    for (unsigned i = i0; i < i1; i++) item_geo[i] = geo[items[i]];
}

}  // namespace

cudaError_t launch_grid_build(const f4 *geo, const f2 *end, int n, const GridFrame &fr, unsigned *cnt, unsigned *cell_start,
                              unsigned *items, f4 *item_geo, cudaStream_t s) {
    if (n <= 0 || fr.nx <= 0) return cudaSuccess;
    const int n_cells = fr.nx * fr.ny;
    cudaError_t e = cudaMemsetAsync(cnt, 0, (size_t)n_cells * sizeof(unsigned), s);
    if (e != cudaSuccess) return e;
    const int wall_blocks = (n + 127) / 128;
    grid_count_kernel<<<wall_blocks, 128, 0, s>>>(geo, end, n, fr, cnt);
    grid_scan_kernel<<<1, 1024, 0, s>>>(cnt, cell_start, n_cells);
    grid_fill_kernel<<<wall_blocks, 128, 0, s>>>(geo, end, n, fr, cell_start, cnt, items);
    grid_finish_kernel<<<(n_cells + 127) / 128, 128, 0, s>>>(cell_start, items, item_geo, geo, n_cells);
    return cudaGetLastError();
}

}  // namespace rar
