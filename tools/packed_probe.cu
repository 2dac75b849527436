// packed_probe.cu -- does FFMA2 / FMUL2 / FADD2 (sm_100 packed FP32) pay in the wall-test inner loop?
// The brute-force nearest-hit scan of the trace kernel (csrc/rar_ray.cuh: wall_test + wall_pass, 4 walls per iteration,
// broadcast LDS.128 of the staged planes) in its scalar form and with two walls per packed instruction.  Both count the
// walls that pass the filter (the exact evaluation of survivors is left out: it is rare in the maze) and must agree.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/_build/packed_probe tools/packed_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { return (u64)__float_as_uint(lo) | ((u64)__float_as_uint(hi) << 32); }
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

constexpr int N = 2048;  // walls (32 KB + 32 KB of shared memory)

__global__ void __launch_bounds__(1024) scalar_scan(const float4 *geo, int reps, unsigned *out) {
    extern __shared__ float4 sg[];
    for (int i = threadIdx.x; i < N; i += blockDim.x) sg[i] = geo[i];
    __syncthreads();
    const float ang = (blockIdx.x * blockDim.x + threadIdx.x) * 1e-3f;
    const float ox = 50.f + 0.001f * threadIdx.x, oy = 50.f, dx = cosf(ang), ndy = -sinf(ang);
    float bound_m = 30.0f;
    unsigned cnt = 0;
    for (int r = 0; r < reps; r++) {
        for (int w = 0; w < N; w += 4) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float4 s = sg[w + k];
                const float v1x = ox - s.x, v1y = oy - s.y;
                const float dotP = __fmaf_rn(s.z, ndy, s.w * dx);
                const float num2 = __fmaf_rn(v1x, ndy, v1y * dx);
                const float num1 = __fmaf_rn(s.z, v1y, -(s.w * v1x));
                const float rr = bound_m * dotP;
                const bool p = (fabsf(__fmaf_rn(2.0f, num2, -dotP)) <= fabsf(dotP)) & (fabsf(__fmaf_rn(2.0f, num1, -rr)) <= fabsf(rr));
                cnt += p ? 1u : 0u;
            }
        }
        bound_m *= 0.999f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = cnt;
}

// pair planes: A[p] = {x0, x1, y0, y1}, B[p] = {z0, z1, w0, w1} of walls 2p, 2p+1
__global__ void __launch_bounds__(1024) packed_scan(const float4 *pa, const float4 *pb, int reps, unsigned *out) {
    extern __shared__ float4 sg[];
    float4 *sa = sg, *sb = sg + N / 2;
    for (int i = threadIdx.x; i < N / 2; i += blockDim.x) { sa[i] = pa[i]; sb[i] = pb[i]; }
    __syncthreads();
    const unsigned sa_addr = (unsigned)__cvta_generic_to_shared(sa), sb_addr = (unsigned)__cvta_generic_to_shared(sb);
    const float ang = (blockIdx.x * blockDim.x + threadIdx.x) * 1e-3f;
    const float ox = 50.f + 0.001f * threadIdx.x, oy = 50.f, dx = cosf(ang), ndy = -sinf(ang);
    const u64 ox2 = pk(ox, ox), oy2 = pk(oy, oy), dx2 = pk(dx, dx), ndy2 = pk(ndy, ndy), dy2 = pk(-ndy, -ndy), two = pk(2.f, 2.f);
    float bound_m = 30.0f;
    unsigned cnt = 0;
    for (int r = 0; r < reps; r++) {
        const u64 b2 = pk(bound_m, bound_m);
        for (int p = 0; p < N / 2; p += 2) {
#pragma unroll
            for (int k = 0; k < 2; k++) {
                u64 sx, sy, sz, sw;  // one LDS.128 per plane: two packed pairs each
                asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(sx), "=l"(sy) : "r"(sa_addr + 16u * (unsigned)(p + k)));
                asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(sz), "=l"(sw) : "r"(sb_addr + 16u * (unsigned)(p + k)));
                const u64 nv1x = sub2(sx, ox2);                 // -(ox - sx), exact
                const u64 v1y = sub2(oy2, sy);
                const u64 ndotP = fma2(sz, dy2, mul2(sw, pk(-dx, -dx)));   // -(sz*ndy + sw*dx): every term negated, exact
                const u64 num2 = fma2(nv1x, dy2, mul2(v1y, dx2));          // v1x*ndy + v1y*dx
                const u64 num1 = fma2(sz, v1y, mul2(sw, nv1x));            // sz*v1y - sw*v1x
                const u64 c1 = fma2(two, num2, ndotP);                     // 2 num2 - dotP
                const u64 nr = mul2(b2, ndotP);                            // -(bound_m * dotP)
                const u64 c2b = fma2(two, num1, nr);                       // 2 num1 - r
                const bool p0 = (fabsf(lo(c1)) <= fabsf(lo(ndotP))) & (fabsf(lo(c2b)) <= fabsf(lo(nr)));
                const bool p1 = (fabsf(hi(c1)) <= fabsf(hi(ndotP))) & (fabsf(hi(c2b)) <= fabsf(hi(nr)));
                cnt += (p0 ? 1u : 0u) + (p1 ? 1u : 0u);
            }
        }
        bound_m *= 0.999f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = cnt;
}

template <class F>
static double time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount, reps = 64;
    float4 *h = (float4 *)malloc(N * sizeof(float4)), *ha = (float4 *)malloc(N / 2 * sizeof(float4)), *hb = (float4 *)malloc(N / 2 * sizeof(float4));
    srand(1);
    for (int i = 0; i < N; i++) { float x = rand() % 1000 * 0.1f, y = rand() % 1000 * 0.1f; h[i] = make_float4(x, y, (rand() % 200 - 100) * 0.05f, (rand() % 200 - 100) * 0.05f); }
    for (int p = 0; p < N / 2; p++) { ha[p] = make_float4(h[2*p].x, h[2*p+1].x, h[2*p].y, h[2*p+1].y); hb[p] = make_float4(h[2*p].z, h[2*p+1].z, h[2*p].w, h[2*p+1].w); }
    float4 *d, *da, *db; unsigned *o1, *o2;
    cudaMalloc(&d, N * 16); cudaMalloc(&da, N * 8); cudaMalloc(&db, N * 8); cudaMalloc(&o1, sms * 1024 * 4); cudaMalloc(&o2, sms * 1024 * 4);
    cudaMemcpy(d, h, N * 16, cudaMemcpyHostToDevice); cudaMemcpy(da, ha, N * 8, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, N * 8, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(scalar_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, N * 16);
    cudaFuncSetAttribute(packed_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, N * 16);
    const double t1 = time_ms([&] { scalar_scan<<<sms, 1024, N * 16>>>(d, reps, o1); });
    const double t2 = time_ms([&] { packed_scan<<<sms, 1024, N * 16>>>(da, db, reps, o2); });
    unsigned *r1 = (unsigned *)malloc(sms * 1024 * 4), *r2 = (unsigned *)malloc(sms * 1024 * 4);
    cudaMemcpy(r1, o1, sms * 1024 * 4, cudaMemcpyDeviceToHost); cudaMemcpy(r2, o2, sms * 1024 * 4, cudaMemcpyDeviceToHost);
    long long diff = 0, tot = 0;
    for (int i = 0; i < sms * 1024; i++) { diff += r1[i] != r2[i]; tot += r1[i]; }
    const double tests = (double)sms * 1024 * reps * N;
    printf("scalar %.3f ms  %.3e tests/s   packed %.3f ms  %.3e tests/s   speed-up %.3f   survivors %lld  threads that differ %lld\n", t1, tests / (t1 * 1e-3),
           t2, tests / (t2 * 1e-3), t1 / t2, tot, diff);
    return 0;
}
