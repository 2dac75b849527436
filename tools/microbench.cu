// microbench.cu -- issue-rate micro-peaks on the device, used to choose the roofline denominator of the
// ray stage (SURVEY.md 8d asks for measured FFMA / shared-memory-broadcast peaks).  Standalone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096

// mode 0: FFMA with immediate multiplier/addend; 1: FFMA with three register operands;
// 2: FMUL reg; 3: FADD reg; 4: FFMA reg + FADD reg interleaved; 5: FFMA reg + FSETP interleaved
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, const float *in) {
    float a = in[0], b = in[1], c2 = in[2], d = in[3];
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = threadIdx.x * 1e-3f + j;
    int cnt = 0;
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (MODE == 0) x[j] = __fmaf_rn(x[j], 0.999f, 1e-4f);
                if (MODE == 1) x[j] = __fmaf_rn(x[j], a, b);
                if (MODE == 2) x[j] = __fmul_rn(x[j], a);
                if (MODE == 3) x[j] = __fadd_rn(x[j], b);
                if (MODE == 4) { x[j] = (j & 1) ? __fmaf_rn(x[j], a, b) : __fadd_rn(x[j], c2); }
                if (MODE == 5) { x[j] = __fmaf_rn(x[j], a, b); cnt += (x[j] > d) ? 1 : 0; }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += x[j];
    if (s == 12345.678f || cnt == -1) out[0] = s;
}

// Packed FP32 (sm_100: fma/add/mul .f32x2 -> FFMA2 / FADD2 / FMUL2, two IEEE binary32 results per lane and instruction).
// mode 0: FFMA2 three-register; 1: FMUL2; 2: FADD2; 3: FFMA2 + scalar FSETP interleaved 1:1; 4: FFMA2 + FFMA 1:1
__device__ __forceinline__ unsigned long long pk(float lo, float hi) {
    return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
template <int MODE>
__global__ void __launch_bounds__(256) k2(float *out, const float *in) {
    const unsigned long long a = pk(in[0], in[0]), b = pk(in[1], in[1]);
    const float d = in[3], as = in[0], bs = in[1];
    unsigned long long x[8];
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { x[j] = pk(threadIdx.x * 1e-3f + j, threadIdx.x * 2e-3f + j); y[j] = threadIdx.x * 3e-3f + j; }
    int cnt = 0;
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (MODE == 0 || MODE == 3 || MODE == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[j]) : "l"(a), "l"(b));
                if (MODE == 1) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x[j]) : "l"(a));
                if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x[j]) : "l"(b));
                if (MODE == 3) cnt += (__uint_as_float((unsigned)x[j]) > d) ? 1 : 0;
                if (MODE == 4) y[j] = __fmaf_rn(y[j], as, bs);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += __uint_as_float((unsigned)x[j]) + __uint_as_float((unsigned)(x[j] >> 32)) + y[j];
    if (s == 12345.678f || cnt == -1) out[0] = s;
}

// shared-memory broadcast LDS.128 rate: every lane reads the same 16 bytes
__global__ void __launch_bounds__(256) lds_bcast(float *out, int n) {
    __shared__ float4 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float s = 0;
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            float4 v = buf[(i * 16 + u + n) & 1023];
            s += v.x + v.w;
        }
    }
    if (s == 12345.678f) out[0] = s;
}

// L2 -> shared memory bulk copy (TMA 1-D, cp.async.bulk + mbarrier), the way trace_kernel.cu stages the wall
// planes: every CTA copies the same `bytes`-long, L2-resident region `reps` times, two copies in flight.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256) tma_stage(const float4 *src, unsigned bytes, int reps, float *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bar[2];
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; b++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float s = 0;
    for (int r = 0; r < reps + 1; r++) {
        if (r < reps && threadIdx.x == 0) {
            const int b = r & 1;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(smem + (size_t)b * bytes)),
                         "l"(src), "r"(bytes), "r"(smem_u32(&bar[b]))
                         : "memory");
        }
        if (r > 0) {  // wait for copy r-1 while copy r is in flight
            const int b = (r - 1) & 1;
            const unsigned parity = ((r - 1) >> 1) & 1;
            unsigned ok = 0;
            while (!ok) {
                asm volatile(
                    "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                    : "=r"(ok)
                    : "r"(smem_u32(&bar[b])), "r"(parity)
                    : "memory");
            }
            s += reinterpret_cast<const float *>(smem + (size_t)b * bytes)[threadIdx.x];
            __syncthreads();  // the buffer is reused two copies later
        }
    }
    if (s == 12345.678f) out[0] = s;
}

template <class F>
double time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount, blocks = sms * 8;
    float *out, *in;
    cudaMalloc(&out, 64);
    cudaMalloc(&in, 64);
    float h[4] = {0.999f, 1e-4f, 1e-5f, 1e30f};
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    const double ops = (double)blocks * 256 * ITER * 32.0;
    const char *names[] = {"FFMA imm-form", "FFMA 3-register", "FMUL register", "FADD register", "FFMA+FADD 1:1", "FFMA+FSETP+IADD"};
    double ms[6];
    ms[0] = time_ms([&] { k<0><<<blocks, 256>>>(out, in); });
    ms[1] = time_ms([&] { k<1><<<blocks, 256>>>(out, in); });
    ms[2] = time_ms([&] { k<2><<<blocks, 256>>>(out, in); });
    ms[3] = time_ms([&] { k<3><<<blocks, 256>>>(out, in); });
    ms[4] = time_ms([&] { k<4><<<blocks, 256>>>(out, in); });
    ms[5] = time_ms([&] { k<5><<<blocks, 256>>>(out, in); });
    printf("device %s, %d SMs\n", p.name, sms);
    for (int m = 0; m < 6; m++)
        printf("%-18s %8.3f ms  %7.2f T lane-ops/s  (%.3f warp-inst/clk/SMSP at 1.965 GHz)\n", names[m], ms[m],
               ops / (ms[m] * 1e-3) / 1e12, ops / 32.0 / (ms[m] * 1e-3) / (sms * 4 * 1.965e9));
    {
        const char *n2[] = {"FFMA2 3-register", "FMUL2 register", "FADD2 register", "FFMA2+FSETP+IADD", "FFMA2+FFMA 1:1"};
        double t[5];
        t[0] = time_ms([&] { k2<0><<<blocks, 256>>>(out, in); });
        t[1] = time_ms([&] { k2<1><<<blocks, 256>>>(out, in); });
        t[2] = time_ms([&] { k2<2><<<blocks, 256>>>(out, in); });
        t[3] = time_ms([&] { k2<3><<<blocks, 256>>>(out, in); });
        t[4] = time_ms([&] { k2<4><<<blocks, 256>>>(out, in); });
        for (int m = 0; m < 5; m++)  // `ops` counts packed INSTRUCTIONS per lane here: each is two FP32 results
            printf("%-18s %8.3f ms  %7.2f T packed lane-instr/s = %7.2f T FP32 results/s  (%.3f packed warp-inst/clk/SMSP)\n", n2[m], t[m],
                   ops / (t[m] * 1e-3) / 1e12, 2.0 * ops / (t[m] * 1e-3) / 1e12, ops / 32.0 / (t[m] * 1e-3) / (sms * 4 * 1.965e9));
    }
    double l = time_ms([&] { lds_bcast<<<blocks, 256>>>(out, 3); });
    const double lds = (double)blocks * 256 * ITER * 16.0;
    printf("LDS.128 broadcast  %8.3f ms  %7.2f T lane-loads/s  (%.3f warp-LDS/clk/SM)\n", l, lds / (l * 1e-3) / 1e12,
           lds / 32.0 / (l * 1e-3) / (sms * 1.965e9));
    // L2 -> smem staging: 64 KB per copy (two buffers = 128 KB of shared memory per CTA, one CTA per SM), 200 copies
    {
        const unsigned bytes = 64 << 10;
        const int reps = 200;
        float4 *src;
        cudaMalloc(&src, bytes);
        cudaMemset(src, 0, bytes);
        cudaFuncSetAttribute(tma_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * bytes);
        double t = time_ms([&] { tma_stage<<<sms, 256, 2 * bytes>>>(src, bytes, reps, out); });
        const double total = (double)sms * reps * bytes;
        printf("TMA L2->smem bulk  %8.3f ms  %7.2f TB/s aggregate  (%.1f B/clk/SM at 1.965 GHz)\n", t, total / (t * 1e-3) / 1e12,
               total / (t * 1e-3) / (sms * 1.965e9));
        cudaFree(src);
    }
    return 0;
}
