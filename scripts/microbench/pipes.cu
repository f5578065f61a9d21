// pipes.cu -- B200 issue-rate microbenchmarks that shape the neighbour-loop kernels:
// warp-instructions per clock per SM for scalar vs packed (f32x2) FP32, compare/select,
// MUFU.RSQ, shared-memory and L1 128-bit loads (uniform vs per-lane addresses).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
typedef unsigned long long u64;

#define KERNEL_PROLOGUE                                       \
    long long t0 = clock64();
#define KERNEL_EPILOGUE(val)                                  \
    long long t1 = clock64();                                 \
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;          \
    if ((val) == 123456.789f) out[0] = (val);

__global__ void k_ffma(float *out, long long *cyc, float b, float c) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    KERNEL_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    KERNEL_EPILOGUE(s)
}
__global__ void k_fadd(float *out, long long *cyc, float b, float c) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    KERNEL_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __fadd_rn(a[i], b);
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    KERNEL_EPILOGUE(s)
}
__global__ void k_fmul(float *out, long long *cyc, float b, float c) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i + 1;
    KERNEL_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __fmul_rn(a[i], b);
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    KERNEL_EPILOGUE(s)
}
__global__ void k_ffma2(float *out, long long *cyc, float b, float c) {
    u64 a[8], bb, cc;
    float2 t = make_float2(b, b), u = make_float2(c, c);
    bb = *(u64 *)&t; cc = *(u64 *)&u;
    for (int i = 0; i < 8; ++i) { float2 v = make_float2(threadIdx.x * 0.001f + i, i); a[i] = *(u64 *)&v; }
    KERNEL_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(bb), "l"(cc));
    }
    float s = 0; for (int i = 0; i < 8; ++i) { float2 v = *(float2 *)&a[i]; s += v.x + v.y; }
    KERNEL_EPILOGUE(s)
}
__global__ void k_fadd2(float *out, long long *cyc, float b, float c) {
    u64 a[8], bb;
    float2 t = make_float2(b, b);
    bb = *(u64 *)&t;
    for (int i = 0; i < 8; ++i) { float2 v = make_float2(threadIdx.x * 0.001f + i, i); a[i] = *(u64 *)&v; }
    KERNEL_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(bb));
    }
    float s = 0; for (int i = 0; i < 8; ++i) { float2 v = *(float2 *)&a[i]; s += v.x + v.y; }
    KERNEL_EPILOGUE(s)
}
__global__ void k_setp_sel(float *out, long long *cyc, float b, float c) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    KERNEL_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("{ .reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %2, %0, p; }" : "+f"(a[i]) : "f"(b), "f"(c));
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    KERNEL_EPILOGUE(s)
}
__global__ void k_rsq(float *out, long long *cyc, float b, float c) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i + 1;
    KERNEL_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("rsqrt.approx.f32 %0, %0;" : "+f"(a[i]));
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    KERNEL_EPILOGUE(s)
}
// ffma interleaved with integer adds: do the two pipes co-issue?
__global__ void k_ffma_iadd(float *out, long long *cyc, float b, float c) {
    float a[4]; int n[4];
    for (int i = 0; i < 4; ++i) { a[i] = threadIdx.x * 0.001f + i; n[i] = threadIdx.x + i; }
    KERNEL_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = fmaf(a[i], b, c);
            asm volatile("add.s32 %0, %0, %1;" : "+r"(n[i]) : "r"(it));
        }
    }
    float s = 0; for (int i = 0; i < 4; ++i) s += a[i] + n[i];
    KERNEL_EPILOGUE(s)
}
template <int MODE>  // 0: uniform address, 1: lane-consecutive 16 B, 2: lane stride 48 B (run starts differ)
__global__ void k_lds128(float *out, long long *cyc, float b, float c) {
    __shared__ float4 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int base = MODE == 0 ? 0 : (MODE == 1 ? lane : lane * 3);
    int acc = 0;
    KERNEL_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 v = buf[(base + it + i * 7) & 1023];
            acc ^= __float_as_int(v.x) ^ __float_as_int(v.y) ^ __float_as_int(v.z) ^ __float_as_int(v.w);
        }
    }
    float s = __int_as_float(acc);
    KERNEL_EPILOGUE(s)
}
template <int MODE>
__global__ void k_ldg128(float *out, long long *cyc, const float4 *__restrict__ g, float c) {
    const int lane = threadIdx.x & 31;
    int base = MODE == 0 ? 0 : (MODE == 1 ? lane : lane * 3);
    int acc = 0;
    KERNEL_PROLOGUE
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 v = __ldg(g + ((base + it + i * 7) & 1023));
            acc ^= __float_as_int(v.x) ^ __float_as_int(v.y) ^ __float_as_int(v.z) ^ __float_as_int(v.w);
        }
    }
    float s = __int_as_float(acc);
    KERNEL_EPILOGUE(s)
}

template <typename F>
void run(const char *name, F launch, int instr_per_iter, int threads, int blocks_per_sm, int sms, long long *dcyc) {
    const int blocks = sms * blocks_per_sm;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(blocks, threads);  // warm
    cudaEventRecord(a);
    launch(blocks, threads);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long *h = new long long[blocks];
    cudaMemcpy(h, dcyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    delete[] h;
    const double warp_instr_per_sm = (double)blocks_per_sm * (threads / 32) * ITERS * instr_per_iter;
    printf("%-22s %7.3f ms  %9.0f cyc/block  %6.3f warp-instr/clk/SM  (%s)\n", name, ms, avg,
           warp_instr_per_sm / avg, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    float *out; long long *cyc; float4 *g;
    cudaMalloc(&out, 64); cudaMalloc(&cyc, sizeof(long long) * sms * 16); cudaMalloc(&g, sizeof(float4) * 1024);
    cudaMemset(g, 0, sizeof(float4) * 1024);
    const int T = 256, B = 8;  // 2048 threads / SM
#define RUN(name, kern, ipi, ...) run(name, [&](int bl, int th) { kern<<<bl, th>>>(out, cyc, __VA_ARGS__); }, ipi, T, B, sms, cyc)
    RUN("ffma (3-reg)", k_ffma, 8, 1.0001f, 0.5f);
    RUN("fadd", k_fadd, 8, 1.0001f, 0.5f);
    RUN("fmul", k_fmul, 8, 1.0001f, 0.5f);
    RUN("ffma2 (f32x2)", k_ffma2, 8, 1.0001f, 0.5f);
    RUN("fadd2 (f32x2)", k_fadd2, 8, 1.0001f, 0.5f);
    RUN("fsetp+fsel (2/it)", k_setp_sel, 16, 1.0001f, 0.5f);
    RUN("mufu.rsq", k_rsq, 8, 1.0001f, 0.5f);
    RUN("ffma+iadd (2/it)", k_ffma_iadd, 8, 1.0001f, 0.5f);
    RUN("lds.128 uniform", k_lds128<0>, 8, 1.0f, 0.5f);
    RUN("lds.128 lane*16B", k_lds128<1>, 8, 1.0f, 0.5f);
    RUN("lds.128 lane*48B", k_lds128<2>, 8, 1.0f, 0.5f);
    RUN("ldg.128 uniform", k_ldg128<0>, 8, g, 0.5f);
    RUN("ldg.128 lane*16B", k_ldg128<1>, 8, g, 0.5f);
    RUN("ldg.128 lane*48B", k_ldg128<2>, 8, g, 0.5f);
    return 0;
}
