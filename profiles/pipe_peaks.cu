// Micro-benchmark of the SM pipes that bound the BAMSE scoring kernels on B200:
//   FFMA (fp32 FMA pipe), FFMA2 (packed fma.rn.f32x2), MUFU.EX2 (XU pipe), DFMA (fp64 pipe).
// MEASURED_PEAKS.json only carries HBM GB/s and bf16 tensor TFLOP/s, neither of which
// bounds this path, so the roofline denominators come from here.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_peaks pipe_peaks.cu && ./pipe_peaks
// Prints one JSON object (copied to profiles/pipe_peaks.json).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ILP = 16;
constexpr int ITERS = 4096;

__global__ void k_ffma(float* out, float a, float b) {
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 123.456f) out[0] = s;
}

__global__ void k_ffma2(float* out, float a, float b) {
    unsigned long long acc[ILP];
    unsigned long long av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        float x = threadIdx.x * 1e-9f + i;
        asm("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(x));
    }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(av), "l"(bv));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
        s += lo + hi;
    }
    if (s == 123.456f) out[0] = s;
}

__global__ void k_mufu(float* out, float a) {
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3f + i * 0.01f + a;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(acc[i]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 123.456f) out[0] = s;
}

__global__ void k_dfma(double* out, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

// mixed: 1 MUFU per 8 FFMA, to see whether the pipes overlap
__global__ void k_mix(float* out, float a, float b) {
    float acc[ILP], e[2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9f + i;
    e[0] = a; e[1] = b;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], a, b);
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e[0]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e[1]));
    }
    float s = e[0] + e[1];
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 123.456f) out[0] = s;
}

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const int blocks = sms * 8, threads = 256;
    float* d; cudaMalloc(&d, 64);
    const double n = (double)blocks * threads * ILP * ITERS;
    double t_ffma = time_ms([&] { k_ffma<<<blocks, threads>>>(d, 1.0001f, 0.5f); });
    double t_ffma2 = time_ms([&] { k_ffma2<<<blocks, threads>>>(d, 1.0001f, 0.5f); });
    double t_mufu = time_ms([&] { k_mufu<<<blocks, threads>>>(d, 0.25f); });
    double t_dfma = time_ms([&] { k_dfma<<<blocks, threads>>>((double*)d, 1.0001, 0.5); });
    double t_mix = time_ms([&] { k_mix<<<blocks, threads>>>(d, 1.0001f, 0.5f); });
    cudaError_t e = cudaDeviceSynchronize();
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"cuda_error\": \"%s\",\n", prop.name, sms, cudaGetErrorString(e));
    printf(" \"ffma_tflops\": %.2f, \"ffma2_tflops\": %.2f, \"xu_tops\": %.3f, \"fp64_tflops\": %.2f,\n",
           2 * n / t_ffma / 1e9, 4 * n / t_ffma2 / 1e9, n / t_mufu / 1e9, 2 * n / t_dfma / 1e9);
    printf(" \"ffma_per_clk_per_sm_at_1965MHz\": %.1f, \"ffma2_fma_per_clk_per_sm_at_1965MHz\": %.1f,"
           " \"mufu_per_clk_per_sm_at_1965MHz\": %.2f, \"dfma_per_clk_per_sm_at_1965MHz\": %.1f,\n",
           n / t_ffma / 1e-3 / sms / 1.965e9, 2 * n / t_ffma2 / 1e-3 / sms / 1.965e9, n / t_mufu / 1e-3 / sms / 1.965e9,
           n / t_dfma / 1e-3 / sms / 1.965e9);
    printf(" \"mix_16ffma_2mufu_ms\": %.3f, \"ffma_only_ms\": %.3f, \"mufu_only_ms_scaled_2_of_16\": %.3f}\n", t_mix,
           t_ffma, t_mufu * 2 / 16);
    return 0;
}
