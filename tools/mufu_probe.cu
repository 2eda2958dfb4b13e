// MUFU.TANH throughput by operand type on sm_100a: fp32 vs f16 / bf16 (the packed tanh.approx.f16x2 / bf16x2 compile to two
// MUFU.TANH.F16 / .BF16 each).  nvcc -gencode arch=compute_100a,code=sm_100a -o mufu_probe mufu_probe.cu && ./mufu_probe
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(float* out, int iters) {
    float a[8]; unsigned h[8];
    for (int i = 0; i < 8; ++i) { a[i] = 0.001f * (threadIdx.x + i); h[i] = 0x3c003800u + threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 1) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
            if (MODE == 2) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h[i]));
            if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int per_inst) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 4096;
    k<MODE><<<148 * 8, 256>>>(out, 16); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 148.0 * 8 * 256 * 8.0 * iters * per_inst;
    printf("%-22s %.3f ms  %.1f results/clk/SM (at 1.965 GHz)\n", name, ms, ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}
int main() { run<0>("tanh.approx.f32", 1); run<1>("tanh.approx.f16x2", 2); run<2>("tanh.approx.bf16x2", 2); run<3>("ex2.approx.f32", 1); return 0; }
