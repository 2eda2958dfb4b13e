// Diagnostic: dependent-issue latency / throughput of FFMA vs FFMA2 as a function of independent chains and warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>

template <int CH, bool PACKED>
__global__ void chain_kernel(float* out, int iters, float a, float b) {
    float2 x[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) x[j] = make_float2((threadIdx.x + j) * 1e-3f, (threadIdx.x + j) * 2e-3f);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                if (PACKED) x[j] = __ffma2_rn(x[j], a2, b2);
                else x[j].x = fmaf(x[j].x, a, b);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < CH; ++j) s += x[j].x + x[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH, bool PACKED> void run(float* buf, int sms, int warps_per_sm) {
    const int iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        chain_kernel<CH, PACKED><<<sms, warps_per_sm * 32>>>(buf, iters, 0.999999f, 1e-7f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2 && ms < best) best = ms;
    }
    const double inst_per_warp = (double)iters * 16 * CH;
    const double clk = best * 1e-3 * 1.965e9;
    const double warps_per_smsp = warps_per_sm / 4.0;
    printf("%s CH=%d warps/SMSP=%.0f : %.2f clk per dependent step, %.3f inst/clk/SMSP\n", PACKED ? "FFMA2" : "FFMA ", CH, warps_per_smsp,
           clk / (iters * 16.0), inst_per_warp * warps_per_smsp / clk);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float* buf; cudaMalloc(&buf, (size_t)p.multiProcessorCount * 1024 * 4);
    for (int w : {4, 8, 12, 16}) {
        run<1, false>(buf, p.multiProcessorCount, w); run<2, false>(buf, p.multiProcessorCount, w); run<4, false>(buf, p.multiProcessorCount, w);
        run<1, true>(buf, p.multiProcessorCount, w); run<2, true>(buf, p.multiProcessorCount, w); run<4, true>(buf, p.multiProcessorCount, w);
    }
    return cudaGetLastError() != cudaSuccess;
}
