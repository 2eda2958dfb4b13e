// Diagnostic: FFMA / FFMA2 throughput when the multiplicand and addend are per-thread REGISTERS
// (3 distinct register operands) rather than uniform / constant operands.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int CH = 8;
template <bool PACKED, int NREG>   // NREG = number of distinct register (non-uniform) source operands besides the accumulator: 0, 1, 2
__global__ void __launch_bounds__(256) k(const float* in, float* out, int iters, float ua, float ub) {
    float2 x[CH], y[CH], z[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        x[j] = make_float2(in[threadIdx.x + j], in[threadIdx.x + j + 64]);
        y[j] = make_float2(in[threadIdx.x + j + 128], in[threadIdx.x + j + 192]);
        z[j] = make_float2(in[threadIdx.x + j + 256], in[threadIdx.x + j + 320]);
    }
    const float2 a2 = make_float2(ua, ua), b2 = make_float2(ub, ub);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                if (PACKED) {
                    if (NREG == 0) x[j] = __ffma2_rn(x[j], a2, b2);
                    else if (NREG == 1) x[j] = __ffma2_rn(x[j], y[j], b2);
                    else x[j] = __ffma2_rn(x[j], y[(j + u) % CH], z[(j + 2 * u + 1) % CH]);
                } else {
                    if (NREG == 0) x[j].x = fmaf(x[j].x, ua, ub);
                    else if (NREG == 1) x[j].x = fmaf(x[j].x, y[j].x, ub);
                    else x[j].x = fmaf(x[j].x, y[(j + u) % CH].x, z[(j + 2 * u + 1) % CH].x);
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < CH; ++j) s += x[j].x + x[j].y + y[j].x + z[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <bool P, int N> void run(const float* in, float* out, int blocks) {
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<P, N><<<blocks, 256>>>(in, out, iters, 0.999f, 1e-7f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2 && ms < best) best = ms;
    }
    const double inst = (double)CH * 8 * iters * blocks * 8;   // warp instructions
    printf("%s, %d register operands besides the accumulator: %.3f warp-inst/clk/SMSP, %.1f TFLOP/s\n", P ? "FFMA2" : "FFMA ", N,
           inst / (best * 1e-3 * 1.965e9) / (blocks / 8) / 4, inst * 32 * (P ? 4 : 2) / (best * 1e-3) / 1e12);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8;
    float *in, *out; cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4); cudaMalloc(&out, (size_t)blocks * 256 * 4);
    run<false, 0>(in, out, blocks); run<false, 1>(in, out, blocks); run<false, 2>(in, out, blocks);
    run<true, 0>(in, out, blocks); run<true, 1>(in, out, blocks); run<true, 2>(in, out, blocks);
    return cudaGetLastError() != cudaSuccess;
}
