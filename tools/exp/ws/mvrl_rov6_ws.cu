// Translation unit of the warp-specialised 6DoF step kernel (rov6_ws_kernel.cuh).  Separate from mvrl_api.cu because
// here the rare-path helpers must be inlined: ptxas cannot allocate registers under setmaxnreg across a call.
#define MVRL_NOINLINE __forceinline__
#include <cuda_runtime.h>
#include "rov6_ws_kernel.cuh"

using namespace mvrl;

// shared-memory opt-in of the kernel on `device`; returns the number of SMs (= persistent CTAs)
int mvrl_rov6_ws_prepare(int device) {
    static int sms[64] = {};
    const int d = (device >= 0 && device < 64) ? device : 63;
    if (sms[d] == 0) {
        int n = 0;
        cudaFuncSetAttribute(rov6_step_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WsShared));
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
        cudaGetLastError();
        sms[d] = n > 0 ? n : 1;
    }
    return sms[d];
}

void mvrl_rov6_ws_launch(const Rov6StepArgs<float>& a, int device, cudaStream_t s) {
    const long cap = mvrl_rov6_ws_prepare(device);
    const long want = ((a.n + 63) / 64 + WS_COMPUTE_WARPS - 1) / WS_COMPUTE_WARPS;
    rov6_step_ws_kernel<<<(unsigned)(want < cap ? want : cap), WS_THREADS, sizeof(WsShared), s>>>(a);
}
