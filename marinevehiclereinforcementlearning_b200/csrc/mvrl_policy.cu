// Fused actor for rollout collection (BASELINE.json config 5; SURVEY.md 8(f) N2 / 8(d) cfg 5): the policy network of
// tag_00_Dec2023_simpleControlTurbulence/main_00_sbl.py:100-105 (MlpPolicy, net_arch [128, 128, 128], GELU) plus the
// Gaussian action head, as ONE kernel that reads the env's structure-of-arrays observation buffer in place and writes the
// structure-of-arrays action buffer the step kernel reads: obs [O][ld] -> 128 -> 128 -> 128 -> A, tanh mean, a = clip(mean +
// std * eps, -1, 1), log-prob.  A rollout step is then two launches (this kernel + the fused env step).
//
// This is the one dense contraction of the path, so it runs on the tensor cores: mma.sync m16n8k16 (bf16 operands, fp32
// accumulate).  Each warp owns 16 environments; the 16 x 128 activations never leave its registers - the accumulator
// fragment of one layer IS the A-operand fragment of the next after bias + GELU + bf16 packing (the m16n8 C layout of two
// adjacent n-tiles coincides with the m16k16 A layout).  The weights (70 KB as bf16, pre-packed on the host in B-fragment
// order so that a warp reads 256 contiguous bytes per mma) sit in shared memory and are loaded once per CTA; the grid is
// persistent (2 CTAs per SM).  K = 128 per layer is far too short for a tcgen05 / TMEM pipeline to pay for its set-up
// (one 128 x 128 x 128 tile per layer and 128 environments): the kernel is bound by the GELU (MUFU.TANH) and the
// shared-memory operand reads, not by MMA issue.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "mvrl_host.h"
#include "mvrl_math.cuh"

using namespace mvrl;

namespace {

constexpr int H = 128;            // hidden width
constexpr int KIN = 16;           // observation dim padded to one k-step
constexpr int NOUT = 8;           // action dim padded to one n-tile
constexpr int TILE = 128;         // environments per CTA tile (8 warps x 16)
constexpr int THREADS = 256;

// byte offsets inside the packed parameter block (device + shared memory)
constexpr int OFF_W1 = 0;                          // [1 kk][16 nt][32 lanes] uint2
constexpr int OFF_W2 = OFF_W1 + 1 * 16 * 256;      // [8][16][32] uint2
constexpr int OFF_W3 = OFF_W2 + 8 * 16 * 256;
constexpr int OFF_W4 = OFF_W3 + 8 * 16 * 256;      // [8][1][32] uint2
constexpr int OFF_B = OFF_W4 + 8 * 1 * 256;        // float b1[128] b2[128] b3[128] b4[8] std[8]
constexpr int PACKED_BYTES = OFF_B + (3 * H + 2 * NOUT) * 4;
static_assert(PACKED_BYTES % 16 == 0, "packed block is copied with 16-byte accesses");

struct PolicyArgs {
    const unsigned char* packed;
    long n, ld;
    const float* obs; float* act; float* logp; float* mean; float* eps;
    int obs_dim, act_dim, deterministic;
    float logp_const;
    unsigned long long seed, env_id0;
    unsigned step;
};

__device__ __forceinline__ void mma_bf16(float (&c)[4], const unsigned (&a)[4], uint2 b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ unsigned pack_bf16(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);   // .x (low half) = lo
    return *reinterpret_cast<const unsigned*>(&v);
}
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// GELU in its tanh form, 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))) = torch gelu(approximate="tanh") - within 1e-3 of
// the erf form that torch.nn.GELU() (the reference's activation_fn, legacy/main_00_sbl.py:100-105) evaluates, i.e. below the
// bf16 operand rounding of this kernel
// bias + GELU of two adjacent columns at once on the packed fp32 instructions of sm_100 (FADD2 / FMUL2 / FFMA2: the
// activation is ~55 % of this kernel's instructions when written per element), then one bf16x2 pack
__device__ __forceinline__ unsigned gelu_pack2(float x0, float x1, float2 b) {
    const float2 x = __fadd2_rn(make_float2(x0, x1), b);
    const float2 x2 = __fmul2_rn(x, x);
    const float2 u = __fmul2_rn(x, __ffma2_rn(x2, make_float2(0.0356774081f, 0.0356774081f), make_float2(0.7978845608f, 0.7978845608f)));
    const float2 th = make_float2(tanh_fast(u.x), tanh_fast(u.y));
    const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
    const float2 y = __ffma2_rn(h, th, h);
    return pack_bf16(y.x, y.y);
}

// accumulators of one layer (+ bias, GELU) -> A fragments of the next: n-tiles 2kk and 2kk + 1 give k-step kk
__device__ __forceinline__ void activate_to_frags(const float (&acc)[16][4], const float* bias, int t, unsigned (&afr)[8][4]) {
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        const float2 b0 = *reinterpret_cast<const float2*>(bias + 16 * kk + 2 * t);
        const float2 b1 = *reinterpret_cast<const float2*>(bias + 16 * kk + 8 + 2 * t);
        const float (&c0)[4] = acc[2 * kk];
        const float (&c1)[4] = acc[2 * kk + 1];
        afr[kk][0] = gelu_pack2(c0[0], c0[1], b0);   // row g,     k = 16 kk + 2t, +1
        afr[kk][1] = gelu_pack2(c0[2], c0[3], b0);   // row g + 8
        afr[kk][2] = gelu_pack2(c1[0], c1[1], b1);   // row g,     k = 16 kk + 8 + 2t, +1
        afr[kk][3] = gelu_pack2(c1[2], c1[3], b1);   // row g + 8
    }
}

__device__ __forceinline__ void hidden_layer(const uint2* w, const unsigned (&afr)[8][4], int lane, float (&acc)[16][4]) {
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f; }
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) mma_bf16(acc[nt], afr[kk], w[(kk * 16 + nt) * 32 + lane]);
    }
}

// two standard normals from one Philox block of (seed, environment, step): Box-Muller on 24-bit uniforms
__device__ __forceinline__ float2 normal_pair(unsigned long long seed, unsigned long long env, unsigned step, unsigned blk) {
    const uint4 w = Philox::draw(seed, env, step, 7u, blk);
    const float u1 = (float(w.x >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0, 1)
    const float u2 = float(w.y >> 8) * (1.0f / 16777216.0f);            // [0, 1)
    const float r = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);
    return make_float2(r * c, r * s);
}

__global__ void __launch_bounds__(THREADS, 2) policy_act_kernel(const __grid_constant__ PolicyArgs a) {
    extern __shared__ uint4 smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(smem_raw);
    {   // parameters: global (L2) -> shared, once per CTA
        const uint4* src = reinterpret_cast<const uint4*>(a.packed);
        for (int i = threadIdx.x; i < PACKED_BYTES / 16; i += THREADS) smem_raw[i] = src[i];
    }
    __syncthreads();
    const uint2* w1 = reinterpret_cast<const uint2*>(smem + OFF_W1);
    const uint2* w2 = reinterpret_cast<const uint2*>(smem + OFF_W2);
    const uint2* w3 = reinterpret_cast<const uint2*>(smem + OFF_W3);
    const uint2* w4 = reinterpret_cast<const uint2*>(smem + OFF_W4);
    const float* bias = reinterpret_cast<const float*>(smem + OFF_B);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const long n_tiles = (a.n + TILE - 1) / TILE;
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long r0 = tile * TILE + warp * 16 + g, r1 = r0 + 8;
        const bool ok0 = r0 < a.n, ok1 = r1 < a.n;
        // ---- layer 1: the observation rows of this thread's two environments, k = 2t, 2t + 1, 2t + 8, 2t + 9
        unsigned afr[8][4];
        {
            float x[2][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = 2 * t + (j & 1) + (j >> 1) * 8;
                const bool kv = k < a.obs_dim;
                x[0][j] = (kv && ok0) ? a.obs[(long)k * a.ld + r0] : 0.0f;
                x[1][j] = (kv && ok1) ? a.obs[(long)k * a.ld + r1] : 0.0f;
            }
            const unsigned a1[4] = {pack_bf16(x[0][0], x[0][1]), pack_bf16(x[1][0], x[1][1]), pack_bf16(x[0][2], x[0][3]), pack_bf16(x[1][2], x[1][3])};
            float acc[16][4];
#pragma unroll
            for (int nt = 0; nt < 16; ++nt) {
                acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
                mma_bf16(acc[nt], a1, w1[nt * 32 + lane]);
            }
            activate_to_frags(acc, bias, t, afr);
        }
        // ---- layers 2 and 3 (128 x 128)
#pragma unroll 1
        for (int layer = 0; layer < 2; ++layer) {
            float acc[16][4];
            hidden_layer(layer == 0 ? w2 : w3, afr, lane, acc);
            activate_to_frags(acc, bias + (layer + 1) * H, t, afr);
        }
        // ---- head: 128 -> A (one n-tile), tanh mean, Gaussian sample
        float c4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) mma_bf16(c4, afr[kk], w4[kk * 32 + lane]);
        const int col = 2 * t;                       // this thread holds columns col, col + 1 of rows r0 (c4[0..1]) and r1 (c4[2..3])
        const float* b4 = bias + 3 * H;
        const float* sd = b4 + NOUT;
        float lp0 = 0.0f, lp1 = 0.0f;
        if (col < a.act_dim) {
            const float m00 = tanh_fast(c4[0] + b4[col]), m01 = tanh_fast(c4[1] + b4[col + 1]);
            const float m10 = tanh_fast(c4[2] + b4[col]), m11 = tanh_fast(c4[3] + b4[col + 1]);
            float2 e0 = make_float2(0.0f, 0.0f), e1 = e0;
            if (!a.deterministic) {
                e0 = normal_pair(a.seed, a.env_id0 + (unsigned long long)r0, a.step, (unsigned)t);
                e1 = normal_pair(a.seed, a.env_id0 + (unsigned long long)r1, a.step, (unsigned)t);
            }
            const bool second = col + 1 < a.act_dim;
            const float a00 = fminf(1.0f, fmaxf(-1.0f, fmaf(sd[col], e0.x, m00))), a01 = fminf(1.0f, fmaxf(-1.0f, fmaf(sd[col + 1], e0.y, m01)));
            const float a10 = fminf(1.0f, fmaxf(-1.0f, fmaf(sd[col], e1.x, m10))), a11 = fminf(1.0f, fmaxf(-1.0f, fmaf(sd[col + 1], e1.y, m11)));
            lp0 = -0.5f * (e0.x * e0.x + (second ? e0.y * e0.y : 0.0f));
            lp1 = -0.5f * (e1.x * e1.x + (second ? e1.y * e1.y : 0.0f));
            if (ok0) {
                a.act[(long)col * a.ld + r0] = a00;
                if (second) a.act[(long)(col + 1) * a.ld + r0] = a01;
                if (a.mean) { a.mean[(long)col * a.ld + r0] = m00; if (second) a.mean[(long)(col + 1) * a.ld + r0] = m01; }
                if (a.eps) { a.eps[(long)col * a.ld + r0] = e0.x; if (second) a.eps[(long)(col + 1) * a.ld + r0] = e0.y; }
            }
            if (ok1) {
                a.act[(long)col * a.ld + r1] = a10;
                if (second) a.act[(long)(col + 1) * a.ld + r1] = a11;
                if (a.mean) { a.mean[(long)col * a.ld + r1] = m10; if (second) a.mean[(long)(col + 1) * a.ld + r1] = m11; }
                if (a.eps) { a.eps[(long)col * a.ld + r1] = e1.x; if (second) a.eps[(long)(col + 1) * a.ld + r1] = e1.y; }
            }
        }
        // log-prob: sum over the action columns = over the four lanes of a quad
        lp0 += __shfl_xor_sync(0xffffffffu, lp0, 1); lp0 += __shfl_xor_sync(0xffffffffu, lp0, 2);
        lp1 += __shfl_xor_sync(0xffffffffu, lp1, 1); lp1 += __shfl_xor_sync(0xffffffffu, lp1, 2);
        if (a.logp != nullptr && t == 0) {
            if (ok0) a.logp[r0] = lp0 + a.logp_const;
            if (ok1) a.logp[r1] = lp1 + a.logp_const;
        }
    }
}

unsigned short bf16_bits(float f) {   // round to nearest even, like __float2bfloat16_rn
    unsigned u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (unsigned short)((u >> 16) | 0x40);   // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return (unsigned short)(u >> 16);
}

// W [N][K] (row = output unit, as torch.nn.Linear stores it) -> B fragments of mma.m16n8k16: for k-step kk, n-tile nt, lane
// (g = lane / 4, t = lane % 4): b0 = (W[8 nt + g][16 kk + 2t], +1), b1 = (W[8 nt + g][16 kk + 8 + 2t], +1); zeros outside N x K
void pack_b_fragments(const float* W, int N, int K, int n_tiles, int k_steps, unsigned char* dst) {
    unsigned* out = reinterpret_cast<unsigned*>(dst);
    auto at = [&](int n, int k) -> unsigned { return (n < N && k < K) ? bf16_bits(W[(size_t)n * K + k]) : 0u; };
    for (int kk = 0; kk < k_steps; ++kk)
        for (int nt = 0; nt < n_tiles; ++nt)
            for (int lane = 0; lane < 32; ++lane) {
                const int g = lane >> 2, t = lane & 3, n = 8 * nt + g, k0 = 16 * kk + 2 * t;
                const size_t i = ((size_t)(kk * n_tiles + nt) * 32 + lane) * 2;
                out[i] = at(n, k0) | (at(n, k0 + 1) << 16);
                out[i + 1] = at(n, k0 + 8) | (at(n, k0 + 9) << 16);
            }
}

}  // namespace

struct MvrlPolicy {
    int device, obs_dim, act_dim, sm_count;
    unsigned char* packed;   // device
    bool has_weights;
    float logp_const;        // -sum(log_std)
};

extern "C" MVRL_API int mvrl_policy_create(MvrlPolicy** out, int device, int obs_dim, int act_dim) {
    if (!out) return mvrl_fail(MVRL_EINVAL, "mvrl_policy_create: null output");
    if (obs_dim < 1 || obs_dim > KIN || act_dim < 1 || act_dim > NOUT)
        return mvrl_fail(MVRL_EINVAL, "mvrl_policy_create: obs_dim must be 1..%d and act_dim 1..%d (got %d, %d)", KIN, NOUT, obs_dim, act_dim);
    { const int rc = mvrl_require_device(device); if (rc != MVRL_OK) return rc; }
    MVRL_ON_DEVICE(device);
    MvrlPolicy* h = new (std::nothrow) MvrlPolicy();
    if (!h) return mvrl_fail(MVRL_EINVAL, "out of host memory");
    h->device = device; h->obs_dim = obs_dim; h->act_dim = act_dim; h->has_weights = false; h->logp_const = 0.f; h->packed = nullptr;
    h->sm_count = 148;
    { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) h->sm_count = v; else cudaGetLastError(); }
    cudaError_t e = cudaMalloc(&h->packed, PACKED_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(policy_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PACKED_BYTES);
    if (e != cudaSuccess) {
        if (h->packed) cudaFree(h->packed);
        delete h;
        return mvrl_fail(MVRL_ECUDA, "mvrl_policy_create: %s", cudaGetErrorString(e));
    }
    *out = h;
    return MVRL_OK;
}

extern "C" MVRL_API int mvrl_policy_destroy(MvrlPolicy* h) {
    if (!h) return MVRL_OK;
    { MvrlDeviceGuard guard(h->device); cudaFree(h->packed); }
    delete h;
    return MVRL_OK;
}

extern "C" MVRL_API int mvrl_policy_set_weights(MvrlPolicy* h, const float* W1, const float* b1, const float* W2, const float* b2,
                                                const float* W3, const float* b3, const float* W4, const float* b4, const float* log_std) {
    if (!h || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !W4 || !b4 || !log_std) return mvrl_fail(MVRL_EINVAL, "mvrl_policy_set_weights: null argument");
    std::vector<unsigned char> buf(PACKED_BYTES, 0);
    pack_b_fragments(W1, H, h->obs_dim, 16, 1, buf.data() + OFF_W1);
    pack_b_fragments(W2, H, H, 16, 8, buf.data() + OFF_W2);
    pack_b_fragments(W3, H, H, 16, 8, buf.data() + OFF_W3);
    pack_b_fragments(W4, h->act_dim, H, 1, 8, buf.data() + OFF_W4);
    float* fb = reinterpret_cast<float*>(buf.data() + OFF_B);
    memcpy(fb, b1, H * 4); memcpy(fb + H, b2, H * 4); memcpy(fb + 2 * H, b3, H * 4);
    double lsum = 0;
    for (int k = 0; k < h->act_dim; ++k) { fb[3 * H + k] = b4[k]; fb[3 * H + NOUT + k] = expf(log_std[k]); lsum += log_std[k]; }
    h->logp_const = (float)(-lsum);
    MVRL_ON_DEVICE(h->device);
    MVRL_CUDA(cudaMemcpy(h->packed, buf.data(), PACKED_BYTES, cudaMemcpyHostToDevice));
    h->has_weights = true;
    return MVRL_OK;
}

extern "C" MVRL_API int mvrl_policy_act(MvrlPolicy* h, int64_t n, int64_t ld, const float* obs, float* act, float* logp, float* mean,
                                        float* eps, uint64_t seed, uint64_t env_id0, uint32_t step, int deterministic, mvrl_stream_t stream) {
    if (!h || !obs || !act) return mvrl_fail(MVRL_EINVAL, "mvrl_policy_act: null argument");
    if (n < 0 || ld < n) return mvrl_fail(MVRL_EINVAL, "mvrl_policy_act: need 0 <= n <= ld");
    if (!h->has_weights) return mvrl_fail(MVRL_EINVAL, "mvrl_policy_act: call mvrl_policy_set_weights first");
    if (n == 0) return MVRL_OK;
    MVRL_ON_DEVICE(h->device);
    PolicyArgs a;
    a.packed = h->packed; a.n = n; a.ld = ld; a.obs = obs; a.act = act; a.logp = logp; a.mean = mean; a.eps = eps;
    a.obs_dim = h->obs_dim; a.act_dim = h->act_dim; a.deterministic = deterministic ? 1 : 0; a.logp_const = h->logp_const;
    a.seed = seed; a.env_id0 = env_id0; a.step = step;
    const int64_t tiles = (n + TILE - 1) / TILE;
    const int64_t cap = 2 * (int64_t)h->sm_count;
    const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
    policy_act_kernel<<<grid, THREADS, PACKED_BYTES, (cudaStream_t)stream>>>(a);
    return mvrl_check_launch("policy_act");
}
