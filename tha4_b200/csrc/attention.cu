// Self-attention of the U-Net bottleneck (unet.py:192-202,230-239): 256 tokens (16x16), 8 heads x 32 channels.
// One CTA per (sample, head); K and V of the head live in shared memory (64 KB), each thread owns one query row
// and runs an online softmax over the 256 keys.  0.4 GFLOP per network: latency-, not throughput-critical.
#include "ops.cuh"
#include "profiler.cuh"

namespace tha4 {
namespace {

constexpr int L = 256, D = 32;

__global__ void __launch_bounds__(L) attention_kernel(const float* __restrict__ qkv, int qkv_ld, int C, int heads,
                                                      float* __restrict__ out, int out_ld) {
    extern __shared__ __align__(16) float sm[];
    float* Ks = sm;            // [L][D]
    float* Vs = sm + L * D;    // [L][D]
    const int n = blockIdx.x / heads, h = blockIdx.x % heads;
    const int t = threadIdx.x;
    const float* base = qkv + (long)n * L * qkv_ld;
    // stage K and V: thread t copies token t
    {
        const float4* kp = reinterpret_cast<const float4*>(base + (long)t * qkv_ld + C + h * D);
        const float4* vp = reinterpret_cast<const float4*>(base + (long)t * qkv_ld + 2 * C + h * D);
#pragma unroll
        for (int j = 0; j < D / 4; ++j) {
            reinterpret_cast<float4*>(Ks + t * D)[j] = kp[j];
            reinterpret_cast<float4*>(Vs + t * D)[j] = vp[j];
        }
    }
    float q[D];
    {
        const float4* qp = reinterpret_cast<const float4*>(base + (long)t * qkv_ld + h * D);
#pragma unroll
        for (int j = 0; j < D / 4; ++j) {
            float4 v = qp[j];
            q[4 * j] = v.x; q[4 * j + 1] = v.y; q[4 * j + 2] = v.z; q[4 * j + 3] = v.w;
        }
    }
    // scale = C_head^-1/4 applied to both q and k (unet.py:197-199)
    const float scale = 1.0f / sqrtf(sqrtf((float)D));
#pragma unroll
    for (int j = 0; j < D; ++j) q[j] *= scale;
    __syncthreads();

    float m = -INFINITY, l = 0.0f, acc[D];
#pragma unroll
    for (int j = 0; j < D; ++j) acc[j] = 0.0f;
    for (int s = 0; s < L; ++s) {
        const float* kr = Ks + s * D;
        float dot = 0.0f;
#pragma unroll
        for (int j = 0; j < D; ++j) dot = fmaf(q[j], kr[j] * scale, dot);
        const float mn = fmaxf(m, dot);
        const float corr = expf(m - mn);
        const float pw = expf(dot - mn);
        l = l * corr + pw;
        const float* vr = Vs + s * D;
#pragma unroll
        for (int j = 0; j < D; ++j) acc[j] = fmaf(acc[j], corr, pw * vr[j]);
        m = mn;
    }
    const float inv = 1.0f / l;
    float4* op = reinterpret_cast<float4*>(out + ((long)n * L + t) * out_ld + h * D);
#pragma unroll
    for (int j = 0; j < D / 4; ++j)
        op[j] = make_float4(acc[4 * j] * inv, acc[4 * j + 1] * inv, acc[4 * j + 2] * inv, acc[4 * j + 3] * inv);
}

}  // namespace

void attention_forward(const View& qkv, int heads, const View& out, cudaStream_t s) {
    THA4_REQUIRE(qkv.H * qkv.W == L && out.C * 3 == qkv.C && out.C / heads == D, "attention: shape (L=256, head dim 32)");
    THA4_REQUIRE(qkv.ld % 4 == 0 && out.ld % 4 == 0, "attention: alignment");
    const size_t smem = 2 * L * D * sizeof(float);
    static bool configured = false;
    if (!configured) {
        THA4_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    ProfScope prof(PROF_ATTN, s);
    attention_kernel<<<qkv.N * heads, L, smem, s>>>(qkv.p, qkv.ld, out.C, heads, out.p, out.ld);
    THA4_LAUNCH_CHECK();
}

}  // namespace tha4
