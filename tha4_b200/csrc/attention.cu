// Self-attention of the U-Net bottleneck (unet.py:192-202,230-239): 256 tokens (16x16), 8 heads x 32 channels.
// One CTA per (sample, head); K and V of the head live in shared memory (64 KB), each thread owns one query row
// and runs an online softmax over the 256 keys.  0.4 GFLOP per network: latency-, not throughput-critical.
#include "ops.cuh"
#include "profiler.cuh"

namespace tha4 {
namespace {

constexpr int L = 256, D = 32;

// grid = N * heads * (L / QPB), block = 256 threads = QPB queries x KSPLIT key splits.  Each thread runs an online softmax
// over its L / KSPLIT keys; the KSPLIT partial (max, sum, acc) triples of a query are merged with warp shuffles.
//   <64, 4, 36>: 4 CTAs per (sample, head) -- the measured default;  row pitch 36: the 4 key-split lanes hit 4 banks.
//   <16, 16, 33>: 16 CTAs per (sample, head) for B=1 latency (128 CTAs instead of 32); row pitch 33: the 16 key-split
//                 lanes read 16 consecutive rows = 16 different banks.  Opt-in (option "attn_split16"), not yet measured.
template <int QPB, int KSPLIT, int DP>
__global__ void __launch_bounds__(QPB * KSPLIT) attention_kernel(const float* __restrict__ qkv, int qkv_ld, int C, int heads,
                                                                 float* __restrict__ out, int out_ld) {
    static_assert(QPB * KSPLIT == L, "one thread per token while staging K / V");
    extern __shared__ __align__(16) float sm[];
    float* Ks = sm;            // [L][DP]
    float* Vs = sm + L * DP;   // [L][DP]
    const int qq = blockIdx.x % (L / QPB);
    const int nh = blockIdx.x / (L / QPB);
    const int n = nh / heads, h = nh % heads;
    const int tid = threadIdx.x;
    const float* base = qkv + (long)n * L * qkv_ld;
    {   // stage K and V of this head: thread tid copies token tid
        const float4* kp = reinterpret_cast<const float4*>(base + (long)tid * qkv_ld + C + h * D);
        const float4* vp = reinterpret_cast<const float4*>(base + (long)tid * qkv_ld + 2 * C + h * D);
#pragma unroll
        for (int j = 0; j < D / 4; ++j) {
            const float4 kv = kp[j], vv = vp[j];
            if (DP % 4 == 0) {
                reinterpret_cast<float4*>(Ks + tid * DP)[j] = kv;
                reinterpret_cast<float4*>(Vs + tid * DP)[j] = vv;
            } else {        // odd pitch: rows are not 16-byte aligned
                float* kd = Ks + tid * DP + 4 * j; float* vd = Vs + tid * DP + 4 * j;
                kd[0] = kv.x; kd[1] = kv.y; kd[2] = kv.z; kd[3] = kv.w;
                vd[0] = vv.x; vd[1] = vv.y; vd[2] = vv.z; vd[3] = vv.w;
            }
        }
    }
    const int t = qq * QPB + tid / KSPLIT;      // query token
    const int ks = tid % KSPLIT;                // key split: keys ks, ks + 4, ks + 8, ...
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
    for (int s = ks; s < L; s += KSPLIT) {
        const float* kr = Ks + s * DP;
        float dot = 0.0f;
#pragma unroll
        for (int j = 0; j < D; ++j) dot = fmaf(q[j], kr[j] * scale, dot);
        const float mn = fmaxf(m, dot);
        const float corr = expf(m - mn);
        const float pw = expf(dot - mn);
        l = l * corr + pw;
        const float* vr = Vs + s * DP;
#pragma unroll
        for (int j = 0; j < D; ++j) acc[j] = fmaf(acc[j], corr, pw * vr[j]);
        m = mn;
    }
    // merge the KSPLIT partials of this query (adjacent lanes)
#pragma unroll
    for (int off = 1; off < KSPLIT; off <<= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, m, off);
        const float lo = __shfl_xor_sync(0xffffffffu, l, off);
        const float mn = fmaxf(m, mo);
        const float ca = expf(m - mn), cb = expf(mo - mn);
        l = l * ca + lo * cb;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const float ao = __shfl_xor_sync(0xffffffffu, acc[j], off);
            acc[j] = acc[j] * ca + ao * cb;
        }
        m = mn;
    }
    if (ks == 0) {
        const float inv = 1.0f / l;
        float4* op = reinterpret_cast<float4*>(out + ((long)n * L + t) * out_ld + h * D);
#pragma unroll
        for (int j = 0; j < D / 4; ++j)
            op[j] = make_float4(acc[4 * j] * inv, acc[4 * j + 1] * inv, acc[4 * j + 2] * inv, acc[4 * j + 3] * inv);
    }
}

bool g_attn_split16 = false;

template <int QPB, int KSPLIT, int DP>
void launch_attention(const View& qkv, int heads, const View& out, cudaStream_t s) {
    const size_t smem = 2 * L * DP * sizeof(float);
    THA4_ENSURE_SMEM((attention_kernel<QPB, KSPLIT, DP>), smem);
    attention_kernel<QPB, KSPLIT, DP><<<qkv.N * heads * (L / QPB), QPB * KSPLIT, smem, s>>>(qkv.p, qkv.ld, out.C, heads, out.p, out.ld);
    THA4_LAUNCH_CHECK();
}

}  // namespace

void attention_enable_split16(bool on) { g_attn_split16 = on; }

void attention_forward(const View& qkv, int heads, const View& out, cudaStream_t s) {
    THA4_REQUIRE(qkv.H * qkv.W == L && out.C * 3 == qkv.C && out.C / heads == D, "attention: shape (L=256, head dim 32)");
    THA4_REQUIRE(qkv.ld % 4 == 0 && out.ld % 4 == 0, "attention: alignment");
    ProfScope prof(PROF_ATTN, s);
    if (g_attn_split16) launch_attention<16, 16, 33>(qkv, heads, out, s);
    else launch_attention<64, 4, 36>(qkv, heads, out, s);
}

}  // namespace tha4
