// Small dense layers: pose / time embedding MLPs and the per-ResBlock FiLM projections (unet.py:137-146,443-452).
// One warp per output element; these are GEMVs with N <= a few hundred rows, far below tensor-core granularity.
#include "ops.cuh"

namespace tha4 {
namespace {

__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ x, int x_ld, int N, int I,
                                                     const float* __restrict__ W, const float* __restrict__ bias, int O,
                                                     int silu_in, float* __restrict__ y, int y_ld) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= N * O) return;
    const int n = warp / O, o = warp - n * O;
    const float* xr = x + (long)n * x_ld;
    const float* wr = W + (long)o * I;
    float acc = 0.0f;
    for (int i = lane; i < I; i += 32) {
        float v = xr[i];
        if (silu_in) v = v / (1.0f + expf(-v));
        acc = fmaf(v, wr[i], acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) y[(long)n * y_ld + o] = acc + (bias ? bias[o] : 0.0f);
}

}  // namespace

void linear_forward(const float* x, int x_ld, int N, int I, const float* W, const float* bias, int O, int silu_in,
                    float* y, int y_ld, cudaStream_t s) {
    const long warps = (long)N * O;
    const int blocks = ceil_div(warps * 32, 256);
    linear_kernel<<<blocks, 256, 0, s>>>(x, x_ld, N, I, W, bias, O, silu_in, y, y_ld);
    THA4_LAUNCH_CHECK();
}

}  // namespace tha4
