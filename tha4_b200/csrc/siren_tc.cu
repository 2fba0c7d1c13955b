// Distilled student networks on the Blackwell paths: SirenFaceMorpher00 (siren_face_morpher_00.py:28-51) and the three
// levels of SirenMorpher03 (siren_morpher_03.py:107-139) as persistent fused-MLP kernels on TMA + tcgen05 + TMEM.
//
// A CTA walks 128-pixel tiles (128 consecutive pixels of one image row).  The activations of a tile live in shared
// memory as the K-major SWIZZLE_128B A operand ([K / 64 chunks][128 rows][128 B]); every sine layer is
//   TMA     weight tiles W[NB rows x 64 k] (fp16, pre-scaled by omega_0 = 30) stream through a ring that runs ahead of
//           the math across layers and tiles (the weight sequence of a tile is fixed);
//   UMMA    D[128 x N] (fp32, TMEM) = A[128 x K] . W^T: one elected thread, N in slices of NB <= 256 columns;
//   drain   the four compute warps (thread = pixel = TMEM lane) read the accumulator with tcgen05.ld, add the bias
//           (+ the per-sample pose bias and the two xy terms on a level's first layer: the tiled pose / position
//           planes of siren_morpher_03.py:92-105 are never materialised), take sin(.) and write the result back INTO the
//           A operand as fp16 -- the layer's output is the next layer's operand, it never leaves the SM.
// Level hand-off (bilinear x2, :121) goes through fp16 NHWC tensors in HBM (it needs a cross-tile halo); level 2 ends in
// the fused tail: 1x1 head (a 16-column MMA) -> grid_sample -> blend -> five NCHW outputs (thread = pixel: the 32 lanes
// of a warp write 32 consecutive pixels = full 128-byte lines).  The mma.sync kernels of siren.cu remain as the
// reference path of the option "siren_tc" = 0 and for the distillation backward (stored-activation forward).
#include "siren.cuh"
#include "gridsample.cuh"
#include "profiler.cuh"
#include "tc_common.cuh"
#include <cuda.h>
#include <map>
#include <mutex>
#include <tuple>

namespace tha4 {
namespace {

using namespace tc;

constexpr int ST_THREADS = 192;            // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: compute
constexpr int ST_TILE = 128;
constexpr int ST_MAXL = 8;                 // GEMM layers per kernel (face: 7 sine + head)
enum { SM_BODY0 = 0, SM_BODY1 = 1, SM_BODY2 = 2, SM_FACE = 3 };

struct StLayer {
    int kpad, npad, nb;                    // K (padded, multiple of 16), N (padded), N slice per MMA / per weight tile
    int sine;                              // 1: sin epilogue into the A operand; 0: linear head (N = 16 columns, raw)
    int bias_off;                          // offset of this layer's bias in the staged bias table (floats)
    int first;                             // 1: add the per-sample bias + xy terms (first layer of body levels 1 / 2)
};
struct StMaps { CUtensorMap w[ST_MAXL]; };

struct StParams {
    int R, B, nl;
    StLayer L[ST_MAXL];
    const float* bias_table; int bias_floats;       // all layers' biases, concatenated (pre-scaled)
    // elementwise first layer (body level 0, face): act = sin(pb[n][c] + wx[c] * x + wy[c] * y), c < e_npad
    int e_npad; const float* e_pb; int e_pb_ld; const float* e_wxy;
    // first GEMM layer of body levels 1 / 2: per-sample bias + xy terms
    const float* f_pb; int f_pb_ld; const float* f_wxy;
    const float* base;                              // affine_grid coordinates of this resolution
    const __half* prev; int prev_c;                 // previous level's output [B, R/2, R/2, prev_c] (bilinear x2 prologue)
    __half* out; int out_c;                         // this level's output [B, R, R, out_c] (levels 0 / 1)
    ImgView image; float* o[5]; int o_f16;          // level 2: tail (o_f16: the five outputs are __half planes, io_dtype = f16)
    float* face_out;                                // face: [B, 4, R, R]
    const float* head_bias;
};

// sin(x) WITHOUT the transcendental unit.  The mma.sync student kernels (siren.cu) and the first version of this file used
// rintf + MUFU.SIN: two XU-pipe operations per output -- and ncu showed the XU pipe, not the tensor pipe, bounding every
// level (profiles/r02_ncu_siren_tc_v1.txt: both implementations ran at ~1.8 sin / clk / SM).  Here: k = round(x / pi) by
// the magic-number trick (FMA pipe), r = x - k pi (two-constant Cody-Waite), sin(r) on [-pi/2, pi/2] as the degree-9
// Taylor polynomial (|error| <= 3.6e-6, far below the fp16 the result is stored in), sign (-1)^k from the parity bit.
// 13 FMA / ALU-pipe instructions, 128 lanes / clk / SM.
__device__ __forceinline__ float st_sin(float x) {
    const float kf = fmaf(x, 0.31830988618379067f, 12582912.0f);          // 1.5 * 2^23: the integer k sits in the low mantissa bits
    const float k = kf - 12582912.0f;
    float r = fmaf(-k, 3.1415927410125732f, x);
    r = fmaf(-k, -8.7422776573475858e-8f, r);
    const float r2 = r * r;
    float p = fmaf(r2, 2.7557319223985893e-6f, -1.9841269841269841e-4f);
    p = fmaf(p, r2, 8.3333333333333332e-3f);
    p = fmaf(p, r2, -1.6666666666666666e-1f);
    const float s = fmaf(p * r2, r, r);
    return __uint_as_float(__float_as_uint(s) ^ ((__float_as_uint(kf) & 1u) << 31));
}

// byte offset of the 16-byte chunk holding channels [c8, c8 + 8) of tile row `row` in the swizzled A operand
__device__ __forceinline__ uint32_t a_off(int row, int c8) {
    const int chunk = c8 >> 6, j = (c8 & 63) >> 3;
    return (uint32_t)(chunk * (ST_TILE * 128) + row * 128 + ((j ^ (row & 7)) << 4));
}

template <int ACH, int NBMAX, int SB, int TMEM_COLS, int MODE>
__global__ void __launch_bounds__(ST_THREADS) siren_tc_kernel(const __grid_constant__ StMaps maps, const StParams p) {
    constexpr int A_BYTES = ACH * ST_TILE * 128;
    constexpr int B_STAGE = NBMAX * 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic (not an integer round trip) keeps the shared address space: LDS / STS, not generic LD / ST
    uint8_t* smA = smem;
    uint8_t* smB = smem + A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smB + SB * B_STAGE);      // b_full[SB], b_empty[SB], a_ready, acc_full
    uint64_t* b_full = bars, *b_empty = bars + SB, *a_ready = bars + 2 * SB, *acc_full = a_ready + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tmem_slot + 4) + ((16u - (tc::smem_u32(tmem_slot + 4) & 15u)) & 15u));   // [bias_floats]
    float* sfirst = sbias + ((p.bias_floats + 3) & ~3);                    // per-tile first-layer bias [npad] + wxy [2 * npad]
    float* sx = sfirst + 3 * 384;                                          // x coordinate of every tile pixel [128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_row = p.R / ST_TILE;
    const long ntiles = (long)p.B * p.R * tiles_per_row;

    if (threadIdx.x == 0) {
        for (int s = 0; s < SB; ++s) { mbar_init(smem_u32(b_full + s), 1); mbar_init(smem_u32(b_empty + s), 1); }
        mbar_init(smem_u32(a_ready), 128); mbar_init(smem_u32(acc_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        for (int l = 0; l < p.nl; ++l) asm volatile("prefetch.tensormap [%0];\n" :: "l"(&maps.w[l]) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.bias_floats; i += ST_THREADS) sbias[i] = __ldg(p.bias_table + i);
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {       // ===== TMA producer: the weight tiles of every layer of every tile, in order =====
            uint32_t it = 0;
            for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int l = 0; l < p.nl; ++l) {
                    const StLayer& L = p.L[l];
                    const int nsl = L.npad / L.nb, nkc = (L.kpad + 63) >> 6;
                    for (int ns = 0; ns < nsl; ++ns)
                        for (int kc = 0; kc < nkc; ++kc, ++it) {
                            const int s = it % SB;
                            mbar_wait(smem_u32(b_empty + s), ((it / SB) & 1) ^ 1);
                            mbar_expect_tx(smem_u32(b_full + s), L.nb * 128);
                            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
                                         :: "r"(smem_u32(smB + s * B_STAGE)), "l"(&maps.w[l]), "r"(smem_u32(b_full + s)), "r"(kc * 64), "r"(ns * L.nb) : "memory");
                        }
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {       // ===== MMA issuer =====
            uint32_t it = 0, ar = 0;
            for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int l = 0; l < p.nl; ++l, ++ar) {
                    const StLayer& L = p.L[l];
                    const int nsl = L.npad / L.nb, nkc = (L.kpad + 63) >> 6;
                    mbar_wait(smem_u32(a_ready), ar & 1);                  // operand written, previous accumulator drained
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t idesc = (1u << 4) | ((uint32_t)(L.nb >> 3) << 17) | ((128u >> 4) << 24);
                    for (int ns = 0; ns < nsl; ++ns)
                        for (int kc = 0; kc < nkc; ++kc, ++it) {
                            const int s = it % SB;
                            mbar_wait(smem_u32(b_full + s), (it / SB) & 1);
                            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                            const uint64_t adesc = make_smem_desc_sw<128>(smem_u32(smA + kc * (ST_TILE * 128)));
                            const uint64_t bdesc = make_smem_desc_sw<128>(smem_u32(smB + s * B_STAGE));
                            const int ksteps = min(4, (L.kpad - kc * 64) >> 4);       // K tail: columns beyond kpad hold stale operand data
                            for (int k = 0; k < ksteps; ++k)
                                umma_f16(tmem_base + (uint32_t)(ns * L.nb), adesc + 2 * k, bdesc + 2 * k, idesc, (kc > 0 || k > 0) ? 1u : 0u);
                            umma_commit(smem_u32(b_empty + s));
                        }
                    umma_commit(smem_u32(acc_full));
                }
        }
    } else {                   // ===== compute warps: thread = pixel (tile row) = TMEM lane =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int te = threadIdx.x - 64;                                   // 0..127, for cooperative loops
        uint32_t af = 0;
        for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int n = (int)(tile / ((long)p.R * tiles_per_row));
            const int rem = (int)(tile - (long)n * p.R * tiles_per_row);
            const int y = rem / tiles_per_row, x0 = (rem - y * tiles_per_row) * ST_TILE;
            const float yv = __ldg(p.base + y);
            sx[te] = __ldg(p.base + x0 + te);
            {   // this tile's per-sample first-layer terms: bias (b + Wpose . pose[n]) and the xy weights
                const bool elementwise = (MODE == SM_BODY0 || MODE == SM_FACE);
                const int np = elementwise ? p.e_npad : p.L[0].npad;
                const float* pb = elementwise ? p.e_pb + (size_t)n * p.e_pb_ld : p.f_pb + (size_t)n * p.f_pb_ld;
                const float* wxy = elementwise ? p.e_wxy : p.f_wxy;
                for (int i = te; i < np; i += 128) {
                    sfirst[i] = __ldg(pb + i);
                    sfirst[384 + 2 * i] = __ldg(wxy + 2 * i); sfirst[384 + 2 * i + 1] = __ldg(wxy + 2 * i + 1);
                }
            }
            asm volatile("bar.sync 1, 128;\n" ::: "memory");
            // ---- prologue: the tile's first operand ----
            if (MODE == SM_BODY0 || MODE == SM_FACE) {
                const int groups = p.e_npad >> 3;
                for (int i = te; i < ST_TILE * groups; i += 128) {
                    const int r = i / groups, c8 = (i - r * groups) << 3;
                    const float xv = sx[r];
                    uint4 pk;
                    __half2* h2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = c8 + 2 * e;
                        const float v0 = sfirst[c] + sfirst[384 + 2 * c] * xv + sfirst[384 + 2 * c + 1] * yv;
                        const float v1 = sfirst[c + 1] + sfirst[384 + 2 * c + 2] * xv + sfirst[384 + 2 * c + 3] * yv;
                        h2[e] = __floats2half2_rn(st_sin(v0), st_sin(v1));
                    }
                    *reinterpret_cast<uint4*>(smA + a_off(r, c8)) = pk;
                }
            } else {
                // bilinear x2 of the previous level (align_corners = False).  Thread = tile pixel: ONE horizontal tap pair, four corner
                // row pointers and four packed-half weights per thread and tile, then per 8-channel group four 16-byte loads,
                // 1 HMUL2 + 3 HFMA2 per channel pair, one 16-byte store into the swizzled operand.  The first version walked
                // (pixel, group) items: an integer division, a lerp_locate, four 64-bit addresses and 32 half->float conversions
                // per item made this prologue 25 % of the level's instructions and -- with its three exposed L2 round trips --
                // 41 % of the compute warps' time, more than the sine layers (ncu source page, profiles/r02_siren_tc_notes.txt).
                // The taps 0 / 0.25 / 0.75 / 1 and their products are exact in fp16; the weighted sum is rounded per operation
                // (<= 2 ulp of the fp16 operand it becomes) instead of once.
                const int Rh = p.R >> 1, CP = p.prev_c, groups = CP >> 3;
                const __half* prev = p.prev + (size_t)n * Rh * Rh * CP;
                const LerpTap ty = lerp_locate(y, 0.5f, Rh);
                const LerpTap tx = lerp_locate(x0 + te, 0.5f, Rh);
                const uint4* pa = reinterpret_cast<const uint4*>(prev + ((size_t)ty.i0 * Rh + tx.i0) * CP);
                const uint4* pb = reinterpret_cast<const uint4*>(prev + ((size_t)ty.i0 * Rh + tx.i1) * CP);
                const uint4* pc = reinterpret_cast<const uint4*>(prev + ((size_t)ty.i1 * Rh + tx.i0) * CP);
                const uint4* pd = reinterpret_cast<const uint4*>(prev + ((size_t)ty.i1 * Rh + tx.i1) * CP);
                const __half2 w00 = __float2half2_rn(ty.l0 * tx.l0), w01 = __float2half2_rn(ty.l0 * tx.l1);
                const __half2 w10 = __float2half2_rn(ty.l1 * tx.l0), w11 = __float2half2_rn(ty.l1 * tx.l1);
#pragma unroll 1
                for (int cg0 = 0; cg0 < groups; cg0 += 4) {
                    uint4 va[4], vb[4], vc[4], vd[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int cg = min(cg0 + u, groups - 1);
                        va[u] = __ldg(pa + cg); vb[u] = __ldg(pb + cg); vc[u] = __ldg(pc + cg); vd[u] = __ldg(pd + cg);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (cg0 + u >= groups) break;
                        const __half2* ah = reinterpret_cast<const __half2*>(&va[u]); const __half2* bh = reinterpret_cast<const __half2*>(&vb[u]);
                        const __half2* ch = reinterpret_cast<const __half2*>(&vc[u]); const __half2* dh = reinterpret_cast<const __half2*>(&vd[u]);
                        uint4 o;
                        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            oh[k] = __hfma2(w11, dh[k], __hfma2(w10, ch[k], __hfma2(w01, bh[k], __hmul2(w00, ah[k]))));
                        *reinterpret_cast<uint4*>(smA + a_off(te, (cg0 + u) * 8)) = o;
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            mbar_arrive(smem_u32(a_ready));
            // ---- the layer chain ----
            for (int l = 0; l < p.nl; ++l, ++af) {
                const StLayer& L = p.L[l];
                mbar_wait(smem_u32(acc_full), af & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                if (L.sine) {
                    const float* lb = sbias + L.bias_off;
                    const float xv = sx[row];
#pragma unroll 1
                    for (int c0 = 0; c0 < L.npad; c0 += 32) {
                        uint32_t acc[32];
                        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
#pragma unroll
                        for (int g8 = 0; g8 < 4; ++g8) {
                            uint4 pk;
                            __half2* h2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int c = c0 + g8 * 8 + 2 * e;
                                float v0 = __uint_as_float(acc[g8 * 8 + 2 * e]), v1 = __uint_as_float(acc[g8 * 8 + 2 * e + 1]);
                                if (L.first) {
                                    v0 += sfirst[c] + sfirst[384 + 2 * c] * xv + sfirst[384 + 2 * c + 1] * yv;
                                    v1 += sfirst[c + 1] + sfirst[384 + 2 * c + 2] * xv + sfirst[384 + 2 * c + 3] * yv;
                                } else {
                                    v0 += lb[c]; v1 += lb[c + 1];
                                }
                                h2[e] = __floats2half2_rn(st_sin(v0), st_sin(v1));
                            }
                            *reinterpret_cast<uint4*>(smA + a_off(row, c0 + g8 * 8)) = pk;
                        }
                    }
                    const bool last = (l == p.nl - 1);
                    if (!last) {
                        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                        mbar_arrive(smem_u32(a_ready));
                    } else {
                        // levels 0 / 1: the last sine layer's output is the level's output tensor (fp16 NHWC)
                        asm volatile("bar.sync 1, 128;\n" ::: "memory");
                        const int groups = p.out_c >> 3;
                        __half* dst = p.out + (((size_t)n * p.R + y) * p.R + x0) * p.out_c;
                        for (int i = te; i < ST_TILE * groups; i += 128) {
                            const int r = i / groups, cg = i - r * groups;
                            *reinterpret_cast<uint4*>(dst + (size_t)r * p.out_c + cg * 8) = *reinterpret_cast<const uint4*>(smA + a_off(r, cg * 8));
                        }
                        asm volatile("bar.sync 1, 128;\n" ::: "memory");      // the next tile's prologue overwrites the operand
                    }
                } else {
                    // linear head: 16 accumulator columns, thread = pixel
                    uint32_t acc[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16), acc);
                    const int x = x0 + row;
                    if (MODE == SM_FACE) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            p.face_out[(((size_t)n * 4 + c) * p.R + y) * p.R + x] = __uint_as_float(acc[c]) + __ldg(p.head_bias + c);
                    } else {
                        float o[7];
#pragma unroll
                        for (int c = 0; c < 7; ++c) o[c] = __uint_as_float(acc[c]) + __ldg(p.head_bias + c);   // grid_change(0,1) alpha(2) colour(3..6)
                        const GsTap t = gs_locate(sx[row], yv, o[0], o[1], p.R, p.R);
                        float w[4];
                        gs_sample<4>(p.image.p + n * p.image.sn, p.image.sc, p.image.sh, p.R, p.R, t, w);
                        const size_t plane = (size_t)p.R * p.R, pix = (size_t)y * p.R + x;
                        const float alpha = o[2];
                        if (p.o_f16) {
                            __half* const* oh = reinterpret_cast<__half* const*>(p.o);
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                oh[0][((size_t)n * 4 + c) * plane + pix] = __float2half_rn((1.0f - alpha) * w[c] + alpha * o[3 + c]);
                                oh[2][((size_t)n * 4 + c) * plane + pix] = __float2half_rn(o[3 + c]);
                                oh[3][((size_t)n * 4 + c) * plane + pix] = __float2half_rn(w[c]);
                            }
                            oh[1][(size_t)n * plane + pix] = __float2half_rn(alpha);
                            oh[4][((size_t)n * 2) * plane + pix] = __float2half_rn(o[0]);
                            oh[4][((size_t)n * 2 + 1) * plane + pix] = __float2half_rn(o[1]);
                        } else {
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                p.o[0][((size_t)n * 4 + c) * plane + pix] = (1.0f - alpha) * w[c] + alpha * o[3 + c];
                                p.o[2][((size_t)n * 4 + c) * plane + pix] = o[3 + c];
                                p.o[3][((size_t)n * 4 + c) * plane + pix] = w[c];
                            }
                            p.o[1][(size_t)n * plane + pix] = alpha;
                            p.o[4][((size_t)n * 2) * plane + pix] = o[0];
                            p.o[4][((size_t)n * 2 + 1) * plane + pix] = o[1];
                        }
                    }
                    asm volatile("bar.sync 1, 128;\n" ::: "memory");          // sx / sfirst / the operand are rewritten by the next tile
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn st_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        THA4_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
        THA4_REQUIRE(ptr != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled unavailable");
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// W: [rows][kpad] fp16 K-major; box {64 k, nb rows}; rows beyond `rows` and k beyond kpad are zero-filled by TMA
CUtensorMap weight_tile_map(const void* W, int rows, int kpad, int nb) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)nb};
    cuuint32_t es[2] = {1, 1};
    CUresult r = st_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(W), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    THA4_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(siren weights) failed: " + std::to_string((int)r));
    return m;
}

template <int ACH, int NBMAX, int SB, int TMEM_COLS, int MODE>
void launch_siren_tc(const StMaps& maps, const StParams& p, int ctas_per_sm, cudaStream_t s) {
    const size_t smem = 1024 + (size_t)ACH * ST_TILE * 128 + (size_t)SB * NBMAX * 128 + (2 * SB + 2) * 8 + 16 + 16 +
                        ((size_t)((p.bias_floats + 3) & ~3) + 3 * 384 + 128) * sizeof(float);
    THA4_REQUIRE(smem <= 227 * 1024, "siren_tc: shared memory budget");
    THA4_ENSURE_SMEM((siren_tc_kernel<ACH, NBMAX, SB, TMEM_COLS, MODE>), smem);
    const long ntiles = (long)p.B * p.R * (p.R / ST_TILE);
    const int grid = (int)std::min<long>(ntiles, 148L * ctas_per_sm);
    siren_tc_kernel<ACH, NBMAX, SB, TMEM_COLS, MODE><<<grid, ST_THREADS, smem, s>>>(maps, p);
    THA4_LAUNCH_CHECK();
}

bool g_siren_tc = true;

}  // namespace

void siren_tc_enable(bool on) { g_siren_tc = on; }
bool siren_tc_enabled() { return g_siren_tc; }

void SirenTcPlan::add(const SirenLayer& l, int nb, int sine, int first) {
    THA4_REQUIRE(nl < 8, "siren_tc: too many layers");
    kpad[nl] = l.KPAD; npad[nl] = sine ? l.NPAD : 16; this->nb[nl] = nb; this->sine[nl] = sine; this->first[nl] = first;
    W[nl] = l.W; rows[nl] = l.NPAD; bias[nl] = l.bias;
    ++nl;
}

void siren_tc_run(Runtime& rt, int mode, const SirenTcPlan& plan, const SirenTcLevel& lv) {
    StParams p{};
    StMaps maps;
    p.R = lv.R; p.B = lv.B; p.nl = plan.nl;
    // bias table (device, rebuilt per call from the layers' bias vectors: tiny)
    int off = 0;
    for (int l = 0; l < plan.nl; ++l) {
        p.L[l].kpad = plan.kpad[l]; p.L[l].npad = plan.npad[l]; p.L[l].nb = plan.nb[l]; p.L[l].sine = plan.sine[l]; p.L[l].first = plan.first[l];
        p.L[l].bias_off = off;
        if (plan.sine[l]) off += plan.npad[l];
        maps.w[l] = weight_tile_map(plan.W[l], plan.rows[l], plan.kpad[l], plan.nb[l]);
    }
    p.bias_floats = off;
    float* table = rt.persist->alloc((size_t)std::max(off, 4));
    for (int l = 0; l < plan.nl; ++l)
        if (plan.sine[l])
            THA4_CUDA_CHECK(cudaMemcpyAsync(table + p.L[l].bias_off, plan.bias[l], plan.npad[l] * sizeof(float), cudaMemcpyDeviceToDevice, rt.stream));
    p.bias_table = table;
    p.e_npad = lv.e_npad; p.e_pb = lv.e_pb; p.e_pb_ld = lv.e_pb_ld; p.e_wxy = lv.e_wxy;
    p.f_pb = lv.f_pb; p.f_pb_ld = lv.f_pb_ld; p.f_wxy = lv.f_wxy;
    p.base = base_grid_table(lv.R);
    p.prev = lv.prev; p.prev_c = lv.prev_c; p.out = lv.out; p.out_c = lv.out_c;
    p.image = lv.image;
    for (int i = 0; i < 5; ++i) p.o[i] = lv.o[i];
    p.o_f16 = lv.o_f16 ? 1 : 0;
    p.face_out = lv.face_out; p.head_bias = lv.head_bias;
    cudaStream_t s = rt.stream;
    ProfScope prof(PROF_SIREN, s);
    //                         A chunks, widest weight tile, ring, TMEM columns
    if (mode == SM_BODY0) launch_siren_tc<6, 192, 4, 512, SM_BODY0>(maps, p, 1, s);
    else if (mode == SM_BODY1) launch_siren_tc<3, 96, 4, 256, SM_BODY1>(maps, p, 2, s);
    else if (mode == SM_BODY2) launch_siren_tc<2, 96, 2, 128, SM_BODY2>(maps, p, 4, s);      // 56 KB per CTA: four tiles in flight per SM
    else launch_siren_tc<2, 128, 2, 128, SM_FACE>(maps, p, 3, s);
}

}  // namespace tha4
