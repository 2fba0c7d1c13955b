// Persistent tcgen05 implicit-GEMM convolution for grids that fill the GPU (no K split):
//   * one CTA per SM walks a static round-robin list of (sample, 8x16 pixel tile, N tile, phase) work items;
//   * the activation HALO of a 32-channel chunk is loaded ONCE per work item and reused by every tap: for each distinct
//     horizontal tap offset dx one 4-D TMA box {32 ch, 8 w, 16 + (ndy-1) h, 1 n} is fetched; a box row is 8 pixels x 128 B
//     = exactly one 1024-byte swizzle atom, so the vertical tap offset dy is a whole-atom shift of the UMMA descriptor
//     start address (always 1024-byte aligned, canonical K-major SWIZZLE_128B).  A 3x3 conv thus moves 3x18 KB of A per
//     chunk instead of 9x16 KB;
//   * weights stream per (chunk, tap) through their own TMA ring;
//   * the fp32 accumulator is double-buffered in TMEM (2 x BN columns): the four epilogue warps drain tile i (bias,
//     residual, NHWC stores, per-channel statistics) while the MMA thread already accumulates tile i+1.
// The non-persistent kernel in conv_tc.cu keeps the small grids (cluster split-K).
#include "conv.cuh"
#include "profiler.cuh"
#include "tc_common.cuh"
#include <cuda.h>
#include <map>
#include <tuple>

namespace tha4 {
namespace {

using namespace tc;

constexpr int PT_W = 8, PT_H = 16;               // 128 output pixels per tile
constexpr int P_THREADS = 192;
constexpr int A_STAGES = 2;
constexpr int MAX_COPIES = 3, MAX_ROWS = PT_H + 2;
constexpr int A_STAGE_BYTES = MAX_COPIES * MAX_ROWS * 1024;     // 54 KB

struct TcpParams {
    float* out; int outH, outW, outC, out_ld;
    const float* bias;
    const float* res; int resH, resW, res_ld, res_mode;
    double* stats; int stats_ld; int stats_rep; long stats_rep_stride;
    int N, MH, MW, tiles_x, tiles_y, tiles_n, nphase, ntaps, cpt, out_mul, rows;
    long total_work;
    signed char dxmin[CONV_MAX_PHASES], ndx[CONV_MAX_PHASES], dymin[CONV_MAX_PHASES];
    signed char tap_copy[CONV_MAX_PHASES][CONV_MAX_TAPS], tap_row[CONV_MAX_PHASES][CONV_MAX_TAPS];
    signed char ph_oy[CONV_MAX_PHASES], ph_ox[CONV_MAX_PHASES];
};


struct Work { int n, ty, tx, nt, phase; };
__device__ __forceinline__ Work decode(long w, const TcpParams& p) {
    Work k;
    k.phase = (int)(w % p.nphase); w /= p.nphase;
    k.nt = (int)(w % p.tiles_n); w /= p.tiles_n;
    k.tx = (int)(w % p.tiles_x); w /= p.tiles_x;
    k.ty = (int)(w % p.tiles_y);
    k.n = (int)(w / p.tiles_y);
    return k;
}

template <int BN, int B_STAGES>
__global__ void __launch_bounds__(P_THREADS, 1) conv_tcp_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB, const TcpParams p) {
    constexpr int B_BYTES = BN * KCH * 4;
    constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smA = smem;
    uint8_t* smB = smA + A_STAGES * A_STAGE_BYTES;
    float* scratch = reinterpret_cast<float*>(smB + B_STAGES * B_BYTES);            // [4][32*33]
    float2* part = reinterpret_cast<float2*>(scratch + 4 * 32 * 33);               // [4][BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(part + 4 * BN);
    uint64_t* a_full = bars, *a_empty = bars + A_STAGES;
    uint64_t* b_full = bars + 2 * A_STAGES, *b_empty = b_full + B_STAGES;
    uint64_t* t_full = b_empty + B_STAGES, *t_empty = t_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < A_STAGES; ++s) { mbar_init(smem_u32(a_full + s), 1); mbar_init(smem_u32(a_empty + s), 1); }
        for (int s = 0; s < B_STAGES; ++s) { mbar_init(smem_u32(b_full + s), 1); mbar_init(smem_u32(b_empty + s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(t_full + s), 1); mbar_init(smem_u32(t_empty + s), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("prefetch.tensormap [%0];\n" :: "l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" :: "l"(&tmB) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int copy_bytes = p.rows * 1024;

    if (warp == 0) {
        if (lane == 0) {   // ===== TMA producer =====
            int a_it = 0, b_it = 0;
            for (long w = blockIdx.x; w < p.total_work; w += gridDim.x) {
                const Work k = decode(w, p);
                const int x0 = k.tx * PT_W + p.dxmin[k.phase], y0 = k.ty * PT_H + p.dymin[k.phase];
                const int ndx = p.ndx[k.phase];
                for (int c = 0; c < p.cpt; ++c) {
                    const int sa = a_it % A_STAGES;
                    mbar_wait(smem_u32(a_empty + sa), ((a_it / A_STAGES) & 1) ^ 1);
                    const uint32_t af = smem_u32(a_full + sa);
                    mbar_expect_tx(af, ndx * copy_bytes);
                    for (int ci = 0; ci < ndx; ++ci)
                        tma_load_4d(smem_u32(smA + sa * A_STAGE_BYTES + ci * copy_bytes), &tmA, c * KCH, x0 + ci, y0, k.n, af);
                    ++a_it;
                    for (int tap = 0; tap < p.ntaps; ++tap) {
                        const int sb = b_it % B_STAGES;
                        mbar_wait(smem_u32(b_empty + sb), ((b_it / B_STAGES) & 1) ^ 1);
                        const uint32_t bf = smem_u32(b_full + sb);
                        mbar_expect_tx(bf, B_BYTES);
                        tma_load_3d(smem_u32(smB + sb * B_BYTES), &tmB, c * KCH, k.nt * BN, k.phase * p.ntaps + tap, bf);
                        ++b_it;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ===== MMA issuer =====
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
            int a_it = 0, b_it = 0, t_it = 0;
            for (long w = blockIdx.x; w < p.total_work; w += gridDim.x) {
                const Work k = decode(w, p);
                const int buf = t_it & 1;
                mbar_wait(smem_u32(t_empty + buf), ((t_it >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                for (int c = 0; c < p.cpt; ++c) {
                    const int sa = a_it % A_STAGES;
                    mbar_wait(smem_u32(a_full + sa), (a_it / A_STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t a_base = smem_u32(smA + sa * A_STAGE_BYTES);
                    for (int tap = 0; tap < p.ntaps; ++tap) {
                        const int sb = b_it % B_STAGES;
                        mbar_wait(smem_u32(b_full + sb), (b_it / B_STAGES) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                        const uint64_t adesc = make_smem_desc(a_base + p.tap_copy[k.phase][tap] * copy_bytes + p.tap_row[k.phase][tap] * 1024);
                        const uint64_t bdesc = make_smem_desc(smem_u32(smB + sb * B_BYTES));
#pragma unroll
                        for (int kk = 0; kk < KCH / 8; ++kk)
                            umma_tf32(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (c > 0 || tap > 0 || kk > 0) ? 1u : 0u);
                        umma_commit(smem_u32(b_empty + sb));
                        ++b_it;
                    }
                    umma_commit(smem_u32(a_empty + sa));
                    ++a_it;
                }
                umma_commit(smem_u32(t_full + buf));
                ++t_it;
            }
        }
    } else {               // ===== epilogue warps =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        float* sc = scratch + q * (32 * 33);
        int t_it = 0;
        for (long w = blockIdx.x; w < p.total_work; w += gridDim.x) {
            const Work k = decode(w, p);
            const int buf = t_it & 1;
            mbar_wait(smem_u32(t_full + buf), (t_it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const int my = k.ty * PT_H + row / PT_W, mx = k.tx * PT_W + row % PT_W;
            const bool valid = my < p.MH && mx < p.MW;
            const int oy = my * p.out_mul + p.ph_oy[k.phase], ox = mx * p.out_mul + p.ph_ox[k.phase];
            float* orow = p.out + (((long)k.n * p.outH + oy) * p.outW + ox) * p.out_ld;
            const int n0 = k.nt * BN;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c0), r);
                if (c0 + 32 >= BN) {          // last read of this accumulator: hand the TMEM buffer back to the MMA thread
                    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(t_empty + buf));
                }
                const int cbase = n0 + c0;
                if (cbase >= p.outC) continue;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                const int cn = min(32, p.outC - cbase);
                if (valid) {
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (j < cn) v[j] += __ldg(p.bias + cbase + j);
                    }
                    if (p.res_mode == RES_SAME || p.res_mode == RES_UP2) {
                        const int ry = p.res_mode == RES_UP2 ? (oy >> 1) : oy, rx = p.res_mode == RES_UP2 ? (ox >> 1) : ox;
                        const float* rr = p.res + (((long)k.n * p.resH + ry) * p.resW + rx) * p.res_ld + cbase;
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (j < cn) v[j] += rr[j];
                    } else if (p.res_mode == RES_DOWN2) {
                        const float* rr = p.res + (((long)k.n * p.resH + 2 * oy) * p.resW + 2 * ox) * p.res_ld + cbase;
                        const long dx1 = p.res_ld, dy1 = (long)p.resW * p.res_ld;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < cn) v[j] += 0.25f * ((rr[j] + rr[dx1 + j]) + (rr[dy1 + j] + rr[dy1 + dx1 + j]));
                    }
                    if (cn == 32) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(orow + cbase + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (j < cn) orow[cbase + j] = v[j];
                    }
                }
                if (p.stats) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) sc[lane * 33 + j] = valid ? v[j] : 0.0f;
                    __syncwarp();
                    float su = 0.0f, sq = 0.0f;
#pragma unroll 8
                    for (int rr = 0; rr < 32; ++rr) { const float t = sc[rr * 33 + lane]; su += t; sq += t * t; }
                    __syncwarp();
                    part[q * BN + c0 + lane] = make_float2(su, sq);
                }
            }
            if (p.stats) {
                asm volatile("bar.sync 1, 128;\n" ::: "memory");
                double* base = p.stats + (long)(w % p.stats_rep) * p.stats_rep_stride + ((long)k.n * p.stats_ld + n0) * 2;
                for (int c = (warp - 2) * 32 + lane; c < BN; c += 128) {
                    if (n0 + c >= p.outC) break;
                    const float2 a = part[c], b = part[BN + c], cc = part[2 * BN + c], d = part[3 * BN + c];
                    atomicAdd(base + 2 * c, (double)a.x + (double)b.x + (double)cc.x + (double)d.x);
                    atomicAdd(base + 2 * c + 1, (double)a.y + (double)b.y + (double)cc.y + (double)d.y);
                }
                asm volatile("bar.sync 1, 128;\n" ::: "memory");
            }
            ++t_it;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
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

using MapKey = std::tuple<const void*, long, long, long, long, long, int>;
std::map<MapKey, CUtensorMap> g_maps;

const CUtensorMap& activation_map(const View& v, int rows) {
    MapKey key{v.p, v.N, v.H, v.W, v.C, v.ld, rows};
    auto it = g_maps.find(key);
    if (it != g_maps.end()) return it->second;
    CUtensorMap m;
    cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
    cuuint64_t strides[3] = {(cuuint64_t)v.ld * 4, (cuuint64_t)v.W * v.ld * 4, (cuuint64_t)v.H * v.W * v.ld * 4};
    cuuint32_t box[4] = {KCH, PT_W, (cuuint32_t)rows, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, v.p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    THA4_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation halo) failed: " + std::to_string((int)r));
    return g_maps.emplace(key, m).first->second;
}

const CUtensorMap& weight_map(const ConvWeights& cw, int bn) {
    MapKey key{cw.w, cw.cin_pad, cw.cout_pad, cw.ntaps, cw.nphase, 1, bn};
    auto it = g_maps.find(key);
    if (it != g_maps.end()) return it->second;
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)cw.cin_pad, (cuuint64_t)cw.cout_pad, (cuuint64_t)cw.ntaps * cw.nphase};
    cuuint64_t strides[2] = {(cuuint64_t)cw.cin_pad * 4, (cuuint64_t)cw.cout_pad * cw.cin_pad * 4};
    cuuint32_t box[3] = {KCH, (cuuint32_t)bn, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, cw.w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    THA4_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r));
    return g_maps.emplace(key, m).first->second;
}

template <int BN, int B_STAGES>
void launch_tcp(const CUtensorMap& ma, const CUtensorMap& mb, const TcpParams& p, int grid, cudaStream_t s) {
    constexpr size_t smem = 1024 + (size_t)A_STAGES * A_STAGE_BYTES + (size_t)B_STAGES * BN * KCH * 4 + 4 * 32 * 33 * 4 + 4 * BN * 8 +
                            (2 * A_STAGES + 2 * B_STAGES + 4) * 8 + 16;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    static bool configured = false;
    if (!configured) {
        THA4_CUDA_CHECK(cudaFuncSetAttribute(conv_tcp_kernel<BN, B_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    conv_tcp_kernel<BN, B_STAGES><<<grid, P_THREADS, smem, s>>>(ma, mb, p);
    THA4_LAUNCH_CHECK();
}

int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        THA4_CUDA_CHECK(cudaGetDevice(&dev));
        THA4_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    }
    return n;
}

bool g_use_persistent = false;   // measured slower than the non-persistent kernel so far (weight traffic dominates): opt-in

}  // namespace

void conv_tcp_enable(bool on) { g_use_persistent = on; }

// persistent path: stride-1 taps, grid large enough that K never needs splitting
bool conv_tcp_supported(const ConvWeights& cw, const ConvArgs& a) {
    if (!g_use_persistent || a.in.f16 || !conv_tc_supported(cw, a) || a.ksplit > 1) return false;
    const int MH = a.out.H / cw.out_mul, MW = a.out.W / cw.out_mul;
    const long tiles = (long)ceil_div(MW, PT_W) * ceil_div(MH, PT_H) * a.in.N * cw.nphase;
    const int bn = (cw.cout_pad % 128 == 0) ? 128 : (cw.cout_pad % 64 == 0 ? 64 : 32);
    return tiles * (cw.cout_pad / bn) >= 120;
}

void conv_tcp_forward(const ConvWeights& cw, const ConvArgs& a, cudaStream_t s) {
    THA4_REQUIRE(conv_tcp_supported(cw, a), "conv_tcp: unsupported configuration");
    THA4_REQUIRE(a.in.C == cw.cin && a.out.C == cw.cout && a.in.N == a.out.N, "conv_tcp: shapes");
    TcpParams p{};
    p.out = a.out.p; p.outH = a.out.H; p.outW = a.out.W; p.outC = a.out.C; p.out_ld = a.out.ld;
    p.bias = cw.bias;
    p.res = a.res.p; p.res_mode = a.res.p ? a.res_mode : RES_NONE;
    p.resH = a.res.H; p.resW = a.res.W; p.res_ld = a.res.ld;
    p.stats = a.out.stats; p.stats_ld = a.out.stats_ld;
    p.stats_rep = std::max(1, a.out.stats_rep); p.stats_rep_stride = a.out.stats_rep_stride;
    p.N = a.in.N; p.out_mul = cw.out_mul;
    p.MH = a.out.H / cw.out_mul; p.MW = a.out.W / cw.out_mul;
    THA4_REQUIRE(p.MH == a.in.H && p.MW == a.in.W, "conv_tcp: geometry");
    p.tiles_x = ceil_div(p.MW, PT_W); p.tiles_y = ceil_div(p.MH, PT_H);
    p.nphase = cw.nphase; p.ntaps = cw.ntaps; p.cpt = cw.cin_pad / KCH;
    int ndy_all = -1;
    for (int ph = 0; ph < cw.nphase; ++ph) {
        int dxmin = 127, dxmax = -127, dymin = 127, dymax = -127;
        for (int t = 0; t < cw.ntaps; ++t) {
            dxmin = std::min<int>(dxmin, cw.dx[ph][t]); dxmax = std::max<int>(dxmax, cw.dx[ph][t]);
            dymin = std::min<int>(dymin, cw.dy[ph][t]); dymax = std::max<int>(dymax, cw.dy[ph][t]);
        }
        THA4_REQUIRE(dxmax - dxmin + 1 <= MAX_COPIES && dymax - dymin + 1 <= 3, "conv_tcp: tap extent");
        if (ndy_all < 0) ndy_all = dymax - dymin + 1;
        THA4_REQUIRE(ndy_all == dymax - dymin + 1, "conv_tcp: phases must share the vertical tap extent");
        p.dxmin[ph] = (signed char)dxmin; p.ndx[ph] = (signed char)(dxmax - dxmin + 1); p.dymin[ph] = (signed char)dymin;
        p.ph_oy[ph] = cw.ph_oy[ph]; p.ph_ox[ph] = cw.ph_ox[ph];
        for (int t = 0; t < cw.ntaps; ++t) {
            p.tap_copy[ph][t] = (signed char)(cw.dx[ph][t] - dxmin);
            p.tap_row[ph][t] = (signed char)(cw.dy[ph][t] - dymin);
        }
    }
    p.rows = PT_H + ndy_all - 1;
    const int bn = (cw.cout_pad % 128 == 0) ? 128 : (cw.cout_pad % 64 == 0 ? 64 : 32);
    p.tiles_n = cw.cout_pad / bn;
    p.total_work = (long)p.tiles_x * p.tiles_y * p.N * p.tiles_n * p.nphase;
    const int grid = (int)std::min<long>(p.total_work, num_sms());
    ProfScope prof(PROF_CONV, s);
    prof_add_work(PROF_CONV, 2.0 * (double)p.N * p.MH * p.MW * cw.cout * cw.cin * cw.ntaps * cw.nphase, 0.0);
    const CUtensorMap& ma = activation_map(a.in, p.rows);
    const CUtensorMap& mb = weight_map(cw, bn);
    if (bn == 128) launch_tcp<128, 4>(ma, mb, p, grid, s);
    else if (bn == 64) launch_tcp<64, 6>(ma, mb, p, grid, s);
    else launch_tcp<32, 8>(ma, mb, p, grid, s);
}

}  // namespace tha4
