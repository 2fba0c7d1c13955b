// 3x3 (stride 1, pad 1) convolution on tcgen05 with HALO REUSE -- the kernel behind most of the teacher's FLOPs in the
// default mode (both 3x3 convs of every U-Net ResBlock, the eleven 512 -> 512 bottleneck convs of the encoder-decoder nets).
//
// conv_tc.cu fetches one 128-pixel x 64-channel activation box PER TAP (9 boxes per channel chunk) and, with the fused
// input normalisation, transforms each of them.  Here the CTA tile is 8 x 16 pixels and one TMA box {64 ch, 10 w, 18 h}
// brings in the tile's 10 x 18 HALO once per channel chunk; the nine taps are nine VIEWS of that shared-memory image:
// the UMMA A-descriptor start address is shifted by ((dy + 1) * 10 + (dx + 1)) pixel rows and the stride between 8-row
// groups (SBO) is the halo pitch (10 rows) -- the tensor core applies the 128-/64-byte swizzle on absolute shared-memory
// addresses, so row-shifted views of a TMA-written image are legal (profiles/r02_umma_row_shift_probe.txt).
//   * activation traffic L2 -> shared memory per channel chunk: 180 rows instead of 9 x 128 (6.4x less);
//   * the pending normalisation (XF: GroupNorm / InstanceNorm affine + FiLM + SiLU / ReLU of the RAW f16 input) is
//     applied to 180 rows once instead of 1152 rows, by the four warps that later drain the accumulator;
//   * weights stream per (chunk, tap) through their own TMA ring, pre-issued ahead of the programmatic-dependency wait;
//   * K (channel chunks) can be split over a thread-block cluster, partials meeting in distributed shared memory
//     (same epilogues as conv_tc.cu: conv_tc_device.cuh).
#include "conv.cuh"
#include "profiler.cuh"
#include "conv_tc_device.cuh"
#include <cuda.h>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

namespace tha4 {
static long long* g_dbg_buf = nullptr;   // THA4_HALO_DEBUG phase stamps

namespace {

using namespace tc;
using namespace tcdev;

constexpr int HT_W = 8, HT_H = 16;                 // CTA tile: 8 x 16 = 128 output pixels
constexpr int HALO_W = HT_W + 2, HALO_H = HT_H + 2, HALO_ROWS = HALO_W * HALO_H;     // 10 x 18 = 180 pixel rows

__device__ __forceinline__ uint64_t make_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout) {
    const uint32_t lo = ((smem_addr & 0x3FFFF) >> 4) | (1u << 16);
    const uint32_t hi = (sbo_bytes >> 4) | (1u << 14) | (layout << 29);
    return ((uint64_t)hi << 32) | lo;
}

__host__ __device__ constexpr int halo_a_bytes(int rowb) { return ((HALO_ROWS * rowb + 1023) / 1024) * 1024; }

// SA / SB: stages of the activation-halo ring / of the weight-tile ring.  OP: OP_F16 (64 channels per chunk, 128-byte rows)
// or OP_F16N (32 channels, 64-byte rows).  p.ksplit = cluster size CS (split over channel chunks), p.cpt = chunks.
// MINB: resident CTAs per SM the register allocation must allow (4 for the single-chunk unsplit variants, whose small
// rings fit four times: the layers at 256x256 / 512x512 are chains of dependent latencies, more CTAs = more overlap)
__host__ __device__ constexpr int halo_min_ctas(int bn, int sa, int cs, int op) {
    return cs > 1 ? 1 : ((sa == (op == OP_F16N ? 2 : 1) && bn <= 64) ? 4 : (bn >= 128 ? 2 : 3));
}
template <int BN, int SA, int SB, int CS, int OP, int XF>
__global__ void __launch_bounds__(TC_THREADS, halo_min_ctas(BN, SA, CS, OP)) conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                const __grid_constant__ CUtensorMap tmO32, const __grid_constant__ CUtensorMap tmO16,
                                                                const __grid_constant__ CUtensorMap tmR, const TcParams p) {
    static_assert(OP != OP_TF32, "halo kernel: f16 operands");
    constexpr int ROWB = op_row_bytes(OP);
    constexpr int KCE = op_kch(OP);
    constexpr int A_BYTES = halo_a_bytes(ROWB);
    constexpr int B_BYTES = BN * ROWB;
    constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    constexpr uint32_t LAYOUT = ROWB == 128 ? 2u : 4u;
    constexpr int NSLOT = CS == 1 ? epi_nslot(BN, (size_t)SA * A_BYTES + (size_t)SB * B_BYTES) : 0;     // TMA-store staging slots of the unsplit epilogue
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic (not an integer round trip) keeps the shared address space: LDS / STS, not generic LD / ST
    uint8_t* smA = smem;
    uint8_t* smB = smem + SA * A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smB + SB * B_BYTES);
    uint64_t* a_full = bars, *a_empty = bars + SA, *a_xf = bars + 2 * SA;
    uint64_t* b_full = bars + 3 * SA, *b_empty = b_full + SB;
    uint64_t* t_full = b_empty + SB;
    uint64_t* res_bars = t_full + 1;                                     // [4]: residual tiles of the unsplit epilogue (epi_direct)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bars + 4);
    float* xf_A = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tmem_slot + 4) + ((16u - (tc::smem_u32(tmem_slot + 4) & 15u)) & 15u));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // developer timing: CTA `slot` of every 97 records clock64 at its phase boundaries (8 stamps per role: dbg[slot][role][8])
    const bool dbg_on = p.dbg != nullptr && (blockIdx.x % 97) == 0 && blockIdx.y == 0 && (blockIdx.x / 97) * CS + blockIdx.z < 32;
    long long* dbg = dbg_on ? p.dbg + ((blockIdx.x / 97) * CS + blockIdx.z) * 32 : nullptr;
#define HSTAMP(role, i) do { if (dbg) dbg[(role) * 8 + (i)] = clock64(); } while (0)
    if (threadIdx.x == 0) HSTAMP(0, 0);
    int tile = blockIdx.x;
    const int tx = tile % p.tiles_x; tile /= p.tiles_x;
    const int ty = tile % p.tiles_y;
    const int n = tile / p.tiles_y;
    const int x0 = tx * HT_W, y0 = ty * HT_H;
    const int n0 = blockIdx.y * BN;
    const int split = blockIdx.z;                                        // rank in the cluster (CS == gridDim.z)
    const int c_per = (p.cpt + CS - 1) / CS;
    const int cb0 = split * c_per;
    const int nc = max(0, min(p.cpt, cb0 + c_per) - cb0);                // channel chunks of this CTA
    const int nb = nc * 9;                                               // weight tiles of this CTA

    if (threadIdx.x == 0) {
        for (int s = 0; s < SA; ++s) { mbar_init(smem_u32(a_full + s), 1); mbar_init(smem_u32(a_empty + s), 1); mbar_init(smem_u32(a_xf + s), 128); }
        for (int s = 0; s < SB; ++s) { mbar_init(smem_u32(b_full + s), 1); mbar_init(smem_u32(b_empty + s), 1); }
        mbar_init(smem_u32(t_full), 1);
        for (int s = 0; s < 4; ++s) mbar_init(smem_u32(res_bars + s), 1);
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
    if (threadIdx.x == 0) HSTAMP(0, 1);
    pdl_trigger();
    // weight tiles of the first ring pass do not depend on the previous kernel: fetch them ahead of the dependency wait
    const int npre = p.pre_b ? min(nb, SB) : 0;
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < npre; ++i) {
            const uint32_t full = smem_u32(b_full + i);
            mbar_expect_tx(full, B_BYTES);
            tma_load_3d(smem_u32(smB + i * B_BYTES), &tmB, (cb0 + i / 9) * KCE, n0, i % 9, full);
        }
    }
    pdl_wait();
    if (threadIdx.x == 0) HSTAMP(0, 2);

    if (nc > 0) {
        if (warp == 0) {
            if (lane == 0) {   // ===== TMA producer: one halo box per chunk, nine weight tiles per chunk =====
                int bi = 0;
                for (int ci = 0; ci < nc; ++ci) {
                    const int sa = ci % SA;
                    mbar_wait(smem_u32(a_empty + sa), ((ci / SA) & 1) ^ 1);
                    mbar_expect_tx(smem_u32(a_full + sa), HALO_ROWS * ROWB);
                    tma_load_4d(smem_u32(smA + sa * A_BYTES), &tmA, (cb0 + ci) * KCE, x0 - 1, y0 - 1, n, smem_u32(a_full + sa));
                    for (int tap = 0; tap < 9; ++tap, ++bi) {
                        if (bi < npre) continue;
                        const int sb = bi % SB;
                        mbar_wait(smem_u32(b_empty + sb), ((bi / SB) & 1) ^ 1);
                        mbar_expect_tx(smem_u32(b_full + sb), B_BYTES);
                        tma_load_3d(smem_u32(smB + sb * B_BYTES), &tmB, (cb0 + ci) * KCE, n0, tap, smem_u32(b_full + sb));
                    }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {   // ===== MMA issuer: 9 taps = 9 row-shifted views of the halo =====
                constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);     // f16 x f16 -> f32, K-major, M = 128
                int bi = 0;
                for (int ci = 0; ci < nc; ++ci) {
                    const int sa = ci % SA;
                    mbar_wait(smem_u32((XF ? a_xf : a_full) + sa), (ci / SA) & 1);
                    if (ci == 0) HSTAMP(1, 0);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t a_base = smem_u32(smA + sa * A_BYTES);
                    for (int tap = 0; tap < 9; ++tap, ++bi) {
                        const int sb = bi % SB;
                        mbar_wait(smem_u32(b_full + sb), (bi / SB) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                        const int shift = (p.dy[0][tap] + 1) * HALO_W + (p.dx[0][tap] + 1);
                        const uint64_t adesc = make_desc_sbo(a_base + shift * ROWB, HALO_W * ROWB, LAYOUT);
                        const uint64_t bdesc = make_smem_desc_sw<ROWB>(smem_u32(smB + sb * B_BYTES));
#pragma unroll
                        for (int k = 0; k < ROWB / 32; ++k)
                            umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (ci > 0 || tap > 0 || k > 0) ? 1u : 0u);
                        umma_commit(smem_u32(b_empty + sb));
                    }
                    umma_commit(smem_u32(a_empty + sa));                  // the halo of this chunk is free when its 9 taps retire
                }
                umma_commit(smem_u32(t_full));
                HSTAMP(1, 1);
            }
        } else {
            if (XF) {          // ===== warps 2-5: normalise each chunk's halo ONCE, in place; then they are the epilogue =====
                const int te = threadIdx.x - 64;
                __half* hA = reinterpret_cast<__half*>(xf_A);
                __half* hB = hA + p.xf_C;
                double2* chs = reinterpret_cast<double2*>(xf_A + 2 * p.xf_C);
                xf_build_coef(p, n, te, hA, hB, chs, cb0 * KCE, (cb0 + nc) * KCE);
                if (te == 0) HSTAMP(2, 0);
                const bool silu = p.xf_act == ACT_SILU || p.xf_act == ACT_SILU_FAST;
                for (int ci = 0; ci < nc; ++ci) {
                    const int sa = ci % SA;
                    mbar_wait(smem_u32(a_full + sa), (ci / SA) & 1);
                    if (te == 0 && ci == 0) HSTAMP(2, 1);
                    const int c0 = (cb0 + ci) * KCE;
                    // items = (halo row, half row): 360 items over 128 threads (3 at most) instead of 180 whole rows (2 at most:
                    // 52 threads did twice the work of the others and set the length of the stage)
                    constexpr int HC = ROWB / 32;                                             // chunks per half row
                    for (int it = te; it < 2 * HALO_ROWS; it += 128) {
                        const int row = it >> 1, half = it & 1;
                        const int hy = row / HALO_W, hx = row - hy * HALO_W;
                        const int iy = y0 - 1 + hy, ix = x0 - 1 + hx;
                        if (iy < 0 || iy >= p.inH || ix < 0 || ix >= p.inW) continue;      // zero padding stays zero
                        const int swz = ROWB == 128 ? (row & 7) : ((row >> 1) & 3);
                        xf_chunks<HC>(smA + sa * A_BYTES + row * ROWB, swz, half * HC, c0, p, hA, hB, silu);
                    }
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                    mbar_arrive(smem_u32(a_xf + sa));
                    if (te == 0 && ci == nc - 1) HSTAMP(2, 2);
                }
            }
            if (threadIdx.x == 64) { mbar_wait(smem_u32(t_full), 0); HSTAMP(2, 3); }
            if (CS > 1) { mbar_wait(smem_u32(t_full), 0); asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }   // own accumulator complete
            else epi_direct<BN, HT_W, NSLOT>(p, tmem_base, smem, smem_u32(t_full), n, y0, x0, n0, 0, 0, warp, lane, &tmO32, &tmO16, &tmR, res_bars);
            if (threadIdx.x == 64) HSTAMP(2, 4);
        }
    }
    if (CS > 1) {
        // barrier A: every CTA of the cluster has its accumulator and idle pipeline buffers -> peers may write into them
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        if (threadIdx.x == 64) HSTAMP(2, 5);
        if (warp >= 2 && nc > 0) epi_push_partial<BN, CS>(tmem_base, smem, split, warp, lane);
        // barrier B: the pushed slices are visible to their owners; nobody touches a peer's memory afterwards
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        if (threadIdx.x == 64) HSTAMP(2, 7);
        if (warp >= 2) epi_cluster_reduce<BN, CS, HT_W>(p, smem, n, y0, x0, n0, 0, split, warp, dbg);
        if (threadIdx.x == 64) HSTAMP(2, 6);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
    if (threadIdx.x == 0) HSTAMP(0, 3);
#undef HSTAMP
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn halo_encode() {
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

using HKey = std::tuple<int, const void*, long, long, long, long, long, int>;
std::map<HKey, CUtensorMap> g_halo_maps;
std::mutex g_halo_mu;

const CUtensorMap& halo_activation_map(const View& v, int op) {
    HKey key{current_device(), v.p, v.N, v.H, v.W, v.C, v.ld, op};
    std::lock_guard<std::mutex> lock(g_halo_mu);
    auto it = g_halo_maps.find(key);
    if (it != g_halo_maps.end()) return it->second;
    CUtensorMap m;
    cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
    cuuint64_t strides[3] = {(cuuint64_t)v.ld * 2, (cuuint64_t)v.W * v.ld * 2, (cuuint64_t)v.H * v.W * v.ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)op_kch(op), (cuuint32_t)HALO_W, (cuuint32_t)HALO_H, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = halo_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, v.p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               op == OP_F16N ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    THA4_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(halo activation) failed: " + std::to_string((int)r));
    return g_halo_maps.emplace(key, m).first->second;
}

const CUtensorMap& halo_weight_map(const ConvWeights& cw, int bn, int op) {
    HKey key{current_device(), cw.w16, cw.cin_pad, cw.cout_pad, cw.ntaps, bn, -1, op};
    std::lock_guard<std::mutex> lock(g_halo_mu);
    auto it = g_halo_maps.find(key);
    if (it != g_halo_maps.end()) return it->second;
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)cw.cin_pad, (cuuint64_t)cw.cout_pad, (cuuint64_t)cw.ntaps};
    cuuint64_t strides[2] = {(cuuint64_t)cw.cin_pad * 2, (cuuint64_t)cw.cout_pad * cw.cin_pad * 2};
    cuuint32_t box[3] = {(cuuint32_t)op_kch(op), (cuuint32_t)bn, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = halo_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, cw.w16, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               op == OP_F16N ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    THA4_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(halo weights) failed: " + std::to_string((int)r));
    return g_halo_maps.emplace(key, m).first->second;
}

// Output tile of the unsplit epilogue: box {32 channels, 16 x 8 pixels}, rows of 128 bytes (fp32) / 64 bytes (f16), swizzled
// like the staging writes of epi_direct.  Returns false when the view cannot be described (alignment).
bool halo_store_map(const View& v, bool f16, const CUtensorMap** out) {
    const int es = f16 ? 2 : 4;
    if (!v.p || (reinterpret_cast<uintptr_t>(v.p) & 15) || ((long)v.ld * es) % 16 != 0) return false;
    HKey key{current_device(), v.p, v.N, v.H, v.W, v.C, v.ld, f16 ? -2 : -3};
    std::lock_guard<std::mutex> lock(g_halo_mu);
    auto it = g_halo_maps.find(key);
    if (it == g_halo_maps.end()) {
        CUtensorMap m;
        cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
        cuuint64_t strides[3] = {(cuuint64_t)v.ld * es, (cuuint64_t)v.W * v.ld * es, (cuuint64_t)v.H * v.W * v.ld * es};
        cuuint32_t box[4] = {32, (cuuint32_t)HT_W, (cuuint32_t)HT_H, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = halo_encode()(&m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, v.p, dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, f16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
        it = g_halo_maps.emplace(key, m).first;
    }
    *out = &it->second;
    return true;
}

struct HaloPlan { int bn, tiles_x, tiles_y, tiles_m, tiles_n, cs, chunks; };

HaloPlan halo_plan(const ConvWeights& cw, const ConvArgs& a, int op) {
    HaloPlan pl;
    pl.tiles_x = ceil_div(a.out.W, HT_W); pl.tiles_y = ceil_div(a.out.H, HT_H);
    pl.tiles_m = pl.tiles_x * pl.tiles_y * a.in.N;
    pl.chunks = cw.cin_pad / op_kch(op);
    pl.bn = (cw.cout_pad % 256 == 0) ? 256 : (cw.cout_pad % 128 == 0 ? 128 : (cw.cout_pad % 64 == 0 ? 64 : 32));
    pl.cs = 1;
    long ctas = (long)pl.tiles_m * (cw.cout_pad / pl.bn);
    if (a.ksplit > 1) {
        while (pl.cs * 2 <= std::min(8, a.ksplit)) pl.cs *= 2;
    } else if (a.ksplit <= 0 && ctas < 120) {
        // Too few tiles to fill the GPU.  Narrow the N tiles first (down to 64 columns: more CTAs and nothing to exchange),
        // then split the channel chunks over a cluster.  The DSMEM exchange moves 128 x bn x 4 x (cs-1)/cs bytes per CTA at
        // ~11 B/clk: at bn = 256, cs = 4..8 it took twice as long as the MMA phase it parallelised (10 000 vs 4 700 cycles,
        // profiles/r02_halo_phase_stamps.txt).
        while (pl.bn > 64 && (long)pl.tiles_m * (cw.cout_pad / pl.bn) < 148 && cw.cout_pad % (pl.bn / 2) == 0) pl.bn /= 2;
        ctas = (long)pl.tiles_m * (cw.cout_pad / pl.bn);
        const int want = std::max(2, (int)((222 + ctas - 1) / ctas));       // >= 2: the cluster variants carry the deep weight ring
        while (pl.cs < 8 && pl.cs * 2 <= want && pl.cs * 2 <= pl.chunks) pl.cs *= 2;
        while (pl.bn > 32 && (long)pl.tiles_m * (cw.cout_pad / pl.bn) * pl.cs < 96 && cw.cout_pad % (pl.bn / 2) == 0) pl.bn /= 2;
    }
    while (pl.cs > 1 && (pl.cs > pl.chunks || ceil_div(pl.chunks, ceil_div(pl.chunks, pl.cs)) != pl.cs)) pl.cs /= 2;   // every rank owns chunks
    pl.tiles_n = cw.cout_pad / pl.bn;
    return pl;
}

template <int OP, int BN, int SA, int SB, int CS, int XF>
void launch_halo(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo32, const CUtensorMap& mo16, const CUtensorMap& mr, const TcParams& p, dim3 grid, cudaStream_t s) {
    constexpr int ROWB = op_row_bytes(OP);
    constexpr size_t ring = (size_t)SA * halo_a_bytes(ROWB) + (size_t)SB * BN * ROWB;
    constexpr size_t smem0 = 1024 + ring + (3 * SA + 2 * SB + 1 + 4) * 8 + 16;
    static_assert(smem0 <= 227 * 1024, "shared memory budget");
    static_assert(ring >= (size_t)4 * 32 * 33 * 4 + 4 * BN * 8, "epilogue scratch must fit in the pipeline buffers");
    static_assert(CS == 1 || ring >= (size_t)128 * BN * 4 + 128 * 8 * 4 + 128 * 4 * 4, "partial tile + statistics scratch must fit");
    const size_t smem = smem0 + (XF ? (size_t)24 * p.xf_C + 32 : 0);
    THA4_REQUIRE(smem <= 227 * 1024, "conv_halo: shared memory budget (fused input normalisation)");
    THA4_ENSURE_SMEM((conv_halo_kernel<BN, SA, SB, CS, OP, XF>), smem);
    launch_pdl(conv_halo_kernel<BN, SA, SB, CS, OP, XF>, grid, dim3(TC_THREADS), smem, s, CS, ma, mb, mo32, mo16, mr, p);
    THA4_LAUNCH_CHECK();
}

int halo_num_sms() {
    static int sms[THA4_MAX_DEVICES] = {};
    const int d = current_device();
    if (!sms[d]) THA4_CUDA_CHECK(cudaDeviceGetAttribute(&sms[d], cudaDevAttrMultiProcessorCount, d));
    return sms[d];
}

// SBD: weight-ring depth of the cluster split-K launches (few CTAs per SM, the ring is what hides the DRAM latency of weights
// that are fetched ahead of the dependency wait); SBS: depth of the unsplit launches (many tiles: a shallow ring keeps
// the CTA small so that 2 - 4 of them share an SM and overlap each other's load -> transform -> MMA -> drain chains).
template <int OP, int BN, int SA, int SBD, int SBS, int XF>
void launch_halo_cs(int cs, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo32, const CUtensorMap& mo16, const CUtensorMap& mr, const TcParams& p, dim3 grid, cudaStream_t s) {
    if constexpr (BN <= 64) {
        if (cs == 1 && p.cpt == 1) {     // one chunk: one halo, ever -> a ring that fits four times per SM
            launch_halo<OP, BN, (OP == OP_F16N ? 2 : 1), SBS, 1, XF>(ma, mb, mo32, mo16, mr, p, grid, s);
            return;
        }
    }
    if (cs == 8) launch_halo<OP, BN, SA, SBD, 8, XF>(ma, mb, mo32, mo16, mr, p, grid, s);
    else if (cs == 4) launch_halo<OP, BN, SA, SBD, 4, XF>(ma, mb, mo32, mo16, mr, p, grid, s);
    else if (cs == 2) launch_halo<OP, BN, SA, SBD, 2, XF>(ma, mb, mo32, mo16, mr, p, grid, s);
    else if ((long)grid.x * grid.y * grid.z <= halo_num_sms())
        // unsplit and at most one CTA per SM (e.g. 256 -> 256 channels at 128 x 128: 128 tiles): nothing shares the SM, so the
        // CTA takes the deep weight ring.  With two 32 KB stages the MMA phase of that layer ran at 294 cycles per 128 x 256 x 16
        // MMA (28 B/clk/SM of weights from L2; the tensor pipe needs 128 cycles): profiles/r02_halo_phase_stamps.txt, section F
        launch_halo<OP, BN, SA, SBD, 1, XF>(ma, mb, mo32, mo16, mr, p, grid, s);
    else launch_halo<OP, BN, SA, SBS, 1, XF>(ma, mb, mo32, mo16, mr, p, grid, s);
}

template <int OP, int XF>
void launch_halo_bn(int bn, int cs, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo32, const CUtensorMap& mo16, const CUtensorMap& mr, const TcParams& p, dim3 grid, cudaStream_t s) {
    constexpr int M = OP == OP_F16N ? 2 : 1;        // 64-byte rows: twice the stages for the same bytes in flight
    if (bn == 256) launch_halo_cs<OP, 256, 2 * M, 4 * M, 2 * M, XF>(cs, ma, mb, mo32, mo16, mr, p, grid, s);        // unsplit: 111 KB -> 2 CTAs / SM
    else if (bn == 128) launch_halo_cs<OP, 128, 2 * M, 6 * M, 3 * M, XF>(cs, ma, mb, mo32, mo16, mr, p, grid, s);   // unsplit:  95 KB -> 2 CTAs / SM
    else if (bn == 64) launch_halo_cs<OP, 64, 2 * M, 8 * M, 3 * M, XF>(cs, ma, mb, mo32, mo16, mr, p, grid, s);     // unsplit:  71 KB -> 3 CTAs / SM
    else launch_halo_cs<OP, 32, 2 * M, 9 * M, 5 * M, XF>(cs, ma, mb, mo32, mo16, mr, p, grid, s);                    // unsplit:  67 KB -> 3 CTAs / SM
}

bool g_use_halo = true;
bool g_tma_store = true;      // option "tma_store": unsplit epilogue through shared-memory staging + TMA stores

}  // namespace

void conv_halo_enable(bool on) { g_use_halo = on; }
void conv_halo_enable_tma_store(bool on) { g_tma_store = on; }

bool conv_halo_supported(const ConvWeights& cw, const ConvArgs& a) {
    if (!g_use_halo || !conv_tc_supported(cw, a)) return false;
    if (!a.in.f16 || cw.ntaps != 9 || cw.nphase != 1 || cw.stride != 1 || cw.out_mul != 1) return false;
    for (int t = 0; t < 9; ++t)
        if (cw.dy[0][t] < -1 || cw.dy[0][t] > 1 || cw.dx[0][t] < -1 || cw.dx[0][t] > 1) return false;
    return a.in.H == a.out.H && a.in.W == a.out.W;
}

bool conv_halo_fuses_stats(const ConvWeights&, const ConvArgs&) { return true; }       // unsplit or cluster split: always final

void conv_halo_forward(const ConvWeights& cw, const ConvArgs& a, cudaStream_t s) {
    THA4_REQUIRE(conv_halo_supported(cw, a), "conv_halo: unsupported configuration");
    THA4_REQUIRE(a.in.C == cw.cin && a.out.C == cw.cout && a.in.N == a.out.N, "conv_halo: shapes");
    const int op = cw.cin_pad % 64 == 0 ? OP_F16 : OP_F16N;
    TcParams p{};
    p.out = a.out.p; p.outH = a.out.H; p.outW = a.out.W; p.outC = a.out.C; p.out_ld = a.out.ld;
    p.out16 = a.out16.p ? a.out16.hp() : nullptr; p.out16_ld = a.out16.ld;
    if (a.out16.p) THA4_REQUIRE(a.out16.f16 && a.out16.N == a.out.N && a.out16.H == a.out.H && a.out16.W == a.out.W && a.out16.C == a.out.C, "conv_halo: f16 output copy geometry");
    p.inH = a.in.H; p.inW = a.in.W; p.inC = a.in.C;
    if (a.nin.on) {
        const ConvNormIn& ni = a.nin;
        THA4_REQUIRE(ni.stats != nullptr && ni.gamma != nullptr && ni.beta != nullptr, "conv_halo: fused input normalisation needs statistics and affine parameters");
        THA4_REQUIRE(ni.C > 0 && ni.C <= a.in.C && ni.C % 8 == 0 && ni.C <= 1024, "conv_halo: normalised channel count");
        THA4_REQUIRE(ni.groups == 0 || (ni.C == a.in.C && ni.C % ni.groups == 0), "conv_halo: GroupNorm spans the whole input");
        p.in_stats = ni.stats; p.in_stats_ld = ni.stats_ld; p.in_stats_rep = std::max(1, ni.stats_rep); p.in_stats_rep_stride = ni.stats_rep_stride;
        p.xf_C = ni.C; p.xf_groups = ni.groups; p.xf_act = ni.act;
        p.xf_inv_cnt = 1.0 / ((double)a.in.H * a.in.W * (ni.groups == 0 ? 1 : ni.C / ni.groups));
        p.xf_gamma = ni.gamma; p.xf_beta = ni.beta; p.xf_film0 = ni.film0; p.xf_film1 = ni.film1; p.xf_film1_ld = ni.film1_ld;
    }
    p.bias = cw.bias;
    p.res = a.res.p; p.res_mode = a.res.p ? a.res_mode : RES_NONE;
    p.resH = a.res.H; p.resW = a.res.W; p.res_ld = a.res.ld;
    p.N = a.in.N; p.out_mul = 1; p.in_mul = 1;
    const HaloPlan pl = halo_plan(cw, a, op);
    p.MH = a.out.H; p.MW = a.out.W; p.tiles_x = pl.tiles_x; p.tiles_y = pl.tiles_y;
    p.pre_b = cw.dynamic ? 0 : 1;
    p.ntaps = 9; p.cpt = pl.chunks;
    if (!cw.w16) { conv_make_half(cw, s); p.pre_b = 0; }
    p.acc_scale = 1.0f / cw.w16_scale;
    for (int t = 0; t < 9; ++t) { p.dy[0][t] = cw.dy[0][t]; p.dx[0][t] = cw.dx[0][t]; }
    p.ph_oy[0] = 0; p.ph_ox[0] = 0;
    p.ksplit = 1;                                    // the epilogues' "split-K through a workspace / atomics" modes are not used here
    p.ws = nullptr;
    p.stats = a.out.stats; p.stats_ld = a.out.stats_ld;
    p.stats_rep = std::max(1, a.out.stats_rep); p.stats_rep_stride = a.out.stats_rep_stride;
    static const bool dbg_env = getenv("THA4_HALO_DEBUG") != nullptr;
    if (dbg_env) {
        if (!g_dbg_buf) { THA4_CUDA_CHECK(cudaMalloc(&g_dbg_buf, 32 * 32 * sizeof(long long))); }
        THA4_CUDA_CHECK(cudaMemsetAsync(g_dbg_buf, 0, 32 * 32 * sizeof(long long), s));
        p.dbg = g_dbg_buf;
    }
    ProfScope prof(PROF_CONV, s);
    prof_add_work(PROF_CONV, 2.0 * (double)p.N * p.MH * p.MW * cw.cout * cw.cin * 9, 0.0);
    const CUtensorMap& ma = halo_activation_map(a.in, op);
    const CUtensorMap& mb = halo_weight_map(cw, pl.bn, op);
    const CUtensorMap* mo32 = &ma;                   // placeholders when an output does not leave through TMA
    const CUtensorMap* mo16 = &ma;
    const CUtensorMap* mr = &ma;
    p.st_tma = 0;
    if (pl.cs == 1 && g_tma_store) {
        if (a.out.p && halo_store_map(a.out, false, &mo32)) p.st_tma |= 1;
        if (a.out16.p && halo_store_map(a.out16, true, &mo16)) p.st_tma |= 2;
        // the residual of a ResBlock's second conv has the geometry of the fp32 output: it arrives through the same box
        if ((p.st_tma & 1) && a.res.p && p.res_mode == RES_SAME && !a.res.f16 && a.res.N == a.out.N && a.res.H == a.out.H &&
            a.res.W == a.out.W && a.res.C >= a.out.C && halo_store_map(a.res, false, &mr)) p.st_tma |= 4;
    }
    p.vec4 = ((!cw.bias || (reinterpret_cast<uintptr_t>(cw.bias) & 15) == 0) &&
              (!a.res.p || ((reinterpret_cast<uintptr_t>(a.res.p) & 15) == 0 && a.res.ld % 4 == 0))) ? 1 : 0;
    dim3 grid(pl.tiles_m, pl.tiles_n, pl.cs);
    if (op == OP_F16) {
        if (a.nin.on) launch_halo_bn<OP_F16, 1>(pl.bn, pl.cs, ma, mb, *mo32, *mo16, *mr, p, grid, s);
        else launch_halo_bn<OP_F16, 0>(pl.bn, pl.cs, ma, mb, *mo32, *mo16, *mr, p, grid, s);
    } else {
        if (a.nin.on) launch_halo_bn<OP_F16N, 1>(pl.bn, pl.cs, ma, mb, *mo32, *mo16, *mr, p, grid, s);
        else launch_halo_bn<OP_F16N, 0>(pl.bn, pl.cs, ma, mb, *mo32, *mo16, *mr, p, grid, s);
    }
    static const bool dbg_all = dbg_env && !strcmp(getenv("THA4_HALO_DEBUG"), "2");
    if (dbg_all) {       // developer: stamps of every launch of a real forward (serialises the stream)
        fprintf(stderr, "halo launch: N %d %dx%d cin %d cout %d | bn %d cs %d chunks %d grid %d x %d | xf %d groups %d act %d res %d out32 %d out16 %d st_tma %d\n",
                p.N, p.MH, p.MW, cw.cin, cw.cout, pl.bn, pl.cs, pl.chunks, pl.tiles_m, pl.tiles_n, a.nin.on ? 1 : 0, p.xf_groups, p.xf_act,
                p.res_mode, p.out ? 1 : 0, p.out16 ? 1 : 0, p.st_tma);
        conv_halo_debug_dump();
    }
}

// developer helper: prints the phase stamps of the last launch (THA4_HALO_DEBUG=1)
void conv_halo_debug_dump() {
    if (!g_dbg_buf) return;
    std::vector<long long> h(32 * 32);
    THA4_CUDA_CHECK(cudaDeviceSynchronize());
    THA4_CUDA_CHECK(cudaMemcpy(h.data(), g_dbg_buf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    for (int c = 0; c < 32; ++c) {
        const long long* d = h.data() + c * 32;
        if (!d[0]) continue;
        auto rel = [&](long long v) { return v ? (long)(v - d[0]) : -1L; };
        fprintf(stderr, "halo slot %2d: tmem %ld pdl %ld | coef %ld a_full %ld xf_done %ld | mma_first %ld mma_commit %ld | t_full %ld epi_done %ld | bar_A %ld pushed+bar_B %ld (loads issued %ld stored %ld) reduce_done %ld | end %ld\n",
                c, rel(d[1]), rel(d[2]), rel(d[16]), rel(d[17]), rel(d[18]), rel(d[8]), rel(d[9]), rel(d[19]), rel(d[20]), rel(d[21]), rel(d[23]), rel(d[25]), rel(d[24]), rel(d[22]), rel(d[3]));
    }
}

}  // namespace tha4
