// Fused decoder tail on the Blackwell paths (default precision mode) -- SURVEY.md section 8 a-T, the kernel BASELINE.json's
// "grid_sample + decoder %HBM-roofline" names.  From the RAW last feature map of a network (f16, NHWC, written by the
// producing conv's epilogue) to every tensor the network returns, one pass, HBM-bound by design:
//
//   TMA     one 4-D box {C ch, 130 w, TR+2 h, 1 n} = the (TR+2) x 130 pixel HALO of a 128 x TR pixel tile lands in shared
//           memory as 128-byte (C = 64) / 64-byte (C = 32) rows with the matching swizzle; out-of-image pixels are
//           zero-filled = the head conv's zero padding.  The nine 16 x C head-weight tiles come in with one more box.
//   warps   apply the pending GroupNorm / InstanceNorm affine + SiLU / ReLU to the halo IN PLACE (the affine is rebuilt
//           per CTA from the statistics the producing conv accumulated: no finalize kernel, no coefficient tensor).
//   UMMA    the 3x3 head conv is 9 taps x C/16 tcgen05.mma (M = 128 pixels of one tile row, N = 16 head channels, K = 16)
//           per tile row, straight from the halo: the A descriptor of tap (dy, dx) is the SAME shared-memory image with
//           its start address shifted by (dy * 130 + dx) rows -- the tensor core applies the swizzle on absolute
//           addresses, so any row shift is legal (profiles/r02_umma_row_shift_probe.txt).  Accumulators: TR x 16 TMEM columns.
//   drain   tcgen05.ld gives every thread the 16 head outputs of ITS pixel (lane = pixel): no shared-memory transpose;
//           sigmoid / tanh, affine_grid + grid_sample (4-tap gather from the planar image), blends, and planar NCHW stores
//           where the 32 lanes of a warp write 32 consecutive pixels = one full 128-byte line per channel.
// Reference: morpher_00.py:53-66, upscaler_02.py:84-96, face_morpher_08.py:170-193, eyebrow_morphing_combiner_00.py:51-72,
// eyebrow_decomposer_00.py:49-64.  The fp32 / strict variant is tail.cu.
#include "ops.cuh"
#include "tail_epilogue.cuh"
#include "profiler.cuh"
#include "tc_common.cuh"
#include <cuda.h>
#include <cmath>
#include <map>
#include <mutex>
#include <tuple>

namespace tha4 {
namespace {

using namespace tc;

constexpr int TT_W = 128, TT_HW = TT_W + 2;          // tile / halo width in pixels
constexpr int TT_THREADS = 128;
constexpr int TT_N = 16;                             // head channels of the MMA (TAIL_CO_PAD = 12 real ones at most)

struct TailTcParams {
    int S, N;
    const double* stats; int stats_ld, stats_rep; long stats_rep_stride;   // statistics of the raw feature map
    int groups, act;
    const float* gamma; const float* beta;
    const float* bias;               // [TAIL_CO_PAD]
    float acc_scale;
    ImgView img0, img1;
    const float* g0; int g0_ld; const float* g1; int g1_ld;     // optional interleaved (NHWC) copies of img0 / img1 (tail_epilogue.cuh)
    const float* base;
    float* o[8];
};

template <int C> struct TailCfg {
    static constexpr int ROWB = 2 * C;                // bytes per pixel row of the operand (f16)
    static constexpr int TR = C == 32 ? 4 : 2;        // tile rows per CTA (shared-memory budget: 3 / 2 CTAs per SM)
    static constexpr int HROWS = (TR + 2) * TT_HW;    // halo pixels
    static constexpr int A_BYTES = ((HROWS * ROWB + 1023) / 1024) * 1024;
    static constexpr int B_BYTES = 9 * TT_N * ROWB;
    static constexpr int TMEM_COLS = TR * TT_N < 32 ? 32 : TR * TT_N;
    static constexpr size_t SMEM = 1024 + A_BYTES + B_BYTES + 2 * C * sizeof(float) + 2 * C * sizeof(double) + 16 * sizeof(float) + 64;
};

template <int KIND, int C>
__global__ void __launch_bounds__(TT_THREADS) tail_tc_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmW,
                                                              const TailTcParams p) {
    using Cfg = TailCfg<C>;
    constexpr int ROWB = Cfg::ROWB, TR = Cfg::TR, NCH = ROWB / 16;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic (not an integer round trip) keeps the shared address space: LDS / STS, not generic LD / ST
    uint8_t* smA = smem;
    uint8_t* smB = smem + Cfg::A_BYTES;
    double* chs = reinterpret_cast<double*>(smB + Cfg::B_BYTES);         // [C][2] folded statistics
    float* cA = reinterpret_cast<float*>(chs + 2 * C);                   // [C] affine
    float* cB = cA + C;
    float* sbias = cB + C;                                               // [16]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + 16);            // tma_full, mma_done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int n = blockIdx.z, x0 = blockIdx.x * TT_W, y0 = blockIdx.y * TR;

    if (tid == 0) {
        mbar_init(smem_u32(bars), 1); mbar_init(smem_u32(bars + 1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("prefetch.tensormap [%0];\n" :: "l"(&tmF) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" :: "l"(&tmW) : "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    if (tid == 0) {        // the head weights do not depend on the previous kernel: fetch them ahead of the dependency wait
        mbar_expect_tx(smem_u32(bars), Cfg::HROWS * ROWB + Cfg::B_BYTES);
        tma_load_3d(smem_u32(smB), &tmW, 0, 0, 0, smem_u32(bars));
    }
    if (tid < 16) sbias[tid] = tid < TAIL_CO_PAD ? __ldg(p.bias + tid) : 0.0f;
    pdl_wait();
    if (tid == 0) tma_load_4d(smem_u32(smA), &tmF, 0, x0 - 1, y0 - 1, n, smem_u32(bars));

    // ---- per-channel affine of the pending normalisation, from the producer's statistics ----
    const int cpg = p.groups == 0 ? 1 : C / p.groups;
    for (int c = tid; c < C; c += TT_THREADS) {
        const double2 v = fold_stat_replicas(p.stats + ((long)n * p.stats_ld + c) * 2, p.stats_rep_stride, p.stats_rep);
        chs[2 * c] = v.x; chs[2 * c + 1] = v.y;
    }
    __syncthreads();
    const bool silu = p.act == ACT_SILU || p.act == ACT_SILU_FAST;
    for (int c = tid; c < C; c += TT_THREADS) {
        const int g0 = (c / cpg) * cpg;
        double su = 0.0, sq = 0.0;
        for (int j = 0; j < cpg; ++j) { su += chs[2 * (g0 + j)]; sq += chs[2 * (g0 + j) + 1]; }
        const double inv_cnt = 1.0 / ((double)p.S * p.S * cpg);          // fp64 for the sums and the cancelling subtraction only
        const double mean = su * inv_cnt;
        const float var = fmaxf((float)fma(sq, inv_cnt, -mean * mean), 0.0f);
        float A = rsqrtf(var + 1e-5f) * __ldg(p.gamma + c);
        float B = __ldg(p.beta + c) - (float)mean * A;
        if (silu) { A *= 0.5f; B *= 0.5f; }                                // silu(v) = h + h * tanh(h), h = v / 2
        cA[c] = A; cB[c] = B;
    }
    __syncthreads();

    // ---- normalise + activate the halo in place (zero padding stays zero) ----
    mbar_wait(smem_u32(bars), 0);
    for (int row = tid; row < Cfg::HROWS; row += TT_THREADS) {
        const int hy = row / TT_HW, hx = row - hy * TT_HW;
        const int gy = y0 - 1 + hy, gx = x0 - 1 + hx;
        if (gy < 0 || gy >= p.S || gx < 0 || gx >= p.S) continue;
        uint8_t* rowp = smA + row * ROWB;
        const int swz = ROWB == 128 ? (row & 7) : ((row >> 1) & 3);        // smA is 1024-byte aligned: the row's XOR term of the swizzle
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            uint4* dp = reinterpret_cast<uint4*>(rowp + ((j ^ swz) << 4));
            uint4 d = *dp;
            const float4 a0 = *reinterpret_cast<const float4*>(cA + 8 * j), a1 = *reinterpret_cast<const float4*>(cA + 8 * j + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(cB + 8 * j), b1 = *reinterpret_cast<const float4*>(cB + 8 * j + 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            __half2* hp = reinterpret_cast<__half2*>(&d);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 v = __half22float2(hp[e]);
                v.x = fmaf(v.x, av[2 * e], bv[2 * e]); v.y = fmaf(v.y, av[2 * e + 1], bv[2 * e + 1]);
                if (silu) {
                    float tx, ty;
                    asm("tanh.approx.f32 %0, %1;\n" : "=f"(tx) : "f"(v.x));
                    asm("tanh.approx.f32 %0, %1;\n" : "=f"(ty) : "f"(v.y));
                    v.x = fmaf(v.x, tx, v.x); v.y = fmaf(v.y, ty, v.y);
                } else if (p.act == ACT_RELU) {
                    v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f);
                }
                hp[e] = __floats2half2_rn(v.x, v.y);
            }
            *dp = d;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");          // generic-proxy writes -> the tensor core's async-proxy reads
    __syncthreads();

    // ---- 3x3 head conv: one elected thread issues TR x 9 x C/16 MMAs on row-shifted views of the halo ----
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        constexpr uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TT_N >> 3) << 17) | ((128u >> 4) << 24);
#pragma unroll 1
        for (int r = 0; r < TR; ++r)
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
                const int dy = tap / 3, dx = tap - 3 * dy;
                const uint64_t adesc = make_smem_desc_sw<ROWB>(smem_u32(smA + ((r + dy) * TT_HW + dx) * ROWB));
                const uint64_t bdesc = make_smem_desc_sw<ROWB>(smem_u32(smB + tap * TT_N * ROWB));
#pragma unroll
                for (int k = 0; k < C / 16; ++k)
                    umma_f16(tmem_base + (uint32_t)(r * TT_N), adesc + 2 * k, bdesc + 2 * k, idesc, (tap > 0 || k > 0) ? 1u : 0u);
            }
        umma_commit(smem_u32(bars + 1));
    }
    mbar_wait(smem_u32(bars + 1), 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

    // ---- drain: thread = pixel column x0 + tid; its 16 head outputs per tile row come straight out of TMEM ----
    const int x = x0 + tid;
#pragma unroll 1
    for (int r0 = 0; r0 < TR; r0 += 2) {
        uint32_t acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(r0 * TT_N), acc);     // two tile rows x 16 columns
        if (x < p.S) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                float o[TAIL_CO_PAD];
#pragma unroll
                for (int j = 0; j < TAIL_CO_PAD; ++j) o[j] = fmaf(__uint_as_float(acc[rr * TT_N + j]), p.acc_scale, sbias[j]);
                tail_epilogue<KIND>(o, n, y0 + r0 + rr, x, p.S, p.img0, p.img1, p.base, p.o[0], p.o[1], p.o[2], p.o[3], p.o[4], p.o[5], p.o[6], p.o[7]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
}


// ---------------------------------------------------------------------------------------------------------------------
// PERSISTENT, software-pipelined version (the default).  The kernel above runs one tile per CTA and its phases -- statistics
// fold, halo TMA, in-place normalisation, MMAs, gather / blend / store drain -- are a serial chain, co-resident CTAs all start
// together, so nothing overlaps (ncu at B=1, Upscaler02 site: 46 us, every unit < 16 % busy).  Here ONE CTA per SM walks a
// contiguous range of tiles with the phases on different warps and different tiles:
//   warp 0      TMA producer: halo boxes into an NS-deep ring (the head weights once)
//   warp 1      MMA issuer: TR x 9 x C/16 tcgen05.mma per tile into one of TWO accumulator slots in TMEM
//   warps 2..   workers (4 * TR warps), two stages on the same warps, two tiles apart, the drain split around the transform:
//               drain A(i):       tcgen05.ld of the pixel's grid_change outputs -> gs_locate -> the four corner pixels requested;
//               transform(i + 2): normalise + activate the landed halo in place (the per-channel affine is rebuilt once per
//                                 SAMPLE, not per tile), every worker thread;
//               drain B(i):       tcgen05.ld again -> tail_epilogue with the corners that arrived meanwhile (thread = pixel).
// (profiles/r02_tail_persist_notes.txt: the seven versions on the way here and what the ncu source page showed for each)
// mbarriers: h_full (TMA -> transform), h_xf (transform -> MMA, one arrival per worker thread), h_empty (MMA commit -> TMA),
// acc_full (MMA commit -> drain), acc_empty (drain -> MMA, one arrival per drain warp once its tcgen05.ld has completed).
template <int C, int TR> struct TailPCfg {
    static constexpr int ROWB = 2 * C;
    static constexpr int HROWS = (TR + 2) * TT_HW;
    static constexpr int A_BYTES = ((HROWS * ROWB + 1023) / 1024) * 1024;
    static constexpr int NS = 3;                                            // halo ring depth (the transform runs two tiles ahead of the drain)
    static_assert(C == 32 || TR == 2, "64-channel sites: TR = 2 (three 66 KB halo slots fit, three 100 KB slots do not)");
    static constexpr int B_BYTES = 9 * TT_N * ROWB;
    static constexpr int ACC_COLS = TR * TT_N;                              // TMEM columns of one accumulator slot
    static constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS; // 64 / 128: a power of two
    static constexpr int DRAIN_WARPS = 4 * TR;                              // one thread per output pixel of a tile
    static constexpr int WORKER_WARPS = DRAIN_WARPS;                        // every worker warp transforms AND drains (18 / 10 warps per CTA: the
                                                                            // register cap stays above what the split drain keeps live)
    static constexpr int THREADS = (2 + WORKER_WARPS) * 32;
    static constexpr int NBARS = 3 * NS + 5;
    static constexpr int MAX_S = 512;                                       // base-grid copy in shared memory
    static constexpr size_t SMEM = 1024 + (size_t)NS * A_BYTES + B_BYTES + 2 * C * sizeof(double) + 2 * C * sizeof(float) + 16 * sizeof(float) +
                                   MAX_S * sizeof(float) + NBARS * sizeof(uint64_t) + 16;
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

template <int KIND, int C, int TR>
__global__ void __launch_bounds__(TailPCfg<C, TR>::THREADS, 1) tail_tc_persist_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmW,
                                                                                      const TailTcParams p) {
    using Cfg = TailPCfg<C, TR>;
    constexpr int ROWB = Cfg::ROWB, NCH = ROWB / 16, NS = Cfg::NS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smA = smem;
    uint8_t* smB = smem + NS * Cfg::A_BYTES;
    double* chs = reinterpret_cast<double*>(smB + Cfg::B_BYTES);         // [C][2] folded statistics
    float* cA = reinterpret_cast<float*>(chs + 2 * C);                   // [C] affine
    float* cB = cA + C;
    float* sbias = cB + C;                                               // [16]
    float* sbase = sbias + 16;                                           // [S] affine_grid coordinates (one dependent global round trip less per pixel)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sbase + Cfg::MAX_S);
    uint64_t* h_full = bars, *h_xf = bars + NS, *h_empty = bars + 2 * NS;
    uint64_t* acc_full = bars + 3 * NS, *acc_empty = acc_full + 2, *w_full = acc_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_x = (p.S + TT_W - 1) / TT_W, tiles_y = p.S / TR, per_n = tiles_x * tiles_y;
    const int total = per_n * p.N;
    const int t_begin = (int)((long)blockIdx.x * total / gridDim.x);
    const int nt = (int)((long)(blockIdx.x + 1) * total / gridDim.x) - t_begin;       // contiguous tiles of this CTA (>= 1: grid <= total)

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(smem_u32(h_full + s), 1); mbar_init(smem_u32(h_xf + s), Cfg::WORKER_WARPS * 32); mbar_init(smem_u32(h_empty + s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(acc_full + a), 1); mbar_init(smem_u32(acc_empty + a), Cfg::DRAIN_WARPS); }
        mbar_init(smem_u32(w_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("prefetch.tensormap [%0];\n" :: "l"(&tmF) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" :: "l"(&tmW) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid >= 64 && tid < 80) sbias[tid - 64] = (tid - 64) < TAIL_CO_PAD ? __ldg(p.bias + (tid - 64)) : 0.0f;     // weights: independent of the previous kernel
    for (int i = tid; i < p.S; i += Cfg::THREADS) sbase[i] = __ldg(p.base + i);                                     // constant table
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    if (tid == 0) {        // the head weights do not depend on the previous kernel: fetch them ahead of the dependency wait
        mbar_expect_tx(smem_u32(w_full), Cfg::B_BYTES);
        tma_load_3d(smem_u32(smB), &tmW, 0, 0, 0, smem_u32(w_full));
    }
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {   // ===== TMA producer =====
            for (int i = 0; i < nt; ++i) {
                const int s = i % NS;
                int t = t_begin + i;
                const int n = t / per_n; t -= n * per_n;
                const int ty = t / tiles_x, tx = t - ty * tiles_x;
                mbar_wait(smem_u32(h_empty + s), ((i / NS) & 1) ^ 1);
                mbar_expect_tx(smem_u32(h_full + s), Cfg::HROWS * ROWB);
                tma_load_4d(smem_u32(smA + s * Cfg::A_BYTES), &tmF, 0, tx * TT_W - 1, ty * TR - 1, n, smem_u32(h_full + s));
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ===== MMA issuer: TR tile rows x 9 taps = row-shifted views of the halo =====
            constexpr uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TT_N >> 3) << 17) | ((128u >> 4) << 24);
            mbar_wait(smem_u32(w_full), 0);
            for (int i = 0; i < nt; ++i) {
                const int s = i % NS, a = i & 1;
                mbar_wait(smem_u32(acc_empty + a), ((i >> 1) & 1) ^ 1);
                mbar_wait(smem_u32(h_xf + s), (i / NS) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint8_t* hA = smA + s * Cfg::A_BYTES;
#pragma unroll 1
                for (int r = 0; r < TR; ++r)
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        const int dy = tap / 3, dx = tap - 3 * dy;
                        const uint64_t adesc = make_smem_desc_sw<ROWB>(smem_u32(hA + ((r + dy) * TT_HW + dx) * ROWB));
                        const uint64_t bdesc = make_smem_desc_sw<ROWB>(smem_u32(smB + tap * TT_N * ROWB));
#pragma unroll
                        for (int k = 0; k < C / 16; ++k)
                            umma_f16(tmem_base + (uint32_t)(a * Cfg::ACC_COLS + r * TT_N), adesc + 2 * k, bdesc + 2 * k, idesc, (tap > 0 || k > 0) ? 1u : 0u);
                    }
                umma_commit(smem_u32(h_empty + s));          // the halo slot is free once these MMAs have read it
                umma_commit(smem_u32(acc_full + a));
            }
        }
    } else {
        // ===== worker warps (2 ..): BOTH remaining stages, on every warp =====
        //   transform(i + 1): normalise + activate the landed halo of the NEXT tile in place, all WORKERS threads;
        //   drain(i):         one thread per output pixel of the current tile (the first 4 * TR warps).
        // The first version gave the transform four dedicated warps: one warp per scheduler, 0.2 IPC on dependent
        // conversions / MUFU, 7 000 cycles per tile and the slowest stage of the pipeline while sixteen drain warps waited
        // (ncu source page, profiles/r02_tail_persist_notes.txt).  A thread owns ONE 16-byte chunk column (8 channels) of the
        // halo rows wt / NCH + k * (WORKERS / NCH): its 16 coefficients live in registers, XU rows are in flight at once.
        constexpr int WORKERS = Cfg::WORKER_WARPS * 32;
        constexpr int RSTEP = WORKERS / NCH;
        constexpr int ITEMS = (Cfg::HROWS + RSTEP - 1) / RSTEP, ITERS = (ITEMS + 5) / 6, XU = (ITEMS + ITERS - 1) / ITERS;
        const int wt = tid - 64, wid = warp - 2;
        const int jc = wt % NCH, r_first = wt / NCH;
        const int cpg = p.groups == 0 ? 1 : C / p.groups;
        const bool silu = p.act == ACT_SILU || p.act == ACT_SILU_FAST;
        const bool relu = p.act == ACT_RELU;
        const int q = warp & 3;                                                  // TMEM lane quadrant this warp may access
        const int dr = wid >> 2;                                                 // drain: tile row (each quadrant appears once per row)
        float av[8], bv[8];
        int cur_n = -1;

        auto transform = [&](int i) {
            const int s = i % NS;
            int t = t_begin + i;
            const int n = t / per_n; t -= n * per_n;
            const int ty = t / tiles_x, tx = t - ty * tiles_x;
            const int x0 = tx * TT_W, y0 = ty * TR;
            if (n != cur_n) {            // per-SAMPLE affine of the pending normalisation, from the producer's statistics
                cur_n = n;
                float g1 = 0.0f, b1 = 0.0f;
                if (wt < C) { g1 = __ldg(p.gamma + wt); b1 = __ldg(p.beta + wt); }     // requested ahead of the statistics
                asm volatile("bar.sync 1, %0;\n" :: "n"(WORKERS) : "memory");           // everyone is done with the previous table
                if (wt < C) {
                    const double2 v = fold_stat_replicas16(p.stats + ((long)n * p.stats_ld + wt) * 2, p.stats_rep_stride, p.stats_rep);
                    chs[2 * wt] = v.x; chs[2 * wt + 1] = v.y;
                }
                asm volatile("bar.sync 1, %0;\n" :: "n"(WORKERS) : "memory");
                if (wt < C) {
                    const int g0 = (wt / cpg) * cpg;
                    double su = 0.0, sq = 0.0;
                    for (int j = 0; j < cpg; ++j) { su += chs[2 * (g0 + j)]; sq += chs[2 * (g0 + j) + 1]; }
                    const double inv_cnt = 1.0 / ((double)p.S * p.S * cpg);          // fp64 for the sums and the cancelling subtraction only
                    const double mean = su * inv_cnt;
                    const float var = fmaxf((float)fma(sq, inv_cnt, -mean * mean), 0.0f);
                    float A = rsqrtf(var + 1e-5f) * g1;
                    float B = b1 - (float)mean * A;
                    if (silu) { A *= 0.5f; B *= 0.5f; }                            // silu(v) = h + h * tanh(h), h = v / 2
                    cA[wt] = A; cB[wt] = B;
                }
                asm volatile("bar.sync 1, %0;\n" :: "n"(WORKERS) : "memory");
#pragma unroll
                for (int e = 0; e < 8; ++e) { av[e] = cA[8 * jc + e]; bv[e] = cB[8 * jc + e]; }
            }
            mbar_wait(smem_u32(h_full + s), (i / NS) & 1);
            uint8_t* hbase = smA + s * Cfg::A_BYTES;
#pragma unroll 1
            for (int row0 = r_first; row0 < Cfg::HROWS; row0 += XU * RSTEP) {
                uint4 d[XU];
                uint4* dp[XU];
                bool ok[XU];
#pragma unroll
                for (int u = 0; u < XU; ++u) {
                    const int row = row0 + u * RSTEP;
                    const int hy = row / TT_HW, hx = row - hy * TT_HW;
                    const int gy = y0 - 1 + hy, gx = x0 - 1 + hx;
                    ok[u] = row < Cfg::HROWS && gy >= 0 && gy < p.S && gx >= 0 && gx < p.S;       // zero padding stays zero
                    const int swz = ROWB == 128 ? (row & 7) : ((row >> 1) & 3);
                    dp[u] = reinterpret_cast<uint4*>(hbase + row * ROWB + ((jc ^ swz) << 4));
                    if (ok[u]) d[u] = *dp[u];
                }
#pragma unroll
                for (int u = 0; u < XU; ++u) {
                    if (!ok[u]) continue;
                    __half2* hp = reinterpret_cast<__half2*>(&d[u]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float2 v = __half22float2(hp[e]);
                        v.x = fmaf(v.x, av[2 * e], bv[2 * e]); v.y = fmaf(v.y, av[2 * e + 1], bv[2 * e + 1]);
                        if (silu) {
                            float tx2, ty2;
                            asm("tanh.approx.f32 %0, %1;\n" : "=f"(tx2) : "f"(v.x));
                            asm("tanh.approx.f32 %0, %1;\n" : "=f"(ty2) : "f"(v.y));
                            v.x = fmaf(v.x, tx2, v.x); v.y = fmaf(v.y, ty2, v.y);
                        } else if (relu) {
                            v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f);
                        }
                        hp[e] = __floats2half2_rn(v.x, v.y);
                    }
                }
#pragma unroll
                for (int u = 0; u < XU; ++u)
                    if (ok[u]) *dp[u] = d[u];
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // generic-proxy writes -> the tensor core's async-proxy reads
            mbar_arrive(smem_u32(h_xf + s));
        };

        // The drain of a tile is split around the transform of the next one: phase A reads the two grid_change outputs of the
        // pixel and REQUESTS the four corner pixels of its sampling tap; the transform runs while they travel; phase B reads
        // the accumulator again (cheap: TMEM) and finishes.  Unsplit, every tile paid one exposed global round trip with all
        // sixteen drain warps waiting on it together (ncu source page: 60 % long-scoreboard in the drain, ~1.6 us per tile).
        // Every drain warp polls acc_full itself: a common named barrier made fifteen warps wait for the slowest (23 % of all
        // samples); with the drain behind the transform the accumulator has usually been complete for a while.
        float4 pre[4];
        bool have_pre = false;
        auto tile_xy = [&](int i, int& n, int& x, int& y) {
            int t = t_begin + i;
            n = t / per_n; t -= n * per_n;
            const int ty = t / tiles_x, tx = t - ty * tiles_x;
            x = tx * TT_W + q * 32 + lane; y = ty * TR + dr;
        };
        auto load_acc = [&](int a, float (&o)[TAIL_CO_PAD]) {
            uint32_t acc[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * Cfg::ACC_COLS + dr * TT_N), acc);
#pragma unroll
            for (int j = 0; j < TAIL_CO_PAD; ++j) o[j] = fmaf(__uint_as_float(acc[j]), p.acc_scale, sbias[j]);
        };
        auto drain_issue = [&](int i) {
            const int a = i & 1;
            int n, x, y;
            tile_xy(i, n, x, y);
            mbar_wait(smem_u32(acc_full + a), (i >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            have_pre = false;
            if (KIND != TAIL_DECOMPOSER && p.g0 != nullptr) {           // warp-uniform
                float o[TAIL_CO_PAD];
                load_acc(a, o);
                if (x < p.S) have_pre = tail_gather_issue<KIND>(o, n, y, x, p.S, sbase, p.g0, p.g0_ld, pre);
            }
        };
        auto drain_finish = [&](int i) {
            const int a = i & 1;
            int n, x, y;
            tile_xy(i, n, x, y);
            float o[TAIL_CO_PAD];
            load_acc(a, o);
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            if (lane == 0) mbar_arrive(smem_u32(acc_empty + a));                 // the accumulator slot may be overwritten
            if (x < p.S)
                tail_epilogue<KIND>(o, n, y, x, p.S, p.img0, p.img1, sbase, p.o[0], p.o[1], p.o[2], p.o[3], p.o[4], p.o[5], p.o[6], p.o[7],
                                    p.g0, p.g0_ld, p.g1, p.g1_ld, have_pre ? &pre : nullptr);
        };

        // The transform runs TWO tiles ahead of the drain (the halo ring is three deep): the MMAs of tile i + 1 then have the
        // whole of drain(i) and transform(i + 2) to complete in.  One tile ahead, with the drain split around the transform,
        // they had only the second half of drain(i): 31 % of all samples were drain warps polling acc_full (ncu, v5).
        transform(0);
        if (nt > 1) transform(1);
        for (int i = 0; i < nt; ++i) {
            drain_issue(i);
            if (i + 2 < nt) transform(i + 2);
            drain_finish(i);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tail_encode() {
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

using TKey = std::tuple<int, const void*, long, long, long, long, int>;
std::map<TKey, CUtensorMap> g_tail_maps;
std::mutex g_tail_maps_mu;

const CUtensorMap& feature_map(const View& f, int TR) {
    TKey key{current_device(), f.p, f.N, f.H, f.C, f.ld, TR};
    std::lock_guard<std::mutex> lock(g_tail_maps_mu);
    auto it = g_tail_maps.find(key);
    if (it != g_tail_maps.end()) return it->second;
    CUtensorMap m;
    cuuint64_t dims[4] = {(cuuint64_t)f.C, (cuuint64_t)f.W, (cuuint64_t)f.H, (cuuint64_t)f.N};
    cuuint64_t strides[3] = {(cuuint64_t)f.ld * 2, (cuuint64_t)f.W * f.ld * 2, (cuuint64_t)f.H * f.W * f.ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)f.C, (cuuint32_t)TT_HW, (cuuint32_t)(TR + 2), 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = tail_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, f.p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               f.C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    THA4_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(tail feature map) failed: " + std::to_string((int)r));
    return g_tail_maps.emplace(key, m).first->second;
}

const CUtensorMap& head_weight_map(const TailWeights& tw) {
    TKey key{current_device(), tw.w16, tw.C, 0, 0, 0, -1};
    std::lock_guard<std::mutex> lock(g_tail_maps_mu);
    auto it = g_tail_maps.find(key);
    if (it != g_tail_maps.end()) return it->second;
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)tw.C, (cuuint64_t)TT_N, 9};
    cuuint64_t strides[2] = {(cuuint64_t)tw.C * 2, (cuuint64_t)TT_N * tw.C * 2};
    cuuint32_t box[3] = {(cuuint32_t)tw.C, (cuuint32_t)TT_N, 9};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = tail_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, tw.w16, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               tw.C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    THA4_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(tail head weights) failed: " + std::to_string((int)r));
    return g_tail_maps.emplace(key, m).first->second;
}

// [9][C][TAIL_CO_PAD] fp32 -> [9][16][C] f16 (K-major B operand: one row per head channel), scaled by a power of two
__global__ void tail_pack_half_kernel(const float* __restrict__ w, __half* __restrict__ h, int C, float scale) {
    const int total = 9 * TT_N * C;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % C, nn = (i / C) % TT_N, tap = i / (C * TT_N);
        h[i] = __float2half_rn(nn < TAIL_CO_PAD ? w[(tap * C + c) * TAIL_CO_PAD + nn] * scale : 0.0f);
    }
}
__global__ void tail_absmax_kernel(const float* __restrict__ w, int n, unsigned* __restrict__ out) {
    float m = 0.0f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(w[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

template <int KIND, int C>
void launch_tail_tc(const TailWeights& tw, const View& f, const NormSpecTail& ns, const ImgView& i0, const ImgView& i1, float* const* o, int nout,
                    cudaStream_t s) {
    using Cfg = TailCfg<C>;
    TailTcParams p{};
    p.S = f.H; p.N = f.N;
    p.stats = f.stats; p.stats_ld = f.stats_ld; p.stats_rep = std::max(1, f.stats_rep); p.stats_rep_stride = f.stats_rep_stride;
    p.groups = ns.groups; p.act = ns.act; p.gamma = ns.gamma; p.beta = ns.beta;
    p.bias = tw.bias; p.acc_scale = 1.0f / tw.w16_scale;
    p.img0 = i0; p.img1 = i1; p.base = base_grid_table(f.H);
    for (int i = 0; i < 8; ++i) p.o[i] = i < nout ? o[i] : nullptr;
    THA4_ENSURE_SMEM((tail_tc_kernel<KIND, C>), Cfg::SMEM);
    dim3 grid(ceil_div(f.W, TT_W), f.H / Cfg::TR, f.N);
    ProfScope prof(PROF_TAIL, s);
    {   // compulsory traffic as SURVEY 8d defines it (fp32 element size): feature map + image(s) read once, every returned tensor written once
        const int out_ch[4] = {15, 18, 24, 24};
        const int img_ch = (KIND == TAIL_COMBINER) ? 8 : 4;
        prof_add_work(PROF_TAIL, 2.0 * f.pixels() * 9 * tw.C * tw.CO, (double)f.pixels() * (f.C + img_ch + out_ch[KIND]) * 4);
    }
    launch_pdl(tail_tc_kernel<KIND, C>, grid, dim3(TT_THREADS), Cfg::SMEM, s, 1, feature_map(f, Cfg::TR), head_weight_map(tw), p);
    THA4_LAUNCH_CHECK();
}

bool g_tail_persist = true;       // option "tail_persist": the persistent pipelined kernel (default) / one tile per CTA

int tail_num_sms() {
    static int sms[THA4_MAX_DEVICES] = {};
    const int d = current_device();
    if (!sms[d]) THA4_CUDA_CHECK(cudaDeviceGetAttribute(&sms[d], cudaDevAttrMultiProcessorCount, d));
    return sms[d];
}

template <int KIND, int C, int TR>
void launch_tail_persist(const TailWeights& tw, const View& f, const NormSpecTail& ns, const ImgView& i0, const ImgView& i1, float* const* o, int nout,
                         cudaStream_t s, const View* g0, const View* g1) {
    using Cfg = TailPCfg<C, TR>;
    TailTcParams p{};
    p.S = f.H; p.N = f.N;
    p.stats = f.stats; p.stats_ld = f.stats_ld; p.stats_rep = std::max(1, f.stats_rep); p.stats_rep_stride = f.stats_rep_stride;
    p.groups = ns.groups; p.act = ns.act; p.gamma = ns.gamma; p.beta = ns.beta;
    p.bias = tw.bias; p.acc_scale = 1.0f / tw.w16_scale;
    p.img0 = i0; p.img1 = i1; p.base = base_grid_table(f.H);
    auto gather_ok = [&](const View* g) {      // fp32 NHWC view of the same geometry, four channels 16-byte aligned, offsets within 32 bits
        return g && g->p && !g->f16 && g->N == f.N && g->H == f.H && g->W == f.W && g->ld % 4 == 0 && (reinterpret_cast<uintptr_t>(g->p) & 15) == 0 &&
               (size_t)f.H * f.W * g->ld < (size_t)1 << 30;
    };
    if (gather_ok(g0)) { p.g0 = g0->p; p.g0_ld = g0->ld; }
    if (gather_ok(g1)) { p.g1 = g1->p; p.g1_ld = g1->ld; }
    for (int i = 0; i < 8; ++i) p.o[i] = i < nout ? o[i] : nullptr;
    THA4_REQUIRE(f.H <= Cfg::MAX_S, "tail_tc: image size");
    THA4_ENSURE_SMEM((tail_tc_persist_kernel<KIND, C, TR>), Cfg::SMEM);
    const long total = (long)ceil_div(f.W, TT_W) * (f.H / TR) * f.N;
    dim3 grid((unsigned)std::min<long>(total, tail_num_sms()));
    ProfScope prof(PROF_TAIL, s);
    {   // compulsory traffic as SURVEY 8d defines it (fp32 element size): feature map + image(s) read once, every returned tensor written once
        const int out_ch[4] = {15, 18, 24, 24};
        const int img_ch = (KIND == TAIL_COMBINER) ? 8 : 4;
        prof_add_work(PROF_TAIL, 2.0 * f.pixels() * 9 * tw.C * tw.CO, (double)f.pixels() * (f.C + img_ch + out_ch[KIND]) * 4);
    }
    launch_pdl(tail_tc_persist_kernel<KIND, C, TR>, grid, dim3(Cfg::THREADS), Cfg::SMEM, s, 1, feature_map(f, TR), head_weight_map(tw), p);
    THA4_LAUNCH_CHECK();
}

template <int KIND>
void launch_tail_tc_c(const TailWeights& tw, const View& f, const NormSpecTail& ns, const ImgView& i0, const ImgView& i1, float* const* o, int nout,
                      cudaStream_t s, const View* g0, const View* g1) {
    if (g_tail_persist) {
        // tile rows per step (32-channel sites): 4 when that still gives every SM two tiles or more, else 2 (more, smaller tiles:
        // the small sites at B = 1 are one latency chain per CTA)
        const long tiles4 = (long)ceil_div(f.W, TT_W) * (f.H / 4) * f.N;
        const bool tr4 = tiles4 >= 2L * tail_num_sms();
        if (tw.C == 32) { if (tr4) launch_tail_persist<KIND, 32, 4>(tw, f, ns, i0, i1, o, nout, s, g0, g1); else launch_tail_persist<KIND, 32, 2>(tw, f, ns, i0, i1, o, nout, s, g0, g1); }
        else            launch_tail_persist<KIND, 64, 2>(tw, f, ns, i0, i1, o, nout, s, g0, g1);        // three halo slots of 66 KB
        return;
    }
    if (tw.C == 32) launch_tail_tc<KIND, 32>(tw, f, ns, i0, i1, o, nout, s);
    else launch_tail_tc<KIND, 64>(tw, f, ns, i0, i1, o, nout, s);
}

}  // namespace

void tail_make_half(TailWeights& tw, cudaStream_t s) {
    if (tw.w16 || !tw.w) return;
    const int nw = 9 * tw.C * TAIL_CO_PAD;
    __half* h = reinterpret_cast<__half*>(tracked_malloc((size_t)9 * TT_N * tw.C * sizeof(__half)));
    unsigned* dmax = reinterpret_cast<unsigned*>(h);
    THA4_CUDA_CHECK(cudaMemsetAsync(dmax, 0, sizeof(unsigned), s));
    tail_absmax_kernel<<<8, 256, 0, s>>>(tw.w, nw, dmax);
    THA4_LAUNCH_CHECK();
    unsigned hmax = 0;
    THA4_CUDA_CHECK(cudaMemcpyAsync(&hmax, dmax, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    THA4_CUDA_CHECK(cudaStreamSynchronize(s));
    float mx; memcpy(&mx, &hmax, sizeof(float));
    float scale = 1.0f;
    if (mx > 0.0f && std::isfinite(mx)) {
        int e = 0; frexpf(mx, &e);
        e = std::max(-24, std::min(8, e));
        scale = ldexpf(1.0f, -e);
    }
    tail_pack_half_kernel<<<16, 256, 0, s>>>(tw.w, h, tw.C, scale);
    THA4_LAUNCH_CHECK();
    tw.w16 = h; tw.w16_scale = scale;
}

void tail_tc_enable_persist(bool on) { g_tail_persist = on; }

bool tail_tc_supported(const TailWeights& tw, const View& feature) {
    return feature.f16 && feature.stats != nullptr && (tw.C == 32 || tw.C == 64) && feature.C == tw.C && feature.ld == tw.C &&
           feature.H == feature.W && feature.H % 4 == 0 && tw.w16 != nullptr && (((uintptr_t)feature.p) & 15) == 0;
}

void tail_tc_forward(TailKind kind, const TailWeights& tw, const View& feature, const NormSpecTail& ns, const ImgView& image0,
                     const ImgView& image1, float* const* outputs, cudaStream_t s, const View* g0, const View* g1) {
    THA4_REQUIRE(tail_tc_supported(tw, feature), "tail_tc: unsupported configuration");
    THA4_REQUIRE(image0.H == feature.H && image0.W == feature.W && image0.C == 4, "tail_tc: image dims");
    THA4_REQUIRE(ns.groups == 0 || tw.C % ns.groups == 0, "tail_tc: groups");
    switch (kind) {
        case TAIL_UNET: launch_tail_tc_c<TAIL_UNET>(tw, feature, ns, image0, image1, outputs, 5, s, g0, g1); break;
        case TAIL_DECOMPOSER: launch_tail_tc_c<TAIL_DECOMPOSER>(tw, feature, ns, image0, image1, outputs, 6, s, g0, g1); break;
        case TAIL_COMBINER: launch_tail_tc_c<TAIL_COMBINER>(tw, feature, ns, image0, image1, outputs, 8, s, g0, g1); break;
        case TAIL_FACE: launch_tail_tc_c<TAIL_FACE>(tw, feature, ns, image0, image1, outputs, 8, s, g0, g1); break;
    }
}

}  // namespace tha4
