// Hardware probe (developer tool, not product): how does tcgen05.mma address a K-major SWIZZLE_128B operand whose
// start address is NOT aligned to the 1024-byte swizzle atom (row-shifted descriptors), with and without the
// descriptor's base_offset field, and with a non-1024 stride between 8-row groups (SBO)?
//
// A is R rows x 64 f16 (128-byte rows) written with the absolute-address swizzle TMA would apply
// (16-byte chunk index XOR ((smem_address >> 7) & 7)).  B selects channels: B[n][k] = (k == n), n < 16, so that one
// K=16 MMA returns D[m][n] = A[row(m)][n].  A[row][ch] = row % 251 for even ch, ch for odd ch: the result tells which
// smem row and which 16-byte chunk the tensor core fetched for every logical (m, n).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma_shift_test scripts/dev/umma_shift_test.cu
#include "../../tha4_b200/csrc/tc_common.cuh"
#include <cstdio>
#include <vector>

namespace tha4 { std::atomic<long> g_kernel_launches{0}; bool g_use_pdl = false; thread_local AllocSink* g_alloc_sink = nullptr;
void* tracked_malloc(size_t) { return nullptr; } }
using namespace tha4;
using namespace tha4::tc;

constexpr int R = 200;

template <int ROWB>
__global__ void __launch_bounds__(128) probe(int shift_rows, int sbo_bytes, int use_base_offset, int kstep, float* out) {
    constexpr int NCH = ROWB / 16, CH = ROWB / 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smA = smem;                         // R rows x 128 B
    uint8_t* smB = smem + 26 * 1024;             // 16 rows x 128 B, 1024-aligned
    uint64_t* bar = reinterpret_cast<uint64_t*>(smB + 2048);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    // fill A with the absolute-address swizzle
    for (int i = tid; i < R * NCH; i += 128) {
        const int row = i / NCH, chunk = i % NCH;
        __half v[8];
        for (int e = 0; e < 8; ++e) { const int ch = chunk * 8 + e; v[e] = __float2half((ch & 1) ? (float)ch : (float)(row % 251)); }
        const uint32_t row_addr = smem_u32(smA) + row * ROWB;
        const uint32_t phys = (uint32_t)chunk ^ (ROWB == 128 ? ((row_addr >> 7) & 7) : ((row_addr >> 7) & 3));
        *reinterpret_cast<uint4*>(smA + row * ROWB + phys * 16) = *reinterpret_cast<uint4*>(v);
    }
    for (int i = tid; i < 16 * NCH; i += 128) {
        const int row = i / NCH, chunk = i % NCH;
        __half v[8];
        for (int e = 0; e < 8; ++e) { const int k = chunk * 8 + e; v[e] = __float2half(k == row + 16 * kstep ? 1.0f : 0.0f); }
        const uint32_t row_addr = smem_u32(smB) + row * ROWB;
        const uint32_t phys = (uint32_t)chunk ^ (ROWB == 128 ? ((row_addr >> 7) & 7) : ((row_addr >> 7) & 3));
        *reinterpret_cast<uint4*>(smB + row * ROWB + phys * 16) = *reinterpret_cast<uint4*>(v);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // generic-proxy writes -> async proxy (UMMA) reads
    if (tid == 0) { mbar_init(smem_u32(bar), 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(slot)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const uint32_t a_addr = smem_u32(smA) + shift_rows * ROWB + kstep * 32;
        const uint32_t b_addr = smem_u32(smB) + kstep * 32;
        uint32_t lo = ((a_addr & 0x3FFFF) >> 4) | (1u << 16);
        uint32_t hi = ((uint32_t)sbo_bytes >> 4) | (1u << 14) | ((ROWB == 128 ? 2u : 4u) << 29);
        if (use_base_offset) hi |= (((a_addr >> 7) & 7u) << 17);          // descriptor bits 49..51
        const uint64_t adesc = ((uint64_t)hi << 32) | lo;
        const uint64_t bdesc = make_smem_desc_sw<ROWB>(b_addr);
        constexpr uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(16 >> 3) << 17) | ((128u >> 4) << 24);
        umma_f16(tmem, adesc, bdesc, idesc, 0u);
        umma_commit(smem_u32(bar));
    }
    mbar_wait(smem_u32(bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), r);
    for (int j = 0; j < 16; ++j) out[tid * 16 + j] = __uint_as_float(r[j]);
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(32) : "memory");
}

template <int ROWB>
int run_all(float* d) {
    const size_t smem = 1024 + 26 * 1024 + 2048 + 64;
    cudaFuncSetAttribute(probe<ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int shifts[] = {0, 1, 2, 3, 8, 10, 19, 37};
    const int sbos[] = {8 * ROWB, 10 * ROWB};
    int all_ok_abs_nobo = 1, all_ok_abs_bo = 1;
    printf("==== %d-byte operand rows (%s) ====\n", ROWB, ROWB == 128 ? "SWIZZLE_128B" : "SWIZZLE_64B");
    for (int kstep = 0; kstep < ROWB / 32; kstep += ROWB / 32 - 1)
    for (int sbo : sbos)
        for (int shift : shifts)
            for (int bo = 0; bo < 2; ++bo) {
                cudaMemset(d, 0xff, 128 * 16 * sizeof(float));
                probe<ROWB><<<1, 128, smem>>>(shift, sbo, bo, kstep, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("kstep %d sbo %d shift %d bo %d: CUDA error %s\n", kstep, sbo, shift, bo, cudaGetErrorString(e)); return 1; }
                std::vector<float> h(128 * 16);
                cudaMemcpy(h.data(), d, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
                // hypothesis: row(m) = shift + (m / 8) * (sbo / 128) + m % 8, chunk un-swizzled by absolute address
                int bad = 0, first_m = -1; float got_row = 0, got_ch = 0;
                for (int m = 0; m < 128; ++m) {
                    const int row = shift + (m / 8) * (sbo / ROWB) + m % 8;
                    for (int n = 0; n < 16; ++n) {
                        const int ch = n + 16 * kstep;
                        const float want = (ch & 1) ? (float)ch : (float)(row % 251);
                        if (h[m * 16 + n] != want) { if (!bad) { first_m = m; got_row = h[m * 16 + 0]; got_ch = h[m * 16 + 1]; } ++bad; }
                    }
                }
                printf("kstep %d sbo %4d shift %2d base_offset %d : %s (%d mismatches", kstep, sbo, shift, bo, bad ? "MISMATCH" : "ok", bad);
                if (bad) printf("; first at m=%d: fetched row-code %.0f chunk-code %.0f, wanted row %d ch %d", first_m, got_row, got_ch,
                                (shift + (first_m / 8) * (sbo / ROWB) + first_m % 8) % 251, 1 + 16 * kstep);
                printf(")\n");
                if (bad) { if (bo) all_ok_abs_bo = 0; else all_ok_abs_nobo = 0; }
            }
    printf("SUMMARY (%d-byte rows) absolute-address swizzle holds for every shift: without base_offset %d, with base_offset %d\n", ROWB, all_ok_abs_nobo, all_ok_abs_bo);
    return 0;
}

int main() {
    float* d = nullptr;
    cudaMalloc(&d, 128 * 16 * sizeof(float));
    if (run_all<128>(d)) return 1;
    if (run_all<64>(d)) return 1;
    return 0;
}
