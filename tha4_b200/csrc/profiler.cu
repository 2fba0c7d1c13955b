#include "profiler.cuh"
#include <vector>

namespace tha4 {
namespace {
struct Span { cudaEvent_t a, b; int cat; };
bool g_on = false;
std::vector<Span> g_spans;
std::vector<cudaEvent_t> g_free;
cudaEvent_t g_open[PROF_NCAT];
double g_us[PROF_NCAT], g_flops[PROF_NCAT], g_bytes[PROF_NCAT];
long g_launches[PROF_NCAT];

cudaEvent_t get_event() {
    if (!g_free.empty()) { cudaEvent_t e = g_free.back(); g_free.pop_back(); return e; }
    cudaEvent_t e;
    THA4_CUDA_CHECK(cudaEventCreate(&e));
    return e;
}
void drain() {
    for (auto& sp : g_spans) {
        THA4_CUDA_CHECK(cudaEventSynchronize(sp.b));
        float ms = 0.f;
        THA4_CUDA_CHECK(cudaEventElapsedTime(&ms, sp.a, sp.b));
        g_us[sp.cat] += 1000.0 * ms;
        g_launches[sp.cat] += 1;
        g_free.push_back(sp.a); g_free.push_back(sp.b);
    }
    g_spans.clear();
}
}  // namespace

void prof_enable(bool on) { g_on = on; }
bool prof_enabled() { return g_on; }
void prof_begin(int cat, cudaStream_t s) {
    g_open[cat] = get_event();
    THA4_CUDA_CHECK(cudaEventRecord(g_open[cat], s));
}
void prof_end(int cat, cudaStream_t s) {
    cudaEvent_t b = get_event();
    THA4_CUDA_CHECK(cudaEventRecord(b, s));
    g_spans.push_back({g_open[cat], b, cat});
}
void prof_add_work(int cat, double flops, double bytes) { if (g_on) { g_flops[cat] += flops; g_bytes[cat] += bytes; } }
double prof_read(int cat, int what) {
    drain();
    if (cat < 0 || cat >= PROF_NCAT) return -1;
    switch (what) { case 0: return g_us[cat]; case 1: return (double)g_launches[cat]; case 2: return g_flops[cat]; default: return g_bytes[cat]; }
}
void prof_reset() {
    drain();
    for (int i = 0; i < PROF_NCAT; ++i) { g_us[i] = g_flops[i] = g_bytes[i] = 0; g_launches[i] = 0; }
}
}  // namespace tha4
