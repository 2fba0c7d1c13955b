// Optional per-kernel-class device timing (CUDA events on the launching stream).  Off by default; bench.py turns it
// on for a separate profiled pass to obtain the roofline numerators/denominators of the dominant kernels.
#pragma once
#include "common.cuh"

namespace tha4 {

enum ProfCat { PROF_CONV = 0, PROF_NORM = 1, PROF_TAIL = 2, PROF_ATTN = 3, PROF_GLUE = 4, PROF_SIREN = 5, PROF_NCAT = 6 };

void prof_enable(bool on);
bool prof_enabled();
void prof_begin(int cat, cudaStream_t s);
void prof_end(int cat, cudaStream_t s);
void prof_add_work(int cat, double flops, double bytes);
// Synchronises outstanding events and returns accumulated microseconds / launches / flops / bytes; `reset` clears.
double prof_read(int cat, int what /*0 us, 1 launches, 2 flops, 3 bytes*/);
void prof_reset();

struct ProfScope {
    int cat; cudaStream_t s; bool on;
    ProfScope(int c, cudaStream_t st) : cat(c), s(st), on(prof_enabled()) { if (on) prof_begin(cat, s); }
    ~ProfScope() { if (on) prof_end(cat, s); }
};

}  // namespace tha4
