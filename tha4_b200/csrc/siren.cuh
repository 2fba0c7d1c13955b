// SIREN student networks (declarations) -- see siren.cu.
#pragma once
#include "nets.cuh"

namespace tha4 {

struct SirenLayer {
    void* W = nullptr;        // __half [NPAD][KPAD], pre-scaled by omega_0
    float* bias = nullptr;    // [NPAD], pre-scaled
    float* wxy = nullptr;     // [NPAD][2]   (first layers only)
    float* wpose = nullptr;   // [NPAD][P]   (first layers only)
    int N = 0, NPAD = 0, KPAD = 0, P = 0;
    void load(const StateDict& sd, const std::string& prefix, int feat, int pose, int kpad, int npad, float scale, cudaStream_t s);
};

class SirenFaceNet {
public:
    void load(const StateDict& sd, cudaStream_t s);
    // pose: [B, >=39] with row stride pose_ld; out: [B,4,128,128] fp32 NCHW
    void forward(Runtime& rt, const float* pose, int pose_ld, int B, float* out);
    bool loaded() const { return loaded_; }
private:
    AllocSink owned_;
    SirenLayer layers_[8], head_;
    bool loaded_ = false;
};

class SirenBodyNet {
public:
    void load(const StateDict& sd, cudaStream_t s);
    // image: [B,4,512,512]; pose: [B,45]; outputs: blended(4) alpha(1) colour(4) warped(4) grid_change(2), fp32 NCHW
    void forward(Runtime& rt, const ImgView& image, const float* pose, int pose_ld, float* const* outputs);
    bool loaded() const { return loaded_; }
private:
    AllocSink owned_;
    SirenLayer l_[3][3], head_;
    bool loaded_ = false;
};

}  // namespace tha4
