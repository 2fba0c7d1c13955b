// Network assembly for the THA4 hot path: weights in library-owned packed layouts, activations from a caching
// device pool, launch sequences written as plain C++ that reads like the reference forward()s.
#pragma once
#include "common.cuh"
#include "conv.cuh"
#include "ops.cuh"
#include <map>
#include <unordered_map>
#include <memory>
#include <string>
#include <vector>

namespace tha4 {

// Exact-size caching device allocator.  A forward pass requests the same sizes in the same order every time, so
// after the first call no cudaMalloc happens and every buffer keeps its address (CUDA-graph friendly).
class Pool {
public:
    ~Pool();
    float* alloc(size_t nfloats);
    double* alloc_f64(size_t n) { return reinterpret_cast<double*>(alloc(2 * n)); }
    void reset();          // every block becomes reusable (stream order makes reuse safe: one stream per ctx call)
    size_t bytes() const { return total_; }
private:
    struct Bucket { std::vector<void*> blocks; size_t next = 0; };
    std::unordered_map<size_t, Bucket> buckets_;   // O(1) per request: a pass makes ~600 of them between kernel launches
    std::vector<void*> all_;
    size_t total_ = 0;
};

struct TensorRef { const float* p = nullptr; std::vector<long> shape; long numel() const; };
using StateDict = std::map<std::string, TensorRef>;

struct NormW { float* gamma = nullptr; float* beta = nullptr; int C = 0; };

struct Runtime {                      // per-call execution context
    Pool* persist;
    Pool* scratch;
    cudaStream_t stream;
    int strict;
    int f16 = 0;                      // normalisation layers hand f16 tensors to the tcgen05 convs (non-strict, tcgen05 on)
    // optional second stream + fork / join events: independent branches of the DAG (the 1x1 skip conv of a ResBlock next to
    // its norm0 -> conv0 chain) run beside the main chain -- also inside a captured graph, where they become parallel branches
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // zero-initialised arena for per-(n,c) statistics (View::stats); bump-allocated, re-zeroed by the caller per pass
    double* stats_base = nullptr;
    size_t stats_cap = 0;
    size_t* stats_off = nullptr;
    double* alloc_stats(size_t n);
};

// ------------------------------------------------------------------ encoder-decoder networks
// EyebrowDecomposer00 / EyebrowMorphingCombiner00 / FaceMorpher08 (poser_encoder_decoder_00.py:43-121,
// face_morpher_08.py:48-202): conv3 + 3 stride-2 convs, bottleneck conv (pose concat) + 5 ResnetBlocks, 3 transposed
// convs, fused head tail.
class EncDecNet {
public:
    EncDecNet(TailKind kind, int size, int in_ch, int pose_ch);
    void load(const StateDict& sd, cudaStream_t s);
    // image0 / image1: see tail.cu (decomposer, face: image1 unused; combiner: image0 = eyebrow layer,
    // image1 = background layer, network input = cat(background, eyebrow)).
    void forward(Runtime& rt, const ImgView& image0, const ImgView& image1, const float* pose, int pose_ld,
                 float* const* outputs);
    int size() const { return S_; }
    int num_outputs() const { return kind_ == TAIL_DECOMPOSER ? 6 : 8; }
    bool loaded() const { return loaded_; }
private:
    void forward_fused(Runtime& rt, const View& x0, const ImgView& image0, const ImgView& image1, const float* pose, int pose_ld,
                       float* const* outputs);
    AllocSink owned_;          // every device allocation made by load()
    TailKind kind_;
    int S_, in_ch_, pose_ch_, pose_pad_;
    bool loaded_ = false;
    ConvWeights down_[4], bott0_, res_[5][2], up_[3];
    NormW down_n_[4], bott0_n_, res_n_[5][2], up_n_[3];
    TailWeights tail_;
};

// ------------------------------------------------------------------ U-Net networks
struct ResBlockW {
    int cin = 0, cout = 0;
    NormW norm0, norm1;
    ConvWeights conv0, conv1, skip;
    bool has_skip = false;
    float* film0 = nullptr;     // [2*cout], constant (t = 0 time embedding, unet.py:365-376)
    int film1_off = 0;          // offset of this block's 2*cout FiLM vector in the batched pose projection
};
struct AttnW { int C = 0; NormW norm; ConvWeights qkv, proj; };

// Morpher00 (morpher_00.py:35-72) and Upscaler02 (upscaler_02.py:37-102) on Unet / UnetWithFirstConvAddition
// (unet.py:438-546,549-658).
class UNetNet {
public:
    UNetNet(bool upscaler, int size, int model_channels, std::vector<int> mults);
    void load(const StateDict& sd, cudaStream_t s);
    // morpher: image = [B,4,S,S]; upscaler: image = rest image, half_posed / half_grid at S/2 (mode_07.py:111-118).
    void forward(Runtime& rt, const ImgView& image, const float* coarse_posed, const float* coarse_grid, int coarse_size,
                 const float* pose, int pose_ld, float* const* outputs);
    int size() const { return S_; }
    bool loaded() const { return loaded_; }
private:
    AllocSink owned_;          // every device allocation made by load()
    void forward_fused(Runtime& rt, const ImgView& image, const float* coarse_posed, const float* coarse_grid, int coarse_size,
                       const float* pose, int pose_ld, float* const* outputs);
    void res_block(Runtime& rt, const ResBlockW& w, const View& x, int mode, const float* film1, const View& out);
    void attn_block(Runtime& rt, const AttnW& w, const View& x, const View& out);
    bool upscaler_;
    int S_, mc_, L_;
    std::vector<int> mults_;
    bool loaded_ = false;
    ConvWeights first_;
    std::vector<ResBlockW> down_res_, down_ds_, mid_res_, up_res_, up_us_;   // up_res_: 2 per level
    std::vector<AttnW> mid_attn_, up_attn_;
    AttnW down_attn_;
    float *cond_w0_ = nullptr, *cond_b0_ = nullptr, *cond_w2_ = nullptr, *cond_b2_ = nullptr;
    float *film1_w_ = nullptr, *film1_b_ = nullptr;
    int film1_total_ = 0;
    NormW last_n_;
    TailWeights tail_;
};

}  // namespace tha4
