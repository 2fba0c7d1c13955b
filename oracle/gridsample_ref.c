/* CPU restatement of the index arithmetic on the THA4 hot path  --  TEST INFRASTRUCTURE (oracle), never shipped.
 *
 * Restates, in plain C compiled with -ffp-contract=off:
 *   - affine_grid(identity, align_corners=False) base coordinates
 *       reference call sites: nn/image_processing_util.py:17-22,50; torch: linspace(-1,1,W)*(W-1)/W
 *   - grid_sample(bilinear, padding_mode=border, align_corners=False) of `image` at base + grid_change
 *       reference: nn/image_processing_util.py:13-24,33-54; nn/face_morpher/face_morpher_08.py:142-153
 *       arithmetic: ATen GridSampler.h grid_sampler_unnormalize / clip_coordinates (torch 2.11 headers :27-35,58-60)
 *   - interpolate(mode='bilinear', align_corners=False)
 *       reference: poser/modes/mode_07.py:102,114-115; nn/siren/morpher/siren_morpher_03.py:121
 * The integer corner indices (and the fp32 lerp weights) computed here are the bit-exact contract that the CUDA
 * kernels in tha4_b200/csrc must reproduce; the sampled values are tolerance-checked against torch.
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see __graft_entry__.build()).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

/* linspace(-1, 1, W) as torch computes it (step = 2/(W-1); symmetric halves), then * (W-1) / W. */
void tha4o_base_grid(int W, float* out) {
    float step = (1.0f - (-1.0f)) / (float)(W - 1);
    int half = W / 2;
    for (int i = 0; i < W; ++i) {
        float v = (i < half) ? (-1.0f + step * (float)i) : (1.0f - step * (float)(W - 1 - i));
        v = v * (float)(W - 1);
        out[i] = v / (float)W;
    }
}

/* source index for one axis: ((g + 1) * size - 1) / 2, clamped to [0, size-1] (border padding). */
static float src_index(float g, int size) {
    float v = ((g + 1.0f) * (float)size - 1.0f) / 2.0f;
    v = fminf((float)(size - 1), fmaxf(v, 0.0f));
    return v;
}

/* image [n,c,h,w], grid_change [n,2,h,w] (channel 0 = x), out [n,c,h,w]; x0,y0 [n,h,w] int32; tx,ty [n,h,w] fp32
 * (tx = ix - x0). Any of x0/y0/tx/ty/out may be NULL. */
void tha4o_grid_sample(const float* image, const float* grid_change, int n, int c, int h, int w,
                       float* out, int32_t* x0o, int32_t* y0o, float* txo, float* tyo) {
    float* bx = (float*)__builtin_alloca(sizeof(float) * (size_t)w);
    float* by = (float*)__builtin_alloca(sizeof(float) * (size_t)h);
    tha4o_base_grid(w, bx);
    tha4o_base_grid(h, by);
    for (int b = 0; b < n; ++b)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                long p = ((long)b * h + y) * w + x;
                float gx = bx[x] + grid_change[((long)b * 2 + 0) * h * w + (long)y * w + x];
                float gy = by[y] + grid_change[((long)b * 2 + 1) * h * w + (long)y * w + x];
                float ix = src_index(gx, w), iy = src_index(gy, h);
                float fx = floorf(ix), fy = floorf(iy);
                int x0 = (int)fx, y0 = (int)fy;
                float tx = ix - fx, ty = iy - fy;
                if (x0o) x0o[p] = x0;
                if (y0o) y0o[p] = y0;
                if (txo) txo[p] = tx;
                if (tyo) tyo[p] = ty;
                if (!out) continue;
                float wnw = ((fx + 1.0f) - ix) * ((fy + 1.0f) - iy), wne = (ix - fx) * ((fy + 1.0f) - iy);
                float wsw = ((fx + 1.0f) - ix) * (iy - fy), wse = (ix - fx) * (iy - fy);
                int x1 = x0 + 1, y1 = y0 + 1;
                for (int ch = 0; ch < c; ++ch) {
                    const float* im = image + ((long)b * c + ch) * h * w;
                    float acc = 0.0f;
                    acc += im[(long)y0 * w + x0] * wnw;
                    if (x1 < w) acc += im[(long)y0 * w + x1] * wne;
                    if (y1 < h) acc += im[(long)y1 * w + x0] * wsw;
                    if (x1 < w && y1 < h) acc += im[(long)y1 * w + x1] * wse;
                    out[((long)b * c + ch) * h * w + (long)y * w + x] = acc;
                }
            }
}

/* interpolate(bilinear, align_corners=False): in [n,c,hi,wi] -> out [n,c,ho,wo]. */
void tha4o_resize_bilinear(const float* in, int n, int c, int hi, int wi, int ho, int wo, float* out) {
    float sy = (float)hi / (float)ho, sx = (float)wi / (float)wo;
    for (int y = 0; y < ho; ++y) {
        float fy = sy * ((float)y + 0.5f) - 0.5f;
        if (fy < 0.0f) fy = 0.0f;
        int y0 = (int)fy;
        int y1 = y0 + ((y0 < hi - 1) ? 1 : 0);
        float ly1 = fy - (float)y0, ly0 = 1.0f - ly1;
        for (int x = 0; x < wo; ++x) {
            float fx = sx * ((float)x + 0.5f) - 0.5f;
            if (fx < 0.0f) fx = 0.0f;
            int x0 = (int)fx;
            int x1 = x0 + ((x0 < wi - 1) ? 1 : 0);
            float lx1 = fx - (float)x0, lx0 = 1.0f - lx1;
            for (long pc = 0; pc < (long)n * c; ++pc) {
                const float* im = in + pc * hi * wi;
                out[pc * ho * wo + (long)y * wo + x] =
                    ly0 * (lx0 * im[(long)y0 * wi + x0] + lx1 * im[(long)y0 * wi + x1]) +
                    ly1 * (lx0 * im[(long)y1 * wi + x0] + lx1 * im[(long)y1 * wi + x1]);
            }
        }
    }
}
