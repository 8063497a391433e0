// Rows N3 / N4 of the scope table: the two other sample formats the reference feeds its model from, and the
// live viewer's per-frame pre / post arithmetic, as small memory-bound device kernels.
//
//   cached_assemble_kernel : load_cached_sample + the tail of __getitem__
//                            (src/foundation_stereo_depth/dataset.py:86-106, 302-311): the npz read-through
//                            cache holds uint8 HWC views and a float16 disparity ALREADY at the model
//                            resolution -> u8 / 255 (float32 division), HWC -> CHW, f16 -> f32, valid_mask,
//                            valid count, and the per-block gray sums the augmentation kernels need.
//   live_preprocess_kernel : preprocess_rgb x 2 + cat (src/live_camera/depth_live_dl.py:225-229, 516-520):
//                            BGR -> RGB, cv2.resize(INTER_LINEAR) on uint8 - OpenCV's 11-bit fixed-point
//                            HResizeLinear / VResizeLinear arithmetic, bit-exact - then / 255, CHW.
//   live_postprocess_kernel: EMA smoothing (depth_live_dl.py:531-538), disparity_to_depth (:371-377),
//                            confidence_from_logvar (:380-381).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sdn {

// grid = (ceil(W/128) * ceil(H/PRE_ROWS), B), block = 128: the same block -> gray-partial mapping as
// decode_resize_kernel, so augment_point_kernel / blur_noise_kernel run unchanged afterwards.
__global__ void __launch_bounds__(128) cached_assemble_kernel(
    const uint8_t* __restrict__ L, const uint8_t* __restrict__ R, const __half* __restrict__ D, int B, int H, int W,
    float* __restrict__ input, float* __restrict__ target, uint8_t* __restrict__ mask,
    unsigned long long* __restrict__ valid_count, const AugParams* __restrict__ aug, float* __restrict__ gray_part,
    int parts_per_view) {
    SDN_PDL_ENTRY();
    const int n = blockIdx.y;
    const int xblocks = (W + 127) / 128;
    const int xb = blockIdx.x % xblocks, yb = blockIdx.x / xblocks;
    const int x = xb * 128 + threadIdx.x;
    const size_t plane = (size_t)H * W;
    const uint8_t* Ln = L + (size_t)n * plane * 3;
    const uint8_t* Rn = R + (size_t)n * plane * 3;
    const __half* Dn = D + (size_t)n * plane;
    float gsumL = 0.f, gsumR = 0.f;
    unsigned int cnt = 0;
    float fbL = 1.f, fbR = 1.f;
    if (aug != nullptr) { fbL = aug[2 * n].brightness; fbR = aug[2 * n + 1].brightness; }
    if (x < W) {
        for (int yy = 0; yy < PRE_ROWS; ++yy) {
            const int y = yb * PRE_ROWS + yy;
            if (y >= H) break;
            const size_t pix = (size_t)y * W + x;
            float l[3], r[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                l[c] = u8_over_255(__ldg(Ln + pix * 3 + c));
                r[c] = u8_over_255(__ldg(Rn + pix * 3 + c));
                input[((size_t)n * 6 + c) * plane + pix] = l[c];
                input[((size_t)n * 6 + 3 + c) * plane + pix] = r[c];
            }
            const float t = __half2float(Dn[pix]);
            target[(size_t)n * plane + pix] = t;
            const bool valid = t > 0.f;
            mask[(size_t)n * plane + pix] = valid ? 1 : 0;
            cnt += (valid && isfinite(t)) ? 1u : 0u;
            if (aug != nullptr) {
                gsumL += gray_of(blend(l[0], 0.f, fbL, 1.f - fbL), blend(l[1], 0.f, fbL, 1.f - fbL), blend(l[2], 0.f, fbL, 1.f - fbL));
                gsumR += gray_of(blend(r[0], 0.f, fbR, 1.f - fbR), blend(r[1], 0.f, fbR, 1.f - fbR), blend(r[2], 0.f, fbR, 1.f - fbR));
            }
        }
    }
    __shared__ float redL[4], redR[4];
    __shared__ unsigned int redC[4];
    for (int o = 16; o > 0; o >>= 1) {
        gsumL += __shfl_xor_sync(0xffffffffu, gsumL, o);
        gsumR += __shfl_xor_sync(0xffffffffu, gsumR, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) { redL[threadIdx.x >> 5] = gsumL; redR[threadIdx.x >> 5] = gsumR; redC[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (gray_part != nullptr) {
            gray_part[(size_t)(2 * n) * parts_per_view + blockIdx.x] = (redL[0] + redL[1]) + (redL[2] + redL[3]);
            gray_part[(size_t)(2 * n + 1) * parts_per_view + blockIdx.x] = (redR[0] + redR[1]) + (redR[2] + redR[3]);
        }
        if (valid_count != nullptr) {
            const unsigned int c = redC[0] + redC[1] + redC[2] + redC[3];
            if (c) atomicAdd(valid_count, (unsigned long long)c);
        }
    }
}

// One axis of cv2.resize(INTER_LINEAR) for uint8 (modules/imgproc/src/resize.cpp): source indices and the
// 11-bit fixed-point weights.  scale = 1. / (dsize / ssize) in double, f = (float)((d + 0.5) * scale - 0.5).
// Horizontally an index outside [0, ssize - 1) collapses to one tap; vertically the two rows are clamped and
// the weights kept.  saturate_cast<short>(w * 2048) rounds half to even.
__device__ __forceinline__ void cv_linear_axis(int d, int ssize, double scale, bool horizontal, int& i0, int& i1,
                                               int& w0, int& w1) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (horizontal) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
    }
    i0 = min(max(s, 0), ssize - 1);
    i1 = min(max(s + 1, 0), ssize - 1);
    w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    w1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

// grid = (ceil(W/128), H, 2 views), block = 128; frames are BGR uint8 [Hs, Ws, 3] (what cv2.VideoCapture / cv2.remap
// hand the reference), out = the model input [1, 6, H, W] float32 (left RGB, right RGB).
__global__ void __launch_bounds__(128) live_preprocess_kernel(const uint8_t* __restrict__ frame_l,
                                                              const uint8_t* __restrict__ frame_r, int Hs, int Ws, int H,
                                                              int W, double scale_x, double scale_y,
                                                              float* __restrict__ out) {
    SDN_PDL_ENTRY();
    const int x = blockIdx.x * 128 + threadIdx.x, y = blockIdx.y, view = blockIdx.z;
    if (x >= W) return;
    const uint8_t* src = view == 0 ? frame_l : frame_r;
    int x0, x1, a0, a1, y0, y1, b0, b1;
    cv_linear_axis(x, Ws, scale_x, true, x0, x1, a0, a1);
    cv_linear_axis(y, Hs, scale_y, false, y0, y1, b0, b1);
    const size_t plane = (size_t)H * W;
    const uint8_t* r0 = src + (size_t)y0 * Ws * 3;
    const uint8_t* r1 = src + (size_t)y1 * Ws * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int sc = 2 - c;   // BGR -> RGB (cv2.cvtColor, depth_live_dl.py:226)
        const int t0 = (int)__ldg(r0 + x0 * 3 + sc) * a0 + (int)__ldg(r0 + x1 * 3 + sc) * a1;   // HResizeLinear
        const int t1 = (int)__ldg(r1 + x0 * 3 + sc) * a0 + (int)__ldg(r1 + x1 * 3 + sc) * a1;
        int v = (((b0 * (t0 >> 4)) >> 16) + ((b1 * (t1 >> 4)) >> 16) + 2) >> 2;                  // VResizeLinear
        v = min(max(v, 0), 255);
        out[((size_t)view * 3 + c) * plane + (size_t)y * W + x] = u8_over_255((uint8_t)v);
    }
}

// One pass over the model outputs of a frame: optional exponential smoothing of the disparity (state kept on the
// device between frames), depth = f*B / d where d is finite and > 1e-6 (NaN elsewhere), confidence =
// exp(-0.5 * logvar).  alpha_p = (float)alpha, alpha_s = (float)(1 - alpha): the reference multiplies float32
// arrays by Python doubles, which numpy applies as float32 scalars.
__global__ void __launch_bounds__(256) live_postprocess_kernel(const float* __restrict__ disp,
                                                               const float* __restrict__ logvar, long long n,
                                                               float* __restrict__ ema_state, int ema_mode,
                                                               float alpha_p, float alpha_s, float focal_baseline,
                                                               float* __restrict__ disp_out,
                                                               float* __restrict__ depth_out,
                                                               float* __restrict__ conf_out) {
    SDN_PDL_ENTRY();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float d = disp[i];
        if (ema_mode == 2) d = __fadd_rn(__fmul_rn(alpha_p, d), __fmul_rn(alpha_s, ema_state[i]));
        if (ema_mode != 0) ema_state[i] = d;    // mode 1: first frame, the state becomes the prediction
        if (disp_out != nullptr) disp_out[i] = d;
        if (depth_out != nullptr) depth_out[i] = (isfinite(d) && d > 1e-6f) ? __fdiv_rn(focal_baseline, d) : __int_as_float(0x7fc00000);
        if (conf_out != nullptr && logvar != nullptr) conf_out[i] = expf(__fmul_rn(-0.5f, logvar[i]));
    }
}

}  // namespace sdn
