// Device input pipeline: the per-sample work the reference does on CPU DataLoader
// workers (src/foundation_stereo_depth/dataset.py), as memory-bound kernels.
//
//   decode_resize_kernel : dataset.py:23-30 (depth_uint8_decoding), :184-193
//                          (_load_rgb: /255 + bilinear, align_corners=False, no
//                          antialias), :195-212 (_load_disparity: decode -> bilinear ->
//                          * W_out/W_in), :305-311 (cat([L,R]), valid_mask = target > 0).
//                          Bit-exact against the CPU path: every multiply / add that
//                          the CPU kernel performs is spelled with an explicit
//                          round-to-nearest intrinsic so nvcc cannot re-contract it.
//   augment_point_kernel : dataset.py:248-270 (_augment_rgb) brightness -> contrast ->
//                          saturation -> hue -> gamma [-> noise -> clamp], semantics of
//                          torchvision/transforms/_functional_tensor.py.
//   blur_noise_kernel    : the 5x5 Gaussian blur (reflect padding) for the few views
//                          whose blur coin came up, then noise + clamp.
// Raw uint8 HWC images are the kernel input: PNG inflate is out of scope.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sdn {

// Per-view photometric parameters (explicit, so the reference's samplers
// dataset.py:214-246 can be replayed exactly by tests).
struct AugParams {
    float brightness;  // f_b
    float contrast;    // f_c
    float saturation;  // f_s
    float hue;         // delta h in [-0.5, 0.5]
    float gamma;       // gamma
    float blur_sigma;  // > 0 => blur with this sigma
    float noise_std;   // > 0 => add N(0,1)*noise_std
    uint32_t noise_seed;
};

// PyTorch area_pixel_compute_source_index + guard_index_and_lambda
// (aten/src/ATen/native/UpSample.h), align_corners = false, in fp32.
__device__ __forceinline__ void bilinear_src(float scale, int dst, int in_size, int out_size, int& i0, int& i1,
                                             float& l0, float& l1) {
    if (in_size == out_size) { i0 = dst; i1 = dst; l0 = 1.f; l1 = 0.f; return; }
    // torch's CPU build contracts scale * (dst + 0.5) - 0.5 into one fused multiply-add
    float real = __fmaf_rn(scale, __fadd_rn((float)dst, 0.5f), -0.5f);
    if (real < 0.f) real = 0.f;
    int idx = (int)floorf(real);
    if (idx > in_size - 1) idx = in_size - 1;
    float lam = __fsub_rn(real, (float)idx);
    lam = fminf(fmaxf(lam, 0.f), 1.f);
    i0 = idx;
    i1 = idx + (idx < in_size - 1 ? 1 : 0);
    l1 = lam;
    l0 = __fsub_rn(1.f, lam);
}

// out = h0*(w0*a + w1*b) + h1*(w0*c + w1*e).  torch's CPU kernel exists in two
// compiled instantiations that fuse the multiply-adds differently (<= 1 ulp
// apart; see oracle/stereo_oracle.py:bilinear_resize):
//   FOURTERM = false ("separable", canonical): fma(h0, fma(w0,a,w1*b), h1*fma(w0,c,w1*e))
//   FOURTERM = true  : w_ij = h_i*w_j; fma(w11,e, fma(w10,c, fma(w00,a, w01*b)))
template <bool FOURTERM>
__device__ __forceinline__ float bilerp(float a, float b, float c, float e, float w0, float w1, float h0, float h1) {
    if (FOURTERM) {
        const float w00 = __fmul_rn(h0, w0), w01 = __fmul_rn(h0, w1), w10 = __fmul_rn(h1, w0), w11 = __fmul_rn(h1, w1);
        return __fmaf_rn(w11, e, __fmaf_rn(w10, c, __fmaf_rn(w00, a, __fmul_rn(w01, b))));
    }
    // w1 == 0 implies w0 == 1 (bilinear_src): fma(1, a, 0 * b) == a bit for bit for finite non-negative inputs
    const float top = w1 == 0.f ? a : __fmaf_rn(w0, a, __fmul_rn(w1, b));
    const float bot = w1 == 0.f ? c : __fmaf_rn(w0, c, __fmul_rn(w1, e));
    return __fmaf_rn(h0, top, __fmul_rn(h1, bot));
}

__device__ __forceinline__ float gray_of(float r, float g, float b) {
    // (0.2989 * r + 0.587 * g + 0.114 * b) evaluated left to right in fp32
    return __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
}
__device__ __forceinline__ float blend(float x, float y, float ratio, float one_minus) {
    // _blend: (ratio * img1 + (1 - ratio) * img2).clamp(0, 1)
    return fminf(fmaxf(__fadd_rn(__fmul_rn(ratio, x), __fmul_rn(one_minus, y)), 0.f), 1.f);
}

// grid = (ceil(W/128) * ceil(H/ROWS_PER_BLOCK), B); block = 128 threads (one output column each).
constexpr int PRE_ROWS = 8;

template <bool FOURTERM>
__global__ void __launch_bounds__(128) decode_resize_kernel(
    const uint8_t* __restrict__ L, const uint8_t* __restrict__ R, const uint8_t* __restrict__ D, int B, int Hs, int Ws,
    int H, int W, float* __restrict__ input, float* __restrict__ target, uint8_t* __restrict__ mask,
    unsigned long long* __restrict__ valid_count, const AugParams* __restrict__ aug, float* __restrict__ gray_part,
    int parts_per_view) {
    SDN_PDL_ENTRY();
    const int n = blockIdx.y;
    const int xblocks = (W + 127) / 128;
    const int xb = blockIdx.x % xblocks;
    const int yb = blockIdx.x / xblocks;
    const int x = xb * 128 + threadIdx.x;
    const float sy = (float)Hs / (float)H;
    const float sx = (float)Ws / (float)W;
    const float wscale = (float)((double)W / (double)Ws);
    const size_t src_img = (size_t)Hs * Ws * 3;
    const uint8_t* Ln = L + (size_t)n * src_img;
    const uint8_t* Rn = R + (size_t)n * src_img;
    const uint8_t* Dn = D + (size_t)n * src_img;
    const size_t plane = (size_t)H * W;
    float gsumL = 0.f, gsumR = 0.f;
    unsigned int cnt = 0;
    float fbL = 1.f, fbR = 1.f;
    if (aug != nullptr) { fbL = aug[2 * n].brightness; fbR = aug[2 * n + 1].brightness; }
    if (x < W) {
        int x0, x1;
        float w0, w1;
        bilinear_src(sx, x, Ws, W, x0, x1, w0, w1);
        for (int yy = 0; yy < PRE_ROWS; ++yy) {
            const int y = yb * PRE_ROWS + yy;
            if (y >= H) break;
            int y0, y1;
            float h0, h1;
            bilinear_src(sy, y, Hs, H, y0, y1, h0, h1);
            const size_t o00 = ((size_t)y0 * Ws + x0) * 3, o01 = ((size_t)y0 * Ws + x1) * 3;
            const size_t o10 = ((size_t)y1 * Ws + x0) * 3, o11 = ((size_t)y1 * Ws + x1) * 3;
            const size_t opix = (size_t)y * W + x;
            float rgbL[3], rgbR[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float a = __fdiv_rn((float)__ldg(Ln + o00 + c), 255.f);
                const float b = __fdiv_rn((float)__ldg(Ln + o01 + c), 255.f);
                const float cc = __fdiv_rn((float)__ldg(Ln + o10 + c), 255.f);
                const float e = __fdiv_rn((float)__ldg(Ln + o11 + c), 255.f);
                rgbL[c] = bilerp<FOURTERM>(a, b, cc, e, w0, w1, h0, h1);
                input[((size_t)n * 6 + c) * plane + opix] = rgbL[c];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float a = __fdiv_rn((float)__ldg(Rn + o00 + c), 255.f);
                const float b = __fdiv_rn((float)__ldg(Rn + o01 + c), 255.f);
                const float cc = __fdiv_rn((float)__ldg(Rn + o10 + c), 255.f);
                const float e = __fdiv_rn((float)__ldg(Rn + o11 + c), 255.f);
                rgbR[c] = bilerp<FOURTERM>(a, b, cc, e, w0, w1, h0, h1);
                input[((size_t)n * 6 + 3 + c) * plane + opix] = rgbR[c];
            }
            float dv[4];
            const size_t offs[4] = {o00, o01, o10, o11};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float r = (float)__ldg(Dn + offs[k]), g = (float)__ldg(Dn + offs[k] + 1),
                            b = (float)__ldg(Dn + offs[k] + 2);
                // exact in fp32 (max 16,646,655 < 2^24): R*255*255 + G*255 + B, then one rounded divide
                const float s = __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(r, 255.f), 255.f), __fmul_rn(g, 255.f)), b);
                dv[k] = __fdiv_rn(s, 1000.f);
            }
            const float t = __fmul_rn(bilerp<FOURTERM>(dv[0], dv[1], dv[2], dv[3], w0, w1, h0, h1), wscale);
            target[(size_t)n * plane + opix] = t;
            const bool valid = t > 0.f;
            mask[(size_t)n * plane + opix] = valid ? 1 : 0;
            cnt += (valid && isfinite(t)) ? 1u : 0u;
            if (aug != nullptr) {
                // brightness-adjusted grayscale, whose image mean adjust_contrast needs
                gsumL += gray_of(blend(rgbL[0], 0.f, fbL, 1.f - fbL), blend(rgbL[1], 0.f, fbL, 1.f - fbL),
                                 blend(rgbL[2], 0.f, fbL, 1.f - fbL));
                gsumR += gray_of(blend(rgbR[0], 0.f, fbR, 1.f - fbR), blend(rgbR[1], 0.f, fbR, 1.f - fbR),
                                 blend(rgbR[2], 0.f, fbR, 1.f - fbR));
            }
        }
    }
    // block reduction (4 warps)
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

// Shared-memory staged variant of decode_resize_kernel (same arithmetic, bit-exact):
// the block first copies the source byte ranges it needs (all three images) into
// shared memory with 128-bit coalesced loads, then computes from there.  Needs 16-byte
// aligned image rows (Ws*3 % 16 == 0 and 16-byte aligned bases) - the host falls back
// to the direct kernel otherwise.  256 threads: 128 output columns x PRE_ROWS rows.
// u8 / 255 correctly rounded (== the reference's float32 division, dataset.py:185) without a divide or a
// table: one Newton step on q = b * fl(1/255) is exact for all 256 inputs (checked exhaustively).
// byte -> float on the FMA pipe: 0x4B000000 | b is the float 8388608 + b exactly (the integer -> float conversion
// instruction runs on the quarter-rate conversion pipe, and this kernel is instruction-bound)
__device__ __forceinline__ float byte_to_float(uint32_t b) { return __uint_as_float(0x4B000000u | b) - 8388608.0f; }
__device__ __forceinline__ float u8_over_255(uint8_t v) {
    const float b = byte_to_float(v);
    const float r = 0.003921568859368563f;   // fl(1/255)
    const float q = __fmul_rn(b, r);
    const float rem = __fmaf_rn(-q, 255.f, b);
    return __fmaf_rn(rem, r, q);
}

// Three consecutive bytes (one HWC pixel) at an arbitrary shared-memory byte address as the low 24 bits of a word:
// two aligned 32-bit loads and a funnel shift instead of three byte loads (the pixel pitch of 9 bytes between
// neighbouring threads already costs ~3 wavefronts per load instruction).
__device__ __forceinline__ uint32_t load_rgb24(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    const uint32_t sh = uint32_t(a & 3) * 8;
    return __funnelshift_r(w[0], w[1], sh) & 0x00FFFFFFu;
}
// depth_uint8_decoding (dataset.py:23-30) of one pixel held as R | G << 8 | B << 16: exact in fp32 (max
// 16,646,655 < 2^24), then one correctly rounded divide.
// The quotient s / 1000 without a divide: q = s * fl(1/1000), one Newton correction - bit-identical to the IEEE
// division for EVERY integer s in [0, 16,646,655] (checked exhaustively on the device, tests/div_probe.cu).
__device__ __forceinline__ float decode_rgb24(uint32_t px) {
    const float r = byte_to_float(px & 0xFFu), g = byte_to_float((px >> 8) & 0xFFu), b = byte_to_float(px >> 16);
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(r, 255.f), 255.f), __fmul_rn(g, 255.f)), b);
    const float rcp = 1.0f / 1000.0f;
    const float q = __fmul_rn(s, rcp);
    return __fmaf_rn(__fmaf_rn(-q, 1000.f, s), rcp, q);
}

constexpr int PRE_THREADS = 256;   // 128 output columns x 2 row phases
template <bool FOURTERM>
__global__ void __launch_bounds__(PRE_THREADS) decode_resize_smem_kernel(
    const uint8_t* __restrict__ L, const uint8_t* __restrict__ R, const uint8_t* __restrict__ D, int B, int Hs, int Ws,
    int H, int W, float* __restrict__ input, float* __restrict__ target, uint8_t* __restrict__ mask,
    unsigned long long* __restrict__ valid_count, const AugParams* __restrict__ aug, float* __restrict__ gray_part,
    int parts_per_view, int max_rows, int row_bytes) {
    SDN_PDL_ENTRY();
    // A CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... of the (sample, row block, column block) grid;
    // with a second shared-memory buffer the source rows of tile t+1 stream in (cp.async, every 16-byte chunk in
    // flight at once) while tile t is computed.  Measured: the kernel is bound by its compute phase (latency of
    // the conversion / interpolation chains), not by staging - one persistent 512-thread CTA per SM with two
    // buffers ran 0.96 ms where three independent 256-thread CTAs per SM (one tile each, launched that way by
    // the host: gridDim.x == number of tiles, one buffer) run 0.68 ms - so occupancy wins over pipelining here.
    extern __shared__ __align__(16) uint8_t sm_all[];
    const int xblocks = (W + 127) / 128;
    const int tiles_per_sample = parts_per_view;                 // == xblocks * yblocks
    const int num_tiles = B * tiles_per_sample;
    const size_t buf_bytes = (size_t)3 * max_rows * row_bytes;
    const float sy = (float)Hs / (float)H;
    const float sx = (float)Ws / (float)W;
    const float wscale = (float)((double)W / (double)Ws);
    const size_t src_img = (size_t)Hs * Ws * 3;
    const size_t plane = (size_t)H * W;
    const size_t pitch = (size_t)Ws * 3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    struct Foot { int n, part, ox0, oy0, row_first, nrows, byte0, chunks; };
    auto footprint = [&](int t) {
        Foot f;
        f.n = t / tiles_per_sample;
        f.part = t - f.n * tiles_per_sample;
        const int xb = f.part % xblocks, yb = f.part / xblocks;
        f.ox0 = xb * 128;
        f.oy0 = yb * PRE_ROWS;
        const int ox1 = min(f.ox0 + 127, W - 1), oy1 = min(f.oy0 + PRE_ROWS - 1, H - 1);
        int a0, a1, b0, b1;
        float t0, t1;
        bilinear_src(sx, f.ox0, Ws, W, a0, a1, t0, t1);
        bilinear_src(sx, ox1, Ws, W, b0, b1, t0, t1);
        const int col_first = a0, col_last = b1;
        bilinear_src(sy, f.oy0, Hs, H, a0, a1, t0, t1);
        bilinear_src(sy, oy1, Hs, H, b0, b1, t0, t1);
        f.row_first = a0;
        f.nrows = b1 - a0 + 1;
        f.byte0 = (col_first * 3) & ~15;                       // 16-byte aligned inside the row
        const int byte1 = min(((col_last + 1) * 3 + 15) & ~15, Ws * 3);
        f.chunks = (byte1 - f.byte0) >> 4;
        return f;
    };
    // one warp per (image, source row) line, lanes over its 16-byte chunks
    auto stage = [&](const Foot& f, uint8_t* sm) {
        for (int line = warp; line < 3 * f.nrows; line += PRE_THREADS / 32) {
            const int im = line >= 2 * f.nrows ? 2 : (line >= f.nrows ? 1 : 0);
            const int rr = line - im * f.nrows;
            const uint8_t* base = (im == 0 ? L : (im == 1 ? R : D)) + (size_t)f.n * src_img;
            const uint4* src = reinterpret_cast<const uint4*>(base + (size_t)(f.row_first + rr) * pitch + f.byte0);
            uint4* dst = reinterpret_cast<uint4*>(sm + (size_t)(im * max_rows + rr) * row_bytes);
            for (int ck = lane; ck < f.chunks; ck += 32) {
                const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst + ck));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + ck) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    __shared__ float redL[PRE_THREADS / 32], redR[PRE_THREADS / 32];
    __shared__ unsigned int redC[PRE_THREADS / 32];
    __shared__ int s_y0[PRE_ROWS], s_y1[PRE_ROWS];       // vertical taps of the tile's rows: once per tile, not per thread
    __shared__ float s_h0[PRE_ROWS], s_h1[PRE_ROWS];
    int t = blockIdx.x;
    if (t >= num_tiles) return;
    Foot cur = footprint(t);
    stage(cur, sm_all);
    int buf = 0;
    for (; t < num_tiles; t += gridDim.x) {
        uint8_t* sm = sm_all + (size_t)buf * buf_bytes;
        const int tn = t + gridDim.x;
        Foot nxt = cur;
        if (tn < num_tiles) {
            nxt = footprint(tn);
            stage(nxt, sm_all + (size_t)(buf ^ 1) * buf_bytes);      // (that buffer was released by the barrier below)
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        if (threadIdx.x < PRE_ROWS) {
            int y0, y1;
            float h0, h1;
            bilinear_src(sy, min(cur.oy0 + (int)threadIdx.x, H - 1), Hs, H, y0, y1, h0, h1);
            s_y0[threadIdx.x] = y0 - cur.row_first; s_y1[threadIdx.x] = y1 - cur.row_first;
            s_h0[threadIdx.x] = h0; s_h1[threadIdx.x] = h1;
        }
        __syncthreads();

        const int n = cur.n, byte0 = cur.byte0;
        float* const in_n = input + (size_t)n * 6 * plane;
        float* const tg_n = target + (size_t)n * plane;
        uint8_t* const mk_n = mask + (size_t)n * plane;
        const int x = cur.ox0 + (threadIdx.x & 127);
        float gsumL = 0.f, gsumR = 0.f;
        unsigned int cnt = 0;
        float fbL = 1.f, fbR = 1.f;
        if (aug != nullptr) { fbL = aug[2 * n].brightness; fbR = aug[2 * n + 1].brightness; }
        if (x < W) {
            int x0, x1;
            float w0, w1;
            bilinear_src(sx, x, Ws, W, x0, x1, w0, w1);
            const int c0 = x0 * 3 - byte0, c1 = x1 * 3 - byte0;
            for (int yy = (threadIdx.x >> 7); yy < PRE_ROWS; yy += PRE_THREADS / 128) {
                const int y = cur.oy0 + yy;
                if (y >= H) break;
                const int y0 = s_y0[yy], y1 = s_y1[yy];
                const float h0 = s_h0[yy], h1 = s_h1[yy];
                const int opix = y * W + x;
                // A tap whose weight is exactly 0 (integer scale ratios: 960 -> 320 samples column 3x+1 with
                // w1 = 0) contributes fma(w0, a, 0 * b) = the same bits for ANY finite b: reuse the other tap
                // instead of fetching and converting it.  Same for the vertical pair.
                const bool skip_x = w1 == 0.f, skip_y = h1 == 0.f;
                float rgb[2][3];
#pragma unroll
                for (int im = 0; im < 2; ++im) {
                    const uint8_t* r0 = sm + (im * max_rows + y0) * row_bytes;
                    const uint8_t* r1 = sm + (im * max_rows + y1) * row_bytes;
                    const uint32_t p00 = load_rgb24(r0 + c0);
                    const uint32_t p01 = skip_x ? p00 : load_rgb24(r0 + c1);
                    const uint32_t p10 = skip_y ? p00 : load_rgb24(r1 + c0);
                    const uint32_t p11 = skip_y ? p01 : (skip_x ? p10 : load_rgb24(r1 + c1));
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float a = u8_over_255((uint8_t)(p00 >> (8 * c)));
                        const float b = skip_x ? a : u8_over_255((uint8_t)(p01 >> (8 * c)));
                        const float cc = skip_y ? a : u8_over_255((uint8_t)(p10 >> (8 * c)));
                        const float e = skip_y ? b : (skip_x ? cc : u8_over_255((uint8_t)(p11 >> (8 * c))));
                        rgb[im][c] = bilerp<FOURTERM>(a, b, cc, e, w0, w1, h0, h1);
                        in_n[(size_t)(im * 3 + c) * plane + opix] = rgb[im][c];
                    }
                }
                const uint8_t* d0 = sm + (2 * max_rows + y0) * row_bytes;
                const uint8_t* d1 = sm + (2 * max_rows + y1) * row_bytes;
                float dv[4];
                dv[0] = decode_rgb24(load_rgb24(d0 + c0));
                dv[1] = skip_x ? dv[0] : decode_rgb24(load_rgb24(d0 + c1));
                dv[2] = skip_y ? dv[0] : decode_rgb24(load_rgb24(d1 + c0));
                dv[3] = skip_y ? dv[1] : (skip_x ? dv[2] : decode_rgb24(load_rgb24(d1 + c1)));
                const float tv = __fmul_rn(bilerp<FOURTERM>(dv[0], dv[1], dv[2], dv[3], w0, w1, h0, h1), wscale);
                tg_n[opix] = tv;
                const bool valid = tv > 0.f;
                mk_n[opix] = valid ? 1 : 0;
                cnt += (valid && isfinite(tv)) ? 1u : 0u;
                if (aug != nullptr) {
                    gsumL += gray_of(blend(rgb[0][0], 0.f, fbL, 1.f - fbL), blend(rgb[0][1], 0.f, fbL, 1.f - fbL),
                                     blend(rgb[0][2], 0.f, fbL, 1.f - fbL));
                    gsumR += gray_of(blend(rgb[1][0], 0.f, fbR, 1.f - fbR), blend(rgb[1][1], 0.f, fbR, 1.f - fbR),
                                     blend(rgb[1][2], 0.f, fbR, 1.f - fbR));
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            gsumL += __shfl_xor_sync(0xffffffffu, gsumL, o);
            gsumR += __shfl_xor_sync(0xffffffffu, gsumR, o);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if (lane == 0) { redL[warp] = gsumL; redR[warp] = gsumR; redC[warp] = cnt; }
        __syncthreads();      // also: every thread is done reading this tile's buffer
        if (threadIdx.x == 0) {
            if (gray_part != nullptr) {
                float a = 0.f, b = 0.f;
                for (int w8 = 0; w8 < PRE_THREADS / 32; ++w8) { a += redL[w8]; b += redR[w8]; }
                gray_part[(size_t)(2 * n) * parts_per_view + cur.part] = a;
                gray_part[(size_t)(2 * n + 1) * parts_per_view + cur.part] = b;
            }
            if (valid_count != nullptr) {
                unsigned int c = 0;
                for (int w8 = 0; w8 < PRE_THREADS / 32; ++w8) c += redC[w8];
                if (c) atomicAdd(valid_count, (unsigned long long)c);
            }
        }
        cur = nxt;
        buf ^= 1;
    }
}

// ---------------------------------------------------------------- Philox RNG
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// three standard normals for (view, pixel)
__device__ __forceinline__ void normal3(uint32_t seed, uint32_t view, uint32_t pix, float (&z)[3]) {
    uint32_t r[4];
    philox4x32_10(pix, view, 0x5DEECE66u, 0u, seed, 0xB5297A4Du, r);
    const float u0 = ((float)(r[0] >> 8) + 0.5f) * (1.f / 16777216.f);
    const float u1 = ((float)(r[1] >> 8) + 0.5f) * (1.f / 16777216.f);
    const float u2 = ((float)(r[2] >> 8) + 0.5f) * (1.f / 16777216.f);
    const float u3 = ((float)(r[3] >> 8) + 0.5f) * (1.f / 16777216.f);
#ifdef SDN_AUG_PRECISE
    const float ra = sqrtf(-2.f * logf(u0)), rb = sqrtf(-2.f * logf(u2));
    float s, c;
    sincospif(2.f * u1, &s, &c);
    z[0] = ra * c;
    z[1] = ra * s;
    z[2] = rb * cospif(2.f * u3);
#else
    // Box-Muller on the special-function unit: the noise is specified statistically, not bit-wise
    const float ra = __fsqrt_rn(-2.f * __logf(u0)), rb = __fsqrt_rn(-2.f * __logf(u2));
    float s, c;
    __sincosf(6.283185307179586f * u1, &s, &c);
    z[0] = ra * c;
    z[1] = ra * s;
    z[2] = rb * __cosf(6.283185307179586f * u3);
#endif
}

__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// brightness .. gamma for one pixel, torchvision tensor semantics
__device__ __forceinline__ void augment_pixel(float (&v)[3], const AugParams& a, float gray_mean) {
    const float fb = a.brightness, fc = a.contrast, fs = a.saturation;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = blend(v[c], 0.f, fb, 1.f - fb);
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = blend(v[c], gray_mean, fc, 1.f - fc);
    const float g = gray_of(v[0], v[1], v[2]);
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = blend(v[c], g, fs, 1.f - fs);
    // adjust_hue: _rgb2hsv -> (h + delta) mod 1 -> _hsv2rgb
    const float r = v[0], gg = v[1], b = v[2];
    const float maxc = fmaxf(r, fmaxf(gg, b)), minc = fminf(r, fminf(gg, b));
    const bool eqc = maxc == minc;
    const float cr = maxc - minc;
#ifdef SDN_AUG_PRECISE
    const float s = __fdiv_rn(cr, eqc ? 1.f : maxc);
    const float div = eqc ? 1.f : cr;
    const float rc = __fdiv_rn(maxc - r, div), gc = __fdiv_rn(maxc - gg, div), bc = __fdiv_rn(maxc - b, div);
#else
    // The photometric chain is checked against the reference to 1e-5 absolute (torch's own CPU / GPU
    // results differ by more), so the quotients and the gamma power use the special-function unit
    // (~2 ulp): ~250 instead of ~1100 instructions per pixel, which is what made this kernel ALU-bound.
    const float s = __fdividef(cr, eqc ? 1.f : maxc);
    const float inv = __fdividef(1.f, eqc ? 1.f : cr);
    const float rc = (maxc - r) * inv, gc = (maxc - gg) * inv, bc = (maxc - b) * inv;
#endif
    const float hr = (maxc == r) ? (bc - gc) : 0.f;
    const float hg = ((maxc == gg) && (maxc != r)) ? __fadd_rn(2.f, rc) - bc : 0.f;
    const float hb = ((maxc != gg) && (maxc != r)) ? __fadd_rn(4.f, gc) - rc : 0.f;
    float h = __fadd_rn(__fadd_rn(hr, hg), hb);
#ifdef SDN_AUG_PRECISE
    h = fmodf(__fadd_rn(__fdiv_rn(h, 6.f), 1.f), 1.f);
#else
    h = __fadd_rn(__fmul_rn(h, 0.16666667163372040f), 1.f);   // in [5/6, 11/6]: fmod(x, 1) == x - floor(x), exactly
    h = h - floorf(h);
#endif
    h = __fadd_rn(h, a.hue);
    h = h - floorf(h);  // python-style remainder by 1.0
    if (h >= 1.f) h = 0.f;
    const float h6 = __fmul_rn(h, 6.f);
    const float fi = floorf(h6);
    const float f = __fsub_rn(h6, fi);
    int i = ((int)fi) % 6;
    if (i < 0) i += 6;
    const float p = fminf(fmaxf(__fmul_rn(maxc, __fsub_rn(1.f, s)), 0.f), 1.f);
    const float q = fminf(fmaxf(__fmul_rn(maxc, __fsub_rn(1.f, __fmul_rn(s, f))), 0.f), 1.f);
    const float t = fminf(fmaxf(__fmul_rn(maxc, __fsub_rn(1.f, __fmul_rn(s, __fsub_rn(1.f, f)))), 0.f), 1.f);
    const float vv = maxc;
    float ro, go, bo;
    switch (i) {
        case 0: ro = vv; go = t; bo = p; break;
        case 1: ro = q; go = vv; bo = p; break;
        case 2: ro = p; go = vv; bo = t; break;
        case 3: ro = p; go = q; bo = vv; break;
        case 4: ro = t; go = p; bo = vv; break;
        default: ro = vv; go = p; bo = q; break;
    }
    // adjust_gamma: (1.0 * x ** gamma).clamp(0, 1)
    // x in [0, 1], gamma > 0: exp2(gamma * log2(x)) with the 1-ulp log2f / exp2f is within 1e-6 of
    // powf at a third of the instructions (log2f(0) = -inf -> 0, as powf)
#ifdef SDN_AUG_PRECISE
    v[0] = fminf(fmaxf(exp2f(a.gamma * log2f(ro)), 0.f), 1.f);
    v[1] = fminf(fmaxf(exp2f(a.gamma * log2f(go)), 0.f), 1.f);
    v[2] = fminf(fmaxf(exp2f(a.gamma * log2f(bo)), 0.f), 1.f);
#else
    v[0] = fminf(fmaxf(fast_ex2(a.gamma * __log2f(ro)), 0.f), 1.f);
    v[1] = fminf(fmaxf(fast_ex2(a.gamma * __log2f(go)), 0.f), 1.f);
    v[2] = fminf(fmaxf(fast_ex2(a.gamma * __log2f(bo)), 0.f), 1.f);
#endif
}

// grid = (ceil(H*W/256), 2*B); in place on input[B,6,H,W]; blurred views are
// written (post-gamma) to blur_tmp[view] instead and finished by blur_noise_kernel.
__global__ void __launch_bounds__(256) augment_point_kernel(float* __restrict__ input, int B, int H, int W,
                                                            const AugParams* __restrict__ aug,
                                                            const float* __restrict__ gray_part, int parts_per_view,
                                                            float* __restrict__ blur_tmp) {
    SDN_PDL_ENTRY();
    const int view = blockIdx.y;  // 2*n + {0: left, 1: right}
    const AugParams a = aug[view];
    __shared__ float s_mean;
    if (threadIdx.x < 32) {
        // fixed-shape fp64 reduction of the per-block partials by one warp: deterministic and ~100 cycles
        double s = 0.0;
        for (int i = threadIdx.x; i < parts_per_view; i += 32) s += (double)gray_part[(size_t)view * parts_per_view + i];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) s_mean = (float)(s / ((double)H * (double)W));
    }
    __syncthreads();
    const size_t plane = (size_t)H * W;
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= plane) return;
    float* base = input + ((size_t)(view >> 1) * 6 + (view & 1) * 3) * plane;
    float v[3] = {base[pix], base[plane + pix], base[2 * plane + pix]};
    augment_pixel(v, a, s_mean);
    if (a.blur_sigma > 0.f) {
        float* t = blur_tmp + (size_t)view * 3 * plane;
        t[pix] = v[0]; t[plane + pix] = v[1]; t[2 * plane + pix] = v[2];
        return;
    }
    if (a.noise_std > 0.f) {
        float z[3];
        normal3(a.noise_seed, (uint32_t)view, (uint32_t)pix, z);
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = fmaf(z[c], a.noise_std, v[c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) base[c * plane + pix] = fminf(fmaxf(v[c], 0.f), 1.f);
}

// four standard normals per Philox call (two Box-Muller pairs, no output wasted)
__device__ __forceinline__ void normal4(uint32_t seed, uint32_t view, uint32_t idx, uint32_t stream, float (&z)[4]) {
    uint32_t r[4];
    philox4x32_10(idx, view, 0x5DEECE66u + stream, 0x2545F491u, seed, 0xB5297A4Du, r);
    const float u0 = ((float)(r[0] >> 8) + 0.5f) * (1.f / 16777216.f);
    const float u1 = ((float)(r[1] >> 8) + 0.5f) * (1.f / 16777216.f);
    const float u2 = ((float)(r[2] >> 8) + 0.5f) * (1.f / 16777216.f);
    const float u3 = ((float)(r[3] >> 8) + 0.5f) * (1.f / 16777216.f);
#ifdef SDN_AUG_PRECISE
    const float ra = sqrtf(-2.f * logf(u0)), rb = sqrtf(-2.f * logf(u2));
    float s0, c0, s1, c1;
    sincospif(2.f * u1, &s0, &c0);
    sincospif(2.f * u3, &s1, &c1);
#else
    const float ra = __fsqrt_rn(-2.f * __logf(u0)), rb = __fsqrt_rn(-2.f * __logf(u2));
    float s0, c0, s1, c1;
    __sincosf(6.283185307179586f * u1, &s0, &c0);
    __sincosf(6.283185307179586f * u3, &s1, &c1);
#endif
    z[0] = ra * c0; z[1] = ra * s0; z[2] = rb * c1; z[3] = rb * s1;
}

// Per-view mean of the brightness-adjusted gray image (adjust_contrast's blend target), ONCE per view: one warp
// sums the per-tile partials in a fixed order in fp64 (deterministic).  Every 256-pixel block of the point kernel
// used to redo this reduction behind a barrier, which cost more than the block's own arithmetic.
__global__ void __launch_bounds__(256) view_mean_kernel(const float* __restrict__ gray_part, int parts_per_view, int views,
                                                        double inv_pixels, float* __restrict__ view_mean) {
    SDN_PDL_ENTRY();
    const int view = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (view >= views) return;
    double s = 0.0;
    for (int i = lane; i < parts_per_view; i += 32) s += (double)gray_part[(size_t)view * parts_per_view + i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) view_mean[view] = (float)(s * inv_pixels);
}

// Same chain as augment_point_kernel, FOUR pixels per thread: 128-bit loads / stores of each colour plane and three
// Philox calls for the quad's twelve normals (instead of four calls with a quarter of their output discarded).
// grid = (ceil(H*W / 1024), 2*B); needs H*W % 4 == 0 (16-byte aligned planes).
__global__ void __launch_bounds__(256) augment_point4_kernel(float* __restrict__ input, int B, int H, int W,
                                                             const AugParams* __restrict__ aug,
                                                             const float* __restrict__ view_mean,
                                                             float* __restrict__ blur_tmp) {
    SDN_PDL_ENTRY();
    const int view = blockIdx.y;
    const AugParams a = aug[view];
    const float mean = __ldg(view_mean + view);
    const size_t plane = (size_t)H * W;
    const size_t quad = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (quad * 4 >= plane) return;
    float* base = input + ((size_t)(view >> 1) * 6 + (view & 1) * 3) * plane;
    float4 ch[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) ch[c] = reinterpret_cast<const float4*>(base + c * plane)[quad];
    float px[4][3] = {{ch[0].x, ch[1].x, ch[2].x}, {ch[0].y, ch[1].y, ch[2].y}, {ch[0].z, ch[1].z, ch[2].z}, {ch[0].w, ch[1].w, ch[2].w}};
#pragma unroll
    for (int j = 0; j < 4; ++j) augment_pixel(px[j], a, mean);
    float* dst = base;
    if (a.blur_sigma > 0.f) {
        dst = blur_tmp + (size_t)view * 3 * plane;      // post-gamma values: blur_noise_kernel finishes the view
    } else {
        if (a.noise_std > 0.f) {
            float z[12];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float z4[4];
                normal4(a.noise_seed, (uint32_t)view, (uint32_t)quad, (uint32_t)k, z4);
#pragma unroll
                for (int i = 0; i < 4; ++i) z[4 * k + i] = z4[i];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 3; ++c) px[j][c] = fmaf(z[3 * j + c], a.noise_std, px[j][c]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 3; ++c) px[j][c] = fminf(fmaxf(px[j][c], 0.f), 1.f);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
        reinterpret_cast<float4*>(dst + c * plane)[quad] = make_float4(px[0][c], px[1][c], px[2][c], px[3][c]);
}

// 5x5 Gaussian (outer product of the normalised 1-D kernel, reflect padding,
// torchvision gaussian_blur) + noise + clamp, only for views with blur_sigma > 0.
// grid = (ceil(W/32), ceil(H/8), 2*B), block = (32, 8).
template <int KS>
__global__ void __launch_bounds__(256) blur_noise_kernel(float* __restrict__ input, int B, int H, int W,
                                                         const AugParams* __restrict__ aug,
                                                         const float* __restrict__ blur_tmp) {
    SDN_PDL_ENTRY();
    const int view = blockIdx.z;
    const AugParams a = aug[view];
    if (!(a.blur_sigma > 0.f)) return;
    constexpr int R = KS / 2;
    __shared__ float tile[3][8 + 2 * R][32 + 2 * R];
    __shared__ float k1d[KS];
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        float pdf[KS], sum = 0.f;
        for (int i = 0; i < KS; ++i) {
            const float xv = -0.5f * (KS - 1) + (float)i;
            const float q = __fdiv_rn(xv, a.blur_sigma);
            pdf[i] = expf(__fmul_rn(-0.5f, __fmul_rn(q, q)));
            sum = __fadd_rn(sum, pdf[i]);
        }
        for (int i = 0; i < KS; ++i) k1d[i] = __fdiv_rn(pdf[i], sum);
    }
    const size_t plane = (size_t)H * W;
    const float* src = blur_tmp + (size_t)view * 3 * plane;
    const int x0 = blockIdx.x * 32 - R, y0 = blockIdx.y * 8 - R;
    for (int idx = threadIdx.y * 32 + threadIdx.x; idx < 3 * (8 + 2 * R) * (32 + 2 * R); idx += 256) {
        const int c = idx / ((8 + 2 * R) * (32 + 2 * R));
        const int rem = idx % ((8 + 2 * R) * (32 + 2 * R));
        const int ty = rem / (32 + 2 * R), tx = rem % (32 + 2 * R);
        int yy = y0 + ty, xx = x0 + tx;
        // reflect (no edge repeat): -1 -> 1, H -> H-2
        if (yy < 0) yy = -yy;
        if (yy >= H) yy = 2 * H - 2 - yy;
        if (xx < 0) xx = -xx;
        if (xx >= W) xx = 2 * W - 2 - xx;
        yy = min(max(yy, 0), H - 1);
        xx = min(max(xx, 0), W - 1);
        tile[c][ty][tx] = src[c * plane + (size_t)yy * W + xx];
    }
    __syncthreads();
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t pix = (size_t)y * W + x;
    float z[3] = {0.f, 0.f, 0.f};
    if (a.noise_std > 0.f) normal3(a.noise_seed, (uint32_t)view, (uint32_t)pix, z);
    float* base = input + ((size_t)(view >> 1) * 6 + (view & 1) * 3) * plane;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < KS; ++i)
#pragma unroll
            for (int j = 0; j < KS; ++j)
                acc = fmaf(__fmul_rn(k1d[i], k1d[j]), tile[c][threadIdx.y + i][threadIdx.x + j], acc);
        if (a.noise_std > 0.f) acc = fmaf(z[c], a.noise_std, acc);
        base[c * plane + pix] = fminf(fmaxf(acc, 0.f), 1.f);
    }
}

}  // namespace sdn
