// GPU side of the reference's per-sample input pipeline (irc:1132-1177, :803-863) after the image file has been decoded:
// cv2.resize(INTER_AREA) of the 8-bit frame, BGR->RGB, /255, clip, paired horizontal flip, [-1, 1] mapping - batched, so that
// eight GPUs are not fed by four DataLoader workers resizing on the host (irc:107).
//
// The arithmetic follows OpenCV's implementation operation for operation (see oracle/input_pipeline.py for the restated
// algorithm and its pinning against cv2 and the reference loaders): every float product and sum below is an explicit
// round-to-nearest intrinsic, so neither --use_fast_math nor FMA contraction can change a bit of the result.
#include "irc_common.cuh"
#include "../../include/irc_b200.h"

namespace {

// mode 0: separable float tables (general scale factors)   dst = cvRound( sum_j yw_j * ( sum_k S[yi_j][xi_k] * xw_k ) )
// mode 1: integer factors                                    dst = cvRound( float(sum) * float(1 / area) )
// mode 2: 2 x 2                                              dst = (a + b + c + d + 2) >> 2
// One thread = one destination pixel, all channels (C <= 4); threads of a warp walk consecutive dx.
__global__ void __launch_bounds__(256) resize_area_u8_kernel(const unsigned char* __restrict__ src, int Hs, int Ws, int C, int Hd, int Wd,
                                                             const int* __restrict__ xi, const float* __restrict__ xw, int kx,
                                                             const int* __restrict__ yi, const float* __restrict__ yw, int ky, int mode, float inv_area,
                                                             unsigned char* __restrict__ dst, int* __restrict__ img_max) {
    irc::pdl_prologue();
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y, n = blockIdx.z;
    int vmax = 0;
    if (dx < Wd) {
        const unsigned char* S = src + (long long)n * Hs * Ws * C;
        int q[4] = {0, 0, 0, 0};
        if (mode == 0) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int j = 0; j < ky; ++j) {
                const float beta = __ldg(yw + dy * ky + j);
                const unsigned char* row = S + (long long)__ldg(yi + dy * ky + j) * Ws * C;
                float buf[4] = {0.f, 0.f, 0.f, 0.f};
                for (int k = 0; k < kx; ++k) {
                    const float alpha = __ldg(xw + dx * kx + k);
                    const unsigned char* px = row + (long long)__ldg(xi + dx * kx + k) * C;
                    for (int c = 0; c < C; ++c) buf[c] = __fadd_rn(buf[c], __fmul_rn((float)px[c], alpha));
                }
                for (int c = 0; c < C; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(beta, buf[c]));
            }
            for (int c = 0; c < C; ++c) q[c] = min(max(__float2int_rn(acc[c]), 0), 255);
        } else {
            const int fy = Hs / Hd, fx = Ws / Wd;
            int acc[4] = {0, 0, 0, 0};
            for (int a = 0; a < fy; ++a) {
                const unsigned char* row = S + ((long long)(dy * fy + a) * Ws + (long long)dx * fx) * C;
                for (int b = 0; b < fx; ++b)
                    for (int c = 0; c < C; ++c) acc[c] += row[b * C + c];
            }
            for (int c = 0; c < C; ++c)
                q[c] = mode == 2 ? ((acc[c] + 2) >> 2) : min(max(__float2int_rn(__fmul_rn((float)acc[c], inv_area)), 0), 255);
        }
        unsigned char* D = dst + (((long long)n * Hd + dy) * Wd + dx) * C;
        for (int c = 0; c < C; ++c) { D[c] = (unsigned char)q[c]; vmax = max(vmax, q[c]); }
    }
    if (img_max) {
        // integer maximum: order-independent, so the atomic is deterministic
        for (int o = 16; o > 0; o >>= 1) vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        if ((threadIdx.x & 31) == 0 && vmax > 0) atomicMax(img_max + n, vmax);
    }
}

// uint8 NHWC (C = 1 or 3) -> fp32 NCHW in [-1, 1]: x/255 (IEEE division; skipped when img_max[n] <= 1, irc:1142), clip, optional
// channel swap (BGR frames as cv2.imread delivers them), optional per-image horizontal flip, then x * 2 - 1 (irc:1174-1175)
__global__ void __launch_bounds__(256) u8_to_pm1_kernel(const unsigned char* __restrict__ src, int H, int W, int C, int swap_rb,
                                                        const unsigned char* __restrict__ flip, const int* __restrict__ img_max, float* __restrict__ out) {
    irc::pdl_prologue();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, n = blockIdx.z;
    if (x >= W) return;
    const bool f = flip && flip[n];
    const bool scale = !img_max || img_max[n] > 1;
    const unsigned char* px = src + (((long long)n * H + y) * W + (f ? W - 1 - x : x)) * C;
    for (int c = 0; c < C; ++c) {
        const int cs = (swap_rb && C == 3) ? 2 - c : c;
        float v = (float)px[cs];
        if (scale) v = __fdiv_rn(v, 255.0f);
        v = fminf(fmaxf(v, 0.f), 1.f);
        out[(((long long)n * C + c) * H + y) * W + x] = __fsub_rn(__fmul_rn(v, 2.0f), 1.0f);
    }
}

}  // namespace

extern "C" int irc_resize_area_u8(const unsigned char* src, int n_img, int Hs, int Ws, int C, int Hd, int Wd, const int* xi, const float* xw, int kx,
                                  const int* yi, const float* yw, int ky, int mode, unsigned char* dst, int* img_max, void* stream) {
    if (!src || !dst || n_img <= 0 || C < 1 || C > 4 || Hd <= 0 || Wd <= 0 || Hd > Hs || Wd > Ws || n_img > 65535 || Hd > 65535)
        return irc_set_error(IRC_ERR_BAD_ARG, "irc_resize_area_u8: bad extents (INTER_AREA shrinking of 1..4-channel 8-bit frames only)");
    if (mode == 0 && (!xi || !xw || !yi || !yw || kx < 1 || ky < 1)) return irc_set_error(IRC_ERR_BAD_ARG, "irc_resize_area_u8: tables required in mode 0");
    if (mode != 0 && (Hs % Hd || Ws % Wd || (mode == 2 && (Hs != 2 * Hd || Ws != 2 * Wd)) || (mode != 1 && mode != 2)))
        return irc_set_error(IRC_ERR_BAD_ARG, "irc_resize_area_u8: mode 1 needs integer factors, mode 2 exactly 2 x 2");
    if (img_max) cudaMemsetAsync(img_max, 0, sizeof(int) * n_img, (cudaStream_t)stream);
    const float inv_area = mode == 1 ? 1.0f / (float)((Hs / Hd) * (Ws / Wd)) : 0.f;
    const int tb = Wd >= 256 ? 256 : ((Wd + 31) / 32) * 32;
    irc::launch(resize_area_u8_kernel, dim3((Wd + tb - 1) / tb, Hd, n_img), tb, 0, (cudaStream_t)stream, src, Hs, Ws, C, Hd, Wd, xi, xw, kx, yi, yw, ky, mode,
                inv_area, dst, img_max);
    return irc_check_launch("irc_resize_area_u8");
}

extern "C" int irc_u8_to_pm1(const unsigned char* src, int n_img, int H, int W, int C, int swap_rb, const unsigned char* flip, const int* img_max, float* out,
                             void* stream) {
    if (!src || !out || n_img <= 0 || C < 1 || C > 4 || n_img > 65535 || H > 65535) return irc_set_error(IRC_ERR_BAD_ARG, "irc_u8_to_pm1: bad args");
    const int tb = W >= 256 ? 256 : ((W + 31) / 32) * 32;
    irc::launch(u8_to_pm1_kernel, dim3((W + tb - 1) / tb, H, n_img), tb, 0, (cudaStream_t)stream, src, H, W, C, swap_rb, flip, img_max, out);
    return irc_check_launch("irc_u8_to_pm1");
}
