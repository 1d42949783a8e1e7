// Separable table stencil on fp32 NCHW planes: the stand-alone Downsample (irc:269-310) and
// UpsampleAA (irc:313-355) modules and their backward passes (transposed tables).
//   out[p][Y][X] (+)= sum_i sum_j wy[Y][i] * wx[X][j] * in[p][iy[Y][i]][ix[X][j]]
// One block = one output row segment of one plane; the (at most 3..6) source rows it needs are
// staged through shared memory with coalesced 128-bit loads, so every input byte is read from
// DRAM once and the horizontal taps come out of shared memory.
#include "irc_common.cuh"
#include "../../include/irc_b200.h"

using namespace irc;

namespace {

constexpr int kMaxK = 8;

struct StP {
    const float* in; float* out;
    int planes, Hi, Wi, Ho, Wo;
    const int* ty_idx; const float* ty_w; int ky;
    const int* tx_idx; const float* tx_w; int kx;
    int accumulate;
};

// block: 256 threads; grid: (x segments, Ho, planes)
__global__ void __launch_bounds__(256) stencil_kernel(const StP p) {
    extern __shared__ float rows[];            // [ky][Wi]
    const int Y = blockIdx.y;
    const long long plane = blockIdx.z;
    const float* src = p.in + plane * (long long)p.Hi * p.Wi;
    int iy[kMaxK]; float wy[kMaxK];
#pragma unroll
    for (int i = 0; i < kMaxK; ++i) {
        iy[i] = i < p.ky ? __ldg(p.ty_idx + Y * p.ky + i) : 0;
        wy[i] = i < p.ky ? __ldg(p.ty_w + Y * p.ky + i) : 0.f;
    }
    // vertical pass into shared memory: v[x] = sum_i wy_i * in[iy_i][x]   (coalesced over x)
    for (int x = threadIdx.x; x < p.Wi; x += blockDim.x) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxK; ++i)
            if (wy[i] != 0.f) a += wy[i] * __ldg(src + (long long)iy[i] * p.Wi + x);
        rows[x] = a;
    }
    __syncthreads();
    float* dst = p.out + (plane * p.Ho + Y) * (long long)p.Wo;
    for (int X = threadIdx.x; X < p.Wo; X += blockDim.x) {
        float a = 0.f;
        for (int j = 0; j < p.kx; ++j) {
            const float w = __ldg(p.tx_w + X * p.kx + j);
            if (w != 0.f) a += w * rows[__ldg(p.tx_idx + X * p.kx + j)];
        }
        if (p.accumulate) dst[X] += a; else dst[X] = a;
    }
}

}  // namespace

extern "C" int irc_stencil_nchw(const float* in, float* out, int planes, int Hi, int Wi, int Ho, int Wo, const int* ty_idx, const float* ty_w,
                                int ky, const int* tx_idx, const float* tx_w, int kx, int accumulate, void* stream) {
    if (!in || !out || !ty_idx || !ty_w || !tx_idx || !tx_w) return irc_set_error(IRC_ERR_BAD_ARG, "irc_stencil_nchw: null");
    if (ky > kMaxK || ky < 1 || kx < 1) return irc_set_error(IRC_ERR_BAD_ARG, "irc_stencil_nchw: 1 <= ky <= 8 required");
    if (planes > 65535 * 64 || Ho > 65535) return irc_set_error(IRC_ERR_BAD_ARG, "irc_stencil_nchw: extent too large");
    if ((size_t)Wi * sizeof(float) > 160 * 1024) return irc_set_error(IRC_ERR_BAD_ARG, "irc_stencil_nchw: row too wide for shared memory");
    StP p;
    p.in = in; p.out = out; p.planes = planes; p.Hi = Hi; p.Wi = Wi; p.Ho = Ho; p.Wo = Wo;
    p.ty_idx = ty_idx; p.ty_w = ty_w; p.ky = ky; p.tx_idx = tx_idx; p.tx_w = tx_w; p.kx = kx; p.accumulate = accumulate;
    const size_t smem = (size_t)Wi * sizeof(float);
    static bool attr = false;
    if (smem > 48 * 1024 && !attr) {
        cudaFuncSetAttribute(stencil_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        attr = true;
    }
    // gridDim.z is limited to 65535: planes beyond that are folded into y by the caller (not needed at our sizes)
    if (planes > 65535) return irc_set_error(IRC_ERR_BAD_ARG, "irc_stencil_nchw: more than 65535 planes");
    stencil_kernel<<<dim3(1, Ho, planes), 256, smem, (cudaStream_t)stream>>>(p);
    return irc_check_launch("irc_stencil_nchw");
}
