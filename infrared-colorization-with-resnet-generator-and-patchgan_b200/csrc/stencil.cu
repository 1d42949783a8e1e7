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
    irc::pdl_prologue();
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

// ---------------------------------------------------------------------------------
// Streaming version: one thread = one output column of kPPT planes; it walks down a strip of output rows and keeps the
// last K horizontally filtered source rows in registers (source rows enter in order, each is read once per output
// column, the loads of the next source row are in flight while the current output row is produced).  Lanes run along x,
// so loads and stores are coalesced 128-byte rows; the x-taps of neighbouring lanes overlap in L1.  DRAM traffic =
// input + output once.
// ---------------------------------------------------------------------------------
constexpr int kMaxStrip = 32;
// planes per thread: kPPT x K independent 4-byte loads in flight per thread and source row
template <int K> struct Ppt { static constexpr int v = 4; };        // (8 planes for K <= 3 measured slower: 166 / 246 us vs 158 / 205)

template <int K>
__global__ void __launch_bounds__(256) stencil_stream_kernel(const StP p, int strip, int TX) {
    irc::pdl_prologue();
    constexpr int kPPT = Ppt<K>::v;
    __shared__ int s_hi[kMaxStrip];
    __shared__ int s_lo0;
    __shared__ float s_wd[kMaxStrip][K];
    const int ya = blockIdx.y * strip;
    const int rows = min(strip, p.Ho - ya);
    if ((int)threadIdx.x < rows) {
        const int y = ya + threadIdx.x;
        int lo = 0x7fffffff, hi = -1;
        for (int i = 0; i < p.ky; ++i)
            if (__ldg(p.ty_w + y * p.ky + i) != 0.f) { const int q = __ldg(p.ty_idx + y * p.ky + i); lo = min(lo, q); hi = max(hi, q); }
        s_hi[threadIdx.x] = hi;
        if (threadIdx.x == 0) s_lo0 = lo;
#pragma unroll
        for (int s = 0; s < K; ++s) {
            const int r = hi + 1 - K + s;
            float w = 0.f;
            for (int i = 0; i < p.ky; ++i) {
                const float wi = __ldg(p.ty_w + y * p.ky + i);
                if (wi != 0.f && __ldg(p.ty_idx + y * p.ky + i) == r) w += wi;
            }
            s_wd[threadIdx.x][s] = w;
        }
    }
    __syncthreads();
    const int X = blockIdx.x * TX + (threadIdx.x % TX);
    const long long plane0 = ((long long)blockIdx.z * (blockDim.x / TX) + threadIdx.x / TX) * kPPT;
    if (X >= p.Wo || plane0 >= p.planes) return;
    const int np = (int)min((long long)kPPT, p.planes - plane0);       // planes of this thread (the last group may be ragged)
    float wx[K]; int ix[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const bool h = j < p.kx;
        wx[j] = h ? __ldg(p.tx_w + X * p.kx + j) : 0.f;
        ix[j] = h ? __ldg(p.tx_idx + X * p.kx + j) : 0;
    }
    const long long pin = (long long)p.Hi * p.Wi, pout = (long long)p.Ho * p.Wo;
    const float* src = p.in + plane0 * pin;
    float* dst = p.out + plane0 * pout + X;
    float hb[kPPT][K];
#pragma unroll
    for (int u = 0; u < kPPT; ++u)
#pragma unroll
        for (int s = 0; s < K; ++s) hb[u][s] = 0.f;
    int top = s_lo0;
    const int last = s_hi[rows - 1];
    float raw[kPPT][K];
    auto fetch = [&](int r) {
        const float* row = src + (long long)r * p.Wi;
#pragma unroll
        for (int u = 0; u < kPPT; ++u)
#pragma unroll
            for (int j = 0; j < K; ++j) raw[u][j] = u < np ? __ldg(row + u * pin + ix[j]) : 0.f;
    };
    if (top <= last) fetch(top);
    for (int t = 0; t < rows; ++t) {
        const int hi = s_hi[t];
        while (top <= hi) {
            float h[kPPT];
#pragma unroll
            for (int u = 0; u < kPPT; ++u) {
                h[u] = 0.f;
#pragma unroll
                for (int j = 0; j < K; ++j) h[u] = fmaf(wx[j], raw[u][j], h[u]);
            }
            ++top;
            if (top <= last) fetch(top);
#pragma unroll
            for (int u = 0; u < kPPT; ++u) {
#pragma unroll
                for (int s = 0; s + 1 < K; ++s) hb[u][s] = hb[u][s + 1];
                hb[u][K - 1] = h[u];
            }
        }
        float* d = dst + (long long)(ya + t) * p.Wo;
#pragma unroll
        for (int u = 0; u < kPPT; ++u) {
            float a = 0.f;
#pragma unroll
            for (int s = 0; s < K; ++s) a = fmaf(s_wd[t][s], hb[u][s], a);
            if (u < np) { if (p.accumulate) d[u * pout] += a; else d[u * pout] = a; }
        }
    }
}

}  // namespace

/* Streaming variant of irc_stencil_nchw for tables whose last source row is non-decreasing in the output row and whose
 * non-zero entries of every row span at most `window` <= 8 consecutive source rows (the caller checks its own tables). */
extern "C" int irc_stencil_nchw_stream(const float* in, float* out, int planes, int Hi, int Wi, int Ho, int Wo, const int* ty_idx,
                                       const float* ty_w, int ky, const int* tx_idx, const float* tx_w, int kx, int window, int accumulate,
                                       void* stream) {
    if (!in || !out || !ty_idx || !ty_w || !tx_idx || !tx_w) return irc_set_error(IRC_ERR_BAD_ARG, "irc_stencil_nchw_stream: null");
    if (window < ky || window < kx || window > 8 || ky < 1 || kx < 1) return irc_set_error(IRC_ERR_BAD_ARG, "irc_stencil_nchw_stream: bad window");
    StP p;
    p.in = in; p.out = out; p.planes = planes; p.Hi = Hi; p.Wi = Wi; p.Ho = Ho; p.Wo = Wo;
    p.ty_idx = ty_idx; p.ty_w = ty_w; p.ky = ky; p.tx_idx = tx_idx; p.tx_w = tx_w; p.kx = kx; p.accumulate = accumulate;
    int TX = 32;
    while (TX < 256 && TX < Wo) TX *= 2;
    const int kk = window <= 2 ? 2 : (window <= 3 ? 3 : (window <= 4 ? 4 : (window <= 6 ? 6 : 8)));
    const int ppt = 4;
    const int pb = 256 / TX * ppt;                        // planes per block
    const long long gz = ((long long)planes + pb - 1) / pb;
    const int gx = (Wo + TX - 1) / TX;
    // rows per strip: keep >= ~8 blocks per SM in flight, at most kMaxStrip rows
    long long strip = (long long)gx * gz * Ho / ((long long)irc_num_sms() * 8);
    if (strip > kMaxStrip) strip = kMaxStrip;
    if (strip < 4) strip = 4;
    const int gy = (Ho + (int)strip - 1) / (int)strip;
    if (gz > 65535 || gy > 65535) return irc_set_error(IRC_ERR_BAD_ARG, "irc_stencil_nchw_stream: extent too large");
    const dim3 grid(gx, gy, (unsigned)gz);
    cudaStream_t st = (cudaStream_t)stream;
    if (kk == 2) irc::launch(stencil_stream_kernel<2>, grid, 256, 0, st, p, (int)strip, TX);
    else if (kk == 3) irc::launch(stencil_stream_kernel<3>, grid, 256, 0, st, p, (int)strip, TX);
    else if (kk == 4) irc::launch(stencil_stream_kernel<4>, grid, 256, 0, st, p, (int)strip, TX);
    else if (kk == 6) irc::launch(stencil_stream_kernel<6>, grid, 256, 0, st, p, (int)strip, TX);
    else irc::launch(stencil_stream_kernel<8>, grid, 256, 0, st, p, (int)strip, TX);
    return irc_check_launch("irc_stencil_nchw_stream");
}

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
    irc::launch(stencil_kernel, dim3(1, Ho, planes), 256, smem, (cudaStream_t)stream, p);
    return irc_check_launch("irc_stencil_nchw");
}
