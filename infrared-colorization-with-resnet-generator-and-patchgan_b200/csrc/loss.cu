// Loss kernels on fp32 NCHW images: fused L1 + TV (value and gradient in one pass),
// Gaussian-window SSIM forward/backward as shared-memory separable stencils, hinge / GAN
// terms, VGG feature L1, and the test-mode quantise + MAE/MSE pass.
#include <stdlib.h>

#include "irc_common.cuh"
#include "../../include/irc_b200.h"

using namespace irc;

namespace {

__device__ __forceinline__ float sgnf(float v) { return (v > 0.f) - (v < 0.f); }

// ---------------------------------------------------------------- L1 + TV
// sums[0] += sum|f-r|, sums[1] += sum|f[y+1]-f[y]|, sums[2] += sum|f[x+1]-f[x]|
// dfake = w_l1*sign(f-r) + w_tvv*d(TV_v) + w_tvh*d(TV_h)      (irc:686-694, irc:1664, :1672)
__global__ void pixel_loss_kernel(const float* __restrict__ f, const float* __restrict__ r, long long planes, int H, int W,
                                  float w_l1, float w_tvv, float w_tvh, float* sums, float* dfake) {
    irc::pdl_prologue();
    __shared__ float sh[32];
    const long long total = planes * H * W;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const int y = (int)((i / W) % H);
        const float v = f[i];
        float g = 0.f;
        if (r) { const float d = v - r[i]; a0 += fabsf(d); g += w_l1 * sgnf(d); }
        if (y + 1 < H) { const float d = f[i + W] - v; a1 += fabsf(d); g -= w_tvv * sgnf(d); }
        if (y > 0) g += w_tvv * sgnf(v - f[i - W]);
        if (x + 1 < W) { const float d = f[i + 1] - v; a2 += fabsf(d); g -= w_tvh * sgnf(d); }
        if (x > 0) g += w_tvh * sgnf(v - f[i - 1]);
        if (dfake) dfake[i] = g;
    }
    a0 = block_sum(a0, sh); if (threadIdx.x == 0) atomicAdd(sums + 0, a0);
    a1 = block_sum(a1, sh); if (threadIdx.x == 0) atomicAdd(sums + 1, a1);
    a2 = block_sum(a2, sh); if (threadIdx.x == 0) atomicAdd(sums + 2, a2);
}

// Vectorised version (W % 4 == 0, 16-byte aligned planes): one thread = 4 consecutive pixels of a row; the rows above and
// below come in as float4 (L1/L2 hits: every row is read by three row-neighbours), the two horizontal neighbours as scalars.
__global__ void __launch_bounds__(256) pixel_loss_vec_kernel(const float* __restrict__ f, const float* __restrict__ r, long long planes, int H, int W,
                                                              float w_l1, float w_tvv, float w_tvh, float* sums, float* dfake) {
    irc::pdl_prologue();
    __shared__ float sh[32];
    const int W4 = W >> 2;
    const long long total = planes * H * W4;
    const float4* f4 = reinterpret_cast<const float4*>(f);
    const float4* r4 = reinterpret_cast<const float4*>(r);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < total; v += (long long)gridDim.x * blockDim.x) {
        // 32-bit index arithmetic (the host checks total < 2^31): a 64-bit division costs ~100 instructions per vector
        const unsigned row = (unsigned)v / (unsigned)W4;
        const int x4 = (int)((unsigned)v - row * (unsigned)W4);
        const int y = (int)(row % (unsigned)H);
        const float4 c4 = __ldg(f4 + v);
        const float c[4] = {c4.x, c4.y, c4.z, c4.w};
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        if (r) {
            const float4 t4 = __ldg(r4 + v);
            const float t[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float d = c[k] - t[k]; a0 += fabsf(d); g[k] += w_l1 * sgnf(d); }
        }
        if (y + 1 < H) {
            const float4 d4 = __ldg(f4 + v + W4);
            const float dn[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float d = dn[k] - c[k]; a1 += fabsf(d); g[k] -= w_tvv * sgnf(d); }
        }
        if (y > 0) {
            const float4 u4 = __ldg(f4 + v - W4);
            const float up[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) g[k] += w_tvv * sgnf(c[k] - up[k]);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) { const float d = c[k + 1] - c[k]; a2 += fabsf(d); const float sg = w_tvh * sgnf(d); g[k] -= sg; g[k + 1] += sg; }
        if (x4 + 1 < W4) { const float d = __ldg(f + v * 4 + 4) - c[3]; a2 += fabsf(d); g[3] -= w_tvh * sgnf(d); }
        if (x4 > 0) g[0] += w_tvh * sgnf(c[0] - __ldg(f + v * 4 - 1));
        if (dfake) reinterpret_cast<float4*>(dfake)[v] = make_float4(g[0], g[1], g[2], g[3]);
    }
    a0 = block_sum(a0, sh); if (threadIdx.x == 0) atomicAdd(sums + 0, a0);
    a1 = block_sum(a1, sh); if (threadIdx.x == 0) atomicAdd(sums + 1, a1);
    a2 = block_sum(a2, sh); if (threadIdx.x == 0) atomicAdd(sums + 2, a2);
}

// Row-blocked version (W % 4 == 0): one thread = a 4-pixel vector of FOUR consecutive rows.  The six rows it needs (one above,
// one below) and the four target rows are ten independent 16-byte loads in flight per thread, the rows above / below are
// fetched 1.5x per output instead of 3x, and the horizontal neighbours come from the adjacent lanes by shuffle (first / last
// lane of a warp: one scalar load).  (A warp-per-column-tile streaming walk with a rolling 3-row window was also measured:
// slower - one row of prefetch per warp is too little memory-level parallelism.)
constexpr int kPLRows = 4;

__device__ __forceinline__ float wsgn(float w, float e) { return e != 0.f ? copysignf(w, e) : 0.f; }      // w * sign(e), sign(0) = 0

// One 4-pixel x kPLRows block.  kEdge = false: the block and its two halo rows lie inside the plane and the whole warp is
// live, so no bounds predicate is evaluated; kEdge = true: first / last rows of a plane and the ragged last warp of a row.
// The kernel is issue-bound (ncu: DRAM bytes == algorithmic, 54 % issue utilisation at 2 resident blocks per SM), so every
// difference is formed ONCE and its signed weight goes to both pixels it couples, and addresses are 32-bit element offsets.
template <bool kEdge>
__device__ __forceinline__ void pixel_loss_block(const float* __restrict__ fs, const float* __restrict__ rs, float* __restrict__ gs, int H, int W, int W4,
                                                 int x4, int y0, bool live, int lane, float w_l1, float w_tvv, float w_tvh, float& a0, float& a1, float& a2) {
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* fp = reinterpret_cast<const float4*>(fs) + x4;
    const float4* rp = reinterpret_cast<const float4*>(rs) + x4;
    float4 row[kPLRows + 2], tg[kPLRows];
    const int o0 = (y0 - 1) * W4;
#pragma unroll
    for (int j = 0; j < kPLRows + 2; ++j) {
        const int y = y0 - 1 + j;
        row[j] = (!kEdge || (live && y >= 0 && y < H)) ? __ldg(fp + o0 + j * W4) : z4;
    }
#pragma unroll
    for (int j = 0; j < kPLRows; ++j) tg[j] = (rs && (!kEdge || (live && y0 + j < H))) ? __ldg(rp + o0 + (j + 1) * W4) : z4;
    float g[kPLRows][4];
#pragma unroll
    for (int j = 0; j < kPLRows; ++j) { g[j][0] = g[j][1] = g[j][2] = g[j][3] = 0.f; }
    // vertical pairs (row y0 - 1 + j, row y0 + j), j = 0 .. kPLRows; a pair is counted by the block that owns its upper row
#pragma unroll
    for (int j = 0; j <= kPLRows; ++j) {
        const int ya = y0 - 1 + j;
        if (kEdge && (ya < 0 || ya + 1 >= H || !live)) continue;
        const float up_[4] = {row[j].x, row[j].y, row[j].z, row[j].w}, dn_[4] = {row[j + 1].x, row[j + 1].y, row[j + 1].z, row[j + 1].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float e = dn_[k] - up_[k];
            const float sg = wsgn(w_tvv, e);
            if (j >= 1) { a1 += fabsf(e); g[j - 1][k] -= sg; }
            if (j < kPLRows) g[j][k] += sg;
        }
    }
#pragma unroll
    for (int j = 0; j < kPLRows; ++j) {
        const int y = y0 + j;
        const float4 cur = row[j + 1];
        float lf = __shfl_up_sync(0xffffffffu, cur.w, 1), rt = __shfl_down_sync(0xffffffffu, cur.x, 1);
        if (kEdge && (!live || y >= H)) continue;
        const bool has_l = x4 > 0, has_r = x4 + 1 < W4;
        if (lane == 0 && has_l) lf = __ldg(fs + y * W + x4 * 4 - 1);
        if (lane == 31 && has_r) rt = __ldg(fs + y * W + x4 * 4 + 4);
        const float c[4] = {cur.x, cur.y, cur.z, cur.w}, t[4] = {tg[j].x, tg[j].y, tg[j].z, tg[j].w};
        if (rs) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float e = c[k] - t[k]; a0 += fabsf(e); g[j][k] += wsgn(w_l1, e); }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) { const float e = c[k + 1] - c[k]; a2 += fabsf(e); const float sg = wsgn(w_tvh, e); g[j][k] -= sg; g[j][k + 1] += sg; }
        if (has_r) { const float e = rt - c[3]; a2 += fabsf(e); g[j][3] -= wsgn(w_tvh, e); }
        if (has_l) g[j][0] += wsgn(w_tvh, c[0] - lf);
        if (gs) reinterpret_cast<float4*>(gs)[y * W4 + x4] = make_float4(g[j][0], g[j][1], g[j][2], g[j][3]);
    }
}

// Row-blocked L1 + TV (W % 4 == 0, plane < 2^31 elements): one thread = a 4-pixel vector of FOUR consecutive rows.  The six rows
// it needs and the four target rows are ten independent 16-byte loads in flight per thread, the rows above / below are fetched
// 1.5x per output instead of 3x, and the horizontal neighbours come from the adjacent lanes by shuffle (first / last lane of a
// warp: one scalar load).  (A warp-per-column-tile streaming walk with a rolling 3-row window was also measured: slower - one
// row of prefetch per warp is too little memory-level parallelism.)
__global__ void __launch_bounds__(256) pixel_loss_rows_kernel(const float* __restrict__ f, const float* __restrict__ r, int planes, int H, int W,
                                                               float w_l1, float w_tvv, float w_tvh, float* sums, float* dfake) {
    irc::pdl_prologue();
    __shared__ float sh[32];
    const int W4 = W >> 2, lane = threadIdx.x & 31;
    const int W4p = (W4 + 31) & ~31;                        // whole warps per row block: shuffles need converged lanes
    const int yblocks = (H + kPLRows - 1) / kPLRows;
    const unsigned per_plane = (unsigned)yblocks * (unsigned)W4p;
    const long long total = (long long)planes * per_plane;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < total; v += (long long)gridDim.x * blockDim.x) {
        const int plane = (int)(v / per_plane);
        const unsigned in_plane = (unsigned)(v - (long long)plane * per_plane);
        const int yb = (int)(in_plane / (unsigned)W4p), x4 = (int)(in_plane - (unsigned)yb * W4p), y0 = yb * kPLRows;
        const long long po = (long long)plane * H * W;
        const float* fs = f + po;
        const float* rs = r ? r + po : nullptr;
        float* gs = dfake ? dfake + po : nullptr;
        const bool warp_full = (x4 | 31) < W4;              // uniform per warp: every lane of the warp has a vector inside the row
        if (y0 > 0 && y0 + kPLRows < H && warp_full)
            pixel_loss_block<false>(fs, rs, gs, H, W, W4, x4, y0, true, lane, w_l1, w_tvv, w_tvh, a0, a1, a2);
        else
            pixel_loss_block<true>(fs, rs, gs, H, W, W4, x4, y0, x4 < W4, lane, w_l1, w_tvv, w_tvh, a0, a1, a2);
    }
    a0 = block_sum(a0, sh); if (threadIdx.x == 0) atomicAdd(sums + 0, a0);
    a1 = block_sum(a1, sh); if (threadIdx.x == 0) atomicAdd(sums + 1, a1);
    a2 = block_sum(a2, sh); if (threadIdx.x == 0) atomicAdd(sums + 2, a2);
}

// ---------------------------------------------------------------- SSIM (irc:714-750)
constexpr int TW = 32, TH = 16, R = 5, K = 11;
constexpr int LW = TW + 2 * R, LH = TH + 2 * R;

struct Win { float g[K]; };

// pass 1: per-pixel SSIM value summed per image + the three sensitivity maps
__global__ void __launch_bounds__(256)
ssim_fwd_kernel(const float* __restrict__ img1, const float* __restrict__ img2, int C, int H, int W, float scale, float shift,
                Win win, float* sums, float* Ga, float* Gb, float* Gc) {
    irc::pdl_prologue();
    __shared__ float xs[LH][LW], ys[LH][LW];
    __shared__ float hs[5][LH][TW];
    __shared__ float red[32];
    const int plane = blockIdx.z;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const float* p1 = img1 + (long long)plane * H * W;
    const float* p2 = img2 + (long long)plane * H * W;
    for (int i = threadIdx.x; i < LH * LW; i += blockDim.x) {
        const int ly = i / LW, lx = i % LW;
        const int y = y0 + ly - R, x = x0 + lx - R;
        float a = 0.f, b = 0.f;
        if (y >= 0 && y < H && x >= 0 && x < W) { a = p1[(long long)y * W + x] * scale + shift; b = p2[(long long)y * W + x] * scale + shift; }
        xs[ly][lx] = a; ys[ly][lx] = b;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LH * TW; i += blockDim.x) {
        const int ly = i / TW, lx = i % TW;
        float m1 = 0, m2 = 0, e11 = 0, e22 = 0, e12 = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float a = xs[ly][lx + k], b = ys[ly][lx + k], w = win.g[k];
            m1 += w * a; m2 += w * b; e11 += w * a * a; e22 += w * b * b; e12 += w * a * b;
        }
        hs[0][ly][lx] = m1; hs[1][ly][lx] = m2; hs[2][ly][lx] = e11; hs[3][ly][lx] = e22; hs[4][ly][lx] = e12;
    }
    __syncthreads();
    float local = 0.f;
    for (int i = threadIdx.x; i < TH * TW; i += blockDim.x) {
        const int ly = i / TW, lx = i % TW;
        const int y = y0 + ly, x = x0 + lx;
        if (y >= H || x >= W) continue;
        float m1 = 0, m2 = 0, e11 = 0, e22 = 0, e12 = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float w = win.g[k];
            m1 += w * hs[0][ly + k][lx]; m2 += w * hs[1][ly + k][lx]; e11 += w * hs[2][ly + k][lx];
            e22 += w * hs[3][ly + k][lx]; e12 += w * hs[4][ly + k][lx];
        }
        const float C1 = 1e-4f, C2 = 9e-4f;
        const float s1 = e11 - m1 * m1, s2 = e22 - m2 * m2, s12 = e12 - m1 * m2;
        const float A1 = 2.f * m1 * m2 + C1, A2 = 2.f * s12 + C2, B1 = m1 * m1 + m2 * m2 + C1, B2 = s1 + s2 + C2;
        const float inv = __fdiv_rn(1.f, B1 * B2);
        const float S = A1 * A2 * inv;
        local += S;
        if (Ga) {
            const long long o = (long long)plane * H * W + (long long)y * W + x;
            const float dS_de11 = -__fdiv_rn(S, B2);
            const float dS_de12 = 2.f * A1 * inv;
            Ga[o] = 2.f * m2 * A2 * inv - 2.f * m1 * __fdiv_rn(S, B1) - 2.f * m1 * dS_de11 - m2 * dS_de12;
            Gb[o] = dS_de11;
            Gc[o] = dS_de12;
        }
    }
    local = block_sum(local, red);
    if (threadIdx.x == 0) atomicAdd(sums + plane / C, local);
}

// pass 2: dimg1 (+)= coef * ( w*Ga + 2 x (w*Gb) + y (w*Gc) )
__global__ void __launch_bounds__(256)
ssim_bwd_kernel(const float* __restrict__ img1, const float* __restrict__ img2, int H, int W, float scale, float shift, Win win,
                const float* __restrict__ Ga, const float* __restrict__ Gb, const float* __restrict__ Gc, float coef, float* dimg, int accumulate) {
    irc::pdl_prologue();
    __shared__ float gs[3][LH][LW];
    __shared__ float hs[3][LH][TW];
    const int plane = blockIdx.z;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const long long base = (long long)plane * H * W;
    for (int i = threadIdx.x; i < LH * LW; i += blockDim.x) {
        const int ly = i / LW, lx = i % LW;
        const int y = y0 + ly - R, x = x0 + lx - R;
        float a = 0.f, b = 0.f, c = 0.f;
        if (y >= 0 && y < H && x >= 0 && x < W) { const long long o = base + (long long)y * W + x; a = Ga[o]; b = Gb[o]; c = Gc[o]; }
        gs[0][ly][lx] = a; gs[1][ly][lx] = b; gs[2][ly][lx] = c;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LH * TW; i += blockDim.x) {
        const int ly = i / TW, lx = i % TW;
        float a = 0, b = 0, c = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) { const float w = win.g[k]; a += w * gs[0][ly][lx + k]; b += w * gs[1][ly][lx + k]; c += w * gs[2][ly][lx + k]; }
        hs[0][ly][lx] = a; hs[1][ly][lx] = b; hs[2][ly][lx] = c;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TH * TW; i += blockDim.x) {
        const int ly = i / TW, lx = i % TW;
        const int y = y0 + ly, x = x0 + lx;
        if (y >= H || x >= W) continue;
        float a = 0, b = 0, c = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) { const float w = win.g[k]; a += w * hs[0][ly + k][lx]; b += w * hs[1][ly + k][lx]; c += w * hs[2][ly + k][lx]; }
        const long long o = base + (long long)y * W + x;
        const float xv = img1[o] * scale + shift, yv = img2[o] * scale + shift;
        const float g = coef * scale * (a + 2.f * xv * b + yv * c);
        if (accumulate) dimg[o] += g; else dimg[o] = g;
    }
}

// ---------------------------------------------------------------- SSIM, streaming form
// The tiled kernels above spend one shared-memory load per multiply-add (the vertical pass reads its 55 operands from shared
// memory) and filter 26 rows for every 16 they output: they are issue-bound at ~5 TFMA/s.  Here a thread owns ONE image column
// and walks down a strip of rows: each input row is staged once per block in shared memory (double-buffered, one barrier per
// row), the thread filters it horizontally (11 taps read from shared memory, products formed in registers) and pushes the five
// row-filtered values into an 11-deep ring held in REGISTERS (the row loop is unrolled by 11, so every ring slot is a fixed
// register); the vertical pass is then 55 multiply-adds on registers.  ~1.7x fewer issue slots per output pixel.
constexpr int SW = 128;                  // columns per block = threads per block
constexpr int SLW = SW + 2 * R;

template <bool kMaps>
__global__ void __launch_bounds__(SW)
ssim_fwd_stream_kernel(const float* __restrict__ img1, const float* __restrict__ img2, int C, int H, int W, int strip, float scale, float shift,
                       Win win, float* sums, float* Ga, float* Gb, float* Gc) {
    irc::pdl_prologue();
    __shared__ float ra[2][SLW], rb[2][SLW];
    __shared__ float red[32];
    const int plane = blockIdx.z, tid = threadIdx.x;
    const int x0 = blockIdx.x * SW, y0 = blockIdx.y * strip, y1 = min(y0 + strip, H);
    const int x = x0 + tid;
    const long long base = (long long)plane * H * W;
    const float* p1 = img1 + base;
    const float* p2 = img2 + base;
    const int nrows = y1 - y0 + 2 * R;                    // input rows y0 - R .. y1 - 1 + R
    float ring[5][K];
    float local = 0.f;
    for (int rbase = 0; rbase < nrows; rbase += K) {
#pragma unroll
        for (int ph = 0; ph < K; ++ph) {
            const int r = rbase + ph;
            if (r >= nrows) break;
            const int iy = y0 - R + r, bsel = r & 1;
            {   // stage input row iy (scaled; zero outside the image: the reference zero-pads the scaled images, irc:727)
                const bool yok = iy >= 0 && iy < H;
                for (int lx = tid; lx < SLW; lx += SW) {
                    const int gx = x0 - R + lx;
                    float a = 0.f, b = 0.f;
                    if (yok && gx >= 0 && gx < W) { a = __ldg(p1 + (long long)iy * W + gx) * scale + shift; b = __ldg(p2 + (long long)iy * W + gx) * scale + shift; }
                    ra[bsel][lx] = a; rb[bsel][lx] = b;
                }
            }
            __syncthreads();
            float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float a = ra[bsel][tid + k], b = rb[bsel][tid + k], w = win.g[k];
                const float wa = w * a, wb = w * b;
                m1 += wa; m2 += wb; e11 = fmaf(wa, a, e11); e22 = fmaf(wb, b, e22); e12 = fmaf(wa, b, e12);
            }
            ring[0][ph] = m1; ring[1][ph] = m2; ring[2][ph] = e11; ring[3][ph] = e22; ring[4][ph] = e12;
            if (r >= 2 * R) {
                const int y = y0 + r - 2 * R;
                float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float w = win.g[k];
#pragma unroll
                    for (int m = 0; m < 5; ++m) v[m] = fmaf(w, ring[m][(ph + 1 + k) % K], v[m]);
                }
                if (x < W) {
                    const float C1 = 1e-4f, C2 = 9e-4f;
                    const float mm1 = v[0], mm2 = v[1];
                    const float s1 = v[2] - mm1 * mm1, s2 = v[3] - mm2 * mm2, s12 = v[4] - mm1 * mm2;
                    const float A1 = 2.f * mm1 * mm2 + C1, A2 = 2.f * s12 + C2, B1 = mm1 * mm1 + mm2 * mm2 + C1, B2 = s1 + s2 + C2;
                    const float inv = __fdiv_rn(1.f, B1 * B2);
                    const float S = A1 * A2 * inv;
                    local += S;
                    if (kMaps) {
                        const long long o = base + (long long)y * W + x;
                        const float dS_de11 = -__fdiv_rn(S, B2);
                        const float dS_de12 = 2.f * A1 * inv;
                        Ga[o] = 2.f * mm2 * A2 * inv - 2.f * mm1 * __fdiv_rn(S, B1) - 2.f * mm1 * dS_de11 - mm2 * dS_de12;
                        Gb[o] = dS_de11;
                        Gc[o] = dS_de12;
                    }
                }
            }
        }
    }
    local = block_sum(local, red);
    if (threadIdx.x == 0) atomicAdd(sums + plane / C, local);
}

// dimg1 (+)= coef * scale * ( w*Ga + 2 x (w*Gb) + y (w*Gc) ), same streaming structure with three maps
__global__ void __launch_bounds__(SW)
ssim_bwd_stream_kernel(const float* __restrict__ img1, const float* __restrict__ img2, int H, int W, int strip, float scale, float shift, Win win,
                       const float* __restrict__ Ga, const float* __restrict__ Gb, const float* __restrict__ Gc, float coef, float* dimg, int accumulate) {
    irc::pdl_prologue();
    __shared__ float rg[2][3][SLW];
    const int plane = blockIdx.z, tid = threadIdx.x;
    const int x0 = blockIdx.x * SW, y0 = blockIdx.y * strip, y1 = min(y0 + strip, H);
    const int x = x0 + tid;
    const long long base = (long long)plane * H * W;
    const int nrows = y1 - y0 + 2 * R;
    float ring[3][K];
    for (int rbase = 0; rbase < nrows; rbase += K) {
#pragma unroll
        for (int ph = 0; ph < K; ++ph) {
            const int r = rbase + ph;
            if (r >= nrows) break;
            const int iy = y0 - R + r, bsel = r & 1;
            {
                const bool yok = iy >= 0 && iy < H;
                for (int lx = tid; lx < SLW; lx += SW) {
                    const int gx = x0 - R + lx;
                    float a = 0.f, b = 0.f, c = 0.f;
                    if (yok && gx >= 0 && gx < W) { const long long o = base + (long long)iy * W + gx; a = __ldg(Ga + o); b = __ldg(Gb + o); c = __ldg(Gc + o); }
                    rg[bsel][0][lx] = a; rg[bsel][1][lx] = b; rg[bsel][2][lx] = c;
                }
            }
            __syncthreads();
            float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) { const float w = win.g[k]; a = fmaf(w, rg[bsel][0][tid + k], a); b = fmaf(w, rg[bsel][1][tid + k], b); c = fmaf(w, rg[bsel][2][tid + k], c); }
            ring[0][ph] = a; ring[1][ph] = b; ring[2][ph] = c;
            if (r >= 2 * R) {
                const int y = y0 + r - 2 * R;
                float v[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float w = win.g[k];
#pragma unroll
                    for (int m = 0; m < 3; ++m) v[m] = fmaf(w, ring[m][(ph + 1 + k) % K], v[m]);
                }
                if (x < W) {
                    const long long o = base + (long long)y * W + x;
                    const float xv = __ldg(img1 + o) * scale + shift, yv = __ldg(img2 + o) * scale + shift;
                    const float g = coef * scale * (v[0] + 2.f * xv * v[1] + yv * v[2]);
                    if (accumulate) dimg[o] += g; else dimg[o] = g;
                }
            }
        }
    }
}

// rows per strip of the streaming SSIM kernels: long strips amortise the 10 halo rows, short ones fill the machine.  Returns 0
// when even 64-row strips leave SMs idle: small batches (B=16 at 256 x 256) stay on the tiled kernels, which were measured
// faster there (70 vs 92 us forward) because 16-row strips filter 26 rows for every 16 they output.
int ssim_strip(int H, int W, int planes) {
    const long long cols = (long long)planes * ((W + SW - 1) / SW);
    int strip = 128;
    while (strip > 64 && cols * ((H + strip - 1) / strip) < (long long)irc_num_sms() * 6) strip >>= 1;
    return cols * ((H + strip - 1) / strip) < (long long)irc_num_sms() * 6 ? 0 : strip;
}

// ---------------------------------------------------------------- hinge / GAN (irc:1647-1649, :1662)
// mode 0: first n_real entries are D(real), the rest D(fake):
//         sums[0] += relu(1-p), sums[1] += relu(1+p); dpred = -w_real*[1-p>0] | +w_fake*[1+p>0]
// mode 1: sums[2] += p; dpred = -w_real
__global__ void hinge_kernel(const float* __restrict__ pred, long long n_total, long long n_real, int mode, float w_real, float w_fake,
                             float* sums, float* dpred) {
    irc::pdl_prologue();
    __shared__ float sh[32];
    float a = 0.f, b = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_total; i += (long long)gridDim.x * blockDim.x) {
        const float p = pred[i];
        if (mode == 1) { a += p; if (dpred) dpred[i] = -w_real; }
        else if (i < n_real) { const float t = 1.f - p; a += fmaxf(t, 0.f); if (dpred) dpred[i] = t > 0.f ? -w_real : 0.f; }
        else { const float t = 1.f + p; b += fmaxf(t, 0.f); if (dpred) dpred[i] = t > 0.f ? w_fake : 0.f; }
    }
    a = block_sum(a, sh); if (threadIdx.x == 0) atomicAdd(sums + (mode == 1 ? 2 : 0), a);
    if (mode == 0) { b = block_sum(b, sh); if (threadIdx.x == 0) atomicAdd(sums + 1, b); }
}

// ---------------------------------------------------------------- VGG feature L1 (irc:1667-1669)
// feat rows [0,R) = features of fake, rows [R,2R) = features of the target (both post-ReLU)
// dz[q][c] = w * sign(f - r) * [f > 0]
__global__ void feat_l1_kernel(const bf16* __restrict__ feat, long long rows_half, long long ld, int C, float w, float* sums, bf16* dz, long long ld_dz) {
    irc::pdl_prologue();
    __shared__ float sh[32];
    const int C8 = C >> 3;
    const long long total = rows_half * C8;
    float acc = 0.f;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C8) * 8;
        const long long q = idx / C8;
        const uint4 uf = __ldg(reinterpret_cast<const uint4*>(feat + q * ld + c));
        const uint4 ur = __ldg(reinterpret_cast<const uint4*>(feat + (q + rows_half) * ld + c));
        const uint32_t wf[4] = {uf.x, uf.y, uf.z, uf.w}, wr[4] = {ur.x, ur.y, ur.z, ur.w};
        uint32_t o[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const float2 a = unpack_bf16x2(wf[h]), b = unpack_bf16x2(wr[h]);
            const float d0 = a.x - b.x, d1 = a.y - b.y;
            acc += fabsf(d0) + fabsf(d1);
            o[h] = pack_bf16x2(a.x > 0.f ? w * sgnf(d0) : 0.f, a.y > 0.f ? w * sgnf(d1) : 0.f);
        }
        if (dz) *reinterpret_cast<uint4*>(dz + q * ld_dz + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) atomicAdd(sums, acc);
}

// ---------------------------------------------------------------- test-mode core (irc:865-876, irc:1197-1205)
// u8 = trunc(clip((x+1)/2, 0, 1) * 255) in HWC order; sums[n] = (sum |u8/255 - gt|, sum (u8/255 - gt)^2)
__global__ void quantize_metrics_kernel(const float* __restrict__ fake, const float* __restrict__ gt, int C, int H, int W,
                                        unsigned char* u8, double* sums) {
    irc::pdl_prologue();
    __shared__ float sh[32];
    const int n = blockIdx.y;
    const long long hw = (long long)H * W;
    float a = 0.f, b = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw * C; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i / hw);
        const long long pix = i % hw;
        float v = (fake[(long long)n * C * hw + i] + 1.0f) * 0.5f;
        v = fminf(fmaxf(v, 0.f), 1.f);
        const unsigned char q = (unsigned char)(v * 255.0f);
        if (u8) u8[((long long)n * hw + pix) * C + c] = q;
        if (gt) {
            const float d = __fdiv_rn((float)q, 255.0f) - gt[(long long)n * C * hw + i];
            a += fabsf(d); b += d * d;
        }
    }
    if (gt) {
        a = block_sum(a, sh); if (threadIdx.x == 0) atomicAdd(sums + n * 2, (double)a);
        b = block_sum(b, sh); if (threadIdx.x == 0) atomicAdd(sums + n * 2 + 1, (double)b);
    }
}

// acc[j] += coef[j][n] + sum_i coef[j][i] * s[i]   (rows <= 32, one warp per row; fixed summation order)
__global__ void accumulate_kernel(const float* __restrict__ s, int n, const float* __restrict__ coef, int rows, double* __restrict__ acc) {
    irc::pdl_prologue();
    const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (j >= rows) return;
    const float* c = coef + (long long)j * (n + 1);
    double a = 0.0;
    for (int i = lane; i < n; i += 32) a += (double)c[i] * (double)s[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) acc[j] += a + (double)c[n];
}

// C == 3, H*W % 4 == 0: one thread = 4 consecutive pixels, three float4 loads per tensor (coalesced per plane) and 12 contiguous
// output bytes (HWC), no 64-bit divisions
__global__ void __launch_bounds__(256) quantize_metrics3_kernel(const float* __restrict__ fake, const float* __restrict__ gt, long long hw, unsigned char* u8,
                                                                 double* sums) {
    irc::pdl_prologue();
    __shared__ float sh[32];
    const int n = blockIdx.y;
    const long long hw4 = hw >> 2;
    const float4* fp = reinterpret_cast<const float4*>(fake + (long long)n * 3 * hw);
    const float4* gp = gt ? reinterpret_cast<const float4*>(gt + (long long)n * 3 * hw) : nullptr;
    float a = 0.f, b = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw4; i += (long long)gridDim.x * blockDim.x) {
        unsigned char q[4][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float4 v4 = __ldg(fp + c * hw4 + i);
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
            float t[4] = {0.f, 0.f, 0.f, 0.f};
            if (gp) { const float4 g4 = __ldg(gp + c * hw4 + i); t[0] = g4.x; t[1] = g4.y; t[2] = g4.z; t[3] = g4.w; }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float x = (v[k] + 1.0f) * 0.5f;
                x = fminf(fmaxf(x, 0.f), 1.f);
                q[k][c] = (unsigned char)(x * 255.0f);
                if (gp) { const float d = __fdiv_rn((float)q[k][c], 255.0f) - t[k]; a += fabsf(d); b += d * d; }
            }
        }
        if (u8) {
            const unsigned char* qq = &q[0][0];
            uint32_t w[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) w[j] = (uint32_t)qq[4 * j] | ((uint32_t)qq[4 * j + 1] << 8) | ((uint32_t)qq[4 * j + 2] << 16) | ((uint32_t)qq[4 * j + 3] << 24);
            uint32_t* o = reinterpret_cast<uint32_t*>(u8 + ((long long)n * hw + i * 4) * 3);
            o[0] = w[0]; o[1] = w[1]; o[2] = w[2];
        }
    }
    if (gt) {
        a = block_sum(a, sh); if (threadIdx.x == 0) atomicAdd(sums + n * 2, (double)a);
        b = block_sum(b, sh); if (threadIdx.x == 0) atomicAdd(sums + n * 2 + 1, (double)b);
    }
}

// ---------------------------------------------------------------- SSIM metric (irc:1208-1215)
// skimage.metrics.structural_similarity(gt, pred, data_range=1.0, channel_axis=2) with its defaults: 7 x 7 uniform window,
// sample covariance (N / (N - 1)), K1 = 0.01, K2 = 0.03, float64 arithmetic, mean over the map cropped by 3 pixels, mean over
// channels.  Only the cropped (valid) region is evaluated, so the filter's border mode never matters.  pred = u8 / 255 as
// float32 (irc:1413), gt float32 in [0, 1].  sums[n] += sum over channels and valid pixels of S.
constexpr int kSmW = 32, kSmH = 16, kSmWin = 7;
__global__ void __launch_bounds__(256) ssim_metric_kernel(const unsigned char* __restrict__ u8, const float* __restrict__ gt, int H, int W, double* sums) {
    irc::pdl_prologue();
    __shared__ float tx[kSmH + kSmWin - 1][kSmW + kSmWin - 1], ty[kSmH + kSmWin - 1][kSmW + kSmWin - 1];
    __shared__ double hs[5][kSmH + kSmWin - 1][kSmW];
    __shared__ double red[8];
    const int n = blockIdx.z / 3, c = blockIdx.z % 3;
    const int x0 = blockIdx.x * kSmW, y0 = blockIdx.y * kSmH;          // first valid-region output of the tile = input offset
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const long long hw = (long long)H * W;
    for (int i = tid; i < (kSmH + 6) * (kSmW + 6); i += 256) {
        const int r = i / (kSmW + 6), q = i - r * (kSmW + 6);
        const int y = y0 + r, x = x0 + q;
        float a = 0.f, b = 0.f;
        if (y < H && x < W) {
            a = __ldg(gt + ((long long)n * 3 + c) * hw + (long long)y * W + x);
            b = __fdiv_rn((float)u8[((long long)n * hw + (long long)y * W + x) * 3 + c], 255.0f);
        }
        tx[r][q] = a; ty[r][q] = b;
    }
    __syncthreads();
    for (int i = tid; i < (kSmH + 6) * kSmW; i += 256) {
        const int r = i / kSmW, q = i - r * kSmW;
        double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
#pragma unroll
        for (int k = 0; k < kSmWin; ++k) {
            const double a = (double)tx[r][q + k], b = (double)ty[r][q + k];
            sx += a; sy += b; sxx += a * a; syy += b * b; sxy += a * b;
        }
        hs[0][r][q] = sx; hs[1][r][q] = sy; hs[2][r][q] = sxx; hs[3][r][q] = syy; hs[4][r][q] = sxy;
    }
    __syncthreads();
    const double NP = 49.0, cov_norm = NP / (NP - 1.0), C1 = 1e-4, C2 = 9e-4;
    double acc = 0.0;
    for (int rr = threadIdx.y; rr < kSmH; rr += 8) {
        const int oy = y0 + rr, ox = x0 + threadIdx.x;
        if (oy >= H - 6 || ox >= W - 6) continue;
        double s[5] = {0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < kSmWin; ++k)
#pragma unroll
            for (int j = 0; j < 5; ++j) s[j] += hs[j][rr + k][threadIdx.x];
        const double ux = s[0] / NP, uy = s[1] / NP, uxx = s[2] / NP, uyy = s[3] / NP, uxy = s[4] / NP;
        const double vx = cov_norm * (uxx - ux * ux), vy = cov_norm * (uyy - uy * uy), vxy = cov_norm * (uxy - ux * uy);
        acc += ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) red[threadIdx.y] = acc;
    __syncthreads();
    if (tid == 0) {
        double t = 0;
        for (int i = 0; i < 8; ++i) t += red[i];
        atomicAdd(sums + n, t);
    }
}

int grid_for(long long total, int threads, int per_sm) {
    long long b = (total + threads - 1) / threads;
    const long long cap = (long long)irc_num_sms() * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

extern "C" int irc_pixel_loss(const float* fake, const float* target, int n_img, int C, int H, int W, float w_l1, float w_tvv, float w_tvh,
                              float* sums, float* dfake, void* stream) {
    if (!fake || !sums) return irc_set_error(IRC_ERR_BAD_ARG, "irc_pixel_loss: null");
    const long long planes = (long long)n_img * C;
    const bool vec = W % 4 == 0 && !((uintptr_t)fake & 15) && !((uintptr_t)target & 15) && !((uintptr_t)dfake & 15) &&
                     planes * H * (W / 4) < (1LL << 31);
    static int mode = -1;      // IRC_PIXEL_LOSS=vec selects the older one-row kernel (kept for comparison)
    if (mode < 0) { const char* e = getenv("IRC_PIXEL_LOSS"); mode = (e && e[0] == 'v') ? 1 : 0; }
    const long long rb_total = planes * ((H + kPLRows - 1) / kPLRows) * (((W / 4) + 31) & ~31);
    if (vec && mode == 0 && (long long)H * W < (1LL << 31) && planes < (1LL << 31))
        irc::launch(pixel_loss_rows_kernel, grid_for(rb_total, 256, 8), 256, 0, (cudaStream_t)stream, fake, target, (int)planes, H, W, w_l1, w_tvv, w_tvh, sums, dfake);
    else if (vec) irc::launch(pixel_loss_vec_kernel, grid_for(planes * H * (W / 4), 256, 8), 256, 0, (cudaStream_t)stream, fake, target, planes, H, W, w_l1, w_tvv, w_tvh, sums, dfake);
    else irc::launch(pixel_loss_kernel, grid_for(planes * H * W, 256, 8), 256, 0, (cudaStream_t)stream, fake, target, planes, H, W, w_l1, w_tvv, w_tvh, sums, dfake);
    return irc_check_launch("irc_pixel_loss");
}

static int ssim_mode() {      // IRC_SSIM=tiled selects the older shared-memory tiled kernels (kept for comparison)
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("IRC_SSIM"); mode = (e && e[0] == 't') ? 1 : 0; }
    return mode;
}

extern "C" int irc_ssim_fwd(const float* img1, const float* img2, int n_img, int C, int H, int W, float scale, float shift, const float* window11,
                            float* sums, float* ga, float* gb, float* gc, void* stream) {
    if (!img1 || !img2 || !sums || !window11) return irc_set_error(IRC_ERR_BAD_ARG, "irc_ssim_fwd: null");
    Win w; for (int i = 0; i < K; ++i) w.g[i] = window11[i];
    const int strip = ssim_strip(H, W, n_img * C);
    if (ssim_mode() == 0 && n_img * C <= 65535 && strip) {
        dim3 grid((W + SW - 1) / SW, (H + strip - 1) / strip, n_img * C);
        if (ga) irc::launch(ssim_fwd_stream_kernel<true>, grid, SW, 0, (cudaStream_t)stream, img1, img2, C, H, W, strip, scale, shift, w, sums, ga, gb, gc);
        else irc::launch(ssim_fwd_stream_kernel<false>, grid, SW, 0, (cudaStream_t)stream, img1, img2, C, H, W, strip, scale, shift, w, sums, ga, gb, gc);
        return irc_check_launch("irc_ssim_fwd");
    }
    dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, n_img * C);
    irc::launch(ssim_fwd_kernel, grid, 256, 0, (cudaStream_t)stream, img1, img2, C, H, W, scale, shift, w, sums, ga, gb, gc);
    return irc_check_launch("irc_ssim_fwd");
}

extern "C" int irc_ssim_bwd(const float* img1, const float* img2, int n_img, int C, int H, int W, float scale, float shift, const float* window11,
                            const float* ga, const float* gb, const float* gc, float coef, float* dimg1, int accumulate, void* stream) {
    if (!img1 || !img2 || !ga || !gb || !gc || !dimg1) return irc_set_error(IRC_ERR_BAD_ARG, "irc_ssim_bwd: null");
    Win w; for (int i = 0; i < K; ++i) w.g[i] = window11[i];
    const int strip = ssim_strip(H, W, n_img * C);
    if (ssim_mode() == 0 && n_img * C <= 65535 && strip) {
        dim3 grid((W + SW - 1) / SW, (H + strip - 1) / strip, n_img * C);
        irc::launch(ssim_bwd_stream_kernel, grid, SW, 0, (cudaStream_t)stream, img1, img2, H, W, strip, scale, shift, w, ga, gb, gc, coef, dimg1, accumulate);
        return irc_check_launch("irc_ssim_bwd");
    }
    dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, n_img * C);
    irc::launch(ssim_bwd_kernel, grid, 256, 0, (cudaStream_t)stream, img1, img2, H, W, scale, shift, w, ga, gb, gc, coef, dimg1, accumulate);
    return irc_check_launch("irc_ssim_bwd");
}

extern "C" int irc_hinge(const float* pred, long long n_total, long long n_real, int mode, float w_real, float w_fake, float* sums, float* dpred, void* stream) {
    if (!pred || !sums) return irc_set_error(IRC_ERR_BAD_ARG, "irc_hinge: null");
    irc::launch(hinge_kernel, grid_for(n_total, 256, 1), 256, 0, (cudaStream_t)stream, pred, n_total, n_real, mode, w_real, w_fake, sums, dpred);
    return irc_check_launch("irc_hinge");
}

extern "C" int irc_feat_l1(const void* feat, long long rows_half, long long ld, int C, float w, float* sums, void* dz, long long ld_dz, void* stream) {
    if (!feat || !sums || C % 8) return irc_set_error(IRC_ERR_BAD_ARG, "irc_feat_l1: bad args");
    irc::launch(feat_l1_kernel, grid_for(rows_half * (C / 8), 256, 8), 256, 0, (cudaStream_t)stream, (const bf16*)feat, rows_half, ld, C, w, sums, (bf16*)dz, ld_dz);
    return irc_check_launch("irc_feat_l1");
}

extern "C" int irc_quantize_metrics(const float* fake, const float* gt, int n_img, int C, int H, int W, unsigned char* u8, double* sums, void* stream) {
    if (!fake || (gt && !sums)) return irc_set_error(IRC_ERR_BAD_ARG, "irc_quantize_metrics: null");
    if (gt) cudaMemsetAsync(sums, 0, sizeof(double) * 2 * n_img, (cudaStream_t)stream);
    const long long hw = (long long)H * W;
    if (C == 3 && hw % 4 == 0 && !((uintptr_t)fake & 15) && !((uintptr_t)gt & 15) && !((uintptr_t)u8 & 3)) {
        int bx = grid_for(hw / 4, 256, 8);
        if ((long long)bx * n_img < irc_num_sms() * 4) bx = grid_for(hw / 4, 256, 64);
        irc::launch(quantize_metrics3_kernel, dim3(bx, n_img), 256, 0, (cudaStream_t)stream, fake, gt, hw, u8, sums);
        return irc_check_launch("irc_quantize_metrics");
    }
    int bx = grid_for((long long)C * H * W, 256, 4);
    irc::launch(quantize_metrics_kernel, dim3(bx, n_img), 256, 0, (cudaStream_t)stream, fake, gt, C, H, W, u8, sums);
    return irc_check_launch("irc_quantize_metrics");
}

extern "C" int irc_ssim_metric(const unsigned char* u8, const float* gt, int n_img, int H, int W, double* sums, void* stream) {
    if (!u8 || !gt || !sums || n_img <= 0) return irc_set_error(IRC_ERR_BAD_ARG, "irc_ssim_metric: null");
    if (H < kSmWin || W < kSmWin) return irc_set_error(IRC_ERR_BAD_ARG, "irc_ssim_metric: images must be at least 7 x 7 (the window of structural_similarity)");
    if ((long long)n_img * 3 > 65535) return irc_set_error(IRC_ERR_BAD_ARG, "irc_ssim_metric: too many images in one call");
    cudaMemsetAsync(sums, 0, sizeof(double) * n_img, (cudaStream_t)stream);
    irc::launch(ssim_metric_kernel, dim3((W - 6 + kSmW - 1) / kSmW, (H - 6 + kSmH - 1) / kSmH, n_img * 3), dim3(32, 8), 0, (cudaStream_t)stream, u8, gt, H, W, sums);
    return irc_check_launch("irc_ssim_metric");
}

extern "C" int irc_accumulate(const float* s, int n, const float* coef, int rows, double* acc, void* stream) {
    if (!s || !coef || !acc || n <= 0 || rows <= 0 || rows > 32) return irc_set_error(IRC_ERR_BAD_ARG, "irc_accumulate: bad args");
    irc::launch(accumulate_kernel, 1, rows * 32, 0, (cudaStream_t)stream, s, n, coef, rows, acc);
    return irc_check_launch("irc_accumulate");
}
