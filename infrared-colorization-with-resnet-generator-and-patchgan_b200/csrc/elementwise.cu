// Memory-bound passes on NHWC bf16 frames: InstanceNorm statistics / apply / backward,
// separable table stencils (anti-aliased Downsample and UpsampleAA, halo folding), all
// vectorised 8 channels (16 bytes) per thread and coalesced along the channel axis.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "irc_common.cuh"
#include "../../include/irc_b200.h"

using namespace irc;
namespace cg = cooperative_groups;

namespace {

struct View {
    const bf16* p; long long ld; int off, hp, wp, oy, ox, s2d_c;
    __device__ __forceinline__ const bf16* at(int n, int y, int x, int c) const {
        const int Y = y + oy, X = x + ox;
        if (s2d_c) {
            const long long r = (long long)(n * hp + (Y >> 1)) * wp + (X >> 1);
            return p + r * ld + off + ((Y & 1) * 2 + (X & 1)) * s2d_c + c;
        }
        return p + ((long long)(n * hp + Y) * wp + X) * ld + off + c;
    }
};

inline View mk(const irc_view& v) {
    View r; r.p = (const bf16*)v.ptr; r.ld = v.ld; r.off = v.chan_off; r.hp = v.hp; r.wp = v.wp; r.oy = v.oy; r.ox = v.ox; r.s2d_c = v.s2d_c;
    return r;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
    unpack8(__ldg(reinterpret_cast<const uint4*>(p)), v);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ int reflect_idx(int i, int n) {
    if (i < 0) i = -i;
    if (i > n - 1) i = 2 * (n - 1) - i;
    return i;
}
// mean / rstd of 8 channels from the (sum, sum of squares) pairs
// eps < 0: `stats` already holds (mean, 1/std) pairs - the effective moments irc_bn_finalize builds for nn.BatchNorm2d (batch
// statistics and the affine parameters folded into one subtract-and-scale), so every normalise / backward kernel serves both norms
__device__ __forceinline__ void moments8(const float* stats, int n, int C, int c, float inv_cnt, float eps, float (&mu)[8], float (&rs)[8]) {
    const float4* sp = reinterpret_cast<const float4*>(stats + ((long long)n * C + c) * 2);
    if (eps < 0.f) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const float4 s = __ldg(sp + g);
            mu[g * 2] = s.x; rs[g * 2] = s.y; mu[g * 2 + 1] = s.z; rs[g * 2 + 1] = s.w;
        }
        return;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const float4 s = __ldg(sp + g);
        const float m0 = s.x * inv_cnt, m1 = s.z * inv_cnt;
        mu[g * 2] = m0; mu[g * 2 + 1] = m1;
        rs[g * 2] = rsqrtf(fmaxf(s.y * inv_cnt - m0 * m0, 0.f) + eps);
        rs[g * 2 + 1] = rsqrtf(fmaxf(s.w * inv_cnt - m1 * m1, 0.f) + eps);
    }
}
__device__ __forceinline__ float actf(float v, int act, float slope) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v > 0.f ? v : v * slope;
    return v;
}
__device__ __forceinline__ float dactf(float v, int act, float slope) {
    if (act == 1) return v > 0.f ? 1.f : 0.f;
    if (act == 2) return v > 0.f ? 1.f : slope;
    return 1.f;
}

// ---------------------------------------------------------------------------------
// row index table: row_img[q] = n for live rows, -1 for the padding ring
// ---------------------------------------------------------------------------------
__global__ void row_index_kernel(short* out, int n_img, int hp, int wp, int y0, int y1, int x0, int x1) {
    irc::pdl_prologue();
    const long long total = (long long)n_img * hp * wp;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(q % wp);
        const int y = (int)((q / wp) % hp);
        const int n = (int)(q / ((long long)wp * hp));
        out[q] = (y >= y0 && y < y1 && x >= x0 && x < x1) ? (short)n : (short)-1;
    }
}

// Second stage of the order-fixed reductions, folded into the first kernel: the last block of image n to finish (ticket
// counter) adds the per-chunk partials in chunk order (bit-reproducible) and re-arms the counter for the next launch.
__device__ __forceinline__ void finalize_by_last_block(const float* part, float* out, int n, int C2, unsigned* counters) {
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&counters[n], 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int i = threadIdx.x; i < C2; i += blockDim.x) {
        float a = 0.f;
        for (unsigned c = 0; c < gridDim.x; ++c) a += __ldcg(part + ((long long)c * gridDim.y + n) * C2 + i);
        out[(long long)n * C2 + i] = a;
    }
    if (threadIdx.x == 0) counters[n] = 0u;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// The row-walking reductions below are latency-bound with plain loads (ncu: 32 warps/SM x 2 x 512 B in flight = 3.3 TB/s,
// long-scoreboard stalls 15 per issue).  They therefore prefetch kPipeD pixels ahead per thread with cp.async into
// PRIVATE shared-memory slots (slot [stage][tensor][thread]: no barrier needed, only cp.async.wait_group), which puts
// 4x the bytes in flight without a single extra register.
constexpr int kPipeD = 4;

// ---------------------------------------------------------------------------------
// per-(image, channel) sum and sum of squares over the H x W pixels of a view
// block = (C/8) channel vectors x L pixel lanes; grid = (chunks, N)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 4) in_stats_kernel(View z, int C, int H, int W, float* part, float* out, unsigned* counters) {
    irc::pdl_prologue();
    extern __shared__ float sh[];
    const int C8 = C >> 3;
    const int cv = threadIdx.x % C8, lane = threadIdx.x / C8, L = blockDim.x / C8;
    const int n = blockIdx.y;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ss[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int y = blockIdx.x; y < H; y += gridDim.x) {
#pragma unroll 4
        for (int x = lane; x < W; x += L) {
            float v[8];
            load8(z.at(n, y, x, cv * 8), v);
#pragma unroll
            for (int j = 0; j < 8; ++j) { s[j] += v[j]; ss[j] += v[j] * v[j]; }
        }
    }
    // reduce over pixel lanes through shared memory
    float* shs = sh;                          // [L][C8][16]
    if (lane < L) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { shs[(lane * C8 + cv) * 16 + j * 2] = s[j]; shs[(lane * C8 + cv) * 16 + j * 2 + 1] = ss[j]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C8 * 16; i += blockDim.x) {
        float a = 0.f;
        for (int l = 0; l < L; ++l) a += shs[l * C8 * 16 + i];
        part[((long long)blockIdx.x * gridDim.y + n) * C * 2 + i] = a;
    }
    if (counters) finalize_by_last_block(part, out, n, C * 2, counters);
}

// pipelined variant for plain (non space-to-depth) views
__global__ void __launch_bounds__(256, 4) in_stats_pipe_kernel(View z, int C, int H, int W, float* part, float* out, unsigned* counters) {
    irc::pdl_prologue();
    extern __shared__ float sh[];             // [kPipeD][threads] uint4 slots, then reused as [L][C8][16] floats
    const int nt = blockDim.x;
    uint4* slot = reinterpret_cast<uint4*>(sh) + threadIdx.x;
    const int C8 = C >> 3;
    const int cv = threadIdx.x % C8, lane = threadIdx.x / C8, L = blockDim.x / C8;
    const int n = blockIdx.y;
    const bf16* base = z.at(n, 0, 0, cv * 8);
    const long long rstride = (long long)z.wp * z.ld;
    const int pstride = (int)z.ld;
    int iy = blockIdx.x, ix = lane, cy = blockIdx.x, cx = lane;
#pragma unroll
    for (int st = 0; st < kPipeD; ++st) {
        if (iy < H) { cp_async16(slot + st * nt, base + iy * rstride + ix * pstride); ix += L; if (ix >= W) { ix = lane; iy += gridDim.x; } }
        cp_async_commit();
    }
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ss[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int st = 0;
    while (cy < H) {
        cp_async_wait<kPipeD - 1>();
        const uint4 r = slot[st * nt];
        if (iy < H) { cp_async16(slot + st * nt, base + iy * rstride + ix * pstride); ix += L; if (ix >= W) { ix = lane; iy += gridDim.x; } }
        cp_async_commit();
        float v[8];
        unpack8(r, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += v[j]; ss[j] = fmaf(v[j], v[j], ss[j]); }
        cx += L; if (cx >= W) { cx = lane; cy += gridDim.x; }
        st = st + 1 == kPipeD ? 0 : st + 1;
    }
    cp_async_wait<0>();
    __syncthreads();                          // every thread is done with its slots: the buffer becomes the reduction scratch
    float* shs = sh;
#pragma unroll
    for (int j = 0; j < 8; ++j) { shs[(lane * C8 + cv) * 16 + j * 2] = s[j]; shs[(lane * C8 + cv) * 16 + j * 2 + 1] = ss[j]; }
    __syncthreads();
    for (int i = threadIdx.x; i < C8 * 16; i += blockDim.x) {
        float a = 0.f;
        for (int l = 0; l < L; ++l) a += shs[l * C8 * 16 + i];
        part[((long long)blockIdx.x * gridDim.y + n) * C * 2 + i] = a;
    }
    if (counters) finalize_by_last_block(part, out, n, C * 2, counters);
}

// out[j] = sum over chunks of part[chunk][j]: deterministic second stage of the reductions.  One warp per output: lane l
// adds chunks l, l + 32, ... in order, then a fixed shuffle tree combines the lanes (bit-reproducible, and not a single
// block walking hundreds of chunks serially)
__global__ void sum_chunks_kernel(const float* __restrict__ part, int chunks, long long n, float* __restrict__ out) {
    irc::pdl_prologue();
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long j = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); j < n; j += warps) {
        float a = 0.f;
        for (int c = lane; c < chunks; c += 32) a += part[(long long)c * n + j];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        if (lane == 0) out[j] = a;
    }
}

// ---------------------------------------------------------------------------------
// gather: dst = halo( sum_ij wy_i wx_j * pre(src)[ty_i, tx_j] (+ src2[...]) + res )
// ---------------------------------------------------------------------------------
struct GatherP {
    View src, src2, res, dst;
    int has2, has_res, C, n_img;
    const float* stats; float inv_cnt, eps; int act; float slope;
    const int* ty_idx; const float* ty_w; int ky;
    const int* tx_idx; const float* tx_w; int kx;
    int H, W, pad, halo_mode, dst_s2d;
    float slope_eff;     // activation as max(v,0) + slope_eff*min(v,0)
};

template <bool kIdent>
__global__ void __launch_bounds__(256, 4) gather_kernel(const GatherP p) {
    irc::pdl_prologue();
    // one block per (image, padded row); thread = (channel vector, pixel lane): nothing is divided in the loops
    const int C8 = p.C >> 3;
    const int cv = threadIdx.x % C8, lane = threadIdx.x / C8, L = blockDim.x / C8;
    const int c = cv * 8;
    const int Hp = p.H + 2 * p.pad, Wp = p.W + 2 * p.pad;
    for (int row = blockIdx.x; row < p.n_img * Hp; row += gridDim.x) {
        const int n = row / Hp, Y = row - n * Hp;
        int y = Y - p.pad;
        const bool halo_y = y < 0 || y >= p.H;
        if (halo_y) y = reflect_idx(y, p.H);
        float mu[8], rs[8];
        if (p.stats) moments8(p.stats, n, p.C, c, p.inv_cnt, p.eps, mu, rs);
        for (int X = lane; X < Wp; X += L) {
            int x = X - p.pad;
            const bool halo = halo_y || x < 0 || x >= p.W;
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (!halo || p.halo_mode == 1) {
                if (x < 0 || x >= p.W) x = reflect_idx(x, p.W);
                const int ky = kIdent ? 1 : p.ky, kx = kIdent ? 1 : p.kx;
                for (int i = 0; i < ky; ++i) {
                    const int iy = (!kIdent && p.ty_idx) ? __ldg(p.ty_idx + y * p.ky + i) : y;
                    const float wy = (!kIdent && p.ty_w) ? __ldg(p.ty_w + y * p.ky + i) : 1.f;
                    if (wy == 0.f) continue;
                    for (int j = 0; j < kx; ++j) {
                        const int ix = (!kIdent && p.tx_idx) ? __ldg(p.tx_idx + x * p.kx + j) : x;
                        const float w = wy * ((!kIdent && p.tx_w) ? __ldg(p.tx_w + x * p.kx + j) : 1.f);
                        if (w == 0.f) continue;
                        float v[8];
                        load8(p.src.at(n, iy, ix, c), v);
                        if (p.stats) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[k] = actf((v[k] - mu[k]) * rs[k], p.act, p.slope);
                        } else if (p.act) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[k] = actf(v[k], p.act, p.slope);
                        }
                        if (p.has2) {
                            float u[8];
                            load8(p.src2.at(n, iy, ix, c), u);
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[k] += u[k];
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) acc[k] += w * v[k];
                    }
                }
                if (p.has_res) {
                    float u[8];
                    load8(p.res.at(n, y, x, c), u);
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] += u[k];
                }
            }
            bf16* d;
            if (p.dst_s2d) {
                // padded pixel (Y, X) -> 2x2 block (Y/2, X/2), channel group (Y&1)*2 + (X&1)
                const long long r = ((long long)n * (Hp >> 1) + (Y >> 1)) * (Wp >> 1) + (X >> 1);
                d = const_cast<bf16*>(p.dst.p) + r * p.dst.ld + p.dst.off + ((Y & 1) * 2 + (X & 1)) * p.C + c;
            } else {
                d = const_cast<bf16*>(p.dst.p) + ((long long)(n * p.dst.hp + Y - p.pad + p.dst.oy) * p.dst.wp + (X - p.pad + p.dst.ox)) * p.dst.ld + p.dst.off + c;
            }
            store8(d, acc);
        }
    }
}

// ---------------------------------------------------------------------------------
// Identity-table gather (InstanceNorm apply + activation (+ residual) into a frame with its ring) with the source pixels
// prefetched kPipeD ahead per thread through private cp.async slots: same latency argument as the reductions.
// ---------------------------------------------------------------------------------
template <bool kRes>
__global__ void __launch_bounds__(256, 4) gather_ident_pipe_kernel(const GatherP p) {
    irc::pdl_prologue();
    extern __shared__ uint4 gpslots[];
    constexpr int NT = kRes ? 2 : 1;
    const int nt = blockDim.x;
    uint4* slot = gpslots + threadIdx.x;
    const int C8 = p.C >> 3;
    const int cv = threadIdx.x % C8, lane = threadIdx.x / C8, L = blockDim.x / C8;
    const int c = cv * 8;
    const int Hp = p.H + 2 * p.pad, Wp = p.W + 2 * p.pad;
    const int rows = p.n_img * Hp;
    // decode of padded pixel (row, X): source coordinates, or live = false for the zero ring
    auto decode = [&](int row, int X, int& n, int& y, int& x) -> bool {
        n = row / Hp;
        const int Y = row - n * Hp;
        y = Y - p.pad; x = X - p.pad;
        const bool halo = y < 0 || y >= p.H || x < 0 || x >= p.W;
        if (halo) {
            if (p.halo_mode != 1) return false;
            y = reflect_idx(y, p.H); x = reflect_idx(x, p.W);
        }
        return true;
    };
    int irow = blockIdx.x, iX = lane, crow = blockIdx.x, cX = lane;
    auto issue = [&](int st) {
        if (irow < rows) {
            int n, y, x;
            if (decode(irow, iX, n, y, x)) {
                cp_async16(slot + (st * NT + 0) * nt, p.src.at(n, y, x, c));
                if (kRes) cp_async16(slot + (st * NT + 1) * nt, p.res.at(n, y, x, c));
            }
            iX += L; if (iX >= Wp) { iX = lane; irow += gridDim.x; }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int st = 0; st < kPipeD; ++st) issue(st);
    float mu[8], rs[8];
    int cur_n = -1, st = 0;
    while (crow < rows) {
        int n, y, x;
        const bool live = decode(crow, cX, n, y, x);
        if (n != cur_n && p.stats) {
            cur_n = n;
            moments8(p.stats, n, p.C, c, p.inv_cnt, p.eps, mu, rs);
#pragma unroll
            for (int k = 0; k < 8; ++k) mu[k] = -mu[k] * rs[k];
        }
        cp_async_wait<kPipeD - 1>();
        const uint4 zr = slot[(st * NT + 0) * nt];
        uint4 rr = make_uint4(0, 0, 0, 0);
        if (kRes) rr = slot[(st * NT + 1) * nt];
        issue(st);
        float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (live) {
            unpack8(zr, v);
            if (p.stats) {
#pragma unroll
                for (int k = 0; k < 8; ++k) { const float t = fmaf(v[k], rs[k], mu[k]); v[k] = fmaxf(t, 0.f) + p.slope_eff * fminf(t, 0.f); }
            }
            if (kRes) {
                float u[8];
                unpack8(rr, u);
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] += u[k];
            }
        }
        const int Y = crow - n * Hp;
        bf16* d = const_cast<bf16*>(p.dst.p) + ((long long)(n * p.dst.hp + Y - p.pad + p.dst.oy) * p.dst.wp + (cX - p.pad + p.dst.ox)) * p.dst.ld + p.dst.off + c;
        store8(d, v);
        cX += L; if (cX >= Wp) { cX = lane; crow += gridDim.x; }
        st = st + 1 == kPipeD ? 0 : st + 1;
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------
// Identity-table gather, register version: the cp.async-slot kernel above decodes (row, column) twice per 16-byte vector
// (once to issue, once to consume) and spends ~200 instructions per vector - ncu: issue slots 78 % busy at 0.61 of the HBM peak
// (profiles/r5/ncu_gathers_summary.txt).  Here a block owns whole padded frame rows: the row is decoded once (block-uniform), every
// thread streams four interior pixels per step with its four (or eight) 16-byte loads issued back to back, the activation is a
// template parameter, and the 2 * pad ring columns of the row are done afterwards by the first lanes (their sources are L1 / L2 hits).
// ---------------------------------------------------------------------------------
template <int kAct>      // 0: copy, 1: normalise + ReLU, 2: normalise + max(t, 0) + slope_eff * min(t, 0)
__device__ __forceinline__ void ident_vec(const uint4& z, const float (&mu)[8], const float (&rs)[8], float slope_eff, float (&v)[8]) {
    unpack8(z, v);
    if (kAct == 1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaxf(fmaf(v[k], rs[k], mu[k]), 0.f);
    } else if (kAct == 2) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float t = fmaf(v[k], rs[k], mu[k]); v[k] = fmaf(slope_eff, fminf(t, 0.f), fmaxf(t, 0.f)); }
    }
}

template <bool kRes, int kAct>
__global__ void __launch_bounds__(256, kRes ? 3 : 4) gather_ident_lean_kernel(const GatherP p) {
    irc::pdl_prologue();
    constexpr int U = 4;
    const int C8 = p.C >> 3;
    const int cv = threadIdx.x % C8, lane = threadIdx.x / C8, L = blockDim.x / C8;
    const int c = cv * 8;
    const int pad = p.pad, H = p.H, W = p.W;
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    const int rows = p.n_img * Hp;
    const int sld = (int)p.src.ld, dld = (int)p.dst.ld, rld = kRes ? (int)p.res.ld : 0;
    float mu[8], rs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { mu[k] = 0.f; rs[k] = 1.f; }
    int cur_n = -1;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int n = row / Hp, Y = row - n * Hp;
        int y = Y - pad;
        const bool yhalo = y < 0 || y >= H;
        if (yhalo) y = reflect_idx(y, H);
        // padded pixel (Y, X) of the destination frame sits at drow + X * dld
        bf16* drow = const_cast<bf16*>(p.dst.p) + ((long long)(n * p.dst.hp + Y - pad + p.dst.oy) * p.dst.wp + (p.dst.ox - pad)) * p.dst.ld + p.dst.off + c;
        if (yhalo && p.halo_mode != 1) {
            for (int X = lane; X < Wp; X += L) *reinterpret_cast<uint4*>(drow + X * dld) = make_uint4(0, 0, 0, 0);
            continue;
        }
        if (kAct != 0 && n != cur_n) {
            cur_n = n;
            moments8(p.stats, n, p.C, c, p.inv_cnt, p.eps, mu, rs);
#pragma unroll
            for (int k = 0; k < 8; ++k) mu[k] = -mu[k] * rs[k];
        }
        const bf16* srow = p.src.at(n, y, 0, c);
        const bf16* rrow = kRes ? p.res.at(n, y, 0, c) : nullptr;
        for (int x0 = lane; x0 < W; x0 += U * L) {
            uint4 z[U], r[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int x = x0 + u * L;
                if (x < W) {
                    z[u] = __ldg(reinterpret_cast<const uint4*>(srow + x * sld));
                    if (kRes) r[u] = __ldg(reinterpret_cast<const uint4*>(rrow + x * rld));
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int x = x0 + u * L;
                if (x < W) {
                    float v[8];
                    ident_vec<kAct>(z[u], mu, rs, p.slope_eff, v);
                    if (kRes) {
                        float w[8];
                        unpack8(r[u], w);
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[k] += w[k];
                    }
                    store8(drow + (x + pad) * dld, v);
                }
            }
        }
        // ring columns of this row: X in [0, pad) and [W + pad, W + 2 pad)
        for (int j = lane; j < 2 * pad; j += L) {
            const int X = j < pad ? j : W + j;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (p.halo_mode == 1) {
                const int x = reflect_idx(X - pad, W);
                ident_vec<kAct>(__ldg(reinterpret_cast<const uint4*>(srow + x * sld)), mu, rs, p.slope_eff, v);
                if (kRes) {
                    float w[8];
                    load8(rrow + x * rld, w);
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] += w[k];
                }
            }
            store8(drow + X * dld, v);
        }
    }
}

// ---------------------------------------------------------------------------------
// Lean table gather: K x K taps with K a compile-time constant (2: reflection fold / Downsample^T, 3: Downsample and
// UpsampleAA, 6: UpsampleAA^T).  One block per output row: the row's y-entries and source row pointers are computed once,
// the x-entries of a pixel sit in registers, and the K*K 16-byte loads per output vector hit L1/L2 (every source pixel
// is shared by neighbouring outputs).  No per-element integer division, no per-tap table reads.
// ---------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256, 3) gather_lean_kernel(const GatherP p) {
    irc::pdl_prologue();
    const int C8 = p.C >> 3;
    const int cv = threadIdx.x % C8, lane = threadIdx.x / C8, L = blockDim.x / C8;
    const int c = cv * 8;
    const int Hp = p.H + 2 * p.pad, Wp = p.W + 2 * p.pad;
    for (int row = blockIdx.x; row < p.n_img * Hp; row += gridDim.x) {
        const int n = row / Hp, Y = row - n * Hp;
        int y = Y - p.pad;
        const bool halo_y = y < 0 || y >= p.H;
        if (halo_y) y = reflect_idx(y, p.H);
        bf16* drow = const_cast<bf16*>(p.dst.p) + ((long long)(n * p.dst.hp + Y - p.pad + p.dst.oy) * p.dst.wp + (p.dst.ox - p.pad)) * p.dst.ld + p.dst.off + c;
        if (halo_y && p.halo_mode == 0) {
            const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int X = lane; X < Wp; X += L) store8(drow + (long long)X * p.dst.ld, z);
            continue;
        }
        float mu[8], rs[8];
        if (p.stats) {
            moments8(p.stats, n, p.C, c, p.inv_cnt, p.eps, mu, rs);
#pragma unroll
            for (int k = 0; k < 8; ++k) mu[k] = -mu[k] * rs[k];
        }
        // y entries and source row base pointers of this output row
        float wy[K]; const bf16* r1[K]; const bf16* r2[K];
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const bool h = p.ty_idx && i < p.ky;
            wy[i] = h ? __ldg(p.ty_w + y * p.ky + i) : ((!p.ty_idx && i == 0) ? 1.f : 0.f);
            const int iy = h ? __ldg(p.ty_idx + y * p.ky + i) : y;
            r1[i] = p.src.at(n, iy, 0, c);
            r2[i] = p.has2 ? p.src2.at(n, iy, 0, c) : nullptr;
        }
        for (int X = lane; X < Wp; X += L) {
            int x = X - p.pad;
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const bool halo_x = x < 0 || x >= p.W;
            if (!halo_x || p.halo_mode == 1) {
                if (halo_x) x = reflect_idx(x, p.W);
                float wx[K]; long long ox[K], ox2[K];
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const bool h = p.tx_idx && j < p.kx;
                    wx[j] = h ? __ldg(p.tx_w + x * p.kx + j) : ((!p.tx_idx && j == 0) ? 1.f : 0.f);
                    const int ix = h ? __ldg(p.tx_idx + x * p.kx + j) : x;
                    ox[j] = (long long)ix * p.src.ld;
                    ox2[j] = (long long)ix * p.src2.ld;
                }
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    if (wy[i] == 0.f) continue;
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        if (wx[j] == 0.f) continue;
                        const float w = wy[i] * wx[j];
                        float v[8];
                        load8(r1[i] + ox[j], v);
                        if (p.stats) {          // mu[] holds -mean*rstd here: one FFMA per element, ReLU = one FMNMX
                            if (p.act == 1) {
#pragma unroll
                                for (int k = 0; k < 8; ++k) v[k] = fmaxf(fmaf(v[k], rs[k], mu[k]), 0.f);
                            } else {
#pragma unroll
                                for (int k = 0; k < 8; ++k) { const float t = fmaf(v[k], rs[k], mu[k]); v[k] = fmaxf(t, 0.f) + p.slope_eff * fminf(t, 0.f); }
                            }
                        }
                        if (p.has2) {
                            float u[8];
                            load8(r2[i] + ox2[j], u);
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[k] += u[k];
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) acc[k] += w * v[k];
                    }
                }
                if (p.has_res) {
                    float u[8];
                    load8(p.res.at(n, y, x, c), u);
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] += u[k];
                }
            }
            store8(drow + (long long)X * p.dst.ld, acc);
        }
    }
}

// ---------------------------------------------------------------------------------
// Streaming separable gather (the anti-aliased Downsample / UpsampleAA stencils and their transposes).
// One thread = 8 channels of one output column; it walks down a strip of output rows keeping the last K horizontally
// filtered source rows in registers.  Source rows enter strictly in order (the host checks that the last source row
// of the y-table is non-decreasing), each is read, normalised and x-filtered ONCE per output column, the raw 16-byte
// loads of the next source row are in flight while the current output row is y-filtered and stored.  Per output vector
// this costs (kx loads + kx x-taps) per NEW source row + K y-taps, instead of ky*kx loads/normalisations.
// The padding ring of the destination is written by the threads that own the bordering pixels (zeros, or the
// reflected value when halo_mode == 1).
// ---------------------------------------------------------------------------------
constexpr int kMaxStrip = 32;

template <int K, bool kNorm, bool kTwo>
__global__ void __launch_bounds__(256, K <= 3 ? 3 : 2) gather_stream_kernel(const GatherP p, int strip) {
    irc::pdl_prologue();
    __shared__ int s_hi[kMaxStrip];
    __shared__ int s_lo0;
    __shared__ float s_wd[kMaxStrip][K];
    const int C8 = p.C >> 3;
    const int cv = threadIdx.x % C8, lane = threadIdx.x / C8, L = blockDim.x / C8;
    const int c = cv * 8;
    const int n = blockIdx.z;
    const int ya = blockIdx.y * strip;
    const int rows = min(strip, p.H - ya);
    const int x = blockIdx.x * L + lane;
    if ((int)threadIdx.x < rows) {
        // dense y-weights of output row y over the window [hi+1-K, hi] of source rows
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
    if (x >= p.W) return;
    // x entries of this output column
    float wx[K]; int ox[K], ox2[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const bool h = j < p.kx;
        wx[j] = h ? __ldg(p.tx_w + x * p.kx + j) : 0.f;
        const int ix = h ? __ldg(p.tx_idx + x * p.kx + j) : 0;
        ox[j] = ix * (int)p.src.ld;
        ox2[j] = kTwo ? ix * (int)p.src2.ld : 0;
    }
    float mu[8], rs[8];
    if (kNorm) {
        moments8(p.stats, n, p.C, c, p.inv_cnt, p.eps, mu, rs);
#pragma unroll
        for (int k = 0; k < 8; ++k) mu[k] = -mu[k] * rs[k];
    }
    const bf16* srow = p.src.at(n, 0, 0, c);
    const long long sstep = (long long)p.src.wp * p.src.ld;
    const bf16* srow2 = kTwo ? p.src2.at(n, 0, 0, c) : nullptr;
    const long long sstep2 = kTwo ? (long long)p.src2.wp * p.src2.ld : 0;
    // destination: interior pixel (y, x) -> padded (y + pad, x + pad); extra ring column / row owned by this thread
    const int pad = p.pad, W = p.W, H = p.H;
    int eX = -1;
    if (pad) {
        if (p.halo_mode == 1) eX = (x >= 1 && x <= pad) ? pad - x : ((x >= W - 1 - pad && x <= W - 2) ? pad + 2 * (W - 1) - x : -1);
        else eX = x < pad ? x : (x >= W - pad ? x + 2 * pad : -1);
    }
    bf16* dbase = const_cast<bf16*>(p.dst.p) + ((long long)(n * p.dst.hp + p.dst.oy - pad) * p.dst.wp + (p.dst.ox - pad)) * p.dst.ld + p.dst.off + c;
    const long long dstep = (long long)p.dst.wp * p.dst.ld;

    bf16* dptr = dbase + (ya + pad) * dstep + (long long)(x + pad) * p.dst.ld;      // advances one frame row per output row
    const long long dex = eX >= 0 ? (long long)(eX - (x + pad)) * p.dst.ld : 0;
    float hb[K][8];
#pragma unroll
    for (int s = 0; s < K; ++s)
#pragma unroll
        for (int k = 0; k < 8; ++k) hb[s][k] = 0.f;
    int top = s_lo0;
    const int last = s_hi[rows - 1];
    uint4 raw[K], raw2[K];
    // unused table slots have weight 0 and index 0: they load pixel 0 of the row and add nothing, which is cheaper than a
    // branch per tap; the row pointers advance by one frame row per fetch (rows enter in order)
    const bf16* rp = srow + top * sstep;
    const bf16* rp2 = kTwo ? srow2 + top * sstep2 : nullptr;
    auto fetch = [&]() {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            raw[j] = __ldg(reinterpret_cast<const uint4*>(rp + ox[j]));
            if (kTwo) raw2[j] = __ldg(reinterpret_cast<const uint4*>(rp2 + ox2[j]));
        }
        rp += sstep;
        if (kTwo) rp2 += sstep2;
    };
    if (top <= last) fetch();
    for (int t = 0; t < rows; ++t) {
        const int hi = s_hi[t];
        while (top <= hi) {
            float h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int j = 0; j < K; ++j) {
                {
                    float v[8];
                    unpack8(raw[j], v);
                    if (kNorm) {
                        if (p.act == 1) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[k] = fmaxf(fmaf(v[k], rs[k], mu[k]), 0.f);
                        } else {
#pragma unroll
                            for (int k = 0; k < 8; ++k) { const float u = fmaf(v[k], rs[k], mu[k]); v[k] = fmaxf(u, 0.f) + p.slope_eff * fminf(u, 0.f); }
                        }
                    }
                    if (kTwo) {
                        float u[8];
                        unpack8(raw2[j], u);
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[k] += u[k];
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) h[k] = fmaf(wx[j], v[k], h[k]);
                }
            }
            ++top;
            if (top <= last) fetch();
#pragma unroll
            for (int s = 0; s + 1 < K; ++s)
#pragma unroll
                for (int k = 0; k < 8; ++k) hb[s][k] = hb[s + 1][k];
#pragma unroll
            for (int k = 0; k < 8; ++k) hb[K - 1][k] = h[k];
        }
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int s = 0; s < K; ++s) {
            const float w = s_wd[t][s];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fmaf(w, hb[s][k], acc[k]);
        }
        const int y = ya + t;
        const uint4 val = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
        *reinterpret_cast<uint4*>(dptr) = val;
        if (pad && (eX >= 0 || y <= pad || y >= H - 1 - pad)) {
            const uint4 ring = p.halo_mode == 1 ? val : make_uint4(0, 0, 0, 0);
            int eY;
            if (p.halo_mode == 1) eY = (y >= 1 && y <= pad) ? pad - y : ((y >= H - 1 - pad && y <= H - 2) ? pad + 2 * (H - 1) - y : -1);
            else eY = y < pad ? y : (y >= H - pad ? y + 2 * pad : -1);
            if (eX >= 0) *reinterpret_cast<uint4*>(dptr + dex) = ring;
            if (eY >= 0) {
                bf16* erow = dptr + (long long)(eY - (y + pad)) * dstep;
                *reinterpret_cast<uint4*>(erow) = ring;
                if (eX >= 0) *reinterpret_cast<uint4*>(erow + dex) = ring;
            }
        }
        dptr += dstep;
    }
}

// ---------------------------------------------------------------------------------
// Staged-rows variant of the streaming gather.  In the kernel above every output column issues its own kx 16-byte global loads per
// new source row, one row ahead: ncu shows 2.4 - 4.4 long-scoreboard stall cycles per issue at 23 - 34 % occupancy (the register
// windows), i.e. the DRAM latency is not covered (profiles/r5/ncu_gathers_summary.txt).  Here the block copies the source-row
// segment its L output columns need with cp.async into a kRowsStg-deep shared-memory ring (raw bf16, no registers, three rows in
// flight per block), one __syncthreads per source row; the x-taps are 16-byte shared-memory reads.  Each source pixel crosses
// L2 -> SM once per block instead of kx / stride times.
// ---------------------------------------------------------------------------------
constexpr int kRowsNS = 3;      // staged vectors per thread and source row: the segment spans at most 3 L columns (host-checked bound)
constexpr int kRowsStg = 4;     // ring depth

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void cp_async16_s(uint32_t saddr, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gmem) : "memory");
}

template <int K, int kNorm, bool kTwo>      // kNorm: 0 raw, 1 normalise + ReLU, 2 normalise + max(t, 0) + slope_eff * min(t, 0)
__global__ void __launch_bounds__(256, K <= 3 ? 3 : 2) gather_rows_kernel(const GatherP p, int strip, int sw_max) {
    irc::pdl_prologue();
    extern __shared__ uint4 s_ring[];            // [kRowsStg][kTwo ? 2 : 1][sw_max][C8]
    __shared__ int s_hi[kMaxStrip];
    __shared__ int s_lo0, s_cmin, s_cmax;
    __shared__ float s_wd[kMaxStrip][K];
    const int C8 = p.C >> 3;
    const int cv = threadIdx.x % C8, lane = threadIdx.x / C8, L = blockDim.x / C8;
    const int c = cv * 8;
    const int n = blockIdx.z;
    const int ya = blockIdx.y * strip;
    const int rows = min(strip, p.H - ya);
    const int x = blockIdx.x * L + lane;
    const bool xok = x < p.W;
    if (threadIdx.x == 0) { s_cmin = 0x7fffffff; s_cmax = -1; }
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
    // x entries of this output column (unused table slots: weight 0, staged column 0)
    float wx[K]; int ix[K];
    int mn = 0x7fffffff, mx = -1;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const bool h = xok && j < p.kx;
        wx[j] = h ? __ldg(p.tx_w + x * p.kx + j) : 0.f;
        ix[j] = h ? __ldg(p.tx_idx + x * p.kx + j) : 0;
        if (wx[j] != 0.f) { mn = min(mn, ix[j]); mx = max(mx, ix[j]); }
    }
    __syncthreads();
    if (cv == 0 && mx >= 0) { atomicMin(&s_cmin, mn); atomicMax(&s_cmax, mx); }
    __syncthreads();
    const int cmin = s_cmin, sw = s_cmax - cmin + 1;       // sw <= sw_max (checked by the host from the table bandwidth)
    const uint32_t ring_s = smem_u32(s_ring);              // shared-space byte addresses, computed once
    uint32_t xo[K];                                        // byte offsets of the taps inside a staged row
#pragma unroll
    for (int j = 0; j < K; ++j) xo[j] = (uint32_t)((wx[j] != 0.f ? ix[j] - cmin : 0) * C8 + cv) * 16u;
    float mu[8], rs[8];
    if (kNorm) {
        moments8(p.stats, n, p.C, c, p.inv_cnt, p.eps, mu, rs);
#pragma unroll
        for (int k = 0; k < 8; ++k) mu[k] = -mu[k] * rs[k];
    }
    const int sld = (int)p.src.ld, sld2 = kTwo ? (int)p.src2.ld : 0;
    const long long sstep = (long long)p.src.wp * p.src.ld;
    const long long sstep2 = kTwo ? (long long)p.src2.wp * p.src2.ld : 0;
    // destination
    const int pad = p.pad, W = p.W, H = p.H;
    int eX = -1;
    if (pad && xok) {
        if (p.halo_mode == 1) eX = (x >= 1 && x <= pad) ? pad - x : ((x >= W - 1 - pad && x <= W - 2) ? pad + 2 * (W - 1) - x : -1);
        else eX = x < pad ? x : (x >= W - pad ? x + 2 * pad : -1);
    }
    bf16* dbase = const_cast<bf16*>(p.dst.p) + ((long long)(n * p.dst.hp + p.dst.oy - pad) * p.dst.wp + (p.dst.ox - pad)) * p.dst.ld + p.dst.off + c;
    const long long dstep = (long long)p.dst.wp * p.dst.ld;
    bf16* dptr = dbase + (ya + pad) * dstep + (long long)(x + pad) * p.dst.ld;
    const long long dex = eX >= 0 ? (long long)(eX - (x + pad)) * p.dst.ld : 0;
    float hb[K][8];
#pragma unroll
    for (int s = 0; s < K; ++s)
#pragma unroll
        for (int k = 0; k < 8; ++k) hb[s][k] = 0.f;
    int top = s_lo0;
    const int last = s_hi[rows - 1];
    // staged columns of this thread: cmin + lane + i * L, channel vector cv (blockDim is a multiple of C8)
    const bf16* rp = p.src.at(n, top, cmin + lane, c);
    const bf16* rp2 = kTwo ? p.src2.at(n, top, cmin + lane, c) : nullptr;
    const uint32_t slot_b = (uint32_t)((kTwo ? 2 : 1) * sw_max * C8) * 16u;      // bytes per ring slot
    const uint32_t two_b = (uint32_t)(sw_max * C8) * 16u;                         // offset of the second source inside a slot
    const uint32_t ring_end = ring_s + kRowsStg * slot_b;
    const uint32_t step_b = (uint32_t)(L * C8) * 16u;                             // staged vector i of this thread: + i * step_b
    const int gstep = L * sld, gstep2 = L * sld2;
    bool on[kRowsNS];
#pragma unroll
    for (int i = 0; i < kRowsNS; ++i) on[i] = lane + i * L < sw;
    int nxt = top;                                         // next source row to issue
    uint32_t idst = ring_s + (uint32_t)(lane * C8 + cv) * 16u;      // ... and where its first vector goes
    auto issue = [&]() {
        if (nxt <= last) {
#pragma unroll
            for (int i = 0; i < kRowsNS; ++i)
                if (on[i]) {
                    cp_async16_s(idst + i * step_b, rp + i * gstep);
                    if (kTwo) cp_async16_s(idst + two_b + i * step_b, rp2 + i * gstep2);
                }
            rp += sstep;
            if (kTwo) rp2 += sstep2;
            ++nxt;
            idst += slot_b;
            if (idst >= ring_end) idst -= kRowsStg * slot_b;
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < kRowsStg - 1; ++s) issue();
    uint32_t cur = ring_s;                                 // slot of source row `top`
    const float slope_eff = p.slope_eff;
    for (int t = 0; t < rows; ++t) {
        const int hi = s_hi[t];
        while (top <= hi) {
            cp_async_wait<kRowsStg - 2>();       // this thread's copies of row `top` have landed ...
            __syncthreads();                     // ... and so have everyone else's; the slot read last iteration is free
            issue();
            float h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int j = 0; j < K; ++j) {
                float v[8];
                unpack8(lds128(cur + xo[j]), v);
                if (kNorm == 1) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = fmaxf(fmaf(v[k], rs[k], mu[k]), 0.f);
                } else if (kNorm == 2) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) { const float u = fmaf(v[k], rs[k], mu[k]); v[k] = fmaf(slope_eff, fminf(u, 0.f), fmaxf(u, 0.f)); }
                }
                if (kTwo) {
                    float u[8];
                    unpack8(lds128(cur + two_b + xo[j]), u);
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] += u[k];
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) h[k] = fmaf(wx[j], v[k], h[k]);
            }
#pragma unroll
            for (int s = 0; s + 1 < K; ++s)
#pragma unroll
                for (int k = 0; k < 8; ++k) hb[s][k] = hb[s + 1][k];
#pragma unroll
            for (int k = 0; k < 8; ++k) hb[K - 1][k] = h[k];
            ++top;
            cur += slot_b;
            if (cur >= ring_end) cur = ring_s;
        }
        if (xok) {
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int s = 0; s < K; ++s) {
                const float w = s_wd[t][s];
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fmaf(w, hb[s][k], acc[k]);
            }
            const int y = ya + t;
            const uint4 val = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
            *reinterpret_cast<uint4*>(dptr) = val;
            if (pad && (eX >= 0 || y <= pad || y >= H - 1 - pad)) {
                const uint4 ring = p.halo_mode == 1 ? val : make_uint4(0, 0, 0, 0);
                int eY;
                if (p.halo_mode == 1) eY = (y >= 1 && y <= pad) ? pad - y : ((y >= H - 1 - pad && y <= H - 2) ? pad + 2 * (H - 1) - y : -1);
                else eY = y < pad ? y : (y >= H - pad ? y + 2 * pad : -1);
                if (eX >= 0) *reinterpret_cast<uint4*>(dptr + dex) = ring;
                if (eY >= 0) {
                    bf16* erow = dptr + (long long)(eY - (y + pad)) * dstep;
                    *reinterpret_cast<uint4*>(erow) = ring;
                    if (eX >= 0) *reinterpret_cast<uint4*>(erow + dex) = ring;
                }
            }
        }
        dptr += dstep;
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------
// Shared-memory tiled gather for real stencils (ky*kx > 1): one block = TY x TX output pixels x 32 channels.
// The source patch the tile needs is loaded once (coalesced 16-byte loads), normalised / activated / summed with
// src2 once per source pixel, kept in shared memory as fp32, and every output pixel then takes its taps from
// shared memory.  Global traffic = patch + tile instead of taps x tile.
// ---------------------------------------------------------------------------------
constexpr int kTileCC = 32;   // channels per block

template <int K>
__global__ void __launch_bounds__(256) gather_tiled_kernel(const GatherP p, int TY, int TX, int maxNy, int maxNx) {
    irc::pdl_prologue();
    extern __shared__ float patch[];                 // [ny*nx][kTileCC]
    __shared__ int ys[64], xs[64];                   // interior coordinate of every tile row / column (-1 = zero ring, -2 = outside)
    __shared__ int box[4];                           // lo_y, hi_y, lo_x, hi_x
    const int CV = kTileCC / 8;
    const int Hp = p.H + 2 * p.pad, Wp = p.W + 2 * p.pad;
    const int cchunks = p.C / kTileCC;
    const int n = blockIdx.z / cchunks, c0 = (blockIdx.z % cchunks) * kTileCC;
    const int Y0 = blockIdx.y * TY, X0 = blockIdx.x * TX;
    const int t = threadIdx.x;
    if (t < 4) box[t] = (t & 1) ? -1 : 0x7fffffff;
    __syncthreads();
    if (t < TY + TX) {
        const bool isy = t < TY;
        const int P = isy ? Y0 + t : X0 + (t - TY);
        const int lim = isy ? Hp : Wp, ext = isy ? p.H : p.W;
        int v = -2;
        if (P < lim) {
            v = P - p.pad;
            if (v < 0 || v >= ext) v = p.halo_mode == 1 ? reflect_idx(v, ext) : -1;
        }
        (isy ? ys : xs)[isy ? t : t - TY] = v;
        if (v >= 0) {
            const int* idx = isy ? p.ty_idx : p.tx_idx; const float* w = isy ? p.ty_w : p.tx_w; const int k = isy ? p.ky : p.kx;
            int lo = 0x7fffffff, hi = -1;
            if (!idx) { lo = hi = v; }
            else for (int i = 0; i < k; ++i) if (__ldg(w + v * k + i) != 0.f) { const int q = __ldg(idx + v * k + i); lo = min(lo, q); hi = max(hi, q); }
            if (hi >= 0) { atomicMin(&box[isy ? 0 : 2], lo); atomicMax(&box[isy ? 1 : 3], hi); }
        }
    }
    __syncthreads();
    const int lo_y = box[0], lo_x = box[2];
    const int ny = box[1] - lo_y + 1, nx = box[3] - lo_x + 1;
    const int cv = t % CV, c = c0 + cv * 8;
    if (box[1] >= 0 && box[3] >= 0) {
        if (ny > maxNy || nx > maxNx) { if (t == 0) printf("irc: gather tile patch %dx%d exceeds %dx%d\n", ny, nx, maxNy, maxNx); __trap(); }
        float mu[8], rs[8];
        if (p.stats) {
            moments8(p.stats, n, p.C, c, p.inv_cnt, p.eps, mu, rs);
#pragma unroll
            for (int k = 0; k < 8; ++k) mu[k] = -mu[k] * rs[k];
        }
        for (int item = t / CV; item < ny * nx; item += blockDim.x / CV) {
            const int py = item / nx, px = item - py * nx;
            float v[8];
            load8(p.src.at(n, lo_y + py, lo_x + px, c), v);
            if (p.stats) {
#pragma unroll
                for (int k = 0; k < 8; ++k) { const float tt = fmaf(v[k], rs[k], mu[k]); v[k] = fmaxf(tt, 0.f) + p.slope_eff * fminf(tt, 0.f); }
            } else if (p.act) {
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = actf(v[k], p.act, p.slope);
            }
            if (p.has2) {
                float u[8];
                load8(p.src2.at(n, lo_y + py, lo_x + px, c), u);
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] += u[k];
            }
            float4* d = reinterpret_cast<float4*>(patch + (size_t)item * kTileCC + cv * 8);
            d[0] = make_float4(v[0], v[1], v[2], v[3]); d[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
    __syncthreads();
    for (int item = t / CV; item < TY * TX; item += blockDim.x / CV) {
        const int r = item / TX, q = item - r * TX;
        const int y = ys[r], x = xs[q];
        if (y == -2 || x == -2) continue;
        const int Y = Y0 + r, X = X0 + q;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (y >= 0 && x >= 0) {
            // the K (index, weight) pairs of this output row / column live in registers; zero weights mark padding entries
            int oy[K], ox[K]; float wy[K], wx[K];
#pragma unroll
            for (int i = 0; i < K; ++i) {
                const bool hy = p.ty_idx && i < p.ky, hx = p.tx_idx && i < p.kx;
                wy[i] = hy ? __ldg(p.ty_w + y * p.ky + i) : ((!p.ty_idx && i == 0) ? 1.f : 0.f);
                wx[i] = hx ? __ldg(p.tx_w + x * p.kx + i) : ((!p.tx_idx && i == 0) ? 1.f : 0.f);
                oy[i] = ((hy ? __ldg(p.ty_idx + y * p.ky + i) : y) - lo_y) * nx;
                ox[i] = (hx ? __ldg(p.tx_idx + x * p.kx + i) : x) - lo_x;
            }
#pragma unroll
            for (int i = 0; i < K; ++i) {
                if (wy[i] == 0.f) continue;
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    if (wx[j] == 0.f) continue;
                    const float w = wy[i] * wx[j];
                    const float4* s4 = reinterpret_cast<const float4*>(patch + (size_t)(oy[i] + ox[j]) * kTileCC + cv * 8);
                    const float4 a = s4[0], b = s4[1];
                    acc[0] += w * a.x; acc[1] += w * a.y; acc[2] += w * a.z; acc[3] += w * a.w;
                    acc[4] += w * b.x; acc[5] += w * b.y; acc[6] += w * b.z; acc[7] += w * b.w;
                }
            }
            if (p.has_res) {
                float u[8];
                load8(p.res.at(n, y, x, c), u);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += u[k];
            }
        }
        bf16* d;
        if (p.dst_s2d) {
            const long long rr = ((long long)n * (Hp >> 1) + (Y >> 1)) * (Wp >> 1) + (X >> 1);
            d = const_cast<bf16*>(p.dst.p) + rr * p.dst.ld + p.dst.off + ((Y & 1) * 2 + (X & 1)) * p.C + c;
        } else {
            d = const_cast<bf16*>(p.dst.p) + ((long long)(n * p.dst.hp + Y - p.pad + p.dst.oy) * p.dst.wp + (X - p.pad + p.dst.ox)) * p.dst.ld + p.dst.off + c;
        }
        store8(d, acc);
    }
}

// ---------------------------------------------------------------------------------
// InstanceNorm(+activation) backward.
//   g      = sum_ij wy wx (gsrc + gsrc2)[ty_i, tx_j]      (gradient w.r.t. the activated output)
//   gd     = g * act'(xhat),  xhat = (z - mu) * rstd
//   reduce : bsum[n][c] = (sum gd, sum gd*xhat)
//   apply  : dz = rstd * (gd - bsum0/cnt - xhat * bsum1/cnt)
// Without stats: dz = g * act'(z) (layers without a norm).
// ---------------------------------------------------------------------------------
struct InBwdP {
    View z, g1, g2, dz;
    int has2, C, n_img, H, W;
    const float* stats; float inv_cnt, eps; int act; float slope;
    const int* ty_idx; const float* ty_w; int ky;
    const int* tx_idx; const float* tx_w; int kx;
    float* bsum;
    float* part;
    unsigned* counters;
};

template <bool kIdent>
__device__ __forceinline__ void bwd_gather(const InBwdP& p, int n, int y, int x, int c, float (&g)[8]) {
    if (kIdent) {
        load8(p.g1.at(n, y, x, c), g);
        if (p.has2) {
            float u[8];
            load8(p.g2.at(n, y, x, c), u);
#pragma unroll
            for (int k = 0; k < 8; ++k) g[k] += u[k];
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = 0.f;
    for (int i = 0; i < p.ky; ++i) {
        const int iy = p.ty_idx ? __ldg(p.ty_idx + y * p.ky + i) : y;
        const float wy = p.ty_w ? __ldg(p.ty_w + y * p.ky + i) : 1.f;
        if (wy == 0.f) continue;
        for (int j = 0; j < p.kx; ++j) {
            const int ix = p.tx_idx ? __ldg(p.tx_idx + x * p.kx + j) : x;
            const float w = wy * (p.tx_w ? __ldg(p.tx_w + x * p.kx + j) : 1.f);
            if (w == 0.f) continue;
            float v[8];
            load8(p.g1.at(n, iy, ix, c), v);
            if (p.has2) {
                float u[8];
                load8(p.g2.at(n, iy, ix, c), u);
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] += u[k];
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) g[k] += w * v[k];
        }
    }
}

template <bool kIdent>
__global__ void __launch_bounds__(256, 4) in_bwd_reduce_kernel(const InBwdP p) {
    irc::pdl_prologue();
    extern __shared__ float sh[];
    const int C8 = p.C >> 3;
    const int cv = threadIdx.x % C8, lane = threadIdx.x / C8, L = blockDim.x / C8;
    const int n = blockIdx.y, c = cv * 8;
    float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    {
        float mu[8], rs[8];
        moments8(p.stats, n, p.C, c, p.inv_cnt, p.eps, mu, rs);
        for (int y = blockIdx.x; y < p.H; y += gridDim.x) {
            for (int x = lane; x < p.W; x += L) {
                float g[8], zv[8];
                bwd_gather<kIdent>(p, n, y, x, c, g);
                load8(p.z.at(n, y, x, c), zv);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float xh = (zv[k] - mu[k]) * rs[k];
                    const float gd = g[k] * dactf(xh, p.act, p.slope);
                    s1[k] += gd; s2[k] += gd * xh;
                }
            }
        }
    }
    if (lane < L) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { sh[(lane * C8 + cv) * 16 + j * 2] = s1[j]; sh[(lane * C8 + cv) * 16 + j * 2 + 1] = s2[j]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C8 * 16; i += blockDim.x) {
        float a = 0.f;
        for (int l = 0; l < L; ++l) a += sh[l * C8 * 16 + i];
        p.part[((long long)blockIdx.x * gridDim.y + n) * p.C * 2 + i] = a;
    }
    if (p.counters) finalize_by_last_block(p.part, p.bsum, n, p.C * 2, p.counters);
}

template <bool kIdent>
__global__ void __launch_bounds__(256, 4) in_bwd_apply_kernel(const InBwdP p) {
    irc::pdl_prologue();
    const int C8 = p.C >> 3;
    const int cv = threadIdx.x % C8, lane = threadIdx.x / C8, L = blockDim.x / C8;
    const int c = cv * 8;
    for (int row = blockIdx.x; row < p.n_img * p.H; row += gridDim.x) {
        const int n = row / p.H, y = row - n * p.H;
        float mu[8], rs[8], b1[8], b2[8];
        if (p.stats) {
            moments8(p.stats, n, p.C, c, p.inv_cnt, p.eps, mu, rs);
            const float4* bp = reinterpret_cast<const float4*>(p.bsum + ((long long)n * p.C + c) * 2);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 b = __ldg(bp + q);
                b1[q * 2] = b.x * p.inv_cnt; b2[q * 2] = b.y * p.inv_cnt; b1[q * 2 + 1] = b.z * p.inv_cnt; b2[q * 2 + 1] = b.w * p.inv_cnt;
            }
        }
        for (int x = lane; x < p.W; x += L) {
            float g[8], zv[8], o[8];
            bwd_gather<kIdent>(p, n, y, x, c, g);
            load8(p.z.at(n, y, x, c), zv);
            if (p.stats) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float xh = (zv[k] - mu[k]) * rs[k];
                    const float gd = g[k] * dactf(xh, p.act, p.slope);
                    o[k] = rs[k] * (gd - b1[k] - xh * b2[k]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = g[k] * dactf(zv[k], p.act, p.slope);
            }
            store8(const_cast<bf16*>(p.dz.at(n, y, x, c)), o);
        }
    }
}

// ---------------------------------------------------------------------------------
// Pipelined identity-table versions of the two passes (plain views; one or two gradient sources)
// ---------------------------------------------------------------------------------
template <bool kTwo>
__global__ void __launch_bounds__(256, 4) in_bwd_reduce_pipe_kernel(const InBwdP p) {
    irc::pdl_prologue();
    extern __shared__ float sh[];
    constexpr int NT = kTwo ? 3 : 2;
    const int nt = blockDim.x;
    uint4* slot = reinterpret_cast<uint4*>(sh) + threadIdx.x;         // [kPipeD][NT][nt]
    const int C8 = p.C >> 3;
    const int cv = threadIdx.x % C8, lane = threadIdx.x / C8, L = blockDim.x / C8;
    const int n = blockIdx.y, c = cv * 8;
    const bf16* gb = p.g1.at(n, 0, 0, c); const long long grs = (long long)p.g1.wp * p.g1.ld; const int gps = (int)p.g1.ld;
    const bf16* zb = p.z.at(n, 0, 0, c); const long long zrs = (long long)p.z.wp * p.z.ld; const int zps = (int)p.z.ld;
    const bf16* hb = kTwo ? p.g2.at(n, 0, 0, c) : nullptr; const long long hrs = kTwo ? (long long)p.g2.wp * p.g2.ld : 0; const int hps = kTwo ? (int)p.g2.ld : 0;
    int iy = blockIdx.x, ix = lane, cy = blockIdx.x, cx = lane;
    auto issue = [&](int st) {
        if (iy < p.H) {
            cp_async16(slot + (st * NT + 0) * nt, gb + iy * grs + ix * gps);
            cp_async16(slot + (st * NT + 1) * nt, zb + iy * zrs + ix * zps);
            if (kTwo) cp_async16(slot + (st * NT + 2) * nt, hb + iy * hrs + ix * hps);
            ix += L; if (ix >= p.W) { ix = lane; iy += gridDim.x; }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int st = 0; st < kPipeD; ++st) issue(st);
    float mu[8], rs[8];
    moments8(p.stats, n, p.C, c, p.inv_cnt, p.eps, mu, rs);
#pragma unroll
    for (int k = 0; k < 8; ++k) mu[k] = -mu[k] * rs[k];
    float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int st = 0;
    while (cy < p.H) {
        cp_async_wait<kPipeD - 1>();
        const uint4 gr = slot[(st * NT + 0) * nt], zr = slot[(st * NT + 1) * nt];
        uint4 hr = make_uint4(0, 0, 0, 0);
        if (kTwo) hr = slot[(st * NT + 2) * nt];
        issue(st);
        float g[8], zv[8];
        unpack8(gr, g); unpack8(zr, zv);
        if (kTwo) {
            float u[8];
            unpack8(hr, u);
#pragma unroll
            for (int k = 0; k < 8; ++k) g[k] += u[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float xh = fmaf(zv[k], rs[k], mu[k]);
            const float gd = g[k] * dactf(xh, p.act, p.slope);
            s1[k] += gd; s2[k] = fmaf(gd, xh, s2[k]);
        }
        cx += L; if (cx >= p.W) { cx = lane; cy += gridDim.x; }
        st = st + 1 == kPipeD ? 0 : st + 1;
    }
    cp_async_wait<0>();
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) { sh[(lane * C8 + cv) * 16 + j * 2] = s1[j]; sh[(lane * C8 + cv) * 16 + j * 2 + 1] = s2[j]; }
    __syncthreads();
    for (int i = threadIdx.x; i < C8 * 16; i += blockDim.x) {
        float a = 0.f;
        for (int l = 0; l < L; ++l) a += sh[l * C8 * 16 + i];
        p.part[((long long)blockIdx.x * gridDim.y + n) * p.C * 2 + i] = a;
    }
    if (p.counters) finalize_by_last_block(p.part, p.bsum, n, p.C * 2, p.counters);
}

template <bool kTwo>
__global__ void __launch_bounds__(256, 4) in_bwd_apply_pipe_kernel(const InBwdP p) {
    irc::pdl_prologue();
    extern __shared__ float sh[];
    constexpr int NT = kTwo ? 3 : 2;
    const int nt = blockDim.x;
    uint4* slot = reinterpret_cast<uint4*>(sh) + threadIdx.x;
    const int C8 = p.C >> 3;
    const int cv = threadIdx.x % C8, lane = threadIdx.x / C8, L = blockDim.x / C8;
    const int c = cv * 8;
    const int rows = p.n_img * p.H;
    int irow = blockIdx.x, ix = lane, crow = blockIdx.x, cx = lane;
    auto issue = [&](int st) {
        if (irow < rows) {
            const int n = irow / p.H, y = irow - n * p.H;
            cp_async16(slot + (st * NT + 0) * nt, p.g1.at(n, y, ix, c));
            cp_async16(slot + (st * NT + 1) * nt, p.z.at(n, y, ix, c));
            if (kTwo) cp_async16(slot + (st * NT + 2) * nt, p.g2.at(n, y, ix, c));
            ix += L; if (ix >= p.W) { ix = lane; irow += gridDim.x; }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int st = 0; st < kPipeD; ++st) issue(st);
    float mu[8], rs[8], b1[8], b2[8];
    int cur_n = -1, st = 0;
    while (crow < rows) {
        const int n = crow / p.H, y = crow - n * p.H;
        if (n != cur_n && p.stats) {
            cur_n = n;
            moments8(p.stats, n, p.C, c, p.inv_cnt, p.eps, mu, rs);
            const float4* bp = reinterpret_cast<const float4*>(p.bsum + ((long long)n * p.C + c) * 2);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 b = __ldg(bp + q);
                b1[q * 2] = b.x * p.inv_cnt; b2[q * 2] = b.y * p.inv_cnt; b1[q * 2 + 1] = b.z * p.inv_cnt; b2[q * 2 + 1] = b.w * p.inv_cnt;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) mu[k] = -mu[k] * rs[k];
        }
        cp_async_wait<kPipeD - 1>();
        const uint4 gr = slot[(st * NT + 0) * nt], zr = slot[(st * NT + 1) * nt];
        uint4 hr = make_uint4(0, 0, 0, 0);
        if (kTwo) hr = slot[(st * NT + 2) * nt];
        issue(st);
        float g[8], zv[8], o[8];
        unpack8(gr, g); unpack8(zr, zv);
        if (kTwo) {
            float u[8];
            unpack8(hr, u);
#pragma unroll
            for (int k = 0; k < 8; ++k) g[k] += u[k];
        }
        if (p.stats) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float xh = fmaf(zv[k], rs[k], mu[k]);
                const float gd = g[k] * dactf(xh, p.act, p.slope);
                o[k] = rs[k] * (gd - b1[k] - xh * b2[k]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = g[k] * dactf(zv[k], p.act, p.slope);
        }
        store8(const_cast<bf16*>(p.dz.at(n, y, cx, c)), o);
        cx += L; if (cx >= p.W) { cx = lane; crow += gridDim.x; }
        st = st + 1 == kPipeD ? 0 : st + 1;
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------
// L2-resident single-launch InstanceNorm(+activation) backward for LARGE maps (identity tables, plain views).
// The two-pass version reads g and z twice from HBM (5 tensor units moved for 3 algorithmic: both tensors are several
// times the L2).  Here the CTAs form K groups; a group walks its images one at a time: phase 1 streams the image's g and z
// (a few MB: they stay in the 126 MB L2) into the (sum gd, sum gd*xhat) partials, the group meets at a counter barrier,
// every CTA adds the group's partials in a fixed order (bit-reproducible), and phase 2 re-reads the same pixels - now L2
// hits - and writes dz.  Only K images are in flight, so HBM sees g and z once: 3 units.  The grid never exceeds what is
// co-resident (checked on the host with the occupancy API), so the spin barrier cannot starve; it is also bounded.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <bool kTwo, int kD>
__global__ void __launch_bounds__(256, 2) in_bwd_l2_kernel(const InBwdP p, int K, int Gc, float* part, unsigned* sync) {
    irc::pdl_prologue();
    extern __shared__ float sh[];
    constexpr int NT = kTwo ? 3 : 2;
    const int nt = blockDim.x;
    uint4* slot = reinterpret_cast<uint4*>(sh) + threadIdx.x;         // [kD][NT][nt]
    __shared__ float tot[512];                                         // group totals of the current image: [C][2]
    const int C8 = p.C >> 3, C2 = p.C * 2;
    const int cv = threadIdx.x % C8, lane = threadIdx.x / C8, L = blockDim.x / C8;
    const int c = cv * 8;
    const int group = blockIdx.x % K, j = blockIdx.x / K;              // CTA j of its group
    unsigned* bar = sync + group;
    unsigned* exitc = sync + 32 + group;
    const long long grs = (long long)p.g1.wp * p.g1.ld, zrs = (long long)p.z.wp * p.z.ld, drs = (long long)p.dz.wp * p.dz.ld;
    const long long hrs = kTwo ? (long long)p.g2.wp * p.g2.ld : 0;
    const int gps = (int)p.g1.ld, zps = (int)p.z.ld, dps = (int)p.dz.ld, hps = kTwo ? (int)p.g2.ld : 0;
    unsigned it = 0;
    for (int n = group; n < p.n_img; n += K, ++it) {
        const bf16* gb = p.g1.at(n, 0, 0, c);
        const bf16* zb = p.z.at(n, 0, 0, c);
        const bf16* hb = kTwo ? p.g2.at(n, 0, 0, c) : nullptr;
        bf16* db = const_cast<bf16*>(p.dz.at(n, 0, 0, c));
        float mu[8], rs[8];
        moments8(p.stats, n, p.C, c, p.inv_cnt, p.eps, mu, rs);
#pragma unroll
        for (int k = 0; k < 8; ++k) mu[k] = -mu[k] * rs[k];
        // ---------------- phase 1: partial sums over this CTA's rows (j, j + Gc, ...)
        float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        {
            int iy = j, ix = lane, cy = j, cx = lane;
            auto issue = [&](int st) {
                if (iy < p.H) {
                    cp_async16(slot + (st * NT + 0) * nt, gb + iy * grs + ix * gps);
                    cp_async16(slot + (st * NT + 1) * nt, zb + iy * zrs + ix * zps);
                    if (kTwo) cp_async16(slot + (st * NT + 2) * nt, hb + iy * hrs + ix * hps);
                    ix += L; if (ix >= p.W) { ix = lane; iy += Gc; }
                }
                cp_async_commit();
            };
#pragma unroll
            for (int st = 0; st < kD; ++st) issue(st);
            int st = 0;
            while (cy < p.H) {
                cp_async_wait<kD - 1>();
                const uint4 gr = slot[(st * NT + 0) * nt], zr = slot[(st * NT + 1) * nt];
                uint4 hr = make_uint4(0, 0, 0, 0);
                if (kTwo) hr = slot[(st * NT + 2) * nt];
                issue(st);
                float g[8], zv[8];
                unpack8(gr, g); unpack8(zr, zv);
                if (kTwo) {
                    float u[8];
                    unpack8(hr, u);
#pragma unroll
                    for (int k = 0; k < 8; ++k) g[k] += u[k];
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float xh = fmaf(zv[k], rs[k], mu[k]);
                    const float gd = g[k] * dactf(xh, p.act, p.slope);
                    s1[k] += gd; s2[k] = fmaf(gd, xh, s2[k]);
                }
                cx += L; if (cx >= p.W) { cx = lane; cy += Gc; }
                st = st + 1 == kD ? 0 : st + 1;
            }
            cp_async_wait<0>();
        }
        __syncthreads();                                  // slots are reused as the lane-reduction scratch
#pragma unroll
        for (int q = 0; q < 8; ++q) { sh[(lane * C8 + cv) * 16 + q * 2] = s1[q]; sh[(lane * C8 + cv) * 16 + q * 2 + 1] = s2[q]; }
        __syncthreads();
        float* mine = part + ((long long)((it & 1) * K + group) * Gc + j) * C2;
        for (int i = threadIdx.x; i < C2; i += blockDim.x) {
            float a = 0.f;
            for (int l = 0; l < L; ++l) a += sh[l * C8 * 16 + i];
            __stcg(mine + i, a);
        }
        // ---------------- group barrier (monotonic counter; all CTAs of the grid are co-resident)
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            atomicAdd(bar, 1u);
            const unsigned target = (it + 1) * (unsigned)Gc;
            const long long t0 = clock64();
            while (ld_acquire_u32(bar) < target) {
                __nanosleep(64);
                if (clock64() - t0 > 4000000000LL) { printf("irc: in_bwd_l2 group barrier timed out (block %d)\n", (int)blockIdx.x); __trap(); }
            }
        }
        __syncthreads();
        // every CTA adds the group's partials itself, in a fixed order: the 256 threads split the Gc partials of each value
        // into 256 / C2 interleaved subsets (independent loads, 8 in flight), the subsets are combined in subset order
        const float* gp = part + (long long)((it & 1) * K + group) * Gc * C2;
        {
            const int nh = blockDim.x / C2 > 0 ? blockDim.x / C2 : 1;      // 2 (C = 64), 1 (C >= 128)
            const int i0 = threadIdx.x % C2, h = threadIdx.x / C2;
            for (int i = i0; i < C2; i += blockDim.x) {                     // one trip unless C2 > 256
                float a = 0.f;
                int q = h;
                for (; q + 7 * nh < Gc; q += 8 * nh) {
                    float v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = __ldcg(gp + (long long)(q + u * nh) * C2 + i);
#pragma unroll
                    for (int u = 0; u < 8; ++u) a += v[u];
                }
                for (; q < Gc; q += nh) a += __ldcg(gp + (long long)q * C2 + i);
                sh[h * C2 + i] = a;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < C2; i += blockDim.x) {
                float a = 0.f;
                for (int hh = 0; hh < nh; ++hh) a += sh[hh * C2 + i];
                tot[i] = a;
                if (j == 0 && p.bsum) p.bsum[(long long)n * C2 + i] = a;
            }
        }
        __syncthreads();
        float b1[8], b2[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { b1[k] = tot[(c + k) * 2] * p.inv_cnt; b2[k] = tot[(c + k) * 2 + 1] * p.inv_cnt; }
        // ---------------- phase 2: the same pixels again (L2 hits), dz out
        {
            int iy = j, ix = lane, cy = j, cx = lane;
            auto issue = [&](int st) {
                if (iy < p.H) {
                    cp_async16(slot + (st * NT + 0) * nt, gb + iy * grs + ix * gps);
                    cp_async16(slot + (st * NT + 1) * nt, zb + iy * zrs + ix * zps);
                    if (kTwo) cp_async16(slot + (st * NT + 2) * nt, hb + iy * hrs + ix * hps);
                    ix += L; if (ix >= p.W) { ix = lane; iy += Gc; }
                }
                cp_async_commit();
            };
#pragma unroll
            for (int st = 0; st < kD; ++st) issue(st);
            int st = 0;
            while (cy < p.H) {
                cp_async_wait<kD - 1>();
                const uint4 gr = slot[(st * NT + 0) * nt], zr = slot[(st * NT + 1) * nt];
                uint4 hr = make_uint4(0, 0, 0, 0);
                if (kTwo) hr = slot[(st * NT + 2) * nt];
                issue(st);
                float g[8], zv[8], o[8];
                unpack8(gr, g); unpack8(zr, zv);
                if (kTwo) {
                    float u[8];
                    unpack8(hr, u);
#pragma unroll
                    for (int k = 0; k < 8; ++k) g[k] += u[k];
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float xh = fmaf(zv[k], rs[k], mu[k]);
                    const float gd = g[k] * dactf(xh, p.act, p.slope);
                    o[k] = rs[k] * (gd - b1[k] - xh * b2[k]);
                }
                store8(db + cy * drs + cx * dps, o);
                cx += L; if (cx >= p.W) { cx = lane; cy += Gc; }
                st = st + 1 == kD ? 0 : st + 1;
            }
            cp_async_wait<0>();
        }
        __syncthreads();                                  // `tot` and the slots are rewritten by the next image
    }
    // re-arm the counters for the next launch: the last CTA of the group to leave resets them
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(exitc, 1u) == (unsigned)Gc - 1) { *bar = 0u; *exitc = 0u; __threadfence(); }
    }
}

// ---------------------------------------------------------------------------------
// Single-pass InstanceNorm(+activation) backward for small feature maps (H*W <= 4096: the ResNet bottleneck and the
// PatchGAN layers).  One thread-block CLUSTER owns one (image, 32-channel group): its CTAs split the pixels, every thread
// keeps its <= 8 pixels of g and z (16 bytes each) in REGISTERS, the (sum gd, sum gd*xhat) partials are combined inside
// the CTA by shuffles + shared memory and across the cluster through distributed shared memory (fixed order:
// bit-reproducible), and dz is produced from the registers.  g and z are read once, dz written once: 3 passes over the
// tensor instead of 5, one launch instead of two; with fold_pad > 0 the transpose of ReflectionPad2d(fold_pad) is applied
// while g is loaded (border pixels add the ring pixels that mirror onto them), replacing a third launch.
// ---------------------------------------------------------------------------------
constexpr int kFusedNP = 8;      // pixels per thread
constexpr int kFusedCC = 32;     // channels per cluster
constexpr int kFusedLanes = 64;  // pixel lanes per CTA (256 threads = 64 lanes x 4 channel vectors)

__device__ __forceinline__ int mirror_src(int v, int n, int p) {
    // interior coordinate v of an axis of length n: the ring coordinate (in interior units, < 0 or >= n) that
    // ReflectionPad2d(p) mirrors onto v, or INT_MIN when there is none
    if (v >= 1 && v <= p) return -v;
    if (v >= n - 1 - p && v <= n - 2) return 2 * (n - 1) - v;
    return INT_MIN;
}


// Sum 16 per-thread values over the 8 lanes of a warp that share (lane & 3) - lane bits 2..4 - as a reduce-scatter:
// every step halves the number of values a thread carries (8 + 4 + 2 shuffles instead of 3 x 16).  On return the thread
// holds the complete sums of values idx and idx + 1 in (r0, r1).  Fixed order: bit-reproducible.
__device__ __forceinline__ void reduce16_scatter(const float (&v)[16], float& r0, float& r1, int& idx) {
    const int l = threadIdx.x & 31;
    const bool b4 = l & 16, b3 = l & 8, b2 = l & 4;
    float a[8], b[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float keep = b4 ? v[8 + j] : v[j], give = b4 ? v[j] : v[8 + j];
        a[j] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float keep = b3 ? a[4 + j] : a[j], give = b3 ? a[j] : a[4 + j];
        b[j] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
    }
    {
        const float k0 = b2 ? b[2] : b[0], g0 = b2 ? b[0] : b[2], k1 = b2 ? b[3] : b[1], g1 = b2 ? b[1] : b[3];
        r0 = k0 + __shfl_xor_sync(0xffffffffu, g0, 4);
        r1 = k1 + __shfl_xor_sync(0xffffffffu, g1, 4);
    }
    idx = (b4 ? 8 : 0) + (b3 ? 4 : 0) + (b2 ? 2 : 0);
}

// CTA + cluster stage of the fused kernels: v = this thread's 16 partial sums, layout [k][2] for its channel vector.
// Returns with tot[cv*16 + j] = sum over the whole cluster.  part must stay alive until the final cluster.sync().
__device__ __forceinline__ void cluster_sum16(cg::cluster_group& cluster, const float (&v)[16], float (*wsum)[64], float* part, float* tot) {
    const int tid = threadIdx.x, cv = tid & 3;
    float r0, r1; int idx;
    reduce16_scatter(v, r0, r1, idx);
    wsum[tid >> 5][cv * 16 + idx] = r0; wsum[tid >> 5][cv * 16 + idx + 1] = r1;
    __syncthreads();
    if (tid < 64) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) a += wsum[w][tid];
        part[tid] = a;
    }
    cluster.sync();
    if (tid < 64) {
        const unsigned CL = cluster.num_blocks();
        float a = 0.f;
        for (unsigned r = 0; r < CL; ++r) a += cluster.map_shared_rank(part, r)[tid];
        tot[tid] = a;
    }
    __syncthreads();
}

// pixel walk of a thread of the fused kernels: pixel p0 + lane + i*64 as (y, x) plus the 32-bit element offsets of that
// pixel in up to three (non space-to-depth) views, advanced without divisions or 64-bit multiplies
struct VStep {
    const bf16* base; int dx, wrap;      // pointer of pixel (0,0), elements per x step, extra elements when a row wraps
    __device__ __forceinline__ void init(const View& v, int n, int c, int W) { base = v.at(n, 0, 0, c); dx = (int)v.ld; wrap = (v.wp - W) * (int)v.ld; }
    __device__ __forceinline__ int at(int y, int x, int W) const { return y * (W * dx + wrap) + x * dx; }
};
struct PixWalk {
    int y, x, W, o0, o1, o2;
    __device__ __forceinline__ PixWalk(int pix, int W_, const VStep& a, const VStep& b, const VStep& c) : W(W_) {
        y = pix / W_; x = pix - y * W_;
        o0 = a.at(y, x, W_); o1 = b.at(y, x, W_); o2 = c.at(y, x, W_);
    }
    __device__ __forceinline__ void next(const VStep& a, const VStep& b, const VStep& c) {
        x += kFusedLanes; o0 += kFusedLanes * a.dx; o1 += kFusedLanes * b.dx; o2 += kFusedLanes * c.dx;
        while (x >= W) { x -= W; ++y; o0 += a.wrap; o1 += b.wrap; o2 += c.wrap; }
    }
};

template <int kAct>
__device__ __forceinline__ float dact_t(float xh, float slope) {
    if (kAct == 1) return xh > 0.f ? 1.f : 0.f;
    if (kAct == 2) return xh > 0.f ? 1.f : slope;
    return 1.f;
}

// Pixel i (0..7) of a thread: group i covers the 64 consecutive pixels [p0 + 64 i, p0 + 64 i + 64) and the thread's lane is
// ROTATED by 8 per group.  With a fixed lane the threads that sit on a mirrored column (x = 1, x = W - 2 when W == 64) would
// carry the reflection fold of all eight of their pixels - eight dependent trips to L2 while everyone else waits at the
// cluster barrier; rotated, every thread owns at most one or two border pixels.  Any permutation inside a group keeps the
// accesses of a warp contiguous (8 consecutive lanes, possibly wrapped once).
struct RotPix {
    int y, x, o0, o1, o2;
    bool ok;
    __device__ __forceinline__ RotPix(int i, int p0, int p1, int lane, int W, float invW, const VStep& a, const VStep& b, const VStep& c) {
        const int pix = p0 + i * kFusedLanes + ((lane + 8 * i) & (kFusedLanes - 1));
        ok = pix < p1;
        y = __float2int_rd(((float)pix + 0.5f) * invW);        // exact for pix < 2^20 (H * W <= 4096 here)
        x = pix - y * W;
        o0 = a.at(y, x, W); o1 = b.at(y, x, W); o2 = c.at(y, x, W);
    }
};

template <bool kFold, int kAct>
__global__ void __launch_bounds__(256, 3) in_bwd_fused_kernel(const InBwdP p, int fold_pad) {
    irc::pdl_prologue();
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned CL = cluster.num_blocks(), rank = cluster.block_rank();
    extern __shared__ uint4 raw[];      // [2][kFusedNP][256]: g then z, slot [i][tid] is private to thread tid
    __shared__ float part[64];          // this CTA's sums: [cv][k][2]
    __shared__ float wsum[8][64];
    __shared__ float tot[64];
    const int tid = threadIdx.x, cv = tid & 3, lane = tid >> 2;
    const int n = blockIdx.z, c = blockIdx.y * kFusedCC + cv * 8;
    const int HW = p.H * p.W;
    const int P = (HW + CL - 1) / CL;
    const int p0 = rank * P, p1 = min(p0 + P, HW);
    const float invW = 1.f / (float)p.W;
    uint4* gs = raw + tid;
    uint4* zs = raw + kFusedNP * 256 + tid;
    // all loads of the thread go straight to its shared-memory slots (no register staging): 16 x 16 bytes in flight
    VStep vg, vz, vd;
    vg.init(p.g1, n, c, p.W); vz.init(p.z, n, c, p.W); vd.init(p.dz, n, c, p.W);
#pragma unroll
    for (int i = 0; i < kFusedNP; ++i) {
        const RotPix w(i, p0, p1, lane, p.W, invW, vg, vz, vd);
        if (w.ok) {
            cp_async16(gs + i * 256, vg.base + w.o0);
            cp_async16(zs + i * 256, vz.base + w.o1);
        } else {
            gs[i * 256] = make_uint4(0, 0, 0, 0); zs[i * 256] = make_uint4(0, 0, 0, 0);     // g = 0 adds nothing
        }
    }
    float mu[8], rs[8];
    moments8(p.stats, n, p.C, c, p.inv_cnt, p.eps, mu, rs);
#pragma unroll
    for (int k = 0; k < 8; ++k) mu[k] = -mu[k] * rs[k];
    if (kFold) {
        // transpose of ReflectionPad2d(fold_pad): border pixels add the ring pixels that mirror onto them.  The (up to three)
        // ring reads of a pixel are independent and issued together, while the cp.async of the interior are still in flight.
#pragma unroll 2
        for (int i = 0; i < kFusedNP; ++i) {
            const RotPix w(i, p0, p1, lane, p.W, invW, vg, vz, vd);
            const int y = w.y, x = w.x;
            if (!w.ok || (y > fold_pad && y < p.H - 1 - fold_pad && x > fold_pad && x < p.W - 1 - fold_pad)) continue;
            const int my = mirror_src(y, p.H, fold_pad), mx = mirror_src(x, p.W, fold_pad);
            if (my == INT_MIN && mx == INT_MIN) continue;
            const uint4 zero = make_uint4(0, 0, 0, 0);
            const uint4 u0 = my != INT_MIN ? __ldg(reinterpret_cast<const uint4*>(p.g1.at(n, my, x, c))) : zero;
            const uint4 u1 = mx != INT_MIN ? __ldg(reinterpret_cast<const uint4*>(p.g1.at(n, y, mx, c))) : zero;
            const uint4 u2 = (my != INT_MIN && mx != INT_MIN) ? __ldg(reinterpret_cast<const uint4*>(p.g1.at(n, my, mx, c))) : zero;
            cp_async_wait_all();              // the thread's own interior values (a no-op after the first border pixel)
            float a[8], v[8];
            unpack8(gs[i * 256], a);
            unpack8(u0, v);
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] += v[k];
            unpack8(u1, v);
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] += v[k];
            unpack8(u2, v);
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] += v[k];
            // rounded to bf16 like the stand-alone fold, so both paths give the same bits
            gs[i * 256] = make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
        }
    }
    cp_async_wait_all();
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
#pragma unroll 2
    for (int i = 0; i < kFusedNP; ++i) {
        float g[8], zv[8];
        unpack8(gs[i * 256], g); unpack8(zs[i * 256], zv);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float xh = fmaf(zv[k], rs[k], mu[k]);
            const float gd = g[k] * dact_t<kAct>(xh, p.slope);
            acc[2 * k] += gd; acc[2 * k + 1] = fmaf(gd, xh, acc[2 * k + 1]);
        }
    }
    cluster_sum16(cluster, acc, wsum, part, tot);
    if (rank == 0 && tid < 64 && p.bsum) p.bsum[((long long)n * p.C + blockIdx.y * kFusedCC) * 2 + tid] = tot[tid];
    float b1[8], b2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { b1[k] = tot[cv * 16 + k * 2] * p.inv_cnt; b2[k] = tot[cv * 16 + k * 2 + 1] * p.inv_cnt; }
#pragma unroll 2
    for (int i = 0; i < kFusedNP; ++i) {
        const RotPix w(i, p0, p1, lane, p.W, invW, vg, vz, vd);
        if (w.ok) {
            float g[8], zv[8], o[8];
            unpack8(gs[i * 256], g); unpack8(zs[i * 256], zv);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float xh = fmaf(zv[k], rs[k], mu[k]);
                const float gd = g[k] * dact_t<kAct>(xh, p.slope);
                o[k] = rs[k] * (gd - b1[k] - xh * b2[k]);
            }
            store8(const_cast<bf16*>(vd.base) + w.o2, o);
        }
    }
    cluster.sync();      // `part` must outlive the remote reads of every peer
}

// ---------------------------------------------------------------------------------
// Forward twin of the kernel above: InstanceNorm statistics + normalise + activation (+ residual) + padding ring in ONE
// pass for small maps.  The cluster of an (image, 32-channel group) loads z once (cp.async into per-thread shared-memory
// slots), reduces (sum, sum of squares) through distributed shared memory in a fixed order, writes the statistics the
// backward pass needs, and produces the activated frame - interior and ring - from the staged values.
// ---------------------------------------------------------------------------------
template <int kAct>
__global__ void __launch_bounds__(256, 4) in_apply_fused_kernel(const GatherP p, float* stats_out) {
    irc::pdl_prologue();
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned CL = cluster.num_blocks(), rank = cluster.block_rank();
    extern __shared__ uint4 raw[];      // [kFusedNP][256]: z
    __shared__ float part[64];
    __shared__ float wsum[8][64];
    __shared__ float tot[64];
    const int tid = threadIdx.x, cv = tid & 3, lane = tid >> 2;
    const int n = blockIdx.z, c = blockIdx.y * kFusedCC + cv * 8;
    const int HW = p.H * p.W;
    const int P = (HW + CL - 1) / CL;
    const int p0 = rank * P, p1 = min(p0 + P, HW);
    const int np = p0 + lane < p1 ? (p1 - p0 - lane + kFusedLanes - 1) / kFusedLanes : 0;
    uint4* zs = raw + tid;
    const int pad = p.pad, W = p.W, H = p.H;
    VStep vs, vr, vd;
    vs.init(p.src, n, c, W); vr.init(p.has_res ? p.res : p.src, n, c, W); vd.init(p.dst, n, c, W);
    {
        PixWalk w(p0 + lane, W, vs, vr, vd);
#pragma unroll
        for (int i = 0; i < kFusedNP; ++i) {
            if (i < np) {
                cp_async16(zs + i * 256, vs.base + w.o0);
                // the residual is not staged (that would cost the fourth CTA per SM): pulled towards L2 now, read in the apply loop
                if (p.has_res) asm volatile("prefetch.global.L2 [%0];" ::"l"(vr.base + w.o1));
            } else {
                zs[i * 256] = make_uint4(0, 0, 0, 0);
            }
            w.next(vs, vr, vd);
        }
    }
    cp_async_wait_all();
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
#pragma unroll 2
    for (int i = 0; i < kFusedNP; ++i) {
        float v[8];
        unpack8(zs[i * 256], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc[2 * k] += v[k]; acc[2 * k + 1] = fmaf(v[k], v[k], acc[2 * k + 1]); }
    }
    cluster_sum16(cluster, acc, wsum, part, tot);
    if (rank == 0 && tid < 64) stats_out[((long long)n * p.C + blockIdx.y * kFusedCC) * 2 + tid] = tot[tid];     // (sum, sum of squares) pairs
    float mu[8], rs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float m = tot[cv * 16 + k * 2] * p.inv_cnt;
        rs[k] = rsqrtf(fmaxf(tot[cv * 16 + k * 2 + 1] * p.inv_cnt - m * m, 0.f) + p.eps);
        mu[k] = -m * rs[k];
    }
    bf16* dbase = const_cast<bf16*>(vd.base);          // interior pixel (0,0) of the destination view
    const int dld = (int)p.dst.ld, dstep = p.dst.wp * dld;
    PixWalk w(p0 + lane, W, vs, vr, vd);
#pragma unroll 2
    for (int i = 0; i < kFusedNP; ++i, w.next(vs, vr, vd)) {
        if (i >= np) continue;
        const int y = w.y, x = w.x;
        float v[8];
        unpack8(zs[i * 256], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float t = fmaf(v[k], rs[k], mu[k]);
            v[k] = kAct == 1 ? fmaxf(t, 0.f) : (kAct == 2 ? fmaxf(t, 0.f) + p.slope * fminf(t, 0.f) : t);
        }
        if (p.has_res) {
            float u[8];
            load8(vr.base + w.o1, u);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] += u[k];
        }
        const uint4 val = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        bf16* dpix = dbase + w.o2;
        *reinterpret_cast<uint4*>(dpix) = val;
        if (pad && (y <= pad || y >= H - 1 - pad || x <= pad || x >= W - 1 - pad)) {
            // ring pixels owned by this border pixel: the reflected copies (halo_mode 1) or zeros (halo_mode 0)
            const uint4 ring = p.halo_mode == 1 ? val : make_uint4(0, 0, 0, 0);
            int eX, eY;
            if (p.halo_mode == 1) {
                eX = (x >= 1 && x <= pad) ? pad - x : ((x >= W - 1 - pad && x <= W - 2) ? pad + 2 * (W - 1) - x : -1);
                eY = (y >= 1 && y <= pad) ? pad - y : ((y >= H - 1 - pad && y <= H - 2) ? pad + 2 * (H - 1) - y : -1);
            } else {
                eX = x < pad ? x : (x >= W - pad ? x + 2 * pad : -1);
                eY = y < pad ? y : (y >= H - pad ? y + 2 * pad : -1);
            }
            // eX / eY are padded coordinates; the interior pixel sits at padded (y + pad, x + pad)
            const int ddx = (eX - (x + pad)) * dld, ddy = (eY - (y + pad)) * dstep;
            if (eX >= 0) *reinterpret_cast<uint4*>(dpix + ddx) = ring;
            if (eY >= 0) {
                *reinterpret_cast<uint4*>(dpix + ddy) = ring;
                if (eX >= 0) *reinterpret_cast<uint4*>(dpix + ddy + ddx) = ring;
            }
        }
    }
    cluster.sync();
}

// ---------------------------------------------------------------------------------
// 2x2 max pool (VGG trunk, irc:664) on frames, and its backward fused with the ReLU mask
// ---------------------------------------------------------------------------------
// One block row = one output row of one image (block-uniform decode, no per-element division); thread = (channel vector, pixel
// lane); two output pixels per step with their eight 16-byte loads issued together.  Non space-to-depth views only.
__global__ void __launch_bounds__(256, 4) maxpool_kernel(View src, View dst, int C, int n_img, int Ho, int Wo) {
    irc::pdl_prologue();
    const int C8 = C >> 3;
    const int cv = threadIdx.x % C8, lane = threadIdx.x / C8, L = blockDim.x / C8;
    const int c = cv * 8;
    const int sld = (int)src.ld, dld = (int)dst.ld;
    const int srow = src.wp * sld;
    for (int row = blockIdx.x; row < n_img * Ho; row += gridDim.x) {
        const int n = row / Ho, y = row - n * Ho;
        const bf16* s0 = src.at(n, 2 * y, 0, c);
        bf16* d0 = const_cast<bf16*>(dst.at(n, y, 0, c));
        for (int x0 = lane; x0 < Wo; x0 += 2 * L) {
            uint4 u[2][4];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int x = x0 + q * L;
                if (x < Wo) {
                    const bf16* sp = s0 + 2 * x * sld;
                    u[q][0] = __ldg(reinterpret_cast<const uint4*>(sp));
                    u[q][1] = __ldg(reinterpret_cast<const uint4*>(sp + sld));
                    u[q][2] = __ldg(reinterpret_cast<const uint4*>(sp + srow));
                    u[q][3] = __ldg(reinterpret_cast<const uint4*>(sp + srow + sld));
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int x = x0 + q * L;
                if (x < Wo) {
                    float m[8], v[8];
                    unpack8(u[q][0], m);
#pragma unroll
                    for (int t = 1; t < 4; ++t) {
                        unpack8(u[q][t], v);
#pragma unroll
                        for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], v[k]);
                    }
                    store8(d0 + x * dld, m);
                }
            }
        }
    }
}
// dsrc[n, y, x] = g[n, y/2, x/2] if src[n,y,x] is the first maximum of its window and src > 0, else 0
__global__ void __launch_bounds__(256, 4) maxpool_bwd_kernel(View src, View g, View dsrc, int C, int n_img, int Ho, int Wo) {
    irc::pdl_prologue();
    const int C8 = C >> 3;
    const int cv = threadIdx.x % C8, lane = threadIdx.x / C8, L = blockDim.x / C8;
    const int c = cv * 8;
    const int sld = (int)src.ld, gld = (int)g.ld, dld = (int)dsrc.ld;
    const int srow = src.wp * sld, drow = dsrc.wp * dld;
    for (int row = blockIdx.x; row < n_img * Ho; row += gridDim.x) {
        const int n = row / Ho, y = row - n * Ho;
        const bf16* s0 = src.at(n, 2 * y, 0, c);
        const bf16* g0 = g.at(n, y, 0, c);
        bf16* d0 = const_cast<bf16*>(dsrc.at(n, 2 * y, 0, c));
        for (int x = lane; x < Wo; x += L) {
            const bf16* sp = s0 + 2 * x * sld;
            const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(sp)), u1 = __ldg(reinterpret_cast<const uint4*>(sp + sld));
            const uint4 u2 = __ldg(reinterpret_cast<const uint4*>(sp + srow)), u3 = __ldg(reinterpret_cast<const uint4*>(sp + srow + sld));
            const uint4 ug = __ldg(reinterpret_cast<const uint4*>(g0 + x * gld));
            float v[4][8], gv[8];
            unpack8(u0, v[0]); unpack8(u1, v[1]); unpack8(u2, v[2]); unpack8(u3, v[3]); unpack8(ug, gv);
            int arg[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                int a = 0; float m = v[0][k];
#pragma unroll
                for (int t = 1; t < 4; ++t) if (v[t][k] > m) { m = v[t][k]; a = t; }
                arg[k] = m > 0.f ? a : -1;
            }
            bf16* dp = d0 + 2 * x * dld;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float o[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = arg[k] == t ? gv[k] : 0.f;
                store8(dp + (t >> 1) * drow + (t & 1) * dld, o);
            }
        }
    }
}

// column sums of a bf16 [rows][ld] matrix slice -> fp32 [C] (bias gradients)
__global__ void colsum_kernel(const bf16* a, long long rows, long long ld, int off, int C, const short* row_img, float* part) {
    irc::pdl_prologue();
    const int c = blockIdx.y * 32 + (threadIdx.x & 31);
    const int lane_r = threadIdx.x >> 5, R = blockDim.x >> 5;
    float s = 0.f;
    if (c < C)
        for (long long r = blockIdx.x * (long long)R + lane_r; r < rows; r += (long long)gridDim.x * R)
            if (!row_img || row_img[r] >= 0) s += __bfloat162float(a[r * ld + off + c]);
    __shared__ float sh[32][33];
    sh[lane_r][threadIdx.x & 31] = s;
    __syncthreads();
    if (lane_r == 0 && c < C) {
        float t = 0.f;
        for (int i = 0; i < R; ++i) t += sh[i][threadIdx.x & 31];
        part[(long long)blockIdx.x * C + c] = t;
    }
}

// Vectorised form (C, off, ld multiples of 8): thread = (channel vector, row lane), 16-byte loads, four rows in flight per thread;
// the row lanes of a block are combined through shared memory in a fixed order.  (The scalar kernel above reads two bytes per
// thread and row: 54 us for the 69 MB of model.0's bias gradient.)
__global__ void __launch_bounds__(256) colsum_vec_kernel(const bf16* __restrict__ a, long long rows, long long ld, int off, int C,
                                                        const short* __restrict__ row_img, float* __restrict__ part) {
    irc::pdl_prologue();
    extern __shared__ float cs_sh[];          // [L][C]
    const int C8 = C >> 3;
    const int cv = threadIdx.x % C8, lane = threadIdx.x / C8, L = blockDim.x / C8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const long long step = (long long)gridDim.x * L;
    const bf16* base = a + off + cv * 8;
    for (long long r0 = blockIdx.x * (long long)L + lane; r0 < rows; r0 += 4 * step) {
        uint4 u[4]; bool on[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long long r = r0 + q * step;
            on[q] = r < rows && (!row_img || __ldg(row_img + r) >= 0);
            if (on[q]) u[q] = __ldg(reinterpret_cast<const uint4*>(base + r * ld));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (on[q]) {
                float v[8];
                unpack8(u[q], v);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += v[k];
            }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) cs_sh[lane * C + cv * 8 + k] = acc[k];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float t = 0.f;
        for (int i = 0; i < L; ++i) t += cs_sh[i * C + c];
        part[(long long)blockIdx.x * C + c] = t;
    }
}

// In-place transpose of ReflectionPad2d(p) on a frame that holds the gradient w.r.t. the padded tensor: every interior
// pixel within p of the border adds the ring pixels that mirror onto it (ring pixels are only read, each thread writes
// its own pixel: race-free) and clears them.  Touches O(p * perimeter) pixels instead of a full pass.
__global__ void fold_inplace_kernel(bf16* g, long long ld, int off, int C, int n_img, int H, int W, int p) {
    irc::pdl_prologue();
    // Only pixels within p of the border receive anything: the 2p rows next to the top / bottom border (all columns) and,
    // on every other row, the 2p columns next to the side borders.  grid = (chunks of that work list, images); one pixel
    // per thread, so the pass costs one load-add-store round trip.
    const int C8 = C >> 3;
    const int cv = threadIdx.x % C8, lane = threadIdx.x / C8, L = blockDim.x / C8;
    const int c = cv * 8;
    const int Hp = H + 2 * p, Wp = W + 2 * p;
    const int n = blockIdx.y;
    const int NA = 2 * p * W, NB = (H - 2 * p) * 2 * p;
    const int item = blockIdx.x * L + lane;
    if (item >= NA + NB) return;
    int y, x;
    if (item < NA) {
        const int r = item / W;
        x = item - r * W;
        y = r < p ? 1 + r : H - 1 - p + (r - p);
    } else {
        const int j = item - NA, jr = j / (2 * p), k = j - jr * 2 * p;
        y = jr == 0 ? 0 : (jr == H - 2 * p - 1 ? H - 1 : p + jr);        // the rows that are not among the 2p border rows
        x = k < p ? 1 + k : W - 1 - p + (k - p);
    }
    const int my = (y >= 1 && y <= p) ? p - y : ((y >= H - 1 - p && y <= H - 2) ? 2 * (H - 1) - y + p : -1);   // mirrored padded row
    const int mx = (x >= 1 && x <= p) ? p - x : ((x >= W - 1 - p && x <= W - 2) ? 2 * (W - 1) - x + p : -1);
    if (my < 0 && mx < 0) return;
    bf16* base = g + (long long)n * Hp * Wp * ld + off + c;
    float acc[8], v[8];
    bf16* self = base + ((long long)(y + p) * Wp + (x + p)) * ld;
    // every ring pixel mirrors onto exactly one interior pixel, so the thread that consumes it also clears it
    // (the folded frame can then serve as a zero-ring addend / operand).  All (up to four) loads go out together; the
    // additions keep the order self + row mirror + column mirror + corner.
    bf16* q0 = my >= 0 ? base + ((long long)my * Wp + (x + p)) * ld : nullptr;
    bf16* q1 = mx >= 0 ? base + ((long long)(y + p) * Wp + mx) * ld : nullptr;
    bf16* q2 = (my >= 0 && mx >= 0) ? base + ((long long)my * Wp + mx) * ld : nullptr;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    const uint4 us = *reinterpret_cast<const uint4*>(self);
    const uint4 u0 = q0 ? *reinterpret_cast<const uint4*>(q0) : zero4;
    const uint4 u1 = q1 ? *reinterpret_cast<const uint4*>(q1) : zero4;
    const uint4 u2 = q2 ? *reinterpret_cast<const uint4*>(q2) : zero4;
    unpack8(us, acc);
    unpack8(u0, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += v[k];
    unpack8(u1, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += v[k];
    unpack8(u2, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += v[k];
    if (q0) *reinterpret_cast<uint4*>(q0) = zero4;
    if (q1) *reinterpret_cast<uint4*>(q1) = zero4;
    if (q2) *reinterpret_cast<uint4*>(q2) = zero4;
    store8(self, acc);
}

int grid_for(long long total, int threads) {
    long long b = (total + threads - 1) / threads;
    const long long cap = (long long)irc_num_sms() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

int check_view(const irc_view& v, const char* what) {
    if (!v.ptr) return irc_set_error(IRC_ERR_BAD_ARG, "%s: null view", what);
    if (((uintptr_t)v.ptr & 15) || (v.ld % 8) || (v.chan_off % 8)) return irc_set_error(IRC_ERR_BAD_ARG, "%s: view must be 16-byte aligned (ld %% 8, chan_off %% 8)", what);
    return IRC_OK;
}

// block shape for the per-(n,c) reductions
constexpr int kCounterFloats = 512;     // tail of the reduction workspace: 256 per-image ticket counters + 256 words of group-barrier state

// block = (C/8 channel vectors) x L pixel lanes
void row_block(int C, int W, int& threads, int& L) {
    const int C8 = C / 8;
    L = 256 / C8; if (L < 1) L = 1;
    if (L > W) L = W;
    threads = L * C8;
}

void reduce_shape(int C, int H, int W, int n_img, long long work_floats, int& threads, int& L, int& chunks, size_t& smem) {
    const int C8 = C / 8;
    row_block(C, W, threads, L);
    long long want = ((long long)irc_num_sms() * 4 + n_img - 1) / n_img;
    long long maxc = H;
    chunks = (int)(want < maxc ? want : maxc);
    const long long cap = work_floats / ((long long)n_img * C * 2);
    if (chunks > cap) chunks = (int)cap;
    if (chunks < 1) chunks = 1;
    smem = (size_t)L * C8 * 16 * sizeof(float);
}

// ---------------------------------------------------------------------------------
// nn.BatchNorm2d (norm='batch', irc:158-159) on top of the InstanceNorm kernels
// ---------------------------------------------------------------------------------
// Forward: y = gamma * (x - mean_B) * rstd_B + beta = (x - mean_eff) * rs_eff with rs_eff = gamma * rstd_B and
// mean_eff = mean_B - beta / rs_eff, the same for every image of a batch-statistics group.  One thread per channel turns the
// per-image (sum, sum of squares) into those effective moments, and updates the running statistics as PyTorch does
// (momentum, unbiased variance), `updates` times (the reference runs the generator twice per iteration on the same batch).
__global__ void bn_finalize_kernel(const float* __restrict__ stats, int n_img, int group, int C, float cnt_per_img, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* running_mean, float* running_var, float momentum, float eps, int training,
                                   int updates, float* __restrict__ eff) {
    irc::pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float M = (float)group * cnt_per_img;
    float rm = running_mean[c], rv = running_var[c];
    for (int g0 = 0; g0 < n_img; g0 += group) {
        float mean, var;
        if (training) {
            float s1 = 0.f, s2 = 0.f;
            for (int n = g0; n < g0 + group; ++n) { s1 += stats[((long long)n * C + c) * 2]; s2 += stats[((long long)n * C + c) * 2 + 1]; }
            mean = s1 / M; var = fmaxf(s2 / M - mean * mean, 0.f);
            for (int u = 0; u < updates; ++u) {
                rm = (1.f - momentum) * rm + momentum * mean;
                rv = (1.f - momentum) * rv + momentum * var * (M / fmaxf(M - 1.f, 1.f));
            }
        } else { mean = rm; var = rv; }
        float sc = gamma[c] * rsqrtf(var + eps);
        if (fabsf(sc) < 1e-20f) sc = sc < 0.f ? -1e-20f : 1e-20f;
        const float me = mean - beta[c] / sc;
        for (int n = g0; n < g0 + group; ++n) { eff[((long long)n * C + c) * 2] = me; eff[((long long)n * C + c) * 2 + 1] = sc; }
    }
    if (training) { running_mean[c] = rm; running_var[c] = rv; }
}

// Backward: the reduce pass leaves per image s1 = sum gd, s2 = sum gd * y (y = the affine output, gd = g * act'(y)).  With
// S1, S2 their sums over the group: dbeta = S1, dgamma = (S2 - beta S1) / gamma, and
// dx = rs_eff * (gd - S1/M - xhat * dgamma/M), xhat = (y - beta) / gamma  ==  rs_eff * (gd - A/M - y * B/M) with
// B = (S2 - beta S1) / gamma^2, A = S1 - beta B: exactly the apply pass's formula, so it only needs (A, B) in place of (s1, s2).
__global__ void bn_bwd_fix_kernel(float* __restrict__ bsum, int n_img, int group, int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                  float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
    irc::pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float ga = gamma[c];
    if (fabsf(ga) < 1e-20f) ga = ga < 0.f ? -1e-20f : 1e-20f;
    const float be = beta[c];
    float dg = 0.f, db = 0.f;
    for (int g0 = 0; g0 < n_img; g0 += group) {
        float S1 = 0.f, S2 = 0.f;
        for (int n = g0; n < g0 + group; ++n) { S1 += bsum[((long long)n * C + c) * 2]; S2 += bsum[((long long)n * C + c) * 2 + 1]; }
        const float dgam = (S2 - be * S1) / ga;
        const float B = dgam / ga, A = S1 - be * B;
        for (int n = g0; n < g0 + group; ++n) { bsum[((long long)n * C + c) * 2] = A; bsum[((long long)n * C + c) * 2 + 1] = B; }
        dg += dgam; db += S1;
    }
    if (accumulate) { dgamma[c] += dg; dbeta[c] += db; } else { dgamma[c] = dg; dbeta[c] = db; }
}

}  // namespace

extern "C" int irc_row_index(short* row_img, int n_img, int hp, int wp, int y0, int y1, int x0, int x1, void* stream) {
    if (!row_img) return irc_set_error(IRC_ERR_BAD_ARG, "irc_row_index: null");
    const long long total = (long long)n_img * hp * wp;
    irc::launch(row_index_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, row_img, n_img, hp, wp, y0, y1, x0, x1);
    return irc_check_launch("irc_row_index");
}

extern "C" int irc_in_stats(const irc_view* z, int C, int n_img, int H, int W, float* stats, float* work, long long work_floats, void* stream) {
    int rc = check_view(*z, "irc_in_stats"); if (rc) return rc;
    if (C % 8 || C > 2048) return irc_set_error(IRC_ERR_BAD_ARG, "irc_in_stats: C must be a multiple of 8");
    int threads, L, chunks; size_t smem;
    // the last kCounterFloats floats of `work` are the (zero-initialised, self re-arming) ticket counters
    const bool staged = work && work_floats > kCounterFloats && n_img <= kCounterFloats;
    reduce_shape(C, H, W, n_img, staged ? work_floats - kCounterFloats : 0, threads, L, chunks, smem);
    if (chunks == 1) {
        irc::launch(in_stats_kernel, dim3(1, n_img), threads, smem, (cudaStream_t)stream, mk(*z), C, H, W, stats, nullptr, nullptr);
        return irc_check_launch("irc_in_stats");
    }
    if (!z->s2d_c) {
        const size_t smem_p = smem > (size_t)kPipeD * threads * 16 ? smem : (size_t)kPipeD * threads * 16;
        irc::launch(in_stats_pipe_kernel, dim3(chunks, n_img), threads, smem_p, (cudaStream_t)stream, mk(*z), C, H, W, work, stats,
                    (unsigned*)(work + work_floats - kCounterFloats));
        return irc_check_launch("irc_in_stats");
    }
    irc::launch(in_stats_kernel, dim3(chunks, n_img), threads, smem, (cudaStream_t)stream, mk(*z), C, H, W, work, stats,
                                                                                 (unsigned*)(work + work_floats - kCounterFloats));
    return irc_check_launch("irc_in_stats");
}

extern "C" int irc_gather(const irc_gather_args* a, void* stream) {
    int rc = check_view(a->src, "irc_gather src"); if (rc) return rc;
    rc = check_view(a->dst, "irc_gather dst"); if (rc) return rc;
    if (a->C % 8) return irc_set_error(IRC_ERR_BAD_ARG, "irc_gather: C %% 8");
    GatherP p;
    p.src = mk(a->src); p.dst = mk(a->dst);
    p.has2 = a->src2.ptr != nullptr; if (p.has2) { rc = check_view(a->src2, "irc_gather src2"); if (rc) return rc; p.src2 = mk(a->src2); }
    p.has_res = a->res.ptr != nullptr; if (p.has_res) { rc = check_view(a->res, "irc_gather res"); if (rc) return rc; p.res = mk(a->res); }
    p.C = a->C; p.n_img = a->n_img;
    p.stats = a->stats; p.inv_cnt = a->cnt > 0 ? 1.f / a->cnt : 0.f; p.eps = a->eps; p.act = a->act; p.slope = a->slope;
    p.ty_idx = a->ty_idx; p.ty_w = a->ty_w; p.ky = a->ky > 0 ? a->ky : 1;
    p.tx_idx = a->tx_idx; p.tx_w = a->tx_w; p.kx = a->kx > 0 ? a->kx : 1;
    p.H = a->H; p.W = a->W; p.pad = a->pad; p.halo_mode = a->halo_mode; p.dst_s2d = a->dst_s2d;
    p.slope_eff = a->act == 1 ? 0.f : (a->act == 2 ? a->slope : 1.f);
    if (p.dst_s2d && (((p.H + 2 * p.pad) | (p.W + 2 * p.pad)) & 1)) return irc_set_error(IRC_ERR_BAD_ARG, "irc_gather: space-to-depth needs even padded extents");
    if (a->tile_y == -3) {
        // streaming separable kernel: tile_x = window of source rows (span of the y-table), patch_y = rows per strip
        const int K = a->tile_x, strip = a->patch_y;
        if (!p.ty_idx || !p.tx_idx || p.dst_s2d || p.src.s2d_c || (p.has2 && p.src2.s2d_c) || p.has_res || (p.act && !p.stats) ||
            K < p.ky || K < p.kx || K > 6 || strip < 1 || strip > kMaxStrip || (p.pad && (p.W <= 2 * p.pad + 1 || p.H <= 2 * p.pad + 1)) ||
            (p.has2 && p.stats))
            return irc_set_error(IRC_ERR_BAD_ARG, "irc_gather(stream): unsupported combination");
        int threads, L;
        row_block(p.C, p.W, threads, L);
        dim3 grid((p.W + L - 1) / L, (p.H + strip - 1) / strip, p.n_img);
        cudaStream_t st = (cudaStream_t)stream;
        const bool nrm = p.stats != nullptr;
        // staged-rows kernel (default; IRC_GATHER_STREAM=regs selects the register-only one below): patch_x = bandwidth of the
        // x-table (max over output columns x, x' with |x - x'| < L of the source-column span), which bounds the staged segment
        static int rows_mode = -1;
        if (rows_mode < 0) { const char* e = getenv("IRC_GATHER_STREAM"); rows_mode = (e && e[0] == 'r') ? 0 : 1; }
        const int sw_max = a->patch_x;
        // (normalise + up-sample stays on the register kernel: it is bound by the instructions of its output rows, not by load
        // latency, and the ring's extra shared-memory traffic costs 3 % there - profiles/r5/gathers_rows_vs_regs.txt)
        const bool up_norm = nrm && p.H > p.src.hp;
        if (rows_mode && !up_norm && sw_max > 0 && sw_max <= kRowsNS * L) {
            const size_t smem = (size_t)kRowsStg * (p.has2 ? 2 : 1) * sw_max * (p.C / 8) * 16;
            if (smem <= 48 * 1024) {
#define IRC_ROWS(KK) do { \
            if (p.has2) irc::launch(gather_rows_kernel<KK, 0, true>, grid, threads, smem, st, p, strip, sw_max); \
            else if (nrm && p.act == 1) irc::launch(gather_rows_kernel<KK, 1, false>, grid, threads, smem, st, p, strip, sw_max); \
            else if (nrm) irc::launch(gather_rows_kernel<KK, 2, false>, grid, threads, smem, st, p, strip, sw_max); \
            else irc::launch(gather_rows_kernel<KK, 0, false>, grid, threads, smem, st, p, strip, sw_max); } while (0)
                if (K <= 2) IRC_ROWS(2); else if (K <= 3) IRC_ROWS(3); else if (K <= 4) IRC_ROWS(4); else IRC_ROWS(6);
#undef IRC_ROWS
                return irc_check_launch("irc_gather(rows)");
            }
        }
#define IRC_STREAM(KK) do { \
            if (p.has2) irc::launch(gather_stream_kernel<KK, false, true>, grid, threads, 0, st, p, strip); \
            else if (nrm) irc::launch(gather_stream_kernel<KK, true, false>, grid, threads, 0, st, p, strip); \
            else irc::launch(gather_stream_kernel<KK, false, false>, grid, threads, 0, st, p, strip); } while (0)
        if (K <= 2) IRC_STREAM(2); else if (K <= 3) IRC_STREAM(3); else if (K <= 4) IRC_STREAM(4); else IRC_STREAM(6);
#undef IRC_STREAM
        return irc_check_launch("irc_gather(stream)");
    }
    const int kmax0 = p.ky > p.kx ? p.ky : p.kx;
    if (a->tile_y <= 0 && (p.ty_idx || p.tx_idx) && !p.dst_s2d && !p.src.s2d_c && !(p.has2 && p.src2.s2d_c) && (p.stats || !p.act) && kmax0 <= 6 &&
        a->tile_y != -2) {
        int threads, L;
        row_block(p.C, p.W + 2 * p.pad, threads, L);
        const long long rows = (long long)p.n_img * (p.H + 2 * p.pad);
        const unsigned grid = (unsigned)(rows < 65535 * 16 ? rows : 65535 * 16);
        cudaStream_t st = (cudaStream_t)stream;
        if (kmax0 <= 2) irc::launch(gather_lean_kernel<2>, grid, threads, 0, st, p);
        else if (kmax0 <= 3) irc::launch(gather_lean_kernel<3>, grid, threads, 0, st, p);
        else irc::launch(gather_lean_kernel<6>, grid, threads, 0, st, p);
        return irc_check_launch("irc_gather(lean)");
    }
    if (a->tile_y > 0 && a->tile_x > 0 && (p.ty_idx || p.tx_idx) && p.C % kTileCC == 0) {
        if (a->tile_y + a->tile_x > 64 || a->tile_y > 64 || a->tile_x > 64) return irc_set_error(IRC_ERR_BAD_ARG, "irc_gather: tile too large");
        const size_t smem = (size_t)a->patch_y * a->patch_x * kTileCC * sizeof(float);
        if (smem > 200 * 1024) return irc_set_error(IRC_ERR_BAD_ARG, "irc_gather: tile patch does not fit shared memory");
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(gather_tiled_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(gather_tiled_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(gather_tiled_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(gather_tiled_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            attr = true;
        }
        const int Hp = p.H + 2 * p.pad, Wp = p.W + 2 * p.pad;
        dim3 grid((Wp + a->tile_x - 1) / a->tile_x, (Hp + a->tile_y - 1) / a->tile_y, p.n_img * (p.C / kTileCC));
        const int kmax = p.ky > p.kx ? p.ky : p.kx;
        cudaStream_t st = (cudaStream_t)stream;
        if (kmax <= 2) irc::launch(gather_tiled_kernel<2>, grid, 256, smem, st, p, a->tile_y, a->tile_x, a->patch_y, a->patch_x);
        else if (kmax <= 3) irc::launch(gather_tiled_kernel<3>, grid, 256, smem, st, p, a->tile_y, a->tile_x, a->patch_y, a->patch_x);
        else if (kmax <= 6) irc::launch(gather_tiled_kernel<6>, grid, 256, smem, st, p, a->tile_y, a->tile_x, a->patch_y, a->patch_x);
        else if (kmax <= 8) irc::launch(gather_tiled_kernel<8>, grid, 256, smem, st, p, a->tile_y, a->tile_x, a->patch_y, a->patch_x);
        else return irc_set_error(IRC_ERR_BAD_ARG, "irc_gather: tables wider than 8 entries");
        return irc_check_launch("irc_gather(tiled)");
    }
    int threads, L;
    row_block(p.C, p.W + 2 * p.pad, threads, L);
    const long long rows = (long long)p.n_img * (p.H + 2 * p.pad);
    const unsigned grid = (unsigned)(rows < 65535 * 16 ? rows : 65535 * 16);
    if (!p.ty_idx && !p.tx_idx && !p.has2 && !p.dst_s2d && !p.src.s2d_c && !(p.has_res && p.res.s2d_c) && (p.stats || !p.act) && !getenv("IRC_NO_GATHER_PIPE")) {
        static int lean_mode = -1;      // IRC_GATHER_IDENT=pipe: the cp.async-slot kernel (kept for the A/B in profiles/)
        if (lean_mode < 0) { const char* e = getenv("IRC_GATHER_IDENT"); lean_mode = (e && e[0] == 'p') ? 0 : 1; }
        if (lean_mode) {
            const unsigned gl = (unsigned)(rows < (long long)irc_num_sms() * 8 ? rows : (long long)irc_num_sms() * 8);
            const int am = !p.stats ? 0 : (p.act == 1 ? 1 : 2);
            cudaStream_t st = (cudaStream_t)stream;
#define IRC_IDENT(RES) do { \
                if (am == 0) irc::launch(gather_ident_lean_kernel<RES, 0>, gl, threads, 0, st, p); \
                else if (am == 1) irc::launch(gather_ident_lean_kernel<RES, 1>, gl, threads, 0, st, p); \
                else irc::launch(gather_ident_lean_kernel<RES, 2>, gl, threads, 0, st, p); } while (0)
            if (p.has_res) IRC_IDENT(true); else IRC_IDENT(false);
#undef IRC_IDENT
            return irc_check_launch("irc_gather(ident)");
        }
        const size_t smem = (size_t)kPipeD * (p.has_res ? 2 : 1) * threads * 16;
        const unsigned gp = (unsigned)(rows < (long long)irc_num_sms() * 16 ? rows : (long long)irc_num_sms() * 16);
        if (p.has_res) irc::launch(gather_ident_pipe_kernel<true>, gp, threads, smem, (cudaStream_t)stream, p);
        else irc::launch(gather_ident_pipe_kernel<false>, gp, threads, smem, (cudaStream_t)stream, p);
        return irc_check_launch("irc_gather(pipe)");
    }
    if (!p.ty_idx && !p.tx_idx) irc::launch(gather_kernel<true>, grid, threads, 0, (cudaStream_t)stream, p);
    else irc::launch(gather_kernel<false>, grid, threads, 0, (cudaStream_t)stream, p);
    return irc_check_launch("irc_gather");
}

extern "C" int irc_in_apply_fused(const irc_gather_args* a, float* stats_out, void* stream) {
    int rc = check_view(a->src, "irc_in_apply_fused src"); if (rc) return rc;
    rc = check_view(a->dst, "irc_in_apply_fused dst"); if (rc) return rc;
    GatherP p;
    p.src = mk(a->src); p.dst = mk(a->dst);
    p.has2 = 0;
    p.has_res = a->res.ptr != nullptr; if (p.has_res) { rc = check_view(a->res, "irc_in_apply_fused res"); if (rc) return rc; p.res = mk(a->res); }
    p.C = a->C; p.n_img = a->n_img;
    p.stats = nullptr; p.inv_cnt = a->cnt > 0 ? 1.f / a->cnt : 0.f; p.eps = a->eps; p.act = a->act; p.slope = a->slope;
    p.ty_idx = nullptr; p.ty_w = nullptr; p.ky = 1; p.tx_idx = nullptr; p.tx_w = nullptr; p.kx = 1;
    p.H = a->H; p.W = a->W; p.pad = a->pad; p.halo_mode = a->halo_mode; p.dst_s2d = 0;
    p.slope_eff = a->act == 1 ? 0.f : (a->act == 2 ? a->slope : 1.f);
    const long long hw = (long long)p.H * p.W;
    if (!stats_out || a->src2.ptr || a->ty_idx || a->tx_idx || a->dst_s2d || p.src.s2d_c || p.dst.s2d_c || (p.has_res && p.res.s2d_c) || p.C % kFusedCC ||
        hw > 8 * kFusedNP * kFusedLanes || a->act < 0 || a->act > 2 || (p.pad && (p.W <= 2 * p.pad + 1 || p.H <= 2 * p.pad + 1)) || a->cnt != (float)hw)
        return irc_set_error(IRC_ERR_BAD_ARG, "irc_in_apply_fused: needs identity tables, one source, C %% 32 == 0, H*W <= 4096, cnt == H*W");
    unsigned cl = 1;
    while ((long long)cl * kFusedNP * kFusedLanes < hw) cl *= 2;
    const dim3 grid(cl, p.C / kFusedCC, p.n_img);
    // only z is staged: 32 KB per CTA, four CTAs per SM
    const size_t smem_max = kFusedNP * 256 * sizeof(uint4);
    const size_t smem = smem_max;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(in_apply_fused_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        cudaFuncSetAttribute(in_apply_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        cudaFuncSetAttribute(in_apply_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        attr = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (a->act == 1) irc::launch_cluster(in_apply_fused_kernel<1>, grid, 256, smem, st, cl, p, stats_out);
    else if (a->act == 2) irc::launch_cluster(in_apply_fused_kernel<2>, grid, 256, smem, st, cl, p, stats_out);
    else irc::launch_cluster(in_apply_fused_kernel<0>, grid, 256, smem, st, cl, p, stats_out);
    return irc_check_launch("irc_in_apply_fused");
}

static int fill_bwd(const irc_in_bwd_args* a, InBwdP& p) {
    int rc = check_view(a->z, "irc_in_bwd z"); if (rc) return rc;
    rc = check_view(a->g1, "irc_in_bwd g1"); if (rc) return rc;
    p.z = mk(a->z); p.g1 = mk(a->g1);
    p.has2 = a->g2.ptr != nullptr; if (p.has2) { rc = check_view(a->g2, "irc_in_bwd g2"); if (rc) return rc; p.g2 = mk(a->g2); }
    if (a->C % 8) return irc_set_error(IRC_ERR_BAD_ARG, "irc_in_bwd: C %% 8");
    p.C = a->C; p.n_img = a->n_img; p.H = a->H; p.W = a->W;
    p.stats = a->stats; p.inv_cnt = a->cnt > 0 ? 1.f / a->cnt : 0.f; p.eps = a->eps; p.act = a->act; p.slope = a->slope;
    p.ty_idx = a->ty_idx; p.ty_w = a->ty_w; p.ky = a->ky > 0 ? a->ky : 1;
    p.tx_idx = a->tx_idx; p.tx_w = a->tx_w; p.kx = a->kx > 0 ? a->kx : 1;
    p.bsum = a->bsum;
    return IRC_OK;
}

extern "C" int irc_in_bwd_reduce(const irc_in_bwd_args* a, void* stream) {
    InBwdP p; int rc = fill_bwd(a, p); if (rc) return rc;
    if (!p.stats || !p.bsum) return irc_set_error(IRC_ERR_BAD_ARG, "irc_in_bwd_reduce: stats and bsum required");
    int threads, L, chunks; size_t smem;
    const bool staged = a->work && a->work_floats > kCounterFloats && p.n_img <= kCounterFloats;
    reduce_shape(p.C, p.H, p.W, p.n_img, staged ? a->work_floats - kCounterFloats : 0, threads, L, chunks, smem);
    p.part = chunks == 1 ? p.bsum : a->work;
    p.counters = chunks == 1 ? nullptr : (unsigned*)(a->work + a->work_floats - kCounterFloats);
    const bool plain = !p.ty_idx && !p.tx_idx && !p.z.s2d_c && !p.g1.s2d_c && !(p.has2 && p.g2.s2d_c);
    static bool attr = false;
    if (!attr) {       // 48 KB of slots + the static ticket flag exceed the default limit
        cudaFuncSetAttribute(in_bwd_reduce_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024);
        cudaFuncSetAttribute(in_bwd_reduce_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024);
        attr = true;
    }
    if (plain) {
        const size_t need = (size_t)kPipeD * (p.has2 ? 3 : 2) * threads * 16;
        const size_t smem_p = smem > need ? smem : need;
        if (p.has2) irc::launch(in_bwd_reduce_pipe_kernel<true>, dim3(chunks, p.n_img), threads, smem_p, (cudaStream_t)stream, p);
        else irc::launch(in_bwd_reduce_pipe_kernel<false>, dim3(chunks, p.n_img), threads, smem_p, (cudaStream_t)stream, p);
    } else if (!p.ty_idx && !p.tx_idx) irc::launch(in_bwd_reduce_kernel<true>, dim3(chunks, p.n_img), threads, smem, (cudaStream_t)stream, p);
    else irc::launch(in_bwd_reduce_kernel<false>, dim3(chunks, p.n_img), threads, smem, (cudaStream_t)stream, p);
    return irc_check_launch("irc_in_bwd_reduce");
}

extern "C" int irc_in_bwd_apply(const irc_in_bwd_args* a, void* stream) {
    InBwdP p; int rc = fill_bwd(a, p); if (rc) return rc;
    rc = check_view(a->dz, "irc_in_bwd dz"); if (rc) return rc;
    p.dz = mk(a->dz);
    if (p.stats && !p.bsum) return irc_set_error(IRC_ERR_BAD_ARG, "irc_in_bwd_apply: bsum required with stats");
    int threads, L;
    row_block(p.C, p.W, threads, L);
    const long long rows = (long long)p.n_img * p.H;
    const unsigned grid = (unsigned)(rows < 65535 * 16 ? rows : 65535 * 16);
    const bool plain = !p.ty_idx && !p.tx_idx && !p.z.s2d_c && !p.g1.s2d_c && !(p.has2 && p.g2.s2d_c) && !p.dz.s2d_c;
    if (plain) {
        const size_t smem_p = (size_t)kPipeD * (p.has2 ? 3 : 2) * threads * 16;
        // a few rows per block so that the prefetch pipeline runs across row boundaries
        const long long gcap = (long long)irc_num_sms() * 16;          // (4..32 blocks per SM measured: no difference)
        const unsigned gp = (unsigned)(rows < gcap ? rows : gcap);
        if (p.has2) irc::launch(in_bwd_apply_pipe_kernel<true>, gp, threads, smem_p, (cudaStream_t)stream, p);
        else irc::launch(in_bwd_apply_pipe_kernel<false>, gp, threads, smem_p, (cudaStream_t)stream, p);
    } else if (!p.ty_idx && !p.tx_idx) irc::launch(in_bwd_apply_kernel<true>, grid, threads, 0, (cudaStream_t)stream, p);
    else irc::launch(in_bwd_apply_kernel<false>, grid, threads, 0, (cudaStream_t)stream, p);
    return irc_check_launch("irc_in_bwd_apply");
}

extern "C" int irc_in_bwd_l2(const irc_in_bwd_args* a, int groups, void* stream) {
    InBwdP p; int rc = fill_bwd(a, p); if (rc) return rc;
    rc = check_view(a->dz, "irc_in_bwd_l2 dz"); if (rc) return rc;
    p.dz = mk(a->dz);
    const int C8 = p.C / 8;
    if (!p.stats || p.ty_idx || p.tx_idx || p.z.s2d_c || p.g1.s2d_c || p.dz.s2d_c || (p.has2 && p.g2.s2d_c) || p.C > 256 || 256 % C8 || p.W < 256 / C8 ||
        !a->work || groups < 1)
        return irc_set_error(IRC_ERR_BAD_ARG, "irc_in_bwd_l2: needs stats, identity tables, plain views, C in {64,128,256}, W >= 2048/C, a workspace, groups >= 1");
    // prefetch depth: 4 pixels per thread in flight (8 measured no faster: IRC_INBWD_L2_DEPTH=8)
    static int depth = 0;
    if (!depth) { const char* e = getenv("IRC_INBWD_L2_DEPTH"); depth = e ? atoi(e) : 4; if (depth != 8) depth = 4; }
    const size_t smem = (size_t)depth * (p.has2 ? 3 : 2) * 256 * 16;
    static int per_sm[2] = {0, 0};
    if (!per_sm[p.has2]) {
        cudaFuncSetAttribute(in_bwd_l2_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        cudaFuncSetAttribute(in_bwd_l2_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        cudaFuncSetAttribute(in_bwd_l2_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        cudaFuncSetAttribute(in_bwd_l2_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        int n0 = 0;
        cudaError_t e;
        if (depth == 8) e = p.has2 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n0, in_bwd_l2_kernel<true, 8>, 256, smem)
                                   : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n0, in_bwd_l2_kernel<false, 8>, 256, smem);
        else e = p.has2 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n0, in_bwd_l2_kernel<true, 4>, 256, smem)
                        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n0, in_bwd_l2_kernel<false, 4>, 256, smem);
        if (e != cudaSuccess || n0 < 1) return irc_check_launch("cudaOccupancyMaxActiveBlocksPerMultiprocessor(in_bwd_l2)");
        per_sm[p.has2] = n0 > 2 ? 2 : n0;
    }
    // K image groups, K a divisor of the image count (every group then walks the same number of images)
    int K = groups < p.n_img ? groups : p.n_img;
    if (K > 32) K = 32;
    while (p.n_img % K) --K;
    int Gc = irc_num_sms() * per_sm[p.has2] / K;           // the whole grid must be co-resident (spin barrier)
    if (Gc > p.H) Gc = p.H;
    const long long need = 2LL * K * Gc * p.C * 2;
    if (Gc < 1 || a->work_floats < need + kCounterFloats) return irc_set_error(IRC_ERR_BAD_ARG, "irc_in_bwd_l2: workspace too small (%lld floats needed)", need + kCounterFloats);
    unsigned* sync = (unsigned*)(a->work + a->work_floats - kCounterFloats) + 256;
    cudaStream_t st = (cudaStream_t)stream;
    if (depth == 8) {
        if (p.has2) irc::launch(in_bwd_l2_kernel<true, 8>, K * Gc, 256, smem, st, p, K, Gc, a->work, sync);
        else irc::launch(in_bwd_l2_kernel<false, 8>, K * Gc, 256, smem, st, p, K, Gc, a->work, sync);
    } else {
        if (p.has2) irc::launch(in_bwd_l2_kernel<true, 4>, K * Gc, 256, smem, st, p, K, Gc, a->work, sync);
        else irc::launch(in_bwd_l2_kernel<false, 4>, K * Gc, 256, smem, st, p, K, Gc, a->work, sync);
    }
    return irc_check_launch("irc_in_bwd_l2");
}

extern "C" int irc_in_bwd_fused(const irc_in_bwd_args* a, int fold_pad, void* stream) {
    InBwdP p; int rc = fill_bwd(a, p); if (rc) return rc;
    rc = check_view(a->dz, "irc_in_bwd_fused dz"); if (rc) return rc;
    p.dz = mk(a->dz);
    const long long hw = (long long)p.H * p.W;
    if (!p.stats || p.has2 || p.ty_idx || p.tx_idx || p.g1.s2d_c || p.z.s2d_c || p.dz.s2d_c || p.C % kFusedCC || hw > 8 * kFusedNP * kFusedLanes || fold_pad < 0 ||
        (fold_pad && (2 * fold_pad + 2 > p.H || 2 * fold_pad + 2 > p.W)))
        return irc_set_error(IRC_ERR_BAD_ARG, "irc_in_bwd_fused: needs stats, one source, identity tables, C %% 32 == 0 and H*W <= 4096");
    unsigned cl = 1;
    while ((long long)cl * kFusedNP * kFusedLanes < hw) cl *= 2;
    const dim3 grid(cl, p.C / kFusedCC, p.n_img);
    const size_t smem = 2 * kFusedNP * 256 * sizeof(uint4);       // 64 KB
    static bool attr = false;
    if (!attr) {
#define IRC_ATTR(F, A) cudaFuncSetAttribute(in_bwd_fused_kernel<F, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
        IRC_ATTR(true, 0); IRC_ATTR(true, 1); IRC_ATTR(true, 2); IRC_ATTR(false, 0); IRC_ATTR(false, 1); IRC_ATTR(false, 2);
#undef IRC_ATTR
        attr = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (p.act < 0 || p.act > 2) return irc_set_error(IRC_ERR_BAD_ARG, "irc_in_bwd_fused: act must be 0, 1 or 2");
#define IRC_GO(F, A) irc::launch_cluster(in_bwd_fused_kernel<F, A>, grid, 256, smem, st, cl, p, fold_pad)
    if (fold_pad) { if (p.act == 1) IRC_GO(true, 1); else if (p.act == 2) IRC_GO(true, 2); else IRC_GO(true, 0); }
    else { if (p.act == 1) IRC_GO(false, 1); else if (p.act == 2) IRC_GO(false, 2); else IRC_GO(false, 0); }
#undef IRC_GO
    return irc_check_launch("irc_in_bwd_fused");
}

extern "C" int irc_maxpool2(const irc_view* src, const irc_view* dst, int C, int n_img, int Ho, int Wo, void* stream) {
    int rc = check_view(*src, "irc_maxpool2 src"); if (rc) return rc;
    rc = check_view(*dst, "irc_maxpool2 dst"); if (rc) return rc;
    if (C % 8 || src->s2d_c || dst->s2d_c) return irc_set_error(IRC_ERR_BAD_ARG, "irc_maxpool2: C %% 8 == 0, plain (not space-to-depth) views");
    int threads, L;
    row_block(C, Wo, threads, L);
    const long long rows = (long long)n_img * Ho;
    const unsigned grid = (unsigned)(rows < (long long)irc_num_sms() * 8 ? rows : (long long)irc_num_sms() * 8);
    irc::launch(maxpool_kernel, grid, threads, 0, (cudaStream_t)stream, mk(*src), mk(*dst), C, n_img, Ho, Wo);
    return irc_check_launch("irc_maxpool2");
}

extern "C" int irc_maxpool2_bwd(const irc_view* src, const irc_view* g, const irc_view* dsrc, int C, int n_img, int Ho, int Wo, void* stream) {
    int rc = check_view(*src, "irc_maxpool2_bwd src"); if (rc) return rc;
    rc = check_view(*g, "irc_maxpool2_bwd g"); if (rc) return rc;
    rc = check_view(*dsrc, "irc_maxpool2_bwd dsrc"); if (rc) return rc;
    if (C % 8 || src->s2d_c || g->s2d_c || dsrc->s2d_c) return irc_set_error(IRC_ERR_BAD_ARG, "irc_maxpool2_bwd: C %% 8 == 0, plain (not space-to-depth) views");
    int threads, L;
    row_block(C, Wo, threads, L);
    const long long rows = (long long)n_img * Ho;
    const unsigned grid = (unsigned)(rows < (long long)irc_num_sms() * 8 ? rows : (long long)irc_num_sms() * 8);
    irc::launch(maxpool_bwd_kernel, grid, threads, 0, (cudaStream_t)stream, mk(*src), mk(*g), mk(*dsrc), C, n_img, Ho, Wo);
    return irc_check_launch("irc_maxpool2_bwd");
}

extern "C" int irc_colsum(const void* a, long long rows, long long ld, int chan_off, int C, const short* row_img, float* out, float* work,
                          long long work_floats, void* stream) {
    if (!a || !out) return irc_set_error(IRC_ERR_BAD_ARG, "irc_colsum: null");
    long long bx = (rows + 31) / 32; if (bx > irc_num_sms() * 4) bx = irc_num_sms() * 4;
    const long long cap = work && work_floats > kCounterFloats ? (work_floats - kCounterFloats) / C : 0;
    if (bx > cap) bx = cap;
    if (bx <= 1) {
        irc::launch(colsum_kernel, dim3(1, (C + 31) / 32), 1024, 0, (cudaStream_t)stream, (const bf16*)a, rows, ld, chan_off, C, row_img, out);
        return irc_check_launch("irc_colsum");
    }
    if (C % 8 == 0 && chan_off % 8 == 0 && ld % 8 == 0 && !((uintptr_t)a & 15) && C <= 2048 && 256 % (C / 8) == 0) {
        const int L = 256 / (C / 8);
        long long bv = (rows + 4 * L - 1) / (4 * L); if (bv > irc_num_sms() * 4) bv = irc_num_sms() * 4;
        if (bv > cap) bv = cap;
        if (bv < 1) bv = 1;
        irc::launch(colsum_vec_kernel, (unsigned)bv, 256, (size_t)L * C * sizeof(float), (cudaStream_t)stream, (const bf16*)a, rows, ld, chan_off, C, row_img, work);
        irc::launch(sum_chunks_kernel, grid_for((long long)C * 32, 256), 256, 0, (cudaStream_t)stream, work, (int)bv, C, out);
        return irc_check_launch("irc_colsum(vec)");
    }
    irc::launch(colsum_kernel, dim3((unsigned)bx, (C + 31) / 32), 1024, 0, (cudaStream_t)stream, (const bf16*)a, rows, ld, chan_off, C, row_img, work);
    irc::launch(sum_chunks_kernel, grid_for((long long)C * 32, 256), 256, 0, (cudaStream_t)stream, work, (int)bx, C, out);
    return irc_check_launch("irc_colsum");
}

extern "C" int irc_fold_inplace(void* g, long long ld, int chan_off, int C, int n_img, int H, int W, int p, void* stream) {
    if (!g || C % 8 || ((uintptr_t)g & 15) || ld % 8 || chan_off % 8) return irc_set_error(IRC_ERR_BAD_ARG, "irc_fold_inplace: bad args");
    if (p < 1 || 2 * p + 2 > H || 2 * p + 2 > W) return irc_set_error(IRC_ERR_BAD_ARG, "irc_fold_inplace: image too small for the pad width");
    int threads, L;
    row_block(C, W, threads, L);
    if (n_img > 65535) return irc_set_error(IRC_ERR_BAD_ARG, "irc_fold_inplace: extent too large");
    const long long items = 2LL * p * W + (long long)(H - 2 * p) * 2 * p;
    irc::launch(fold_inplace_kernel, dim3((unsigned)((items + L - 1) / L), n_img), threads, 0, (cudaStream_t)stream, (bf16*)g, ld, chan_off, C, n_img, H, W, p);
    return irc_check_launch("irc_fold_inplace");
}

extern "C" int irc_bn_finalize(const float* stats, int n_img, int group, int C, float cnt_per_img, const float* gamma, const float* beta, float* running_mean,
                               float* running_var, float momentum, float eps, int training, int updates, float* eff, void* stream) {
    if (!gamma || !beta || !running_mean || !running_var || !eff || (training && !stats) || n_img <= 0 || group <= 0 || n_img % group || C <= 0)
        return irc_set_error(IRC_ERR_BAD_ARG, "irc_bn_finalize: bad args (n_img must be a multiple of the statistics group)");
    irc::launch(bn_finalize_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream, stats, n_img, group, C, cnt_per_img, gamma, beta, running_mean, running_var,
                momentum, eps, training, updates < 1 ? 1 : updates, eff);
    return irc_check_launch("irc_bn_finalize");
}

extern "C" int irc_bn_bwd_fix(float* bsum, int n_img, int group, int C, const float* gamma, const float* beta, float* dgamma, float* dbeta, int accumulate,
                              void* stream) {
    if (!bsum || !gamma || !beta || !dgamma || !dbeta || n_img <= 0 || group <= 0 || n_img % group || C <= 0)
        return irc_set_error(IRC_ERR_BAD_ARG, "irc_bn_bwd_fix: bad args");
    irc::launch(bn_bwd_fix_kernel, (C + 127) / 128, 128, 0, (cudaStream_t)stream, bsum, n_img, group, C, gamma, beta, dgamma, dbeta, accumulate);
    return irc_check_launch("irc_bn_bwd_fix");
}
