// Layout passes around the degenerate convolutions: the layers whose GEMM
// has a tiny K (inc 1->64 7x7, VGG conv1_1 3->64, D model.0 4->64) are fed by an im2col
// operand of 64 columns; the layers with a tiny N (outc 64->3 7x7, D model.11 512->1) are
// computed as a GEMM over one kernel axis followed by a shifted tap reduction.
#include <math.h>

#include "irc_common.cuh"
#include "../../include/irc_b200.h"

using namespace irc;

namespace {

__device__ __forceinline__ int reflect_idx(int i, int n) {
    if (i < 0) i = -i;
    if (i > n - 1) i = 2 * (n - 1) - i;
    return i;
}

struct RowMap {
    int mode, n_img, Ho, Wo;
    // decode flat row -> (n, oy, ox); returns false for padding rows
    __device__ __forceinline__ bool decode(long long q, int& n, int& oy, int& ox) const {
        if (mode == 0) {
            ox = (int)(q % Wo); q /= Wo; oy = (int)(q % Ho); n = (int)(q / Ho);
            return true;
        } else if (mode == 1) {
            const int wp = Wo + 2, hp = Ho + 2;
            ox = (int)(q % wp) - 1; q /= wp; oy = (int)(q % hp) - 1; n = (int)(q / hp);
            return oy >= 0 && oy < Ho && ox >= 0 && ox < Wo;
        } else {
            const int wb = (Wo + 2) >> 1, hb = (Ho + 2) >> 1;
            const int sub = (int)(q & 3); q >>= 2;
            const int xb = (int)(q % wb); q /= wb; const int yb = (int)(q % hb); n = (int)(q / hb);
            oy = 2 * yb + (sub >> 1) - 1; ox = 2 * xb + (sub & 1) - 1;
            return oy >= 0 && oy < Ho && ox >= 0 && ox < Wo;
        }
    }
    __device__ __forceinline__ long long encode(int n, int oy, int ox) const {
        if (mode == 0) return ((long long)n * Ho + oy) * Wo + ox;
        if (mode == 1) return ((long long)n * (Ho + 2) + oy + 1) * (Wo + 2) + ox + 1;
        const int wb = (Wo + 2) >> 1, hb = (Ho + 2) >> 1;
        const int yp = oy + 1, xp = ox + 1;
        return ((((long long)n * hb + (yp >> 1)) * wb + (xp >> 1)) << 2) + ((yp & 1) * 2 + (xp & 1));
    }
    __host__ __device__ long long rows() const {
        if (mode == 0) return (long long)n_img * Ho * Wo;
        return (long long)n_img * (Ho + 2) * (Wo + 2);
    }
    // the flat rows of one image form `lines()` lines of `line_len()` consecutive rows
    __host__ __device__ int lines() const { return mode == 0 ? Ho : (mode == 1 ? Ho + 2 : (Ho + 2) >> 1); }
    __host__ __device__ int line_len() const { return mode == 0 ? Wo : (mode == 1 ? Wo + 2 : ((Wo + 2) >> 1) * 4); }
    __device__ __forceinline__ bool decode_line(int line, int i, int& oy, int& ox) const {
        if (mode == 0) { oy = line; ox = i; return true; }
        if (mode == 1) { oy = line - 1; ox = i - 1; }
        else { const int sub = i & 3; oy = 2 * line + (sub >> 1) - 1; ox = 2 * (i >> 2) + (sub & 1) - 1; }
        return oy >= 0 && oy < Ho && ox >= 0 && ox < Wo;
    }
};

struct Im2colP {
    const float* src1; const float* src2; int c1, c2;
    const float* scale; const float* shift;
    int H, W, k, stride, pad, pad_mode;
    RowMap rm;
    bf16* dst; short* row_img;
};

__global__ void __launch_bounds__(256) im2col_kernel(const Im2colP p, int R, int Wp, int LPB) {
    // One block per (image, group of LPB consecutive lines of the row order).  The R input rows the group needs are staged
    // once in shared memory - all channels, padding / reflection and the per-channel affine resolved while filling, with
    // coalesced loads - so a thread (one item of a line x one group of 8 columns) assembles its 16 bytes with 8
    // shared-memory reads at register-resident offsets: no border cases, no scattered global loads.
    irc::pdl_prologue();
    extern __shared__ float tile[];                 // [C][R][Wp]
    const int C = p.c1 + p.c2;
    const int K = p.k * p.k * C;
    const int g = threadIdx.x & 7;
    int off[8]; unsigned valid = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = g * 8 + j;
        off[j] = 0;
        if (col < K) {
            const int c = col % C, rs = col / C, r = rs / p.k, s_ = rs % p.k;
            valid |= 1u << j;
            off[j] = (c * R + r) * Wp + s_;
        }
    }
    const int nl = p.rm.lines(), ll = p.rm.line_len();
    const int ngrp = (nl + LPB - 1) / LPB;
    const long long hw = (long long)p.H * p.W;
    for (int bg = blockIdx.x; bg < p.rm.n_img * ngrp; bg += gridDim.x) {
        const int n = bg / ngrp, line0 = (bg - n * ngrp) * LPB;
        const int nlines = min(LPB, nl - line0);
        // first output row of the group and the input row it starts at
        const int oy0 = p.rm.mode == 0 ? line0 : (p.rm.mode == 1 ? line0 - 1 : 2 * line0 - 1);
        const int y_lo = oy0 * p.stride - p.pad;
        __syncthreads();                            // the previous group's readers are done with the tile
        // one warp per staged row (channel c, input row r): the row decode happens once, lanes run along x
        for (int t = threadIdx.x >> 5; t < C * R; t += blockDim.x >> 5) {
            const int c = t / R, r = t - c * R;
            int y = y_lo + r;
            if (p.pad_mode == 1) y = reflect_idx(y, p.H);
            const bool oky = y >= 0 && y < p.H;
            const float* srow = (c < p.c1 ? p.src1 + ((long long)n * p.c1 + c) * hw : p.src2 + ((long long)n * p.c2 + (c - p.c1)) * hw) + (long long)(oky ? y : 0) * p.W;
            const float sc = p.scale ? __ldg(p.scale + c) : 1.f, sh = p.scale ? __ldg(p.shift + c) : 0.f;
            float* trow = tile + t * Wp;
            for (int xx = threadIdx.x & 31; xx < Wp; xx += 32) {
                int x = xx - p.pad;
                if (p.pad_mode == 1) x = reflect_idx(x, p.W);
                float v = 0.f;
                if (oky && x >= 0 && x < p.W) v = fmaf(__ldg(srow + x), sc, sh);
                trow[xx] = v;
            }
        }
        __syncthreads();
        for (int l = 0; l < nlines; ++l) {
            const int line = line0 + l;
            const long long qbase = ((long long)n * nl + line) * ll;
            for (int i = threadIdx.x >> 3; i < ll; i += blockDim.x >> 3) {
                const long long q = qbase + i;
                int oy, ox;
                const bool live = p.rm.decode_line(line, i, oy, ox);
                if (g == 0 && p.row_img) p.row_img[q] = live ? (short)n : (short)-1;
                float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                if (live) {
                    const int base = (oy - oy0) * p.stride * Wp + ox * p.stride;
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = (valid >> j) & 1 ? tile[base + off[j]] : 0.f;
                }
                *reinterpret_cast<uint4*>(p.dst + q * 64 + g * 8) =
                    make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// Direct small-K convolution (inc 1->64 7x7 irc:458-463, VGG conv1_1 3->64 3x3 irc:664, D model.0 4->64 4x4 s2 irc:600)
// ------------------------------------------------------------------------------------
// The im2col operand of these layers (64 bf16 slots per output pixel) is 16..30x larger than the image it is built from;
// writing it to HBM and reading it back in a one-tap GEMM makes the layer a pair of layout passes (VGG conv1_1 at 2 x 16
// images of 256 x 256: 272 MB written, 272 MB read, 272 MB written: 0.32 ms).  Here the block stages the input rows of a group
// of output lines in shared memory exactly as im2col_kernel does, and every warp multiplies 16 output pixels x K slots,
// gathered from that tile straight into mma.sync A fragments, with the [64][K] weight matrix held in registers as B fragments;
// bias / activation / ring zeroing happen on the accumulators and the [16][64] bf16 tile leaves through a per-warp
// shared-memory transpose as whole 128-byte rows.  HBM traffic = the image once + the output once.  The operand itself is
// only written (same pass, same fragments) when the caller wants it for the weight gradient (dst != NULL).
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all_() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit_() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kSkPitch = 36;      // u32 per staged output row (32 + 4): conflict-free fragment stores and 16-byte row reads

// activation on a packed bf16 pair AFTER the fp32 -> bf16 rounding (ReLU and LeakyReLU with 0 <= slope < 1 commute with it up to the
// rounding of slope * v, far below the bf16 step of the stored value): one or two packed instructions instead of six scalar ones
__device__ __forceinline__ uint32_t act_bf16x2(uint32_t v, int act, __nv_bfloat162 slope2) {
    __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&v);
    if (act == 1) x = __hmax2(x, __float2bfloat162_rn(0.f));
    else if (act == 2) x = __hmax2(x, __hmul2(x, slope2));
    return *reinterpret_cast<uint32_t*>(&x);
}

template <int KS, int MODE, bool KEEP, bool EPI>      // K = 16 * KS slots; row order of RowMap; KEEP: also write the operand E; EPI: bias / activation
__global__ void __launch_bounds__(256, 2) smallk_conv_fwd_kernel(const Im2colP p, const bf16* __restrict__ w, const float* __restrict__ bias, int act, float slope,
                                                                 bf16* __restrict__ out, int R, int Wp, int LPB, int nbuf) {
    irc::pdl_prologue();
    extern __shared__ float tile_base[];            // NBUF x [C][R][Wp] fp32 tiles, then 8 warps x (1 + KEEP) x [16][kSkPitch] u32 staging
    const int C = p.c1 + p.c2;
    const int K = p.k * p.k * C;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int tile_floats = C * R * Wp;
    uint32_t* stage_o = reinterpret_cast<uint32_t*>(tile_base + (size_t)nbuf * tile_floats) + warp * ((KEEP ? 2 : 1) * 16 * kSkPitch);
    uint32_t* stage_e = stage_o + 16 * kSkPitch;
    // operand columns this thread gathers: per k-step the pairs (2t, 2t+1) and (2t+8, 2t+9); columns >= K meet zero weights
    // (their tile reads stay inside the tile: offset 0), and are masked out of the operand copy
    int off[KS][4]; uint32_t cmask[KS][2];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = ks * 16 + 2 * t + (j & 1) + (j >> 1) * 8;
            off[ks][j] = 0;
            if ((j & 1) == 0) cmask[ks][j >> 1] = 0u;
            if (col < K) {
                const int c = col % C, rs = col / C, r = rs / p.k, s_ = rs % p.k;
                off[ks][j] = (c * R + r) * Wp + s_;
                cmask[ks][j >> 1] |= (j & 1) ? 0xffff0000u : 0x0000ffffu;
            }
        }
    // B fragments: B[k][n] = w[n][k]; (b0, b1) of n-tile j, k-step ks = w[8j + g][16ks + 2t .. +1], [.. + 8 .. +9]
    uint32_t bf[KS][8][2];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t* wr = reinterpret_cast<const uint32_t*>(w + (size_t)(8 * j + g) * 64 + ks * 16 + 2 * t);
            bf[ks][j][0] = __ldg(wr); bf[ks][j][1] = __ldg(wr + 4);
        }
    const __nv_bfloat162 slope2 = __float2bfloat162_rn(slope);
    const int Ho = p.rm.Ho, Wo = p.rm.Wo;
    const int nl = p.rm.lines(), ll = p.rm.line_len();
    const int ngrp = (nl + LPB - 1) / LPB;
    const long long hw = (long long)p.H * p.W;
    const int sWp = p.stride * Wp;
    // tile fill with 4-byte cp.async (no load -> store dependence: every copy of a tile is in flight at once); interior cells in a
    // branch-free loop, the 2 * pad border columns (zeros or reflected pixels) by the first lanes.  With nbuf == 2 the NEXT group's
    // tile is fetched while the current one is multiplied, so the global-load latency never sits between two barriers.
    auto fill = [&](float* tile, int bg) {
        const int n = bg / ngrp, line0 = (bg - n * ngrp) * LPB;
        const int oy0 = MODE == 0 ? line0 : (MODE == 1 ? line0 - 1 : 2 * line0 - 1);
        const int y_lo = oy0 * p.stride - p.pad;
        for (int tt = warp; tt < C * R; tt += 8) {
            const int c = tt / R, r = tt - c * R;
            int y = y_lo + r;
            if (p.pad_mode == 1) y = reflect_idx(y, p.H);
            float* trow = tile + tt * Wp;
            if (y < 0 || y >= p.H) {
                for (int xx = lane; xx < Wp; xx += 32) trow[xx] = 0.f;
                continue;
            }
            const float* srow = (c < p.c1 ? p.src1 + ((long long)n * p.c1 + c) * hw : p.src2 + ((long long)n * p.c2 + (c - p.c1)) * hw) + (long long)y * p.W;
            for (int x = lane; x < p.W; x += 32) cp_async4(trow + p.pad + x, srow + x);
            if (lane < 2 * p.pad) {
                const int xx = lane < p.pad ? lane : p.W + lane;                 // left / right border column of the padded row
                if (p.pad_mode == 1) cp_async4(trow + xx, srow + reflect_idx(xx - p.pad, p.W));
                else trow[xx] = 0.f;
            }
        }
        cp_async_commit_();
    };
    auto affine = [&](float* tile, int bg) {
        // per-channel affine (VGG input normalisation) on the cells this thread copied itself; zero padding stays zero
        const int n = bg / ngrp, line0 = (bg - n * ngrp) * LPB;
        const int oy0 = MODE == 0 ? line0 : (MODE == 1 ? line0 - 1 : 2 * line0 - 1);
        const int y_lo = oy0 * p.stride - p.pad;
        for (int tt = warp; tt < C * R; tt += 8) {
            const int c = tt / R, r = tt - c * R;
            const int y = y_lo + r;
            if (p.pad_mode != 1 && (y < 0 || y >= p.H)) continue;
            const float sc = __ldg(p.scale + c), sh = __ldg(p.shift + c);
            float* trow = tile + tt * Wp;
            for (int x = lane; x < p.W; x += 32) trow[p.pad + x] = fmaf(trow[p.pad + x], sc, sh);
            if (p.pad_mode == 1 && lane < 2 * p.pad) { const int xx = lane < p.pad ? lane : p.W + lane; trow[xx] = fmaf(trow[xx], sc, sh); }
        }
    };
    const int total = p.rm.n_img * ngrp;
    int cur = 0;
    if ((int)blockIdx.x < total) fill(tile_base, blockIdx.x);
    for (int bg = blockIdx.x; bg < total; bg += gridDim.x) {
        const int n = bg / ngrp, line0 = (bg - n * ngrp) * LPB;
        const int nlines = min(LPB, nl - line0);
        const int oy0 = MODE == 0 ? line0 : (MODE == 1 ? line0 - 1 : 2 * line0 - 1);
        float* tile = tile_base + (size_t)cur * tile_floats;
        const int nxt = bg + gridDim.x;
        if (nbuf == 2) {
            if (nxt < total) { fill(tile_base + (size_t)(cur ^ 1) * tile_floats, nxt); cp_async_wait_<1>(); } else cp_async_wait_<0>();
        } else {
            cp_async_wait_<0>();
        }
        if (p.scale) affine(tile, bg);
        __syncthreads();
        // the group's rows are one contiguous range of the flat row order: chunks of 16 rows go round-robin over the warps;
        // (l, ii) = (line within the group, position in the line) of this thread's first row, advanced without divisions
        const int nrows = nlines * ll;
        const long long qbase = ((long long)n * nl + line0) * ll;
        int l = 0, ii = warp * 16 + g;
        while (ii >= ll) { ii -= ll; ++l; }
        for (int i0 = warp * 16; i0 < nrows; i0 += 8 * 16) {
            int base[2]; bool live[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int lh = l, ih = ii + 8 * h;
                if (ih >= ll) { ih -= ll; ++lh; }
                int ry, ox;                          // output row relative to oy0, output column
                if (MODE == 0) { ry = lh; ox = ih; }
                else if (MODE == 1) { ry = lh; ox = ih - 1; }
                else { const int sub = ih & 3; ry = 2 * lh + (sub >> 1); ox = 2 * (ih >> 2) + (sub & 1) - 1; }
                const int oy = oy0 + ry;
                live[h] = (i0 + g + 8 * h < nrows) && (MODE == 0 || (oy >= 0 && oy < Ho && ox >= 0 && ox < Wo));
                base[h] = live[h] ? ry * sWp + ox * p.stride : 0;
                if (t == 0 && p.row_img && i0 + g + 8 * h < nrows) p.row_img[qbase + i0 + g + 8 * h] = live[h] ? (short)n : (short)-1;
            }
            float acc[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const float* t0 = tile + base[0];
                const float* t1 = tile + base[1];
                const uint32_t a[4] = {pack_bf16x2(t0[off[ks][0]], t0[off[ks][1]]), pack_bf16x2(t1[off[ks][0]], t1[off[ks][1]]),
                                       pack_bf16x2(t0[off[ks][2]], t0[off[ks][3]]), pack_bf16x2(t1[off[ks][2]], t1[off[ks][3]])};
#pragma unroll
                for (int j = 0; j < 8; ++j) mma_bf16_16816(acc[j], a, bf[ks][j][0], bf[ks][j][1]);
                if (KEEP) {
                    const uint32_t m0 = live[0] ? 0xffffffffu : 0u, m1 = live[1] ? 0xffffffffu : 0u;
                    stage_e[g * kSkPitch + ks * 8 + t] = a[0] & cmask[ks][0] & m0; stage_e[(g + 8) * kSkPitch + ks * 8 + t] = a[1] & cmask[ks][0] & m1;
                    stage_e[g * kSkPitch + ks * 8 + 4 + t] = a[2] & cmask[ks][1] & m0; stage_e[(g + 8) * kSkPitch + ks * 8 + 4 + t] = a[3] & cmask[ks][1] & m1;
                }
            }
            if (KEEP && KS < 4) {
#pragma unroll
                for (int ks = KS; ks < 4; ++ks) {
                    stage_e[g * kSkPitch + ks * 8 + t] = 0u; stage_e[(g + 8) * kSkPitch + ks * 8 + t] = 0u;
                    stage_e[g * kSkPitch + ks * 8 + 4 + t] = 0u; stage_e[(g + 8) * kSkPitch + ks * 8 + 4 + t] = 0u;
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t lo, hi;
                if (EPI) {
                    const float2 bv = bias ? __ldg(reinterpret_cast<const float2*>(bias + 8 * j + 2 * t)) : make_float2(0.f, 0.f);      // L1-resident
                    lo = act_bf16x2(pack_bf16x2(acc[j][0] + bv.x, acc[j][1] + bv.y), act, slope2);
                    hi = act_bf16x2(pack_bf16x2(acc[j][2] + bv.x, acc[j][3] + bv.y), act, slope2);
                } else {
                    lo = pack_bf16x2(acc[j][0], acc[j][1]); hi = pack_bf16x2(acc[j][2], acc[j][3]);
                }
                stage_o[g * kSkPitch + j * 4 + t] = live[0] ? lo : 0u;
                stage_o[(g + 8) * kSkPitch + j * 4 + t] = live[1] ? hi : 0u;
            }
            __syncwarp();
            // 16 rows x 128 bytes: lane -> (row, 16-byte chunk), four rows per instruction, consecutive rows are consecutive in memory
            {
                const int r = lane >> 3, ch = lane & 7;
                bf16* op = out + (qbase + i0 + r) * 64 + ch * 8;
                bf16* ep = KEEP ? p.dst + (qbase + i0 + r) * 64 + ch * 8 : nullptr;
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    if (i0 + it * 4 + r < nrows) {
                        *reinterpret_cast<uint4*>(op + it * 4 * 64) = *reinterpret_cast<const uint4*>(stage_o + (it * 4 + r) * kSkPitch + ch * 4);
                        if (KEEP) *reinterpret_cast<uint4*>(ep + it * 4 * 64) = *reinterpret_cast<const uint4*>(stage_e + (it * 4 + r) * kSkPitch + ch * 4);
                    }
                }
            }
            __syncwarp();
            ii += 8 * 16;
            while (ii >= ll) { ii -= ll; ++l; }
        }
        __syncthreads();                            // every warp is done reading this tile
        if (nbuf == 2) cur ^= 1;
        else if (nxt < total) fill(tile_base, nxt);
    }
}

struct Col2imP {
    const bf16* de; long long ld;
    int C, c_first, c_out, H, W, k, stride, pad;
    const float* scale;
    RowMap rm;
    float* out; int accumulate;
};

__global__ void col2im_kernel(const Col2imP p) {
    irc::pdl_prologue();
    const long long total = (long long)p.rm.n_img * p.c_out * p.H * p.W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        long long t = idx;
        const int x = (int)(t % p.W); t /= p.W;
        const int y = (int)(t % p.H); t /= p.H;
        const int co = (int)(t % p.c_out);
        const int n = (int)(t / p.c_out);
        const int c = p.c_first + co;
        float acc = 0.f;
        for (int r = 0; r < p.k; ++r) {
            const int ty = y + p.pad - r;
            if (ty < 0 || ty % p.stride) continue;
            const int oy = ty / p.stride;
            if (oy >= p.rm.Ho) continue;
            for (int s = 0; s < p.k; ++s) {
                const int tx = x + p.pad - s;
                if (tx < 0 || tx % p.stride) continue;
                const int ox = tx / p.stride;
                if (ox >= p.rm.Wo) continue;
                acc += __bfloat162float(p.de[p.rm.encode(n, oy, ox) * p.ld + (r * p.k + s) * p.C + c]);
            }
        }
        if (p.scale) acc *= __ldg(p.scale + c);
        if (p.accumulate) p.out[idx] += acc; else p.out[idx] = acc;
    }
}

// Row-staged version: one block per (image, output row).  The (at most k) operand rows that contribute are read with
// coalesced 16-byte loads - only the 16-byte chunks that hold the k*C columns of that kernel row - and transposed into
// shared memory as T[kernel row][column][ox] (fp32), so that the k*k taps of an output pixel are conflict-free reads with
// lanes running along x.  Same summation order as col2im_kernel (r outer, s inner): identical bits.
template <int S>      // stride as a compile-time constant: the tap tests below are shifts and masks, not integer divisions
__global__ void __launch_bounds__(256) col2im_rows_kernel(const Col2imP p, int pitch) {
    irc::pdl_prologue();
    extern __shared__ float T[];                  // [k][k*C][pitch]
    const int n = blockIdx.y, y = blockIdx.x;
    const int kC = p.k * p.C, Wo = p.rm.Wo;
    for (int r = 0; r < p.k; ++r) {
        const int ty = y + p.pad - r;
        if (ty < 0 || ty % S || ty / S >= p.rm.Ho) continue;          // uniform over the block
        const int oy = ty / S;
        const int ch_lo = (r * kC) >> 3, nch = (((r + 1) * kC - 1) >> 3) - ch_lo + 1;
        float* Tr = T + (size_t)r * kC * pitch;
        for (int idx = threadIdx.x; idx < Wo * nch; idx += blockDim.x) {
            const int ox = idx / nch, g = ch_lo + (idx - ox * nch);
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(p.de + p.rm.encode(n, oy, ox) * p.ld + g * 8));
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = g * 8 + j - r * kC;
                if (col >= 0 && col < kC) {
                    const uint32_t h = (j & 1) ? (w[j >> 1] & 0xffff0000u) : (w[j >> 1] << 16);
                    Tr[col * pitch + ox] = __uint_as_float(h);
                }
            }
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < p.c_out * p.W; idx += blockDim.x) {
        const int co = idx / p.W, x = idx - co * p.W;
        const int c = p.c_first + co;
        float acc = 0.f;
        for (int r = 0; r < p.k; ++r) {
            const int ty = y + p.pad - r;
            if (ty < 0 || ty % S || ty / S >= p.rm.Ho) continue;
            const float* Tr = T + ((size_t)r * kC + c) * pitch;
            for (int s_ = 0; s_ < p.k; ++s_) {
                const int tx = x + p.pad - s_;
                if (tx < 0 || tx % S) continue;
                const int ox = tx / S;
                if (ox >= Wo) continue;
                acc += Tr[(size_t)s_ * p.C * pitch + ox];
            }
        }
        if (p.scale) acc *= __ldg(p.scale + c);
        const long long o = (((long long)n * p.c_out + co) * p.H + y) * p.W + x;
        if (p.accumulate) p.out[o] += acc; else p.out[o] = acc;
    }
}

struct TapP {
    int nshift, nco;
    int shifts[IRC_MAX_TAPS];
    signed char dy[IRC_MAX_TAPS], dx[IRC_MAX_TAPS];   // shift = dy*wp + dx with |dx| < wp/2
    int n_img, H, W, hp, wp, oy, ox;
    int live_cols_only;      // 1: only the 16-byte column groups that hold tap columns are written (the rest stays as the caller zeroed it)
};

// out[n][co][y][x] = act(bias[co] + sum_j P[q + shift_j][j*nco + co])
__global__ void __launch_bounds__(256) tap_reduce_kernel(const float* P, long long ldp, const float* bias, int act, float* out, const TapP t) {
    irc::pdl_prologue();
    // grid = (x chunks, rows, images): no per-element index arithmetic
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, n = blockIdx.z;
    if (x >= t.W) return;
    const long long q = ((long long)n * t.hp + y + t.oy) * t.wp + x + t.ox;
    for (int co = 0; co < t.nco; ++co) {
        float acc = bias ? __ldg(bias + co) : 0.f;
        for (int j = 0; j < t.nshift; ++j) acc += __ldg(P + (q + t.shifts[j]) * ldp + j * t.nco + co);
        if (act == 3) acc = tanhf(acc);
        out[(((long long)n * t.nco + co) * t.H + y) * t.W + x] = acc;
    }
}

// E[q][j*nco + co] = g'[pixel(q - shift_j)][co];  g' = g * (1 - yv^2) when yv is given (tanh').
// A shifted position that leaves its frame line lands in the padding ring (the ring is at least as wide as the
// largest horizontal shift), i.e. on a zero, so the (dy, dx) decomposition needs no wrap-around handling.
__global__ void __launch_bounds__(256) tap_expand_kernel(const float* g, const float* yv, bf16* E, const TapP t, int ndy, float* rowsum) {
    irc::pdl_prologue();
    // One block per (image, padded row).  The <= 8 distinct source rows the taps of this row read (one per distinct dy) are
    // staged in shared memory as g' = g * (1 - y^2), all channels, zero where the row falls outside the image; a thread
    // (pixel X, group of 8 columns) then assembles its 16 bytes from shared memory at register-resident offsets.
    extern __shared__ float T[];                 // [ndy][nco][W]
    __shared__ int s_dy[8];
    const int ncol = t.nshift * t.nco;
    const int grp = threadIdx.x & 7;
    if (threadIdx.x == 0) {                      // distinct dy values in first-appearance order (same rule as below)
        int m = 0;
        for (int j = 0; j < t.nshift; ++j) {
            bool seen = false;
            for (int i = 0; i < m; ++i) seen |= s_dy[i] == t.dy[j];
            if (!seen && m < 8) s_dy[m++] = t.dy[j];
        }
    }
    __syncthreads();
    int cdx[8], cbase[8]; unsigned valid = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int col = grp * 8 + k;
        cdx[k] = cbase[k] = 0;
        if (col < ncol) {
            const int j = col / t.nco, co = col - j * t.nco;
            int sl = 0;
            for (int i = 0; i < ndy; ++i) if (s_dy[i] == t.dy[j]) sl = i;
            valid |= 1u << k; cdx[k] = t.dx[j] + t.ox; cbase[k] = (sl * t.nco + co) * t.W;
        }
    }
    const long long hw = (long long)t.H * t.W;
    for (int row = blockIdx.x; row < t.n_img * t.hp; row += gridDim.x) {
        const int n = row / t.hp, Y = row - n * t.hp;
        __syncthreads();
        // staging with ALL threads and every global load of the row in flight at once (a warp per source row made the 8 dependent
        // load latencies of its row the critical path: 45 % barrier stalls, profiles/r3); the row sums come from shared memory after
        const int tot = ndy * t.nco * t.W;
        for (int e0 = threadIdx.x; e0 < tot; e0 += 4 * 256) {
            float gv[4], yy[4]; int er[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + u * 256;
                gv[u] = 0.f; yy[u] = 0.f; er[u] = -1;
                if (e < tot) {
                    const int r = e / t.W, x = e - r * t.W;
                    const int sl = r / t.nco, co = r - sl * t.nco;
                    const int y = Y - s_dy[sl] - t.oy;
                    er[u] = e;
                    if (y >= 0 && y < t.H) {
                        const long long o = ((long long)n * t.nco + co) * hw + (long long)y * t.W + x;
                        gv[u] = __ldg(g + o);
                        if (yv) yy[u] = __ldg(yv + o);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) if (er[u] >= 0) T[er[u]] = gv[u] * (1.f - yy[u] * yy[u]);
        }
        if (rowsum) {
            __syncthreads();
            // bias gradient: every image row is staged exactly once in the slot of the first vertical shift (host-checked), so the
            // per-(frame row, channel) sums of that slot add up to sum g' over the image; irc_tap_expand adds them in a fixed order
            const int r = threadIdx.x >> 5;
            if (r < t.nco) {
                float acc = 0.f;
                for (int x = threadIdx.x & 31; x < t.W; x += 32) acc += T[r * t.W + x];
                acc = warp_sum(acc);
                if ((threadIdx.x & 31) == 0) rowsum[(long long)row * t.nco + r] = acc;
            }
        }
        __syncthreads();
        if (t.live_cols_only && !valid) continue;        // (uniform per 8-column group; the next iteration's barrier still sees every thread)
        for (int X = threadIdx.x >> 3; X < t.wp; X += blockDim.x >> 3) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int x = X - cdx[k];
                v[k] = (((valid >> k) & 1) && x >= 0 && x < t.W) ? T[cbase[k] + x] : 0.f;
            }
            *reinterpret_cast<uint4*>(E + ((long long)row * t.wp + X) * 64 + grp * 8) =
                make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        }
    }
}

// Horizontal taps only (one vertical shift), NS x NCO <= 24 columns known at compile time - the backward expansion of the generator's
// output head (7 taps x 3 channels, irc:527-531).  One thread = one frame pixel: per tap one bounds test and NCO shared-memory reads
// at compile-time register positions, three 16-byte stores (the generic kernel spends ~180 instructions per 8-column group on
// per-column predicates: 62 us, issue-bound, profiles/r3).  Blocks walk frame rows; staging as in the generic kernel.
template <int NS, int NCO>
__global__ void __launch_bounds__(256) tap_expand_row_kernel(const float* __restrict__ g, const float* __restrict__ yv, bf16* __restrict__ E, const TapP t,
                                                             float* rowsum) {
    irc::pdl_prologue();
    extern __shared__ float T[];                 // [NCO][W]
    constexpr int NV = (NS * NCO + 7) / 8;       // 16-byte groups that hold tap columns
    const long long hw = (long long)t.H * t.W;
    const int dy0 = t.dy[0];
    for (int row = blockIdx.x; row < t.n_img * t.hp; row += gridDim.x) {
        const int n = row / t.hp, Y = row - n * t.hp;
        const int y = Y - dy0 - t.oy;
        const bool ok = y >= 0 && y < t.H;
        __syncthreads();
        const int tot = NCO * t.W;
        for (int e0 = threadIdx.x; e0 < tot; e0 += 4 * 256) {
            float gv[4], yy[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + u * 256;
                gv[u] = 0.f; yy[u] = 0.f;
                if (ok && e < tot) {
                    const int co = e / t.W, x = e - co * t.W;
                    const long long o = ((long long)n * NCO + co) * hw + (long long)y * t.W + x;
                    gv[u] = __ldg(g + o);
                    if (yv) yy[u] = __ldg(yv + o);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) if (e0 + u * 256 < tot) T[e0 + u * 256] = gv[u] * (1.f - yy[u] * yy[u]);
        }
        __syncthreads();
        if (rowsum && (threadIdx.x >> 5) < NCO) {
            const int r = threadIdx.x >> 5;
            float acc = 0.f;
            for (int x = threadIdx.x & 31; x < t.W; x += 32) acc += T[r * t.W + x];
            acc = warp_sum(acc);
            if ((threadIdx.x & 31) == 0) rowsum[(long long)row * NCO + r] = acc;
        }
        for (int X = threadIdx.x; X < t.wp; X += 256) {
            float v[NV * 8];
#pragma unroll
            for (int i = 0; i < NV * 8; ++i) v[i] = 0.f;
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                const int x = X - t.dx[j] - t.ox;
                if (x >= 0 && x < t.W) {
#pragma unroll
                    for (int co = 0; co < NCO; ++co) v[j * NCO + co] = T[co * t.W + x];
                }
            }
            uint4* ep = reinterpret_cast<uint4*>(E + ((long long)row * t.wp + X) * 64);
#pragma unroll
            for (int q = 0; q < NV; ++q)
                ep[q] = make_uint4(pack_bf16x2(v[q * 8], v[q * 8 + 1]), pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]), pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]),
                                   pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]));
            // whole 32-byte sectors: a half-written sector costs the L2 a read-modify-write
            constexpr int NVS = (NV + 1) & ~1;
#pragma unroll
            for (int q = NV; q < (NVS < 8 ? NVS : 8); ++q) ep[q] = make_uint4(0, 0, 0, 0);
            if (!t.live_cols_only) {
#pragma unroll
                for (int q = NVS; q < 8; ++q) ep[q] = make_uint4(0, 0, 0, 0);
            }
        }
    }
}

// part[b][c] = partial sum over block b's slice of sum_{n, pixels} g[n][c][.] * (1 - yv^2); chan_sum_final adds the
// partials in a fixed order (bit-reproducible)
__global__ void chan_sum_kernel(const float* g, const float* yv, int n_img, int C, long long hw, float* part) {
    irc::pdl_prologue();
    __shared__ float sh[32];
    const int c = blockIdx.y;
    float s = 0.f;
    const long long total = (long long)n_img * hw;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / hw, r = i - n * hw;
        const long long o = (n * C + c) * hw + r;
        float v = __ldg(g + o);
        if (yv) { const float yy = __ldg(yv + o); v *= (1.f - yy * yy); }
        s += v;
    }
    s = block_sum(s, sh);
    if (threadIdx.x == 0) part[(long long)blockIdx.x * C + c] = s;
}
// out[c] = sum over the rows of part[row][c]: thread t adds rows t, t + 1024, ... in order, then a fixed tree over the 1024 partials
__global__ void __launch_bounds__(1024) rowsum_final_kernel(const float* __restrict__ part, int nrows, int C, float* __restrict__ out) {
    irc::pdl_prologue();
    __shared__ float sh[1024];
    for (int c = 0; c < C; ++c) {
        float a = 0.f;
        for (int r = threadIdx.x; r < nrows; r += 1024) a += part[(long long)r * C + c];
        sh[threadIdx.x] = a;
        __syncthreads();
        for (int o = 512; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[c] = sh[0];
        __syncthreads();
    }
}
__global__ void chan_sum_final(const float* part, int nb, int C, float* out) {
    irc::pdl_prologue();
    const int c = threadIdx.x;
    if (c < C) { float a = 0.f; for (int b = 0; b < nb; ++b) a += part[(long long)b * C + c]; out[c] = a; }
}

int grid_for(long long total, int threads) {
    long long b = (total + threads - 1) / threads;
    const long long cap = (long long)irc_num_sms() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

extern "C" long long irc_im2col_rows(int row_mode, int n_img, int Ho, int Wo) {
    RowMap rm; rm.mode = row_mode; rm.n_img = n_img; rm.Ho = Ho; rm.Wo = Wo;
    return rm.rows();
}

extern "C" int irc_im2col(const irc_im2col_args* a, void* stream) {
    if (!a->src1 || !a->dst) return irc_set_error(IRC_ERR_BAD_ARG, "irc_im2col: null");
    const int C = a->c1 + (a->src2 ? a->c2 : 0);
    if (a->k * a->k * C > 64) return irc_set_error(IRC_ERR_BAD_ARG, "irc_im2col: k*k*C must be <= 64");
    if (a->row_mode == 2 && ((a->Ho | a->Wo) & 1)) return irc_set_error(IRC_ERR_BAD_ARG, "irc_im2col: row_mode 2 needs even output extents");
    Im2colP p;
    p.src1 = a->src1; p.src2 = a->src2; p.c1 = a->c1; p.c2 = a->src2 ? a->c2 : 0;
    p.scale = a->scale; p.shift = a->shift;
    p.H = a->H; p.W = a->W; p.k = a->k; p.stride = a->stride; p.pad = a->pad; p.pad_mode = a->pad_mode;
    p.rm.mode = a->row_mode; p.rm.n_img = a->n_img; p.rm.Ho = a->Ho; p.rm.Wo = a->Wo;
    p.dst = (bf16*)a->dst; p.row_img = a->row_img;
    // LPB consecutive lines share one staged tile (amortises the fill and its two barriers); a line holds one output row
    // (two in the space-to-depth order), so the tile needs (rows_out - 1) * stride + k input rows
    const int LPB = a->row_mode == 2 ? 2 : 8;
    const int rows_out = LPB * (a->row_mode == 2 ? 2 : 1);
    const int R = (rows_out - 1) * a->stride + a->k, Wp = a->W + 2 * a->pad;
    const size_t smem = (size_t)C * R * Wp * sizeof(float);
    if (smem > 200 * 1024) return irc_set_error(IRC_ERR_BAD_ARG, "irc_im2col: line tile does not fit shared memory");
    static size_t attr = 48 * 1024;
    if (smem > attr) { cudaFuncSetAttribute(im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = 200 * 1024; }
    const long long nblk = (long long)p.rm.n_img * ((p.rm.lines() + LPB - 1) / LPB);
    irc::launch(im2col_kernel, (unsigned)(nblk < 1048576 ? nblk : 1048576), 256, smem, (cudaStream_t)stream, p, R, Wp, LPB);
    return irc_check_launch("irc_im2col");
}

/* Direct small-K convolution: same geometry arguments as irc_im2col; computes out[q][0..64) = act(bias + sum_k E[q][k] * w[n][k])
 * for the im2col operand E without materialising it (args->dst may be NULL; when given it ALSO receives E, for the weight gradient). */
extern "C" int irc_smallk_conv_fwd(const irc_im2col_args* a, const void* w, const float* bias, int act, float slope, void* out, void* stream) {
    if (!a || !a->src1 || !w || !out) return irc_set_error(IRC_ERR_BAD_ARG, "irc_smallk_conv_fwd: null");
    const int C = a->c1 + (a->src2 ? a->c2 : 0);
    const int K = a->k * a->k * C;
    if (K > 64) return irc_set_error(IRC_ERR_BAD_ARG, "irc_smallk_conv_fwd: k*k*C must be <= 64");
    if (a->row_mode == 2 && ((a->Ho | a->Wo) & 1)) return irc_set_error(IRC_ERR_BAD_ARG, "irc_smallk_conv_fwd: row_mode 2 needs even output extents");
    if (((uintptr_t)out & 15) || ((uintptr_t)a->dst & 15) || ((uintptr_t)w & 3)) return irc_set_error(IRC_ERR_BAD_ARG, "irc_smallk_conv_fwd: misaligned operand");
    Im2colP p;
    p.src1 = a->src1; p.src2 = a->src2; p.c1 = a->c1; p.c2 = a->src2 ? a->c2 : 0;
    p.scale = a->scale; p.shift = a->shift;
    p.H = a->H; p.W = a->W; p.k = a->k; p.stride = a->stride; p.pad = a->pad; p.pad_mode = a->pad_mode;
    p.rm.mode = a->row_mode; p.rm.n_img = a->n_img; p.rm.Ho = a->Ho; p.rm.Wo = a->Wo;
    p.dst = (bf16*)a->dst; p.row_img = a->row_img;
    // lines per block: as many as keep the tile within half of the shared memory (two resident blocks overlap one block's
    // tile fill with the other's arithmetic); wide images fall back to fewer lines, then to one block per SM
    const bool keep = a->dst != nullptr;
    const int Wp = a->W + 2 * a->pad;
    const size_t stage_bytes = 8 * (keep ? 2 : 1) * 16 * kSkPitch * sizeof(uint32_t);
    int LPB = a->row_mode == 2 ? 2 : 4, R = 0, nbuf = 2;
    size_t smem = 0, tile_bytes = 0;
    for (;; LPB >>= 1) {
        const int rows_out = LPB * (a->row_mode == 2 ? 2 : 1);
        R = (rows_out - 1) * a->stride + a->k;
        tile_bytes = (size_t)C * R * Wp * sizeof(float);
        smem = 2 * tile_bytes + stage_bytes;
        if (smem <= 100 * 1024 || LPB == 1) break;
    }
    if (smem > 100 * 1024) { nbuf = 1; smem = tile_bytes + stage_bytes; }        // wide images: one tile, filled between the groups
    if (smem > 200 * 1024) return irc_set_error(IRC_ERR_BAD_ARG, "irc_smallk_conv_fwd: line tile does not fit shared memory");
    // persistent blocks (two per SM): each walks its groups of lines with the next tile in flight
    const long long nblk = (long long)p.rm.n_img * ((p.rm.lines() + LPB - 1) / LPB);
    const long long cap = 2LL * irc_num_sms();
    const unsigned grid = (unsigned)(nblk < cap ? nblk : cap);
    const int ks = K <= 32 ? 2 : 4;
    bool ok = false;
    const bool epi = bias != nullptr || act != 0;
    if (act < 0 || act > 2 || (act == 2 && !(slope >= 0.f && slope < 1.f)))
        return irc_set_error(IRC_ERR_BAD_ARG, "irc_smallk_conv_fwd: act must be 0, 1 (ReLU) or 2 (LeakyReLU with 0 <= slope < 1)");
#define IRC_SMALLK_CASE2(KS_, MODE_, KEEP_, EPI_)                                                                                                  \
    if (!ok && ks == KS_ && a->row_mode == MODE_ && keep == KEEP_ && epi == EPI_) {                                                               \
        static bool attr = false;                                                                                                                  \
        if (!attr) { cudaFuncSetAttribute(smallk_conv_fwd_kernel<KS_, MODE_, KEEP_, EPI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; } \
        irc::launch(smallk_conv_fwd_kernel<KS_, MODE_, KEEP_, EPI_>, grid, 256, smem, (cudaStream_t)stream, p, (const bf16*)w, bias, act, slope, (bf16*)out, R, Wp, LPB, nbuf); \
        ok = true;                                                                                                                                 \
    }
#define IRC_SMALLK_CASE(KS_, MODE_, KEEP_) IRC_SMALLK_CASE2(KS_, MODE_, KEEP_, false) IRC_SMALLK_CASE2(KS_, MODE_, KEEP_, true)
    IRC_SMALLK_CASE(2, 0, false) IRC_SMALLK_CASE(2, 0, true) IRC_SMALLK_CASE(2, 1, false) IRC_SMALLK_CASE(2, 1, true)
    IRC_SMALLK_CASE(2, 2, false) IRC_SMALLK_CASE(2, 2, true) IRC_SMALLK_CASE(4, 0, false) IRC_SMALLK_CASE(4, 0, true)
    IRC_SMALLK_CASE(4, 1, false) IRC_SMALLK_CASE(4, 1, true) IRC_SMALLK_CASE(4, 2, false) IRC_SMALLK_CASE(4, 2, true)
#undef IRC_SMALLK_CASE
#undef IRC_SMALLK_CASE2
    if (!ok) return irc_set_error(IRC_ERR_BAD_ARG, "irc_smallk_conv_fwd: bad row_mode %d", a->row_mode);
    return irc_check_launch("irc_smallk_conv_fwd");
}

extern "C" int irc_col2im(const irc_col2im_args* a, void* stream) {
    if (!a->de || !a->out) return irc_set_error(IRC_ERR_BAD_ARG, "irc_col2im: null");
    Col2imP p;
    p.de = (const bf16*)a->de; p.ld = a->ld; p.C = a->C; p.c_first = a->c_first; p.c_out = a->c_out;
    p.H = a->H; p.W = a->W; p.k = a->k; p.stride = a->stride; p.pad = a->pad; p.scale = a->scale;
    p.rm.mode = a->row_mode; p.rm.n_img = a->n_img; p.rm.Ho = a->Ho; p.rm.Wo = a->Wo;
    p.out = a->out; p.accumulate = a->accumulate;
    const long long total = (long long)a->n_img * a->c_out * a->H * a->W;
    const int pitch = a->Wo | 1;                                  // odd pitch: the transposing stores spread over the banks
    const size_t smem = (size_t)a->k * a->k * a->C * pitch * sizeof(float);
    if (a->k * a->k * a->C <= 64 && !(a->ld % 8) && !((uintptr_t)a->de & 15) && smem <= 160 * 1024 && a->H <= 65535 && a->n_img <= 65535 &&
        (a->stride == 1 || a->stride == 2)) {
        static size_t attr = 48 * 1024;
        if (smem > attr) {
            cudaFuncSetAttribute(col2im_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
            cudaFuncSetAttribute(col2im_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
            attr = 160 * 1024;
        }
        if (a->stride == 1) irc::launch(col2im_rows_kernel<1>, dim3(a->H, a->n_img), 256, smem, (cudaStream_t)stream, p, pitch);
        else irc::launch(col2im_rows_kernel<2>, dim3(a->H, a->n_img), 256, smem, (cudaStream_t)stream, p, pitch);
        return irc_check_launch("irc_col2im(rows)");
    }
    irc::launch(col2im_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, p);
    return irc_check_launch("irc_col2im");
}

static int fill_tap(const irc_tap_args* a, TapP& t) {
    if (a->nshift <= 0 || a->nshift > IRC_MAX_TAPS || a->nshift * a->nco > 64) return irc_set_error(IRC_ERR_BAD_ARG, "irc_tap: nshift*nco must be <= 64");
    t.nshift = a->nshift; t.nco = a->nco;
    for (int i = 0; i < a->nshift; ++i) {
        const int dy = a->dy[i], dx = a->dx[i];
        if (dy < -127 || dy > 127 || dx < -127 || dx > 127) return irc_set_error(IRC_ERR_BAD_ARG, "irc_tap: shift out of range");
        t.shifts[i] = dy * a->wp + dx;
        t.dy[i] = (signed char)dy; t.dx[i] = (signed char)dx;
    }
    t.n_img = a->n_img; t.H = a->H; t.W = a->W; t.hp = a->hp; t.wp = a->wp; t.oy = a->oy; t.ox = a->ox;
    t.live_cols_only = a->live_cols_only;
    return IRC_OK;
}

extern "C" int irc_tap_reduce(const irc_tap_args* a, const float* P, long long ldp, const float* bias, int act, float* out, void* stream) {
    TapP t; int rc = fill_tap(a, t); if (rc) return rc;
    if (!P || !out) return irc_set_error(IRC_ERR_BAD_ARG, "irc_tap_reduce: null");
    if (t.H > 65535 || t.n_img > 65535) return irc_set_error(IRC_ERR_BAD_ARG, "irc_tap_reduce: extent too large");
    const int tb = t.W >= 256 ? 256 : ((t.W + 31) / 32) * 32;
    irc::launch(tap_reduce_kernel, dim3((t.W + tb - 1) / tb, t.H, t.n_img), tb, 0, (cudaStream_t)stream, P, ldp, bias, act, out, t);
    return irc_check_launch("irc_tap_reduce");
}

extern "C" int irc_tap_expand(const irc_tap_args* a, const float* g, const float* y, void* E, float* dbias, float* work, long long work_floats,
                              void* stream) {
    TapP t; int rc = fill_tap(a, t); if (rc) return rc;
    if (!g || !E) return irc_set_error(IRC_ERR_BAD_ARG, "irc_tap_expand: null");
    const long long nrow = (long long)t.n_img * t.hp;
    int ndy = 0, seen_dy[8];
    for (int j = 0; j < t.nshift; ++j) {
        bool seen = false;
        for (int i = 0; i < ndy; ++i) seen |= seen_dy[i] == t.dy[j];
        if (!seen) { if (ndy == 8) return irc_set_error(IRC_ERR_BAD_ARG, "irc_tap_expand: more than 8 distinct vertical shifts"); seen_dy[ndy++] = t.dy[j]; }
    }
    const size_t smem = (size_t)ndy * t.nco * t.W * sizeof(float);
    if (smem > 48 * 1024) return irc_set_error(IRC_ERR_BAD_ARG, "irc_tap_expand: row tile does not fit shared memory");
    // a block pays ~250 instructions per thread of column decoding before its first row: let it walk several rows
    const long long cap = 8LL * irc_num_sms();
    // bias gradient from the staged rows (no second read of g and y) when every image row is staged exactly once under the first
    // vertical shift, i.e. its frame row y + dy0 + oy exists for all y, and the workspace holds one partial per (frame row, channel)
    const int dy0 = t.dy[0];
    const bool fused_dbias = dbias && work && work_floats >= nrow * t.nco && t.nco <= 8 && dy0 + t.oy >= 0 && t.H - 1 + dy0 + t.oy < t.hp;
    if (ndy == 1 && t.nshift == 7 && t.nco == 3)
        irc::launch(tap_expand_row_kernel<7, 3>, (unsigned)(nrow < cap ? nrow : cap), 256, smem, (cudaStream_t)stream, g, y, (bf16*)E, t, fused_dbias ? work : nullptr);
    else
        irc::launch(tap_expand_kernel, (unsigned)(nrow < cap ? nrow : cap), 256, smem, (cudaStream_t)stream, g, y, (bf16*)E, t, ndy, fused_dbias ? work : nullptr);
    rc = irc_check_launch("irc_tap_expand"); if (rc) return rc;
    if (fused_dbias) {
        irc::launch(rowsum_final_kernel, 1, 1024, 0, (cudaStream_t)stream, (const float*)work, (int)nrow, t.nco, dbias);
        rc = irc_check_launch("irc_tap_expand(dbias)");
    } else if (dbias) {
        const long long hw = (long long)t.H * t.W;
        long long nb = ((long long)t.n_img * hw + 4095) / 4096;
        if (nb > 512) nb = 512;
        if (!work || work_floats < nb * t.nco) nb = 1;
        float* part = nb == 1 ? dbias : work;
        irc::launch(chan_sum_kernel, dim3((unsigned)nb, t.nco), 256, 0, (cudaStream_t)stream, g, y, t.n_img, t.nco, hw, part);
        if (nb > 1) irc::launch(chan_sum_final, 1, 64, 0, (cudaStream_t)stream, work, (int)nb, t.nco, dbias);
        rc = irc_check_launch("irc_tap_expand(dbias)");
    }
    return rc;
}
