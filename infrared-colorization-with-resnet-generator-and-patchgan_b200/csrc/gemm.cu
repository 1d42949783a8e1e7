// tcgen05 / TMEM / TMA implicit-GEMM kernels (sm_100a).
//
//   conv_gemm : out[q][n] = epilogue( sum_t sum_c A[q + tap_t][c] * W[n][t*Cin + c] )
//               A is an NHWC bf16 activation buffer seen as a flat [rows][channels] matrix
//               ("framed" layout: the padding ring is stored, so every conv tap of a
//               stride-1 convolution is a constant row shift).  Used for the forward
//               convolutions (irc:458-531, irc:598-630, irc:664) and, with negated shifts
//               and transposed weights, for their data gradients.
//   tn_gemm   : out[t][m][n] = sum_q A[q + sa_t][m] * B[q + sb_t][n]   (reduction over rows)
//               = weight gradients; both operands are consumed MN-major straight from the
//               NHWC buffers, no transposed copies.
//
// Both: one CTA = 192 threads = TMA producer warp, MMA issuer warp, 4 epilogue warps;
// operands staged by TMA into SWIZZLE_128B shared memory, accumulators in TMEM.
#include <stdlib.h>
#include <string.h>

#include "irc_common.cuh"
#include "../../include/irc_b200.h"

using namespace irc;

namespace {

constexpr int kThreads = 192;          // tn_gemm / tap-run kernels: producer, MMA, 4 epilogue warps
constexpr int kConvThreads = 320;      // conv_gemm: producer, MMA, 8 epilogue warps (two per TMEM lane quarter)
constexpr int kBM = 128;          // rows (pixels) per tile = UMMA M
constexpr int kBK = 64;           // bf16 elements per 128-byte swizzled row
constexpr int kMaxSmem = 232448;  // 227 KB

// ------------------------------------------------------------------------------------
// host: tensor-map encoding through the driver entry point (no link-time libcuda)
// ------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// 2-D bf16 matrix [rows][ld] (row-major), box = box_rows x 64 elements, 128-byte swizzle.
int make_map(CUtensorMap* m, const void* base, long long rows, int ld, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return irc_set_error(IRC_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
    if (((uintptr_t)base & 15) || (ld % 8)) return irc_set_error(IRC_ERR_BAD_ARG, "TMA operand must be 16-byte aligned with ld %% 8 == 0");
    cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return irc_set_error(IRC_ERR_DRIVER, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return IRC_OK;
}

// ------------------------------------------------------------------------------------
// conv_gemm
// ------------------------------------------------------------------------------------
struct ConvParams {
    long long rows;
    int n_tiles, bn, k_chunks, a_chan_off, ntaps, stages;
    int k16;         // 16-column reduction steps issued per 64-column chunk (4; fewer when the operand's last columns are structural zeros)
    int mt;          // M sub-tiles of 128 rows per tile (1 or 2): two sub-tiles share every B stage
    int nbuf;        // TMEM accumulator buffers (2 = epilogue overlaps the next tile's MMAs)
    int taps[IRC_MAX_TAPS];
    void* out;
    long long out_ld;
    int out_chan_off, out_fp32;
    const float* bias;
    int act;
    float slope;
    const short* row_img;
    const bf16* mask;
    long long mask_ld;
    int mask_chan_off;
    float mask_slope;
    const bf16* addend;      // optional bf16 [rows][addend_ld]: added to the accumulator (residual-stream gradient)
    long long addend_ld;
    int addend_chan_off;
    // InstanceNorm statistics from the epilogue (staged bf16 path only): per 128-row sub-tile and output channel the (sum, sum of
    // squares) of the bf16-ROUNDED values of the sub-tile's live rows, image slot 0 -> stats_part[subtile][n_out][2]; the rows of
    // a sub-tile that already belong to the next image -> stats_edge[image][n_out][2].  irc_conv_stats_finalize adds them up in
    // sub-tile order (bit-reproducible).  Null = off.
    float* stats_part;
    float* stats_edge;
    int rows_per_img;
    int n_out;
    // tile geometry: tile t covers rows [t * tile_stride + row_bias, ... + 128 * mt); stride < 128 * mt = overlapping tiles
    int tile_stride, row_bias;
    // horizontal tap reduction fused into the epilogue (tiny-Cout k x k convolutions computed as a GEMM over the kernel rows):
    // out[n][co][y][x] = act(bias[co] + sum_j acc[q + j - halo][j * nco + co]); tiles overlap by 2 * halo rows
    float* tap_out;
    int tap_nshift, tap_nco, tap_H, tap_W, tap_hp, tap_wp, tap_oy, tap_ox, tap_act;
    const float* tap_scale;  // optional per-plane factor
    int tap_accumulate;      // 1: tap_out += result
    int nstg;       // staging tiles of the TMA-store epilogue (2, or 1 when shared memory is needed for pipeline stages)
    int tma_store;  // 1: bf16 rows leave through a swizzled shared-memory staging tile and TMA stores (coalesced)
};

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v > 0.f ? v : v * slope;
    return v;
}

// ---- epilogue math: branch-free per element (one warp per SM sub-partition cannot hide branch / latency chains) ----
struct EpiCtx {
    float slope_eff;     // activation as max(v,0) + slope_eff * min(v,0): 1 = identity, 0 = ReLU, s = LeakyReLU(s)
    const float* sbias;  // bias staged in shared memory, or null
};

// mask / addend rows of one 32-column chunk, fetched EARLY (before the barrier and the TMEM wait of the staged epilogue) so that
// their global-load latency overlaps that wait instead of sitting on the per-chunk critical path of shallow-K layers
struct EpiPre { uint4 m[4]; uint4 a[4]; };

__device__ __forceinline__ void epi_prefetch(const ConvParams& p, EpiPre& pre, long long row, int col, bool in_range) {
    const bool live = in_range;
    if (p.mask && live) {
        const uint4* mp = reinterpret_cast<const uint4*>(p.mask + row * p.mask_ld + p.mask_chan_off + col);
#pragma unroll
        for (int g = 0; g < 4; ++g) pre.m[g] = __ldg(mp + g);
    }
    if (p.addend && live) {
        const uint4* ap = reinterpret_cast<const uint4*>(p.addend + row * p.addend_ld + p.addend_chan_off + col);
#pragma unroll
        for (int g = 0; g < 4; ++g) pre.a[g] = __ldg(ap + g);
    }
}

__device__ __forceinline__ void epi_math32(const ConvParams& p, const EpiCtx& e, const uint32_t (&r)[32], float (&v)[32], long long row, int col,
                                           bool live, const EpiPre* pre = nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (e.sbias) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const float4 b = *reinterpret_cast<const float4*>(e.sbias + col + g * 4);
            v[g * 4] += b.x; v[g * 4 + 1] += b.y; v[g * 4 + 2] += b.z; v[g * 4 + 3] += b.w;
        }
    }
    if (p.mask && live) {
        const uint4* mp = reinterpret_cast<const uint4*>(p.mask + row * p.mask_ld + p.mask_chan_off + col);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const uint4 mv = pre ? pre->m[g] : __ldg(mp + g);
            const uint32_t w[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 f = unpack_bf16x2(w[q]);
                v[g * 8 + q * 2] *= f.x > 0.f ? 1.f : p.mask_slope;
                v[g * 8 + q * 2 + 1] *= f.y > 0.f ? 1.f : p.mask_slope;
            }
        }
    }
    if (p.addend && live) {
        const uint4* ap = reinterpret_cast<const uint4*>(p.addend + row * p.addend_ld + p.addend_chan_off + col);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const uint4 av = pre ? pre->a[g] : __ldg(ap + g);
            const uint32_t w[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 f = unpack_bf16x2(w[q]);
                v[g * 8 + q * 2] += f.x; v[g * 8 + q * 2 + 1] += f.y;
            }
        }
    }
    if (p.act) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f) + e.slope_eff * fminf(v[j], 0.f);
    }
    if (!live) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
}

// Direct epilogue (fp32 outputs, or tiles narrower than one 64-channel swizzle row): rows are stored straight from registers.
// `half` selects which of the two warps of this TMEM lane quarter takes the 32-column chunk.
__device__ __forceinline__ void epilogue_subtile(const ConvParams& p, const EpiCtx& e, long long row, int n0, uint32_t taddr, int half) {
    const bool in_range = row < p.rows;
    int img = 0;
    if (p.row_img) img = in_range ? (int)p.row_img[row] : -1;
    const bool live = in_range && img >= 0;
    const int nchunk = p.bn >> 5;
    for (int ci = (nchunk > 1 ? half : 0); ci < nchunk; ci += (nchunk > 1 ? 2 : 1)) {
        if (nchunk == 1 && half) break;
        const int c0 = ci * 32;
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
        float v[32];
        epi_math32(p, e, r, v, row, n0 + c0, live);
        if (in_range) {
            if (p.out_fp32) {
                float4* op = reinterpret_cast<float4*>((float*)p.out + row * p.out_ld + p.out_chan_off + n0 + c0);
#pragma unroll
                for (int g = 0; g < 8; ++g) op[g] = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
            } else {
                uint4* op = reinterpret_cast<uint4*>((bf16*)p.out + row * p.out_ld + p.out_chan_off + n0 + c0);
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    op[g] = make_uint4(pack_bf16x2(v[g * 8], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                                       pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
            }
        }
    }
}

// Staged epilogue of one 128-row sub-tile (bf16 output, bn % 64 == 0): the eight epilogue warps convert 64 columns at a
// time (two warps per 32-row quarter, 32 columns each) into a SWIZZLE_128B [128 rows][64 ch] shared-memory tile with
// conflict-free 16-byte stores, and one thread hands it to the TMA store engine; two staging tiles alternate.
// One-step-ahead prefetch of the epilogue's global operands (row_img entry, mask / addend rows).  They do not depend on the
// accumulator, so prefetch instructions for the NEXT 64-column chunk / sub-tile / tile are issued while the current one is
// converted (no registers held: the real loads then hit L2); the first ones of a kernel go out before the wait for the first
// accumulator.  Measured need (scripts/prof_epi.py): with the loads
// issued inside the chunk, a ReLU mask costs the 64 -> 64 full-resolution data gradient 93 us on top of 190, bias + ReLU + ring 34 us.
struct EpiNext { bool on; };      // (no payload: the next step's operands are pulled into L2 / L1 by prefetch instructions, not registers)

__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

// prefetch the epilogue operands of (row, 32-column chunk `col`): 64 bytes of mask / addend per thread, one row_img entry
__device__ __forceinline__ void epi_fetch_next(const ConvParams& p, long long row, int col) {
    if (row >= p.rows) return;
    if (p.row_img) prefetch_l2(p.row_img + row);
    if (p.mask) prefetch_l2(p.mask + row * p.mask_ld + p.mask_chan_off + col);
    if (p.addend) prefetch_l2(p.addend + row * p.addend_ld + p.addend_chan_off + col);
}

__device__ __forceinline__ void epilogue_subtile_staged(const ConvParams& p, const EpiCtx& e, const CUtensorMap* tmOut, long long row0, int r_in_tile,
                                                        int n0, uint32_t taddr, int half, uint8_t* stg, uint32_t& stg_iter, bool store_thread,
                                                        EpiNext* carry = nullptr, long long next_row0 = -1, int next_n0 = 0) {
    const long long row = row0 + r_in_tile;
    const bool in_range = row < p.rows;
    int img = 0;
    if (p.row_img) img = in_range ? (int)p.row_img[row] : -1;
    const bool live = in_range && img >= 0;
    for (int c0 = 0; c0 < p.bn; c0 += 64) {
        uint8_t* buf = stg + (p.nstg == 2 ? (stg_iter & 1) : 0) * (kBM * 128);
        const int cc = c0 + half * 32;
        uint32_t r[32];
        tmem_ld32(taddr + cc, r);                         // TMEM read overlaps the wait for the staging tile
        EpiPre pre;
        epi_prefetch(p, pre, row, n0 + cc, in_range);     // ... and so do the mask / addend rows of this chunk (ring rows included:
                                                          // no dependence on the row_img load); L2 hits when the step before prefetched them
        if (carry != nullptr) {
            // the step after this one: next chunk of this sub-tile (same row), else the first chunk of the next sub-tile / tile
            if (c0 + 64 < p.bn) epi_fetch_next(p, row, n0 + cc + 64);
            else if (next_row0 >= 0) epi_fetch_next(p, next_row0 + r_in_tile, next_n0 + half * 32);
        }
        if (store_thread) { if (p.nstg == 2) bulk_wait_group_read<1>(); else bulk_wait_group_read<0>(); }   // buffer drained by its last TMA store
        named_bar_sync(1, 256);
        tmem_ld_wait();
        float v[32];
        epi_math32(p, e, r, v, row, n0 + cc, live, &pre);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int chunk = (half * 4 + g) ^ (r_in_tile & 7);           // 128-byte swizzle: 16-byte chunk index XOR row % 8
            *reinterpret_cast<uint4*>(buf + r_in_tile * 128 + chunk * 16) =
                make_uint4(pack_bf16x2(v[g * 8], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                           pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
        }
        fence_proxy_async();
        named_bar_sync(1, 256);
        if (store_thread) {
            tma_store_2d(tmOut, buf, p.out_chan_off + n0 + c0, (int)row0);
            bulk_commit_group();
        }
        if (p.stats_part) {
            // column sums of the staged tile (the values the consumer will read back, i.e. after the bf16 rounding; dead rows
            // were stored as zeros).  Warp w takes columns 8w..8w+7; lane = (row quarter << 3) | column; the quarters walk their
            // 32 rows from different offsets so that the four rows read together sit in different 16-byte swizzle chunks.
            const int et = (int)threadIdx.x - 64;
            const int col = (et >> 5) * 8 + (et & 7), rq = (et >> 3) & 3;
            const int rb = p.rows_per_img - (int)((unsigned)row0 % (unsigned)p.rows_per_img);      // tile-local rows >= rb belong to the next image
            float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll 8
            for (int i = 0; i < 32; ++i) {
                const int r = rq * 32 + ((i + 2 * rq) & 31);
                const unsigned short u = *reinterpret_cast<const unsigned short*>(buf + r * 128 + (((col >> 3) ^ (r & 7)) << 4) + ((col & 7) << 1));
                const float f = __uint_as_float((uint32_t)u << 16);
                if (r < rb) { s0 += f; q0 = fmaf(f, f, q0); } else { s1 += f; q1 = fmaf(f, f, q1); }
            }
            // quarters combined in a fixed order: (q0 + q1) + (q2 + q3)
            s0 += __shfl_xor_sync(0xffffffffu, s0, 8); q0 += __shfl_xor_sync(0xffffffffu, q0, 8);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 8); q1 += __shfl_xor_sync(0xffffffffu, q1, 8);
            s0 += __shfl_xor_sync(0xffffffffu, s0, 16); q0 += __shfl_xor_sync(0xffffffffu, q0, 16);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 16); q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
            if (rq == 0 && row0 < p.rows) {
                const int ch = n0 + c0 + col;
                *reinterpret_cast<float2*>(p.stats_part + ((row0 >> 7) * p.n_out + ch) * 2) = make_float2(s0, q0);
                if (rb < kBM) {
                    const long long img1 = (unsigned)row0 / (unsigned)p.rows_per_img + 1;
                    *reinterpret_cast<float2*>(p.stats_edge + (img1 * p.n_out + ch) * 2) = make_float2(s1, q1);
                }
            }
        }
        ++stg_iter;
    }
}

// Epilogue of the horizontal-tap mode (fp32, bn = 32): all sub-tiles of the tile are copied from TMEM to a shared-memory
// strip (row stride 21..: odd, conflict-free), the accumulator is released, and every thread then reduces the taps of its
// output rows, adds the bias, applies tanh and writes fp32 NCHW planes (consecutive rows = consecutive x: coalesced).
template <int MT>
__device__ __forceinline__ void epilogue_tapsum(const ConvParams& p, long long row0, uint32_t tmem_acc, int quarter, int half, float* strip, uint64_t* tempty_bar) {
    const int ncol = p.tap_nshift * p.tap_nco;           // <= 32
    const int ld = ncol | 1;                             // odd row stride
    const int lane = threadIdx.x & 31;
    named_bar_sync(2, 256);                              // the strip of the previous tile has been consumed
#pragma unroll 1
    for (int m = half; m < MT; m += 2) {
        uint32_t r[32];
        tmem_ld32(tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(m * p.bn), r);
        tmem_ld_wait();
        float* dst = strip + (m * kBM + quarter * 32 + lane) * ld;
#pragma unroll
        for (int j = 0; j < 32; ++j) if (j < ncol) dst[j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    named_bar_sync(2, 256);
    if (lane == 0) mbar_arrive(tempty_bar);              // the MMAs of the next tile may overwrite the accumulator
    const int halo = (p.tap_nshift - 1) >> 1;
    const int et = (int)threadIdx.x - 64;
    const long long hw = (long long)p.tap_H * p.tap_W;
    const int img_rows = p.tap_hp * p.tap_wp;            // frame rows per image (the host checks rows < 2^31: 32-bit index math)
    for (int r = halo + et; r < MT * kBM - halo; r += 256) {
        const long long q = row0 + r;
        if (q < 0 || q >= p.rows) continue;
        const int qi = (int)q;
        const int n = qi / img_rows;
        const int rem = qi - n * img_rows;
        const int Y = rem / p.tap_wp, X = rem - Y * p.tap_wp;
        const int y = Y - p.tap_oy, x = X - p.tap_ox;
        if (y < 0 || y >= p.tap_H || x < 0 || x >= p.tap_W) continue;
        for (int co = 0; co < p.tap_nco; ++co) {
            float acc = p.bias ? __ldg(p.bias + co) : 0.f;
            for (int j = 0; j < p.tap_nshift; ++j) acc += strip[(r + j - halo) * ld + j * p.tap_nco + co];
            if (p.tap_act == 3) acc = tanhf(acc);
            if (p.tap_scale) acc *= __ldg(p.tap_scale + co);
            float* o = p.tap_out + ((long long)n * p.tap_nco + co) * hw + (long long)y * p.tap_W + x;
            // accumulate: every output element has exactly one producer in this launch, so a reduction instruction gives the same
            // bits as load-add-store without putting a global-load round trip on every tile's critical path
            if (p.tap_accumulate) atomicAdd(o, acc); else *o = acc;
        }
    }
}

// Epilogue role of the persistent conv kernels: 8 warps (two per TMEM lane quarter) drain the accumulator of every tile this
// CTA owns - staged TMA-store epilogue, direct stores, or the horizontal tap reduction - and hand the buffer back to the MMA warp.
template <int MT>
__device__ __forceinline__ void conv_epilogue_loop(const ConvParams& p, const CUtensorMap* tmOut, uint32_t tmem_base, uint64_t* tfull, uint64_t* tempty,
                                                   uint8_t* stg, const float* sbias, long long total_tiles) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t acc_cols = (uint32_t)(MT * p.bn);
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;             // which of the two warps of the quarter
    const int r_in_tile = quarter * 32 + lane;
    EpiCtx e;
    e.slope_eff = p.act == 1 ? 0.f : (p.act == 2 ? p.slope : 1.f);
    e.sbias = p.bias ? sbias : nullptr;
    int acc = 0; uint32_t acc_phase = 0;
    uint32_t stg_iter = 0;
    const bool store_thread = warp == 2 && lane == 0;
    EpiNext carry;
    carry.on = true;
    const bool use_carry = p.tma_store && !p.tap_out && (p.mask != nullptr || p.addend != nullptr || p.row_img != nullptr);
    if (use_carry && (long long)blockIdx.x < total_tiles)
        epi_fetch_next(p, ((long long)blockIdx.x / p.n_tiles) * p.tile_stride + p.row_bias + r_in_tile,
                       (int)((long long)blockIdx.x % p.n_tiles) * p.bn + half * 32);
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n0 = (int)(tile % p.n_tiles) * p.bn;
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        if (p.tap_out) {
            epilogue_tapsum<MT>(p, (tile / p.n_tiles) * p.tile_stride + p.row_bias, tmem_base + (uint32_t)acc * acc_cols, quarter, half,
                                reinterpret_cast<float*>(stg), &tempty[acc]);
            if (++acc == p.nbuf) { acc = 0; acc_phase ^= 1; }
            continue;
        }
#pragma unroll 1
        for (int m = 0; m < MT; ++m) {
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * acc_cols + (uint32_t)(m * p.bn);
            const long long srow0 = (tile / p.n_tiles) * p.tile_stride + p.row_bias + m * kBM;
            if (p.tma_store) {
                long long nrow0 = -1; int nn0 = n0;
                if (use_carry) {
                    if (m + 1 < MT) nrow0 = srow0 + kBM;
                    else if (tile + gridDim.x < total_tiles) {
                        const long long t2 = tile + gridDim.x;
                        nrow0 = (t2 / p.n_tiles) * p.tile_stride + p.row_bias;
                        nn0 = (int)(t2 % p.n_tiles) * p.bn;
                    }
                }
                epilogue_subtile_staged(p, e, tmOut, srow0, r_in_tile, n0, taddr, half, stg, stg_iter, store_thread, use_carry ? &carry : nullptr, nrow0, nn0);
            } else epilogue_subtile(p, e, srow0 + r_in_tile, n0, taddr, half);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        if (++acc == p.nbuf) { acc = 0; acc_phase ^= 1; }
    }
}

template <int MT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
    irc::pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int tile_rows = kBM * MT;
    constexpr int a_box = tile_rows > 256 ? 256 : tile_rows;
    const int stage_a = tile_rows * 128;
    const int stage_b = p.bn * 128;
    const int stage_bytes = stage_a + stage_b;
    const int S = p.stages;
    uint64_t* bars = (uint64_t*)(smem + (size_t)S * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint64_t* tempty = bars + 2 * S + 2;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 4);
    uint8_t* stg = smem + (size_t)S * stage_bytes + 1024;      // 2 x 16 KB staging tiles of the TMA-store epilogue

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m_tiles = (p.rows + p.tile_stride - 1) / p.tile_stride;
    const uint32_t acc_cols = (uint32_t)(MT * p.bn);      // TMEM columns of one accumulator buffer
    const long long total_tiles = m_tiles * p.n_tiles;
    const int num_kb = p.ntaps * p.k_chunks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (p.tma_store) tma_prefetch_desc(&tmOut);
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
        fence_barrier_init();
    }
    __shared__ __align__(16) float sbias[512];
    if (p.bias) for (int i = threadIdx.x; i < p.n_out; i += blockDim.x) sbias[i] = p.bias[i];
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        {
            int stage = 0; uint32_t phase = 0;
            for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const long long row0 = (tile / p.n_tiles) * p.tile_stride + p.row_bias;
                const int n0 = (int)(tile % p.n_tiles) * p.bn;
                for (int t = 0; t < p.ntaps; ++t) {
                    const long long arow = row0 + p.taps[t];
                    for (int kc = 0; kc < p.k_chunks; ++kc) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx(&full[stage], stage_bytes);
                            uint8_t* sa = smem + (size_t)stage * stage_bytes;
                            for (int r = 0; r < tile_rows; r += a_box)      // TMA boxes are at most 256 rows
                                tma_load_2d(sa + r * 128, &tmA, &full[stage], p.a_chan_off + kc * kBK, (int)arow + r);
                            tma_load_2d(sa + stage_a, &tmB, &full[stage], (t * p.k_chunks + kc) * kBK, n0);
                        }
                        __syncwarp();
                        if (++stage == S) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        {
            const uint32_t idesc = umma_idesc_bf16(kBM, p.bn, 0, 0);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_cols;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t bdesc = umma_desc_sw128(sa + stage_a, 16);
                    if (elect_one()) {
                        // consecutive MMAs go to different accumulators: a chain of dependent accumulations into one
                        // TMEM tile costs ~130 cycles per MMA regardless of N (measured, scripts/prof_mma.py), so narrow
                        // tiles need 2-4 independent chains in flight to reach the tensor-pipe rate
                        const uint64_t adesc = umma_desc_sw128(sa, 16);
                        const uint32_t accum = kb != 0;
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k) {
                            if (k >= p.k16) break;              // operand columns past k_live are structural zeros (args.k_live)
#pragma unroll
                            for (int m = 0; m < MT; ++m)
                                umma_bf16(d_tmem + (uint32_t)(m * p.bn), adesc + (uint64_t)(m * (kBM * 128 / 16) + k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                          k == 0 ? accum : 1u);
                        }
                        umma_commit(&empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit(&tfull[acc]);
                __syncwarp();
                if (++acc == p.nbuf) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        conv_epilogue_loop<MT>(p, &tmOut, tmem_base, tfull, tempty, stg, sbias, total_tiles);
    }

    if (p.tma_store && warp == 2 && lane == 0) bulk_wait_group<0>();      // all output tiles have left shared memory
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}


// ------------------------------------------------------------------------------------
// conv_gemm on CTA pairs: tcgen05.mma.cta_group::2, tile = 256 rows x bn columns per pair
// ------------------------------------------------------------------------------------
// Same roles and the same epilogue as conv_gemm_kernel<1>; per pipeline stage each CTA stages its own 128 rows of A and HALF of
// the weight tile (bn / 2 rows of B), so a stage costs 16 KB + bn * 64 B instead of 16 KB + bn * 128 B: a third less TMA / L2
// traffic at bn = 256 and room for more stages.  Barrier protocol (all barriers live at the same offset in both CTAs):
//   full[s]   (leader's; count 1 + tx)   both CTAs' TMA loads report their bytes to the LEADER's barrier
//   empty[s]  (one per CTA; count 1)     tcgen05.commit multicast to both CTAs when the MMAs that read stage s are done
//   tfull[a]  (one per CTA; count 1)     commit multicast: accumulator a complete (each CTA drains its own 128 TMEM lanes)
//   tempty[a] (leader's; count 16)       the 8 epilogue warps of BOTH CTAs arrive (the peer's remotely)
__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
    irc::pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int stage_a = kBM * 128;
    const int half_bn = p.bn >> 1;
    const int stage_b = half_bn * 128;
    const int stage_bytes = stage_a + stage_b;
    const int S = p.stages;
    uint64_t* bars = (uint64_t*)(smem + (size_t)S * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint64_t* tempty = bars + 2 * S + 2;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 4);
    uint8_t* stg = smem + (size_t)S * stage_bytes + 1024;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const long long m_tiles = (p.rows + 2 * kBM - 1) / (2 * kBM);
    const long long total_tiles = m_tiles * p.n_tiles;
    const uint32_t acc_cols = (uint32_t)p.bn;
    const int num_kb = p.ntaps * p.k_chunks;
    const long long pair0 = blockIdx.x >> 1, pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (p.tma_store) tma_prefetch_desc(&tmOut);
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 16); }
        fence_barrier_init();
    }
    __shared__ __align__(16) float sbias[512];
    if (p.bias) for (int i = threadIdx.x; i < p.n_out; i += blockDim.x) sbias[i] = p.bias[i];
    if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // both CTAs' barriers are initialised before any remote arrival / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        int stage = 0; uint32_t phase = 0;
        for (long long tile = pair0; tile < total_tiles; tile += pairs) {
            const long long row0 = (tile / p.n_tiles) * (2 * kBM) + (long long)rank * kBM;
            const int n0 = (int)(tile % p.n_tiles) * p.bn + (int)rank * half_bn;
            for (int t = 0; t < p.ntaps; ++t) {
                const long long arow = row0 + p.taps[t];
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (elect_one()) {
                        const uint32_t fb = mapa_u32(smem_u32(&full[stage]), 0);
                        if (leader) mbar_expect_tx(&full[stage], 2 * stage_bytes);
                        uint8_t* sa = smem + (size_t)stage * stage_bytes;
                        tma_load_2d_pair(sa, &tmA, fb, p.a_chan_off + kc * kBK, (int)arow);
                        tma_load_2d_pair(sa + stage_a, &tmB, fb, (t * p.k_chunks + kc) * kBK, n0);
                    }
                    __syncwarp();
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            const uint32_t idesc = umma_idesc_bf16(2 * kBM, p.bn, 0, 0);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (long long tile = pair0; tile < total_tiles; tile += pairs) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_cols;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t adesc = umma_desc_sw128(sa, 16);
                    const uint64_t bdesc = umma_desc_sw128(sa + stage_a, 16);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k)
                            umma_bf16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_commit_pair(&empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit_pair(&tfull[acc]);
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (both CTAs: own 128 accumulator rows) =====================
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int r_in_tile = quarter * 32 + lane;
        EpiCtx e;
        e.slope_eff = p.act == 1 ? 0.f : (p.act == 2 ? p.slope : 1.f);
        e.sbias = p.bias ? sbias : nullptr;
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t stg_iter = 0;
        const bool store_thread = warp == 2 && lane == 0;
        const uint32_t tempty_leader[2] = {mapa_u32(smem_u32(&tempty[0]), 0), mapa_u32(smem_u32(&tempty[1]), 0)};
        EpiNext carry;
        carry.on = true;
        const bool use_carry = p.tma_store && (p.mask != nullptr || p.addend != nullptr || p.row_img != nullptr);
        if (use_carry && pair0 < total_tiles)
            epi_fetch_next(p, (pair0 / p.n_tiles) * (2 * kBM) + (long long)rank * kBM + r_in_tile, (int)(pair0 % p.n_tiles) * p.bn + half * 32);
        for (long long tile = pair0; tile < total_tiles; tile += pairs) {
            const int n0 = (int)(tile % p.n_tiles) * p.bn;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * acc_cols;
            const long long srow0 = (tile / p.n_tiles) * (2 * kBM) + (long long)rank * kBM;
            if (p.tma_store) {
                long long nrow0 = -1; int nn0 = n0;
                if (use_carry && tile + pairs < total_tiles) {
                    const long long t2 = tile + pairs;
                    nrow0 = (t2 / p.n_tiles) * (2 * kBM) + (long long)rank * kBM;
                    nn0 = (int)(t2 % p.n_tiles) * p.bn;
                }
                epilogue_subtile_staged(p, e, &tmOut, srow0, r_in_tile, n0, taddr, half, stg, stg_iter, store_thread, use_carry ? &carry : nullptr, nrow0, nn0);
            } else epilogue_subtile(p, e, srow0 + r_in_tile, n0, taddr, half);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(tempty_leader[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    if (p.tma_store && warp == 2 && lane == 0) bulk_wait_group<0>();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // the peer's shared memory and TMEM are read by the leader's MMAs until its last commit
    if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

// ------------------------------------------------------------------------------------
// conv_gemm with A-tile reuse across horizontal taps ("tap runs")
// ------------------------------------------------------------------------------------
// Taps whose row shifts are consecutive integers (the kw taps of one kernel row) read the same pixels shifted by one
// frame row each.  The plain kernel fetches the A tile once per tap: with Cout <= 128 that is 9 x 16 KB of L2 -> shared-memory
// traffic per 128 x 64-channel block for only 64..128 MMA columns, and the layer runs at the L2 slice bandwidth
// (~6300 B/clk chip-wide), not at the tensor rate (V.2 forward: 2.96 GB in 242 us = 12.2 TB/s, 0.47 of the bf16 peak).
// Here ONE box of 128 * MT + 8 rows is staged per (run, channel chunk) and the L taps of the run issue their MMAs from it
// through descriptors whose start address is advanced by 128 bytes (one pixel row) per tap: a third of the A traffic for
// 3 x 3 kernels.  The weight tiles travel through their own ring (one stage per tap and channel chunk).
struct RunParams {
    int nruns;
    int run_first[IRC_MAX_TAPS];   // smallest row shift of the run
    int run_len[IRC_MAX_TAPS];
    int run_tap[IRC_MAX_TAPS][8];  // original tap index (-> weight K offset) of each member, ascending shift
    int sa_stages, sb_stages, a_stage_bytes, base_off_mode;
};

// Measured on B200 (scripts/try_reuse.py): the 128-byte swizzle phase is taken from the absolute shared-memory
// address, so a descriptor whose start is advanced by whole 128-byte rows needs NO base-offset field (setting
// base_offset = (addr >> 7) & 7 gives wrong products).
__device__ __forceinline__ uint64_t umma_desc_sw128_shifted(uint32_t smem_addr, int base_off_mode) {
    uint64_t d = umma_desc_sw128(smem_addr, 16);
    if (base_off_mode == 1) d |= (uint64_t)((smem_addr >> 7) & 7) << 49;
    return d;
}

template <int MT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_runs_kernel(const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA8,
                      const __grid_constant__ CUtensorMap tmB, const ConvParams p, const RunParams rp) {
    irc::pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int tile_rows = kBM * MT;
    constexpr int a_box = tile_rows > 256 ? 256 : tile_rows;
    const int SA = rp.sa_stages, SB = rp.sb_stages;
    const int stage_b = p.bn * 128;
    uint8_t* smA = smem;
    uint8_t* smB = smem + (size_t)SA * rp.a_stage_bytes;
    uint64_t* bars = (uint64_t*)(smB + (size_t)SB * stage_b);
    uint64_t* fullA = bars;
    uint64_t* emptyA = fullA + SA;
    uint64_t* fullB = emptyA + SA;
    uint64_t* emptyB = fullB + SB;
    uint64_t* tfull = emptyB + SB;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
    uint8_t* stg = (uint8_t*)bars + 1024;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m_tiles = (p.rows + p.tile_stride - 1) / p.tile_stride;
    const long long total_tiles = m_tiles * p.n_tiles;
    const uint32_t acc_cols = (uint32_t)(MT * p.bn);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmA8);
        tma_prefetch_desc(&tmB);
        if (p.tma_store) tma_prefetch_desc(&tmOut);
        for (int s = 0; s < SA; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
        for (int s = 0; s < SB; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
        fence_barrier_init();
    }
    __shared__ __align__(16) float sbias[512];
    if (p.bias) for (int i = threadIdx.x; i < p.n_out; i += blockDim.x) sbias[i] = p.bias[i];
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
        const uint32_t a_bytes = (uint32_t)(tile_rows + 8) * 128u;
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const long long row0 = (tile / p.n_tiles) * p.tile_stride + p.row_bias;
            const int n0 = (int)(tile % p.n_tiles) * p.bn;
            for (int r = 0; r < rp.nruns; ++r) {
                const int arow = (int)(row0 + rp.run_first[r]);
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait(&emptyA[sa], pa ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&fullA[sa], a_bytes);
                        uint8_t* dst = smA + (size_t)sa * rp.a_stage_bytes;
                        for (int q = 0; q < tile_rows; q += a_box)
                            tma_load_2d(dst + q * 128, &tmA, &fullA[sa], p.a_chan_off + kc * kBK, arow + q);
                        tma_load_2d(dst + tile_rows * 128, &tmA8, &fullA[sa], p.a_chan_off + kc * kBK, arow + tile_rows);   // the rows the shifted taps reach into
                    }
                    __syncwarp();
                    if (++sa == SA) { sa = 0; pa ^= 1; }
                    for (int j = 0; j < rp.run_len[r]; ++j) {
                        mbar_wait(&emptyB[sb], pb ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx(&fullB[sb], stage_b);
                            tma_load_2d(smB + (size_t)sb * stage_b, &tmB, &fullB[sb], (rp.run_tap[r][j] * p.k_chunks + kc) * kBK, n0);
                        }
                        __syncwarp();
                        if (++sb == SB) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = umma_idesc_bf16(kBM, p.bn, 0, 0);
        int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            mbar_wait(&tempty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_cols;
            uint32_t started = 0;
            for (int r = 0; r < rp.nruns; ++r) {
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait(&fullA[sa], pa);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smA + (size_t)sa * rp.a_stage_bytes);
                    for (int j = 0; j < rp.run_len[r]; ++j) {
                        mbar_wait(&fullB[sb], pb);
                        tc_fence_after();
                        const uint64_t adesc = umma_desc_sw128_shifted(a_addr + (uint32_t)j * 128u, rp.base_off_mode);
                        const uint64_t bdesc = umma_desc_sw128(smem_u32(smB + (size_t)sb * stage_b), 16);
                        if (elect_one()) {
                            // consecutive MMAs go to different accumulators (independent chains, see conv_gemm_kernel)
#pragma unroll
                            for (int k = 0; k < kBK / 16; ++k) {
#pragma unroll
                                for (int m = 0; m < MT; ++m)
                                    umma_bf16(d_tmem + (uint32_t)(m * p.bn), adesc + (uint64_t)(m * (kBM * 128 / 16) + k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                              k == 0 ? started : 1u);
                            }
                            umma_commit(&emptyB[sb]);
                        }
                        __syncwarp();
                        started = 1;
                        if (++sb == SB) { sb = 0; pb ^= 1; }
                    }
                    if (elect_one()) umma_commit(&emptyA[sa]);
                    __syncwarp();
                    if (++sa == SA) { sa = 0; pa ^= 1; }
                }
            }
            if (elect_one()) umma_commit(&tfull[acc]);
            __syncwarp();
            if (++acc == p.nbuf) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        conv_epilogue_loop<MT>(p, &tmOut, tmem_base, tfull, tempty, stg, sbias, total_tiles);
    }

    if (p.tma_store && warp == 2 && lane == 0) bulk_wait_group<0>();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------
// conv_gemm with the kw taps of a kernel row packed along N ("packed taps", 64 output channels, 3 x 3 kernels)
// ------------------------------------------------------------------------------------
// In SS mode every tcgen05.mma re-reads its 4 KB A slab from shared memory, so an M128 x N64 x K16 MMA costs ~94 cycles
// against a 32-cycle tensor floor (profiles/r3/ncu_narrow_v2_*): the N = 64 layers sit at half of the sustained peak no matter
// how the operands reach shared memory.  Here the three taps of one kernel row share ONE unshifted A slab: their weight tiles
// are stacked along N (3 x 64 = 192 accumulator columns: P_j[q] = sum_k A[q + mid][k] W_j[k], j = 0..2) and the epilogue adds
// the three column groups shifted by one row each: out[q] = P_0[q - 1] + P_1[q] + P_2[q + 1].  Rows = TMEM lanes, so the shift is
// a warp shuffle; the rows next to a lane-quarter boundary travel through a small shared-memory exchange buffer and tiles
// overlap by two rows (126 outputs per 128-row tile).  A third of the A reads per FLOP and a third of the MMA instructions.
constexpr int kPackRows = kBM - 2;      // outputs per tile

struct PackParams {
    int nruns;
    int run_mid[IRC_MAX_TAPS / 3];     // row shift of the middle tap of each run
    int run_tap[IRC_MAX_TAPS / 3][3];  // original tap indices (-> weight K offset), ascending shift
};

__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_pack_kernel(const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p,
                      const PackParams pk) {
    irc::pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int stage_a = kBM * 128;             // 128 rows x 64 channels
    constexpr int tap_b = 64 * 128;                // one tap's weight tile: 64 output channels x 64 channels
    constexpr int stage_bytes = stage_a + 3 * tap_b;
    constexpr uint32_t acc_cols = 192;
    const int S = p.stages;
    uint64_t* bars = (uint64_t*)(smem + (size_t)S * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint64_t* tempty = bars + 2 * S + 2;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 4);
    float* xbuf = reinterpret_cast<float*>((uint8_t*)bars + 256);              // [2 groups][2 passes][2 dirs][4 quarters][32] boundary rows (4 KB)
    uint8_t* stg = smem + (size_t)S * stage_bytes + 8192;                      // 2 x 16 KB staging tiles

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long total_tiles = (p.rows + kPackRows - 1) / kPackRows;
    const int num_kb = pk.nruns * p.k_chunks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmOut);
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        fence_barrier_init();
    }
    __shared__ __align__(16) float sbias[512];
    if (p.bias) for (int i = threadIdx.x; i < p.n_out; i += blockDim.x) sbias[i] = p.bias[i];
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0; uint32_t phase = 0;
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const long long row0 = tile * kPackRows - 1;
            for (int r = 0; r < pk.nruns; ++r) {
                const int arow = (int)(row0 + pk.run_mid[r]);
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&full[stage], stage_bytes);
                        uint8_t* sa = smem + (size_t)stage * stage_bytes;
                        tma_load_2d(sa, &tmA, &full[stage], p.a_chan_off + kc * kBK, arow);
#pragma unroll
                        for (int j = 0; j < 3; ++j)
                            tma_load_2d(sa + stage_a + j * tap_b, &tmB, &full[stage], (pk.run_tap[r][j] * p.k_chunks + kc) * kBK, 0);
                    }
                    __syncwarp();
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = umma_idesc_bf16(kBM, 192, 0, 0);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            mbar_wait(&tempty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_cols;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint64_t adesc = umma_desc_sw128(sa, 16);
                const uint64_t bdesc = umma_desc_sw128(sa + stage_a, 16);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k)
                        umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(&empty[stage]);
                }
                __syncwarp();
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(&tfull[acc]);
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ===================== epilogue: shifted sum of the three column groups, then the usual per-row math =====================
        // Two groups of four warps; group g owns accumulator buffer g, i.e. every second tile of this CTA, and walks its 64 output
        // channels in two passes of 32 (3 x 32 accumulator columns in registers per pass).  While one group is in its shuffle /
        // pack / store phase the other one is reading TMEM, so the 64 B/clk TMEM read path (1536 cycles per tile) stays busy:
        // with all eight warps on one tile the two phases alternated and a tile took ~4500 cycles against a 1900-cycle main
        // loop at 64 input channels (profiles/r3/narrow_v2_with_packed.txt).
        const int quarter = warp & 3;                 // TMEM lane quarter
        const int grp = (warp - 2) >> 2;              // epilogue group == accumulator buffer
        const int r_in_tile = quarter * 32 + lane;
        const uint32_t bar_id = 1 + grp;
        EpiCtx e;
        e.slope_eff = p.act == 1 ? 0.f : (p.act == 2 ? p.slope : 1.f);
        e.sbias = p.bias ? sbias : nullptr;
        uint32_t acc_phase = 0;
        const bool st_thread = (warp == 2 + 4 * grp) && lane == 0;
        uint8_t* buf = stg + grp * (kBM * 128);
        float* xg = xbuf + grp * 512;                 // [pass][dir][quarter][32]
        for (long long tile = blockIdx.x + (long long)grp * gridDim.x; tile < total_tiles; tile += 2LL * gridDim.x) {
            const long long row0 = tile * kPackRows - 1;
            const long long row = row0 + r_in_tile;
            const bool valid = r_in_tile >= 1 && r_in_tile <= kPackRows && row < p.rows;
            int img = 0;
            if (p.row_img) img = valid ? (int)p.row_img[row] : -1;      // issued before the accumulator wait: latency hidden
            const bool live = valid && img >= 0;
            mbar_wait(&tfull[grp], acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)grp * acc_cols + (uint32_t)(half * 32);
                uint32_t r0[32], r1[32], r2[32];
                tmem_ld32(taddr, r0);
                tmem_ld32(taddr + 64, r1);
                tmem_ld32(taddr + 128, r2);
                tmem_ld_wait();
                if (half == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[grp]);          // this warp holds the last of its accumulator columns
                }
                float* xp = xg + half * 256;
                if (lane == 31) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) xp[quarter * 32 + j] = __uint_as_float(r0[j]);          // P_0 row read by quarter + 1
                }
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) xp[128 + quarter * 32 + j] = __uint_as_float(r2[j]);    // P_2 row read by quarter - 1
                }
                if (half == 0 && st_thread) bulk_wait_group_read<0>();      // this group's staging tile of its previous tile has left
                named_bar_sync(bar_id, 128);
                uint32_t c[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float up = __shfl_up_sync(0xffffffffu, __uint_as_float(r0[j]), 1);
                    float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[j]), 1);
                    if (lane == 0 && quarter > 0) up = xp[(quarter - 1) * 32 + j];
                    if (lane == 31 && quarter < 3) dn = xp[128 + (quarter + 1) * 32 + j];
                    c[j] = __float_as_uint(up + __uint_as_float(r1[j]) + dn);
                }
                float v[32];
                epi_math32(p, e, c, v, row, half * 32, live);
                if (r_in_tile >= 1 && r_in_tile <= kPackRows) {
                    const int sr = r_in_tile - 1;                      // staging row: the 126 outputs of the tile start at row 0
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int chunk = (half * 4 + g) ^ (sr & 7);
                        *reinterpret_cast<uint4*>(buf + sr * 128 + chunk * 16) =
                            make_uint4(pack_bf16x2(v[g * 8], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                                       pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
                    }
                }
            }
            fence_proxy_async();
            named_bar_sync(bar_id, 128);
            if (st_thread) {
                tma_store_2d(&tmOut, buf, p.out_chan_off, (int)(row0 + 1));
                bulk_commit_group();
            }
            acc_phase ^= 1;
        }
        if (st_thread) bulk_wait_group<0>();
    }

    if (warp == 2 && lane == 0) bulk_wait_group<0>();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------
// tn_gemm (weight gradients)
// ------------------------------------------------------------------------------------
struct TnParams {
    long long k_rows;        // reduction length (rows)
    int m, n;                // valid output extents
    int bn;                  // tile N (multiple of 64)
    int m_tiles, n_tiles, ntaps, splits, stages;
    int groups;              // tap groups: one CTA accumulates TPC taps into TPC independent TMEM tiles
    int merge_taps;          // bn == 64: the taps of a group go through one MMA of N = 64 * taps (IRC_TN_MERGE=0 turns it off)
    int tmem_cols;
    int a_chan_off, b_chan_off;
    int a_shift[IRC_MAX_TAPS];
    int b_shift[IRC_MAX_TAPS];
    float* out;
    long long out_tap_stride, out_m_stride, out_n_stride, out_split_stride;
};

// TPC taps per CTA: the A tile (output-gradient rows) is staged once per k-block and used by TPC MMAs chains that
// accumulate into TPC different TMEM tiles, issued round-robin so that consecutive MMAs never depend on each other
// (a single chain of narrow MMAs is latency-bound at ~130 cycles each, scripts/prof_mma.py).
template <int TPC>
__global__ void __launch_bounds__(kThreads, 1)
tn_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TnParams p) {
    irc::pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int stage_a = kBK * 128 * 2;             // 2 groups of 64 channels x 64 rows
    const int groups_b = p.bn / 64;
    const int tile_b = kBK * 128 * groups_b;       // one tap's B tile
    const int stage_bytes = stage_a + TPC * tile_b;
    const int S = p.stages;
    uint64_t* bars = (uint64_t*)(smem + (size_t)S * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // work decomposition: blockIdx.x -> (split, tap group, m_tile, n_tile)
    int w = blockIdx.x;
    const int n_tile = w % p.n_tiles; w /= p.n_tiles;
    const int m_tile = w % p.m_tiles; w /= p.m_tiles;
    const int group = w % p.groups; w /= p.groups;
    const int split = w;
    const int tap0 = group * TPC;
    const int ntap = (p.ntaps - tap0) < TPC ? (p.ntaps - tap0) : TPC;      // active taps of this group
    const long long kb_total = (p.k_rows + kBK - 1) / kBK;
    const long long kb_per = (kb_total + p.splits - 1) / p.splits;
    const long long kb_begin = (long long)split * kb_per;
    long long kb_end = kb_begin + kb_per; if (kb_end > kb_total) kb_end = kb_total;
    const long long my_kb = kb_end > kb_begin ? kb_end - kb_begin : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        const uint32_t bytes = (uint32_t)(stage_a + ntap * tile_b);
        for (long long kb = kb_begin; kb < kb_begin + my_kb; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&full[stage], bytes);
                uint8_t* sa = smem + (size_t)stage * stage_bytes;
                const long long r0 = kb * kBK;
                for (int g = 0; g < 2; ++g)
                    tma_load_2d(sa + g * (kBK * 128), &tmA, &full[stage], p.a_chan_off + m_tile * kBM + g * 64, (int)(r0 + p.a_shift[tap0]));
                for (int t = 0; t < ntap; ++t)
                    for (int g = 0; g < groups_b; ++g)
                        tma_load_2d(sa + stage_a + t * tile_b + g * (kBK * 128), &tmB, &full[stage], p.b_chan_off + n_tile * p.bn + g * 64,
                                    (int)(r0 + p.b_shift[tap0 + t]));
            }
            __syncwarp();
            if (++stage == S) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        const uint32_t idesc = umma_idesc_bf16(kBM, p.bn, 1, 1);
        const bool merge_taps = TPC > 1 && p.bn == 64 && p.merge_taps;
        int stage = 0; uint32_t phase = 0;
        for (long long kb = 0; kb < my_kb; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
            const uint64_t adesc = umma_desc_sw128(sa, kBK * 128);
            const uint32_t accum = kb != 0;
            if (elect_one()) {
                if (merge_taps) {
                    // bn == 64: the taps' B tiles are consecutive 8 KB blocks = the 64-channel groups of ONE wider MN-major operand, and
                    // their accumulators are consecutive 64-column blocks: up to four taps go through a single MMA of N = 64 * taps.
                    // Every MMA re-reads its A slab from shared memory (~64 B/clk), so 4 x (A + 64 columns of B) costs 4 x 96 cycles,
                    // A + 256 columns 192.
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        for (int t0 = 0; t0 < ntap; t0 += 4) {
                            const int nt = ntap - t0 < 4 ? ntap - t0 : 4;
                            const uint64_t bdesc = umma_desc_sw128(sa + stage_a + t0 * tile_b, kBK * 128);
                            umma_bf16(tmem_base + (uint32_t)(t0 * 64), adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128),
                                      umma_idesc_bf16(kBM, 64 * nt, 1, 1), k == 0 ? accum : 1u);
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {   // 16 reduction rows = 2048 bytes per step
#pragma unroll
                        for (int t = 0; t < TPC; ++t) {
                            if (t < ntap) {
                                const uint64_t bdesc = umma_desc_sw128(sa + stage_a + t * tile_b, kBK * 128);
                                umma_bf16(tmem_base + (uint32_t)(t * p.bn), adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, k == 0 ? accum : 1u);
                            }
                        }
                    }
                }
                umma_commit(&empty[stage]);
            }
            __syncwarp();
            if (++stage == S) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(tfull);
        __syncwarp();
    } else {
        const int quarter = warp & 3;
        const int m = m_tile * kBM + quarter * 32 + lane;
        if (my_kb > 0) {
            mbar_wait(tfull, 0);
            tc_fence_after();
        }
        for (int t = 0; t < ntap; ++t) {
            float* obase = p.out + (long long)split * p.out_split_stride + (long long)(tap0 + t) * p.out_tap_stride + (long long)m * p.out_m_stride;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * p.bn);
            for (int c0 = 0; c0 < p.bn; c0 += 32) {
                uint32_t r[32];
                if (my_kb > 0) {
                    tmem_ld32(taddr + c0, r);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
                if (m < p.m) {
                    const int nb = n_tile * p.bn + c0;
                    if (p.out_n_stride == 1 && nb + 32 <= p.n && ((((uintptr_t)(obase + nb)) & 15) == 0)) {
                        float4* o4 = reinterpret_cast<float4*>(obase + nb);
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            o4[g] = make_float4(__uint_as_float(r[g * 4]), __uint_as_float(r[g * 4 + 1]), __uint_as_float(r[g * 4 + 2]), __uint_as_float(r[g * 4 + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int n = nb + j;
                            if (n < p.n) obase[(long long)n * p.out_n_stride] = __uint_as_float(r[j]);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// Weight-gradient GEMM on CTA pairs (cta_group::2, M = 256): CTA r of a pair stages output-gradient channels
// [256 mt + 128 r, +128) (its own A tile) and HALF of the bn input channels of every tap's B tile; each CTA ends up with its 128
// accumulator rows in its own TMEM.  Same barrier protocol as conv_gemm_pair_kernel.  Used when the layer has >= 256 output
// channels (ResNet blocks, down2, PatchGAN model.5 / model.8).
template <int TPC>
__global__ void __launch_bounds__(kThreads, 1)
tn_gemm_cta2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TnParams p) {
    irc::pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int stage_a = kBK * 128 * 2;             // 2 groups of 64 channels x 64 rows (this CTA's 128 channels)
    const int groups_h = p.bn / 128;               // 64-channel groups of this CTA's half of the B tile
    const int tile_b = kBK * 128 * groups_h;
    const int stage_bytes = stage_a + TPC * tile_b;
    const int S = p.stages;
    uint64_t* bars = (uint64_t*)(smem + (size_t)S * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    int w = blockIdx.x >> 1;
    const int n_tile = w % p.n_tiles; w /= p.n_tiles;
    const int m_tile = w % p.m_tiles; w /= p.m_tiles;          // 256-channel tiles
    const int group = w % p.groups; w /= p.groups;
    const int split = w;
    const int tap0 = group * TPC;
    const int ntap = (p.ntaps - tap0) < TPC ? (p.ntaps - tap0) : TPC;
    const long long kb_total = (p.k_rows + kBK - 1) / kBK;
    const long long kb_per = (kb_total + p.splits - 1) / p.splits;
    const long long kb_begin = (long long)split * kb_per;
    long long kb_end = kb_begin + kb_per; if (kb_end > kb_total) kb_end = kb_total;
    const long long my_kb = kb_end > kb_begin ? kb_end - kb_begin : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        const uint32_t bytes = 2u * (uint32_t)(stage_a + ntap * tile_b);
        const int a_c0 = p.a_chan_off + m_tile * 2 * kBM + (int)rank * kBM;
        const int b_c0 = p.b_chan_off + n_tile * p.bn + (int)rank * (p.bn >> 1);
        for (long long kb = kb_begin; kb < kb_begin + my_kb; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) {
                const uint32_t fb = mapa_u32(smem_u32(&full[stage]), 0);
                if (leader) mbar_expect_tx(&full[stage], bytes);
                uint8_t* sa = smem + (size_t)stage * stage_bytes;
                const long long r0 = kb * kBK;
                for (int g = 0; g < 2; ++g)
                    tma_load_2d_pair(sa + g * (kBK * 128), &tmA, fb, a_c0 + g * 64, (int)(r0 + p.a_shift[tap0]));
                for (int t = 0; t < ntap; ++t)
                    for (int g = 0; g < groups_h; ++g)
                        tma_load_2d_pair(sa + stage_a + t * tile_b + g * (kBK * 128), &tmB, fb, b_c0 + g * 64, (int)(r0 + p.b_shift[tap0 + t]));
            }
            __syncwarp();
            if (++stage == S) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        if (leader) {
            const uint32_t idesc = umma_idesc_bf16(2 * kBM, p.bn, 1, 1);
            int stage = 0; uint32_t phase = 0;
            for (long long kb = 0; kb < my_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint64_t adesc = umma_desc_sw128(sa, kBK * 128);
                const uint32_t accum = kb != 0;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
#pragma unroll
                        for (int t = 0; t < TPC; ++t) {
                            if (t < ntap) {
                                const uint64_t bdesc = umma_desc_sw128(sa + stage_a + t * tile_b, kBK * 128);
                                umma_bf16_pair(tmem_base + (uint32_t)(t * p.bn), adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, k == 0 ? accum : 1u);
                            }
                        }
                    }
                    umma_commit_pair(&empty[stage]);
                }
                __syncwarp();
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit_pair(tfull);
            __syncwarp();
        }
    } else {
        const int quarter = warp & 3;
        const int m = m_tile * 2 * kBM + (int)rank * kBM + quarter * 32 + lane;
        if (my_kb > 0) {
            mbar_wait(tfull, 0);
            tc_fence_after();
        }
        for (int t = 0; t < ntap; ++t) {
            float* obase = p.out + (long long)split * p.out_split_stride + (long long)(tap0 + t) * p.out_tap_stride + (long long)m * p.out_m_stride;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * p.bn);
            for (int c0 = 0; c0 < p.bn; c0 += 32) {
                uint32_t r[32];
                if (my_kb > 0) {
                    tmem_ld32(taddr + c0, r);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
                if (m < p.m) {
                    const int nb = n_tile * p.bn + c0;
                    if (p.out_n_stride == 1 && nb + 32 <= p.n && ((((uintptr_t)(obase + nb)) & 15) == 0)) {
                        float4* o4 = reinterpret_cast<float4*>(obase + nb);
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            o4[g] = make_float4(__uint_as_float(r[g * 4]), __uint_as_float(r[g * 4 + 1]), __uint_as_float(r[g * 4 + 2]), __uint_as_float(r[g * 4 + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int n = nb + j;
                            if (n < p.n) obase[(long long)n * p.out_n_stride] = __uint_as_float(r[j]);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair(tmem_base, p.tmem_cols);
}

// Weight gradients of layers with at most 64 output channels (up2_conv, inc, model.0): a 64-row A tile would leave half of
// every M=128 MMA idle.  Since  dW_t[m][n] = sum_q dz[q][m] x[q + s_t][n] = sum_q' dz[q' - s_t][m] x[q'][n],  the shift can be
// put on dz instead of x: the two 64-channel groups of one A tile are then loaded at the (negated) shifts of TWO taps, the
// B tile (x, unshifted) is shared by all taps, and accumulator rows 0..63 / 64..127 hold taps 2i / 2i+1.  P pairs per CTA
// accumulate into P independent TMEM tiles (same latency argument as TPC above).
template <int P>
__global__ void __launch_bounds__(kThreads, 1)
tn_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TnParams p) {
    irc::pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int tile_a = kBK * 128 * 2;              // one pair: 2 taps x 64 channels x 64 rows
    const int groups_b = p.bn / 64;
    const int tile_b = kBK * 128 * groups_b;
    const int stage_bytes = P * tile_a + tile_b;
    const int S = p.stages;
    uint64_t* bars = (uint64_t*)(smem + (size_t)S * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int w = blockIdx.x;
    const int n_tile = w % p.n_tiles; w /= p.n_tiles;
    const int group = w % p.groups; w /= p.groups;
    const int split = w;
    const int npairs = (p.ntaps + 1) / 2;
    const int pair0 = group * P;
    const int np = (npairs - pair0) < P ? (npairs - pair0) : P;
    const long long kb_total = (p.k_rows + kBK - 1) / kBK;
    const long long kb_per = (kb_total + p.splits - 1) / p.splits;
    const long long kb_begin = (long long)split * kb_per;
    long long kb_end = kb_begin + kb_per; if (kb_end > kb_total) kb_end = kb_total;
    const long long my_kb = kb_end > kb_begin ? kb_end - kb_begin : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        const uint32_t bytes = (uint32_t)(np * tile_a + tile_b);
        for (long long kb = kb_begin; kb < kb_begin + my_kb; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&full[stage], bytes);
                uint8_t* sa = smem + (size_t)stage * stage_bytes;
                const long long r0 = kb * kBK;
                for (int i = 0; i < np; ++i)
                    for (int g = 0; g < 2; ++g) {
                        // an odd tap count leaves the last half pair without a tap: load it far outside the tensor (zero fill)
                        const int tap = 2 * (pair0 + i) + g;
                        const long long row = tap < p.ntaps ? r0 - p.b_shift[tap] : -(1LL << 30);
                        tma_load_2d(sa + i * tile_a + g * (kBK * 128), &tmA, &full[stage], p.a_chan_off, (int)row);
                    }
                for (int g = 0; g < groups_b; ++g)
                    tma_load_2d(sa + P * tile_a + g * (kBK * 128), &tmB, &full[stage], p.b_chan_off + n_tile * p.bn + g * 64, (int)r0);
            }
            __syncwarp();
            if (++stage == S) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        const uint32_t idesc = umma_idesc_bf16(kBM, p.bn, 1, 1);
        int stage = 0; uint32_t phase = 0;
        for (long long kb = 0; kb < my_kb; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
            const uint64_t bdesc = umma_desc_sw128(sa + P * tile_a, kBK * 128);
            const uint32_t accum = kb != 0;
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
#pragma unroll
                    for (int i = 0; i < P; ++i) {
                        if (i < np) {
                            const uint64_t adesc = umma_desc_sw128(sa + i * tile_a, kBK * 128);
                            umma_bf16(tmem_base + (uint32_t)(i * p.bn), adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, k == 0 ? accum : 1u);
                        }
                    }
                }
                umma_commit(&empty[stage]);
            }
            __syncwarp();
            if (++stage == S) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(tfull);
        __syncwarp();
    } else {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;            // accumulator row: tap (row >> 6) of the pair, channel row & 63
        const int m = row & 63;
        if (my_kb > 0) {
            mbar_wait(tfull, 0);
            tc_fence_after();
        }
        for (int i = 0; i < np; ++i) {
            const int tap = 2 * (pair0 + i) + (row >> 6);
            float* obase = p.out + (long long)split * p.out_split_stride + (long long)tap * p.out_tap_stride + (long long)m * p.out_m_stride;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(i * p.bn);
            for (int c0 = 0; c0 < p.bn; c0 += 32) {
                uint32_t r[32];
                if (my_kb > 0) {
                    tmem_ld32(taddr + c0, r);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
                if (m < p.m && tap < p.ntaps) {
                    const int nb = n_tile * p.bn + c0;
                    if (p.out_n_stride == 1 && nb + 32 <= p.n && ((((uintptr_t)(obase + nb)) & 15) == 0)) {
                        float4* o4 = reinterpret_cast<float4*>(obase + nb);
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            o4[g] = make_float4(__uint_as_float(r[g * 4]), __uint_as_float(r[g * 4 + 1]), __uint_as_float(r[g * 4 + 2]), __uint_as_float(r[g * 4 + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int n = nb + j;
                            if (n < p.n) obase[(long long)n * p.out_n_stride] = __uint_as_float(r[j]);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// pairs per CTA of the paired weight-gradient GEMM: as many as keep >= 4 pipeline stages and fit the 512 TMEM columns
int auto_pairs(int bn, int ntaps) {
    const int npairs = (ntaps + 1) / 2;
    int best = 1;
    for (int c = 1; c <= 4; ++c) {
        if (c > npairs || c * bn > 512) break;
        const int stage = c * kBK * 128 * 2 + kBK * 128 * (bn / 64);
        if ((kMaxSmem - 2048) / stage >= 4) best = c;
    }
    return best;
}
bool tn_pair_mode(int m, int ntaps, bool zero_a_shift) { return m <= 64 && ntaps >= 2 && zero_a_shift; }

bool g_attr_conv = false, g_attr_tn = false, g_attr_runs = false;

// taps per CTA pair of the cta_group::2 weight-gradient GEMM.  Measured (profiles/gemm_breakdown_r2_tnpair*.csv): one tap per
// pair wins (ResNet blocks 0.81 -> 0.86 of the sustained peak, model.8 0.80 -> 0.84); two taps per pair halve the number of
// pairs per split, so the one-wave rule doubles the split count and the partial traffic (0.72).  IRC_TN_PAIR=2 forces two.
int tn_pair_taps(int bn, int ntaps) {
    const char* e = getenv("IRC_TN_PAIR");
    if (e && atoi(e) == 2) return (2 * bn <= 512 && ntaps >= 2) ? 2 : 1;
    return 1;
}
// pairs pay off for 256-wide tiles; with 128 input channels (down2) the single-CTA kernel with several taps per CTA is faster
bool tn_pair_shape(int m, int n, int bn, bool same_a) {
    const char* e = getenv("IRC_TN_PAIR");
    if (e && atoi(e) == 0) return false;
    return same_a && m % (2 * kBM) == 0 && bn == 256 && n % 256 == 0;
}

// taps per CTA of the weight-gradient GEMM: as many independent accumulation chains as keep >= 4 pipeline stages
// (16 KB + tpc * bn * 128 B per stage), preferring an exact divisor of the tap count
int auto_tpc(int bn, int ntaps) {
    const int cap = 320 / bn;
    const int cand[3] = {4, 3, 2};
    if (ntaps <= 1 || cap < 2) return 1;
    for (int c : cand) if (c <= cap && ntaps % c == 0) return c;
    for (int c : cand) if (c <= cap) return c;      // ragged last group
    return 1;
}

// stats[n][c] = (sum, sum of squares) of image n: its sub-tiles' slot-0 partials in sub-tile order, preceded by the edge
// partial when the image starts inside a sub-tile.  Block = 64 channels x 16 segments of the sub-tile range.
__global__ void __launch_bounds__(1024) conv_stats_finalize_kernel(const float* __restrict__ part, const float* __restrict__ edge, int rows_per_img, int n_out,
                                                                   float* __restrict__ stats) {
    irc::pdl_prologue();
    __shared__ float2 seg[16][64];
    const int n = blockIdx.x, c = blockIdx.y * 64 + (threadIdx.x & 63), sg = threadIdx.x >> 6;
    const long long first_row = (long long)n * rows_per_img, end_row = first_row + rows_per_img;
    const long long s_begin = (first_row + kBM - 1) / kBM, s_end = (end_row + kBM - 1) / kBM;      // sub-tiles whose slot 0 is image n
    const long long cnt = s_end - s_begin, per = (cnt + 15) / 16;
    long long s0 = s_begin + sg * per, s1 = s0 + per; if (s1 > s_end) s1 = s_end;
    float a = 0.f, b = 0.f;
    if (c < n_out) {
        for (long long sidx = s0; sidx < s1; ++sidx) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(part + (sidx * n_out + c) * 2));
            a += v.x; b += v.y;
        }
    }
    seg[sg][threadIdx.x & 63] = make_float2(a, b);
    __syncthreads();
    if (sg == 0 && c < n_out) {
        float x = 0.f, y = 0.f;
        if (first_row % kBM) { const float2 v = __ldg(reinterpret_cast<const float2*>(edge + ((long long)n * n_out + c) * 2)); x = v.x; y = v.y; }
        for (int g = 0; g < 16; ++g) { x += seg[g][threadIdx.x & 63].x; y += seg[g][threadIdx.x & 63].y; }
        *reinterpret_cast<float2*>(stats + ((long long)n * n_out + c) * 2) = make_float2(x, y);
    }
}

// group the taps into runs of consecutive row shifts (ascending), at most 8 long
void build_runs(const int* taps, int ntaps, RunParams& rp) {
    int order[IRC_MAX_TAPS];
    for (int i = 0; i < ntaps; ++i) order[i] = i;
    for (int i = 1; i < ntaps; ++i) { int k = order[i], j = i; while (j > 0 && taps[order[j - 1]] > taps[k]) { order[j] = order[j - 1]; --j; } order[j] = k; }
    rp.nruns = 0;
    for (int i = 0; i < ntaps;) {
        int len = 1;
        while (i + len < ntaps && len < 8 && taps[order[i + len]] == taps[order[i]] + len) ++len;
        rp.run_first[rp.nruns] = taps[order[i]];
        rp.run_len[rp.nruns] = len;
        for (int j = 0; j < len; ++j) rp.run_tap[rp.nruns][j] = order[i + j];
        ++rp.nruns;
        i += len;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------
extern "C" int irc_conv_gemm(const irc_conv_gemm_args* a, void* stream) {
    if (!a || !a->a || !a->w || !a->out) return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: null pointer");
    if (a->cin <= 0 || a->cin % 64) return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: cin must be a positive multiple of 64 (got %d)", a->cin);
    if (a->ntaps <= 0 || a->ntaps > IRC_MAX_TAPS) return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: ntaps out of range");
    if (a->n_out <= 0 || a->n_out % 32 || a->n_out > 512) return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: n_out must be a multiple of 32, at most 512 (got %d)", a->n_out);
    int bn = a->bn;
    if (bn <= 0) {
        bn = a->n_out;
        if (bn > 256) {
            bn = 256;
            while (a->n_out % bn) bn -= 32;
        }
    }
    if (bn % 32 || bn > 256 || a->n_out % bn) return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: bad tile width %d for n_out %d", bn, a->n_out);
    const int esz = a->out_fp32 ? 4 : 2;
    if (((uintptr_t)a->out & 15) || ((long long)a->out_ld * esz) % 16 || (a->out_chan_off * esz) % 16)
        return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: output rows must be 16-byte aligned");
    if (a->mask && (((uintptr_t)a->mask & 15) || (a->mask_ld % 8) || (a->mask_chan_off % 8)))
        return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: mask rows must be 16-byte aligned");

    // two 128-row sub-tiles per tile share each weight stage (halves the B bytes per MMA) whenever that still
    // leaves at least one tile per SM
    const int sms = irc_num_sms();
    int mt = a->mt;
    // ... and both accumulator buffers still fit TMEM (an exposed epilogue costs more than the saved bytes: measured)
    if (mt <= 0) {
        mt = 1;
        for (int c = 4; c > 1; c >>= 1)
            if (2 * c * bn <= 512 && ((a->a_rows + c * kBM - 1) / (c * kBM)) * (a->n_out / bn) >= sms) { mt = c; break; }
    }
    const bool tapsum = a->tap_out != nullptr;
    if (tapsum) {
        if (a->tap_nshift < 1 || !(a->tap_nshift & 1) || a->tap_nshift * a->tap_nco > 32 || a->n_out != 32 || a->tap_nco < 1 || a->mask || a->addend || a->row_img ||
            (long long)a->tap_hp * a->tap_wp <= 0 || a->a_rows % ((long long)a->tap_hp * a->tap_wp) || a->a_rows >= (1LL << 31) ||
            (a->tap_act != 0 && a->tap_act != 3))
            return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm(tap mode): needs an odd number of horizontal taps with nshift * nco <= 32 = n_out, frames of hp x wp rows, act 0 or 3 (tanh)");
        if (mt != 1) mt = 2;
    }
    if (mt != 1 && mt != 2 && mt != 4) return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: mt must be 0, 1, 2 or 4");
    if (mt * bn > 512) return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: mt * bn exceeds the 512 TMEM columns");
    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, a->a, a->a_rows, a->a_ld, kBM * mt > 256 ? 256 : kBM * mt);
    if (rc) return rc;
    rc = make_map(&tmB, a->w, a->n_out, a->ntaps * a->cin, bn);
    if (rc) return rc;

    ConvParams p;
    p.mt = mt;
    p.nbuf = (2 * mt * bn <= 512) ? 2 : 1;
    p.rows = a->a_rows;
    p.bn = bn;
    p.n_tiles = a->n_out / bn;
    p.k_chunks = a->cin / 64;
    if (a->k_live < 0 || a->k_live > 64 || (a->k_live > 0 && a->cin != 64))
        return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: k_live (live operand columns, the rest structural zeros) needs cin == 64 and 0 <= k_live <= 64");
    p.k16 = a->k_live > 0 ? (a->k_live + 15) / 16 : 4;
    p.a_chan_off = a->a_chan_off;
    p.ntaps = a->ntaps;
    for (int i = 0; i < a->ntaps; ++i) p.taps[i] = a->taps[i];
    p.out = a->out; p.out_ld = a->out_ld; p.out_chan_off = a->out_chan_off; p.out_fp32 = a->out_fp32;
    p.bias = a->bias; p.act = a->act; p.slope = a->slope;
    p.row_img = a->row_img;
    p.mask = (const bf16*)a->mask; p.mask_ld = a->mask_ld; p.mask_chan_off = a->mask_chan_off; p.mask_slope = a->mask_slope;
    p.addend = (const bf16*)a->addend; p.addend_ld = a->addend_ld; p.addend_chan_off = a->addend_chan_off;
    if (a->addend && (((uintptr_t)a->addend & 15) || (a->addend_ld % 8) || (a->addend_chan_off % 8)))
        return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: addend rows must be 16-byte aligned");
    p.n_out = a->n_out;
    p.tile_stride = kBM * mt; p.row_bias = 0;
    p.stats_part = a->stats_part; p.stats_edge = a->stats_edge; p.rows_per_img = a->rows_per_img;
    p.tap_out = a->tap_out; p.tap_nshift = a->tap_nshift; p.tap_nco = a->tap_nco; p.tap_H = a->tap_H; p.tap_W = a->tap_W; p.tap_hp = a->tap_hp;
    p.tap_wp = a->tap_wp; p.tap_oy = a->tap_oy; p.tap_ox = a->tap_ox; p.tap_act = a->tap_act;
    p.tap_scale = a->tap_scale; p.tap_accumulate = a->tap_accumulate;
    if (tapsum) {
        const int halo = (a->tap_nshift - 1) / 2;
        p.tile_stride = kBM * mt - 2 * halo; p.row_bias = -halo;
    }
    // coalesced TMA-store epilogue for bf16 outputs whose tile width is a multiple of one 64-channel swizzle row
    p.tma_store = (!tapsum && !a->out_fp32 && bn % 64 == 0 && a->epilogue_direct == 0) ? 1 : 0;
    if (p.stats_part) {
        if (!p.tma_store || !a->row_img || !a->stats_edge || a->rows_per_img < kBM)
            return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: epilogue statistics need the staged bf16 epilogue (tile width %% 64 == 0), a row_img table "
                                                  "(ring rows stored as zeros), an edge buffer and images of at least 128 rows");
    }
    CUtensorMap tmOut = tmB;
    if (p.tma_store) {
        rc = make_map(&tmOut, a->out, a->a_rows, (int)a->out_ld, kBM);
        if (rc) return rc;
    }
    const int stage_bytes = kBM * mt * 128 + bn * 128;
    const int kStatic = 2048;                              // sbias
    p.nstg = 2;
    if (p.tma_store && (kMaxSmem - kStatic - 2048 - (2 * kBM * 128 + 1024)) / stage_bytes < 4 &&
        (kMaxSmem - kStatic - 2048 - (kBM * 128 + 1024)) / stage_bytes >= 4) p.nstg = 1;
    const int stg_bytes = tapsum ? kBM * mt * ((a->tap_nshift * a->tap_nco) | 1) * 4 + 1024 : (p.tma_store ? p.nstg * kBM * 128 + 1024 : 0);
    int stages = (kMaxSmem - kStatic - 2048 - stg_bytes) / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: tile does not fit shared memory");
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 2048 + stg_bytes;
    if (!g_attr_conv) {
        if (cudaFuncSetAttribute(conv_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 2048) != cudaSuccess ||
            cudaFuncSetAttribute(conv_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 2048) != cudaSuccess ||
            cudaFuncSetAttribute(conv_gemm_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 2048) != cudaSuccess)
            return irc_check_launch("cudaFuncSetAttribute(conv_gemm)");
        g_attr_conv = true;
    }
    const long long tiles = ((a->a_rows + p.tile_stride - 1) / p.tile_stride) * p.n_tiles;
    const int grid = (int)(tiles < sms ? tiles : sms);
    // CTA pairs (cta_group::2) win where the weight tile is wide and the reduction deep (measured per layer,
    // profiles/gemm_breakdown_r2_pair.csv: ResNet blocks 0.84 -> 0.93 of the sustained peak, VGG conv3_x 0.91 -> 1.00, up1 0.77 -> 0.82)
    const int kred = a->ntaps * a->cin;
    const bool pair_auto = ((bn == 256 && kred >= 1024) || (bn == 128 && kred >= 2048)) && a->a_rows >= (long long)sms * 128;
    // tap-run variant: taps with consecutive row shifts share one staged A box (reuse = 0 off, 1 on wherever taps form runs,
    // -1 auto: the layers the pair kernel does not take, i.e. the L2-bandwidth-bound ones with <= 128 output columns)
    RunParams rp;
    build_runs(a->taps, a->ntaps, rp);
    int max_len = 1;
    for (int r = 0; r < rp.nruns; ++r) if (rp.run_len[r] > max_len) max_len = rp.run_len[r];
    // packed taps (reuse = 2, or auto): 64 output channels, every run exactly three taps long (3 x 3 kernels in either role)
    bool pack_ok = a->n_out == 64 && bn == 64 && !tapsum && p.tma_store && !p.stats_part && rp.nruns * 3 == a->ntaps && rp.nruns <= IRC_MAX_TAPS / 3;
    for (int r = 0; r < rp.nruns && pack_ok; ++r) pack_ok = rp.run_len[r] == 3;
    // auto: only with >= 2 channel chunks per tap - with one, the 12 MMAs of a tile (~1900 cycles) finish before the epilogue has
    // read its 128 x 192 fp32 accumulator columns out of TMEM (~1500 cycles at 64 B/clk) and done the shifted sum (measured: VGG
    // conv1_2 forward 208 -> 269 us, while up2 192 -> 64 goes 272 -> 214 us and down1's data gradient 197 -> 142 us)
    if (pack_ok && (a->reuse == 2 || (a->reuse < 0 && a->cin >= 128 && a->a_rows >= (long long)sms * 128))) {
        PackParams pk;
        pk.nruns = rp.nruns;
        for (int r = 0; r < rp.nruns; ++r) {
            pk.run_mid[r] = rp.run_first[r] + 1;
            for (int j = 0; j < 3; ++j) pk.run_tap[r][j] = rp.run_tap[r][j];
        }
        CUtensorMap tmA1, tmB64, tmOut126;
        rc = make_map(&tmA1, a->a, a->a_rows, a->a_ld, kBM);
        if (rc) return rc;
        rc = make_map(&tmB64, a->w, a->n_out, a->ntaps * a->cin, 64);
        if (rc) return rc;
        rc = make_map(&tmOut126, a->out, a->a_rows, (int)a->out_ld, kPackRows);
        if (rc) return rc;
        ConvParams q = p; q.mt = 1; q.nbuf = 2; q.nstg = 2;
        const int sbytes = kBM * 128 + 3 * 64 * 128;
        int st = (kMaxSmem - kStatic - 1024 - 8192 - 2 * kBM * 128) / sbytes;
        if (st > 6) st = 6;
        q.stages = st;
        const size_t smem2 = (size_t)st * sbytes + 8192 + 2 * kBM * 128 + 1024;
        static bool attr_pack = false;
        if (!attr_pack) {
            if (cudaFuncSetAttribute(conv_gemm_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 2048) != cudaSuccess)
                return irc_check_launch("cudaFuncSetAttribute(conv_gemm_pack)");
            attr_pack = true;
        }
        const long long tiles_p = (a->a_rows + kPackRows - 1) / kPackRows;
        irc::launch<1>(conv_gemm_pack_kernel, (int)(tiles_p < sms ? tiles_p : sms), kConvThreads, smem2, (cudaStream_t)stream, tmOut126, tmA1, tmB64, q, pk);
        return irc_check_launch("irc_conv_gemm(packed taps)");
    }
    // (128 output columns: only with one channel chunk per tap - measured, same box: conv2_1 forward 109 -> 101 us, but conv2_2
    // forward, two chunks, 143 -> 156..168 us)
    const bool runs_auto = max_len > 1 && !pair_auto && (bn <= 64 || (bn <= 128 && a->cin == 64)) && a->a_rows >= (long long)sms * 128;
    if (max_len > 1 && !tapsum && (a->reuse == 1 || a->reuse == 3 || (a->reuse < 0 && runs_auto))) {
        rp.a_stage_bytes = (kBM * mt + 8) * 128;
        rp.base_off_mode = a->reuse == 3 ? 1 : 0;      // 3 = the (wrong) base-offset encoding, kept for the experiment script
        const int stage_b = bn * 128;
        const int stg_b = p.tma_store ? 2 * kBM * 128 + 1024 : 0;
        const int budget = kMaxSmem - kStatic - 2048 - stg_b;
        int sb = 2 * max_len; if (sb > 8) sb = 8;
        while (sb > 2 && budget - sb * stage_b < 2 * rp.a_stage_bytes) --sb;
        int sa = (budget - sb * stage_b) / rp.a_stage_bytes;
        if (sa > 6) sa = 6;
        if (sa < 2) return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: tap-run tile does not fit shared memory");
        const int extra = (budget - sa * rp.a_stage_bytes) / stage_b;
        if (extra > sb) sb = extra > 8 ? 8 : extra;
        rp.sa_stages = sa; rp.sb_stages = sb;
        CUtensorMap tmA8;
        rc = make_map(&tmA8, a->a, a->a_rows, a->a_ld, 8);
        if (rc) return rc;
        ConvParams q = p; q.nstg = 2;
        const size_t smem2 = (size_t)sa * rp.a_stage_bytes + (size_t)sb * stage_b + 2048 + stg_b;
        if (!g_attr_runs) {
            if (cudaFuncSetAttribute(conv_gemm_runs_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 2048) != cudaSuccess ||
                cudaFuncSetAttribute(conv_gemm_runs_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 2048) != cudaSuccess ||
                cudaFuncSetAttribute(conv_gemm_runs_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 2048) != cudaSuccess)
                return irc_check_launch("cudaFuncSetAttribute(conv_gemm_runs)");
            g_attr_runs = true;
        }
        cudaStream_t st = (cudaStream_t)stream;
        if (mt == 1) irc::launch<1>(conv_gemm_runs_kernel<1>, grid, kConvThreads, smem2, st, tmOut, tmA, tmA8, tmB, q, rp);
        else if (mt == 2) irc::launch<1>(conv_gemm_runs_kernel<2>, grid, kConvThreads, smem2, st, tmOut, tmA, tmA8, tmB, q, rp);
        else irc::launch<1>(conv_gemm_runs_kernel<4>, grid, kConvThreads, smem2, st, tmOut, tmA, tmA8, tmB, q, rp);
        return irc_check_launch("irc_conv_gemm(runs)");
    }
    // CTA pairs (cta_group::2): IRC_CONV_PAIR=1 all eligible launches, =0 never, default: auto (see pair_auto)
    static int pair_mode = -2;
    if (pair_mode == -2) { const char* e = getenv("IRC_CONV_PAIR"); pair_mode = e ? atoi(e) : -1; }
    const bool pair_ok = !tapsum && (bn == 64 || bn == 128 || bn == 256) && a->mt <= 1 && sms >= 2;
    if (pair_ok && (pair_mode == 1 || (pair_mode == -1 && pair_auto))) {
        ConvParams q = p; q.mt = 1; q.nbuf = 2; q.tile_stride = 2 * kBM; q.row_bias = 0;
        CUtensorMap tmA1, tmBh;
        rc = make_map(&tmA1, a->a, a->a_rows, a->a_ld, kBM);
        if (rc) return rc;
        rc = make_map(&tmBh, a->w, a->n_out, a->ntaps * a->cin, bn / 2);
        if (rc) return rc;
        const int sb2 = kBM * 128 + (bn / 2) * 128;
        q.nstg = 2;
        const int stg2 = q.tma_store ? q.nstg * kBM * 128 + 1024 : 0;
        int st2 = (kMaxSmem - kStatic - 2048 - stg2) / sb2;
        if (st2 > 8) st2 = 8;
        q.stages = st2;
        const size_t smem2 = (size_t)st2 * sb2 + 2048 + stg2;
        static bool attr2 = false;
        if (!attr2) {
            if (cudaFuncSetAttribute(conv_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - 2048) != cudaSuccess)
                return irc_check_launch("cudaFuncSetAttribute(conv_gemm_pair)");
            attr2 = true;
        }
        const long long tiles2 = ((a->a_rows + 2 * kBM - 1) / (2 * kBM)) * q.n_tiles;
        const long long max_pairs = sms / 2;
        const unsigned grid2 = 2u * (unsigned)(tiles2 < max_pairs ? tiles2 : max_pairs);
        irc::launch_cluster(conv_gemm_pair_kernel, grid2, kConvThreads, smem2, (cudaStream_t)stream, 2, tmOut, tmA1, tmBh, q);
        return irc_check_launch("irc_conv_gemm(pair)");
    }
    if (mt == 1) irc::launch<1>(conv_gemm_kernel<1>, grid, kConvThreads, smem, (cudaStream_t)stream, tmOut, tmA, tmB, p);
    else if (mt == 2) irc::launch<1>(conv_gemm_kernel<2>, grid, kConvThreads, smem, (cudaStream_t)stream, tmOut, tmA, tmB, p);
    else irc::launch<1>(conv_gemm_kernel<4>, grid, kConvThreads, smem, (cudaStream_t)stream, tmOut, tmA, tmB, p);
    return irc_check_launch("irc_conv_gemm");
}

extern "C" int irc_tn_gemm(const irc_tn_gemm_args* a, void* stream) {
    if (!a || !a->a || !a->b || !a->out) return irc_set_error(IRC_ERR_BAD_ARG, "irc_tn_gemm: null pointer");
    if (a->ntaps <= 0 || a->ntaps > IRC_MAX_TAPS) return irc_set_error(IRC_ERR_BAD_ARG, "irc_tn_gemm: ntaps out of range");
    if (a->m <= 0 || a->n <= 0 || a->splits <= 0) return irc_set_error(IRC_ERR_BAD_ARG, "irc_tn_gemm: bad extents");
    int bn = a->bn;
    if (bn <= 0) { bn = ((a->n + 63) / 64) * 64; if (bn > 256) bn = 256; }
    if (bn % 64 || bn > 256) return irc_set_error(IRC_ERR_BAD_ARG, "irc_tn_gemm: tile width must be a multiple of 64 <= 256");
    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, a->a, a->a_rows, a->a_ld, kBK);
    if (rc) return rc;
    rc = make_map(&tmB, a->b, a->b_rows, a->b_ld, kBK);
    if (rc) return rc;
    TnParams p;
    p.k_rows = a->k_rows; p.m = a->m; p.n = a->n; p.bn = bn;
    p.m_tiles = (a->m + kBM - 1) / kBM;
    p.n_tiles = (a->n + bn - 1) / bn;
    p.ntaps = a->ntaps; p.splits = a->splits;
    p.a_chan_off = a->a_chan_off; p.b_chan_off = a->b_chan_off;
    for (int i = 0; i < a->ntaps; ++i) { p.a_shift[i] = a->a_shift[i]; p.b_shift[i] = a->b_shift[i]; }
    p.out = a->out; p.out_tap_stride = a->out_tap_stride; p.out_m_stride = a->out_m_stride;
    p.out_n_stride = a->out_n_stride; p.out_split_stride = a->out_split_stride;
    static int merge_mode = -1;
    if (merge_mode < 0) { const char* e = getenv("IRC_TN_MERGE"); merge_mode = (e && e[0] == '0') ? 0 : 1; }
    p.merge_taps = merge_mode;
    bool zero_a = true;
    for (int i = 0; i < a->ntaps; ++i) zero_a = zero_a && a->a_shift[i] == 0;
    if (a->tpc <= 0 && tn_pair_mode(a->m, a->ntaps, zero_a) && !getenv("IRC_TN_NOPAIR")) {
        const int P = auto_pairs(bn, a->ntaps);
        p.groups = ((a->ntaps + 1) / 2 + P - 1) / P;
        p.m_tiles = 1;
        int cols = 32; while (cols < P * bn) cols <<= 1;
        p.tmem_cols = cols;
        const int stage_bytes = P * kBK * 128 * 2 + kBK * 128 * (bn / 64);
        int stages = (kMaxSmem - 2048) / stage_bytes;
        if (stages > 8) stages = 8;
        p.stages = stages;
        const size_t smem = (size_t)stages * stage_bytes + 2048;
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(tn_gemm_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
            cudaFuncSetAttribute(tn_gemm_pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
            cudaFuncSetAttribute(tn_gemm_pair_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
            cudaFuncSetAttribute(tn_gemm_pair_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
            attr = true;
        }
        const unsigned grid = (unsigned)((long long)p.n_tiles * p.groups * p.splits);
        cudaStream_t st = (cudaStream_t)stream;
        switch (P) {
            case 1: irc::launch<1>(tn_gemm_pair_kernel<1>, grid, kThreads, smem, st, tmA, tmB, p); break;
            case 2: irc::launch<1>(tn_gemm_pair_kernel<2>, grid, kThreads, smem, st, tmA, tmB, p); break;
            case 3: irc::launch<1>(tn_gemm_pair_kernel<3>, grid, kThreads, smem, st, tmA, tmB, p); break;
            default: irc::launch<1>(tn_gemm_pair_kernel<4>, grid, kThreads, smem, st, tmA, tmB, p); break;
        }
        return irc_check_launch("irc_tn_gemm(pair)");
    }
    // taps per CTA: all taps of a group must share the A shift (true for weight gradients: a_shift == 0)
    int tpc = a->tpc;
    bool same_a = true;
    for (int i = 1; i < a->ntaps; ++i) same_a = same_a && a->a_shift[i] == a->a_shift[0];
    if (tpc <= 0) {
        tpc = same_a ? auto_tpc(bn, a->ntaps) : 1;
    }
    if (!(tpc == 1 || tpc == 2 || tpc == 3 || tpc == 4 || tpc == 8) || tpc * bn > 512 || (tpc > 1 && !same_a))
        return irc_set_error(IRC_ERR_BAD_ARG, "irc_tn_gemm: unsupported taps-per-CTA %d for tile width %d", tpc, bn);
    // CTA pairs (cta_group::2) for layers with >= 256 output channels and a tile width that splits into two 64-channel-group halves
    static int tn_pair_mode = -2;
    if (tn_pair_mode == -2) { const char* e = getenv("IRC_TN_PAIR"); tn_pair_mode = e ? atoi(e) : -1; }
    if (tn_pair_mode != 0 && a->tpc <= 0 && tn_pair_shape(a->m, a->n, bn, same_a) && irc_num_sms() >= 2) {
        int tp = tn_pair_taps(bn, a->ntaps);
        p.groups = (a->ntaps + tp - 1) / tp;
        p.m_tiles = a->m / (2 * kBM);
        int cols2 = 32; while (cols2 < tp * bn) cols2 <<= 1;
        p.tmem_cols = cols2;
        const int sb = kBK * 128 * 2 + tp * kBK * 128 * (bn / 128);
        int st = (kMaxSmem - 2048) / sb;
        if (st > 8) st = 8;
        p.stages = st;
        const size_t smem2 = (size_t)st * sb + 2048;
        static bool attr2 = false;
        if (!attr2) {
            if (cudaFuncSetAttribute(tn_gemm_cta2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess ||
                cudaFuncSetAttribute(tn_gemm_cta2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess)
                return irc_check_launch("cudaFuncSetAttribute(tn_gemm_cta2)");
            attr2 = true;
        }
        const unsigned grid2 = 2u * (unsigned)((long long)p.m_tiles * p.n_tiles * p.groups * p.splits);
        cudaStream_t st_ = (cudaStream_t)stream;
        if (tp == 2) irc::launch_cluster(tn_gemm_cta2_kernel<2>, grid2, kThreads, smem2, st_, 2, tmA, tmB, p);
        else irc::launch_cluster(tn_gemm_cta2_kernel<1>, grid2, kThreads, smem2, st_, 2, tmA, tmB, p);
        return irc_check_launch("irc_tn_gemm(cta pair)");
    }
    p.groups = (a->ntaps + tpc - 1) / tpc;
    int cols = 32; while (cols < tpc * bn) cols <<= 1;
    p.tmem_cols = cols;
    const int stage_bytes = kBK * 128 * 2 + tpc * kBK * 128 * (bn / 64);
    int stages = (kMaxSmem - 2048) / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) return irc_set_error(IRC_ERR_BAD_ARG, "irc_tn_gemm: stage does not fit shared memory");
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 2048;
    if (!g_attr_tn) {
        if (cudaFuncSetAttribute(tn_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess ||
            cudaFuncSetAttribute(tn_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess ||
            cudaFuncSetAttribute(tn_gemm_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess ||
            cudaFuncSetAttribute(tn_gemm_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess ||
            cudaFuncSetAttribute(tn_gemm_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess)
            return irc_check_launch("cudaFuncSetAttribute(tn_gemm)");
        g_attr_tn = true;
    }
    const unsigned grid = (unsigned)((long long)p.m_tiles * p.n_tiles * p.groups * p.splits);
    cudaStream_t st = (cudaStream_t)stream;
    switch (tpc) {
        case 1: irc::launch<1>(tn_gemm_kernel<1>, grid, kThreads, smem, st, tmA, tmB, p); break;
        case 2: irc::launch<1>(tn_gemm_kernel<2>, grid, kThreads, smem, st, tmA, tmB, p); break;
        case 3: irc::launch<1>(tn_gemm_kernel<3>, grid, kThreads, smem, st, tmA, tmB, p); break;
        case 4: irc::launch<1>(tn_gemm_kernel<4>, grid, kThreads, smem, st, tmA, tmB, p); break;
        default: irc::launch<1>(tn_gemm_kernel<8>, grid, kThreads, smem, st, tmA, tmB, p); break;
    }
    return irc_check_launch("irc_tn_gemm");
}

// number of CTAs irc_tn_gemm launches per split with the automatic tiling (lets the caller pick `splits` for one wave)
extern "C" int irc_tn_gemm_ctas(int m, int n, int ntaps, int same_a_shift) {
    int bn = ((n + 63) / 64) * 64; if (bn > 256) bn = 256;
    if (tn_pair_mode(m, ntaps, same_a_shift != 0) && !getenv("IRC_TN_NOPAIR")) {
        const int P = auto_pairs(bn, ntaps);
        return ((n + bn - 1) / bn) * (((ntaps + 1) / 2 + P - 1) / P);
    }
    if (tn_pair_shape(m, n, bn, same_a_shift != 0)) {
        const int tp = tn_pair_taps(bn, ntaps);
        return 2 * (m / (2 * kBM)) * (n / bn) * ((ntaps + tp - 1) / tp);       // CTAs (two per pair)
    }
    const int tpc = same_a_shift ? auto_tpc(bn, ntaps) : 1;
    return ((m + kBM - 1) / kBM) * ((n + bn - 1) / bn) * ((ntaps + tpc - 1) / tpc);
}

// floats of the per-sub-tile statistics partials irc_conv_gemm writes for a [rows][n_out] output (args->stats_part); the edge
// buffer (args->stats_edge) needs (n_img + 1) * n_out * 2 floats
extern "C" long long irc_conv_stats_workspace_floats(long long rows, int n_out) {
    return ((rows + kBM - 1) / kBM) * (long long)n_out * 2;
}

extern "C" int irc_conv_stats_finalize(const float* part, const float* edge, int n_img, int rows_per_img, int n_out, float* stats, void* stream) {
    if (!part || !edge || !stats || n_img <= 0 || rows_per_img < kBM || n_out <= 0) return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_stats_finalize: bad args");
    irc::launch(conv_stats_finalize_kernel, dim3(n_img, (n_out + 63) / 64), 1024, 0, (cudaStream_t)stream, part, edge, rows_per_img, n_out, stats);
    return irc_check_launch("irc_conv_stats_finalize");
}

// nn.ConvTranspose2d(Cin, Cout, 3, stride=2, padding=1, output_padding=1) forward (irc:495-499, :512-516) on a framed input
// (pad >= 1, ZERO ring): output pixel (2y + a, 2x + b) only sees inputs (y + dy, x + dx), dy, dx in {0, 1}, so the transposed
// convolution is ONE stride-1 implicit GEMM with four taps whose 4 * Cout output columns are the four sub-pixel phases
// (depth-to-space order: column (a * 2 + b) * Cout + co).  w: bf16 [4 * cout][4 * cin] packed as (phase, co) x (tap dy * 2 + dx, ci)
// with the kernel element (a + 1 - 2 dy, b + 1 - 2 dx) or zero; bias4: fp32 [4 * cout] (the bias repeated per phase) or NULL;
// out: bf16 [rows][out_ld], columns [out_chan_off, out_chan_off + 4 * cout).  Launched as ceil(4 * cout / 512) irc_conv_gemm calls.
extern "C" int irc_convT2d_fwd(const void* x, long long rows, int x_ld, int x_chan_off, int cin, int wp, const void* w, int cout, const float* bias4,
                               void* out, long long out_ld, int out_chan_off, void* stream) {
    if (!x || !w || !out || cin <= 0 || cin % 64 || cout <= 0 || (4 * cout) % 32 || wp < 2)
        return irc_set_error(IRC_ERR_BAD_ARG, "irc_convT2d_fwd: cin must be a multiple of 64, 4 * cout a multiple of 32, wp >= 2");
    const int n_total = 4 * cout;
    int blk = n_total;
    if (blk > 512) { blk = 512; while (n_total % blk) blk -= 32; }
    for (int c0 = 0; c0 < n_total; c0 += blk) {
        irc_conv_gemm_args g;
        memset(&g, 0, sizeof(g));
        g.a = x; g.a_rows = rows; g.a_ld = x_ld; g.a_chan_off = x_chan_off; g.cin = cin;
        g.ntaps = 4; g.taps[0] = 0; g.taps[1] = 1; g.taps[2] = wp; g.taps[3] = wp + 1;
        g.w = (const char*)w + (size_t)c0 * 4 * cin * 2; g.n_out = blk;
        g.out = out; g.out_ld = out_ld; g.out_chan_off = out_chan_off + c0;
        g.bias = bias4 ? bias4 + c0 : nullptr;
        g.reuse = 0;
        const int rc = irc_conv_gemm(&g, stream);
        if (rc) return rc;
    }
    return IRC_OK;
}
