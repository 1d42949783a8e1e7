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
#include "irc_common.cuh"
#include "../../include/irc_b200.h"

using namespace irc;

namespace {

constexpr int kThreads = 192;
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
    float* stats;   // [n_img][n_out][2] or null
    int n_out;
};

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v > 0.f ? v : v * slope;
    return v;
}

__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int tile_rows = kBM * p.mt;
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

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m_tiles = (p.rows + tile_rows - 1) / tile_rows;
    const uint32_t acc_cols = (uint32_t)(p.mt * p.bn);      // TMEM columns of one accumulator buffer
    const long long total_tiles = m_tiles * p.n_tiles;
    const int num_kb = p.ntaps * p.k_chunks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const long long row0 = (tile / p.n_tiles) * tile_rows;
                const int n0 = (int)(tile % p.n_tiles) * p.bn;
                for (int t = 0; t < p.ntaps; ++t) {
                    const long long arow = row0 + p.taps[t];
                    for (int kc = 0; kc < p.k_chunks; ++kc) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        mbar_expect_tx(&full[stage], stage_bytes);
                        uint8_t* sa = smem + (size_t)stage * stage_bytes;
                        tma_load_2d(sa, &tmA, &full[stage], p.a_chan_off + kc * kBK, (int)arow);
                        tma_load_2d(sa + stage_a, &tmB, &full[stage], (t * p.k_chunks + kc) * kBK, n0);
                        if (++stage == S) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
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
                    for (int m = 0; m < p.mt; ++m) {
                        const uint64_t adesc = umma_desc_sw128(sa + m * (kBM * 128), 16);
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k)
                            umma_bf16(d_tmem + (uint32_t)(m * p.bn), adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[acc]);
                if (++acc == p.nbuf) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (4 warps, one TMEM lane quarter each) =====================
        const int quarter = warp & 3;
        const int r_in_tile = quarter * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int n0 = (int)(tile % p.n_tiles) * p.bn;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            for (int m = 0; m < p.mt; ++m) {
                const long long row = (tile / p.n_tiles) * tile_rows + m * kBM + r_in_tile;
                const bool in_range = row < p.rows;
                int img = 0;
                if (p.row_img) img = in_range ? (int)p.row_img[row] : -1;
                const bool live = in_range && img >= 0;
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * acc_cols + (uint32_t)(m * p.bn);
                for (int c0 = 0; c0 < p.bn; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c0, r);
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += __ldg(p.bias + n0 + c0 + j);
                    }
                    if (p.mask && live) {
                        const uint4* mp = reinterpret_cast<const uint4*>(p.mask + row * p.mask_ld + p.mask_chan_off + n0 + c0);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const uint4 mv = __ldg(mp + g);
                            const uint32_t w[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                const float2 f = unpack_bf16x2(w[h]);
                                if (!(f.x > 0.f)) v[g * 8 + h * 2] *= p.mask_slope;
                                if (!(f.y > 0.f)) v[g * 8 + h * 2 + 1] *= p.mask_slope;
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = live ? apply_act(v[j], p.act, p.slope) : 0.f;
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
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == p.nbuf) { acc = 0; acc_phase ^= 1; }
        }
    }

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
    int a_chan_off, b_chan_off;
    int a_shift[IRC_MAX_TAPS];
    int b_shift[IRC_MAX_TAPS];
    float* out;
    long long out_tap_stride, out_m_stride, out_n_stride, out_split_stride;
};

__global__ void __launch_bounds__(kThreads, 1)
tn_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int stage_a = kBK * 128 * 2;             // 2 groups of 64 channels x 64 rows
    const int groups_b = p.bn / 64;
    const int stage_b = kBK * 128 * groups_b;
    const int stage_bytes = stage_a + stage_b;
    const int S = p.stages;
    uint64_t* bars = (uint64_t*)(smem + (size_t)S * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // work decomposition: blockIdx.x -> (split, tap, m_tile, n_tile)
    int w = blockIdx.x;
    const int n_tile = w % p.n_tiles; w /= p.n_tiles;
    const int m_tile = w % p.m_tiles; w /= p.m_tiles;
    const int tap = w % p.ntaps; w /= p.ntaps;
    const int split = w;
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
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (long long kb = kb_begin; kb < kb_begin + my_kb; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_expect_tx(&full[stage], stage_bytes);
                uint8_t* sa = smem + (size_t)stage * stage_bytes;
                const long long r0 = kb * kBK;
                for (int g = 0; g < 2; ++g)
                    tma_load_2d(sa + g * (kBK * 128), &tmA, &full[stage], p.a_chan_off + m_tile * kBM + g * 64, (int)(r0 + p.a_shift[tap]));
                for (int g = 0; g < groups_b; ++g)
                    tma_load_2d(sa + stage_a + g * (kBK * 128), &tmB, &full[stage], p.b_chan_off + n_tile * p.bn + g * 64, (int)(r0 + p.b_shift[tap]));
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(kBM, p.bn, 1, 1);
            int stage = 0; uint32_t phase = 0;
            for (long long kb = 0; kb < my_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint64_t adesc = umma_desc_sw128(sa, kBK * 128);
                const uint64_t bdesc = umma_desc_sw128(sa + stage_a, kBK * 128);
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k)   // 16 reduction rows = 2048 bytes per step
                    umma_bf16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (kb | k) != 0);
                umma_commit(&empty[stage]);
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
            umma_commit(tfull);
        }
    } else {
        const int quarter = warp & 3;
        const int m = m_tile * kBM + quarter * 32 + lane;
        float* obase = p.out + (long long)split * p.out_split_stride + (long long)tap * p.out_tap_stride + (long long)m * p.out_m_stride;
        if (my_kb > 0) {
            mbar_wait(tfull, 0);
            tc_fence_after();
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
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
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = n_tile * p.bn + c0 + j;
                    if (n < p.n) obase[(long long)n * p.out_n_stride] = __uint_as_float(r[j]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

bool g_attr_conv = false, g_attr_tn = false;

}  // namespace

// ------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------
extern "C" int irc_conv_gemm(const irc_conv_gemm_args* a, void* stream) {
    if (!a || !a->a || !a->w || !a->out) return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: null pointer");
    if (a->cin <= 0 || a->cin % 64) return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: cin must be a positive multiple of 64 (got %d)", a->cin);
    if (a->ntaps <= 0 || a->ntaps > IRC_MAX_TAPS) return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: ntaps out of range");
    if (a->n_out <= 0 || a->n_out % 32) return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: n_out must be a multiple of 32 (got %d)", a->n_out);
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
    if (mt <= 0) mt = (4 * bn <= 512 && ((a->a_rows + 2 * kBM - 1) / (2 * kBM)) * (a->n_out / bn) >= sms) ? 2 : 1;
    if (mt != 1 && mt != 2) return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: mt must be 0, 1 or 2");
    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, a->a, a->a_rows, a->a_ld, kBM * mt);
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
    p.a_chan_off = a->a_chan_off;
    p.ntaps = a->ntaps;
    for (int i = 0; i < a->ntaps; ++i) p.taps[i] = a->taps[i];
    p.out = a->out; p.out_ld = a->out_ld; p.out_chan_off = a->out_chan_off; p.out_fp32 = a->out_fp32;
    p.bias = a->bias; p.act = a->act; p.slope = a->slope;
    p.row_img = a->row_img;
    p.mask = (const bf16*)a->mask; p.mask_ld = a->mask_ld; p.mask_chan_off = a->mask_chan_off; p.mask_slope = a->mask_slope;
    p.stats = nullptr; p.n_out = a->n_out;
    const int stage_bytes = kBM * mt * 128 + bn * 128;
    int stages = (kMaxSmem - 2048) / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) return irc_set_error(IRC_ERR_BAD_ARG, "irc_conv_gemm: tile does not fit shared memory");
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 2048;
    if (!g_attr_conv) {
        if (cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess)
            return irc_check_launch("cudaFuncSetAttribute(conv_gemm)");
        g_attr_conv = true;
    }
    const long long tiles = ((a->a_rows + kBM * mt - 1) / (kBM * mt)) * p.n_tiles;
    const int grid = (int)(tiles < sms ? tiles : sms);
    conv_gemm_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(tmA, tmB, p);
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
    const int stage_bytes = kBK * 128 * 2 + kBK * 128 * (bn / 64);
    int stages = (kMaxSmem - 2048) / stage_bytes;
    if (stages > 8) stages = 8;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 2048;
    if (!g_attr_tn) {
        if (cudaFuncSetAttribute(tn_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess)
            return irc_check_launch("cudaFuncSetAttribute(tn_gemm)");
        g_attr_tn = true;
    }
    const long long grid = (long long)p.m_tiles * p.n_tiles * p.ntaps * p.splits;
    tn_gemm_kernel<<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(tmA, tmB, p);
    return irc_check_launch("irc_tn_gemm");
}
