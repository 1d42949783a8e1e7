// Fused multi-tensor Adam over one flat fp32 parameter arena (torch.optim.Adam defaults,
// irc:1601-1604, :1651, :1681), weight packing for the GEMM operand layouts, and the
// deterministic reduction of split weight-gradient partials back into OIHW order.
#include "irc_common.cuh"
#include "../../include/irc_b200.h"

using namespace irc;

namespace {

// hyper (device, fp64) = {lr, beta1, beta2, eps, lr_scale, grad_scale}; *step = optimizer steps taken so far.
// The step counter lives on the device and is advanced by adam_bump_kernel right behind this kernel on the same stream, so a
// replayed CUDA graph computes the right bias corrections 1 - beta^t without any per-step host write (a pinned staging
// buffer rewritten by a host that runs ahead of the GPU would race with its own earlier copies).
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                            const double* __restrict__ hyper, const long long* __restrict__ step_count) {
    irc::pdl_prologue();
    const double t = (double)(*step_count + 1);
    const double bc1d = 1.0 - pow(hyper[1], t), bc2d = 1.0 - pow(hyper[2], t);
    const float b1 = (float)hyper[1], b2 = (float)hyper[2], eps = (float)hyper[3], gs = (float)hyper[5];
    const float step = (float)(hyper[0] * hyper[4] / bc1d), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2d));
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        float* pe = &pp.x; const float* ge = &gg.x; float* me = &mm.x; float* ve = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = ge[k] * gs;
            me[k] = b1 * me[k] + (1.f - b1) * gk;
            ve[k] = b2 * ve[k] + (1.f - b2) * gk * gk;
            pe[k] -= step * __fdiv_rn(me[k], sqrtf(ve[k]) * inv_sqrt_bc2 + eps);
        }
        reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gk = g[i] * gs;
        const float mk = b1 * m[i] + (1.f - b1) * gk, vk = b2 * v[i] + (1.f - b2) * gk * gk;
        m[i] = mk; v[i] = vk;
        p[i] -= step * __fdiv_rn(mk, sqrtf(vk) * inv_sqrt_bc2 + eps);
    }
}

__global__ void adam_bump_kernel(long long* step_count) {
    irc::pdl_prologue();
    if (threadIdx.x == 0 && blockIdx.x == 0) *step_count += 1;
}

__global__ void pack_kernel(const float* __restrict__ src, const int* __restrict__ map, long long n, bf16* __restrict__ dst) {
    irc::pdl_prologue();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int j = map[i];
        dst[i] = __float2bfloat16(j >= 0 ? src[j] : 0.f);
    }
}

// 8 outputs per thread: two 16-byte map loads, eight gathers in flight, one 16-byte store (n % 8 == 0, aligned buffers)
__global__ void pack8_kernel(const float* __restrict__ src, const int* __restrict__ map, long long n8, bf16* __restrict__ dst) {
    irc::pdl_prologue();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const int4 a = __ldg(reinterpret_cast<const int4*>(map) + 2 * i), b = __ldg(reinterpret_cast<const int4*>(map) + 2 * i + 1);
        const int j[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = j[k] >= 0 ? __ldg(src + j[k]) : 0.f;
        reinterpret_cast<uint4*>(dst)[i] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
}

__global__ void gather_f32_kernel(const float* __restrict__ src, const int* __restrict__ map, long long n, float* __restrict__ dst) {
    irc::pdl_prologue();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int j = map[i];
        dst[i] = j >= 0 ? src[j] : 0.f;
    }
}

__global__ void gather_sum_kernel(const float* __restrict__ src, const int* __restrict__ map, long long n, int splits, long long split_stride,
                                  float* __restrict__ dst) {
    irc::pdl_prologue();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int j = map[i];
        float a = 0.f;
        if (j >= 0) for (int s = 0; s < splits; ++s) a += src[(long long)s * split_stride + j];
        dst[i] = a;
    }
}

// several gather_sum jobs in one launch: element i of the concatenated index space belongs to the job whose
// [start, start + n) range contains it (binary search over <= a few dozen jobs; the table sits in L1 / constant cache)
__global__ void gather_sum_multi_kernel(const irc_sum_job* __restrict__ jobs, int njobs, long long total) {
    irc::pdl_prologue();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = njobs - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (jobs[mid].start <= i) lo = mid; else hi = mid - 1;
        }
        const irc_sum_job jb = jobs[lo];
        const long long e = i - jb.start;
        const int j = jb.map[e];
        float a = 0.f;
        if (j >= 0) {
            const float* sp = jb.src + j;
            int s = 0;
            for (; s + 4 <= jb.splits; s += 4) {          // four independent loads in flight, added in split order
                const float v0 = __ldg(sp), v1 = __ldg(sp + jb.split_stride), v2 = __ldg(sp + 2 * jb.split_stride), v3 = __ldg(sp + 3 * jb.split_stride);
                a += v0; a += v1; a += v2; a += v3;
                sp += 4 * jb.split_stride;
            }
            for (; s < jb.splits; ++s) { a += __ldg(sp); sp += jb.split_stride; }
        }
        jb.dst[e] = a;
    }
}

int grid_for(long long total, int threads) {
    long long b = (total + threads - 1) / threads;
    const long long cap = (long long)irc_num_sms() * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ---------------------------------------------------------------------------------
// Structured packing of standard convolution weights (OIHW fp32 -> the two bf16 GEMM operands)
// ---------------------------------------------------------------------------------
// The generic pack gathers every operand slot through an index map: 4 bytes of map per slot, and - because the forward operand
// W_f[n][t][k] walks the OIHW tensor with a stride of kh*kw floats - one 32-byte sector per 4-byte read.  For the regular layers
// (all 3 x 3 / 4 x 4 convolutions of the generator's trunk and model.8) the permutation is a small transpose: a block stages the
// [16 output channels][64 input channels][T taps] slab of one layer with coalesced reads and writes whole 128-byte rows of W_f and
// 32-byte runs of the data-gradient operand W_d[k][t][n].  One launch covers every layer of a network through a job table.
struct PackStdJob { long long src, dst_f, dst_d; int N, K, T, kd, first_block, nblocks; };

__global__ void __launch_bounds__(256) pack_std_kernel(const float* __restrict__ arena, bf16* __restrict__ packed, const PackStdJob* __restrict__ jobs, int njobs) {
    irc::pdl_prologue();
    extern __shared__ float slab[];           // [16][64 * T + 1]
    int j = 0;
    while (j + 1 < njobs && (int)blockIdx.x >= jobs[j + 1].first_block) ++j;
    const PackStdJob job = jobs[j];
    const int b = blockIdx.x - job.first_block;
    const int kblocks = job.K >> 6;
    const int n0 = (b / kblocks) * 16, k0 = (b % kblocks) * 64;
    const int T = job.T, run = 64 * T, pitch = run + 1;
    // OIHW: element (n, k, t) at src + (n * K + k) * T + t; for fixed n the 64 k x T taps are one contiguous run
    for (int i = threadIdx.x; i < 16 * run; i += 256) {
        const int nn = i / run, r = i - nn * run;
        slab[nn * pitch + r] = (n0 + nn < job.N) ? __ldg(arena + job.src + ((long long)(n0 + nn) * job.K + k0) * T + r) : 0.f;
    }
    __syncthreads();
    // W_f[n][t * K + k]: rows of 64 k (128 bytes) per (n, t); one thread writes 8 k
    for (int i = threadIdx.x; i < 16 * T * 8; i += 256) {
        const int kg = i & 7, t = (i >> 3) % T, nn = (i >> 3) / T;
        if (n0 + nn >= job.N) continue;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = slab[nn * pitch + (kg * 8 + q) * T + t];
        *reinterpret_cast<uint4*>(packed + job.dst_f + ((long long)(n0 + nn) * T + t) * job.K + k0 + kg * 8) =
            make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
    // W_d[k][t * kd + n]: runs of 16 n (32 bytes) per (k, t); one thread writes 8 n
    for (int i = threadIdx.x; i < 64 * T * 2; i += 256) {
        const int ng = i & 1, t = (i >> 1) % T, kk = (i >> 1) / T;
        if (n0 + ng * 8 >= job.N) continue;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = slab[(ng * 8 + q) * pitch + kk * T + t];
        *reinterpret_cast<uint4*>(packed + job.dst_d + ((long long)(k0 + kk) * T + t) * job.kd + n0 + ng * 8) =
            make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
}

}  // namespace

extern "C" int irc_adam(float* p, const float* g, float* m, float* v, long long n, const double* hyper, long long* step_count, void* stream) {
    if (!p || !g || !m || !v || !hyper || !step_count) return irc_set_error(IRC_ERR_BAD_ARG, "irc_adam: null");
    if (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) return irc_set_error(IRC_ERR_BAD_ARG, "irc_adam: arenas must be 16-byte aligned");
    irc::launch(adam_kernel, grid_for(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream, p, g, m, v, n, hyper, (const long long*)step_count);
    irc::launch(adam_bump_kernel, 1, 32, 0, (cudaStream_t)stream, step_count);
    return irc_check_launch("irc_adam");
}

extern "C" int irc_pack_bf16(const float* src, const int* map, long long n, void* dst, void* stream) {
    if (!src || !map || !dst) return irc_set_error(IRC_ERR_BAD_ARG, "irc_pack_bf16: null");
    if (n % 8 == 0 && !((uintptr_t)map & 15) && !((uintptr_t)dst & 15))
        irc::launch(pack8_kernel, grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream, src, map, n / 8, (bf16*)dst);
    else
        irc::launch(pack_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, src, map, n, (bf16*)dst);
    return irc_check_launch("irc_pack_bf16");
}

extern "C" int irc_gather_sum(const float* src, const int* map, long long n, int splits, long long split_stride, float* dst, void* stream) {
    if (!src || !map || !dst) return irc_set_error(IRC_ERR_BAD_ARG, "irc_gather_sum: null");
    irc::launch(gather_sum_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, src, map, n, splits, split_stride, dst);
    return irc_check_launch("irc_gather_sum");
}

extern "C" int irc_gather_sum_multi(const irc_sum_job* jobs_dev, int njobs, long long total, void* stream) {
    if (!jobs_dev || njobs <= 0 || total <= 0) return irc_set_error(IRC_ERR_BAD_ARG, "irc_gather_sum_multi: empty job table");
    irc::launch(gather_sum_multi_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, jobs_dev, njobs, total);
    return irc_check_launch("irc_gather_sum_multi");
}

extern "C" int irc_gather_f32(const float* src, const int* map, long long n, float* dst, void* stream) {
    if (!src || !map || !dst) return irc_set_error(IRC_ERR_BAD_ARG, "irc_gather_f32: null");
    irc::launch(gather_f32_kernel, grid_for(n, 256), 256, 0, (cudaStream_t)stream, src, map, n, dst);
    return irc_check_launch("irc_gather_f32");
}

/* Structured packing of standard convolution layers (see pack_std_kernel): jobs = device array of `njobs` records
 * {long long src, dst_f, dst_d; int N, K, T, kd, first_block, nblocks} (element offsets into `arena` / `packed`; K % 64 == 0,
 * N % 8 == 0, T <= 16; first_block = running sum of nblocks = ceil(N / 16) * (K / 64)); total_blocks = their sum. */
extern "C" int irc_pack_std(const float* arena, void* packed, const void* jobs, int njobs, int total_blocks, int max_taps, void* stream) {
    if (!arena || !packed || !jobs || njobs <= 0 || total_blocks <= 0 || max_taps <= 0 || max_taps > 16)
        return irc_set_error(IRC_ERR_BAD_ARG, "irc_pack_std: bad args");
    const size_t smem = (size_t)16 * (64 * max_taps + 1) * sizeof(float);
    static size_t attr = 48 * 1024;
    if (smem > attr) { cudaFuncSetAttribute(pack_std_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024); attr = 80 * 1024; }
    irc::launch(pack_std_kernel, total_blocks, 256, smem, (cudaStream_t)stream, arena, (bf16*)packed, (const PackStdJob*)jobs, njobs);
    return irc_check_launch("irc_pack_std");
}
