/* C ABI of libirc_sm100.so — the B200-native kernels under the IR-colorization hot path.
 *
 * The reference (Code/ir_colorization.py, cited irc:LINE) has no FFI: its hot path is a
 * chain of PyTorch library calls.  Each entry point below replaces one family of those
 * calls; the comment on each says which.  Conventions:
 *   - plain device pointers + extents + a cudaStream_t passed as void*; no torch types;
 *   - every function returns 0 on success, non-zero IRC_ERR_* otherwise; the message is
 *     available from irc_last_error(); nothing throws, allocates or synchronises;
 *   - there is no CPU path: irc_arch_check() fails on anything but sm_100.
 *
 * Activation buffers are NHWC bf16 "frames": [N][H+2p][W+2p][C] with the padding ring
 * stored, viewed as a flat [rows][C] matrix, so that a stride-1 convolution tap is a
 * constant row shift and one 2-D TMA box fetches the operand of any tap.
 */
#ifndef IRC_B200_H
#define IRC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define IRC_MAX_TAPS 64

/* build identification / device gate */
int irc_version(void);
int irc_arch_check(void);              /* 0 iff the current device is sm_100 */
const char* irc_last_error(void);

/* Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators,
 * TMA-fed).  out[q][n] = epi( sum_t sum_c a[q + taps[t]][a_chan_off + c] * w[n][t*cin + c] ).
 * Replaces nn.Conv2d forward (irc:458-531 generator, irc:598-630 discriminator, irc:664 VGG
 * trunk) and, with negated taps and transposed weights, the conv data-gradient that
 * loss.backward() (irc:1650, irc:1680) runs through cuDNN/oneDNN.
 * Epilogue: +bias, * (mask>0 ? 1 : mask_slope), activation, rows with row_img<0 zeroed. */
typedef struct irc_conv_gemm_args {
    const void* a;          /* bf16 [a_rows][a_ld] */
    long long a_rows;
    int a_ld, a_chan_off, cin;      /* cin % 64 == 0 */
    int ntaps;
    int taps[IRC_MAX_TAPS];         /* row shift of each tap (may be negative) */
    const void* w;          /* bf16 [n_out][ntaps*cin] */
    int n_out;              /* % 32 == 0 */
    void* out;              /* bf16 or fp32 [a_rows][out_ld] */
    long long out_ld;
    int out_chan_off, out_fp32;
    const float* bias;      /* [n_out] or NULL */
    int act;                /* 0 none, 1 ReLU, 2 LeakyReLU(slope) */
    float slope;
    const short* row_img;   /* [a_rows] image index of each row, <0 = padding ring (written as 0); NULL = all live */
    const void* mask;       /* bf16 [a_rows][mask_ld] or NULL */
    long long mask_ld;
    int mask_chan_off;
    float mask_slope;
    int bn;                 /* tile width, 0 = auto */
} irc_conv_gemm_args;
int irc_conv_gemm(const irc_conv_gemm_args* args, void* stream);

/* Weight-gradient GEMM: out[s][t][m][n] = sum_{q in split s} a[q + a_shift[t]][a_chan_off + m] *
 * b[q + b_shift[t]][b_chan_off + n].  Replaces the conv weight-gradient of loss.backward()
 * (irc:1650, irc:1680).  Split partial sums are reduced by irc_gather_sum. */
typedef struct irc_tn_gemm_args {
    const void* a; long long a_rows; int a_ld, a_chan_off, m;
    const void* b; long long b_rows; int b_ld, b_chan_off, n;
    long long k_rows;
    int ntaps;
    int a_shift[IRC_MAX_TAPS];
    int b_shift[IRC_MAX_TAPS];
    float* out;
    long long out_tap_stride, out_m_stride, out_n_stride, out_split_stride;
    int splits;
    int bn;                 /* tile width (multiple of 64), 0 = auto */
} irc_tn_gemm_args;
int irc_tn_gemm(const irc_tn_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif
