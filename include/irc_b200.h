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
 * Epilogue: +bias, * (mask>0 ? 1 : mask_slope), +addend, activation, rows with row_img<0 zeroed. */
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
    const void* addend;     /* bf16 [a_rows][addend_ld] or NULL: added to the accumulator before the activation */
    long long addend_ld;
    int addend_chan_off;
    int bn;                 /* tile width, 0 = auto */
    int mt;                 /* 128-row sub-tiles per tile sharing each weight stage: 1, 2, 0 = auto */
    int reuse;              /* taps with consecutive shifts: 1 = share one staged A box (tap runs), 2 = pack the three taps of a kernel
                             * row along N with a shifted sum in the epilogue (64 outputs, 3 x 3), 0 = neither, -1 = auto */
    int epilogue_direct;    /* 1 = store rows straight from registers (default 0: swizzled smem staging + TMA stores) */
    /* InstanceNorm statistics from the epilogue (nn.InstanceNorm2d after the convolution, irc:161): per 128-row sub-tile and
     * output channel the (sum, sum of squares) of the stored bf16 values; needs row_img (ring rows are stored as zeros) and
     * bf16 output with a tile width that is a multiple of 64.  stats_part: irc_conv_stats_workspace_floats(a_rows, n_out)
     * floats; stats_edge: (n_img + 1) * n_out * 2 floats; rows_per_img = hp * wp of the output frame.  NULL = off.
     * irc_conv_stats_finalize turns the partials into stats[n][c] = (sum, sum of squares) in a fixed order. */
    float* stats_part;
    float* stats_edge;
    int rows_per_img;
    /* Horizontal tap reduction + bias + tanh fused into the epilogue, for k x k convolutions with a tiny Cout computed as a
     * GEMM over the k kernel ROWS (n_out = 32 columns = (kernel column j, output channel co) -> j * tap_nco + co):
     *   tap_out[n][co][y][x] = act(bias[co] + sum_j acc[q + j - (tap_nshift-1)/2][j * tap_nco + co]),  q = frame row of (n, y, x)
     * fp32 NCHW out (the generator's output head, irc:527-531: 7 x 7 reflect conv 64 -> 3 + tanh).  Tiles overlap by
     * tap_nshift - 1 rows; `out` is not written.  tap_act: 0 none, 3 tanh.  NULL = off.
     * The same mode computes the data gradient of a k x k convolution with a tiny Cin (VGG conv1_1, irc:664 under irc:1680):
     * tap_scale (fp32 [tap_nco] or NULL) multiplies the result per output plane and tap_accumulate != 0 adds it to tap_out. */
    float* tap_out;
    int tap_nshift, tap_nco, tap_H, tap_W, tap_hp, tap_wp, tap_oy, tap_ox, tap_act;
    const float* tap_scale;
    int tap_accumulate;
    /* cin == 64 only: operand columns [k_live, 64) of `a` are structural zeros (e.g. the tap-expanded gradient of the 3-channel output
     * head: 21 live columns), so only ceil(k_live / 16) of the four 16-column reduction steps are issued.  0 = all columns. */
    int k_live;
} irc_conv_gemm_args;
int irc_conv_gemm(const irc_conv_gemm_args* args, void* stream);
/* nn.ConvTranspose2d(cin, cout, 3, stride=2, padding=1, output_padding=1) forward (the generator's up-sampling layers with
 * no_antialias_up=True, irc:495-499 / :512-516) on a framed input with a ZERO ring: one stride-1 implicit GEMM with the four
 * taps {0, 1, wp, wp + 1}; the 4 * cout output columns of input row q = (n, y, x) are the four sub-pixel phases of the output,
 * column (a * 2 + b) * cout + co = output pixel (2y + a, 2x + b) (depth-to-space order; read it back with an irc_view whose
 * s2d_c = cout).  w: bf16 [4 * cout][4 * cin] = (phase, co) x (tap dy * 2 + dx, ci) holding kernel element
 * (a + 1 - 2 dy, b + 1 - 2 dx) of the (cin, cout, 3, 3) PyTorch weight, or zero when that index leaves [0, 2];
 * bias4: fp32 [4 * cout] (bias repeated per phase) or NULL.  The data gradient is irc_conv_gemm with negated taps and the
 * transposed operand, the weight gradient irc_tn_gemm - exactly as for a forward convolution. */
int irc_convT2d_fwd(const void* x, long long rows, int x_ld, int x_chan_off, int cin, int wp, const void* w, int cout, const float* bias4,
                    void* out, long long out_ld, int out_chan_off, void* stream);
long long irc_conv_stats_workspace_floats(long long rows, int n_out);
int irc_conv_stats_finalize(const float* part, const float* edge, int n_img, int rows_per_img, int n_out, float* stats, void* stream);

/* Weight-gradient GEMM: out[s][t][m][n] = sum_{q in split s} a[q + a_shift[t]][a_chan_off + m] *
 * b[q + b_shift[t]][b_chan_off + n].  Replaces the conv weight-gradient of loss.backward()
 * (irc:1650, irc:1680).  Split partial sums are reduced by irc_gather_sum.  The launch uses
 * ceil(m/128) * ceil(n/bn) * ceil(ntaps/tpc) * splits CTAs; irc_tn_gemm_ctas() returns that count for `splits` = 1. */
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
    int tpc;                /* taps accumulated per CTA into independent TMEM tiles (1,2,3,4,8; tpc*bn <= 512), 0 = auto */
} irc_tn_gemm_args;
int irc_tn_gemm(const irc_tn_gemm_args* args, void* stream);
int irc_tn_gemm_ctas(int m, int n, int ntaps, int same_a_shift);


/* ---- memory-bound passes on NHWC bf16 frames ---------------------------------------- */

/* An NHWC bf16 image set inside a flat [rows][ld] buffer: pixel (n,y,x), channel c lives at
 * row (n*hp + y + oy)*wp + (x + ox), column chan_off + c (unless s2d_c > 0, see below). */
typedef struct irc_view {
    const void* ptr;
    long long ld;
    int chan_off, hp, wp, oy, ox;
    int s2d_c;              /* >0: the buffer holds 2x2 space-to-depth blocks of an image with s2d_c
                               channels (hp, wp count blocks; oy, ox offset the un-blocked pixel) */
} irc_view;

/* row_img[q] = image index for rows inside [y0,y1)x[x0,x1) of each hp x wp image, else -1. */
int irc_row_index(short* row_img, int n_img, int hp, int wp, int y0, int y1, int x0, int x1, void* stream);

/* InstanceNorm statistics (nn.InstanceNorm2d, irc:161): stats[n][c] = (sum, sum of squares)
 * over the H x W pixels of view z.  Reductions here are two-stage and order-fixed (bit-reproducible):
 * `work` (optional, work_floats floats, ZERO-INITIALISED by the caller once) holds the per-chunk partials and, in its
 * last 256 floats, self re-arming ticket counters: the last block of an image adds the partials in chunk order.
 * Without it one block per image runs. */
int irc_in_stats(const irc_view* z, int C, int n_img, int H, int W, float* stats, float* work, long long work_floats, void* stream);

/* Separable table gather into a frame:
 *   dst[n,Y,X,:] = halo( sum_ij ty_w[y][i] tx_w[x][j] * pre(src)[n, ty_idx[y][i], tx_idx[x][j], :] (+ src2[same]) + res[n,y,x,:] )
 * pre = InstanceNorm with `stats` (if given) followed by `act`.  NULL tables = identity.
 * With binomial / bilinear tables this is Downsample (irc:307-310) and UpsampleAA
 * (irc:350-355) fused with the preceding norm+activation; with identity tables it is the
 * InstanceNorm/ReLU/residual apply of ResnetBlock (irc:417-418); with fold tables it is the
 * backward of ReflectionPad2d.  halo_mode: 0 zero ring, 1 reflect ring of width `pad`.
 * dst_s2d: write the zero-padded image as 2x2 space-to-depth blocks (stride-2 conv operand). */
typedef struct irc_gather_args {
    irc_view src, src2, res, dst;
    int C, n_img;
    const float* stats; float cnt, eps; int act; float slope;
    const int* ty_idx; const float* ty_w; int ky;
    const int* tx_idx; const float* tx_w; int kx;
    int H, W, pad, halo_mode, dst_s2d;
    /* optional shared-memory tiling for real stencils: output tile tile_y x tile_x whose source bounding box is at
     * most patch_y x patch_x (the caller knows its tables); tile_y = 0: register-table kernel (one thread per output
     * vector, taps from L1/L2); tile_y = -2: generic fallback */
    int tile_y, tile_x, patch_y, patch_x;
} irc_gather_args;
int irc_gather(const irc_gather_args* args, void* stream);
/* InstanceNorm statistics + irc_gather's normalise / activate / residual / ring in ONE launch for small maps (H*W <= 4096,
 * C % 32 == 0, identity tables, one source, cnt == H*W): a thread-block cluster per (image, 32 channels) stages z once,
 * reduces through distributed shared memory and writes both stats_out[n][c] = (sum, sum of squares) and the frame.
 * args->stats is ignored. */
int irc_in_apply_fused(const irc_gather_args* args, float* stats_out, void* stream);

/* InstanceNorm(+activation) backward (autograd of irc:161 + ReLU/LeakyReLU inside
 * loss.backward()).  g = table gather of (g1 + g2); see elementwise.cu. */
typedef struct irc_in_bwd_args {
    irc_view z, g1, g2, dz;
    int C, n_img, H, W;
    const float* stats; float cnt, eps; int act; float slope;
    const int* ty_idx; const float* ty_w; int ky;
    const int* tx_idx; const float* tx_w; int kx;
    float* bsum;            /* [n][C][2] result of the reduce pass */
    float* work;            /* optional partials workspace of the reduce pass */
    long long work_floats;
} irc_in_bwd_args;
int irc_in_bwd_reduce(const irc_in_bwd_args* args, void* stream);
int irc_in_bwd_apply(const irc_in_bwd_args* args, void* stream);
/* reduce + apply in ONE launch for small maps (H*W <= 4096, C % 32 == 0, one source, no tables, stats given): a
 * thread-block cluster per (image, 32 channels) keeps g and z in registers and exchanges the sums through distributed
 * shared memory.  fold_pad > 0: g1 views a frame that holds the gradient w.r.t. ReflectionPad2d(fold_pad) of the map;
 * the ring pixels are folded onto the interior while loading (the ring itself is left untouched).  bsum is optional. */
int irc_in_bwd_fused(const irc_in_bwd_args* args, int fold_pad, void* stream);
/* Same result as irc_in_bwd_reduce + irc_in_bwd_apply in ONE launch for maps larger than the cluster kernel takes: the CTAs
 * form `groups` groups, each group walks its images one at a time (partial sums, a group barrier, apply), so that the second
 * read of g and z hits the L2 and HBM sees every tensor once (3 units instead of 5).  Identity tables, plain views,
 * C in {64, 128, 256}; args->work must hold 4 * groups * ctas_per_group * C floats + 512 (zero-initialised once). */
int irc_in_bwd_l2(const irc_in_bwd_args* args, int groups, void* stream);

/* nn.BatchNorm2d (get_norm_layer('batch'), irc:158-159) on the same kernels.  irc_bn_finalize turns the per-image (sum, sum of
 * squares) of irc_in_stats / the conv epilogue into effective moments eff[n][c] = (mean_eff, rs_eff) with
 * gamma * (x - mean_B) * rstd_B + beta == (x - mean_eff) * rs_eff - batch statistics over groups of `group` consecutive images
 * (one group per forward call of the reference), or the running statistics when training == 0 - and updates running_mean /
 * running_var (momentum, unbiased variance) `updates` times.  Pass eff as `stats` with eps < 0 to irc_gather,
 * irc_in_bwd_reduce and irc_in_bwd_apply (cnt = group * H * W).  irc_bn_bwd_fix goes between the reduce and the apply pass: it
 * sums the per-image partials of bsum over each group, writes the gradients of gamma and beta, and replaces bsum by the pair the
 * apply pass needs for the batch-norm input gradient. */
int irc_bn_finalize(const float* stats, int n_img, int group, int C, float cnt_per_img, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float momentum, float eps, int training, int updates, float* eff, void* stream);
int irc_bn_bwd_fix(float* bsum, int n_img, int group, int C, const float* gamma, const float* beta, float* dgamma, float* dbeta, int accumulate, void* stream);

/* Backward of nn.ReflectionPad2d(p) in place on a frame holding the gradient w.r.t. the padded tensor: interior pixels
 * within p of the border receive the ring pixels that mirror onto them; the ring is cleared. */
int irc_fold_inplace(void* g, long long ld, int chan_off, int C, int n_img, int H, int W, int p, void* stream);

/* nn.MaxPool2d(2,2) of the VGG trunk (torchvision vgg16.features[4], [9]; irc:664) and its
 * backward fused with the mask of the ReLU that precedes it. */
int irc_maxpool2(const irc_view* src, const irc_view* dst, int C, int n_img, int Ho, int Wo, void* stream);
int irc_maxpool2_bwd(const irc_view* src, const irc_view* g, const irc_view* dsrc, int C, int n_img, int Ho, int Wo, void* stream);

/* out[c] = sum over rows (with row_img[row] >= 0 when given) of a[row][chan_off + c]  (bias gradients) */
int irc_colsum(const void* a, long long rows, long long ld, int chan_off, int C, const short* row_img, float* out, float* work,
               long long work_floats, void* stream);

/* ---- degenerate convolutions (tiny K or tiny N) --------------------------------------- */

/* im2col of a small-channel fp32 NCHW input (optionally two tensors concatenated on the
 * channel axis, optionally a per-channel affine) into a bf16 [rows][64] operand, column
 * (r*k+s)*C + c.  Feeds inc (irc:458-463), D model.0 on cat[ir,rgb] (irc:600, :1639-1643) and
 * VGG conv1_1 incl. its input normalisation (irc:679-682) as one-tap GEMMs.
 * row_mode: 0 rows=(n,oy,ox); 1 framed with a one-pixel ring; 2 ring + 2x2 sub-pixel order
 * (so that the [rows][64] GEMM output is the space-to-depth operand of a stride-2 conv). */
typedef struct irc_im2col_args {
    const float* src1; int c1;
    const float* src2; int c2;
    const float* scale; const float* shift;
    int n_img, H, W, k, stride, pad, pad_mode, Ho, Wo, row_mode;
    void* dst;
    short* row_img;         /* optional: receives the live-row table of this row order */
} irc_im2col_args;
long long irc_im2col_rows(int row_mode, int n_img, int Ho, int Wo);
int irc_im2col(const irc_im2col_args* args, void* stream);

/* The same three layers WITHOUT the operand round trip (replaces nn.Conv2d of irc:458-463, :600, :664 outright): the block
 * stages the image rows in shared memory, gathers the 64-slot operand rows straight into mma.sync fragments and applies
 * out[q][0..64) = act(bias + sum_k E[q][k] * w[n][k]), rows outside the image stored as zeros.  w: bf16 [64][64] (slot order of
 * irc_im2col); out: bf16 [rows][64] in the row order of args->row_mode.  args->dst may be NULL; when given it ALSO receives the
 * operand E (kept for the weight gradient).  act: 0 none, 1 ReLU, 2 LeakyReLU(slope). */
int irc_smallk_conv_fwd(const irc_im2col_args* args, const void* w, const float* bias, int act, float slope, void* out, void* stream);

/* Transpose of irc_im2col (zero padding only): out[n][c][y][x] (+)= sum de[row][(r*k+s)*C+c]. */
typedef struct irc_col2im_args {
    const void* de; long long ld;
    int C, c_first, c_out, n_img, H, W, k, stride, pad, Ho, Wo, row_mode;
    const float* scale;
    float* out; int accumulate;
} irc_col2im_args;
int irc_col2im(const irc_col2im_args* args, void* stream);

/* Shifted tap reduction / expansion around a GEMM over one kernel axis (shift_j = dy[j]*wp + dx[j] rows; the
 * padding ring of the frame must be at least max|dx| wide):
 *   reduce: out[n][co][y][x] = act(bias[co] + sum_j P[q(n,y,x) + shift_j][j*nco + co]), act 3 = tanh
 *   expand: E[q][j*nco + co] = g[pixel(q - shift_j)][co] * (1 - y^2 if y given); dbias[co] = sum g'
 * outc + tanh (irc:527-531) and D model.11 (irc:629), forward and backward. */
typedef struct irc_tap_args {
    int nshift, nco;
    int dy[IRC_MAX_TAPS], dx[IRC_MAX_TAPS];
    int n_img, H, W, hp, wp, oy, ox;
    int live_cols_only;     /* expand: write only the 32-byte column sectors that hold tap columns (zero-filled up to the sector end);
                             * the caller keeps the other columns zero */
} irc_tap_args;
int irc_tap_reduce(const irc_tap_args* t, const float* P, long long ldp, const float* bias, int act, float* out, void* stream);
int irc_tap_expand(const irc_tap_args* t, const float* g, const float* y, void* E, float* dbias, float* work, long long work_floats,
                   void* stream);

/* ---- losses (fp32 NCHW) --------------------------------------------------------------- */

/* One pass over fake/target: sums[0..2] += (sum|f-t|, sum|d_rows f|, sum|d_cols f|) and
 * dfake = w_l1*sign(f-t) + TV sign stencils scaled by w_tvv / w_tvh.  nn.L1Loss (irc:1664) and
 * tv_loss (irc:686-694) with their autograd.  target NULL = TV only; dfake NULL = values only. */
int irc_pixel_loss(const float* fake, const float* target, int n_img, int C, int H, int W, float w_l1, float w_tvv, float w_tvh,
                   float* sums, float* dfake, void* stream);

/* ssim_loss_torch (irc:714-750): 11-tap Gaussian separable window, zero padding 5.  Inputs are
 * mapped x = img*scale + shift first (irc:1675-1676 uses (x+1)/2).  fwd: sums[n] += sum of the
 * SSIM map of image n, and (if ga != NULL) the maps dS/dmu1, dS/dE[x^2], dS/dE[xy]; bwd:
 * dimg1 (+)= coef*scale*(w*ga + 2x w*gb + y w*gc).  window11 is a HOST array of the 11 taps. */
int irc_ssim_fwd(const float* img1, const float* img2, int n_img, int C, int H, int W, float scale, float shift, const float* window11,
                 float* sums, float* ga, float* gb, float* gc, void* stream);
int irc_ssim_bwd(const float* img1, const float* img2, int n_img, int C, int H, int W, float scale, float shift, const float* window11,
                 const float* ga, const float* gb, const float* gc, float coef, float* dimg1, int accumulate, void* stream);

/* Hinge discriminator loss (irc:1647-1649; mode 0) and generator GAN term (irc:1662; mode 1)
 * with their gradients w.r.t. the score maps. */
int irc_hinge(const float* pred, long long n_total, long long n_real, int mode, float w_real, float w_fake, float* sums, float* dpred, void* stream);

/* F.l1_loss(vgg(fake), vgg(rgb)) (irc:1667-1669) on the bf16 feature frame holding both
 * halves; writes the gradient w.r.t. the pre-ReLU conv3_3 output of the fake half. */
int irc_feat_l1(const void* feat, long long rows_half, long long ld, int C, float w, float* sums, void* dz, long long ld_dz, void* stream);

/* tensor_to_rgb_image + compute_metrics core (irc:865-876, irc:1197-1205), batched on device:
 * u8[n][y][x][c] = trunc(clip((fake+1)/2,0,1)*255); sums[n] = (sum|u8/255-gt|, sum(u8/255-gt)^2). */
int irc_quantize_metrics(const float* fake, const float* gt, int n_img, int C, int H, int W, unsigned char* u8, double* sums, void* stream);

/* The SSIM column of compute_metrics (irc:1208-1215): skimage.metrics.structural_similarity(gt, pred, data_range=1.0,
 * channel_axis=2) with its defaults (7 x 7 uniform window, sample covariance, K1 = 0.01, K2 = 0.03, float64, 3-pixel crop),
 * batched on device.  u8: [n][H][W][3] predictions (irc_quantize_metrics), gt: fp32 [n][3][H][W] in [0, 1];
 * sums[n] = sum over the 3 channels and the (H - 6) x (W - 6) valid positions of the SSIM map (divide by their count). */
int irc_ssim_metric(const unsigned char* u8, const float* gt, int n_img, int H, int W, double* sums, void* stream);

/* Running loss sums on the device (irc:1683-1697 averages every step of an epoch; this removes the per-step .item()
 * synchronisation): acc[j] += coef[j][n] + sum_i coef[j][i] * s[i], coef is [rows][n+1] fp32, rows <= 32. */
int irc_accumulate(const float* s, int n, const float* coef, int rows, double* acc, void* stream);

/* ---- input pipeline (after the image file has been decoded) ---------------------------------------------------------- */

/* cv2.resize(src, (Wd, Hd), interpolation=cv2.INTER_AREA) of n_img 8-bit frames [n][Hs][Ws][C] (C = 1..4), bit-exact with
 * OpenCV's implementation (KAISTPairDataset._read_ir / _read_rgb, irc:1132-1158; load_ir_image / load_rgb_image, irc:803-852).
 * mode 0: general scale factors, separable tables (xi/xw: [Wd][kx], yi/yw: [Hd][ky] from computeResizeAreaTab, weight 0 =
 * unused slot); mode 1: integer factors; mode 2: exactly 2 x 2.  img_max (optional, [n_img]) receives the maximum byte of
 * each resized frame (the reference divides an IR frame by 255 only if its maximum exceeds 1, irc:1142). */
int irc_resize_area_u8(const unsigned char* src, int n_img, int Hs, int Ws, int C, int Hd, int Wd, const int* xi, const float* xw, int kx,
                       const int* yi, const float* yw, int ky, int mode, unsigned char* dst, int* img_max, void* stream);
/* uint8 [n][H][W][C] -> fp32 [n][C][H][W] in [-1, 1]: x / 255 (skipped for images with img_max <= 1 when img_max is given), clip,
 * channel swap for BGR frames, per-image horizontal flip (the paired augmentation of irc:1166-1168), x * 2 - 1 (irc:1174-1175). */
int irc_u8_to_pm1(const unsigned char* src, int n_img, int H, int W, int C, int swap_rb, const unsigned char* flip, const int* img_max, float* out,
                  void* stream);

/* ---- optimizer / parameter layout ------------------------------------------------------ */

/* torch.optim.Adam step (irc:1651, :1681; defaults of irc:1601-1604) over a flat arena.
 * hyper (device, fp64) = {lr, beta1, beta2, eps, lr_scale, grad_scale}: the step uses lr*lr_scale (LambdaLR factor,
 * irc:1606-1609) and g*grad_scale (1/world_size when the gradients were sum-all-reduced).
 * step_count (device): optimizer steps taken so far; the bias corrections use *step_count + 1 and a second tiny launch
 * on the same stream advances it, so a captured CUDA graph replays correctly without host writes. */
int irc_adam(float* p, const float* g, float* m, float* v, long long n, const double* hyper, long long* step_count, void* stream);
/* dst[i] = map[i] >= 0 ? src[map[i]] : 0 in fp32 (per-phase copies of a transposed-conv bias for the GEMM epilogue). */
int irc_gather_f32(const float* src, const int* map, long long n, float* dst, void* stream);
/* dst[i] = bf16(map[i] >= 0 ? src[map[i]] : 0): OIHW fp32 parameters -> packed GEMM operands. */
int irc_pack_bf16(const float* src, const int* map, long long n, void* dst, void* stream);
/* The same for the regular layers without the index map: OIHW fp32 -> W_f[n][t*K + k] and the data-gradient operand
 * W_d[k][t*kd + n] (bf16) through a shared-memory transpose, all layers of a network in one launch.  jobs: device array of
 * {long long src, dst_f, dst_d; int N, K, T, kd, first_block, nblocks} (element offsets into arena / packed; K % 64 == 0,
 * N % 16 == 0, T <= 16; nblocks = N / 16 * K / 64, first_block = their running sum); max_taps = the largest T. */
int irc_pack_std(const float* arena, void* packed, const void* jobs, int njobs, int total_blocks, int max_taps, void* stream);
/* dst[i] = sum_s src[s*split_stride + map[i]]: split weight-gradient partials -> OIHW fp32. */
int irc_gather_sum(const float* src, const int* map, long long n, int splits, long long split_stride, float* dst, void* stream);
/* The same for a table of jobs in ONE launch (all weight gradients of a network, irc:1650 / :1680).  `jobs_dev` is a DEVICE
 * array of njobs entries whose `start` fields are the exclusive prefix sums of `n`; total = sum of n. */
typedef struct irc_sum_job {
    const float* src; const int* map; float* dst;
    long long n, split_stride, start;
    int splits, pad_;
} irc_sum_job;
int irc_gather_sum_multi(const irc_sum_job* jobs_dev, int njobs, long long total, void* stream);

/* ---- stand-alone anti-aliased resampling on fp32 NCHW ---------------------------------- */

/* out[p][Y][X] (+)= sum_ij ty_w[Y][i] tx_w[X][j] in[p][ty_idx[Y][i]][tx_idx[X][j]] over `planes` = N*C planes.
 * With the binomial stride-2 tables this is Downsample.forward (irc:307-310); with blur*bilinear tables
 * UpsampleAA.forward (irc:350-355); with the transposed tables their backward passes. */
int irc_stencil_nchw(const float* in, float* out, int planes, int Hi, int Wi, int Ho, int Wo, const int* ty_idx, const float* ty_w, int ky,
                     const int* tx_idx, const float* tx_w, int kx, int accumulate, void* stream);

/* Same contract, streaming kernel (rolling register window over the source rows, input and output touched once in
 * DRAM).  Requires tables whose last contributing source row is non-decreasing in Y and whose rows span at most
 * `window` <= 8 consecutive source rows / x-entries: the binomial and blur*bilinear tables and their transposes do. */
int irc_stencil_nchw_stream(const float* in, float* out, int planes, int Hi, int Wi, int Ho, int Wo, const int* ty_idx, const float* ty_w,
                            int ky, const int* tx_idx, const float* tx_w, int kx, int window, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif
