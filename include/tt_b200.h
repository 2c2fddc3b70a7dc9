/*
 * tt_b200.h — C ABI of libtt_b200.so: the B200-native (sm_100a) operator library for the
 * ThinkTwice per-frame forward path.  It replaces `open_loop_training/ops` of the reference
 * (OpenDriveLab/ThinkTwice) and the third-party native ops the path crosses (mmcv._ext
 * ms_deform_attn / deform_conv / hard_voxelize, spconv, cuDNN/cuBLAS through torch.nn).
 *
 * Conventions (SURVEY.md §8b "External native ABIs"):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says "host";
 *   - activations are fp32 channels-last: a feature map is [N][H][W][ld] floats, the logical
 *     channels of a tensor live at [coff, coff + C) inside the pixel's `ld` floats, so a concat
 *     is just several producers writing into one buffer;
 *   - every call is asynchronous on `stream` (a cudaStream_t), takes no locks, keeps no global
 *     state except the last-error string, and NEVER calls exit(): 0 = ok, negative = tt_status;
 *   - the caller owns and allocates all buffers.
 *
 * Each entry cites the reference interface it replaces (paths relative to
 * /root/reference/open_loop_training).
 */
#ifndef TT_B200_H_
#define TT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tt_stream_t; /* cudaStream_t */

typedef enum { TT_OK = 0, TT_ERR_INVALID = -1, TT_ERR_CUDA = -2, TT_ERR_UNSUPPORTED = -3 } tt_status;
enum { TT_ACT_NONE = 0, TT_ACT_RELU = 1, TT_ACT_GELU = 2, TT_ACT_SIGMOID = 3, TT_ACT_SOFTPLUS = 4,
       TT_ACT_SOFTPLUS_CLAMP = 5 /* clamp(softplus(x), min=1e-3): thinktwice_decoder.py:484 */ };
enum { TT_RES_NONE = 0, TT_RES_SAME = 1, TT_RES_UP2_NEAREST = 2 };

int tt_version(void);
const char* tt_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long tt_launch_count(void);
/* diagnosis knobs for kernel micro-benchmarks (tools/conv_bench.py); 0 in normal operation */
void tt_debug_set(int flags);

/* ------------------------------------------------------------------------------------------
 * (1) voxel pooling — drop-in for the reference's only in-tree native op.
 * Replaces: ops/voxel_pooling/src/voxel_pooling_forward.cpp:21-22 and
 *           src/voxel_pooling_forward_cuda.cu:38-42 `voxel_pooling_forward_kernel_launcher`.
 * Same argument list and ownership rules (caller zero-fills output_features [B][Y][X][C] and
 * fills pos_memo [B][P][3] with -1; the op accumulates in place) but returns a status instead
 * of exiting.  geom_xyz int32 [B][P][3] (x, y, z); input_features fp32 [B][P][C].
 * Like the reference it accumulates with fp32 atomics (sum order unspecified), but with coalesced
 * 128-bit feature loads and one 128-bit vector reduction per 4 channels.  The model itself uses
 * tt_lift_splat below, which has no floating-point atomics.
 * `workspace` must hold tt_voxel_pooling_workspace_bytes(...) bytes (may be NULL when that is 0).
 */
size_t tt_voxel_pooling_workspace_bytes(int batch_size, int num_points, int num_voxel_x, int num_voxel_y);
int tt_voxel_pooling_forward(int batch_size, int num_points, int num_channels, int num_voxel_x, int num_voxel_y,
                             int num_voxel_z, const int* geom_xyz, const float* input_features,
                             float* output_features, int* pos_memo, void* workspace, tt_stream_t stream);

/* Fused lift-splat: lss.py:582-632 (depth softmax, outer product with the context features,
 * frustum geometry, voxel index, scatter-add) without the (B,N,D,H,W,C) intermediate.
 *   depth_logits : [BN][fH][fW][ld_d] channels-last, D logits at [d_coff, d_coff + D)
 *   context      : [BN][fH][fW][ld_c] channels-last, C features at [c_coff, c_coff + C)
 *   mats         : [BN][2][16] fp32 row-major 4x4: ida^-1 and sensor2ego*intrin^-1 (lss.py:496,502)
 *   frustum_u [fW], frustum_v [fH], frustum_d [D]: the image-plane / depth coordinates of the
 *                  reference's `frustum` buffer (lss.py:454-471), passed in so both sides use the same floats
 *   bev          : [B][Y][X][bev_ld] (+bev_coff), OVERWRITTEN (not accumulated); if `anti_transpose`
 *                  the rot90(flip) of encoder_decoder_framework.py:241 is folded into the store.
 * voxel index = trunc((p - lower) / size) exactly as lss.py:630-631 (`.int()` truncates toward 0).
 */
typedef struct {
  int B, N, D, fH, fW, C;
  int ld_d, d_coff, ld_c, c_coff;
  float lower[3], size[3];       /* lower = (voxel_coord - voxel_size/2) as computed in fp32 by lss.py:630 */
  int X, Y, Z;
  int bev_ld, bev_coff;
  int anti_transpose;
} tt_lift_splat_desc;
size_t tt_lift_splat_workspace_bytes(const tt_lift_splat_desc* d);
int tt_lift_splat(const tt_lift_splat_desc* d, const float* depth_logits, const float* context, const float* mats,
                  const float* frustum_u, const float* frustum_v, const float* frustum_d, float* bev, void* workspace,
                  tt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (2) implicit-GEMM convolution / linear layer.
 * Replaces every torch.nn.Conv2d / ConvTranspose2d(k2s2) / Linear on the path (cuDNN/cuBLAS via
 * torch) and, with `gather`, spconv SubMConv3d / SparseConv3d (N5).
 *   y[pix(m)][y_coff + n] = act( sum_k A[m][k] * w[k][n] + bias[n] + res[...] + res2[...] )
 *   A[m][k], k = tap * Cin_g + c : the input pixel of output row m under tap (kh, kw), channel
 *   g*Cin_g + c, zero outside the image; or, when gather != NULL, row gather[m*taps + tap] of x
 *   (-1 = zero row) — the sparse-conv rulebook.
 *   w: impl 0/1: packed [taps * Cin_g][Cout] fp32 (BatchNorm already folded in); impl 2/3 (tcgen05): packed
 *      [2][Cout][taps][Cin] = `hi = RN_tf32(w)` plane followed by the `lo = RN_tf32(w - hi)` plane.  bias [Cout] or NULL.
 * Output row m = (n, oh, ow) is stored at pixel (n, oh*oy_mul + oy_add, ow*ox_mul + ox_add) of a
 * [N][yH][yW][y_ld] buffer (a k2s2 transposed conv is four such calls).
 * If m_count != NULL only the first *m_count rows (a device int) are computed.
 * `impl`: 0/1 = SIMT fp32 FFMA (exact), 2 = tcgen05 single-pass TF32 (operands rounded to nearest TF32),
 *         3 = tcgen05 3xTF32 (hi*hi + hi*lo + lo*hi, fp32-class); 4 = scaled-split fp16 through tt_conv2d_f16s.  impl 2/3 need stride 1 or 2, groups 1, channels,
 *         x_ld, x_coff % 4 == 0 and 16-byte aligned x / w / y (TMA reads the activations in place and splits them in
 *         shared memory; no workspace); else TT_ERR_UNSUPPORTED.
 *         For impl 0/1 the workspace is optional: when given (size from tt_conv2d_workspace_bytes) small-M / large-K
 *         layers run split-K with a deterministic fixed-order reduction.
 */
typedef struct {
  int N, H, W, Cin, x_ld, x_coff;
  long long x_nstride, y_nstride;  /* floats between consecutive images; 0 = dense (H*W*x_ld, yH*yW*y_ld) */
  long long x_hstride;             /* floats between consecutive input rows; 0 = dense (W*x_ld).  With x_ld < Cin the
                                    * "channels" of a pixel run on into the next pixels of the row: a KW-wide tap row
                                    * of a thin-channel conv becomes ONE K slab (row-packed stem conv, see lss.py) */
  int Cout, KH, KW, stride, pad, dil, groups;
  int OH, OW;
  int y_ld, y_coff, yH, yW, oy_mul, oy_add, ox_mul, ox_add;
  int act;
  int bias_n_mod;                 /* 0: bias[Cout]; k > 0: per-image bias table, row (n % k) of bias[k][Cout] */
  int res_mode, res_ld, res_coff; /* res: same pixel grid as the output rows (SAME) or coarser (UP2_NEAREST) */
  int res_H, res_W;               /* UP2_NEAREST: residual map size; source pixel = floor(o * res / O) */
  int res2_ld, res2_coff;         /* optional second SAME residual */
  int taps;                       /* gather mode: taps per row (KH=KW=1 then); else ignored */
  int M;                          /* gather mode: row capacity; else ignored (N*OH*OW) */
  int impl;
} tt_conv_desc;
size_t tt_conv2d_workspace_bytes(const tt_conv_desc* d);
int tt_conv2d(const tt_conv_desc* d, const float* x, const float* w, const float* bias, const float* res,
              const float* res2, const int* gather, const int* m_count, float* y, void* workspace, tt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (2b) the same contraction on SCALED-SPLIT FP16 operands (impl 4, csrc/gemm_conv_f16s.cu) — the default tensor-core
 * engine.  Replaces the same reference calls as (2): every torch.nn.Conv2d / ConvTranspose2d / Linear on the path
 * (lss.py, encoder_decoder_framework.py:81-138, thinktwice_decoder.py) and spconv's convolutions (lidarnet.py:42-52).
 * A "split" tensor is two __half planes over the SAME [N][H][W][ld] element grid as its fp32 twin:
 *   hi = RN_f16(x),  lo' = RN_f16((x - hi) * 2048)   =>   x = hi + lo' / 2048 to 2^-22 relative (|x| < 65504).
 * `x_split` / `y_split` point at the hi plane (already advanced to the tensor's channel offset); the lo' plane lies
 * `*_plane` halves further.  Split activations are written by the producing kernel's epilogue (`y_split` here), so a
 * consumer's TMA loads are tensor-core ready; tt_split_f16 converts fp32 rows produced by other kernels.
 *   w_split: [2][Cout][taps][Cin8] halves, Cin8 = Cin rounded up to 8 (zero padded), plane 0 = hi, plane 1 = lo'.
 *   products: hi*hi + (hi*lo' + lo'*hi) / 2048, fp32 accumulation drained into registers every 128 K elements.
 *   y (fp32) and y_split are both optional outputs (at least one) — a tensor only convolutions read needs no fp32 copy;
 *   each residual is given EITHER as fp32 (res / res2) or as split planes (res_split / res2_split, same element offsets).
 * Needs groups 1, stride 1|2, x_ld % 8 == 0 (halves), y_ld % 4 == 0, 16-byte aligned pointers, else TT_ERR_UNSUPPORTED.
 */
typedef struct {
  const void* x_split; long long x_plane;     /* input: hi plane pointer (at the channel offset), halves to the lo' plane */
  const void* w_split;                        /* [2][Cout][taps][Cin8] */
  const float* bias;                          /* fp32 [Cout] (or table, bias_n_mod) or NULL */
  const float* res;  const void* res_split;  long long res_plane;
  const float* res2; const void* res2_split; long long res2_plane;
  float* y; void* y_split; long long y_plane; /* outputs */
} tt_f16s_io;
int tt_conv2d_f16s(const tt_conv_desc* d, const tt_f16s_io* io, tt_stream_t stream);
/* Thin 1x1 stride-1 convolution (Cin <= 64, Cout <= 64, no residual, contiguous pixels) as a per-pixel fp32 mat-vec: one thread per
 * pixel reads its channel row from the split planes `io->x_split`, multiplies with the fp32 K-major weights `w_kmajor` [Cin][Cout] (the
 * tt_conv2d impl 0/1 layout) and writes fp32 rows and / or planes.  For maps so large and channels so few that 128 x 64 tensor-core tiles
 * are all overhead (the segmentation head at 224 x 448). */
int tt_pointwise_f16s(const tt_conv_desc* d, const tt_f16s_io* io, const float* w_kmajor, tt_stream_t stream);
/* fp32 rows [rows][x_ld] (first `cols` columns) -> split planes [2][rows][y_ld]; only the first *row_count rows when given */
int tt_split_f16(const float* x, long long x_ld, void* y_split, long long y_plane, long long y_ld, long long rows, int cols,
                 const int* row_count, tt_stream_t stream);
/* NCHW fp32 images (C <= 8) -> zero-bordered channels-last split planes [2][N][out_H][out_W][8 halves] at pixel offset (top, left): the
 * input layout of the row-packed stem convolution on the f16s engine (replaces tt_nchw_to_nhwc_padded + tt_split_f16; the border
 * is never written: zero the buffer once). */
int tt_image_to_split8(const float* x, void* y_split, long long y_plane, int N, int C, int H, int W, int out_H, int out_W, int top,
                       int left, tt_stream_t stream);
/* tt_upsample2x_bilinear_ac (lss.py:267) with the result written as split planes [2][N][2H][2W][C] only */
int tt_upsample2x_bilinear_ac_split(const float* x, void* y_split, long long y_plane, int N, int H, int W, int C, tt_stream_t stream);
/* split planes -> fp32 rows (hi + lo' / 2048): for a non-convolution kernel that must read a tensor stored as planes only */
int tt_merge_f16(const void* x_split, long long x_plane, long long x_ld, float* y, long long y_ld, long long rows, int cols,
                 tt_stream_t stream);
/* epilogue threads that had to clamp a value to +-65504 since the last reset (host int; synchronises `stream`) */
int tt_f16s_saturation_count(unsigned int* out_host, int reset, tt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (3) memory-bound helpers (torch elementwise / pooling / interpolation ops)
 */
/* NCHW fp32 -> channels-last with `ld` floats per pixel (channels >= C are zero-filled up to cpad) */
/* same, into a larger [N][out_H][out_W][y_ld] buffer at pixel offset (top, left): physical zero padding for
 * row-packed convolutions (the border is never written: zero it once). */
int tt_nchw_to_nhwc_padded(const float* x, float* y, int N, int C, int H, int W, int y_ld, int cpad, int out_H,
                           int out_W, int top, int left, tt_stream_t stream);
int tt_nchw_to_nhwc(const float* x, float* y, int N, int C, int H, int W, int y_ld, int y_coff, int cpad,
                    tt_stream_t stream);
int tt_nhwc_to_nchw(const float* x, int x_ld, int x_coff, float* y, int N, int C, int H, int W, tt_stream_t stream);
/* F.max_pool2d(x, 3, stride 2, padding 1) — mmdet ResNet stem */
int tt_maxpool3x3s2(const float* x, float* y, int N, int H, int W, int C, tt_stream_t stream);
/* nn.Upsample(scale_factor=2, bilinear, align_corners=True) — lss.py:267 */
int tt_upsample2x_bilinear_ac(const float* x, float* y, int N, int H, int W, int C, tt_stream_t stream);
/* mean over H*W -> [N][C] — lss.py:80 AdaptiveAvgPool2d */
int tt_global_avgpool(const float* x, int x_ld, int x_coff, float* y, int N, int HW, int C, tt_stream_t stream);
/* y[n][p][y_coff + c] = v[n][c] (bilinear resize of a 1x1 map == broadcast, lss.py:100-103) */
int tt_broadcast_rows(const float* v, float* y, int N, int HW, int C, int y_ld, int y_coff, tt_stream_t stream);
/* y = x * sigmoid(g[n][c]) — SELayer, lss.py:158 (g holds conv_expand output, pre-sigmoid) */
int tt_se_gate(const float* x, const float* g, float* y, int N, int HW, int C, tt_stream_t stream);
/* s[n][c] = 0.5*mean_hw + 0.5*max_hw — code/utils.py:90-91 */
int tt_se_pool(const float* x, float* s, int N, int HW, int C, tt_stream_t stream);
/* y = relu(x * sigmoid(g[n][c]) + shortcut) — code/utils.py:96,118-120 */
int tt_se_apply(const float* x, const float* g, const float* shortcut, int sc_ld, int sc_coff, float* y, int y_ld,
                int y_coff, int N, int HW, int C, tt_stream_t stream);
/* out[i][j] = x[H-1-j][W-1-i] per image/channel: rot90(flip(x,[2]),1,[2,3]) — framework:241,246 */
int tt_anti_transpose(const float* x, float* y, int N, int S, int C, tt_stream_t stream);
/* dst[r][c] = src[src_row(r)][c], src_row = (r / rdiv) % rmod; strides in floats */
int tt_copy2d(const float* src, int src_ld, float* dst, int dst_ld, int rows, int cols, int rdiv, int rmod,
              tt_stream_t stream);
/* LayerNorm over the last dim (eps 1e-5), rows x D, output stride y_ld */
int tt_layernorm(const float* x, int x_ld, const float* gamma, const float* beta, float* y, int y_ld, int rows,
                 int D, const int* row_count, tt_stream_t stream);
/* elementwise: op 0: y = a + b; 1: y = (1 - a) * b; 2: y = (1 - a) * b + a * c; 3: y = act(a) */
int tt_eltwise(int op, int act, const float* a, int a_ld, const float* b, int b_ld, const float* c, int c_ld, float* y,
               int y_ld, int rows, int cols, tt_stream_t stream);
int tt_fill(float* y, float v, long long n, tt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (4) deformable convolution v1 — mmcv DeformConv2dPack (lss.py:189-197), 3x3, pad 1, stride 1,
 * deform_groups 1.  Produces the sampled columns [N*H*W][groups][9][C/groups]; the grouped GEMM is
 * tt_conv2d(KH=KW=1, Cin = 9*C, groups).  `offset`: [N][H][W][off_ld], 18 channels (dy, dx per tap).
 */
int tt_dcn_im2col(const float* x, const float* offset, int off_ld, float* col, int N, int H, int W, int C, int groups,
                  tt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (5) LiDAR branch: mmcv hard Voxelization + mmdet3d HardSimpleVFE + spconv rulebooks
 * (lidarnet.py:87-96, cfg thinktwice.py:161-176).
 */
typedef struct {
  int B, P, F;                   /* frames, points per frame, floats per point (5) */
  float lower[3], vsize[3];
  int grid[3];                   /* x, y, z voxel counts (672, 672, 70) */
  int zmax;                      /* voxels with z index >= zmax are dropped (sparse_shape z = 41) */
  int max_points, max_voxels;    /* 10, 160000 (eval) */
  int cap;                       /* capacity of the site arrays */
} tt_voxelize_desc;
size_t tt_voxelize_workspace_bytes(const tt_voxelize_desc* d);
/* feats [cap][F] = mean of the first max_points points (input order) of each voxel; coords [cap][4]
 * (b, z, y, x); *count = number of voxels.  Site order is unspecified (results do not depend on it). */
int tt_voxelize_mean(const tt_voxelize_desc* d, const float* points, float* feats, int* coords, int* count,
                     void* workspace, tt_stream_t stream);

typedef struct {
  int B;
  int in_shape[3], out_shape[3]; /* z, y, x */
  int k[3], s[3], p[3];
  int subm;                      /* 1: output sites = input sites */
  int cap_in, cap_out;
  int table_size;                /* power of two >= 2 * max(cap_in, cap_out) */
} tt_rulebook_desc;
size_t tt_rulebook_workspace_bytes(const tt_rulebook_desc* d);
/* Builds out_coords/out_count (copied from the input for subm) and, whichever is non-NULL:
 *   nbr [cap_out][kvol]                      : input row per (output row, tap) or -1 (output-stationary gather), and/or
 *   pairs_in / pairs_out [kvol][cap_out], pair_count [kvol] : tap-major (input row, output row) lists. */
int tt_sparse_rulebook(const tt_rulebook_desc* d, const int* in_coords, const int* in_count, int* out_coords,
                       int* out_count, int* nbr, int* pairs_in, int* pairs_out, int* pair_count, void* workspace,
                       tt_stream_t stream);
/* Tap-major sparse convolution: out[o] = act(sum_tap w[tap] . in[i] + bias + res[o]) over the rulebook's pairs; one
 * launch walks every (tap, pair tile) and accumulates into the bias-initialised output rows with red.add.f32 (the
 * summation order across taps is not fixed: results repeat to fp32 rounding, not bitwise).
 * impl 0/1: SIMT fp32, w = [kvol][Cin][Cout] fp32 (BatchNorm folded).
 * impl 2/3: tcgen05 TF32 / 3xTF32 gather-GEMM, w = [2][Cout][kvol][Cin] hi / lo planes (as tt_conv2d); needs
 *           Cin, Cout >= 32 and % 4 == 0, in_ld / out_ld % 4 == 0, 16-byte aligned pointers, else TT_ERR_UNSUPPORTED.
 * res: optional [cap_out][res_ld] (SparseBasicBlock identity). */
typedef struct {
  int Cin, Cout, kvol;
  int in_ld, out_ld, res_ld;
  int cap_out, pair_cap;         /* pair lists are [kvol][pair_cap] */
  int act;
  int impl;
} tt_sparse_conv_desc;
int tt_sparse_conv(const tt_sparse_conv_desc* d, const float* feats_in, const float* w, const float* bias,
                   const float* res, const int* pairs_in, const int* pairs_out, const int* pair_count,
                   const int* out_count, float* feats_out, tt_stream_t stream);
/* tt_sparse_conv on scaled-split fp16 operands (see 2b): feats_in_split [2][cap_in][in_ld] halves; writes the fp32 rows
 * and, when out_split != NULL, their split planes [2][cap_out][out_ld].  w_split = [2][Cout][kvol][Cin8]. */
int tt_sparse_conv_f16s(const tt_sparse_conv_desc* d, const void* feats_in_split, long long in_plane, const void* w_split,
                        const float* bias, const float* res, const int* pairs_in, const int* pairs_out, const int* pair_count,
                        const int* out_count, float* feats_out, void* out_split, long long out_plane, tt_stream_t stream);
/* Output-stationary sparse convolution on scaled-split fp16 operands: ONE launch, no atomics.  A work item is a tile of 128 output
 * rows; K runs over all kvol taps through the neighbour table `nbr` [cap_out][kvol] (input row under each tap, -1 = none: a zero
 * row), so the sum of an output row is formed in TMEM in a fixed order and stored once with bias + residual + activation fused
 * (fp32 rows `io->y` and / or split planes `io->y_split`; residual as fp32 `io->res` or planes `io->res_split`).  Pays for empty
 * (row, tap) slots with zero rows — the better form once a fair share of the taps is populated (dilated / dense-ish clouds);
 * tt_sparse_conv_f16s (tap-major pair lists, red.add) is the form for very sparse neighbourhoods.  io->x_split = input planes
 * [2][cap_in][in_ld]; w_split = [2][Cout][kvol][Cin8]; d->cap_out = row capacity, *out_count = rows that exist. */
int tt_sparse_conv_os_f16s(const tt_sparse_conv_desc* d, const tt_f16s_io* io, const int* nbr, const int* out_count,
                           tt_stream_t stream);
/* dense()[N][C][D][H][W].view(N, C*D, H, W) (lidarnet.py:53-56) written channels-last, channel = c*D + z,
 * with the framework's anti-transpose (framework:246) folded in when asked; `dense` must be zero-filled. */
int tt_sparse_to_bev(const float* feats, const int* coords, const int* count, int cap, int C, int D, int H, int W,
                     int anti_transpose, float* dense, tt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (6) Look module (thinktwice_decoder.py:88-187) and multi-scale deformable attention
 * (mmcv._ext.ms_deform_attn_forward, multi_scale_deformable_attn_function.py:72-78, 468-526).
 */
typedef struct {
  int B, num_cams, num_query;    /* 4 cams, 120 look points = (T waypoints + T static points) x 15 z levels */
  int T;                         /* pred_len (4) */
  float img_w, img_h;
  int levels;
  int lvl_h[4], lvl_w[4];
  int C;                         /* 256 channels per FPN level */
  int q_dim;                     /* 4 + 3 + emb_dim + meas_dim + flat_dim = 519 */
  int emb_dim, meas_dim, flat_dim;
  int max_len_cap;               /* rows per (b, cam) in the rebatched buffers (>= num_query) */
} tt_look_desc;
/* thinktwice_decoder.py:155-160 + :88-114,129-137: build the look points from the current waypoints
 * wp [B][T][2], project them into every camera (lidar2img [B][cams][16], clamp depth to 1e-5, ida
 * [B][cams][16], normalise by the image size), validity mask, ordered compaction.
 * Outputs: ref_cam [B][cams][Q][2], order [B][cams][Q] (compacted query ids), counts [B][cams], *max_len. */
int tt_look_project(const tt_look_desc* d, const float* wp, const float* lidar2img, const float* ida, float* ref_cam,
                    int* order, int* counts, int* max_len, tt_stream_t stream);
/* thinktwice_decoder.py:117-127,140-150,161-171: rows [B*cams][cap][rows_ld] = [query | bilinear samples of the 4
 * FPN levels at the ref point (feature index c*levels + l)], zero rows past counts; ref_rebatch [B*cams][cap][2].
 * query = [softplus ctrl 4 | xyz 3 | temporal / static embedding | measurement feat | flattened BEV feat].
 * mlvl: host array of 4 device pointers, each [B*cams][h][w][C] channels-last. */
int tt_look_rebatch(const tt_look_desc* d, const float* wp, const float* ctrl_sp, const float* temporal_emb,
                    const float* static_emb, const float* meas, const float* flat, const float* const* mlvl,
                    const float* ref_cam, const int* order, const int* counts, float* rows, int rows_ld,
                    float* ref_rebatch, tt_stream_t stream);
typedef struct {
  int BN, rows_cap, heads, levels, points, dh; /* 8 heads, 4 levels, 8 points, 32 */
  int lvl_h[4], lvl_w[4], lvl_start[4];
  int num_keys;
  int value_ld, value_coff;      /* floats per key row of `value` (0 = heads*dh) and first channel: several layers' value projections
                                  * computed by ONE convolution share a buffer */
} tt_msda_desc;
/* value [BN][num_keys][value_ld] (channels value_coff .. value_coff + heads*dh); off [BN*cap][heads*levels*points*2]; logits [BN*cap][heads*levels*points];
 * ref [BN*cap][2]; out [BN*cap][heads*dh].  Softmax over levels*points is fused. */
int tt_msda_forward(const tt_msda_desc* d, const float* value, const float* off, const float* logits, const float* ref,
                    const int* max_len, float* out, tt_stream_t stream);
/* msda:338-342: zero the first B rows of every (b, cam), divide by B, sum rows < max_len -> [B][cams*C] */
int tt_look_reduce(const float* rows, int B, int cams, int cap, int C, const int* max_len, float* out,
                   tt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (7) decoder glue
 */
/* GRU input plane (dense_heads/utils.py:93-104): buf[b][p][0:6] = (wp_t, softplus ctrl_t) of step t */
int tt_gru_input(const float* wp, const float* ctrl_sp, int t, int T, float* buf, int ld, int B, int HW,
                 tt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (8) agent-side pre-processing: the CPU work between the CARLA sensors and forward_inference.
 */
typedef struct {
  int n_img;            /* images in `raw` (T * N per tick, or B * T * N) */
  int H, W;             /* raw camera frame (900 x 1600: thinktwice_agent.py:237) */
  int newH, newW;       /* resize_dims of sample_ida_augmentation (transform.py:264-267): int(H * resize), int(W * resize) */
  int outH, outW;       /* final_dim = the crop window's size (448 x 896) */
  int crop_x, crop_y;   /* crop[:2] (transform.py:268-271) */
  int undistort;        /* cfg["undistort"]: sample the raw frame through map_grid first (transform.py:281-282) */
  float div;            /* 255 (transform.py:162) */
  float mean[3], std[3];/* T.Normalize constants (transform.py:144) */
  int pad_H, pad_W, pad_top, pad_left; /* geometry of the stem-plane output (see tt_image_to_split8); ignored without out_split8 */
} tt_preproc_desc;
/* Replaces IDAImageTransform.__call__ + img_transform + ImageTransformMulti(aug=False) for the test-time branch
 * (code/datasets/pipelines/transform.py:275-341, 346-378, 149-163): raw uint8 [n_img][H][W][3] (RGB, as the agent holds it:
 * thinktwice_agent.py:297-300) -> grid_sample(map_grid, bilinear, zeros, align_corners=False) -> T.Resize((newH, newW)) [bilinear,
 * no antialias: torchvision 0.13.1 of docs/INSTALL.md] -> crop -> /div -> (x - mean) / std.
 * map_grid: fp32 [H][W][2] normalised (x, y) exactly as transform.py:237-240 builds it (one map for every image; may be NULL without
 * undistort).  out_nchw: fp32 [n_img][3][outH][outW] (the `img` tensor of the batch dict) and / or out_split8: the row-packed stem's
 * operand planes [2][n_img][pad_H][pad_W][8 halves] (plane stride split_plane halves; border never written).  Either may be NULL. */
int tt_preprocess_u8(const tt_preproc_desc* d, const uint8_t* raw, const float* map_grid, float* out_nchw, void* out_split8,
                     long long split_plane, tt_stream_t stream);
/* thinktwice_agent.py:340-352: out [n_prev + n_now][4] = [rel_mat (4x4 fp64, row-major, DEVICE) applied to prev's xyz | now], z += z_add (2.5);
 * float64 arithmetic like the numpy original, intensity untouched. */
int tt_lidar_stitch(const float* prev, int n_prev, const float* now, int n_now, const double* rel_mat, double z_add, float* out,
                    tt_stream_t stream);
/* code/datasets/carla_dataset.py:314-328 (union2one): dst [n][5] = [curr2key (4x4 fp32, row-major, DEVICE; NULL = key frame: copy) applied to the
 * (x, y, z, intensity) 4-vector of src [n][4] | timestamp]; the caller places the sweeps back to back (key frame first). */
int tt_points_union(const float* src, int n, const float* curr2key, float timestamp, float* dst, tt_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (9) training-side ops: what open_loop_training/train.py needs from the op library besides the forward.
 */
/* Replaces the indexing backward of ops/voxel_pooling/voxel_pooling.py:57-69: grad_input [B][P][C] (fully written) =
 * grad_output[pos_memo b][:][pos_memo y][pos_memo x] where pos_memo [B][P][3] != -1, else 0.  grad_output is the (B, C, Y, X)
 * tensor autograd hands over, addressed by ELEMENT strides (either memory order, no copy). */
int tt_voxel_pooling_backward(int batch_size, int num_points, int num_channels, const float* grad_output, long long stride_b,
                              long long stride_c, long long stride_y, long long stride_x, const int* pos_memo, float* grad_input,
                              tt_stream_t stream);
/* Replace mmcv's ext_module.ms_deform_attn_forward / ms_deform_attn_backward as the reference calls them
 * (code/model_code/dense_heads/multi_scale_deformable_attn_function.py:141-147, 172-183; im2col_step has no meaning here).
 * d: BN = bs, rows_cap = num_queries, heads, levels (<= 4), points, dh, lvl_h / lvl_w / lvl_start = value_spatial_shapes /
 * value_level_start_index, num_keys; value_ld / value_coff as in tt_msda_forward (0 = dense).
 * value [bs][num_keys][heads][dh]; sampling_loc [bs][q][heads][levels][points][2] (x, y) in [0, 1]; attn_weight [bs][q][heads][levels][points];
 * output / grad_output [bs][q][heads*dh].  Backward: grad_value [bs][num_keys][heads][dh] must be ZERO-filled by the caller (as the
 * reference does: msda:168) and is accumulated with fp32 atomics; grad_sampling_loc / grad_attn_weight are fully written. */
int tt_ms_deform_attn_forward(const tt_msda_desc* d, const float* value, const float* sampling_loc, const float* attn_weight, float* output,
                              tt_stream_t stream);
int tt_ms_deform_attn_backward(const tt_msda_desc* d, const float* value, const float* sampling_loc, const float* attn_weight,
                               const float* grad_output, float* grad_value, float* grad_sampling_loc, float* grad_attn_weight,
                               tt_stream_t stream);

#ifdef __cplusplus
}
/* C++ linkage, mangled exactly like the reference's own definition (ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu:38,
 * declared at src/voxel_pooling_forward.cpp:21-22): the reference's pybind wrapper links against libtt_b200.so unchanged.
 * `cudaStream_t` is `struct CUstream_st*`. */
struct CUstream_st;
void voxel_pooling_forward_kernel_launcher(int batch_size, int num_points, int num_channels, int num_voxel_x, int num_voxel_y,
                                           int num_voxel_z, const int* geom_xyz, const float* input_features,
                                           float* output_features, int* pos_memo, struct CUstream_st* stream);
#endif
#endif /* TT_B200_H_ */
