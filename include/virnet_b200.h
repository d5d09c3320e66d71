/* virnet_b200.h — C ABI of libvirnet_sm100.so (B200 / sm_100a only).
 *
 * The reference (zsyOAOA/VIRNet) is pure Python/PyTorch and has no FFI of its
 * own: the operator boundary of its hot path is the set of torch calls made
 * by networks/VIRNet.py, networks/AttResUNet.py, networks/DnCNN.py,
 * networks/KNet.py and loss/ELBO_simple.py.  Each entry point below names the
 * reference call site(s) it replaces.  The Python host side
 * (virnet_b200/networks/) binds these with ctypes; INTEGRATION.md shows
 * the stub.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *    the name ends in _host;
 *  - `stream` is a cudaStream_t passed as void*;
 *  - nothing is allocated and no ownership is taken; the library keeps one
 *    piece of hidden state, a cache of TMA descriptors keyed by
 *    (pointer, shape, box);
 *  - return value: 0 ok, <0 invalid argument (VK_E_*), >0 a cudaError_t.
 *  - activations are NHWC with a channel pitch `ld*` (elements) that is a
 *    multiple of 16 (bf16) / 8 (fp32); padding channels must hold zeros.
 */
#ifndef VIRNET_B200_H_
#define VIRNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VK_E_BADARG (-1)
#define VK_E_UNSUPPORTED (-2)
#define VK_E_NODRIVER (-3)

/* storage / MMA type */
#define VK_BF16 0 /* bf16 storage, tcgen05 kind::f16, fp32 accumulate */
#define VK_TF32 1 /* fp32 storage, tcgen05 kind::tf32, fp32 accumulate */

/* geometry of the implicit GEMM */
#define VK_CONV3X3_S1 0 /* nn.Conv2d(k=3,s=1,p=1)   AttResUNet.py:43,46,117-119,139 DnCNN.py:22-29 KNet.py:32-34,49 */
#define VK_CONV3X3_S2 1 /* nn.Conv2d(k=3,s=2,p=1)   AttResUNet.py:67 (DownBlock.downsampler) */
#define VK_CONVT2X2_S2 2 /* nn.ConvTranspose2d(k=2,s=2) AttResUNet.py:80 (UpBlock.upsampler): 1x1 GEMM + depth-to-space */
#define VK_CONV1X1 3    /* nn.Conv2d(k=1)          AttResUNet.py:18-25 (AttLayer), KNet.py:17-19 (CALayer) */
#define VK_CONV2X2_S2 4 /* input gradient of VK_CONVT2X2_S2: a 2x2 stride-2 conv over the upsampled grid */
#define VK_CONV3X3_S2_DGRAD 5 /* input gradient of VK_CONV3X3_S2: x = coarse-grid dY, output on the fine
                                 grid (out_h x out_w), computed as 4 output-parity phases (4 launches) */

/* epilogues */
#define VK_EPI_STD 0      /* NHWC: v=acc+bias; v*=lrelu'(mask); v+=resid; out1=v; out2=lrelu(v) */
#define VK_EPI_NCHW_F32 1 /* NCHW fp32: v=acc+bias; [exp(clamp(v))]; [v+=resid]; cropped store */

typedef struct vk_conv_args {
  int32_t dtype; /* VK_BF16 | VK_TF32 */
  int32_t kind;  /* VK_CONV3X3_S1 ... */
  /* input activation, NHWC [n][ih][iw][ldx] */
  const void* x;
  int32_t n, ih, iw, ldx;
  /* packed weights [taps][wrows][ldx] (K-major), wrows multiple of 16 */
  const void* w;
  int32_t wrows;
  const float* bias; /* [cout] fp32 (for CONVT: shared by the 4 quadrants) or NULL */
  /* output: EPI_STD NHWC [n][oh][ow][ldo] (oh,ow follow from kind);
   *         EPI_NCHW_F32 [n][cout][crop_h][crop_w] fp32 */
  int32_t cout; /* valid output channels (per quadrant for CONVT) */
  int32_t ldo;
  int32_t epi;
  const void* resid;
  const void* mask;
  void* out1;
  void* out2;
  float alpha;
  int32_t round_out2;
  int32_t act_expclamp;
  float clamp_lo, clamp_hi;
  int32_t crop_h, crop_w;
  int32_t out_h, out_w; /* VK_CONV3X3_S2_DGRAD only: fine-grid size (the forward conv's input size) */
  /* SFT modulation of out2 only (AttLayer, networks/AttResUNet.py:27-32,54-58), or NULL:
   * out2 = lrelu(v * sft_mul[img][c] + sft_add[img][c]); fp32 [n][sft_ld] per-sample scalars
   * (spatially constant conditioning maps).  Served by the persistent kernel only. */
  const float* sft_mul;
  const float* sft_add;
  int32_t sft_ld;
  int32_t pad_;
  /* tuning overrides, 0 = automatic */
  int32_t force_tiles_per_cta;
  int32_t force_chunk_bytes;
  int32_t force_stages;
  int32_t force_tw;
  int32_t force_impl; /* 0 auto; 1 v1 kernel; 2 persistent v2 kernel; 3 v2 single CTAs; 4 v2 CTA pairs (cta_group::2) */
  int32_t force_nt;   /* v2 3x3 s1 only: filter taps per weight stage (1, 3 or 9) */
  /* optional: device buffer of int64[gridsize][8] receiving per-CTA cycle counters
   * {start, mainloop end, epilogue end, producer wait, mma wait, 0, 0, 0}; NULL in production */
  long long* cta_timing;
} vk_conv_args;

/* Forward / data-gradient convolution as an implicit GEMM on tcgen05.
 * Replaces F.conv2d / F.conv_transpose2d and their input-gradient (dgrad is
 * the same kernel over rotated, transposed packed weights). */
int vk_conv_igemm(const vk_conv_args* args, void* stream);

typedef struct vk_wgrad_args {
  int32_t dtype; /* VK_BF16 | VK_TF32 */
  int32_t kind;  /* geometry of the FORWARD op whose weight gradient is taken */
  /* M operand: NHWC [n][gh][gw][lda], un-shifted; its pixel grid is the GEMM K dimension.
   *   conv  : dY (gradient w.r.t. the conv output), m_valid = Cout
   *   convT : X  (forward input),                  m_valid = Cin           */
  const void* a;
  int32_t n, gh, gw, lda, m_valid;
  /* N operand: NHWC [n][bh][bw][ldb], loaded shifted / strided per tap.
   *   conv  : X (forward input),  n_valid = Cin
   *   convT : dY (upsampled grid), n_valid = Cout                          */
  const void* b;
  int32_t bh, bw, ldb, n_valid;
  float* dw;    /* fp32 [taps][m_valid][n_valid], ACCUMULATED into (zero it first) */
  float* dbias; /* fp32 [m_valid] accumulated column sums of the M operand, or NULL */
  int32_t force_ksplit;
  int32_t force_k_rows;
  int32_t force_stages;
  /* Deterministic split-K (run-to-run bit-identical gradients).  partials != NULL: K slice s stores its partial sums
   * with plain stores into partials[s][taps][m_valid][n_valid] (and dbias_partials[slot][m_valid], slot <
   * bias_slots of vk_conv_wgrad_plan) instead of adding atomically into dw / dbias, which are then not touched;
   * vk_wgrad_unpack_batched sums the slices in order.  max_slices = slabs the buffer holds (the split is clamped to
   * it; a whole SM wave needs at most one slab per SM). */
  int32_t max_slices;
  float* partials;
  float* dbias_partials;
  /* swapped != 0 (3x3 stride-1 conv whose OUTPUT has few channels: tail, last SNet layer): the caller passes the
   * operands the other way round — a = X (m_valid = Cin), b = dY (n_valid = Cout <= 16) — so that the narrow tensor is
   * the N operand (one 16-column MMA atom instead of a 128-row tile that is 97 % padding).  The kernel then labels the
   * taps mirrored (shifting dY by +t equals shifting X by -t) and writes dw transposed, i.e. still as
   * [taps][Cout][Cin]; dbias must be NULL (the column sums of `a` are not the bias gradient here). */
  int32_t swapped;
  int32_t pad_;
} vk_wgrad_args;

/* Weight (and bias) gradient: autograd of F.conv2d / F.conv_transpose2d w.r.t.
 * weight and bias (train_denoising_syn.py:179 loss.backward()). */
int vk_conv_wgrad(const vk_wgrad_args* args, void* stream);

/* The split vk_conv_wgrad will use for these arguments (pointers may be NULL; max_slices > 0 selects the
 * deterministic layout): *slices = K slices = partial slabs written, *bias_slots = rows of dbias_partials written. */
int vk_conv_wgrad_plan(const vk_wgrad_args* args, int32_t* slices, int32_t* bias_slots);

/* dW workspace [taps][M][N] -> parameter layout [M][N][taps] (OIHW for Conv2d,
 * [Cin][Cout][kh][kw] for ConvTranspose2d); accumulate != 0 adds into `out`. */
int vk_wgrad_unpack(const float* ws, float* out, int32_t taps, int32_t m, int32_t n, int32_t accumulate,
                    void* stream);

/* The same for every layer of a network in ONE launch.  descs_dev: device array of vk_unpack_desc. */
typedef struct vk_unpack_desc {
  const float* ws; /* [taps][mn]; nslices > 1: [nslices] such slabs, slice_stride elements apart */
  float* out;      /* [mn][taps] */
  int32_t taps, mn;
  int32_t nslices; /* 0 / 1: one slab; > 1: the slabs are summed in slice order (deterministic split-K reduction) */
  int32_t pad_;
  int64_t slice_stride;
} vk_unpack_desc;
int vk_wgrad_unpack_batched(const void* descs_dev, int32_t ndesc, int64_t max_mn, int32_t accumulate, void* stream);

uint32_t vk_sizeof_wgrad_args(void);

/* ---- super-resolution forward path: small per-sample kernels ------------- */

/* KernelNet head, nn.Conv2d(c, cout, 9, stride 4, padding 4, bias=False) (networks/KNet.py:45):
 * x NCHW fp32 [n][c][h][w], w OIHW fp32 [cout][c][9][9] -> out NHWC `dtype` [n][oh][ow][ld]. */
int vk_knet_head(int32_t dtype, const float* x, const float* w, void* out, int32_t n, int32_t c, int32_t h, int32_t wd,
                 int32_t cout, int32_t ld, void* stream);

/* CALayer + skip of RB_Layer (networks/KNet.py:23-26, 37-39): out = f * sigmoid(W2 lrelu(W1 mean(f) + b1) + b2) + skip.
 * f, skip, out: NHWC `dtype` [n][npix][ld]; w1 [r][c], w2 [c][r] are the 1x1 conv parameters. */
int vk_ca_layer(int32_t dtype, const void* f, const void* skip, const float* w1, const float* b1, const float* w2,
                const float* b2, void* out, int32_t n, int32_t npix, int32_t c, int32_t r, int32_t ld, float alpha,
                void* stream);

/* Global average over NCHW fp32 planes [n][c][hw] + head: exp(clamp(mean, lo, hi)) for channels in exp_mask,
 * tanh(mean) for channels in tanh_mask -> out [n][c].  nn.AdaptiveAvgPool2d of SNet (networks/DnCNN.py:30-33,
 * VIRNet.py:81) and of KNet's tail (networks/KNet.py:49,55-58). */
int vk_gap_head(const float* x, int32_t n, int32_t c, int32_t hw, uint32_t exp_mask, uint32_t tanh_mask, float lo,
                float hi, float* out, void* stream);

/* AttLayer MLP (networks/AttResUNet.py:27-32) on per-sample constant conditioning values extra[n][e]
 * (sqrt applied to the entries in sqrt_mask): mul = sigmoid(Wm f2 + bm), add = Wa f2 + ba, fp32 [n][c]. */
int vk_sft_mlp(const float* extra, int32_t n, int32_t e, uint32_t sqrt_mask, const float* w1, const float* b1,
               int32_t c1, const float* w2, const float* b2, int32_t c2, const float* wm, const float* bm,
               const float* wa, const float* ba, int32_t c, float alpha, float* mul, float* add, void* stream);

/* ---- backward of the same layers (autograd of the reference modules) ---- */

/* SFT backward after the producing dgrad applied lrelu': gx = g * mul (+ resid); dmul/dadd [n][c] += sums over pixels
 * of g * x and g (accumulated: zero them first).  g, x, resid, gx: NHWC `dtype` [n][npix][ld]. */
int vk_sft_bwd(int32_t dtype, const void* g, const void* x, const float* mul, const void* resid, void* gx, float* dmul,
               float* dadd, int32_t n, int32_t npix, int32_t c, int32_t ld, void* stream);
/* Deterministic form (bit-reproducible training, no atomics): every block stores its partial sums into `ws`
 * (at least vk_sft_bwd_det_ws_floats(n, npix, c) floats), a second launch adds them in block order. */
int64_t vk_sft_bwd_det_ws_floats(int32_t n, int32_t npix, int32_t c);
int vk_sft_bwd_det(int32_t dtype, const void* g, const void* x, const float* mul, const void* resid, void* gx,
                   float* dmul, float* dadd, int32_t n, int32_t npix, int32_t c, int32_t ld, float* ws,
                   int64_t ws_floats, void* stream);

/* AttLayer MLP backward: parameter gradients are ACCUMULATED into gw1..gba (same shapes as the parameters),
 * d_extra [n][e] += dL/d(raw conditioning values). */
int vk_sft_mlp_bwd(const float* extra, int32_t n, int32_t e, uint32_t sqrt_mask, const float* w1, const float* b1,
                   int32_t c1, const float* w2, const float* b2, int32_t c2, const float* wm, const float* bm,
                   const float* wa, const float* ba, int32_t c, float alpha, const float* dmul, const float* dadd,
                   float* gw1, float* gb1, float* gw2, float* gb2, float* gwm, float* gbm, float* gwa, float* gba,
                   float* d_extra, void* stream);

/* Every AttLayer of the network in one launch (forward: fills mul/add; backward: consumes dmul/dadd, accumulates the
 * parameter gradients and d_extra).  descs_dev: device array of vk_sft_desc, one per layer; requires c1 + c2 <= c. */
typedef struct vk_sft_desc {
  const float *w1, *b1, *w2, *b2, *wm, *bm, *wa, *ba;      /* conv1, conv2, mul_conv, add_conv of the AttLayer */
  float *gw1, *gb1, *gw2, *gb2, *gwm, *gbm, *gwa, *gba;    /* their gradients (accumulated) */
  float *mul, *add, *dmul, *dadd;                          /* [n][c] each */
  int32_t c1, c2, c, pad_;
} vk_sft_desc;
int vk_sft_mlp_batched(const void* descs_dev, int32_t n_layers, int32_t max_c, const float* extra, int32_t n, int32_t e,
                       uint32_t sqrt_mask, float alpha, void* stream);
int vk_sft_mlp_bwd_batched(const void* descs_dev, int32_t n_layers, int32_t max_c, const float* extra, int32_t n,
                           int32_t e, uint32_t sqrt_mask, float alpha, float* d_extra, void* stream);
/* Deterministic form: block (sample, layer) stores its gradients into private slots of `ws`, a second launch adds the
 * sample slots in sample order (parameter gradients) and the layer slots in layer order (d_extra).
 * params_per_sample = sum over the layers of c1*e + c1 + c2*c1 + c2 + 2*(c*c2 + c);
 * ws_floats >= n * params_per_sample + n_layers * n * e. */
int vk_sft_mlp_bwd_batched_det(const void* descs_dev, int32_t n_layers, int32_t max_c, const float* extra, int32_t n,
                               int32_t e, uint32_t sqrt_mask, float alpha, float* d_extra, int64_t params_per_sample,
                               float* ws, int64_t ws_floats, void* stream);
uint32_t vk_sizeof_sft_desc(void);

/* ---- SFT modulation with SPATIALLY VARYING conditioning maps (csrc/vk_sft_spatial.cu) ----
 * Source of the conditioning channels at a pixel of any U-Net level: `ec` per-sample constants cst[n][ec] followed by
 * `em` map channels map[n][em][eh][ew] that are nearest-upsampled by esf to the un-padded full resolution hh x ww,
 * reflect-padded (bottom / right) to hp x wp and nearest-resized to the level's grid; sqrt applied to the channels in
 * sqrt_mask.  Replaces kinfo.repeat / F.interpolate(sigma.sqrt()) / cat (networks/VIRNet.py:87-95),
 * util_net.pad_input and F.interpolate(extra_maps, size, 'nearest') (networks/AttResUNet.py:147-168). */
typedef struct vk_extra_src {
  const float* cst; /* [n][ec] fp32 or NULL */
  const float* map; /* [n][em][eh][ew] fp32 or NULL */
  int32_t ec, em, eh, ew, esf;
  uint32_t sqrt_mask;
  int32_t hh, ww; /* un-padded full-resolution size */
  int32_t hp, wp; /* reflect-padded full-resolution size (multiple of every level's grid) */
} vk_extra_src;

/* out = lrelu(x * mul + add) with (mul, add) = AttLayer(conditioning at that pixel): AttLayer.forward
 * (networks/AttResUNet.py:27-32) + the modulation of AttResBlock.forward (:54-58) in one pass.
 * x, out: NHWC `dtype` [n][h][w][ld]; w1 [c1][ec+em], w2 [c2][c1], wm / wa [c][c2] fp32 (the nn.Conv2d 1x1 weights). */
typedef struct vk_sft_apply_args {
  int32_t dtype;
  int32_t n, h, w, c, ld;
  int32_t c1, c2;
  const void* x;
  void* out;
  const float *w1, *b1, *w2, *b2, *wm, *bm, *wa, *ba;
  vk_extra_src extra;
  float alpha;
  int32_t round_tf32; /* fp32 storage: round the output to TF32 (it feeds a tcgen05 kind::tf32 MMA) */
} vk_sft_apply_args;
int vk_sft_apply(const vk_sft_apply_args* args, void* stream);
uint32_t vk_sizeof_sft_apply_args(void);

/* Backward of vk_sft_apply (autograd of networks/AttResUNet.py:27-32,54-58).  g = dL/d(x*mul+add) (the producing
 * dgrad kernel already applied lrelu').  Writes gx = g * mul (+ resid), and the per-pixel operands of the four pixel-K
 * GEMMs that form the AttLayer's parameter gradients with vk_conv_wgrad(VK_CONV1X1): dm [..][ld] (with g: mul_conv /
 * add_conv), f2, dq2 [..][ld2] and f1, dq1 [..][ld1] (conv2), ev [..][16] (conv1 input).  The gradient w.r.t. the
 * conditioning values is accumulated (atomicAdd, zero first) into d_cst [n][ec] and d_map [n][em][eh][ew]. */
typedef struct vk_sft_apply_bwd_args {
  int32_t dtype;
  int32_t n, h, w, c, ld;
  int32_t c1, c2, ld1, ld2;
  const void *g, *x, *resid;
  void *gx, *dm, *f2, *dq2, *f1, *dq1, *ev;
  const float *w1, *b1, *w2, *b2, *wm, *bm, *wa, *ba;
  float *d_cst, *d_map;
  vk_extra_src extra;
  float alpha;
  int32_t pad_;
} vk_sft_apply_bwd_args;
int vk_sft_apply_bwd(const vk_sft_apply_bwd_args* args, void* stream);
uint32_t vk_sizeof_sft_apply_bwd_args(void);

/* Gradient of the head convolution's packed input w.r.t. its conditioning channels [c, c + ec + em) (NHWC `dtype`
 * [n][hp][wp][ld]), folded through reflect padding / nearest up-sampling / sqrt into d_cst and d_map (accumulated). */
int vk_extra_head_grad(int32_t dtype, const void* g, int32_t n, int32_t ld, int32_t c, const vk_extra_src* extra,
                       float* d_cst, float* d_map, void* stream);

/* vk_pack_input with mixed conditioning channels: out NHWC `dtype` [n][hp][wp][ld] =
 * [img (nearest x sf, reflect padded) | constants | maps | 0 ...]  (networks/VIRNet.py:83-96, AttResUNet.py:147-153) */
int vk_pack_input_mixed(int32_t dtype, const float* img, int32_t n, int32_t c, int32_t h, int32_t w, int32_t sf,
                        const vk_extra_src* extra, void* out, int32_t ld, void* stream);

/* CALayer + skip backward: df = g * s + dy / npix (the skip gradient is g); parameter gradients accumulated. */
int vk_ca_layer_bwd(int32_t dtype, const void* g, const void* f, const float* w1, const float* b1, const float* w2,
                    const float* b2, void* df, float* gw1, float* gb1, float* gw2, float* gb2, int32_t n, int32_t npix,
                    int32_t c, int32_t r, int32_t ld, float alpha, void* stream);
/* Deterministic form: per-sample parameter-gradient slots in `ws` (at least n * (2 r c + r + c) floats), added in
 * sample order by a second launch. */
int vk_ca_layer_bwd_det(int32_t dtype, const void* g, const void* f, const float* w1, const float* b1, const float* w2,
                        const float* b2, void* df, float* gw1, float* gb1, float* gw2, float* gb2, int32_t n,
                        int32_t npix, int32_t c, int32_t r, int32_t ld, float alpha, float* ws, int64_t ws_floats,
                        void* stream);

/* Backward of vk_gap_head: gx NHWC `dtype` [n][hw][ld] = gout[n][c] * head'(outv[n][c]) / hw (0 for c >= C). */
int vk_gap_head_bwd(int32_t dtype, const float* gout, const float* outv, int32_t n, int32_t c, int32_t hw,
                    uint32_t exp_mask, uint32_t tanh_mask, float lo, float hi, void* gx, int32_t ld, void* stream);

/* Weight gradient of vk_knet_head, ACCUMULATED into gw [cout][c][9][9]; g NHWC `dtype` [n][oh][ow][ld]. */
int vk_knet_head_wgrad(int32_t dtype, const float* x, const void* g, float* gw, int32_t n, int32_t c, int32_t h,
                       int32_t wd, int32_t cout, int32_t ld, void* stream);
/* Deterministic form: per-sample slots in `ws` (at least n * cout * c * 81 floats), added in sample order by a
 * second launch (no atomics). */
int vk_knet_head_wgrad_det(int32_t dtype, const float* x, const void* g, float* gw, int32_t n, int32_t c, int32_t h,
                           int32_t wd, int32_t cout, int32_t ld, float* ws, int64_t ws_floats, void* stream);

/* F.interpolate(x, scale_factor=sf, mode="nearest") on NCHW fp32 (networks/VIRNet.py:83). */
int vk_upsample_nearest(const float* x, float* out, int32_t n, int32_t c, int32_t h, int32_t w, int32_t sf,
                        void* stream);

/* ---- real-noise denoising trainer: data-side operators ------------------- */

/* utils/util_denoising.py:54-63 (noise_estimate_fun): out = clamp_min(window (*) (noisy - gt)^2, floor) per plane with
 * reflect padding; `window` is the k_size x k_size normalised Gaussian of inverse_gamma_kernel (:24-35), fp32 on the
 * device; tensors are [planes][h][w] fp32 (planes = N * C). */
int vk_noise_estimate(const float* noisy, const float* gt, const float* window, int32_t k_size, float* out,
                      int32_t planes, int32_t h, int32_t w, float floor_, void* stream);

/* ---- evaluation-side kernels (csrc/vk_eval.cu) ----
 * 8-fold flip / rotate self-ensemble (scripts/denoising_virnet_real_sidd.py:120-136,
 * dnd_submission_py/pytorch_wrapper.py:17-32; utils/util_image.py:391-466 data_aug_np / inverse_data_aug_np):
 * vk_aug8 gathers the 8 augmentations of `planes` = N*C fp32 planes [h][w] into out_a [4][planes][h][w] (modes 0, 1,
 * 4, 5) and out_b [4][planes][w][h] (modes 2, 3, 6, 7) for ONE batched forward (two when h != w); vk_aug8_merge
 * averages the inversely augmented network outputs (accumulation order 0..7, x 1/8; clip01: DND's final clip). */
int vk_aug8(const float* in, float* out_a, float* out_b, int64_t planes, int32_t h, int32_t w, void* stream);
int vk_aug8_merge(const float* in_a, const float* in_b, float* out, int64_t planes, int32_t h, int32_t w,
                  int32_t clip01, void* stream);

/* skimage.img_as_ubyte(clamp(x, 0, 1)) of the eval scripts: NCHW fp32 -> [n][h][w][c] uint8, rint(x * 255) in fp32. */
int vk_to_u8(const float* in, uint8_t* out, int32_t n, int32_t c, int32_t h, int32_t w, void* stream);

/* utils/util_image.py:68-89 calculate_psnr on uint8 HWC images (device pointers): *ssd_out = exact sum of squared
 * differences over the border-cropped region, on RGB (all c channels) or on the rounded Y channel of MATLAB's
 * rgb2ycbcr (:129-153) when ycbcr != 0.  PSNR = 20 log10(255 / sqrt(ssd / count)) is formed by the caller. */
int vk_psnr_u8(const uint8_t* a, const uint8_t* b, int32_t h, int32_t w, int32_t c, int32_t border, int32_t ycbcr,
               uint64_t* ssd_out, void* stream);

/* utils/util_image.py:16-66 calculate_ssim: sums_out[ch] = sum of the SSIM map over the valid (h-2b-10) x (w-2b-10)
 * region per channel (one entry when ycbcr), fp64; window121_host = the 11x11 Gaussian window (HOST pointer). */
int vk_ssim_u8(const uint8_t* a, const uint8_t* b, int32_t h, int32_t w, int32_t c, int32_t border, int32_t ycbcr,
               const double* window121_host, double* sums_out, void* stream);

/* datasets/data_tools.py:21-30 (MixUp_AUG.aug): out_x[i] = lam[i] * x[i] + (1 - lam[i]) * x[perm[i]] for x in {a, b};
 * [n][per_sample] fp32, per_sample a multiple of 4; perm int64 [n], lam fp32 [n] on the device; out must not alias in. */
int vk_mixup(const float* a, const float* b, const int64_t* perm, const float* lam, float* out_a, float* out_b, int32_t n,
             int64_t per_sample, void* stream);

/* ---- device-side synthesis of denoising training batches ----------------- */

/* datasets/DenoisingDatasets.py:180-253 (SimulateTrain.__getitem__) after cropping: patches uint8 [n][p][p][c] (RGB, HWC);
 * params fp64 [n][6] = {center_h, center_w, scale, up, down, iid_level} (scale <= 0: constant 'iid' map = iid_level);
 * aug int32 [n] in 0..7 (utils/util_image.py:391-436); noise fp32 [n][p][p][c] ~ N(0,1) (the reference's
 * torch.randn(im_gt.shape)); outputs NCHW fp32: im_noisy, im_gt [n][c][p][p], sigma_gt [n][1][p][p] = max(sigma^2, 1e-10). */
int vk_synth_denoise(const uint8_t* patches, const double* params, const int32_t* aug, const float* noise, int32_t n,
                     int32_t p, int32_t c, int32_t clip, float* im_noisy, float* im_gt, float* sigma_gt, void* stream);

/* ---- super-resolution negative ELBO ------------------------------------- */

/* loss/ELBO_simple.py:82-138 (elbo_sisr) with its helpers (:55-80), utils/util_sisr.py:26-58 (sigma2kernel) and
 * :127-144 (conv_multi_kernel_tensor), and ResizeRight/resize_right.py:29-76 as the dense 1-D operators rh / rw.
 * One call computes the 8 scalar terms AND the gradients of `loss` w.r.t. mu, sigma_est and kinfo_est.
 * All tensors are fp32, images NCHW.  The three random draws the reference makes inside the loss are inputs:
 * gamma_draw [n][2] ~ Gamma(kappa0-1, 1), rho_draw [n] ~ N(0,1), z_draw like mu ~ N(0,1).
 * prior_mean / prior_logmean [n]: per-sample mean of sigma_prior and of log(sigma_prior) (equal to
 * sigma_prior and its log for the N x 1 x 1 x 1 Gaussian-noise prior of train_SISR.py:202).
 * Every reduction goes through per-block slots of the workspace that are added in a fixed order (no atomics):
 * the terms and all gradients are bit-identical from run to run.
 * terms[8] = loss, lh, kl_rnet, kl_snet, kl_knet, kl_knet0, kl_knet1, kl_knet2 (the reference's return order);
 * kernel [n][k_size*k_size] is the re-sampled blur kernel (last entry of the reference's detail list). */
typedef struct vk_elbo_sisr_args {
  const float* mu;            /* [n][c][H][W] */
  const float* im_hr;         /* [n][c][H][W] */
  const float* im_lr;         /* [n][c][h][w] */
  const float* sigma_est;     /* [n] */
  const float* kinfo_est;     /* [n][3] */
  const float* kinfo_gt;      /* [n][3] */
  const float* prior_mean;    /* [n] */
  const float* prior_logmean; /* [n] */
  const float* gamma_draw;    /* [n][2] */
  const float* rho_draw;      /* [n] */
  const float* z_draw;        /* [n][c][H][W] */
  const float* rh;            /* [h][H] down-sampling operator along rows */
  const float* rw;            /* [w][W] down-sampling operator along columns */
  float* d_mu;                /* [n][c][H][W] */
  float* d_sigma;             /* [n] */
  float* d_kinfo;             /* [n][3] */
  float* kernel;              /* [n][k_size*k_size] */
  float* terms;               /* [8] */
  void* ws;                   /* workspace of at least vk_elbo_sisr_ws_bytes() bytes */
  int64_t ws_bytes;
  int32_t n, c, H, W, h, w, k_size;
  float center;               /* kernel centre: k_size/2, + 0.5*(sf - k_size%2) when shifted (util_sisr.py:43-46) */
  float alpha0, digamma_am1;  /* digamma(alpha0 - 1) */
  float kappa0, r2, eps2, pk0, pk1;
} vk_elbo_sisr_args;

int64_t vk_elbo_sisr_ws_bytes(int32_t n, int32_t c, int32_t H, int32_t W, int32_t h, int32_t w, int32_t k_size);
uint32_t vk_sizeof_elbo_sisr_args(void);
int vk_elbo_sisr(const vk_elbo_sisr_args* args, void* stream);

/* Device-side synthesis of SISR training pairs: datasets/SISRDatasets.py:86-104 (GeneralTrainFloder.__getitem__ after
 * the crop / augmentation, Gaussian-noise branch): im_blur = down(clip(conv(im_hr, kernel[n]), 0, 1)) with
 * scipy.ndimage.convolve(mode='reflect') semantics and the rh / rw down-sampling operators (Direct or ResizeRight
 * bicubic), im_lr = clip(im_blur + noise * std[n], 0, 1).  NCHW fp32; kernels [n][k*k]; noise [n][c][h][w]. */
int64_t vk_sisr_degrade_ws_bytes(int32_t n, int32_t c, int32_t H, int32_t W, int32_t w);
int vk_sisr_degrade(const float* im_hr, const float* kernels, int32_t k_size, const float* rh, const float* rw,
                    const float* noise, const float* std, float* im_blur, float* im_lr, void* ws, int64_t ws_bytes,
                    int32_t n, int32_t c, int32_t H, int32_t W, int32_t h, int32_t w, void* stream);

/* ---- HBM-bound kernels ------------------------------------------------- */

/* NCHW fp32 image (+ conditioning channels) -> NHWC `dtype`, reflect-padded bottom/right
 * to (hp, wp), nearest-upsampled by sf, channel-padded with zeros to ld.
 * Replaces util_net.pad_input (utils/util_net.py:20-25), torch.cat([x, extra]) (AttResUNet.py:152-153),
 * sigma.sqrt() (VIRNet.py:44,92-94), F.interpolate(nearest) and .repeat (VIRNet.py:83-95).
 * extra: NCHW fp32 map [n][e][eh][ew] (extra_is_map=1, sampled at (y/esf, x/esf)) or per-sample
 * constants [n][e]; bit i of extra_sqrt_mask applies sqrt to extra channel i. */
int vk_pack_input(int32_t dtype, const float* img, int32_t n, int32_t c, int32_t h, int32_t w, int32_t sf,
                  const float* extra, int32_t e, int32_t extra_is_map, int32_t extra_sqrt_mask, int32_t eh,
                  int32_t ew, int32_t esf, void* out, int32_t hp, int32_t wp, int32_t ld, void* stream);

/* NCHW fp32 gradient [n][c][h][w] -> NHWC `dtype` [n][hp][wp][ld], zero outside the crop
 * (autograd of `tail(x)[..., :h, :w]`, AttResUNet.py:173). */
int vk_pack_grad(int32_t dtype, const float* g, int32_t n, int32_t c, int32_t h, int32_t w, void* out, int32_t hp,
                 int32_t wp, int32_t ld, void* stream);

/* Chain rule through sigma = exp(clamp(l, log_lo, log_hi)), s = sqrt(sigma) and the reflect
 * padding of s (VIRNet.py:43-44, util_net.py:20-25):
 *   g_l = [in range] * sigma * (g_sigma + fold(g_in[..., chan]) / (2 sqrt(sigma)))
 * sigma, g_sigma: NCHW fp32 [n][sc][h][w] (g_sigma may be NULL); g_in: NHWC `dtype`
 * [n][hp][wp][ld_in] or NULL; out: NHWC `dtype` [n][h][w][ld]. */
int vk_sigma_head_bwd(int32_t dtype, const float* sigma, const float* g_sigma, const void* g_in, int32_t ld_in,
                      int32_t chan, int32_t hp, int32_t wp, void* out, int32_t ld, int32_t n, int32_t sc,
                      int32_t h, int32_t w, float log_lo, float log_hi, void* stream);

/* elbo_denoising_simple (loss/ELBO_simple.py:23-53), forward and gradient in one pass.
 * All tensors NCHW fp32; sigma / beta0 have sc (1 or c) channels; the prior parameter used is
 * beta0[i] * beta0_scale (pass sigma_gt and alpha0 to fuse train_denoising_syn.py:172).  out4 = {loss, lh, kl_gauss,
 * kl_Igamma}; d_mu / d_sigma (may be NULL) receive grad_scale * dloss/d(.).  acc_ws: scratch of acc_ws_doubles
 * doubles (3 * VK_REDUCE_MAX_BLOCKS always suffices): one slot per thread block, summed in a fixed order (no
 * atomics: the loss value is run-to-run bit-identical).  digamma_alpha0_m1 = digamma(alpha0 - 1), computed by the
 * caller. */
#define VK_REDUCE_MAX_BLOCKS 592 /* largest grid of the two-pass reductions below (4 blocks per SM) */
int vk_elbo_denoise(const float* mu, const float* sigma, const float* noisy, const float* gt, const float* beta0,
                    float beta0_scale, int32_t n, int32_t c, int32_t sc, int32_t h, int32_t w, float eps2, float alpha0,
                    float digamma_alpha0_m1, float grad_scale, float* d_mu, float* d_sigma, double* acc_ws,
                    int32_t acc_ws_doubles, float* out4, void* stream);

/* One launch packs every layer's fp32 parameters into the K-major GEMM operands.
 * descs_dev: device array of vk_pack_desc; max_elems = largest dst element count. */
typedef struct vk_pack_desc {
  const float* src; /* parameter [dim0][dim1][taps] fp32 */
  void* dst;        /* [dst_taps][rows][ld] `dtype` */
  int32_t dim0, dim1, taps;
  int32_t rows, ld;
  int32_t dst_taps;
  int32_t mode; /* 0 conv fprop, 1 conv dgrad (rotated, transposed), 2 convT fprop, 3 convT dgrad,
                   4 stride-2 conv dgrad (transposed, not rotated) */
  int32_t pad_;
} vk_pack_desc;
int vk_pack_weights(int32_t dtype, const void* descs_dev, int32_t ndesc, int64_t max_elems, int32_t round_tf32,
                    void* stream);

/* out[c] += sum over pixels of x[p][c] (NHWC `dtype`, pitch ld): ConvTranspose2d bias gradient.
 * ws != NULL (ws_floats >= VK_REDUCE_MAX_BLOCKS * c): per-block partial sums go to ws and a second launch adds them
 * in block order (bit-reproducible); ws == NULL: one atomicAdd per (block, channel). */
int vk_channel_sum(int32_t dtype, const void* x, int64_t npix, int32_t ld, int32_t c, float* out, float* ws,
                   int64_t ws_floats, void* stream);

/* Per-sample form: out[s][c] += sum over the npix pixels of sample s (x is [n][npix][ld]). */
int vk_channel_sum_batched(int32_t dtype, const void* x, int32_t n, int64_t npix, int32_t ld, int32_t c, float* out,
                           void* stream);

/* Per-sub-network gradient L2 norm -> clip coefficient -> Adam update over flat fp32 buffers, no
 * host synchronisation (train_denoising_syn.py:182-184).  groups_dev: device array of
 * vk_adam_group; sq_ws: scratch of sq_ws_doubles doubles (ngroups * VK_REDUCE_MAX_BLOCKS always suffices: one
 * slot per (group, thread block), summed in a fixed order by the update kernel — no atomics, the step is
 * bit-reproducible); grads are first multiplied by grad_scale (1/world_size after a sum all-reduce); norms_out
 * (may be NULL) receives the pre-clip norms. */
typedef struct vk_adam_group {
  int64_t begin, end; /* element range in the flat buffers */
  float max_norm;
  int32_t pad_;
} vk_adam_group;
int vk_adam_clip_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const void* groups_dev,
                      int32_t ngroups, int64_t max_group_elems, double* sq_ws, int32_t sq_ws_doubles, float grad_scale,
                      float lr, float beta1, float beta2, float eps, int32_t step, float* norms_out, void* stream);

/* Same, with the per-step scalars read from device memory: hyper_dev = {lr, 1 - beta1^step, sqrt(1 - beta2^step)}.
 * A training step captured in a CUDA graph replays with fresh values written by the host before each launch. */
int vk_adam_clip_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const void* groups_dev,
                          int32_t ngroups, int64_t max_group_elems, double* sq_ws, int32_t sq_ws_doubles,
                          float grad_scale, float beta1, float beta2, float eps, const float* hyper_dev,
                          float* norms_out, void* stream);

/* sizeof(vk_conv_args) as compiled into the library (binding self-check). */
uint32_t vk_sizeof_conv_args(void);

/* Library / build information: returns a static NUL-terminated string. */
const char* vk_version(void);

/* Number of kernel launches issued by this library since load (all entry
 * points); used by bench.py for the "gpu_launches" figure. */
uint64_t vk_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VIRNET_B200_H_ */
