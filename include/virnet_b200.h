/* virnet_b200.h — C ABI of libvirnet_sm100.so (B200 / sm_100a only).
 *
 * The reference (zsyOAOA/VIRNet) is pure Python/PyTorch and has no FFI of its
 * own: the operator boundary of its hot path is the set of torch calls made
 * by networks/VIRNet.py, networks/AttResUNet.py, networks/DnCNN.py,
 * networks/KNet.py and loss/ELBO_simple.py.  Each entry point below names the
 * reference call site(s) it replaces.  The Python host side
 * (virnet_b200/networks/*.py) binds these with ctypes; INTEGRATION.md shows
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
  const float* bias; /* [wrows] or NULL */
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
  /* tuning overrides, 0 = automatic */
  int32_t force_tiles_per_cta;
  int32_t force_chunk_bytes;
  int32_t force_stages;
  int32_t force_tw;
} vk_conv_args;

/* Forward / data-gradient convolution as an implicit GEMM on tcgen05.
 * Replaces F.conv2d / F.conv_transpose2d and their input-gradient (dgrad is
 * the same kernel over rotated, transposed packed weights). */
int vk_conv_igemm(const vk_conv_args* args, void* stream);

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
