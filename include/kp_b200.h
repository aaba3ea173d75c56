/* kp_b200.h — C ABI of the B200-native stage-1 hot path.
 *
 * The reference (YunjiKim/Unsupervised-Keypoint-Learning-for-Guiding-Class-conditional-Video-Prediction)
 * is pure Python on TensorFlow 1.12 and has NO native/FFI boundary of its own; the arithmetic of the
 * stage-1 path lives inside TF ops.  Each entry point below therefore cites the *reference call site*
 * (file:line under /root/reference) whose TF op chain it replaces.  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name says host; the library never allocates,
 *     frees or retains device memory (the caller's allocator owns all buffers, including workspaces);
 *   - tensors are channels-last (NHWC) like the reference's; fp32 at the keypoint boundary;
 *   - all work is enqueued asynchronously on the caller-supplied cudaStream_t (passed as void*);
 *   - return value: 0 = ok, <0 = error (KP_ERR_*), message via kp_last_error() (thread-local);
 *   - no exceptions cross the ABI.
 */
#ifndef KP_B200_H
#define KP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KP_OK 0
#define KP_ERR_INVALID_ARG (-1)
#define KP_ERR_UNSUPPORTED (-2)
#define KP_ERR_CUDA (-3)
#define KP_ERR_DRIVER (-4)

/* ABI version of this header (bumped on any signature change). */
int kp_abi_version(void);
/* Last error message of the calling thread ("" if none). */
const char* kp_last_error(void);
/* Number of CUDA kernels this library has launched in this process (monotonic; for bench accounting). */
unsigned long long kp_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * K1 — keypoint math (utils/model.py)
 * ------------------------------------------------------------------------------------------- */

/* Fused get_coord x2 + stack + get_gaussian_maps.
 *   replaces: utils/model.py:63-70 (get_coord, twice), models/networks/__init__.py:68-71 (stack (x,y)),
 *             utils/model.py:49-60 (get_gaussian_maps) as called from
 *             models/detector_translator_model.py:166-169.
 *   logits [B,H,W,K] f32 -> mu [B,K,2] f32 (x,y in [-1,1]);
 *   prob_x [B,W,K] / prob_y [B,H,K] (get_coord's second return value; nullable);
 *   maps [B,map_h,map_w,K] f32 (nullable: then only the soft-argmax runs).                        */
int kp_softargmax_render_fwd(const float* logits, int B, int H, int W, int K,
                             float* mu, float* prob_x, float* prob_y,
                             float* maps, int map_h, int map_w, float inv_std, void* stream);

/* Backward of the above.  d_maps [B,map_h,map_w,K] (nullable) and d_mu_extra [B,K,2] (nullable, a
 * gradient arriving directly on mu) -> d_logits [B,H,W,K].  mu/prob_x/prob_y are the forward outputs.
 * d_mu_scratch [B,K,2] is caller-owned scratch, only needed for shapes off the fast path (nullable
 * for K=40, W=128).                                                                               */
int kp_softargmax_render_bwd(const float* d_maps, const float* d_mu_extra,
                             const float* mu, const float* prob_x, const float* prob_y,
                             int B, int H, int W, int K, int map_h, int map_w, float inv_std,
                             float* d_logits, float* d_mu_scratch, void* stream);

/* get_gaussian_maps alone (utils/model.py:49-60; callers models/final_model.py:79-92,102-105,
 * models/detector_translator_model.py:176-177).  mu [B,K,2] -> maps [B,h,w,K].                    */
int kp_render_fwd(const float* mu, int B, int K, int h, int w, float inv_std, float* maps, void* stream);

/* Its backward: d_maps [B,h,w,K] (+ optional d_mu_extra) -> d_mu [B,K,2].                          */
int kp_render_bwd(const float* d_maps, const float* d_mu_extra, const float* mu,
                  int B, int K, int h, int w, float inv_std, float* d_mu, void* stream);

/* get_gaussian_maps + colorize_point_maps fused (utils/model.py:42-46 on :49-60; callers
 * models/final_model.py:102-109, models/detector_translator_model.py:207-208).
 * mu [B,K,2], colors [K,3] -> out [B,h,w,3].                                                       */
int kp_render_colorize_fwd(const float* mu, const float* colors, int B, int K, int h, int w,
                           float inv_std, float* out, void* stream);

/* colorize_point_maps on already-materialised maps (utils/model.py:42-46).
 * maps [n_pixels,K], colors [K,3] -> out [n_pixels,3] (max over k of maps*colour).                 */
int kp_colorize_fwd(const float* maps, const float* colors, long long n_pixels, int K, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * C1/C2 — convolutions as tap-GEMMs on the tcgen05 tensor cores
 *
 * One primitive serves every convolution of the path (layers.conv, models/networks/layers.py:4-10, all
 * call sites in models/networks/__init__.py:10-95,144-150; Vgg19.conv_layer, models/networks/vgg.py:48-55)
 * forward AND data-gradient:
 *
 *     out[n,u,v, 0:Cout] = act( bias + sum_{tap t} sum_{source s} sum_c  A_{map(t)+s}[n, u+dh[t], v+dw[t], c] * Wp[:, k(t,s,c)] )
 *
 * where every A map is a strided 4-D view [N][Hd][Wd][C] of a bf16 NHWC tensor (out-of-range reads are
 * zero = TF zero padding), Wp is a packed bf16 matrix [Cout_pad][Ktot] (K-major) and the output is a
 * strided NHWC view (bf16 or f32).  Stride-2 convolutions use four parity views of the input (one per
 * tap parity); stride-2 data-gradients use four launches with strided output views; channel concats are
 * several sources per tap ("virtual concat", no copy).  The host mirror (kp_b200/conv.py) lowers TF
 * SAME-padding convolutions onto this descriptor.
 * Kernel: 128-pixel x BN-channel tiles, TMA (tile mode, hardware swizzle) into a multi-stage shared
 * memory ring, tcgen05.mma (M=128, N=BN, K=16, bf16 x bf16 -> fp32 in TMEM), fused epilogue.
 * ------------------------------------------------------------------------------------------- */
#define KP_MAX_MAPS 4
#define KP_MAX_TAPS 64

#define KP_ACT_NONE 0
#define KP_ACT_RELU 1
#define KP_ACT_LEAKY 2      /* x >= 0 ? x : alpha*x */
#define KP_ACT_SIGMOID 3
#define KP_ACT_SIGMOID_LAST 4 /* channels [0,Cout-1) linear, channel Cout-1 sigmoid (translator heads) */

typedef struct {
    int src;               /* which of the src[] pointers this view reads */
    int C;                 /* channels (multiple of 8) */
    int Wd, Hd;            /* view extent; reads outside [0,Wd) x [0,Hd) x [0,N) return 0 */
    long long off;         /* element offset of view(0,0,0,0) from the source pointer */
    long long sw, sh, sn;  /* element strides of the view (multiples of 8) */
} kp_tap_view;

typedef struct {
    int N;                                   /* images */
    int n_maps;  kp_tap_view map[KP_MAX_MAPS];
    int n_taps, n_src;                       /* each tap reads n_src consecutive maps starting at map_first[t] */
    signed char dh[KP_MAX_TAPS], dw[KP_MAX_TAPS], map_first[KP_MAX_TAPS];
    int CB;                                  /* channel block: 16, 32 or 64 (= swizzle span 32/64/128 B) */
    int Cout_pad, Ktot;                      /* packed weight matrix [Cout_pad][Ktot] bf16 */
    int Ho, Wo;                              /* output extent in tile space */
    long long out_off, out_sw, out_sh, out_sn; /* output view (elements) */
    int Cout;                                /* channels written per pixel */
    int out_f32;                             /* 0: bf16 output, 1: f32 output */
    int act; float alpha;
    int accumulate;                          /* 1: out += result instead of out = result; 2: f32 output that the CALLER has
                                                zeroed and that takes atomic adds (lets long-K launches split K over CTAs,
                                                also into strided views: stride-2 data gradients of img_discr) */
    int TW, TH, TN, BN;                      /* tile: TW*TH*TN == 128 pixels x BN channels; 0 = let the library choose */
    int stat_groups;                         /* batch-norm statistics per batch segment: images [g*N/G, (g+1)*N/G) accumulate
                                                into stats_sum/stats_sq + g*Cout_pad (G = stat_groups, 0/1 = one segment).  Two
                                                calls of a shared-weight network (pose_encoder on image and future_image,
                                                detector_translator_model.py:166-167) run as ONE launch on the concatenated batch
                                                and still normalise each call with its own statistics.                          */
} kp_tapconv_desc;

/* src[i]: bf16 NHWC sources; wpacked: bf16 [Cout_pad][Ktot]; bias: f32 [Cout_pad] (nullable);
 * out: bf16 or f32 view; stats_sum / stats_sq: f32 [stat_groups][Cout_pad] accumulators (nullable) that receive
 * per-channel sum and sum of squares of the PRE-bias accumulators (for batch-norm statistics,
 * models/networks/layers.py:13-14) via atomic adds.                                                */
int kp_tapconv_bf16(const kp_tapconv_desc* desc, const void* const* src, const void* wpacked, const float* bias,
                    void* out, float* stats_sum, float* stats_sq, void* stream);

/* Weight gradient of the same convolutions (the backward of layers.conv w.r.t. its kernel):
 *
 *     dW[tap t][ci][co] += sum_{n,u,v}  X_{map(t)}[n, u+dh[t], v+dw[t], ci] * dY[n,u,v,co]
 *
 * Both operands are read with the SAME 4-D TMA boxes as the forward pass ([pixels][channels], channels
 * contiguous) and fed to tcgen05.mma as MN-major matrices (M = ci, N = co, K = pixels), so no transposed
 * copy of the activations is ever made.  grid = (ci blocks x co blocks, taps, pixel splits); partial sums
 * are accumulated into the f32 HWIO gradient with atomic adds (the caller zeroes dW first).            */
typedef struct {
    int N;
    int n_maps;  kp_tap_view map[KP_MAX_MAPS];   /* X views (one per tap parity for stride 2) */
    kp_tap_view dy;                               /* dY view [N][Ho][Wo][Cout] */
    int n_taps;
    signed char dh[KP_MAX_TAPS], dw[KP_MAX_TAPS], map_first[KP_MAX_TAPS];
    int tap_flat[KP_MAX_TAPS];                   /* kh*k+kw of each tap (row of the HWIO kernel) */
    int CB;                                      /* channel block 16/32/64 (both operands) */
    int Ho, Wo;                                  /* pixel iteration space (= dY extent) */
    int Cin, Cout;                               /* rows / columns written */
    long long dw_off, dw_stap, dw_sci;           /* dW element (t,ci,co) at dw_off + tap_flat[t]*dw_stap + ci*dw_sci + co */
    int splits;                                  /* pixel-range splits (grid.z); 0 = library chooses */
} kp_wgrad_desc;

int kp_tapconv_wgrad_bf16(const kp_wgrad_desc* desc, const void* x, const void* dy, float* dw, void* stream);

/* Weight re-layout for the tap-GEMMs: fp32 HWIO kernel [k*k][cin][cout] -> bf16 K-major [rows_pad][Ktot]
 * (Ktot = T*Kper).  mode 0 (forward): row = co, K = taps x channel-concat segments each padded to the channel
 * block, optional per-output-channel scale (batch-norm folding).  mode 1 (data gradient): row = ci - c0,
 * K = taps x cout padded to the channel block.  Pure data movement + cast, one launch.                 */
typedef struct {
    int mode;                              /* 0 forward, 1 data gradient */
    int T; int tap_flat[KP_MAX_TAPS];      /* taps used, as rows kh*k+kw of the HWIO kernel */
    int cin, cout;                         /* source kernel dims */
    int nseg; int seg_start[3], seg_count[3], seg_kbase[3];   /* mode 0: source channel range -> K offset */
    int c0, rows;                          /* mode 1: first input channel and number of rows */
    int Kper, rows_pad, Ktot;
} kp_pack_desc;
int kp_pack_weights(const float* w, const kp_pack_desc* desc, const float* row_scale, void* dst, void* stream);
/* The same re-layout for MANY kernels in one launch (all convolutions of one optimiser are re-packed right after its
 * Adam step).  `jobs_dev` is a DEVICE array of n_jobs kp_pack_job (built once by the host, block_begin = exclusive
 * prefix sum of kp_pack_job_blocks() over the jobs, total_blocks = the sum).                                  */
typedef struct {
    const float* w;                        /* device: fp32 HWIO kernel */
    void* dst;                             /* device: bf16 [rows_pad][Ktot] */
    kp_pack_desc d;
    int block_begin, n_blocks;
} kp_pack_job;
int kp_pack_job_blocks(const kp_pack_desc* desc);
int kp_pack_weights_batch(const void* jobs_dev, int n_jobs, int total_blocks, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Memory-bound companions (bf16 NHWC, 8-channel vectors; fp32 only at the reference boundary)
 * ------------------------------------------------------------------------------------------- */

/* f32 image [P,3] -> bf16 [P,16]: out[c] = a[c]*x[perm[c]] + b[c] for c<3, zeros above.  Network inputs use the
 * identity; the VGG input folds (x+1)/2*255, RGB->BGR and the mean subtraction
 * (models/detector_translator_model.py:262-263, models/networks/vgg.py:17-19).  a, b, perm are HOST arrays. */
int kp_image_prep(const float* x, long long P, const float* a, const float* b, const int* perm, void* out, void* stream);
/* its adjoint: g bf16 [P,16] -> dx f32 [P,3] (accumulate != 0: +=). */
int kp_image_prep_bwd(const void* g, long long P, const float* a, const int* perm, int accumulate, float* dx, void* stream);

/* W-unrolled variant for the 3-channel first layers (encoder conv_1 7x7, models/networks/__init__.py:10; VGG conv1_1,
 * models/networks/vgg.py:20): out bf16 [N,H,W,Cpad], out[n,h,w,kw*3+c] = a[c]*x[n,h,w+kw-pad_left,perm[c]] + b[c]
 * (0 outside the image, Cpad in {16,32}).  A KHxKW conv over 3 channels then is a KHx1 conv over KW*3 channels with
 * the HWIO kernel reinterpreted in place as [KH][1][KW*3][Cout] (KH taps instead of KH*KW).  _bwd is the adjoint.  */
int kp_image_prep_unrolled(const float* x, int N, int H, int W, int KW, int pad_left, int Cpad, const float* a,
                           const float* b, const int* perm, void* out, void* stream);
int kp_image_prep_unrolled_bwd(const void* g, int N, int H, int W, int KW, int pad_left, int Cpad, const float* a,
                               const int* perm, int accumulate, float* dx, void* stream);

/* tf.contrib.layers.batch_norm (models/networks/layers.py:13-14), training mode, split around the convolution:
 * the conv epilogue accumulates per-channel sum / sum-of-squares of its PRE-bias accumulators
 * (kp_tapconv_bf16 stats_*); kp_bn_finalize turns them into scale/shift (normalising with the biased
 * variance), saves mean/rstd for the backward pass and updates the moving averages (unbiased variance,
 * moving = moving*decay + batch*(1-decay)); moving_mean/moving_var may be NULL (no update).          */
int kp_bn_finalize(const float* stats_sum, const float* stats_sq, const float* conv_bias, const float* gamma,
                   const float* beta, int C, double count, float eps, float decay, float* moving_mean, float* moving_var,
                   float* scale, float* shift, float* save_mean, float* save_rstd, void* stream);
/* y = relu?(x*scale + shift) (scale NULL = identity), optionally followed by tf.image.resize_images x2 (legacy
 * bilinear, models/networks/__init__.py:63,98): x bf16 [N,H,W,C] -> out bf16 [N,H,W,C] or [N,2H,2W,C].  */
int kp_bn_act_apply(const void* x, const float* scale, const float* shift, int relu, int upsample, int N, int H, int W,
                    int C, void* out, void* stream);
/* kp_bn_finalize + kp_bn_act_apply in ONE launch (the training-mode path of layers.batch_norm followed by tf.nn.relu,
 * models/networks/__init__.py:11-12 etc.): every block derives scale/shift from the raw sums, block 0 publishes
 * scale/shift/save_mean/save_rstd and updates the moving averages.  Same argument meaning as the two calls.  */
int kp_bn_stats_apply(const float* stats_sum, const float* stats_sq, const float* conv_bias, const float* gamma,
                      const float* beta, double count, float eps, float decay, float* moving_mean, float* moving_var,
                      float* scale, float* shift, float* save_mean, float* save_rstd, const void* x, int relu, int upsample,
                      int N, int H, int W, int C, void* out, int segments, void* stream);
/* segments > 1: the batch is `segments` equal runs of images that are normalised with their OWN statistics (two calls of
 * a shared-weight network executed as one launch, see kp_tapconv_desc.stat_groups): stats_sum / stats_sq / scale / shift /
 * save_mean / save_rstd are [segments][C], count is the pixel count of ONE segment, and the moving averages receive one
 * update per segment, in order.                                                                           */
/* backward of kp_bn_act_apply + batch-norm statistics: dout (grad of the output, upsampled size if upsample),
 * x (the conv output saved by the forward) -> dbeta, dgamma f32 [C] (this call's sums; zeroed by the library
 * unless prezeroed != 0) and dx bf16 [N,H,W,C].  gbeta_acc / ggamma_acc (nullable): parameter-gradient buffers
 * that additionally receive += dbeta / dgamma.                                                          */
int kp_bn_act_bwd(const void* dout, const void* x, const float* scale, const float* shift, const float* save_mean,
                  const float* save_rstd, int relu, int upsample, int N, int H, int W, int C, float* dbeta, float* dgamma,
                  void* dx, float* gbeta_acc, float* ggamma_acc, int prezeroed, int segments, void* stream);
/* (segments as in kp_bn_stats_apply: scale / shift / save_mean / save_rstd / dbeta / dgamma are [segments][C]; gbeta_acc /
 * ggamma_acc receive the sum over the segments.)                                                              */
/* adjoint of tf.image.resize_images x2 (legacy bilinear, models/networks/__init__.py:63,98) alone:
 * dout bf16 [N,2H,2W,C] -> dact bf16 [N,H,W,C].  The BN backward of the upsampling layers runs this once and then the
 * plain (upsample = 0) kp_bn_act_bwd on dact.                                                                 */
int kp_upsample2x_bwd(const void* dout, int N, int H, int W, int C, void* dact, void* stream);
/* g = dy * (y > 0 ? 1 : alpha): backward of the fused bias+ReLU / bias+leaky_relu epilogues. */
int kp_act_mask_bwd(const void* dy, const void* y, float alpha, long long n_elems, void* g, void* stream);
/* tf.nn.max_pool 2x2 s2 (models/networks/vgg.py:45-46) and its backward (gradient to the first maximum; with
 * relu_mask the ReLU mask of the pooled tensor's producer is applied in the same pass).               */
int kp_maxpool2x2_fwd(const void* x, int N, int H, int W, int C, void* out, void* stream);
int kp_maxpool2x2_bwd(const void* dy, const void* x, int relu_mask, int N, int H, int W, int C, void* dx, void* stream);
/* final = im*mask + crude*(1-mask) (models/detector_translator_model.py:174; clip != 0 adds the clip_by_value
 * of models/final_model.py:98-99).  heads f32 [P,4] = (crude, sigmoid(mask)) from the fused head conv.  */
int kp_mask_compose_fwd(const float* heads, const float* im, long long P, int clip, float* final_out, float* crude_out,
                        float* mask_out, void* stream);
/* d_final f32 [P,3] -> gradient w.r.t. the head pre-activations, bf16 [P,16] (channels 4..15 zero). */
int kp_mask_compose_bwd(const float* d_final, const float* heads, const float* im, long long P, void* d_heads, void* stream);
/* tf.concat on channels with cast to bf16 and zero padding to Ctot (joint embedding,
 * models/detector_translator_model.py:170), and its adjoint.  src/C/is_f32 are HOST arrays.          */
int kp_pack_channels(const void* const* src, const int* C, const int* is_f32, int n, long long P, int Ctot, void* out,
                     void* stream);
int kp_unpack_channels(const void* g, long long P, int Ctot, void* const* dst, const int* C, const int* is_f32, int n,
                       void* stream);
/* L1 feature loss between the gt and the pred VGG features (two bf16 tensors of n_elems each;
 * models/detector_translator_model.py:280-287): *loss += weight*mean|gt-pred|; d_pred (nullable) receives
 * weight/count*sign(pred-gt).                                                                        */
int kp_l1_pair_fwd_bwd(const void* feat_gt, const void* feat_pred, long long n_elems, float weight, float* loss,
                       void* d_pred, void* stream);
/* sigmoid_cross_entropy_with_logits vs a constant label (models/detector_translator_model.py:249-254,265-267):
 * *loss += weight*mean(...); d_logits (nullable) bf16 [n,8], channel 0 = weight/n*(sigmoid(x)-label). */
int kp_bce_logits_fwd_bwd(const float* logits, int n, float label, float weight, float* loss, void* d_logits, void* stream);
/* tf.train.AdamOptimizer update over one flat buffer (models/detector_translator_model.py:198-202):
 * lr_t = lr*sqrt(1-beta2^t)/(1-beta1^t); p -= lr_t*m/(sqrt(v)+eps); grads are multiplied by grad_scale first.
 * lr_t_dev (nullable): device scalar holding a precomputed lr_t that overrides lr/t (CUDA-graph replays).     */
int kp_adam_tf(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
               int t, float grad_scale, const float* lr_t_dev, void* stream);
/* 1x1 convolution of a narrow bf16 activation into fp32: out[p][co] = bias[co] + sum_ci x[p][ci] * w[ci][co]
 *   replaces: the detector head layers.conv(x, n_pts, kernel=1, stride=1) 'conv_0' (models/networks/__init__.py:57-59),
 *             whose logits feed get_coord in fp32.  x bf16 [P,Cin] (Cin = 16), w f32 [Cin,Cout] (the HWIO kernel of a 1x1
 *             convolution as stored; rounded to bf16 on load like every convolution weight of this library), bias f32 [Cout]
 *             (nullable), out f32 [P,Cout], Cout a multiple of 4 up to 40.
 *   HBM-bound (7 FLOP/B): CUDA-core FMAs, fully coalesced 16-byte stores through a shared-memory transpose.  */
int kp_conv1x1_f32(const void* x, const float* w, const float* bias, long long P, int Cin, int Cout, float* out, void* stream);
/* out[c] += sum over pixels of g bf16 [P,C] (bias gradients). */
int kp_channel_sum(const void* g, long long P, int C, float* out, void* stream);
/* out[c] += sum over pixels of g[p,c]^2: with kp_channel_sum the batch statistics of a stand-alone
 * layers.batch_norm (models/networks/layers.py:13-14) whose input does not come out of a convolution epilogue. */
int kp_channel_sumsq(const void* g, long long P, int C, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Input pipeline: the per-frame Pillow chain of the reference's loaders on the device
 * ------------------------------------------------------------------------------------------- */
#define KP_AUG_SIZE 128 /* IMAGE_SIZE of data/image_pair_dataloader.py:13, data/keypoint_dataloader.py:13 */

/* What happens to ONE decoded frame (tightly packed HWC uint8 RGB inside the caller's source buffer).  Filled on the host
 * by kp_augment_plan_host / kp_augment_plan_zero_host, copied to the device by the caller as an array.            */
typedef struct kp_frame_plan {
    long long src_offset;      /* byte offset of the frame in the source buffer */
    int src_w, src_h;
    int rotate;                /* 1: Image.rotate's nearest-neighbour inverse affine map, 16.16 fixed point */
    int a[6];                  /* its coefficients (Geometry.c affine_fixed: a0 a1 a2 / a3 a4 a5, centres folded in) */
    int filter_id;             /* -1 none; 0..5 DETAIL, EDGE_ENHANCE, SMOOTH, SMOOTH_MORE, EDGE_ENHANCE_MORE, BLUR;
                                  6..9 ImageEnhance Sharpness, Brightness, Color, Contrast (utils/data.py:11-33) */
    int zero;                  /* 1: an all-zero frame (data/keypoint_dataloader.py:77-80), nothing is read */
    float factor;              /* enhancement factor of filter_id 6..9 (r_val * 0.1) */
    short xtab[KP_AUG_SIZE];   /* column / row of the (rotated) source frame behind each output column / row after */
    short ytab[KP_AUG_SIZE];   /* resize, crop and flip; -1 = outside (stays 0) */
} kp_frame_plan;

/* HOST function (no CUDA call): the plan of
 *   image.rotate(angle_deg).resize([resize_w, resize_h]).crop((crop_left, crop_top, crop_left + 128, crop_top + 128))
 *        [.transpose(FLIP_LEFT_RIGHT)] -> apply_random_filter's branch filter_id with factor
 *   replaces: data/image_pair_dataloader.py:95-159 (angle_deg 0 and filter_id -1 give the randomness=False branch and
 *             data/keypoint_dataloader.py:71 `im.resize(...).crop(crop_size)` with utils/data.py:38-59 center_crop).
 * Pillow 6.2.0 semantics (the reference's pin): rotate and resize resample NEAREST, the crop box goes through
 * int(round()) (half to even), pixels outside the rotated / resized frame are 0.                                     */
int kp_augment_plan_host(kp_frame_plan* plan, long long src_offset, int src_w, int src_h, int resize_w, int resize_h,
                         double crop_left, double crop_top, int angle_deg, int flip, int filter_id, double factor);
/* HOST function: the plan of one zero frame of the keypoint loader's padding (data/keypoint_dataloader.py:77-80). */
int kp_augment_plan_zero_host(kp_frame_plan* plan);
/* HOST function: n plans in one call from parallel arrays (zero[i] != 0: a zero frame, the other fields of i are ignored);
 * what a loader calls once per batch.                                                                               */
int kp_augment_plan_batch_host(kp_frame_plan* plans, int n, const long long* src_offset, const int* src_w, const int* src_h,
                               const int* resize_w, const int* resize_h, const double* crop_left, const double* crop_top,
                               const int* angle_deg, const int* flip, const int* filter_id, const double* factor,
                               const int* zero);

/* HOST functions: page-lock / release a host range the CALLER owns (the loader's staging buffer is a shared mapping its
 * decode worker processes write into, so it cannot come from cudaHostAlloc).  KP_ERR_CUDA when CUDA refuses; no error
 * state is left behind in that case.                                                                                */
int kp_host_register(void* host_ptr, unsigned long long bytes);
int kp_host_unregister(void* host_ptr);

/* src: device buffer of decoded frames; plans: DEVICE array [n_frames]; out f32 [n_frames,128,128,3] =
 * float32(pixel / 255.0) * 2 - 1 (`image / 255.0` of data/image_pair_dataloader.py:162-165 followed by map_fn :63-69).
 * One launch, byte-exact against Pillow; 49 152 B gathered + 196 608 B written per frame (HBM-bound).               */
int kp_augment_frames(const unsigned char* src, const kp_frame_plan* plans, int n_frames, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KP_B200_H */
