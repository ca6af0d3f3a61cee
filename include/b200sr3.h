/*
 * b200sr3 — C ABI of the B200 (sm_100a) SR3 reverse-diffusion sampler.
 *
 * This is the drop-in boundary for ONE hot path of
 * zouiner/3d-super-resolution-Face-reconstruction: GaussianDiffusion.super_resolution /
 * p_sample_loop driving the model/sr UNet denoiser. The reference has no FFI of its own (it is
 * pure PyTorch), so every entry point below names the reference Python interface it replaces
 * (paths relative to the reference root). Host code (Python, via ctypes) owns all tensors; the
 * library owns only its packed weights, tables and per-(B,R) workspace.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; b200sr3_last_error() then
 *     returns a thread-local, NUL-terminated description.
 *   - "device pointer" = CUDA global memory on the handle's device (torch: tensor.data_ptr()).
 *   - images are fp32, NCHW, contiguous, range [-1, 1], exactly as the reference passes them.
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).
 *   - no function falls back to the CPU; without a usable sm_100 device create() fails.
 */
#ifndef B200SR3_H_
#define B200SR3_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200SR3_API __attribute__((visibility("default")))
#else
#define B200SR3_API
#endif

#define B200SR3_ABI_VERSION 2
#define B200SR3_MAX_LEVELS 8

typedef struct b200sr3_handle b200sr3_handle;

/* The keys define_G reads from opt['sr']['model'] (model/sr/networks.py:83-101). */
typedef struct b200sr3_config {
  int32_t in_channel;                         /* unet.in_channel  (6 = cond 3 + x 3)            */
  int32_t out_channel;                        /* unet.out_channel (3)                           */
  int32_t inner_channel;                      /* unet.inner_channel (64)                        */
  int32_t norm_groups;                        /* unet.norm_groups, 32 when absent (networks.py:89-90) */
  int32_t res_blocks;                         /* unet.res_blocks (2)                            */
  int32_t n_mults;                            /* len(unet.channel_multiplier)                   */
  int32_t channel_mults[B200SR3_MAX_LEVELS];  /* unet.channel_multiplier ([1,2,4,8,8])          */
  int32_t n_attn_res;
  int32_t attn_res[B200SR3_MAX_LEVELS];       /* unet.attn_res ([16])                           */
  int32_t image_size;                         /* diffusion.image_size: only places attention
                                                 (model/sr/sr3_modules/unet.py:192-197,211-220) */
  int32_t conditional;                        /* diffusion.conditional                          */
} b200sr3_config;

/* Noise source of one sampling call (diffusion.py:186,205 draw from torch's global RNG). */
enum {
  B200SR3_NOISE_INJECTED = 1, /* caller supplies [x_T, z_{T-1}, ..., z_1]: T tensors of B*3*R*R */
  B200SR3_NOISE_PHILOX = 2    /* in-kernel Philox4x32-10 keyed by (seed, t, element)            */
};

B200SR3_API const char* b200sr3_last_error(void);
B200SR3_API int b200sr3_abi_version(void);

/* Replaces unet.UNet(...) + diffusion.GaussianDiffusion(...) construction inside
 * define_G (model/sr/networks.py:91-109). Fails if `device` is not compute capability 10.x. */
B200SR3_API int b200sr3_create(const b200sr3_config* cfg, int device, b200sr3_handle** out);
B200SR3_API int b200sr3_destroy(b200sr3_handle* h);

/* Weight ingestion: replaces netG.load_state_dict (lib/trainer_temp.py:172-216,
 * model/sr/model.py:164-195). `key` is the reference state_dict key WITHOUT the
 * "denoise_fn." prefix (e.g. "downs.1.res_block.block1.block.3.weight"); `data` is fp32,
 * host or device, in the reference's layout (conv OIHW, linear [out,in]). The expected key
 * set can be enumerated with b200sr3_num_tensors / b200sr3_tensor_info. */
B200SR3_API int b200sr3_num_tensors(b200sr3_handle* h);
B200SR3_API int b200sr3_tensor_info(b200sr3_handle* h, int index, const char** key, int64_t shape[4], int* ndim);
B200SR3_API int b200sr3_load_tensor(b200sr3_handle* h, const char* key, const float* data,
                        const int64_t* shape, int ndim);
/* Repack to the device layout (bf16, [Cout][tap][Cin] K-major; 1x1 res_conv appended as extra
 * K columns of block2's conv) and build the noise-embedding weight matrix. Fails if a tensor
 * is missing. */
B200SR3_API int b200sr3_finalize_weights(b200sr3_handle* h, void* stream);

/* Replaces GaussianDiffusion.set_new_noise_schedule (diffusion.py:93-142). The caller computes
 * the float64 tables exactly as the reference does and passes HOST arrays:
 * sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, posterior_mean_coef1/2,
 * posterior_log_variance_clipped (fp32[T]) and sqrt_alphas_cumprod_prev (float64[T+1]).
 * Builds the per-timestep noise-embedding bias table (unet.py:18-50,179-184) on the device. */
B200SR3_API int b200sr3_set_schedule(b200sr3_handle* h, int T, const float* sqrt_recip_ac,
                         const float* sqrt_recipm1_ac, const float* coef1, const float* coef2,
                         const float* post_logvar, const double* sqrt_ac_prev, void* stream);

/* Replaces UNet.forward(cat([cond, x], 1), noise_level) (unet.py:235-265, called at
 * diffusion.py:170) with one scalar noise level for the whole batch. Device pointers;
 * eps is fp32 [B, out_channel, R, R]. cond may be NULL when !conditional. */
B200SR3_API int b200sr3_unet_forward(b200sr3_handle* h, const float* cond, const float* x, float noise_level,
                         int B, int R, float* eps, void* stream);

/* Replaces GaussianDiffusion.p_sample(x, t, clip_denoised, condition_x=cond) (diffusion.py:182-187) with
 * the step's noise given (teacher-forced parity). `noise` may be NULL (treated as zeros; it is
 * ignored at t == 0 as in the reference). clip_denoised != 0 clamps x0 to [-1,1] (diffusion.py:175-176, the
 * default and what p_sample_loop always uses); 0 leaves it unclamped. Device pointers. */
B200SR3_API int b200sr3_step(b200sr3_handle* h, const float* cond, const float* x_t, const float* noise,
                 int t, int clip_denoised, int B, int R, float* x_tm1, void* stream);

/* Replaces GaussianDiffusion.p_sample_loop / super_resolution (diffusion.py:189-225), all T
 * steps, batched. out: fp32 [B,3,R,R] = x after t = 0. snapshots (optional): fp32
 * [n_snap,B,3,R,R], x after every t with t % (1 | T/10) == 0 in visiting order, which is what
 * `continous=True` concatenates after cond. noise_mode INJECTED reads `noise` (T*B*3*R*R
 * floats); PHILOX ignores it and uses `seed`. Device pointers.
 * `row_offset` (PHILOX): global index of this call's batch row 0. The stream is keyed by (seed, t, GLOBAL row, y, x),
 * so a face's noise - like the reference's per-image torch.randn draws - does not depend on how one logical batch is
 * sharded over GPUs or chunked over calls: rank g of a sharded run passes its shard start, a chunked caller passes the
 * chunk start, and all of them pass the same seed. 0 for a stand-alone batch. */
B200SR3_API int b200sr3_sample(b200sr3_handle* h, const float* cond, int noise_mode, const float* noise,
                   uint64_t seed, int64_t row_offset, int B, int R, float* out, float* snapshots, void* stream);
B200SR3_API int b200sr3_num_snapshots(b200sr3_handle* h);

/* The sampler's own N(0,1) stream, as the PHILOX mode of b200sr3_sample draws it: out fp32 [B,out_channel,R,R] =
 * the draw keyed (seed, t, rows row_offset .. row_offset+B-1). t = T (the schedule length) is x_T, the
 * `torch.randn(shape)` of diffusion.py:205 (:196 unconditional); 0 < t < T is z_t, the `torch.randn_like(x)` of
 * diffusion.py:186. Used by the statistical tests of the stream and to return x_T where the reference does
 * (unconditional continous=True, diffusion.py:195-196). Device pointer. */
B200SR3_API int b200sr3_philox_normal(b200sr3_handle* h, uint64_t seed, int t, int64_t row_offset, int B, int R,
                   float* out, void* stream);

/* The same call for HOST buffers (the end-to-end path bench.py times as `e2e`): cond is copied
 * host->device, the chain runs with PHILOX noise, out is copied device->host, and the call
 * returns after the copy has completed. cond/out should be pinned for full PCIe speed. */
B200SR3_API int b200sr3_sample_host(b200sr3_handle* h, const float* cond_host, uint64_t seed, int64_t row_offset,
                        int B, int R, float* out_host, void* stream);

/* Introspection for layer-level parity tests: copy the activation a named layer produced in
 * the most recent forward ("downs.0" ... "ups.18", "mid.0", "mid.1"; unet.py module names) to
 * fp32 NCHW. *C/*H/*W are filled in; `dst` may be NULL to query the shape only. */
B200SR3_API int b200sr3_layer_output(b200sr3_handle* h, const char* layer, float* dst, int* C, int* H, int* W,
                         void* stream);

/* Counters: kernels launched by the library inside the last step / sample call, and the number
 * of conv (tcgen05) launches among them. */
B200SR3_API int b200sr3_last_launch_count(b200sr3_handle* h, int64_t* total, int64_t* conv);

/* Measurement aid for bench.py's roofline: runs ONE sampling step of a (B,R) batch eagerly with
 * a CUDA event between consecutive launches and reports, per launch in chain order, the device
 * time (ms), the algorithmic FLOPs (convs: 2*MAC on the REFERENCE graph, SURVEY.md 8d - full-resolution Upsample
 * convs, no identity-shortcut segments), the FLOPs the tensor pipe actually executes (optional), the algorithmic HBM
 * bytes (HBM-bound kernels), and the newline-separated op names. Not used on the sampling path. */
B200SR3_API int b200sr3_profile_step(b200sr3_handle* h, int B, int R, int max_ops, float* ms, double* flops,
                                     double* flops_executed, double* bytes, char* names, int names_len, int* n_ops,
                                     void* stream);

/* Kernel-level entry used by the conv parity tests and by bench.py's roofline probe: one
 * implicit-GEMM convolution on the tcgen05 path. x: fp32 NCHW [B,Cin,H,W]; w: fp32 OIHW
 * [Cout,Cin,k,k] (k = 1 or 3, pad = k/2); stride 1 or 2; upsample2x applies a nearest 2x
 * before the conv (unet.py:58-65); optional residual (fp32 NCHW, output shape) is added in
 * the epilogue. Operands are rounded to bf16, accumulation is fp32, y is fp32 NCHW.
 * If iters > 0 the conv kernel alone is launched `iters` more times and the average device
 * time in milliseconds is written to *avg_ms (CUDA events on `stream`). */
B200SR3_API int b200sr3_conv2d(int device, const float* x, const float* w, const float* bias,
                   const float* residual, int B, int Cin, int H, int W, int Cout, int k,
                   int stride, int upsample2x, float* y, int iters, float* avg_ms, void* stream);

/* Kernel-level entry for the halo-resident conv (csrc/conv_halo.cuh), used by the parity tests and
 * tools/conv_bench.py: y = conv3x3([swish](GroupNorm(cat(x0, x1)))) + bias + conv1x1(cat(r0, r1)),
 * i.e. one reference Block (unet.py:80-91: GN -> Swish -> Conv) with the ResnetBlock shortcut
 * (unet.py:103-110) folded in as extra K segments, and the channel concat of unet.py:261 read as
 * two sources. x0/x1/r0/r1: fp32 NCHW [B,C*,H,W] (x1, r0, r1 optional: pass NULL and 0 channels);
 * gamma/beta: [C0+C1] or NULL for no GroupNorm; w: OIHW [Cout,C0+C1,3,3]; wres: [Cout,Cr0+Cr1,1,1];
 * resample = 1 applies nearest 2x to the (un-normalised) input first (Upsample, unet.py:58-65; no GN, no
 * shortcut); resample = 2 runs the conv with stride 2 (Downsample, unet.py:68-74; one raw source, y is
 * [B,Cout,H/2,W/2] and the shape rule applies to H/2, W/2); 0 = neither. stats_out (optional): [B][Cout][2] per-(image, channel) sum and sum of squares of y as
 * the fused GroupNorm statistics report them. H % 16 == 0, W % 8 == 0, W >= 16, channels % 64 == 0. */
B200SR3_API int b200sr3_conv_block(int device, const float* x0, int C0, const float* x1, int C1,
                   const float* gamma, const float* beta, int groups, int swish, const float* w,
                   const float* bias, const float* r0, int Cr0, const float* r1, int Cr1,
                   const float* wres, int B, int H, int W, int Cout, int resample, float* y,
                   float* stats_out, int iters, float* avg_ms, void* stream);

/* ---- SR -> MICA hand-off on the device. Stand-alone entry points (no handle): all pointers are
 * device pointers, caller-owned; results are bit-identical to the reference's host round trip.
 *
 * b200sr3_tensor2img replaces core/metrics.py:16-42 (tensor2img): x fp32 NCHW [B,C,H,W] (C = 1 or 3, any range) ->
 * img uint8 NHWC [B,H,W,C] = round_half_even(((clamp(x,-1,1)+1)/2)*255), image by image (the reference's 4-D branch
 * builds a make_grid mosaic instead; callers on this path pass single images). */
B200SR3_API int b200sr3_tensor2img(const float* x, int B, int C, int H, int W, uint8_t* img, void* stream);

/* b200sr3_mica_handoff replaces model/sr3d/model.py:372-382 (and :484-485): for each uint8 RGB image [R,R,3]
 *   up224        = cv2.resize(img, (224,224))                          uint8 NHWC [B,224,224,3]   (optional, may be NULL)
 *   image224     = up224 / 255. as CHW                                 fp32 NCHW [B,3,224,224]    (optional; the reference
 *                  holds this in float64, the consumer casts to fp32)
 *   arcface_blob = cv2.dnn.blobFromImages([up224], 1/127.5, (112,112), (127.5,)*3, swapRB=True)[0] (:127-131)
 *                                                                      fp32 NCHW [B,3,112,112]    (optional)
 * OpenCV semantics (INTER_LINEAR fixed point; the 2x down-scale inside blobFromImages is a 2x2 box mean) are
 * reproduced bit for bit. R = 448 (OpenCV's INTER_AREA shortcut on the first resize) is rejected. */
B200SR3_API int b200sr3_mica_handoff(const uint8_t* img, int B, int R, uint8_t* up224, float* image224,
                   float* arcface_blob, void* stream);

/* b200sr3_tensor_blob replaces the model3 variant, model/sr3d/model.py:477-481: create_tensor_blob
 * (:105-124) applied to core/metrics.py:44-50 tensor2tensor_img(x) * 255, i.e. float bilinear
 * (align_corners=False) resize of (clamp(x)-derived pixel - 127.5)/127.5 to 112x112 with R and B swapped.
 * x fp32 NCHW [B,3,R,R] -> arcface_blob fp32 NCHW [B,3,112,112]; float path, parity within 1e-5. */
B200SR3_API int b200sr3_tensor_blob(const float* x, int B, int R, float* arcface_blob, void* stream);

/* ---- MICA identity encoder on the device: everything between the ArcFace blob above and the
 * FLAME decoder. Replaces, for a whole batch,
 *   codedict['arcface'] = F.normalize(self.arcface(arcface_imgs))      model/sr3d/model.py:164-170 (encode_mica)
 *   shape = self.regressor(arcface)                                     model/mica/generator.py:86-88
 * with self.arcface = Arcface() (iResNet-100, model/mica/arcface.py:165-200, eval mode: BatchNorm running statistics,
 * no dropout) and self.regressor = MappingNetwork(z_dim, map_hidden_dim, n_shape, map_layers)
 * (model/mica/generator.py:31-60; the reference builds it with 512 / 300 / 300 / 3, model/sr3d/model.py:68-75).
 * The convs run in bf16 with fp32 accumulation; stated tolerance in tests/test_gpu_arcface.py.
 *
 * Weights: `key` is "arcface." + the reference Arcface state_dict key (e.g. "arcface.layer3.17.bn2.running_var") or
 * "regressor." + the MappingNetwork key ("regressor.network.0.weight", "regressor.output.bias"); fp32, host or device,
 * reference layouts. "...num_batches_tracked" entries are accepted and ignored. finalize folds every BatchNorm that
 * follows a conv into that conv and packs the bf16 operands. */
typedef struct b200sr3_mica b200sr3_mica;
B200SR3_API int b200sr3_mica_create(int device, int z_dim, int map_hidden_dim, int map_layers, int n_shape,
                   b200sr3_mica** out);
B200SR3_API int b200sr3_mica_destroy(b200sr3_mica* h);
B200SR3_API int b200sr3_mica_num_tensors(b200sr3_mica* h);
B200SR3_API int b200sr3_mica_tensor_info(b200sr3_mica* h, int index, const char** key, int64_t shape[4], int* ndim);
B200SR3_API int b200sr3_mica_load_tensor(b200sr3_mica* h, const char* key, const float* data, const int64_t* shape,
                   int ndim);
B200SR3_API int b200sr3_mica_finalize_weights(b200sr3_mica* h, void* stream);
/* arcface_blob: fp32 [B,3,112,112] as b200sr3_mica_handoff / cv2.dnn.blobFromImages produce it (host or device).
 * Outputs, each optional (NULL to skip), host or device: embedding [B,512] = Arcface.forward (arcface.py:179-200),
 * identity [B,512] = F.normalize(embedding), shape_code [B,n_shape] = the regressor's output. */
B200SR3_API int b200sr3_mica_encode(b200sr3_mica* h, const float* arcface_blob, int B, float* embedding,
                   float* identity, float* shape_code, void* stream);
/* Introspection for layer-level parity tests: the residual stream after "stem" or a block ("layer1.0" ... "layer4.2",
 * arcface.py module names) of the most recent encode, as fp32 NCHW. */
B200SR3_API int b200sr3_mica_layer_output(b200sr3_mica* h, const char* layer, float* dst, int* C, int* H, int* W,
                   void* stream);
/* Measurement aid: one eager pass with a CUDA event between launches (per-launch ms, algorithmic FLOPs, names) and
 * the launch counts of the last encode. */
B200SR3_API int b200sr3_mica_profile(b200sr3_mica* h, int B, int max_ops, float* ms, double* flops, char* names,
                   int names_len, int* n_ops, int64_t* launches, int64_t* conv_launches, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SR3_H_ */
