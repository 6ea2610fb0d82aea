/* r2l_b200 — C ABI of the B200-native R2L hot path.
 *
 * The reference (snap-research/R2L) has no FFI: its hot path is the Python module
 * model/nerf_raybased.py executed by stock PyTorch.  Each entry point below replaces the chain of
 * ATen/cuBLAS launches the cited reference lines produce; the Python shim in r2l_b200/ binds them with
 * ctypes.  All pointers are DEVICE pointers unless marked "host"; the library never
 * allocates, never synchronises, and enqueues on the CUDA stream passed as `stream` (a cudaStream_t).
 * Return value: 0 on success, negative on error; r2l_last_error() gives the message for the calling thread.
 */
#ifndef R2L_B200_H_
#define R2L_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define R2L_NUM_PARAMS 5917187 /* floats in the flat state_dict-ordered parameter buffer */

/* input kinds of r2l_forward */
#define R2L_INPUT_RAYS 0 /* in0 = rays_o[N,3], in1 = rays_d[N,3]  (PointSampler.sample_train, nerf_raybased.py:114-126) */
#define R2L_INPUT_PTS 1  /* in0 = pts[N,48]   (output of PointSampler.sample_*, :94-126)                       */
#define R2L_INPUT_X 2    /* in0 = x[N,1008]   (output of PositionalEmbedder.__call__, :198-208)                 */
#define R2L_INPUT_RAYS9 4 /* in0 = rays9[N,9] = (o | d | rgb) rows as the ray shards store them (utils/create_data.py:820-872,
                           * main.py:1305-1311): o and d are read in place at stride 9, in1 is ignored                */

const char* r2l_last_error(void);
int r2l_abi_version(void);

/* Bytes of the packed-weight buffer / of the forward workspace for n_rays. */
size_t r2l_packed_bytes(void);
size_t r2l_fwd_workspace_bytes(int64_t n_rays);

/* params[R2L_NUM_PARAMS] fp32 in state_dict order (head.0.weight, head.0.bias, body.k.body.{0,2}.{weight,bias},
 * tail.0.weight, tail.0.bias; NeRF_v3_2.__init__, nerf_raybased.py:483-537)  ->  fp16 hi/lo tensor-core operand
 * images + fp32 bias tables in `packed`.  Call again after every parameter update. */
int r2l_pack_weights(const float* params, void* packed, void* stream);

/* rgb[N,3] = NeRF_v3_2.forward(PositionalEmbedder(PointSampler.sample(...)))   (nerf_raybased.py:539-544)
 *   z_lo, z_diff : HOST pointers to 16 floats each (only for R2L_INPUT_RAYS): z_vals (t_rand == NULL) or
 *                  `lower` and `upper - lower` of sample_train's stratified jitter (:118-123)
 *   t_rand       : device [N,16] uniform numbers or NULL
 *   workspace    : device scratch of r2l_fwd_workspace_bytes(n_rays) bytes */
int r2l_forward(int input_kind, const float* in0, const float* in1, const float* t_rand, const float* z_lo,
                const float* z_diff, const void* packed, float* rgb, void* workspace, size_t workspace_bytes,
                int64_t n_rays, void* stream);

/* Pose in -> frame out: for every pose c2w[p] (DEVICE [P,3,4] row-major camera-to-world) and pixel
 * (row j, column i) of a height x width frame, the ray PointSampler.__init__/sample_test builds (nerf_raybased.py:80-86,
 * :94-102: dirs = [(i - W/2)/f, -(j - H/2)/f, -1], rays_d = sum_k dirs_k c2w[:, k], rays_o = c2w[:, 3]) goes through the
 * fused forward; nothing but the poses is read and nothing but the image is written (the reference materialises dirs,
 * pts [H W,48] and the encoding [H W,1008] per frame, main.py:300-309).  z_vals: HOST pointer to the 16 depths.
 *   rgb  : [P,H,W,3] fp32 or NULL;  rgb8 : [P,H,W,3] uint8 = to8b(rgb) (nerf_raybased.py:16, main.py:338) or NULL. */
int r2l_render_poses(const float* c2w, int64_t n_poses, int height, int width, float focal, const float* z_vals,
                     const void* packed, float* rgb, uint8_t* rgb8, void* workspace, size_t workspace_bytes, void* stream);

/* ---- training: fused forward that keeps what the backward needs, and the fused backward ----
 * Replaces loss.backward() through NeRF_v3_2 (main.py:1404; autograd of nerf_raybased.py:539-544).
 *   zf        : [N,256] fp32 out, z_43 + h (input of the tail Linear)
 *   fwd_saved : r2l_train_fwd_saved_bytes(n) bytes; the fp16 hi/lo input operand of every Linear
 *   bwd_saved : r2l_train_bwd_saved_bytes(n) bytes; the fp16 hi/lo output-gradient operand of every Linear
 *   grads     : [R2L_NUM_PARAMS] fp32 in state_dict order, OVERWRITTEN with dL/dparams given grad_rgb = dL/drgb
 * `workspace`: r2l_bwd_workspace_bytes(n) bytes.  r2l_backward enqueues three kernels on `stream`; when the chain
 * grid leaves >= 90 SMs idle the weight-gradient kernel runs concurrently on an internal side stream (joined back
 * into `stream` before the call returns control of the stream order). */
size_t r2l_bwd_workspace_bytes(int64_t n_rays);   /* workspace of r2l_backward (>= the forward's) */
size_t r2l_train_fwd_saved_bytes(int64_t n_rays);
size_t r2l_train_bwd_saved_bytes(int64_t n_rays);
int r2l_forward_train(int input_kind, const float* in0, const float* in1, const float* t_rand, const float* z_lo,
                      const float* z_diff, const void* packed, float* rgb, float* zf, void* fwd_saved,
                      void* workspace, size_t workspace_bytes, int64_t n_rays, void* stream);
int r2l_backward(int input_kind, const void* packed, const float* rgb, const float* grad_rgb, const float* zf,
                 const void* fwd_saved, void* bwd_saved, float* grads, void* workspace, size_t workspace_bytes,
                 int64_t n_rays, void* stream);

/* The same backward with the gradient buffer completed in `n_chunks` pieces, TOP of the buffer first (gradients complete
 * tail -> head): chunk 0 = body layers >= split_layers[0] and the tail, chunk i = body layers in
 * [split_layers[i], split_layers[i-1]), last chunk = body layers < split_layers[n_chunks-2] and the head
 * (split_layers: n_chunks - 1 strictly descending body-layer indices in (0, 86); r2l_grad_chunk_range gives the float range
 * [lo, hi) of a chunk in the flat buffer).  Each chunk but the last is followed by an event on the library's side: a
 * communication stream that calls r2l_stream_wait_grad_chunk(i, stream) can all-reduce chunk i while the backward still
 * runs (data-parallel training: replaces the reference's DataParallel reduce-to-GPU-0, main.py:472-479, by chunked,
 * overlapped all-reduces of ONE flat buffer).  The last chunk is complete when the call's work on `stream` is.
 * reserve_sms: SMs the weight-gradient launches leave to the collective's kernel.  Capturable in a CUDA graph. */
int r2l_backward_chunked(int input_kind, const void* packed, const float* rgb, const float* grad_rgb, const float* zf,
                         const void* fwd_saved, void* bwd_saved, float* grads, void* workspace, size_t workspace_bytes,
                         int64_t n_rays, void* stream, int n_chunks, const int* split_layers, int reserve_sms);
int r2l_grad_chunk_range(int n_chunks, const int* split_layers, int chunk, int64_t* lo, int64_t* hi);
int r2l_stream_wait_grad_chunk(int chunk, void* stream);

/* Teacher NeRF (NeRF.forward :377-401 behind run_network :312-334; D=8, W=256, skips=[4], use_viewdirs, multires 10/4).
 *   params : R2L_TEACHER_NUM_PARAMS floats, state_dict order (pts_linears.0-7, views_linears.0, feature_linear,
 *            alpha_linear, rgb_linear; weight then bias)
 *   input  : pts[P,3] + viewdirs[P/samples_per_ray,3] (embeddings built in-kernel), or x_embedded[P,90]
 *   raw    : [P,4] = (rgb, sigma), no output activation */
#define R2L_TEACHER_NUM_PARAMS 595844
size_t r2l_teacher_packed_bytes(void);
int r2l_teacher_pack_weights(const float* params, void* packed, void* stream);
int r2l_teacher_forward(const float* pts, const float* viewdirs, const float* x_embedded, const void* packed, float* raw,
                        int64_t n_points, int64_t samples_per_ray, void* stream);

/* The same query with the sample points built in the kernel: point (r, s) = rays_o[r] + rays_d[r] * z_vals[r, s]
 * (render_rays, utils/create_data.py:486-487 coarse, :517 fine: the reference materialises pts[N,S,3] first).
 * rays_o, rays_d, viewdirs: [N,3]; z_vals: [N,samples_per_ray]; raw: [N,samples_per_ray,4]. */
int r2l_teacher_forward_rays(const float* rays_o, const float* rays_d, const float* viewdirs, const float* z_vals, const void* packed,
                             float* raw, int64_t n_rays, int64_t samples_per_ray, void* stream);

/* raw2outputs (nerf_raybased.py:226-295, raw_noise_std = 0): raw[N,S,4], z_vals[N,S], rays_d[N,3] ->
 * rgb_map[N,3], disp_map[N], acc_map[N], weights[N,S], depth_map[N].  One warp per ray, one pass over HBM. */
int r2l_raw2outputs(const float* raw, const float* z_vals, const float* rays_d, int64_t n_rays, int n_samples,
                    int white_bkgd, float* rgb_map, float* disp_map, float* acc_map, float* weights, float* depth_map,
                    void* stream);

/* Hierarchical resampling on the GPU: sample_pdf (utils/run_nerf_raybased_helpers.py:283-330) on
 * bins = mid-points of z_vals and weights[..., 1:-1], then the sorted merge torch.sort(cat(z_vals, z_samples))
 * (utils/create_data.py:503-515).  u[ray*u_stride + j] are the uniforms (u_stride 0 = one shared row, det=True uses
 * linspace(0,1,n_importance)).  Outputs z_samples[N,n_importance], z_merged[N,n_samples+n_importance] ascending. */
int r2l_sample_pdf_merge(const float* z_vals, const float* weights, const float* u, int64_t u_stride, int64_t n_rays,
                         int n_samples, int n_importance, float* z_samples, float* z_merged, void* stream);

/* sample_pdf alone with the reference's own arguments: bins[N,n_bins], weights[N,n_bins-1] -> z_samples[N,n_importance]. */
int r2l_sample_pdf(const float* bins, const float* weights, const float* u, int64_t u_stride, int64_t n_rays, int n_bins,
                   int n_importance, float* z_samples, void* stream);

/* Dense positional encoding x[N,dim] -> out[N, dim*(2*n_freqs+1)].
 * style 0: PositionalEmbedder.__call__ (:198-208) layout; style 1: Embedder.embed (:54-55, get_embedder :58-73) layout. */
int r2l_positional_embed(const float* x, float* out, int64_t n, int dim, int n_freqs, int style, void* stream);

/* One Adam update of a flat fp32 buffer (torch.optim.Adam defaults: no amsgrad / weight decay; main.py:465,:1406).
 * `step` counts from 1.  28 bytes of HBM traffic per parameter in one pass. */
int r2l_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                  double beta2, double eps, int64_t step, void* stream);

/* Ray-shard reader (HOST function, no CUDA call inside): dst_host[i * floats_per_shard ...] = the
 * float32 payload of the `.npy` file paths[i] (np.save of a C-ordered [rows, 9] array: utils/create_data.py:866-869), read
 * with pread on n_threads native threads.  dst_host is normally a pinned batch buffer, so a batch of --N_rand shards is
 * one call and one H2D copy (replaces np.load + torch.Tensor + collate + pin per shard, dataset/load_blender.py:304-318,
 * main.py:795-808).  A file with another dtype / order / size fails the call; r2l_last_error() names it. */
int r2l_read_ray_shards(const char* const* paths, int n_paths, float* dst_host, int64_t floats_per_shard, int n_threads);

/* The same update with the two step-dependent scalars read from DEVICE memory: hyper[0] = lr / (1 - beta1^step),
 * hyper[1] = 1 / sqrt(1 - beta2^step) (r2l_adam_hyper computes them on the host exactly as r2l_adam_step does).  A train
 * step captured in a CUDA graph is replayed with the learning-rate schedule of main.py:1181-1195 by refreshing 8 bytes. */
int r2l_adam_hyper(double lr, double beta1, double beta2, int64_t step, float* hyper_host /* [2], host */);
/* The same scalars computed on the DEVICE from device-resident counters (one tiny launch): counters[0] (the schedule's
 * global_step, main.py:1175-1195) and counters[1] (Adam's step) are incremented, then hyper[0..1] as above and hyper[2] = lr
 * with lr = lrate * decay_rate^((step - warmup_end_iter) / decay_steps), or the linear warm-up
 * (lrate - warmup_start_lr) / warmup_end_iter * step + warmup_start_lr while step < warmup_end_iter (0 = no warm-up).
 * Nothing the step depends on lives in host memory: a host that enqueues iterations ahead of the GPU, or a replayed CUDA
 * graph, applies the scalars of the iteration being executed. */
int r2l_adam_schedule_dev(double lrate, double warmup_start_lr, double warmup_end_iter, double decay_rate, double decay_steps,
                          double beta1, double beta2, int64_t* counters /* [2], device */, float* hyper /* [3], device */,
                          void* stream);
int r2l_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double beta1, double beta2,
                      double eps, const float* hyper, void* stream);

/* ---- Data-parallel training over NVLink / NVSwitch peer memory (one process per GPU; csrc/dp.cu) ----
 * Replaces the reference's nn.DataParallel round trip (replicate all parameters, gather outputs, reduce gradients to GPU 0,
 * torch.optim.Adam there: main.py:37-42, :472-479, :1403-1406) by ONE kernel per iteration: reduce-scatter of the flat
 * gradient through peer loads, Adam on this rank's 1/world slice, all-gather of the new parameters through peer stores.
 *   r2l_dp_create     allocates this rank's block (gradient buffer | parameter buffer | flags) and returns its CUDA IPC handle
 *                     (r2l_dp_handle_bytes() bytes); exchange the handles between the ranks (e.g. torch.distributed
 *                     all_gather_object) and pass all of them, in rank order, to r2l_dp_connect.
 *   r2l_dp_grads / r2l_dp_params   this rank's buffers (n_params floats each): the backward writes its gradient into the
 *                     first, the forward / r2l_pack_weights read the parameters from the second.
 *   r2l_dp_adam_step  the kernel over the whole buffer; hyper as for r2l_adam_step_dev.  Every rank must call it once per
 *                     iteration, after its backward on the same stream; capturable in a CUDA graph.  On return of the kernel
 *                     every rank's parameter buffer holds the same, bit-identical parameters.  A rank that never arrives
 *                     traps the others after 20 s.
 *   r2l_dp_adam_step_range   the same for the float range [lo, hi) (lo a multiple of 4) of the buffer, e.g. one gradient chunk of
 *                     r2l_backward_chunked launched on a communication stream as soon as that chunk is complete, so that only
 *                     the last chunk's exchange is exposed.  Launches that may run concurrently need distinct `slot`s (0..7);
 *                     grid = CTAs (0 = two per SM).  r2l_dp_slice: the part [slice_lo, slice_hi) of [lo, hi) this rank updates
 *                     (only those parts of exp_avg / exp_avg_sq are used on this rank). */
size_t r2l_dp_handle_bytes(void);
int r2l_dp_create(int rank, int world, int64_t n_params, void* handle_out);
int r2l_dp_connect(const void* all_handles);
void* r2l_dp_grads(void);
void* r2l_dp_params(void);
int r2l_dp_slice(int64_t lo, int64_t hi, int64_t* slice_lo, int64_t* slice_hi);
int r2l_dp_adam_step(float* exp_avg, float* exp_avg_sq, double beta1, double beta2, double eps, const float* hyper, void* stream);
int r2l_dp_adam_step_range(float* exp_avg, float* exp_avg_sq, double beta1, double beta2, double eps, const float* hyper,
                           int64_t lo, int64_t hi, int slot, int grid, void* stream);
int r2l_dp_destroy(void);
int r2l_debug_set_dp_grid(int grid, int variant);   /* debug: CTAs of r2l_dp_adam_step (0 = default, two per SM) and timing switches (0 = production) */

/* loss[0] = loss_scale * sum((rgb - target)^2), grad_rgb = grad_scale * (rgb - target), per_ray_err[r] = mean_c (rgb - target)^2;
 * target row r starts at target[r * target_stride] (3 for a [N,3] tensor, 9 with target = rays9 + 6 for shard rows).
 * With loss_scale = lw_rgb / (3 N) and grad_scale = 2 lw_rgb / (3 N_global) this is img2mse(rgb, target) * lw_rgb
 * (nerf_raybased.py:18, main.py:1377), the dL/drgb autograd derives from it, and the per-ray error the hard-example pool
 * sorts by (main.py:1411-1413), in one launch with no host sync.  grad_rgb / per_ray_err may be NULL.  scratch:
 * r2l_loss_scratch_bytes() bytes of device memory, zero before the first call (the kernel leaves it zero).  The sum is taken
 * in a fixed order: bit-reproducible. */
size_t r2l_loss_scratch_bytes(void);
int r2l_mse_loss_grad(const float* rgb, const float* target, int64_t n_rays, int target_stride, float grad_scale, float loss_scale,
                      float* grad_rgb, float* per_ray_err, float* loss, void* scratch, void* stream);

/* ---- Hard-example ray pool on the device (main.py:1325-1347 draw, :1410-1425 update); both calls are one launch, need no
 * host sync and can be captured in a CUDA graph together with the rest of the iteration.
 *   pool_rows  : [capacity, 9] fp32 rows (o | d | rgb);   pool_state : int32[1] = rays currently in the pool (device).
 * r2l_pool_draw   (pool full): dst_rows[j] = pool_rows[slots_out[j]], j < n_out <= size, slots_out = the first n_out values of a
 *   pseudo-random permutation of [0, size) keyed by `seed` and counters[0] (the device-resident iteration counter of
 *   r2l_adam_schedule_dev): replaces np.random.permutation(len(pool))[:n_hard_out] + the host-side cat (:1330-1340).
 * r2l_pool_update: the n_hard_in rays with the largest per_ray_err among the first n_fresh rows of rays9[.,9] (the fresh part
 *   of the batch, :1411-1414; ties: lowest ray index first) are appended at row pool_state[0] (slots_out == NULL; the kernel
 *   advances pool_state[0]) or overwrite rows slots_out[0 .. n_hard_in) (pool full, :1416-1418).  picked (optional):
 *   int32[n_hard_in] out, the selected ray indices (all rays above the k-th error in index order, then the ties). */
int r2l_pool_draw(const float* pool_rows, const int32_t* pool_state, int n_out, uint64_t seed, const int64_t* counters,
                  float* dst_rows, int32_t* slots_out, void* stream);
int r2l_pool_update(const float* rays9, const float* per_ray_err, int64_t n_fresh, int n_hard_in, float* pool_rows, int32_t* pool_state,
                    const int32_t* slots_out, int32_t* picked, void* stream);
/* HOST: slot j of the permutation r2l_pool_draw uses for a pool of `size` rays at iteration counter `step` (-1 on bad arguments). */
int64_t r2l_pool_slot_host(int64_t j, int64_t size, uint64_t seed, int64_t step);

/* Debug: device buffer [grid][8] of int64 cycle counters filled by the next r2l_forward calls (NULL = off):
 * [0] MMA wait on head A chunks, [1] on body A chunks, [2] on weight stages, [3] producer wait on free stages,
 * [4] MMA-thread total. */
int r2l_debug_set_stats(long long* stats);

/* Launch form of the chain kernels (r2l_b200/csrc/chain.cu): 0 = single CTA per 128-ray tile, 1 = CTA pair (tcgen05
 * cta_group::2, M = 256) with one tile per CTA, 2 = CTA pair sharing one tile (cta_group::2, M = 128, 64 rays per CTA:
 * 0.69 of the latency per tile, the form for small batches), -1 = default = chosen per call (form 2 while the batch
 * leaves SM pairs idle, i.e. tiles <= SMs / 2, form 1 otherwise).  Forms 0 and 1 give bit-identical results; form 2 keeps
 * the residual stream out of the tensor core's truncating accumulator (fresh accumulator per GEMM, fp32 residual add in the
 * epilogue) and agrees with them to ~1e-4 relative, being the more accurate one.  Process-wide; buffers sized by the
 * *_bytes queries fit every form. */
int r2l_set_pair_mode(int mode);

/* Kernels launched by this library (or recorded into a CUDA graph under capture) since the last reset; reset != 0 zeroes
 * the counter.  (r2l_teacher_pack_weights counts as one although it launches two.)  Not thread-safe: a measuring aid. */
long long r2l_debug_launch_count(int reset);

/* Form 2 multiplies every accumulator it reads by (1 + eps) to undo, in expectation, the round-toward-zero of the tensor
 * core's fp32 accumulation: eps_body for the K = 256 GEMMs, eps_head for the K = 1024 head.  The library's defaults are
 * calibrated on B200 (tools/gpu_accum_calibrate.py); this call overrides them for such measurements.  Process-wide. */
int r2l_debug_set_accum_debias(float eps_body, float eps_head);

/* Weight gradients of r2l_backward when the ray batch is cut into pieces that run on different SMs (small batches, where
 * the weight-gradient kernel overlaps the backward chain): 0 (default) = every piece adds its result into `grads` with
 * L2 floating-point reductions - the order in which the <= 8 pieces of a layer are summed is not fixed, so results can
 * differ in the last bits from run to run, as torch's atomicAdd-based backward kernels do; 1 = pieces go to scratch and
 * are summed in index order (bit-reproducible, ~0.1 ms slower per 4096-ray backward).  Process-wide. */
int r2l_set_deterministic(int on);

/* Debug / tuning: schedule of the weight-gradient kernel inside r2l_backward (dw.cu).  Units = the 86 body Linears in
 * the order the backward chain releases them, then 4 head column groups.  When the kernel overlaps the chain, units
 * < t1 run whole, < t2 in 2 ray-tile pieces, < t3 in 4, < t4 in 8, the rest in 16; when it runs after the chain every
 * unit is cut into serial_pieces.  Negative values (0 for serial_pieces) = built-in defaults.  Results are independent of
 * the schedule up to fp32 summation order of the pieces. */
int r2l_debug_set_dw_schedule(int t1, int t2, int t3, int serial_pieces, int t4);

/* Debug: device buffer [grid][5][96] of clock64 stamps for the first tile of each CTA of the next chain launches:
 * row 0 MMA thread starts layer l, 1 MMA thread has issued layer l, 2 epilogue sees accumulator l complete,
 * 3 epilogue published the first k-step of layer l's output, 4 epilogue finished layer l; followed by [90][4]
 * %globaltimer stamps of the weight-gradient kernel (start, flag seen, -, end).  NULL = off. */
int r2l_debug_set_trace(long long* trace);

/* Debug: tensor-pipe micro-benchmark; out_cycles[grid] = cycles for `reps` x 48 tcgen05.mma (N256 K16) in the MMA shape of
 * launch form `form` (0: M128 cta_group::1; 1: M256 cta_group::2; 2: M128 cta_group::2; forms 1, 2: grid even, the leader of
 * pair i writes out_cycles[2 i]).  variant: bit 0 = non-zero operands (zeros otherwise); variant >> 1 = issue pattern (0: chunk by
 * chunk, products grouped; form 2 only: 1 = the chain kernel's operand addresses and issue order, 2 = its addresses, products
 * grouped, 3 = as 1 with a tcgen05.fence::after_thread_sync in front of every MMA pair). */
int r2l_debug_mma_rate(int form, int variant, int reps, int grid, long long* out_cycles, void* stream);

/* Debug / test hook: C[128,256] = A[128,256] * W_l^T through one tcgen05 layer step, l = body layer 0..85. */
int r2l_selftest_layer(const float* A, const void* packed, int layer, float* C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* R2L_B200_H_ */
