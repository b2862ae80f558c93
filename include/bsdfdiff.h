/* bsdfdiff.h -- C ABI of the B200-native neural BSDF sampler (libbsdfdiff.so).
 *
 * This is the drop-in boundary for the reference's hot path (fzy28/BSDF_diffusion_sampling @ f2b615c):
 * everything the reference does between "a batch of incident directions arrives as a tensor" and
 * "outgoing directions + pdf leave as tensors" is behind these entry points.  Plain pointers and
 * sizes only; all data pointers are DEVICE pointers unless a parameter says "host"; every launch is
 * enqueued on the given stream and the call returns without synchronising (CUDA-graph capturable,
 * like tiny-cuda-nn's binding: tiny-cuda-nn/bindings/torch/tinycudann/bindings.cpp:95-96).
 * Return value: 0 on success, a negative BSDFDIFF_E* code otherwise; never throws.
 *
 * Reference interfaces replaced (file:line in the reference checkout):
 *   bsdfdiff_sample   <- network_sampling_disk       rendering/utils/mlp_brdf_sampling.py:17-51
 *                        network_sampling_spherical  rendering/utils/mlp_brdf_sampling.py:106-140
 *                        + tensor part of MyBSDF.sample: rendering/brdf_measured_disk.py:66-82,
 *                          rendering/brdf_measured_spherical.py:76-92, rendering/bsdf_myresult.py:65-84
 *   bsdfdiff_pdf      <- network_pdf_disk            rendering/utils/mlp_brdf_sampling.py:69-103
 *                        network_pdf_spherical       rendering/utils/mlp_brdf_sampling.py:144-181
 *                        + tensor part of MyBSDF.pdf: rendering/brdf_measured_disk.py:112-124,
 *                          rendering/brdf_measured_spherical.py:122-137, rendering/bsdf_myresult.py:115-130
 *   bsdfdiff_flow_forward <- rectify_stage.dosampling learning_repo_cleanup/disk_domain_sampling.py:93-110,
 *                        learning_repo_cleanup/spherical_domain_sampling.py:147-166,
 *                        learning_repo_cleanup/bsdf_correct_sampling.py:147-166 and
 *                        network_sampling_disk_tiny  rendering/utils/mlp_brdf_sampling.py:54-68
 *                        (the T-step loop around tinycudann.Network.forward,
 *                        tiny-cuda-nn/bindings/torch/tinycudann/modules.py:176-192)
 *   bsdfdiff_mlp_forward  <- tinycudann.Network(...)(x)  modules.py:176-192 /
 *                        kernel_mlp_fused  tiny-cuda-nn/src/fully_fused_mlp.cu:499-557 (inference)
 *   bsdfdiff_pack_flow / bsdfdiff_pack_flow_tcnn
 *                     <- load_pytorch_model_to_tinycuda learning_repo_cleanup/utils/utils.py:13-23
 *                        and the torch.load + load_state_dict at rendering/brdf_measured_disk.py:43-51
 */
#ifndef BSDFDIFF_H_
#define BSDFDIFF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSDFDIFF_ABI_VERSION 3

/* domains (state parameterisation of the flow) */
#define BSDFDIFF_DISK       0   /* state = projected (x,y) on the unit disk; net input 25 = [x,y,alpha,PE5(wi)] */
#define BSDFDIFF_SPHERICAL  1   /* state = (theta,phi); net input 26 = [theta,sin phi,cos phi,alpha,PE5(wi)]    */

/* epilogues */
#define BSDFDIFF_EPI_RAW        0   /* wi,wo are [n,2] domain coordinates; returns what mlp_brdf_sampling.py returns */
#define BSDFDIFF_EPI_DISK       1   /* brdf_measured_disk.py plugin: wi/wo [n,3] local frame, out wo3 + pdf*cos(theta_o) */
#define BSDFDIFF_EPI_SPHERICAL  2   /* brdf_measured_spherical.py plugin: wi/wo [n,3], sin/cos masks, pdf/sin(theta_o)   */
#define BSDFDIFF_EPI_BSDF       3   /* bsdf_myresult.py plugin: as 2 without the cos mask, abs() in 1/sin                */

/* arithmetic paths */
#define BSDFDIFF_PREC_FP32  0   /* CUDA-core fp32 (parity path: matches the reference to ~1e-5)                 */
#define BSDFDIFF_PREC_TC16  1   /* tcgen05: fp16 operands, fp32 TMEM accumulators (throughput path)             */
#define BSDFDIFF_PREC_TC16_EXP 2 /* as TC16 but SiLU via fp32 exp instead of tanh.approx (accuracy cross-check)  */

/* return codes: 0 ok, > 0 ok with a note, < 0 error */
#define BSDFDIFF_OK            0
#define BSDFDIFF_OK_FP32_REROUTE 1 /* a PREC_TC16* call whose net shape the tensor-core kernel does not cover (more than 6
                                      hidden layers, sample/pdf with a 64-wide net) ran on the PREC_FP32 CUDA-core kernel  */
#define BSDFDIFF_EINVAL       -1   /* bad argument (null pointer, unsupported shape, T < 1, ...) */
#define BSDFDIFF_EUNSUPPORTED -2   /* shape not supported by the requested precision path        */
#define BSDFDIFF_ECUDA        -3   /* a CUDA runtime call failed (see bsdfdiff_last_cuda_error)   */
#define BSDFDIFF_ENOTSM100    -4   /* device is not compute capability 10.x                       */

#define BSDFDIFF_BASE_FLOATS 308   /* base net blob: W1[16x14] row-major, b1[16], Wo[4x16] row-major, bo[4] */

int         bsdfdiff_abi_version(void);
const char* bsdfdiff_error_string(int code);
int         bsdfdiff_last_cuda_error(void);          /* cudaError_t of the last failure on this thread */
int         bsdfdiff_debug_timeout_flag(void);       /* 1 if a tensor-core kernel ever hit its mbarrier watchdog (syncs) */
/* Tuning builds (-DBSDFDIFF_TC_TRACE) only: copies the per-warp %clock64 stamps of CTA 0's steady-state rounds to the host
 * ([24 warps][96 rounds][8 events] uint64; see flow_tc.cu); returns the number of words, 0 in the shipped build. */
int         bsdfdiff_debug_trace(unsigned long long* out_host, int max_words);

/* ---- weight packing (host side; the blob is then copied to the device by the caller) ------------------------
 * Flow net = bias-free MLP, layers given as row-major [rows,cols] fp32 matrices exactly as the checkpoints
 * store them (nn.Linear.weight): W1 [H,in], W2..Wk [H,H], Wout [2,H];  in = 25 (disk) or 26 (spherical),
 * H = 32 or 64.  The blob holds an fp32 image for the PREC_FP32 path and an fp16 shared-memory image
 * (UMMA canonical K-major core-matrix layout) for the PREC_TC16 path.  The launch entry points take the
 * device copy of the blob plus (hidden, n_hidden) so that no call ever reads device memory from the host. */
size_t bsdfdiff_packed_flow_bytes(int in_dim, int hidden, int n_hidden);
int    bsdfdiff_pack_flow(const float* const* layer_ptrs /*host*/, const int* rows, const int* cols, int n_layers,
                          void* packed_out /*host, bsdfdiff_packed_flow_bytes()*/);
/* Same, from the flat fp16/fp32 parameter vector of tinycudann.Network.params in the layout written by
 * load_pytorch_model_to_tinycuda (first layer padded to in+16-in%16 columns, last to out+16-out%16 rows). */
int    bsdfdiff_pack_flow_tcnn(const float* tcnn_params /*host, fp32 copy of .params*/, int in_dim, int out_dim,
                               int hidden, int n_hidden, void* packed_out /*host*/);

/* ---- sample: x0 ~ base(.|wi); T Euler steps of the flow; pdf = p_base(x0) / prod det(I + dD/dx / T) ---------
 * wi:        [n,2] domain coords (EPI_RAW) or [n,3] local-frame directions (plugin epilogues)
 * x0_replay: optional [n,2] externally supplied base sample (noise replay for parity); NULL = draw with
 *            Philox4x32-10, key = seed, counter = (first_index + i, offset)  -> results do not depend on
 *            how the batch is sharded across GPUs.
 * u_noise:   optional [n,3] uniforms in [0,1) used INSTEAD of the Philox draws: the renderer's own sample stream as the
 *            noise source (Mitsuba hands sample2.x, sample2.y, sample1 to MyBSDF.sample and the reference ignores them,
 *            rendering/brdf_measured_disk.py:59-68).  (u0, u1) feed the Box-Muller pair of the base Gaussian; the von
 *            Mises rejection rounds of the spherical base draw from Philox keyed by the bits of (u0, u1, u2), so the
 *            sample is a pure function of the three uniforms.  Mutually exclusive with x0_replay.
 * out_dir:   [n,2] (EPI_RAW: x_T) or [n,3] (plugin: wo);  out_pdf: [n];  out_x0: optional [n,2].
 * T >= 1; T == 0 together with flow_packed == NULL evaluates the base distribution alone
 * (D_base.sample / D_base.log_prob, rendering/utils/model.py:387-398, 299-317). */
int bsdfdiff_sample(int precision, int domain, int epilogue, int T, int64_t n,
                    const float* wi, const void* flow_packed, int hidden, int n_hidden,
                    const float* base_params, const float* x0_replay, const float* u_noise, uint64_t seed, uint64_t offset,
                    int64_t first_index, float* out_dir, float* out_pdf, float* out_x0, float fix_threshold,
                    void* fix_scratch, void* cuda_stream);

/* ---- conditioning-triggered fp32 fix-up of the PREC_TC16 path (fix_threshold, fix_scratch above and below) ----
 * pdf = p_base / prod_t det J_t: where a step determinant is close to zero (or the later steps amplify an early
 * state error, or pdf()'s base density is steep at the flow's end point) fp16-operand arithmetic cannot meet the
 * stated pdf tolerance -- an ideal fp16-operand / fp32-accumulate evaluation misses it too (profiles/r2_emulate_tc16.txt).
 * With fix_threshold > 0 the tensor-core kernel computes, per query, the conditioning weight
 *     w = min(1, min_t|det_t| / 0.2) * min(1, 16 / prod_t max(1, sigma_max(J_t))) [* min(1, 25 / |grad log p_base|), pdf()]
 * and appends every row with w < fix_threshold -- and, for bsdfdiff_sample, the base sample it started from -- to a
 * device list; a second launch on the same stream recomputes exactly those rows with the PREC_FP32 kernel from the same
 * base sample.  No host synchronisation; CUDA-graph capturable (memset + 2 kernels).
 * fix_scratch: device buffer of bsdfdiff_fixup_scratch_bytes(n) ~ 16 + 12 n bytes ([count][row list][base samples of the
 * listed rows]); after the call its first uint32 holds the number of recomputed rows.  fix_threshold = 0 (or
 * PREC_FP32): single launch, scratch may be NULL.
 * Thresholds with which EVERY material the reference ships meets the stated bars (calibrated on all 77 checkpoints,
 * profiles/r2s_material_sweep.txt; the host package's defaults): disk flows 0.15 (sample) / 1/60 (pdf), measured-spherical
 * 0.25 / 0.5, bsdf 0.125 / 0.25.  A material calibrated on its own usually needs none (0). */
size_t bsdfdiff_fixup_scratch_bytes(int64_t n);

/* ---- pdf: reverse flow from wo; pdf = p_base(x_T | wi) * prod det(I - dD/dx / T) ---------------------------- */
int bsdfdiff_pdf(int precision, int domain, int epilogue, int T, int64_t n,
                 const float* wo, const float* wi, const void* flow_packed, int hidden, int n_hidden,
                 const float* base_params, float* out_pdf, float fix_threshold, void* fix_scratch, void* cuda_stream);

/* ---- planar directions: zero-copy hand-off from a renderer that keeps Vector3f as three arrays ------------------------
 * Mitsuba / Dr.Jit store si.wi and bs.wo as three separate device arrays; the reference converts them with .torch()
 * (an interleaving copy) on the way in and rebuilds mi.Vector3f from three strided slices on the way out
 * (rendering/brdf_measured_disk.py:66,80, rendering/bsdf_myresult.py:65,72).  These entry points read / write the three
 * component arrays in place (DLPack-exported Dr.Jit arrays on the Python side): wi_xyz / wo_xyz / out_xyz are HOST arrays
 * of three DEVICE pointers to [n] floats.  Plugin epilogues only (EPI_DISK / EPI_SPHERICAL / EPI_BSDF); everything else as
 * bsdfdiff_sample / bsdfdiff_pdf, and the results are bit-identical to the interleaved calls. */
int bsdfdiff_sample_planar(int precision, int domain, int epilogue, int T, int64_t n, const float* const wi_xyz[3],
                           const void* flow_packed, int hidden, int n_hidden, const float* base_params,
                           const float* x0_replay, const float* u_noise, uint64_t seed, uint64_t offset,
                           int64_t first_index, float* const out_xyz[3], float* out_pdf, float* out_x0,
                           float fix_threshold, void* fix_scratch, void* cuda_stream);
int bsdfdiff_pdf_planar(int precision, int domain, int epilogue, int T, int64_t n, const float* const wo_xyz[3],
                        const float* const wi_xyz[3], const void* flow_packed, int hidden, int n_hidden,
                        const float* base_params, float* out_pdf, float fix_threshold, void* fix_scratch,
                        void* cuda_stream);

/* ---- log p_base(x | wi): D_base.log_prob (rendering/utils/model.py:393-398 disk, :308-317 spherical) ---------
 * x, wi: [n,2] domain coordinates.  Returned as a LOG density (no exp/log round trip: finite where p underflows). */
int bsdfdiff_base_log_prob(int domain, int64_t n, const float* x, const float* wi, const float* base_params,
                           float* out_logp, void* cuda_stream);

/* ---- forward-only flow (reflow "dosampling"): x_T = x0 + sum_t D(x_t, t/T | wi)/T, no pdf -------------------
 * wi: [n_wi,2] domain coords; query i uses wi[i / wi_repeat] (repeat_interleave without materialising it).
 * x0: [n,2] start state, or NULL to draw it from the base net (base_params required) with Philox;
 * out_x0 (optional) receives the start state.
 * PREC_TC16 covers hidden = 32 (NN_cond_pos_simpler, disk reflow, T = 256) and hidden = 64 with up to 6 hidden layers
 * (NN_cond_pos_spherical_complicate, learning_repo_cleanup/utils/model.py:449-477, spherical / bsdf reflow, T = 128):
 * replaces the tcnn FullyFusedMLP loop of disk_domain_sampling.py:93-110 / spherical_domain_sampling.py:147-166. */
int bsdfdiff_flow_forward(int precision, int domain, int T, int64_t n, const float* wi, int64_t wi_repeat,
                          const void* flow_packed, int hidden, int n_hidden, const float* base_params,
                          const float* x0, uint64_t seed, uint64_t offset, int64_t first_index,
                          float* out_x, float* out_x0, void* cuda_stream);

/* ---- one MLP forward (tinycudann.Network.forward, inference): out[n,2] = MLP(in[n,in_dim]) ------------------ */
int bsdfdiff_mlp_forward(int precision, int64_t n, const float* in, int in_dim, const void* flow_packed,
                         int hidden, int n_hidden, float* out /*[n,2]*/, void* cuda_stream);

/* ---- one wavefront, several materials: a material id per row, ONE launch -------------------
 * Replaces Mitsuba's per-instance dispatch for scenes with many `mybsdf` BSDFs (twelve in
 * rendering/matpreview/disney_bsdf_array0_envmap.xml:35-335: each instance's sample()/pdf() is called with the lanes
 * that hit it) for a renderer that keeps one wavefront with a material id per lane.
 * bsdfdiff_multi_plan buckets the rows by material ON THE DEVICE (histogram, tile table, scatter: no host read) into
 * virtual tiles of <= 128 rows of one material; bsdfdiff_sample_multi / bsdfdiff_pdf_multi walk that tile table in one
 * persistent launch and switch the weight set in shared memory per tile (PREC_TC16: one weight image per in-flight
 * tile, re-staged with one cp.async.bulk when a tile's material differs).  Rows are read and written at their
 * WAVEFRONT position and the Philox counter is first_index + wavefront row, so every row equals the single-material
 * call on that row bit for bit, whatever the bucketing.  Rows whose id is outside [0, n_materials) are inactive lanes:
 * no flow runs for them and their outputs are zeroed.  A plan can be reused by any number of sample / pdf calls on the
 * same material_id column (sample, then pdf for MIS); calls sharing one plan must be stream-ordered.
 * material_id:   [n] int32 (device).   scratch / plan: device buffer of bsdfdiff_multi_scratch_bytes(n, n_materials).
 * flows_packed:  DEVICE array [n_materials] of device pointers to packed flow blobs, all of one (domain, hidden,
 *                n_hidden) shape;  base_params: DEVICE array [n_materials] of device pointers to 308-float base blobs.
 * fix_threshold: as bsdfdiff_sample; NEGATIVE = every material's OWN thresholds, read from its packed blob's header
 *                (bytes 40..51: uint32 magic 0x54584946 "FIXT", float sample threshold, float pdf threshold -- written by the
 *                host after per-material calibration), |fix_threshold| for a blob that carries none
 *                (the plan buffer holds the per-material fix-up lists and, when neither x0_replay
 *                nor out_x0 is given, the base samples the fix-up pass replays).  n_materials <= 255, n < 2^31. */
size_t bsdfdiff_multi_scratch_bytes(int64_t n, int n_materials);
int bsdfdiff_multi_plan(int64_t n, const int32_t* material_id, int n_materials, void* scratch, void* cuda_stream);
int bsdfdiff_sample_multi(int precision, int domain, int epilogue, int T, int64_t n, const float* wi,
                          const void* plan, int n_materials, const void* const* flows_packed,
                          const float* const* base_params, int hidden, int n_hidden, const float* x0_replay,
                          const float* u_noise, uint64_t seed, uint64_t offset, int64_t first_index, float* out_dir,
                          float* out_pdf, float* out_x0, float fix_threshold, void* cuda_stream);
int bsdfdiff_pdf_multi(int precision, int domain, int epilogue, int T, int64_t n, const float* wo, const float* wi,
                       const void* plan, int n_materials, const void* const* flows_packed,
                       const float* const* base_params, int hidden, int n_hidden, float* out_pdf,
                       float fix_threshold, void* cuda_stream);

/* ---- training step of the flow nets: fused forward + backward + Adam ----------------------------------
 * Replaces one iteration of the reference's diffusion_stage / rectify_stage loops
 * (learning_repo_cleanup/disk_domain_sampling.py:49-58,123-131; spherical_domain_sampling.py:55-76; the
 * kernel_mlp_fused_backward analogue, tiny-cuda-nn/src/fully_fused_mlp.cu:150-259):
 *     alpha = linspace(0,1,n) (or the given column);  x_alpha = (1 - alpha) x0 + alpha x1  (spherical: x1.phi first moved
 *     to within pi of x0.phi, net input (theta, sin phi, cos phi));  pred = D(x_alpha, alpha, wi);
 *     loss = mean((pred - (x1 - x0))^2);  grad = dloss/dW;  torch.optim.Adam update (amsgrad off, no weight decay).
 * weights / grad / adam_m / adam_v: device fp32 vectors of bsdfdiff_flow_param_count() elements in the checkpoint layout
 * (linear1.weight [H,in] row-major, linear2.., output.weight [2,H], concatenated): these ARE the master weights.
 * grad must be zero on entry; with apply_update != 0 it is zero again on exit and the weights / moments are updated
 * (step = 1-based step count for the bias correction), with apply_update == 0 it holds dloss/dW (weights untouched).
 * loss_out: device float (the step's loss).  sync_scratch: 4 zero bytes of device memory (left zero).  One launch, no host
 * synchronisation, CUDA-graph capturable.  fp32 arithmetic; hidden 32 (3 or 4 hidden layers) or 64 (up to 6). */
size_t bsdfdiff_flow_param_count(int in_dim, int hidden, int n_hidden);
int bsdfdiff_flow_matching_step(int domain, int hidden, int n_hidden, int64_t n, const float* x0, const float* x1,
                                const float* wi, const float* alpha, float* weights, float* grad, float* adam_m,
                                float* adam_v, float lr, float beta1, float beta2, float eps, int64_t step,
                                int apply_update, float* loss_out, void* sync_scratch, void* cuda_stream);

/* Pretrain stage (disk_domain_sampling.py:14-33, spherical_domain_sampling.py:16-35): loss = -mean(D_base.log_prob(x, wi))
 * over the batch, gradient w.r.t. the 308 base-net parameters (blob order W1 [16,14], b1 [16], Wo [4,16], bo [4]), Adam --
 * one launch, same buffer contract as bsdfdiff_flow_matching_step (base_params are the fp32 master weights). */
int bsdfdiff_base_nll_step(int domain, int64_t n, const float* x /*[n,2] omega_o*/, const float* wi /*[n,2] omega_i*/,
                           float* base_params, float* grad, float* adam_m, float* adam_v, float lr, float beta1,
                           float beta2, float eps, int64_t step, int apply_update, float* loss_out, void* sync_scratch,
                           void* cuda_stream);

/* ---- measured-BSDF ground truth: Mitsuba 3 `measured` eval of an RGL tensor file, and the plugins' weight + firefly clamp ----
 * Replaces  self.bsdf.eval(ctx, si, bs.wo)  (mi.load_dict({"type": "measured", ...}), rendering/brdf_measured_disk.py:36-42,92
 * and rendering/brdf_measured_spherical.py:45-51,100) and the Dr.Jit<->torch round trips of the firefly clamp
 * (brdf_measured_disk.py:93-101, brdf_measured_spherical.py:101-109).  Mitsuba is third-party and absent here: the model is
 * implemented from its published description (Dupuy & Jakob 2018; Mitsuba 3 measured.cpp / distr_2d.h), see
 * csrc/measured.cu and oracle/measured_oracle.py.
 * pack (host): raw tensor-file fields, row-major as stored -- phi_i [n_phi], theta_i [n_theta], ndf [ndf_h, ndf_w],
 * sigma [sig_h, sig_w], vndf [n_phi, n_theta, v_h, v_w], rgb [n_phi, n_theta, 3, r_h, r_w] -> blob (then copied to the device).
 * eval:   out_rgb[n,3] = measured.eval(wi, wo) (= f cos(theta_o); 0 unless cos_i > 0 and cos_o > 0); wi, wo [n,3] local frame.
 * weight: value = eval / bs_pdf * albedo [spherical epilogue: zeroed unless cos_i > 0 and bs_pdf > 0];
 *         out_pdf = luminance(value) < clamp ? bs_pdf : 0;  out_weight = cos_i > 0 and out_pdf > 0 and cos_o > 0 ? value : 0.
 *         epilogue = BSDFDIFF_EPI_DISK | BSDFDIFF_EPI_SPHERICAL (which plugin's tail); bs_pdf = the pdf bsdfdiff_sample returned. */
size_t bsdfdiff_measured_blob_bytes(int n_phi, int n_theta, int ndf_w, int ndf_h, int sig_w, int sig_h,
                                    int v_w, int v_h, int r_w, int r_h);
int    bsdfdiff_measured_pack(const float* phi_i /*host*/, int n_phi, const float* theta_i, int n_theta, const float* ndf,
                              int ndf_w, int ndf_h, const float* sigma, int sig_w, int sig_h, const float* vndf,
                              int v_w, int v_h, const float* rgb, int r_w, int r_h, int jacobian, void* blob_out /*host*/);
int    bsdfdiff_measured_eval(const void* blob, int64_t n, const float* wi, const float* wo, float* out_rgb, void* cuda_stream);
int    bsdfdiff_measured_weight(const void* blob, int epilogue, int64_t n, const float* wi, const float* wo,
                                const float* bs_pdf, float albedo_r, float albedo_g, float albedo_b, float clamp,
                                float* out_weight, float* out_pdf, void* cuda_stream);

/* Static facts about the tensor-core kernel for a given shape (for bench.py / DESIGN.md bookkeeping). */
int bsdfdiff_device_info(int* sm_count, int* cc_major, int* cc_minor);

#ifdef __cplusplus
}
#endif
#endif /* BSDFDIFF_H_ */
