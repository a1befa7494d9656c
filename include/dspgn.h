/*
 * dspgn.h -- C ABI of libdspgn.so: DSP-SLAM's per-object shape-prior Gauss-Newton reconstruction
 * as hand-written CUDA for NVIDIA B200 (sm_100a).
 *
 * Drop-in boundary.  The reference has no native FFI for this path: its C++ LocalMapping thread
 * calls Python through pybind11 (src/LocalMapping.cc:38-40, src/LocalMapping_util.cc:109-110,
 * 179-181, 390-392), and Python issues PyTorch ops.  The entry points below are what a native
 * binding for exactly those calls binds to; dsp_slam_b200/optimizer.py (ctypes) is that binding,
 * and INTEGRATION.md shows the one-file replacement of reconstruct/optimizer.py.
 *
 *   reference call                                            entry point here
 *   --------------------------------------------------------  -------------------------------
 *   reconstruct.utils.get_decoder (deep_sdf/workspace.py:202)  dspgn_decoder_create
 *   Optimizer.__init__           (reconstruct/optimizer.py:27)  dspgn_solver_create
 *   Optimizer.reconstruct_object (reconstruct/optimizer.py:88)  dspgn_reconstruct_batch
 *   Optimizer.estimate_pose_cam_obj (optimizer.py:45)           dspgn_estimate_pose_batch
 *   loss_utils.decode_sdf        (reconstruct/loss_utils.py:51) dspgn_decode_sdf
 *   loss.compute_sdf_loss / compute_render_loss (loss.py:22,46) dspgn_debug_system (test hook)
 *
 * Conventions: every function returns 0 on success or a negative DSPGN_E_* code; per-object soft
 * failures (the reference's is_good=False exits, optimizer.py:130-150) are reported in
 * DspgnObjectOut.status and never through the return code; nothing throws across this ABI.
 * All host buffers are caller-owned, plain float32/int32 with explicit element strides (so the
 * column-major arrays pybind11's Eigen casters produce need no host-side transpose); device
 * buffers are owned by the handles.  One solver = one GPU = one host thread at a time.
 */
#ifndef DSPGN_H_
#define DSPGN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSPGN_MAX_CODE 64
#define DSPGN_MAX_LINEAR 12
#define DSPGN_MAX_CLASSES 4

/* return codes */
#define DSPGN_OK 0
#define DSPGN_E_ARG (-1)      /* bad argument / unsupported decoder shape */
#define DSPGN_E_CUDA (-2)     /* CUDA runtime error; see dspgn_last_error() */
#define DSPGN_E_NOGPU (-3)    /* no usable sm_100 device */
#define DSPGN_E_ALLOC (-4)
#define DSPGN_E_PEER (-5)     /* multi-GPU exchange: a peer never published its results (timeout) */

/* DspgnObjectOut.status (per-object soft failure = the reference's is_good=False exits) */
#define DSPGN_ST_OK 0
#define DSPGN_ST_SDF_NAN 1     /* optimizer.py:135-136 */
#define DSPGN_ST_RENDER_FEW 2  /* loss.py:72-73 (fewer than 10 samples in the unit sphere) */
#define DSPGN_ST_RENDER_NAN 3  /* optimizer.py:149-150 (no band rows -> NaN loss) */
#define DSPGN_ST_SOLVE 4       /* normal matrix not positive definite / non-finite step */
#define DSPGN_ST_BAD_INPUT 5   /* unusable detection (no surface points, too many rays, ...): never evaluated */

/* kernel schedules of a run (results are bit-identical; the per-iteration schedule exists for debugging / profiling) */
#define DSPGN_SCHED_AUTO 0
#define DSPGN_SCHED_LAUNCHES 1   /* one launch per residual term and solve per GN iteration */
#define DSPGN_SCHED_PERSISTENT 2 /* one persistent object-pipelined kernel for all iterations (tensor-core engine only) */

/* decoder engines */
#define DSPGN_ENGINE_AUTO 0
#define DSPGN_ENGINE_SIMT 1    /* fp32 FFMA kernels: on-device ground truth */
#define DSPGN_ENGINE_TC 2      /* tcgen05 tensor-core kernels, 3-pass split-fp16 (fp32-class accuracy) */

typedef struct DspgnDecoder DspgnDecoder;
typedef struct DspgnSolver DspgnSolver;

/* A DeepSDF decoder (deep_sdf/deep_sdf_decoder.py:29-63) with weight-norm already folded:
 * layer k is  y = W[k] x + b[k],  W[k] row-major (out_dim[k], in_dim[k]);
 * ReLU after every layer but the last, tanh after the last (deep_sdf_decoder.py:103-108);
 * at layer `latent_in_layer` the (latent_size+3)-wide input is concatenated after the
 * activations (deep_sdf_decoder.py:87-88); -1 = none. */
typedef struct {
  int32_t latent_size;
  int32_t num_linear;
  int32_t in_dim[DSPGN_MAX_LINEAR];
  int32_t out_dim[DSPGN_MAX_LINEAR];
  int32_t latent_in_layer;
  /* Optional variants of deep_sdf_decoder.py (all zero = the plain decoder above):
   *   cat_kind[k]   what is concatenated AFTER the activations at the input of layer k: 0 nothing, 1 the decoder
   *                 input (latent_in, :87-88; several layers allowed), 2 xyz only (xyz_in_all, :89-90).
   *                 latent_in_layer >= 0 is shorthand for cat_kind[latent_in_layer] = 1.
   *   layer_norm[k] 1: LayerNorm (eps 1e-5) between layer k and its ReLU (:58-63,96-102); gamma/beta through
   *                 dspgn_decoder_create_ex
   *   use_tanh      1: an extra tanh on the last layer before the final one (:93-94,107-108)
   * Decoders that use any of them run on the fp32 SIMT engine (the tcgen05 engine covers the plain shape). */
  int32_t cat_kind[DSPGN_MAX_LINEAR];
  int32_t layer_norm[DSPGN_MAX_LINEAR];
  int32_t use_tanh;
  int32_t reserved_;
} DspgnDecoderSpec;

/* The `optimizer` block of configs/config_*.json as read by reconstruct/optimizer.py:27-43. */
typedef struct {
  float k1, k2, k3, k4;        /* joint_optim.k1..k4 */
  float b1, b2;                /* Huber thresholds: render, sdf */
  float lr;                    /* joint_optim.learning_rate */
  float s_damp;                /* joint_optim.scale_damping */
  int32_t num_iterations;      /* joint_optim.num_iterations */
  int32_t code_len;            /* 32 or 64 (<= latent_size of the decoders) */
  int32_t num_depth_samples;   /* D, <= 64 */
  float cut_off;               /* cut_off_threshold */
  int32_t pose_only_iterations;/* pose_only_optim.num_iterations */
  int32_t sdf_only;            /* 1: skip the render term (BASELINE config 2 "surface-SDF loss") */
  int32_t engine;              /* DSPGN_ENGINE_* */
  int32_t schedule;            /* DSPGN_SCHED_*: 0 = automatic (persistent kernel on the tensor-core engine) */
} DspgnConfig;

/* One detection, host side.  Strides are in elements (floats). */
typedef struct {
  const float* t_cam_obj; int32_t t_rs, t_cs;           /* (4,4) object->camera, Sim(3) */
  const float* pts;  int32_t n_pts;  int32_t pts_rs, pts_cs;    /* (n_pts,3) camera frame */
  const float* rays; int32_t n_rays; int32_t rays_rs, rays_cs;  /* (n_rays,3), foreground first */
  const float* depth; int32_t n_depth;                  /* (n_depth,) foreground depths */
  const float* code;                                    /* (code_len,) initial code or NULL = zeros */
  float scale;                                          /* estimate_pose only: object scale */
  int32_t class_id;                                     /* index into the solver's decoder list */
  /* Optional input construction ON THE DEVICE; all NULL = the arrays above are used as given.
   *   pixels + inv_k   rays[i] = inv_k [u_i, v_i, 1]   (loss_utils.get_rays, reconstruct/loss_utils.py:23-37;
   *                    src/LocalMapping_util.cc:378-386).  (n_rays,2) pixel coordinates replace `rays` (ignored).
   *   t_cam_world      SE(3) world->camera, 4x4 row-major: `pts` are WORLD map points, x_c = R x_w + t
   *                    (LocalMapping_util.cc:344-352), and `t_cam_obj` is the object's WORLD pose T_wo:
   *                    T_co = T_cw T_wo (LocalMapping_util.cc:390). */
  const float* pixels; int32_t pix_rs, pix_cs;
  const float* inv_k;                                   /* 3x3 row-major */
  const float* t_cam_world;                             /* 4x4 row-major */
} DspgnObjectIn;

typedef struct {
  float t_cam_obj[16];            /* row-major (4,4); undefined when status != 0 */
  float code[DSPGN_MAX_CODE];
  float loss;                     /* k1*render + k2*sdf of the last evaluated iteration */
  int32_t status;                 /* DSPGN_ST_* */
  int32_t n_valid;                /* V: ray samples inside the unit sphere, last iteration */
  int32_t n_band;                 /* m: band rows kept, last iteration */
  int32_t iters_done;
  int32_t pad_[3];
} DspgnObjectOut;                 /* 88 floats */

/* device-side result record (same layout), for callers that keep results on the GPU */
#define DSPGN_RESULT_FLOATS 88

const char* dspgn_last_error(void);
int dspgn_version(void);

int dspgn_decoder_create(const DspgnDecoderSpec* spec, const float* const* W, const float* const* b,
                         int device, DspgnDecoder** out);
/* the same with LayerNorm parameters: ln_gamma[k] / ln_beta[k] (out_dim[k] floats) for layers with layer_norm[k] = 1
 * (entries of other layers are ignored; both arrays may be NULL when no layer is normalised) */
int dspgn_decoder_create_ex(const DspgnDecoderSpec* spec, const float* const* W, const float* const* b,
                            const float* const* ln_gamma, const float* const* ln_beta, int device, DspgnDecoder** out);
void dspgn_decoder_destroy(DspgnDecoder* dec);

int dspgn_solver_create(const DspgnConfig* cfg, DspgnDecoder* const* classes, int n_classes,
                        int device, DspgnSolver** out);
void dspgn_solver_destroy(DspgnSolver* s);
/* stream = a cudaStream_t (NULL = legacy default stream). Work is enqueued on it. */
int dspgn_solver_set_stream(DspgnSolver* s, void* cuda_stream);
int dspgn_solver_engine(const DspgnSolver* s);   /* resolved DSPGN_ENGINE_* */
int dspgn_solver_sync(DspgnSolver* s);           /* wait for everything enqueued on the solver's stream */

/* Whole call, host buffers in, host buffers out (upload + all GN iterations + download + sync). */
int dspgn_reconstruct_batch(DspgnSolver* s, int n_obj, const DspgnObjectIn* in, DspgnObjectOut* out);
int dspgn_estimate_pose_batch(DspgnSolver* s, int n_obj, const DspgnObjectIn* in, DspgnObjectOut* out);

/* The same split into its three phases, for callers that keep a batch resident in HBM:
 *   upload:  pack + H2D of the batch into the solver's workspace (async on the stream)
 *   run:     reset state from the uploaded initial poses/codes, run all GN iterations (async);
 *            mode 0 = joint (reconstruct_object), 1 = pose-only (estimate_pose_cam_obj)
 *   results: D2H + stream sync;  results_device: pointer to n_obj*DSPGN_RESULT_FLOATS floats */
int dspgn_upload_batch(DspgnSolver* s, int n_obj, const DspgnObjectIn* in);
int dspgn_run_batch(DspgnSolver* s, int mode);
int dspgn_results(DspgnSolver* s, DspgnObjectOut* out);
const float* dspgn_results_device(DspgnSolver* s);

/* Forward-only decode (loss_utils.decode_sdf): x (n,3) host, strides in elements -> sdf (n,) host. */
int dspgn_decode_sdf(DspgnSolver* s, int class_id, const float* code, const float* x, int n,
                     int x_rs, int x_cs, float* sdf_out);

/* Counters of the last run (for roofline arithmetic): decoder rows evaluated fwd+bwd (SDF rows + band rows) and
 * fwd-only, and the number of kernel launches issued.  Persistent schedule: fwd-only rows = the sum of V over objects
 * and iterations, i.e. the ray samples inside the unit sphere, which is what the reference decodes (loss.py:68,77-78);
 * one-launch-per-term schedule: every n_rays x D sample it evaluates. */
typedef struct {
  int64_t rows_fwd_bwd;
  int64_t rows_fwd_only;
  int64_t kernel_launches;
  float decoder_ms;      /* device time of the decoder kernels of the last run (CUDA events), if timed */
  float total_ms;
  float solve_ms;        /* device time of the per-object solve kernels of the last run, if timed */
  float pad_;
} DspgnCounters;
int dspgn_counters(DspgnSolver* s, DspgnCounters* out);
int dspgn_enable_timing(DspgnSolver* s, int on);

/* ---- Multi-GPU result exchange (the reference reconstructs objects one by one on one GPU,
 * src/LocalMapping_util.cc:165-203 -- objects are independent, so a batch is sharded object-per-GPU).
 * One process per GPU.  There is no collective kernel: rank 0 owns a "gather buffer" in its HBM, exports it
 * with CUDA IPC, every other rank maps it over NVLink/NVSwitch, and the solve step that finishes an object
 * stores the object's 352-byte result record STRAIGHT INTO rank 0's buffer (peer st.global issued from the
 * same kernel that runs the tcgen05 tiles), at the object's slot = its index in the original batch.  A
 * per-rank sequence flag (release, system scope) publishes a finished step; rank 0 waits for all flags with
 * a one-warp kernel on its own stream.  Two slot sets alternate by step parity and rank 0 acknowledges
 * consumed steps, so ranks may run at most one step ahead of rank 0.  `seq` = 1, 2, 3, ... (caller-owned,
 * identical on all ranks).
 *   rank 0:   gather_create -> (handle to the peers by any host channel) ;  ranks 1..: gather_open
 *   per step, every rank:  upload_batch ; gather_bind(slots) ; run_batch_gather(mode, seq)
 *   rank 0:   gather_results(seq, n, out)  (D2H + sync)   or   gather_device(seq) to keep them in HBM */
#define DSPGN_IPC_HANDLE_BYTES 64
typedef struct { unsigned char bytes[DSPGN_IPC_HANDLE_BYTES]; } DspgnIpcHandle;
int dspgn_gather_create(DspgnSolver* s, int n_slots, int world, DspgnIpcHandle* handle_out);
int dspgn_gather_open(DspgnSolver* s, const DspgnIpcHandle* handle, int n_slots, int world, int rank);
/* slots[i] = slot of resident object i (n == resident objects); n == 0: this rank owns no object this step */
int dspgn_gather_bind(DspgnSolver* s, const int32_t* slots, int n);
int dspgn_run_batch_gather(DspgnSolver* s, int mode, int seq);
int dspgn_gather_results(DspgnSolver* s, int seq, int n, DspgnObjectOut* out);
const float* dspgn_gather_device(DspgnSolver* s, int seq);
/* device time the root's wait kernel spent spinning in the last run_batch_gather (ns, after a sync); -1 if n/a */
long long dspgn_gather_wait_ns(DspgnSolver* s);
void dspgn_gather_close(DspgnSolver* s);

/* Test hook: Lie-group exponentials exactly as the solve step applies them (loss_utils.py:129-233):
 * x (n,7) -> out (n,12) row-major 3x4 [sR | J v]; sim3 = 0 ignores x[6] (exp_se3). */
int dspgn_debug_exp(int device, int sim3, const float* x, int n, float* out);

/* Test hook: evaluate one GN iteration at the uploaded initial state of object `obj` WITHOUT
 * updating it and return the assembled system: H (P*P row-major), b (P), dx (P), P = 7+code_len
 * (mode 0) or 6 (mode 1); J_rows/res_rows (may be NULL): the SDF-term Jacobian rows (n_pts,P) and
 * residuals (n_pts) as loss.compute_sdf_loss returns them. */
int dspgn_debug_system(DspgnSolver* s, int obj, int mode, float* H, float* b, float* dx,
                       float* J_rows, float* res_rows, float* losses /* [sdf, render, V, m] */);
/* The same after advancing the uploaded batch `iter` GN iterations from its initial state (iter = 0: identical
 * to dspgn_debug_system): the system the (iter+1)-th iteration solves, for iteration-by-iteration parity
 * against the reference's captured H/b/dx of every iteration (tests/golden/recon_*.npz H_iters[iter]). */
int dspgn_debug_system_iter(DspgnSolver* s, int obj, int mode, int iter, float* H, float* b, float* dx,
                            float* J_rows, float* res_rows, float* losses);

/* Debug: clock64 timeline of CTA 0 of the tensor-core decoder kernel, [4 tiles][18 steps][8 slots]
 * (only when the solver was created with env DSPGN_CLK set). */
int dspgn_debug_clocks(DspgnSolver* s, long long* out, int n);

/* Test hook for the device-side input construction: the resident batch's object `obj` as the kernels see it --
 * t_cam_obj (16, row-major), pts (n_pts*3, xyz interleaved), rays (n_rays*3).  Any pointer may be NULL. */
int dspgn_debug_inputs(DspgnSolver* s, int obj, float* t_cam_obj, float* pts, float* rays);

/* Debug: event log of the last persistent-kernel run (tile begin/end per kind, scan, solve, queue pops), enabled by
 * env DSPGN_CLK at solver creation; returns the number of (timestamp, descriptor) pairs written, or a negative code. */
int dspgn_debug_events(DspgnSolver* s, long long* out, int max_events);

/* Test hook for the tcgen05 operand paths: D[128][n_mma] = A[128][16*k_steps] * B[n_mma][16*k_steps]^T
 * (A through the TMEM split-fp16 path, B through the pre-swizzled shared-memory images). Host buffers. */
int dspgn_tc_selftest(int device, int n_mma, int k_steps, const float* A, const float* B, float* D);

#ifdef __cplusplus
}
#endif
#endif /* DSPGN_H_ */
