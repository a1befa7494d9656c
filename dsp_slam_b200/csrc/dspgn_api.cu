// libdspgn.so host side: handles, weight packing, batch upload, launch sequencing.  C ABI in
// include/dspgn.h.  No torch, no exceptions across the boundary.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <new>
#include <algorithm>

#include "dspgn_common.cuh"
#include "dspgn_simt.cuh"
#include "dspgn_solve.cuh"
#include "dspgn_tc.cuh"

using namespace dspgn;

namespace {

constexpr int kEvCap = 1 << 18;     // events of the persistent kernel's debug log (env DSPGN_CLK)

thread_local std::string g_err;

int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(DSPGN_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return -1; }
    cap = want;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct HostBuf {   // pinned staging
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    if (cudaMallocHost(&p, want) != cudaSuccess) { cudaGetLastError(); return -1; }
    cap = want;
    return 0;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace

struct DspgnDecoder {
  int device = 0;
  bool has_ln = false;
  DspgnDecoderSpec spec{};
  DecoderDev dev{};
  std::vector<void*> allocs;
  TcDecoderHost tc;
};

struct DspgnSolver {
  int device = 0;
  int num_sms = 0;
  int engine = DSPGN_ENGINE_SIMT;
  DspgnConfig cfg{};
  cudaStream_t stream = nullptr;
  std::vector<DspgnDecoder*> classes;
  DevBuf d_decs;
  // resident batch
  int n_obj = 0;
  int tot_pts = 0, tot_rays = 0, tot_fg = 0;
  long long tot_smp = 0;
  int max_rays = 0;
  std::vector<ObjMeta> h_meta;
  HostBuf h_stage;
  DevBuf d_stage;        // one contiguous upload: meta | T_init | code_init | pts | rays | depth
  ObjMeta* d_meta = nullptr; float* d_Tinit = nullptr; float* d_code = nullptr;
  float* d_pts = nullptr; float* d_rays = nullptr; float* d_depth = nullptr;
  DevBuf d_state, d_part_s, d_part_r, d_tbase, d_V, d_m, d_results, d_active;
  DevBuf d_sdf, d_bx, d_bs, d_br;
  DevBuf d_dbg;
  DevBuf d_q_flag, d_q_ctr, d_tiles_left, d_obj_iter;   // persistent-kernel work queue
  int* d_tbase_static = nullptr;   // [n_obj] first 128-row SDF tile of each object (inside the staging block)
  int* d_tbase_r_static = nullptr; // [n_obj] first band-tile partial slot of each object (capacity: its ray-sample tiles + 1)
  int* d_q0_render = nullptr;      // [n_obj] first iteration-0 queue slot of each object in a run with the render term
  int total_tiles128 = 0;          // SDF tiles of the batch
  long long total_ray_tiles128 = 0;// ray-sample tiles of the batch
  int max_tiles128 = 0;            // largest tile count of one term of one object (queue items hold 19 bits)
  bool mega_enabled = true;
  bool compact_rays = true;        // persistent kernel, render term: forward-only tiles over the valid-sample hulls only (env DSPGN_COMPACT_RAYS=0: all n_rays x D samples)
  bool vpre_exact = false;         // env DSPGN_VPRE_EXACT=1: exhaustive valid-range pre-pass (A/B switch)
  DevBuf d_clk, d_ev, d_seg, d_ln, d_vpre;
  bool clk_on = false;
  HostBuf h_results;
  // counters
  DspgnCounters ctr{};
  bool timing = false;
  std::vector<cudaEvent_t> ev;
  size_t ev_used = 0;
  std::vector<cudaEvent_t> ev_solve;
  size_t evs_used = 0;
  cudaEvent_t ev_run0 = nullptr, ev_run1 = nullptr;
  cudaStream_t stream2 = nullptr;    // fork: the ray-sample forward pass runs beside the SDF-row pass
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_upload = nullptr;   // the pinned staging block may be rewritten only after its last H2D copy finished
  bool upload_pending = false;
  bool mega_ran = false;             // the last run used the persistent kernel: check its abort flag with the results
  bool band_rows_pending = false;    // the persistent kernel's band-row total has not been added to ctr yet
  int n_bad = 0;                     // resident objects rejected at upload (status BAD_INPUT, never evaluated)
  // multi-GPU result exchange (rank 0 owns the buffer, the others map it through CUDA IPC)
  struct Gather {
    bool active = false, owner = false;
    unsigned char* base = nullptr;   // [2][n_slots][88] floats | flags[world] | ack
    int n_slots = 0, world = 0, rank = 0;
    size_t off_flags = 0, off_ack = 0;
    DevBuf d_slot_of, d_local;       // d_local: int err | long long wait_ns
    HostBuf h_out;
    int bound_n = -1;
  } gather;
  GatherDev gdev{};                  // exchange arguments of the run being enqueued (slots == nullptr: off)
};

namespace {

// what is concatenated at the input of layer k (latent_in_layer is shorthand for one kind-1 layer)
int cat_kind_of(const DspgnDecoderSpec& s, int k) {
  if (s.cat_kind[k] != 0) return s.cat_kind[k];
  return (k == s.latent_in_layer) ? 1 : 0;
}

int check_spec(const DspgnDecoderSpec& s) {
  if (s.num_linear < 3 || s.num_linear > 9) return fail(DSPGN_E_ARG, "num_linear must be in [3,9]");
  if (s.latent_size < 1 || s.latent_size > DSPGN_MAX_CODE) return fail(DSPGN_E_ARG, "latent_size must be <= 64");
  const int in0 = s.latent_size + 3;
  if (s.in_dim[0] != in0) return fail(DSPGN_E_ARG, "in_dim[0] must equal latent_size+3");
  if (s.out_dim[s.num_linear - 1] != 1) return fail(DSPGN_E_ARG, "last layer must have one output");
  if (s.latent_in_layer != -1 && (s.latent_in_layer < 1 || s.latent_in_layer > s.num_linear - 1))
    return fail(DSPGN_E_ARG, "latent_in_layer must be a layer index >= 1 or -1");
  for (int k = 0; k < s.num_linear; ++k) {
    if (s.in_dim[k] < 1 || s.in_dim[k] > kHid || s.out_dim[k] < 1 || s.out_dim[k] > kHid)
      return fail(DSPGN_E_ARG, "layer widths must be in [1,256]");
    const int ck = cat_kind_of(s, k);
    if (ck < 0 || ck > 2 || (k == 0 && ck != 0)) return fail(DSPGN_E_ARG, "bad cat_kind");
    if (s.layer_norm[k] != 0 && k == s.num_linear - 1) return fail(DSPGN_E_ARG, "the last layer cannot be normalised");
    if (k > 0) {
      const int expect = s.out_dim[k - 1] + (ck == 1 ? in0 : (ck == 2 ? 3 : 0));
      if (s.in_dim[k] != expect) return fail(DSPGN_E_ARG, "layer in_dim inconsistent with previous out_dim/latent_in");
    }
  }
  return 0;
}

int upload_vec(DspgnDecoder* d, const std::vector<float>& h, const float** out) {
  void* p = nullptr;
  CU(cudaMalloc(&p, h.size() * sizeof(float)));
  d->allocs.push_back(p);
  CU(cudaMemcpy(p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  *out = reinterpret_cast<const float*>(p);
  return 0;
}

}  // namespace

extern "C" {

const char* dspgn_last_error(void) { return g_err.c_str(); }
int dspgn_version(void) { return 100; }

int dspgn_decoder_create(const DspgnDecoderSpec* spec, const float* const* W, const float* const* b,
                         int device, DspgnDecoder** out) {
  return dspgn_decoder_create_ex(spec, W, b, nullptr, nullptr, device, out);
}

int dspgn_decoder_create_ex(const DspgnDecoderSpec* spec, const float* const* W, const float* const* b,
                            const float* const* ln_gamma, const float* const* ln_beta, int device, DspgnDecoder** out) {
  if (!spec || !W || !b || !out) return fail(DSPGN_E_ARG, "null argument");
  for (int k = 0; k < spec->num_linear && k < DSPGN_MAX_LINEAR; ++k)
    if (spec->layer_norm[k] && (!ln_gamma || !ln_beta || !ln_gamma[k] || !ln_beta[k])) return fail(DSPGN_E_ARG, "LayerNorm parameters missing");
  if (int rc = check_spec(*spec)) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(DSPGN_E_NOGPU, "no CUDA device"); }
  if (device < 0 || device >= ndev) return fail(DSPGN_E_ARG, "bad device index");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(DSPGN_E_NOGPU, "libdspgn is built for sm_100a (B200) only; found sm_" + std::to_string(prop.major) + std::to_string(prop.minor));
  DspgnDecoder* d = new (std::nothrow) DspgnDecoder();
  if (!d) return fail(DSPGN_E_ALLOC, "oom");
  d->device = device;
  d->spec = *spec;
  DecoderDev& dv = d->dev;
  dv.L = spec->latent_size; dv.n_lin = spec->num_linear; dv.in0 = spec->latent_size + 3;
  const int nl = spec->num_linear;
  // variants: the plain shape (at most one latent_in layer among the hidden layers, nothing else) runs on both engines
  int n_cat1 = 0, cat1_layer = -1;
  bool generic = spec->use_tanh != 0;
  for (int k = 0; k < nl; ++k) {
    dv.cat_kind[k] = cat_kind_of(*spec, k);
    if (dv.cat_kind[k] == 1) { ++n_cat1; cat1_layer = k; }
    if (dv.cat_kind[k] == 2 || spec->layer_norm[k]) generic = true;
  }
  if (n_cat1 > 1 || (n_cat1 == 1 && cat1_layer > nl - 2)) generic = true;
  dv.latent_in = (n_cat1 == 1) ? cat1_layer : -1;
  dv.use_tanh = spec->use_tanh != 0; dv.generic = generic ? 1 : 0;
  d->has_ln = false;
  int rc = 0;
  for (int k = 0; k < nl && rc == 0; ++k) {
    const int nin = spec->in_dim[k], nout = spec->out_dim[k];
    dv.in_dim[k] = nin; dv.out_dim[k] = nout;
    const int in_pad = (nin + kKC - 1) / kKC * kKC, out_pad = (nout + kKC - 1) / kKC * kKC;
    std::vector<float> wf((size_t)in_pad * kHid, 0.f), wb((size_t)out_pad * kHid, 0.f), bb(kHid, 0.f);
    for (int j = 0; j < nout; ++j) {
      for (int i = 0; i < nin; ++i) {
        const float w = W[k][(size_t)j * nin + i];
        wf[(size_t)i * kHid + j] = w;
        wb[(size_t)j * kHid + i] = w;
      }
      bb[j] = b[k][j];
    }
    rc = upload_vec(d, wf, &dv.Wf[k]);
    if (!rc) rc = upload_vec(d, wb, &dv.Wb[k]);
    if (!rc) rc = upload_vec(d, bb, &dv.bias[k]);
    if (!rc && k == nl - 1) {
      std::vector<float> wl(kHid, 0.f);
      for (int i = 0; i < nin; ++i) wl[i] = W[k][i];
      rc = upload_vec(d, wl, &dv.w_last);
    }
    if (!rc && spec->layer_norm[k]) {
      std::vector<float> g(kHid, 0.f), be(kHid, 0.f);
      for (int j = 0; j < nout; ++j) { g[j] = ln_gamma[k][j]; be[j] = ln_beta[k][j]; }
      rc = upload_vec(d, g, &dv.ln_gamma[k]);
      if (!rc) rc = upload_vec(d, be, &dv.ln_beta[k]);
      d->has_ln = true;
    }
  }
  if (!rc && !generic) rc = tc_pack_decoder(*spec, W, b, d->tc, &dv, g_err);     // the tcgen05 engine covers the plain shape
  if (rc) { dspgn_decoder_destroy(d); return rc; }
  *out = d;
  return 0;
}

void dspgn_decoder_destroy(DspgnDecoder* d) {
  if (!d) return;
  cudaSetDevice(d->device);
  for (void* p : d->allocs) cudaFree(p);
  tc_free_decoder(d->tc);
  delete d;
}

int dspgn_solver_create(const DspgnConfig* cfg, DspgnDecoder* const* classes, int n_classes, int device,
                        DspgnSolver** out) {
  if (!cfg || !classes || !out) return fail(DSPGN_E_ARG, "null argument");
  if (n_classes < 1 || n_classes > DSPGN_MAX_CLASSES) return fail(DSPGN_E_ARG, "n_classes must be in [1,4]");
  if (cfg->num_depth_samples < 2 || cfg->num_depth_samples > 64) return fail(DSPGN_E_ARG, "num_depth_samples must be in [2,64]");
  if (cfg->num_iterations < 1) return fail(DSPGN_E_ARG, "num_iterations must be >= 1");
  for (int c = 0; c < n_classes; ++c) {
    if (!classes[c] || classes[c]->device != device) return fail(DSPGN_E_ARG, "decoder/device mismatch");
    if (cfg->code_len < 1 || cfg->code_len > classes[c]->spec.latent_size) return fail(DSPGN_E_ARG, "code_len exceeds decoder latent_size");
    // code_len < latent_size: the trailing latent entries stay zero and are not optimised (optimizer.py:97-100
    // slices code[:code_len]; the reference itself needs code_len == latent size for its decoder input)
  }
  CU(cudaSetDevice(device));
  DspgnSolver* s = new (std::nothrow) DspgnSolver();
  if (!s) return fail(DSPGN_E_ALLOC, "oom");
  // every failure below releases the half-built solver (events, streams, device buffers)
#undef CU
#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      const int rc_ = fail(DSPGN_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
      dspgn_solver_destroy(s);                                                            \
      return rc_;                                                                         \
    }                                                                                     \
  } while (0)
  s->device = device;
  s->cfg = *cfg;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  s->num_sms = prop.multiProcessorCount;
  if (const char* g = getenv("DSPGN_GRID")) { int v = atoi(g); if (v >= 1 && v <= s->num_sms) s->num_sms = v; }   // experiments only
  for (int c = 0; c < n_classes; ++c) s->classes.push_back(classes[c]);
  std::vector<DecoderDev> decs;
  for (auto* d : s->classes) decs.push_back(d->dev);
  if (s->d_decs.reserve(decs.size() * sizeof(DecoderDev))) { dspgn_solver_destroy(s); return fail(DSPGN_E_ALLOC, "cudaMalloc"); }
  CU(cudaMemcpy(s->d_decs.p, decs.data(), decs.size() * sizeof(DecoderDev), cudaMemcpyHostToDevice));
  bool tc_ok = true;
  for (auto* d : s->classes) tc_ok = tc_ok && d->tc.ok;
  int eng = cfg->engine;
  if (eng == DSPGN_ENGINE_AUTO) {
    const char* e = getenv("DSPGN_ENGINE");
    if (e && !strcmp(e, "simt")) eng = DSPGN_ENGINE_SIMT;
    else if (e && !strcmp(e, "tc")) eng = DSPGN_ENGINE_TC;
    else eng = (tc_ok && tc_engine_default()) ? DSPGN_ENGINE_TC : DSPGN_ENGINE_SIMT;
  }
  if (eng == DSPGN_ENGINE_TC && !tc_ok) { dspgn_solver_destroy(s); return fail(DSPGN_E_ARG, "tensor-core engine unavailable for this decoder shape"); }
  s->engine = eng;
  CU(cudaFuncSetAttribute(k_decoder_simt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SimtSmem)));
  for (auto* d : s->classes)
    if (d->has_ln) {      // LayerNorm decoders: per-CTA scratch for the normalised activations (forward -> backward)
      if (s->d_ln.reserve(4 * (size_t)s->num_sms * DSPGN_MAX_LINEAR * kHid * kTP)) { dspgn_solver_destroy(s); return fail(DSPGN_E_ALLOC, "cudaMalloc"); }
      break;
    }
  if (int rc = tc_setup_kernels(g_err)) { dspgn_solver_destroy(s); return rc; }
  if (const char* m = getenv("DSPGN_MEGA")) s->mega_enabled = (m[0] != '0');
  if (const char* m = getenv("DSPGN_COMPACT_RAYS")) s->compact_rays = (m[0] != '0');   // A/B switch of the valid-sample hulls
  if (const char* m = getenv("DSPGN_VPRE_EXACT")) s->vpre_exact = (m[0] != '0');
  if (cfg->schedule == DSPGN_SCHED_LAUNCHES) s->mega_enabled = false;
  else if (cfg->schedule == DSPGN_SCHED_PERSISTENT) s->mega_enabled = true;
  else if (cfg->schedule != DSPGN_SCHED_AUTO) { dspgn_solver_destroy(s); return fail(DSPGN_E_ARG, "bad schedule"); }
  if (getenv("DSPGN_CLK")) {
    const size_t nb = sizeof(long long) * (kClkTiles * kTcMaxSteps * kClkSlots + 16);
    if (s->d_clk.reserve(nb)) { dspgn_solver_destroy(s); return fail(DSPGN_E_ALLOC, "cudaMalloc"); }
    CU(cudaMemset(s->d_clk.p, 0, nb));
    s->clk_on = true;
  }
  CU(cudaEventCreateWithFlags(&s->ev_upload, cudaEventDisableTiming));
  CU(cudaStreamCreateWithFlags(&s->stream2, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
  CU(cudaEventCreate(&s->ev_run0));
  CU(cudaEventCreate(&s->ev_run1));
#undef CU
#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(DSPGN_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
  } while (0)
  *out = s;
  return 0;
}

void dspgn_solver_destroy(DspgnSolver* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  cudaDeviceSynchronize();
  dspgn_gather_close(s);
  for (DevBuf* b : {&s->d_decs, &s->d_stage, &s->d_state, &s->d_part_s, &s->d_part_r, &s->d_tbase, &s->d_V, &s->d_m, &s->d_results, &s->d_active,
                    &s->d_sdf, &s->d_bx, &s->d_bs, &s->d_br, &s->d_dbg, &s->d_clk, &s->d_q_flag, &s->d_q_ctr,
                    &s->d_tiles_left, &s->d_obj_iter, &s->d_ev, &s->d_seg, &s->d_ln, &s->d_vpre}) b->release();
  s->h_stage.release();
  s->h_results.release();
  for (auto e : s->ev) cudaEventDestroy(e);
  for (auto e : s->ev_solve) cudaEventDestroy(e);
  if (s->ev_upload) cudaEventDestroy(s->ev_upload);
  if (s->ev_fork) cudaEventDestroy(s->ev_fork);
  if (s->ev_join) cudaEventDestroy(s->ev_join);
  if (s->stream2) cudaStreamDestroy(s->stream2);
  if (s->ev_run0) cudaEventDestroy(s->ev_run0);
  if (s->ev_run1) cudaEventDestroy(s->ev_run1);
  delete s;
}

int dspgn_solver_set_stream(DspgnSolver* s, void* cuda_stream) {
  if (!s) return fail(DSPGN_E_ARG, "null solver");
  s->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  return 0;
}

int dspgn_solver_engine(const DspgnSolver* s) { return s ? s->engine : DSPGN_E_ARG; }

int dspgn_solver_sync(DspgnSolver* s) {
  if (!s) return fail(DSPGN_E_ARG, "null solver");
  CU(cudaSetDevice(s->device));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

int dspgn_enable_timing(DspgnSolver* s, int on) { if (!s) return fail(DSPGN_E_ARG, "null solver"); s->timing = on != 0; return 0; }

int dspgn_counters(DspgnSolver* s, DspgnCounters* out) {
  if (!s || !out) return fail(DSPGN_E_ARG, "null argument");
  // device time of the last run's kernels, if they have finished (events on the solver's stream)
  float tot = 0.f;
  if (s->ev_run0 && s->ev_run1 && cudaEventElapsedTime(&tot, s->ev_run0, s->ev_run1) == cudaSuccess) s->ctr.total_ms = tot;
  else cudaGetLastError();
  *out = s->ctr;
  return 0;
}

// ---------------------------------------------------------------------------------------------
namespace {

// decode_only: forward-only use (dspgn_decode_sdf) -- no J^T J partials, no band buffers
int upload_batch_impl(DspgnSolver* s, int n_obj, const DspgnObjectIn* in, bool decode_only) {
  if (!s || !in) return fail(DSPGN_E_ARG, "null argument");
  if (n_obj < 1 || n_obj > kMaxObjScan) return fail(DSPGN_E_ARG, "n_obj must be in [1,1024] per resident batch");
  CU(cudaSetDevice(s->device));
  const int D = s->cfg.num_depth_samples;
  s->h_meta.assign(n_obj, ObjMeta{});
  long long tp = 0, tr = 0, tf = 0, ts = 0;
  int max_rays = 0, n_bad = 0, any_build = 0;
  for (int o = 0; o < n_obj; ++o) {
    const DspgnObjectIn& I = in[o];
    // misuse of the API fails the call; an unusable DETECTION only fails that object (status BAD_INPUT ->
    // is_good=False), like the reference's soft exits (optimizer.py:130-150): one bad object must neither abort
    // its batch neighbours nor raise inside the embedded interpreter
    if (!I.t_cam_obj) return fail(DSPGN_E_ARG, "object without a pose");
    if (I.class_id < 0 || I.class_id >= (int)s->classes.size()) return fail(DSPGN_E_ARG, "bad class_id");
    const bool bad = I.n_pts < 1 || !I.pts || I.n_rays < 0 || I.n_depth < 0 || I.n_depth > I.n_rays ||
                     I.n_rays > kScanMaxRays || (I.n_rays > 0 && I.n_depth > 0 && !I.depth) || (I.pixels && !I.inv_k);
    const float* ray_src = I.pixels ? I.pixels : I.rays;
    ObjMeta& M = s->h_meta[o];
    M.bad = bad ? 1 : 0;
    M.pts_off = (int)tp; M.n_pts = bad ? 0 : I.n_pts;
    M.ray_off = (int)tr; M.n_rays = (!bad && ray_src) ? I.n_rays : 0; M.n_fg = (!bad && ray_src) ? I.n_depth : 0;
    M.build = bad ? 0 : ((I.pixels ? 1 : 0) | (I.t_cam_world ? 2 : 0));
    any_build |= M.build;
    M.fg_off = (int)tf; M.smp_off = (int)ts;
    M.class_id = I.class_id; M.scale = I.scale; M.has_code = I.code != nullptr;
    tp += M.n_pts; tr += M.n_rays; tf += M.n_fg; ts += (long long)M.n_rays * D;
    if (M.n_rays > max_rays) max_rays = M.n_rays;
    n_bad += M.bad;
  }
  s->n_bad = n_bad;
  if (tp > (1 << 28) || ts > (1LL << 30)) return fail(DSPGN_E_ARG, "batch too large");
  // one staging block: meta | T_init | code | pts | rays | depth
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t o_meta = 0, o_T = al(o_meta + sizeof(ObjMeta) * n_obj), o_code = al(o_T + 64 * n_obj),
               o_pts = al(o_code + 4 * kMaxCode * (size_t)n_obj), o_rays = al(o_pts + 12 * (size_t)tp),
               o_depth = al(o_rays + 12 * (size_t)tr), o_tb = al(o_depth + 4 * (size_t)tf),
               o_aux = al(o_tb + 3 * 4 * (size_t)n_obj),
               total = al(o_aux + (any_build ? 4 * (size_t)kAuxFloats * n_obj : 0));
  if (s->upload_pending) { CU(cudaEventSynchronize(s->ev_upload)); s->upload_pending = false; }
  if (s->d_stage.cap < total) CU(cudaStreamSynchronize(s->stream));      // kernels of an earlier batch may still read the old block
  if (s->h_stage.reserve(total) || s->d_stage.reserve(total)) return fail(DSPGN_E_ALLOC, "staging allocation failed");
  unsigned char* hb = s->h_stage.as<unsigned char>();
  memcpy(hb + o_meta, s->h_meta.data(), sizeof(ObjMeta) * n_obj);
  float* hT = reinterpret_cast<float*>(hb + o_T);
  float* hC = reinterpret_cast<float*>(hb + o_code);
  float* hP = reinterpret_cast<float*>(hb + o_pts);
  float* hR = reinterpret_cast<float*>(hb + o_rays);
  float* hD = reinterpret_cast<float*>(hb + o_depth);
  int* hTB = reinterpret_cast<int*>(hb + o_tb);
  {
    int acc = 0, mx = 0;
    long long accr = 0, accq = 0;
    int* hTBr = hTB + n_obj;
    int* hQ0 = hTB + 2 * n_obj;
    for (int o = 0; o < n_obj; ++o) {
      const int nt = (s->h_meta[o].n_pts + kTcRows - 1) / kTcRows;
      const int ntf = (int)(((long long)s->h_meta[o].n_rays * D + kTcRows - 1) / kTcRows);
      hTB[o] = acc; hTBr[o] = (int)(accr + o); hQ0[o] = (int)accq;
      acc += nt; accr += ntf; accq += nt + ntf;
      if (nt > mx) mx = nt;
      if (ntf > mx) mx = ntf;
    }
    s->total_tiles128 = acc; s->total_ray_tiles128 = accr; s->max_tiles128 = mx;
  }
  for (int o = 0; o < n_obj; ++o) {
    const DspgnObjectIn& I = in[o];
    const ObjMeta& M = s->h_meta[o];
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) hT[o * 16 + r * 4 + c] = I.t_cam_obj[(size_t)r * I.t_rs + (size_t)c * I.t_cs];
    for (int i = 0; i < kMaxCode; ++i) hC[o * kMaxCode + i] = (I.code && i < s->cfg.code_len) ? I.code[i] : 0.f;
    float* p = hP + 3 * (size_t)M.pts_off;
    if (M.n_pts > 0 && I.pts_cs == 1 && I.pts_rs == 3) memcpy(p, I.pts, 12 * (size_t)M.n_pts);
    else for (int r = 0; r < M.n_pts; ++r)
      for (int c = 0; c < 3; ++c) p[3 * (size_t)r + c] = I.pts[(size_t)r * I.pts_rs + (size_t)c * I.pts_cs];
    float* q = hR + 3 * (size_t)M.ray_off;
    if (M.build & 1) {                      // pixel coordinates: the device turns (u, v, 1) into inv_k [u, v, 1]
      for (int r = 0; r < M.n_rays; ++r) {
        q[3 * (size_t)r] = I.pixels[(size_t)r * I.pix_rs]; q[3 * (size_t)r + 1] = I.pixels[(size_t)r * I.pix_rs + I.pix_cs];
        q[3 * (size_t)r + 2] = 1.f;
      }
    } else
    for (int r = 0; r < M.n_rays; ++r)
      for (int c = 0; c < 3; ++c) q[3 * (size_t)r + c] = I.rays[(size_t)r * I.rays_rs + (size_t)c * I.rays_cs];
    if (any_build) {
      float* ax = reinterpret_cast<float*>(hb + o_aux) + (size_t)o * kAuxFloats;
      for (int i = 0; i < kAuxFloats; ++i) ax[i] = 0.f;
      if (M.build & 1) for (int i = 0; i < 9; ++i) ax[i] = I.inv_k[i];
      if (M.build & 2) for (int i = 0; i < 12; ++i) ax[9 + i] = I.t_cam_world[i];
    }
    if (M.n_fg) memcpy(hD + M.fg_off, I.depth, 4 * (size_t)M.n_fg);
  }
  CU(cudaMemcpyAsync(s->d_stage.p, hb, total, cudaMemcpyHostToDevice, s->stream));
  CU(cudaEventRecord(s->ev_upload, s->stream));
  s->upload_pending = true;
  unsigned char* db = s->d_stage.as<unsigned char>();
  if (any_build) {                          // rays from pixels / world points -> camera frame, once per upload, in place
    BuildArgs ba{};
    ba.meta = reinterpret_cast<ObjMeta*>(db + o_meta); ba.T_init = reinterpret_cast<float*>(db + o_T);
    ba.pts = reinterpret_cast<float*>(db + o_pts); ba.rays = reinterpret_cast<float*>(db + o_rays);
    ba.aux = reinterpret_cast<const float*>(db + o_aux); ba.n_obj = n_obj;
    k_build_inputs<<<n_obj, 256, 0, s->stream>>>(ba);
    CU(cudaGetLastError());
  }
  s->d_meta = reinterpret_cast<ObjMeta*>(db + o_meta);
  s->d_Tinit = reinterpret_cast<float*>(db + o_T);
  s->d_code = reinterpret_cast<float*>(db + o_code);
  s->d_pts = reinterpret_cast<float*>(db + o_pts);
  s->d_rays = reinterpret_cast<float*>(db + o_rays);
  s->d_depth = reinterpret_cast<float*>(db + o_depth);
  s->d_tbase_static = reinterpret_cast<int*>(db + o_tb);
  s->d_tbase_r_static = s->d_tbase_static + n_obj;
  s->d_q0_render = s->d_tbase_static + 2 * n_obj;
  s->n_obj = n_obj; s->tot_pts = (int)tp; s->tot_rays = (int)tr; s->tot_fg = (int)tf; s->tot_smp = ts; s->max_rays = max_rays;
  s->gather.bound_n = -1;
  int bad = 0;
  bad |= s->d_state.reserve(sizeof(ObjState) * n_obj);
  const bool render = !decode_only && !s->cfg.sdf_only;
  if (!decode_only) {
    // per-tile partial sums: one slot per possible tile of each term at the engine's tile height
    const size_t rows_per_tile = (s->engine == DSPGN_ENGINE_TC) ? kTcRows : kTP;
    const size_t tiles_s = (size_t)tp / rows_per_tile + n_obj + 1, tiles_r = (size_t)ts / rows_per_tile + 2 * (size_t)n_obj + 1;
    bad |= s->d_part_s.reserve(4 * (size_t)kAccStride * tiles_s);
    if (render) bad |= s->d_part_r.reserve(4 * (size_t)kAccStride * tiles_r);
    bad |= s->d_active.reserve((size_t)tp + 1);
  }
  bad |= s->d_tbase.reserve(4 * 2 * (size_t)n_obj);
  bad |= s->d_V.reserve(4 * (size_t)n_obj);
  bad |= s->d_m.reserve(4 * (size_t)n_obj);
  bad |= s->d_results.reserve(4 * DSPGN_RESULT_FLOATS * (size_t)n_obj);
  bad |= s->h_results.reserve(4 * DSPGN_RESULT_FLOATS * (size_t)n_obj + 512);
  const size_t smp = (size_t)(ts > 0 ? ts : 1);
  if (render) {
    bad |= s->d_sdf.reserve(4 * smp);
    bad |= s->d_bx.reserve(12 * smp);
    bad |= s->d_bs.reserve(4 * smp);
    bad |= s->d_br.reserve(4 * smp);
  }
  bad |= s->d_dbg.reserve(4 * ((size_t)kPMax * kPMax + 2 * kPMax + 8));
  if (bad) return fail(DSPGN_E_ALLOC, "workspace allocation failed");
  return 0;
}

}  // namespace

int dspgn_upload_batch(DspgnSolver* s, int n_obj, const DspgnObjectIn* in) {
  return upload_batch_impl(s, n_obj, in, false);
}

namespace {

cudaEvent_t next_event(DspgnSolver* s) {
  if (s->ev_used == s->ev.size()) { cudaEvent_t e; cudaEventCreate(&e); s->ev.push_back(e); }
  return s->ev[s->ev_used++];
}

int launch_term(DspgnSolver* s, TermArgs& a, long long rows_upper, cudaStream_t stream = nullptr, bool use_given = false) {
  cudaStream_t st = use_given ? stream : s->stream;
  const bool timed = s->timing && !use_given;
  if (timed) cudaEventRecord(next_event(s), st);
  const long long tile_rows = (s->engine == DSPGN_ENGINE_TC) ? kTcRows : kTP;
  long long tiles = (rows_upper + tile_rows - 1) / tile_rows + s->n_obj;
  if (s->engine == DSPGN_ENGINE_TC) {
    if (int rc = tc_launch_term(a, s->num_sms, tiles, st, g_err)) return rc;
  } else {
    int grid = (int)std::min<long long>(tiles, s->num_sms);
    if (grid < 1) grid = 1;
    k_decoder_simt<<<grid, kThreads, sizeof(SimtSmem), st>>>(a);
  }
  if (timed) cudaEventRecord(next_event(s), st);
  s->ctr.kernel_launches++;
  CU(cudaGetLastError());
  return 0;
}

TermArgs base_term(DspgnSolver* s, int mode) {
  TermArgs a{};
  a.meta = s->d_meta; a.state = s->d_state.as<ObjState>(); a.decs = s->d_decs.as<DecoderDev>();
  a.n_obj = s->n_obj; a.n_classes = (int)s->classes.size(); a.mode = mode;
  a.pts = s->d_pts; a.pt_active = nullptr; a.pt_active_out = nullptr; a.cut_iter = -1; a.rays = s->d_rays;
  a.band_x = s->d_bx.as<float>(); a.band_s = s->d_bs.as<float>(); a.band_r = s->d_br.as<float>();
  a.band_m = s->d_m.as<int>(); a.sdf_out = s->d_sdf.as<float>(); a.V_count = s->d_V.as<int>();
  a.part = (mode == MODE_BAND) ? s->d_part_r.as<float>() : (mode == MODE_SDF ? s->d_part_s.as<float>() : nullptr);
  a.tile_base = (mode == MODE_BAND) ? s->d_tbase.as<int>() + s->n_obj : (mode == MODE_SDF ? s->d_tbase.as<int>() : nullptr);
  a.D = s->cfg.num_depth_samples;
  a.ln_scratch = s->d_ln.as<float>();
  a.dbg_J = nullptr; a.dbg_res = nullptr; a.dbg_obj = -1; a.dbg_P = 0;
  a.dbg_clk = (s->clk_on && mode == MODE_SDF) ? s->d_clk.as<long long>() : nullptr;
  return a;
}

int launch_init(DspgnSolver* s, int pose_only, bool mega = false, bool render = false) {
  InitArgs ia{};
  ia.meta = s->d_meta; ia.state = s->d_state.as<ObjState>(); ia.T_init = s->d_Tinit; ia.code_init = s->d_code;
  ia.V_count = s->d_V.as<int>(); ia.band_m = s->d_m.as<int>();
  ia.pt_active = nullptr; ia.n_obj = s->n_obj; ia.code_len = s->cfg.code_len; ia.D = s->cfg.num_depth_samples;
  ia.pose_only = pose_only;
  ia.gather = s->gdev; ia.results = s->d_results.as<float>(); ia.n_bad = s->n_bad;
  ia.decs = s->d_decs.as<DecoderDev>();
  ia.mega = mega ? 1 : 0;
  if (mega) {
    ia.render = render ? 1 : 0;
    ia.q0_off = render ? s->d_q0_render : s->d_tbase_static; ia.tile_rows = kTcRows;
    ia.q_flag = s->d_q_flag.as<int>();
    ia.q_head = s->d_q_ctr.as<int>(); ia.q_tail = s->d_q_ctr.as<int>() + 32; ia.done_objects = s->d_q_ctr.as<int>() + 64;
    ia.band_rows_total = s->d_q_ctr.as<int>() + 80; ia.abort_flag = s->d_q_ctr.as<int>() + 96;
    ia.pending = s->d_tiles_left.as<int>(); ia.ray_left = s->d_tiles_left.as<int>() + s->n_obj;
    ia.obj_iter = s->d_obj_iter.as<int>();
    ia.total_tiles0 = s->total_tiles128 + (render ? (int)s->total_ray_tiles128 : 0);
    ia.vpre_exact = s->vpre_exact ? 1 : 0;
    ia.rays = s->d_rays; ia.vpre = (render && s->compact_rays) ? s->d_vpre.as<int>() : nullptr;
    ia.valid_rows_total = reinterpret_cast<unsigned long long*>(s->d_q_ctr.as<int>() + 88);
  }
  k_init<<<s->n_obj, 128, 0, s->stream>>>(ia);
  s->ctr.kernel_launches++;
  CU(cudaGetLastError());
  return 0;
}

ScanArgs base_scan(DspgnSolver* s) {
  const DspgnConfig& c = s->cfg;
  ScanArgs sa{};
  sa.meta = s->d_meta; sa.state = s->d_state.as<ObjState>(); sa.rays = s->d_rays; sa.depth_fg = s->d_depth;
  sa.sdf = s->d_sdf.as<float>(); sa.band_x = s->d_bx.as<float>(); sa.band_s = s->d_bs.as<float>();
  sa.band_r = s->d_br.as<float>(); sa.band_m = s->d_m.as<int>(); sa.th = c.cut_off; sa.D = c.num_depth_samples;
  sa.n_obj = s->n_obj;
  sa.V_count = s->d_V.as<int>();
  return sa;
}

// one GN iteration's residual-term kernels (everything before the solve)
int launch_terms(DspgnSolver* s, int pose_only, float* dbg_J, float* dbg_res, int dbg_obj, int iter_index = 0) {
  const DspgnConfig& c = s->cfg;
  const bool render = !pose_only && !c.sdf_only;
  // The SDF-row pass and the forward-only pass over the ray samples are independent: fork the latter onto a
  // second stream (matters for small batches, where each pass is a single wave of tiles) unless per-launch
  // timing is on.
  const bool fork = render && !s->timing;
  if (render) {
    TermArgs f = base_term(s, MODE_RAYFWD);
    if (fork) {
      CU(cudaEventRecord(s->ev_fork, s->stream));
      CU(cudaStreamWaitEvent(s->stream2, s->ev_fork, 0));
      if (int rc = launch_term(s, f, s->tot_smp, s->stream2, true)) return rc;
      CU(cudaEventRecord(s->ev_join, s->stream2));
    } else {
      if (int rc = launch_term(s, f, s->tot_smp)) return rc;
    }
    s->ctr.rows_fwd_only += s->tot_smp;
  }
  {
    TermArgs a = base_term(s, MODE_SDF);
    a.huber_b = pose_only ? INFINITY : c.b2;       // optimizer.py:71 uses raw residuals
    a.pose_only = pose_only;
    if (pose_only) {                                   // optimizer.py:76-78: inlier cut taken after iteration index 4
      if (iter_index == 4) a.pt_active_out = s->d_active.as<uint8_t>();
      if (iter_index > 4) a.pt_active = s->d_active.as<uint8_t>();
    }
    a.dbg_J = dbg_J; a.dbg_res = dbg_res; a.dbg_obj = dbg_obj; a.dbg_P = pose_only ? 6 : 7 + c.code_len;
    if (int rc = launch_term(s, a, s->tot_pts)) return rc;
    s->ctr.rows_fwd_bwd += s->tot_pts;
  }
  if (render) {
    if (fork) CU(cudaStreamWaitEvent(s->stream, s->ev_join, 0));
    ScanArgs sa = base_scan(s);
    k_ray_scan<<<s->n_obj, kScanThreads, 0, s->stream>>>(sa);
    s->ctr.kernel_launches++;
    CU(cudaGetLastError());
    TermArgs b = base_term(s, MODE_BAND);
    b.huber_b = c.b1;
    if (int rc = launch_term(s, b, s->tot_smp)) return rc;
  }
  return 0;
}

SolveArgs base_solve(DspgnSolver* s, int pose_only) {
  const DspgnConfig& c = s->cfg;
  SolveArgs v{};
  v.meta = s->d_meta; v.state = s->d_state.as<ObjState>();
  v.part_s = s->d_part_s.as<float>(); v.part_r = s->d_part_r.as<float>();
  v.base_s = s->d_tbase.as<int>(); v.base_r = s->d_tbase.as<int>() + s->n_obj;
  v.tile_rows = (s->engine == DSPGN_ENGINE_TC) ? kTcRows : kTP;
  v.V_count = s->d_V.as<int>(); v.band_m = s->d_m.as<int>();
  v.prm = SolverParams{c.k1, c.k2, c.k3, c.k4, c.b1, c.b2, c.lr, c.s_damp, c.code_len, c.num_depth_samples, c.cut_off, c.sdf_only};
  v.n_obj = s->n_obj; v.pose_only = pose_only; v.results = s->d_results.as<float>();
  v.gather = s->gdev;
  v.decs = s->d_decs.as<DecoderDev>();
  v.dbg_obj = -1; v.dbg_H = nullptr; v.dbg_b = nullptr; v.dbg_dx = nullptr; v.dbg_loss = nullptr;
  v.dbg_clk = s->clk_on ? s->d_clk.as<long long>() + kClkTiles * kTcMaxSteps * kClkSlots : nullptr;
  return v;
}

}  // namespace

namespace {
int run_batch_impl(DspgnSolver* s, int mode) {
  if (!s) return fail(DSPGN_E_ARG, "null solver");
  if (s->n_obj < 1) return fail(DSPGN_E_ARG, "no batch uploaded");
  if (mode != 0 && mode != 1) return fail(DSPGN_E_ARG, "mode must be 0 or 1");
  CU(cudaSetDevice(s->device));
  const int pose_only = mode;
  const int iters = pose_only ? s->cfg.pose_only_iterations : s->cfg.num_iterations;
  s->ctr = DspgnCounters{};
  s->band_rows_pending = false;
  s->ev_used = 0;
  s->evs_used = 0;
  CU(cudaEventRecord(s->ev_run0, s->stream));
  const bool render = !pose_only && !s->cfg.sdf_only;
  // queue capacity: per iteration every SDF tile, every ray-sample tile and at most as many band tiles again
  const long long items_per_iter = (long long)s->total_tiles128 +
                                   (render ? 2 * s->total_ray_tiles128 + s->tot_rays / kScanChunkRays + s->n_obj : 0);
  const bool mega = s->mega_enabled && s->engine == DSPGN_ENGINE_TC && s->total_tiles128 > 0 &&
                    s->max_tiles128 <= kItemTileMask && s->n_obj <= kItemObjMask + 1 && items_per_iter * iters < (1LL << 27);
  if (mega) {
    // ---- persistent object-pipelined kernel: every GN iteration of every object in ONE launch --------------
    const int cap = (int)(items_per_iter * iters);
    int bad = 0;
    bad |= s->d_q_flag.reserve(4 * (size_t)cap);
    bad |= s->d_q_ctr.reserve(4 * 128);
    bad |= s->d_tiles_left.reserve(4 * 3 * (size_t)s->n_obj);
    const size_t nseg_cap = (size_t)s->tot_rays / kSegRays + 2 * (size_t)s->n_obj + 4;
    if (render) bad |= s->d_seg.reserve(4 * 2 * nseg_cap);
    if (render && s->compact_rays) bad |= s->d_vpre.reserve(4 * ((size_t)s->tot_rays + (size_t)s->n_obj + 4));
    bad |= s->d_obj_iter.reserve(4 * (size_t)s->n_obj);
    if (bad) return fail(DSPGN_E_ALLOC, "queue allocation failed");
    CU(cudaMemsetAsync(s->d_q_flag.p, 0, 4 * (size_t)cap, s->stream));
    if (int rc = launch_init(s, pose_only, true, render)) return rc;
    TermArgs a = base_term(s, MODE_SDF);
    a.huber_b = pose_only ? INFINITY : s->cfg.b2;
    a.huber_b1 = s->cfg.b1;
    a.pose_only = pose_only;
    a.tile_base = s->d_tbase_static;
    a.part_r = s->d_part_r.as<float>(); a.tile_base_r = s->d_tbase_r_static;
    a.dbg_clk = nullptr;
    if (pose_only && iters > 5) { a.pt_active_out = s->d_active.as<uint8_t>(); a.cut_iter = 4; }
    MegaArgs q{};
    q.n_iters = iters; q.q_cap = cap; q.render = render ? 1 : 0;
    q.q_flag = s->d_q_flag.as<int>();
    q.q_head = s->d_q_ctr.as<int>(); q.q_tail = s->d_q_ctr.as<int>() + 32; q.done_objects = s->d_q_ctr.as<int>() + 64;
    q.band_rows_total = s->d_q_ctr.as<int>() + 80;
    q.abort_flag = s->d_q_ctr.as<int>() + 96;
    q.pending = s->d_tiles_left.as<int>(); q.ray_left = s->d_tiles_left.as<int>() + s->n_obj; q.obj_iter = s->d_obj_iter.as<int>();
    q.scan_left = s->d_tiles_left.as<int>() + 2 * s->n_obj;
    q.seg_cnt = s->d_seg.as<int>(); q.seg_prefix = s->d_seg.as<int>() + nseg_cap;
    q.valid_rows_total = reinterpret_cast<unsigned long long*>(s->d_q_ctr.as<int>() + 88);
    q.vpre = (render && s->compact_rays) ? s->d_vpre.as<int>() : nullptr;
    q.vpre_exact = s->vpre_exact ? 1 : 0;
    if (s->clk_on) {
      if (s->d_ev.reserve(8 * (1 + 2 * (size_t)kEvCap))) return fail(DSPGN_E_ALLOC, "cudaMalloc");
      CU(cudaMemsetAsync(s->d_ev.p, 0, 8, s->stream));
      q.ev = s->d_ev.as<long long>(); q.ev_cap = kEvCap;
    }
    SolveArgs v = base_solve(s, pose_only);
    v.ev = q.ev; v.ev_cap = q.ev_cap;
    v.base_s = s->d_tbase_static; v.base_r = s->d_tbase_r_static; v.tile_rows = kTcRows; v.last_iter = 0; v.iter_index = 0; v.dbg_clk = nullptr;
    ScanArgs sa = base_scan(s);
    sa.vpre = q.vpre;
    if (s->timing) cudaEventRecord(next_event(s), s->stream);
    if (render) k_gn_persistent_render<<<s->num_sms, kTcThreads, kTcSmemBytes, s->stream>>>(a, q, v, sa);
    else k_gn_persistent<<<s->num_sms, kTcThreads, kTcSmemBytes, s->stream>>>(a, q, v);
    if (s->timing) cudaEventRecord(next_event(s), s->stream);
    s->ctr.kernel_launches += 1;
    s->ctr.rows_fwd_bwd += (long long)s->tot_pts * iters;
    s->band_rows_pending = render;           // band rows and valid ray samples are counted by the kernel (dspgn_results)
    s->mega_ran = true;
    CU(cudaGetLastError());
    CU(cudaEventRecord(s->ev_run1, s->stream));
    return 0;
  }
  if (int rc = launch_init(s, pose_only)) return rc;
  for (int e = 0; e < iters; ++e) {
    if (int rc = launch_terms(s, pose_only, nullptr, nullptr, -1, e)) return rc;
    SolveArgs v = base_solve(s, pose_only);
    v.last_iter = (e == iters - 1); v.iter_index = e;
    if (s->timing) {
      if (s->evs_used + 2 > s->ev_solve.size()) { cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); s->ev_solve.push_back(e0); s->ev_solve.push_back(e1); }
      cudaEventRecord(s->ev_solve[s->evs_used], s->stream);
    }
    k_solve<<<s->n_obj, kSolveThreads, 0, s->stream>>>(v);
    if (s->timing) { cudaEventRecord(s->ev_solve[s->evs_used + 1], s->stream); s->evs_used += 2; }
    s->ctr.kernel_launches++;
    CU(cudaGetLastError());
  }
  CU(cudaEventRecord(s->ev_run1, s->stream));
  return 0;
}

GatherDev gather_dev(DspgnSolver* s, int seq) {
  DspgnSolver::Gather& G = s->gather;
  GatherDev g{};
  g.slots = reinterpret_cast<float*>(G.base) + (size_t)(seq & 1) * G.n_slots * DSPGN_RESULT_FLOATS;
  g.slot_of = G.d_slot_of.as<int>();
  g.flags = reinterpret_cast<int*>(G.base + G.off_flags);
  g.ack = reinterpret_cast<int*>(G.base + G.off_ack);
  g.err = G.d_local.as<int>();
  g.wait_ns = reinterpret_cast<long long*>(G.d_local.as<unsigned char>() + 8);
  g.rank = G.rank; g.world = G.world; g.seq = seq;
  return g;
}

int gather_layout(DspgnSolver* s, int n_slots, int world, int rank) {
  DspgnSolver::Gather& G = s->gather;
  G.n_slots = n_slots; G.world = world; G.rank = rank;
  const size_t slot_bytes = 2 * (size_t)n_slots * DSPGN_RESULT_FLOATS * 4;
  G.off_flags = (slot_bytes + 255) / 256 * 256;
  G.off_ack = G.off_flags + 4 * (size_t)((world + 63) / 64 * 64);
  if (G.d_local.reserve(64)) return fail(DSPGN_E_ALLOC, "cudaMalloc");
  CU(cudaMemset(G.d_local.p, 0, 64));
  return 0;
}
}  // namespace

int dspgn_run_batch(DspgnSolver* s, int mode) {
  if (s) s->gdev = GatherDev{};
  return run_batch_impl(s, mode);
}

// ---- multi-GPU result exchange ----------------------------------------------------------------------------------
int dspgn_gather_create(DspgnSolver* s, int n_slots, int world, DspgnIpcHandle* handle_out) {
  if (!s || !handle_out || n_slots < 1 || world < 1 || world > 1024) return fail(DSPGN_E_ARG, "bad gather arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) <= DSPGN_IPC_HANDLE_BYTES, "IPC handle size");
  CU(cudaSetDevice(s->device));
  dspgn_gather_close(s);
  if (int rc = gather_layout(s, n_slots, world, 0)) return rc;
  DspgnSolver::Gather& G = s->gather;
  const size_t bytes = G.off_ack + 256;
  void* p = nullptr;
  CU(cudaMalloc(&p, bytes));           // plain cudaMalloc: exportable through cudaIpcGetMemHandle
  G.base = reinterpret_cast<unsigned char*>(p); G.owner = true;
  CU(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  memset(handle_out, 0, sizeof(*handle_out));
  if (world > 1) {
    CU(cudaIpcGetMemHandle(&h, p));
    memcpy(handle_out->bytes, &h, sizeof(h));
  }
  G.active = true;
  return 0;
}

int dspgn_gather_open(DspgnSolver* s, const DspgnIpcHandle* handle, int n_slots, int world, int rank) {
  if (!s || !handle || n_slots < 1 || world < 2 || rank < 1 || rank >= world) return fail(DSPGN_E_ARG, "bad gather arguments");
  CU(cudaSetDevice(s->device));
  dspgn_gather_close(s);
  if (int rc = gather_layout(s, n_slots, world, rank)) return rc;
  DspgnSolver::Gather& G = s->gather;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle->bytes, sizeof(h));
  void* p = nullptr;
  CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));   // rank 0's HBM, reachable over NVLink
  G.base = reinterpret_cast<unsigned char*>(p); G.owner = false;
  G.active = true;
  return 0;
}

void dspgn_gather_close(DspgnSolver* s) {
  if (!s) return;
  DspgnSolver::Gather& G = s->gather;
  if (G.base) {
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    if (G.owner) cudaFree(G.base); else cudaIpcCloseMemHandle(G.base);
    cudaGetLastError();
  }
  G.base = nullptr; G.active = false; G.owner = false; G.bound_n = -1;
  G.d_slot_of.release(); G.d_local.release(); G.h_out.release();
  s->gdev = GatherDev{};
}

int dspgn_gather_bind(DspgnSolver* s, const int32_t* slots, int n) {
  if (!s || n < 0 || (n > 0 && !slots)) return fail(DSPGN_E_ARG, "bad argument");
  DspgnSolver::Gather& G = s->gather;
  if (!G.active) return fail(DSPGN_E_ARG, "no gather buffer (dspgn_gather_create / dspgn_gather_open first)");
  if (n > 0 && n != s->n_obj) return fail(DSPGN_E_ARG, "gather_bind: n must equal the resident batch size");
  for (int i = 0; i < n; ++i)
    if (slots[i] < 0 || slots[i] >= G.n_slots) return fail(DSPGN_E_ARG, "gather_bind: slot out of range");
  CU(cudaSetDevice(s->device));
  if (n > 0) {
    if (G.d_slot_of.cap < 4 * (size_t)n) CU(cudaStreamSynchronize(s->stream));
    if (G.d_slot_of.reserve(4 * (size_t)n)) return fail(DSPGN_E_ALLOC, "cudaMalloc");
    CU(cudaMemcpyAsync(G.d_slot_of.p, slots, 4 * (size_t)n, cudaMemcpyHostToDevice, s->stream));   // pageable source: returns after staging
  }
  G.bound_n = n;
  return 0;
}

int dspgn_run_batch_gather(DspgnSolver* s, int mode, int seq) {
  if (!s || seq < 1) return fail(DSPGN_E_ARG, "bad argument");
  DspgnSolver::Gather& G = s->gather;
  if (!G.active || G.bound_n < 0) return fail(DSPGN_E_ARG, "gather not bound for the resident batch");
  CU(cudaSetDevice(s->device));
  const GatherDev g = gather_dev(s, seq);
  if (G.bound_n > 0) {
    s->gdev = g;
    const int rc = run_batch_impl(s, mode);
    s->gdev = GatherDev{};
    if (rc) return rc;
  }
  k_gather_publish<<<1, 32, 0, s->stream>>>(g, G.bound_n == 0 ? 1 : 0);
  if (G.rank == 0) k_gather_wait<<<1, 32 * ((G.world + 31) / 32), 0, s->stream>>>(g);
  s->ctr.kernel_launches += (G.rank == 0) ? 2 : 1;
  CU(cudaGetLastError());
  return 0;
}

const float* dspgn_gather_device(DspgnSolver* s, int seq) {
  if (!s || !s->gather.active) return nullptr;
  return reinterpret_cast<const float*>(s->gather.base) + (size_t)(seq & 1) * s->gather.n_slots * DSPGN_RESULT_FLOATS;
}

int dspgn_gather_results(DspgnSolver* s, int seq, int n, DspgnObjectOut* out) {
  if (!s || !out || n < 1) return fail(DSPGN_E_ARG, "bad argument");
  DspgnSolver::Gather& G = s->gather;
  if (!G.active || G.rank != 0 || n > G.n_slots) return fail(DSPGN_E_ARG, "gather_results: rank 0 only, n <= n_slots");
  CU(cudaSetDevice(s->device));
  const size_t bytes = sizeof(DspgnObjectOut) * (size_t)n;
  if (G.h_out.reserve(bytes + 64)) return fail(DSPGN_E_ALLOC, "cudaMallocHost");
  CU(cudaMemcpyAsync(G.h_out.p, dspgn_gather_device(s, seq), bytes, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaMemcpyAsync(G.h_out.as<unsigned char>() + bytes, G.d_local.p, 4, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  memcpy(out, G.h_out.p, bytes);
  int err = 0;
  memcpy(&err, G.h_out.as<unsigned char>() + bytes, 4);
  if (err) { cudaMemset(G.d_local.p, 0, 4); return fail(DSPGN_E_PEER, "a rank did not publish its results within the timeout"); }
  return 0;
}

long long dspgn_gather_wait_ns(DspgnSolver* s) {
  if (!s || !s->gather.active) return -1;
  long long v[2] = {0, 0};
  cudaSetDevice(s->device);
  if (cudaStreamSynchronize(s->stream) != cudaSuccess) return -1;
  if (cudaMemcpy(v, s->gather.d_local.p, 16, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return v[1];
}

__global__ void k_debug_exp(const float* x, int n, int sim3, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) exp_sim3_dev(x + 7 * i, sim3 != 0, out + 12 * i);
}

int dspgn_debug_exp(int device, int sim3, const float* x, int n, float* out) {
  if (!x || !out || n < 1) return fail(DSPGN_E_ARG, "bad argument");
  CU(cudaSetDevice(device));
  DevBuf dx, dout;
  if (dx.reserve(28 * (size_t)n) || dout.reserve(48 * (size_t)n)) return fail(DSPGN_E_ALLOC, "cudaMalloc");
  CU(cudaMemcpy(dx.p, x, 28 * (size_t)n, cudaMemcpyHostToDevice));
  k_debug_exp<<<(n + 63) / 64, 64>>>(dx.as<float>(), n, sim3, dout.as<float>());
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(out, dout.p, 48 * (size_t)n, cudaMemcpyDeviceToHost);
  dx.release(); dout.release();
  if (e != cudaSuccess) return fail(DSPGN_E_CUDA, std::string("debug_exp: ") + cudaGetErrorString(e));
  return 0;
}

const float* dspgn_results_device(DspgnSolver* s) { return s ? s->d_results.as<float>() : nullptr; }

int dspgn_results(DspgnSolver* s, DspgnObjectOut* out) {
  if (!s || !out) return fail(DSPGN_E_ARG, "null argument");
  CU(cudaSetDevice(s->device));
  static_assert(sizeof(DspgnObjectOut) == 4 * DSPGN_RESULT_FLOATS, "result record layout");
  const size_t bytes = sizeof(DspgnObjectOut) * (size_t)s->n_obj;
  const bool mega = s->mega_ran;           // the queue counters (abort flag, band-row total) ride on the same copy + sync
  if (s->h_results.reserve(bytes + 512)) return fail(DSPGN_E_ALLOC, "cudaMallocHost");
  int* hq = reinterpret_cast<int*>(s->h_results.as<unsigned char>() + ((bytes + 63) / 64) * 64);
  CU(cudaMemcpyAsync(s->h_results.p, s->d_results.p, bytes, cudaMemcpyDeviceToHost, s->stream));
  if (mega) CU(cudaMemcpyAsync(hq, s->d_q_ctr.as<int>() + 64, 4 * 40, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  memcpy(out, s->h_results.p, bytes);
  if (mega) {
    s->mega_ran = false;
    if (s->band_rows_pending) {              // roofline accounting: what the reference decodes (loss.py:77-78, :143-144)
      s->ctr.rows_fwd_bwd += hq[80 - 64];    // band rows of all iterations
      { long long v; memcpy(&v, hq + (88 - 64), 8); s->ctr.rows_fwd_only += v; }   // V: ray samples inside the unit sphere, all iterations
    }
    s->band_rows_pending = false;
    if (hq[96 - 64]) return fail(DSPGN_E_CUDA, "persistent kernel: a work-queue wait timed out (aborted softly; results incomplete)");
  }
  if (s->timing) {
    float dec = 0.f;
    for (size_t i = 0; i + 1 < s->ev_used; i += 2) { float ms = 0.f; cudaEventElapsedTime(&ms, s->ev[i], s->ev[i + 1]); dec += ms; }
    s->ctr.decoder_ms = dec;
    float sv = 0.f;
    for (size_t i = 0; i + 1 < s->evs_used; i += 2) { float ms = 0.f; cudaEventElapsedTime(&ms, s->ev_solve[i], s->ev_solve[i + 1]); sv += ms; }
    s->ctr.solve_ms = sv;
  }
  float tot = 0.f;
  if (cudaEventElapsedTime(&tot, s->ev_run0, s->ev_run1) == cudaSuccess) s->ctr.total_ms = tot; else cudaGetLastError();
  return 0;
}

int dspgn_reconstruct_batch(DspgnSolver* s, int n_obj, const DspgnObjectIn* in, DspgnObjectOut* out) {
  if (!s || !in || !out || n_obj < 1) return fail(DSPGN_E_ARG, "bad argument");
  // any number of objects: resident batches of at most kMaxObjScan, one after the other
  for (int o0 = 0; o0 < n_obj; o0 += kMaxObjScan) {
    const int n = std::min(kMaxObjScan, n_obj - o0);
    if (int rc = dspgn_upload_batch(s, n, in + o0)) return rc;
    if (int rc = dspgn_run_batch(s, 0)) return rc;
    if (int rc = dspgn_results(s, out + o0)) return rc;
  }
  return 0;
}

int dspgn_estimate_pose_batch(DspgnSolver* s, int n_obj, const DspgnObjectIn* in, DspgnObjectOut* out) {
  if (!s || !in || !out || n_obj < 1) return fail(DSPGN_E_ARG, "bad argument");
  for (int o = 0; o < n_obj; ++o)
    if (!in[o].code || !(in[o].scale > 0.f)) return fail(DSPGN_E_ARG, "estimate_pose needs a code and a positive scale per object");
  for (int o0 = 0; o0 < n_obj; o0 += kMaxObjScan) {
    const int n = std::min(kMaxObjScan, n_obj - o0);
    if (int rc = dspgn_upload_batch(s, n, in + o0)) return rc;
    if (int rc = dspgn_run_batch(s, 1)) return rc;
    if (int rc = dspgn_results(s, out + o0)) return rc;
  }
  return 0;
}

int dspgn_decode_sdf(DspgnSolver* s, int class_id, const float* code, const float* x, int n, int x_rs, int x_cs,
                     float* sdf_out) {
  if (!s || !code || !x || !sdf_out || n < 1) return fail(DSPGN_E_ARG, "bad argument");
  const float I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  DspgnObjectIn in{};
  in.t_cam_obj = I4; in.t_rs = 4; in.t_cs = 1;
  in.pts = x; in.n_pts = n; in.pts_rs = x_rs; in.pts_cs = x_cs;
  in.rays = nullptr; in.n_rays = 0; in.depth = nullptr; in.n_depth = 0;
  in.code = code; in.scale = 1.f; in.class_id = class_id;
  if (int rc = upload_batch_impl(s, 1, &in, true)) return rc;      // forward only: no J^T J partial / band buffers
  if (s->d_sdf.reserve(4 * (size_t)n)) return fail(DSPGN_E_ALLOC, "cudaMalloc");
  s->ctr = DspgnCounters{};
  s->ev_used = 0;
  s->gdev = GatherDev{};
  if (int rc = launch_init(s, 0)) return rc;
  TermArgs a = base_term(s, MODE_PTSFWD);
  if (int rc = launch_term(s, a, n)) return rc;
  s->ctr.rows_fwd_only += n;
  CU(cudaMemcpyAsync(sdf_out, s->d_sdf.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

int dspgn_debug_system(DspgnSolver* s, int obj, int mode, float* H, float* b, float* dx, float* J_rows,
                       float* res_rows, float* losses) {
  return dspgn_debug_system_iter(s, obj, mode, 0, H, b, dx, J_rows, res_rows, losses);
}

int dspgn_debug_system_iter(DspgnSolver* s, int obj, int mode, int iter, float* H, float* b, float* dx, float* J_rows,
                            float* res_rows, float* losses) {
  if (!s || !H || !b || !dx) return fail(DSPGN_E_ARG, "null argument");
  if (obj < 0 || obj >= s->n_obj) return fail(DSPGN_E_ARG, "bad object index");
  if (iter < 0 || iter > 1000) return fail(DSPGN_E_ARG, "bad iteration index");
  CU(cudaSetDevice(s->device));
  const int pose_only = mode;
  const int P = pose_only ? 6 : 7 + s->cfg.code_len;
  const int npts = s->h_meta[obj].n_pts;
  DevBuf dJ;
  if (dJ.reserve(4 * ((size_t)npts * P + npts))) return fail(DSPGN_E_ALLOC, "cudaMalloc");
  float* dJp = dJ.as<float>();
  float* dres = dJp + (size_t)npts * P;
  s->ctr = DspgnCounters{};
  s->ev_used = 0;
  s->gdev = GatherDev{};
  int rc = launch_init(s, pose_only);
  for (int e = 0; e < iter && !rc; ++e) {            // advance the whole batch `iter` GN iterations (per-iteration schedule)
    rc = launch_terms(s, pose_only, nullptr, nullptr, -1, e);
    if (rc) break;
    SolveArgs v = base_solve(s, pose_only);
    v.last_iter = 0; v.iter_index = e;
    k_solve<<<s->n_obj, kSolveThreads, 0, s->stream>>>(v);
    if (cudaGetLastError() != cudaSuccess) rc = fail(DSPGN_E_CUDA, "k_solve launch failed");
  }
  if (!rc) rc = launch_terms(s, pose_only, dJp, dres, obj, iter);
  if (!rc) {
    SolveArgs v = base_solve(s, pose_only);
    float* d = s->d_dbg.as<float>();
    v.dbg_obj = obj; v.dbg_H = d; v.dbg_b = d + kPMax * kPMax; v.dbg_dx = v.dbg_b + kPMax; v.dbg_loss = v.dbg_dx + kPMax;
    v.last_iter = 0; v.iter_index = iter;
    cudaMemsetAsync(d, 0, 4 * ((size_t)kPMax * kPMax + 2 * kPMax + 8), s->stream);
    k_solve<<<s->n_obj, kSolveThreads, 0, s->stream>>>(v);
    if (cudaGetLastError() != cudaSuccess) rc = fail(DSPGN_E_CUDA, "k_solve launch failed");
    if (!rc) {
      cudaMemcpyAsync(H, v.dbg_H, 4 * (size_t)P * P, cudaMemcpyDeviceToHost, s->stream);
      cudaMemcpyAsync(b, v.dbg_b, 4 * (size_t)P, cudaMemcpyDeviceToHost, s->stream);
      cudaMemcpyAsync(dx, v.dbg_dx, 4 * (size_t)P, cudaMemcpyDeviceToHost, s->stream);
      if (losses) cudaMemcpyAsync(losses, v.dbg_loss, 16, cudaMemcpyDeviceToHost, s->stream);
      if (J_rows) cudaMemcpyAsync(J_rows, dJp, 4 * (size_t)npts * P, cudaMemcpyDeviceToHost, s->stream);
      if (res_rows) cudaMemcpyAsync(res_rows, dres, 4 * (size_t)npts, cudaMemcpyDeviceToHost, s->stream);
    }
  }
  cudaError_t e = cudaStreamSynchronize(s->stream);
  dJ.release();
  if (rc) return rc;
  if (e != cudaSuccess) return fail(DSPGN_E_CUDA, std::string("debug_system: ") + cudaGetErrorString(e));
  return 0;
}

int dspgn_debug_clocks(DspgnSolver* s, long long* out, int n) {
  // phase timeline of CTA 0's first tiles of the last SDF-term launch (enabled by env DSPGN_CLK=1 at solver creation)
  if (!s || !out) return fail(DSPGN_E_ARG, "null argument");
  if (!s->clk_on) return fail(DSPGN_E_ARG, "timeline not enabled (DSPGN_CLK)");
  const int have = kClkTiles * kTcMaxSteps * kClkSlots + 16;
  CU(cudaSetDevice(s->device));
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaMemcpy(out, s->d_clk.p, sizeof(long long) * (size_t)std::min(n, have), cudaMemcpyDeviceToHost));
  return 0;
}

int dspgn_debug_inputs(DspgnSolver* s, int obj, float* t_cam_obj, float* pts, float* rays) {
  if (!s || obj < 0 || obj >= s->n_obj) return fail(DSPGN_E_ARG, "bad argument");
  CU(cudaSetDevice(s->device));
  CU(cudaStreamSynchronize(s->stream));
  const ObjMeta& M = s->h_meta[obj];
  if (t_cam_obj) CU(cudaMemcpy(t_cam_obj, s->d_Tinit + 16 * (size_t)obj, 64, cudaMemcpyDeviceToHost));
  if (pts && M.n_pts) CU(cudaMemcpy(pts, s->d_pts + 3 * (size_t)M.pts_off, 12 * (size_t)M.n_pts, cudaMemcpyDeviceToHost));
  if (rays && M.n_rays) CU(cudaMemcpy(rays, s->d_rays + 3 * (size_t)M.ray_off, 12 * (size_t)M.n_rays, cudaMemcpyDeviceToHost));
  return 0;
}

int dspgn_debug_events(DspgnSolver* s, long long* out, int max_events) {
  // event log of the last persistent-kernel run (env DSPGN_CLK=1 at solver creation): returns the number of events,
  // out[2*i] = %globaltimer (ns), out[2*i+1] = kind<<56 | mode<<52 | sm<<40 | object<<24 | tile (or iteration)
  if (!s || !out || max_events < 1) return fail(DSPGN_E_ARG, "bad argument");
  if (!s->clk_on || !s->d_ev.p) return fail(DSPGN_E_ARG, "event log not enabled (DSPGN_CLK)");
  CU(cudaSetDevice(s->device));
  CU(cudaStreamSynchronize(s->stream));
  long long n = 0;
  CU(cudaMemcpy(&n, s->d_ev.p, 8, cudaMemcpyDeviceToHost));
  if (n > kEvCap) n = kEvCap;
  if (n > max_events) n = max_events;
  CU(cudaMemcpy(out, s->d_ev.as<long long>() + 1, 16 * (size_t)n, cudaMemcpyDeviceToHost));
  return (int)n;
}

int dspgn_tc_selftest(int device, int n_mma, int k_steps, const float* A, const float* B, float* D) {
  // D[128][n_mma] = A[128][16*k_steps] * B[n_mma][16*k_steps]^T through the tensor-core operand paths
  if (!A || !B || !D || n_mma < 16 || n_mma > 256 || (n_mma % 16) || k_steps < 1 || k_steps > 16) return fail(DSPGN_E_ARG, "bad selftest shape");
  CU(cudaSetDevice(device));
  if (int rc = tc_setup_kernels(g_err)) return rc;
  const int K = 16 * k_steps;
  std::vector<unsigned char> blob;
  tc_pack_images(blob, n_mma, k_steps, [&](int n, int kk) { return B[(size_t)n * K + kk]; });
  DevBuf dA, dB, dD;
  if (dA.reserve(4 * (size_t)128 * K) || dB.reserve(blob.size()) || dD.reserve(4 * (size_t)128 * n_mma)) return fail(DSPGN_E_ALLOC, "cudaMalloc");
  CU(cudaMemcpy(dA.p, A, 4 * (size_t)128 * K, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dB.p, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  k_tc_selftest<<<1, 128, 2 * kTcStageBytes + 1024>>>(dA.as<float>(), K, dB.as<unsigned char>(), n_mma, k_steps, dD.as<float>());
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(D, dD.p, 4 * (size_t)128 * n_mma, cudaMemcpyDeviceToHost);
  dA.release(); dB.release(); dD.release();
  if (e != cudaSuccess) return fail(DSPGN_E_CUDA, std::string("tc_selftest: ") + cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
