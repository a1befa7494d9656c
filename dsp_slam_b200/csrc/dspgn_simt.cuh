// fp32 SIMT decoder engine: fused  transform -> DeepSDF forward -> backward-to-input -> Jacobian rows
// -> per-tile partial sums of J^T J / J^T r  for one 64-row tile per CTA iteration.  This engine is the on-device ground truth
// (plain FFMA, fp32 accumulation in k order) against which the tcgen05 engine is checked.
//
// Restates: loss.py:22-43 (SDF term), loss.py:143-150 (band rows of the render term),
// loss_utils.py:51-103 (decode / input Jacobian), deep_sdf_decoder.py:75-110, optimizer.py:161-167.
#pragma once
#include "dspgn_common.cuh"

namespace dspgn {

constexpr int kTP = 64;          // rows (points) per tile
constexpr int kThreads = 256;
constexpr int kHid = 256;        // max layer width
constexpr int kKC = 16;          // reduction chunk staged in smem
constexpr int kMaxObjScan = 1024;

struct DecoderDev {
  int L, n_lin, latent_in, in0;            // in0 = L + 3
  int in_dim[DSPGN_MAX_LINEAR], out_dim[DSPGN_MAX_LINEAR];
  const float* Wf[DSPGN_MAX_LINEAR];       // forward, reduction-major  [in_pad16][256]:  Wf[i*256+j] = W[j][i]
  const float* Wb[DSPGN_MAX_LINEAR];       // backward, reduction-major [out_pad16][256]: Wb[i*256+j] = W[i][j]
  const float* bias[DSPGN_MAX_LINEAR];     // [256] zero padded
  const float* w_last;                     // [256] last layer row
  // optional variants (deep_sdf_decoder.py:41-47,58-63,87-102): SIMT engine only
  int cat_kind[DSPGN_MAX_LINEAR];          // input of layer k = [activations | 0: nothing, 1: decoder input, 2: xyz]
  const float* ln_gamma[DSPGN_MAX_LINEAR]; // LayerNorm after layer k (nullptr = none), [256] zero padded
  const float* ln_beta[DSPGN_MAX_LINEAR];
  int use_tanh, generic;                   // generic = any variant in use
  // tcgen05 engine images (dspgn_tc.cuh): pre-swizzled fp16 hi/lo weight chunks + step plan
  const unsigned char* tc_blob;
  TcPlan tc_plan;
};

enum { MODE_SDF = 0, MODE_BAND = 1, MODE_RAYFWD = 2, MODE_PTSFWD = 3 };

struct TermArgs {
  const ObjMeta* meta;
  ObjState* state;
  const DecoderDev* decs;
  int n_obj;
  int n_classes;
  int mode;
  // sources
  const float* pts;          // MODE_SDF: camera-frame points (xyz interleaved)
  const uint8_t* pt_active;  // optional inlier mask (pose-only, optimizer.py:76-78), may be null
  uint8_t* pt_active_out;    // when non-null: write |res| <= 0.05 per point (the cut taken after iteration index 4)
  int cut_iter;              // persistent mode: object iteration at which the cut is recorded (4), -1 = never
  const float* rays;         // MODE_RAYFWD
  const float* band_x;       // MODE_BAND: object-frame points xyz interleaved, per-sample capacity
  const float* band_s;       // de_ds per band row
  const float* band_r;       // residual per band row
  const int* band_m;         // rows per object
  float* sdf_out;            // MODE_RAYFWD: per sample sdf (+inf when outside the unit sphere)
  int* V_count;              // MODE_RAYFWD: valid samples per object
  float* part;               // per-tile partial sums of this term: [tile][kAccStride] (H upper | b | loss, rows)
  int* tile_base;            // [n_obj] first tile of each object in this launch (written by CTA 0)
  float huber_b;
  // persistent kernel with the render term: the band rows' partials / tile bases / Huber threshold (SDF ones above)
  float* part_r; const int* tile_base_r; float huber_b1;
  float* ln_scratch;         // SIMT engine, LayerNorm decoders: per-CTA [layer][256][kTP] normalised activations
  int D;
  int pose_only;             // 1: 6-D se3 Jacobian (no scale column)
  // debug dump of Jacobian rows (external order [pose | code]) for one object
  float* dbg_J; float* dbg_res; int dbg_obj; int dbg_P;
  long long* dbg_clk;         // optional phase timeline of CTA 0 (tensor-core engine)
};

// Device work queue of the persistent object-pipelined kernel (dspgn_tc.cuh).
// item = kind << 29 | object << 19 | tile   (kind: the tile's MODE_*; object < 1024; tile < 2^19; always >= 0)
// A queue slot is ONE word: 0 = not published yet, item + 1 = published (payload and flag in one store / one load).
constexpr int kItemKindShift = 29, kItemObjShift = 19, kItemTileMask = (1 << 19) - 1, kItemObjMask = 1023;
constexpr int kKindScan = 3;   // queue-only kind: per-ray scan of a 64-ray chunk (no GEMM steps); 0..2 = MODE_SDF / MODE_BAND / MODE_RAYFWD
__host__ __device__ __forceinline__ int make_item(int kind, int o, int tile) {
  return (kind << kItemKindShift) | (o << kItemObjShift) | tile;
}
// filler for reserved queue slots that turned out not to be needed (k_init reserves every object's iteration-0 slots
// from host-side upper bounds): consumers skip it.  Never a real item (a scan item's tile index is < 128); + 1 fits an int.
constexpr int kItemNop = 0x7ffffffe;
struct MegaArgs {
  int n_iters;               // GN iterations per object
  int q_cap;                 // total items that can ever be pushed
  int render;                // 1: joint run with the render term (ray-sample tiles -> per-ray scan -> band tiles)
  int* q_flag;               // one word per slot (no wrap-around): 0 = empty, item + 1 = published
  int* q_head; int* q_tail;  // consumer ticket counter / producer reservation counter
  int* pending;              // [n_obj] SDF + band tiles of the object's current iteration still running (+1 while the
                             //         render term has not been expanded into band tiles yet)
  int* ray_left;             // [n_obj] ray-sample tiles of the current iteration still running
  int* scan_left;            // [n_obj] scan items (64-ray chunks) of the current iteration still running
  int* seg_cnt; int* seg_prefix;   // band rows kept per 8-ray segment / their exclusive prefix per object (dspgn_solve.cuh)
  int* obj_iter;             // [n_obj] current iteration of each object
  int* done_objects;         // objects finished (last iteration or frozen)
  int* band_rows_total;      // sum of band rows over all objects and iterations (roofline accounting)
  unsigned long long* valid_rows_total;   // sum of V (ray samples inside the unit sphere) over all objects and iterations
  int vpre_exact;            // 1: the range pre-pass tests all D samples of every ray (debug / A-B switch)
  int* vpre;                 // per ray: (exclusive prefix of the valid-sample hulls << 7) | first valid sample, n_rays + 1
                             // entries per object at ray_off + o (dspgn_solve.cuh: valid_sample_ranges); nullptr = the
                             // forward-only tiles enumerate all n_rays * D samples
  int* abort_flag;           // set when a queue wait timed out: every CTA drains and exits (soft failure, never a trap)
  long long* ev; int ev_cap; // optional event log (env DSPGN_CLK): ev[0] = count, then {globaltimer ns, kind<<48|sm<<32|o<<20|tile}
};
// event kinds of the persistent kernel's debug log
enum { EV_TILE_BEGIN = 0, EV_TILE_END = 1, EV_SCAN_BEGIN = 2, EV_SCAN_END = 3, EV_SOLVE_BEGIN = 4, EV_SOLVE_END = 5, EV_POPPED = 6, EV_FIRST_MMA = 7 };
__device__ __forceinline__ void mega_event(const MegaArgs& q, int kind, int mode, int o, int tile) {
  if (q.ev == nullptr) return;
  const unsigned long long slot = atomicAdd(reinterpret_cast<unsigned long long*>(q.ev), 1ull);
  if ((long long)slot >= q.ev_cap) return;
  unsigned long long t; unsigned sm;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
  q.ev[1 + 2 * slot] = (long long)t;
  q.ev[2 + 2 * slot] = ((long long)kind << 56) | ((long long)mode << 52) | ((long long)sm << 40) | ((long long)o << 24) | (long long)tile;
}

// ---------------------------------------------------------------------------------------------
// tile scheduling shared by all decoder kernels: rows per object -> tiles, scanned per CTA
__device__ __forceinline__ int term_rows(const TermArgs& a, int o) {
  if (a.state[o].status != 0) return 0;
  if (a.mode == MODE_SDF || a.mode == MODE_PTSFWD) return a.meta[o].n_pts;
  if (a.mode == MODE_BAND) return a.band_m[o];
  return a.meta[o].n_rays * a.D;
}

// exclusive scan of tiles per object into s_prefix[0..n_obj]; returns total (all threads)
__device__ inline int build_tile_prefix(const TermArgs& a, int tile_rows, int* s_prefix, int* s_warp) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  int carry = 0;
  for (int base = 0; base < a.n_obj; base += blockDim.x) {
    int o = base + tid;
    int v = (o < a.n_obj) ? (term_rows(a, o) + tile_rows - 1) / tile_rows : 0;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_warp[w];
    int tot = 0;
    for (int w = 0; w < nw; ++w) tot += s_warp[w];
    if (o < a.n_obj) s_prefix[o] = carry + woff + x - v;
    carry += tot;
    __syncthreads();
  }
  if (tid == 0) s_prefix[a.n_obj] = carry;
  __syncthreads();
  if (blockIdx.x == 0 && a.tile_base != nullptr)
    for (int o = tid; o < a.n_obj; o += blockDim.x) a.tile_base[o] = s_prefix[o];
  return carry;
}

__device__ __forceinline__ int find_object(const int* s_prefix, int n_obj, int tile) {
  int lo = 0, hi = n_obj;      // largest o with prefix[o] <= tile
  while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_prefix[mid] <= tile) lo = mid; else hi = mid; }
  return lo;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// acc[jj][pp] = sum_{i<Kred} Wg[i*256 + 8*jg+jj] * in_s[i*kTP + 8*pg+pp]
__device__ __forceinline__ void gemm_rm(const float* __restrict__ Wg, int Kred, const float* __restrict__ in_s,
                                        float* __restrict__ wbuf, float (&acc)[8][8]) {
  const int tid = threadIdx.x, jg = tid >> 3, pg = tid & 7;
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
  const int nch = (Kred + kKC - 1) / kKC;
  // prefetch chunk 0
  {
    const float4* src = reinterpret_cast<const float4*>(Wg);
    float4* dst = reinterpret_cast<float4*>(wbuf);
#pragma unroll
    for (int q = 0; q < 4; ++q) cp_async16(dst + tid + q * kThreads, src + tid + q * kThreads);
    cp_async_commit();
  }
  for (int c = 0; c < nch; ++c) {
    cp_async_wait<0>();
    __syncthreads();                 // chunk c landed for all; everyone is done with chunk c-1's buffer
    if (c + 1 < nch) {
      const float4* src = reinterpret_cast<const float4*>(Wg + (size_t)(c + 1) * kKC * kHid);
      float4* dst = reinterpret_cast<float4*>(wbuf + ((c + 1) & 1) * kKC * kHid);
#pragma unroll
      for (int q = 0; q < 4; ++q) cp_async16(dst + tid + q * kThreads, src + tid + q * kThreads);
      cp_async_commit();
    }
    const float* wb = wbuf + (c & 1) * kKC * kHid + 8 * jg;
    const float* ib = in_s + (size_t)c * kKC * kTP + 8 * pg;
    const int kmax = min(kKC, Kred - c * kKC);
#pragma unroll 4
    for (int kk = 0; kk < kmax; ++kk) {
      float4 w0 = *reinterpret_cast<const float4*>(wb + kk * kHid);
      float4 w1 = *reinterpret_cast<const float4*>(wb + kk * kHid + 4);
      float4 a0 = *reinterpret_cast<const float4*>(ib + kk * kTP);
      float4 a1 = *reinterpret_cast<const float4*>(ib + kk * kTP + 4);
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float x[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(w[a], x[b], acc[a][b]);
    }
  }
  __syncthreads();                   // all reads of in_s / wbuf finished: caller may overwrite in_s
}

struct SimtSmem {
  float act[kHid * kTP];             // feature-major activations / gradients / J rows
  float inp[(kMaxCode + 4) * kTP];   // decoder input [z | x] rows
  float gin[(kMaxCode + 4) * kTP];   // d sdf / d(input) collected from the concat layers (latent_in / xyz_in_all)
  float lnst[DSPGN_MAX_LINEAR * kTP];       // LayerNorm: [layer][p] reciprocal std (the mean is not needed backward)
  float wbuf[2 * kKC * kHid];
  uint8_t mask[8 * kHid * 8];        // ReLU masks: [layer][feature][p/8] bit p%8
  float xo[3 * kTP];
  float yv[kTP], rr[kTP], rscale[kTP];
  float red[4 * kTP];
  int prefix[kMaxObjScan + 1];
  int warp_tmp[32];
};

// LayerNorm backward (deep_sdf_decoder.py:96-102 through autograd): S.act holds the gradient w.r.t. the LN OUTPUT of
// `layer` for its n features (already through the ReLU mask); turn it into the gradient w.r.t. the LN input:
//   g_x = rstd * (gamma g - mean_j(gamma g) - xhat * mean_j(gamma g xhat)).     All threads; caller has synchronised.
struct SimtSmem;
__device__ inline void simt_ln_backward(float* act, float* red, const float* rstd, const float* __restrict__ gamma,
                                        const float* __restrict__ xhat, int n) {
  const int tid = threadIdx.x;
  if (tid < kTP) {
    float s1 = 0.f, s2 = 0.f;
    for (int j = 0; j < n; ++j) {
      const float gg = act[j * kTP + tid] * gamma[j];
      s1 += gg;
      s2 = fmaf(gg, xhat[j * kTP + tid], s2);
    }
    red[tid] = s1 / (float)n;
    red[kTP + tid] = s2 / (float)n;
  }
  __syncthreads();
  for (int idx = tid; idx < n * kTP; idx += kThreads) {
    const int j = idx / kTP, p = idx - j * kTP;
    const float gg = act[idx] * gamma[j];
    act[idx] = rstd[p] * (gg - red[p] - xhat[idx] * red[kTP + p]);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 1) k_decoder_simt(TermArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SimtSmem& S = *reinterpret_cast<SimtSmem*>(smem_raw);
  const int tid = threadIdx.x, jg = tid >> 3, pg = tid & 7;
  const int total_tiles = build_tile_prefix(a, kTP, S.prefix, S.warp_tmp);

  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int o = find_object(S.prefix, a.n_obj, tile);
    const int row0 = (tile - S.prefix[o]) * kTP;
    const ObjMeta M = a.meta[o];
    const ObjState& st = a.state[o];
    const DecoderDev& dec = a.decs[M.class_id];
    const int L = dec.L, in0 = dec.in0, nl = dec.n_lin;
    const int nrows = min(kTP, term_rows(a, o) - row0);

    // ---- phase 0: points in the object frame, decoder input rows -----------------------------
    if (tid < kTP) {
      const int p = tid, r = row0 + p;
      float x = 0.f, y = 0.f, z = 0.f, sc = 0.f, res = 0.f;
      if (p < nrows) {
        if (a.mode == MODE_SDF || a.mode == MODE_PTSFWD) {
          const float* q = a.pts + 3 * (size_t)(M.pts_off + r);
          xform_point(st.T_oc, q[0], q[1], q[2], x, y, z);
          sc = (a.pt_active == nullptr || a.pt_active[M.pts_off + r]) ? 1.f : 0.f;
        } else if (a.mode == MODE_BAND) {
          const size_t s = (size_t)M.smp_off + r;
          x = a.band_x[3 * s]; y = a.band_x[3 * s + 1]; z = a.band_x[3 * s + 2];
          sc = a.band_s[s]; res = a.band_r[s];
        } else {
          const int ray = r / a.D, j = r - ray * a.D;
          const float* q = a.rays + 3 * (size_t)(M.ray_off + ray);
          const float d = lin_depth(st.dmin, st.dmax, st.dstep, j, a.D);
          xform_point(st.T_oc, __fmul_rn(q[0], d), __fmul_rn(q[1], d), __fmul_rn(q[2], d), x, y, z);
          sc = inside_unit_sphere(x, y, z) ? 1.f : 0.f;                // loss.py:68
        }
      }
      S.xo[p] = x; S.xo[kTP + p] = y; S.xo[2 * kTP + p] = z;
      S.rscale[p] = sc; S.rr[p] = res;
    }
    for (int idx = tid; idx < L * kTP; idx += kThreads) S.inp[idx] = st.z[idx / kTP];
    for (int idx = tid; idx < (kMaxCode + 4) * kTP; idx += kThreads) S.gin[idx] = 0.f;
    float* const xhat_all = (a.ln_scratch != nullptr) ? a.ln_scratch + (size_t)blockIdx.x * DSPGN_MAX_LINEAR * kHid * kTP : nullptr;
    __syncthreads();
    if (tid < 3 * kTP) S.inp[L * kTP + tid] = S.xo[tid];
    if (a.mode == MODE_RAYFWD) {
      // whole tile outside the unit sphere: nothing to decode
      int any = __syncthreads_or(tid < kTP && S.rscale[tid] != 0.f);
      if (!any) {
        if (tid < nrows) a.sdf_out[(size_t)M.smp_off + row0 + tid] = INFINITY;
        __syncthreads();
        continue;
      }
    } else {
      __syncthreads();
    }

    float acc[8][8];
    // ---- phase 1: forward -------------------------------------------------------------------
    for (int k = 0; k < nl - 1; ++k) {
      gemm_rm(dec.Wf[k], dec.in_dim[k], (k == 0) ? S.inp : S.act, S.wbuf, acc);
      const int nout = dec.out_dim[k];
      const float* bias = dec.bias[k];
      const float* lng = dec.ln_gamma[k];
      if (lng == nullptr) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int j = 8 * jg + jj;
          if (j < nout) {
            const float bj = bias[j];
            unsigned bits = 0;
            float v[8];
#pragma unroll
            for (int pp = 0; pp < 8; ++pp) {
              float t = acc[jj][pp] + bj;
              bits |= (t > 0.f ? 1u : 0u) << pp;
              v[pp] = fmaxf(t, 0.f);
            }
            S.mask[(k * kHid + j) * 8 + pg] = (uint8_t)bits;
            float4* dst = reinterpret_cast<float4*>(S.act + j * kTP + 8 * pg);
            dst[0] = make_float4(v[0], v[1], v[2], v[3]);
            dst[1] = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
      } else {
        // ---- LayerNorm between the layer and its ReLU (deep_sdf_decoder.py:96-103), eps = 1e-5, biased variance ----
        const float* lnb = dec.ln_beta[k];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int j = 8 * jg + jj;
          if (j < nout) {
            const float bj = bias[j];
            float4* dst = reinterpret_cast<float4*>(S.act + j * kTP + 8 * pg);
            dst[0] = make_float4(acc[jj][0] + bj, acc[jj][1] + bj, acc[jj][2] + bj, acc[jj][3] + bj);
            dst[1] = make_float4(acc[jj][4] + bj, acc[jj][5] + bj, acc[jj][6] + bj, acc[jj][7] + bj);
          }
        }
        __syncthreads();
        if (tid < kTP) {
          float m = 0.f;
          for (int j = 0; j < nout; ++j) m += S.act[j * kTP + tid];
          m /= (float)nout;
          float var = 0.f;
          for (int j = 0; j < nout; ++j) { const float d = S.act[j * kTP + tid] - m; var = fmaf(d, d, var); }
          var /= (float)nout;
          S.red[tid] = m;
          S.lnst[k * kTP + tid] = 1.0f / sqrtf(var + 1e-5f);
        }
        __syncthreads();
        float* xh_out = xhat_all + (size_t)k * kHid * kTP;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int j = 8 * jg + jj;
          if (j < nout) {
            const float gj = lng[j], bj = lnb[j];
            unsigned bits = 0;
            float v[8], xh[8];
#pragma unroll
            for (int pp = 0; pp < 8; ++pp) {
              const int p = 8 * pg + pp;
              xh[pp] = (S.act[j * kTP + p] - S.red[p]) * S.lnst[k * kTP + p];
              const float t = fmaf(xh[pp], gj, bj);
              bits |= (t > 0.f ? 1u : 0u) << pp;
              v[pp] = fmaxf(t, 0.f);
            }
            S.mask[(k * kHid + j) * 8 + pg] = (uint8_t)bits;
            float4* dst = reinterpret_cast<float4*>(S.act + j * kTP + 8 * pg);
            dst[0] = make_float4(v[0], v[1], v[2], v[3]);
            dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            float4* xd = reinterpret_cast<float4*>(xh_out + j * kTP + 8 * pg);      // for the backward pass (L2-resident scratch)
            xd[0] = make_float4(xh[0], xh[1], xh[2], xh[3]);
            xd[1] = make_float4(xh[4], xh[5], xh[6], xh[7]);
          }
        }
      }
      if (dec.cat_kind[k + 1] == 1)          // deep_sdf_decoder.py:87-88: x = cat[x, input]
        for (int idx = tid; idx < in0 * kTP; idx += kThreads) S.act[nout * kTP + idx] = S.inp[idx];
      else if (dec.cat_kind[k + 1] == 2)     // deep_sdf_decoder.py:89-90: x = cat[x, xyz]
        for (int idx = tid; idx < 3 * kTP; idx += kThreads) S.act[nout * kTP + idx] = S.inp[L * kTP + idx];
      __syncthreads();
    }
    {  // last layer (out = 1) + tanh
      const int kin = dec.in_dim[nl - 1];
      const int p = tid & (kTP - 1), part = tid >> 6;
      const int j0 = part * (kHid / 4), j1 = min(kin, j0 + kHid / 4);
      float s = 0.f;
      for (int j = j0; j < j1; ++j) s = fmaf(dec.w_last[j], S.act[j * kTP + p], s);
      S.red[part * kTP + p] = s;
      __syncthreads();
      if (tid < kTP) {
        float t = ((S.red[tid] + S.red[kTP + tid]) + S.red[2 * kTP + tid]) + S.red[3 * kTP + tid];
        t += dec.bias[nl - 1][0];
        float dfac = 1.f;
        if (dec.use_tanh) { t = tanhf(t); dfac = 1.f - t * t; }        // deep_sdf_decoder.py:93-94
        const float y = tanhf(t);                                       // :107-108
        S.yv[tid] = y;
        S.red[tid] = (1.f - y * y) * dfac;                              // d sdf / d(last layer output)
      }
      __syncthreads();
    }
    if (a.mode == MODE_RAYFWD || a.mode == MODE_PTSFWD) {
      int cnt = 0;
      if (tid < nrows) {
        const bool valid = S.rscale[tid] != 0.f;
        const size_t base = (a.mode == MODE_RAYFWD) ? (size_t)M.smp_off : (size_t)M.pts_off;
        a.sdf_out[base + row0 + tid] = valid ? S.yv[tid] : INFINITY;
        cnt = valid ? 1 : 0;
      }
      cnt = __syncthreads_count(cnt);
      if (tid == 0 && cnt && a.mode == MODE_RAYFWD) atomicAdd(a.V_count + o, cnt);
      continue;
    }

    // ---- phase 2: backward to the input ------------------------------------------------------
    {  // seed: g = (1 - y^2) W_last, masked by the last hidden ReLU
      const int kin = dec.in_dim[nl - 1];
      const int ckl = dec.cat_kind[nl - 1];                                // the last layer's input may carry a concat too
      const int ncl = kin - (ckl == 1 ? in0 : (ckl == 2 ? 3 : 0));
      for (int idx = tid; idx < kin * kTP; idx += kThreads) {
        const int j = idx / kTP, p = idx - j * kTP;
        const float g = S.red[p] * dec.w_last[j];
        if (j < ncl) {
          const unsigned bit = (S.mask[((nl - 2) * kHid + j) * 8 + (p >> 3)] >> (p & 7)) & 1u;
          S.act[idx] = bit ? g : 0.f;
        } else {
          S.gin[((ckl == 1 ? 0 : L) + (j - ncl)) * kTP + p] += g;
        }
      }
      __syncthreads();
      if (dec.ln_gamma[nl - 2] != nullptr)
        simt_ln_backward(S.act, S.red + kTP, S.lnst + (nl - 2) * kTP, dec.ln_gamma[nl - 2], xhat_all + (size_t)(nl - 2) * kHid * kTP, ncl);
    }
    for (int k = nl - 2; k >= 0; --k) {
      gemm_rm(dec.Wb[k], dec.out_dim[k], S.act, S.wbuf, acc);
      const int nin = dec.in_dim[k];
      const int ck = dec.cat_kind[k];
      const int ncont = (ck == 1) ? nin - in0 : (ck == 2 ? nin - 3 : nin);   // columns that continue down the chain
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = 8 * jg + jj;
        if (j >= nin) continue;
        float v[8];
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) v[pp] = acc[jj][pp];
        float* dst;
        if (j >= ncont) {                       // concat path (latent_in: whole input, xyz_in_all: xyz) -> d/d(input)
          dst = S.gin + ((ck == 1 ? 0 : L) + (j - ncont)) * kTP + 8 * pg;
#pragma unroll
          for (int pp = 0; pp < 8; ++pp) v[pp] += dst[pp];
        } else if (k > 0) {
          const unsigned bits = S.mask[((k - 1) * kHid + j) * 8 + pg];
#pragma unroll
          for (int pp = 0; pp < 8; ++pp) v[pp] = ((bits >> pp) & 1u) ? v[pp] : 0.f;
          dst = S.act + j * kTP + 8 * pg;
        } else {                                // k == 0: d/d(input) complete; scale rows (loss.py:145)
          const int jrow = (j < L) ? j : (kMaxCode + (j - L));
          dst = S.act + jrow * kTP + 8 * pg;
#pragma unroll
          for (int pp = 0; pp < 8; ++pp) {
            const float t = v[pp] + S.gin[j * kTP + 8 * pg + pp];
            v[pp] = t * S.rscale[8 * pg + pp];
          }
        }
        reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
      __syncthreads();
      if (k > 0 && dec.ln_gamma[k - 1] != nullptr)      // through the LayerNorm of layer k-1 (its ReLU mask is applied above)
        simt_ln_backward(S.act, S.red, S.lnst + (k - 1) * kTP, dec.ln_gamma[k - 1], xhat_all + (size_t)(k - 1) * kHid * kTP, ncont);
    }
    // ---- phase 3: Jacobian rows  J = [code (0..63) | pose (64..70) | 0] ---------------------------
    for (int idx = tid + L * kTP; idx < kMaxCode * kTP; idx += kThreads) S.act[idx] = 0.f;  // code_len < 64
    if (tid < kTP) {
      const int p = tid;
      const float gx = S.act[(kMaxCode + 0) * kTP + p], gy = S.act[(kMaxCode + 1) * kTP + p],
                  gz = S.act[(kMaxCode + 2) * kTP + p];
      const float x = S.xo[p], y = S.xo[kTP + p], z = S.xo[2 * kTP + p];
      // dsdf/dx . [I | -x^ | x]  (loss_utils.py:166-185)  ==  [g, x cross g, g.x]
      S.act[(kMaxCode + 3) * kTP + p] = y * gz - z * gy;
      S.act[(kMaxCode + 4) * kTP + p] = z * gx - x * gz;
      S.act[(kMaxCode + 5) * kTP + p] = x * gy - y * gx;
      S.act[(kMaxCode + 6) * kTP + p] = a.pose_only ? 0.f : (gx * x + gy * y + gz * z);
      S.act[(kMaxCode + 7) * kTP + p] = 0.f;
      float res = (a.mode == MODE_SDF) ? S.yv[p] : S.rr[p];
      const float sc = S.rscale[p];
      if (sc == 0.f && (a.mode == MODE_SDF || p >= nrows)) res = 0.f;
      if (a.pt_active_out != nullptr && a.mode == MODE_SDF && p < nrows)
        a.pt_active_out[M.pts_off + row0 + p] = (sc != 0.f && fabsf(res) <= 0.05f) ? 1 : 0;   // optimizer.py:76-78
      S.yv[p] = res;                                        // raw residual (debug dump)
      S.rr[p] = huber_weight(fabsf(res), a.huber_b) * res;  // loss_utils.py:250-265
    }
    __syncthreads();
    if (a.dbg_J != nullptr && o == a.dbg_obj && a.mode == MODE_SDF) {
      const int P = a.dbg_P, npose = a.pose_only ? 6 : 7;
      for (int idx = tid; idx < nrows * P; idx += kThreads) {
        const int p = idx / P, c = idx - p * P;
        const int ci = (c < npose) ? (kMaxCode + c) : (c - npose);
        a.dbg_J[(size_t)(row0 + p) * P + c] = S.act[ci * kTP + p];
      }
      if (tid < nrows) a.dbg_res[row0 + tid] = S.yv[tid];
    }
    // ---- phase 4: H += J^T J, b += J^T (rho r), loss += sum (rho r)^2  (optimizer.py:161-167) ----
    float* accp = a.part + (size_t)tile * kAccStride;
    if (tid < 171) {
      // upper-triangular 4x4 blocks of the 72x72 matrix: tid -> (bi <= bj)
      int bi = 0, rem = tid;
      while (rem >= 18 - bi) { rem -= 18 - bi; ++bi; }
      const int bj = bi + rem;
      float h[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) h[u][v] = 0.f;
      const float* ra = S.act + (4 * bi) * kTP;
      const float* rb = S.act + (4 * bj) * kTP;
      for (int p = 0; p < kTP; p += 4) {
        float4 A4[4], B4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          A4[u] = *reinterpret_cast<const float4*>(ra + u * kTP + p);
          B4[u] = *reinterpret_cast<const float4*>(rb + u * kTP + p);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v)
            h[u][v] += A4[u].x * B4[v].x + A4[u].y * B4[v].y + A4[u].z * B4[v].z + A4[u].w * B4[v].w;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int r = 4 * bi + u, c = 4 * bj + v;
          if (c >= r && c < kMaxCode + 7) accp[tri_index(r, c)] = h[u][v];
        }
    } else if (tid < 171 + kMaxCode + 7) {
      const int c = tid - 171;
      const float* rj = S.act + c * kTP;
      float s = 0.f;
      for (int p = 0; p < kTP; ++p) s = fmaf(rj[p], S.rr[p], s);
      accp[kAccB + c] = s;
    } else if (tid == 255) {
      float s = 0.f, n = 0.f;
      for (int p = 0; p < kTP; ++p) {
        s = fmaf(S.rr[p], S.rr[p], s);
        n += (a.mode == MODE_SDF) ? S.rscale[p] : (p < nrows ? 1.f : 0.f);
      }
      accp[kAccLoss] = s;
      accp[kAccLoss + 1] = n;
    }
    __syncthreads();
  }
}

}  // namespace dspgn
