// tcgen05 tensor-core decoder engine for B200 (sm_100a).
//
// One persistent CTA per SM walks 128-row tiles.  For each tile the whole DeepSDF forward chain, the
// backward-to-input chain and the Jacobian / J^T J reduction run without leaving the SM:
//
//   * accumulators AND the next layer's A operand live in TMEM (two 256-column regions, ping-pong):
//     the epilogue warps read a finished accumulator (tcgen05.ld), apply bias/ReLU (forward) or the saved
//     ReLU mask (backward), split every fp32 value into fp16 hi + fp16 lo and write the packed pairs back
//     IN PLACE (tcgen05.st); the next GEMM consumes them as a TMEM A operand (tcgen05.mma, A from TMEM);
//   * weights are pre-split (hi/lo fp16), pre-swizzled (128B swizzle, K-major) on the host into exactly
//     the shared-memory image the UMMA descriptor expects, and streamed from L2 through a 4 x 32 KB ring
//     with 1-D bulk copies (cp.async.bulk -> UBLKCP) signalling mbarriers;
//   * every product is formed as  A_hi*W_hi + A_lo*W_hi + A_hi*W_lo  (3 fp16 MMAs, fp32 accumulate):
//     ~2^-21 relative error per product, which keeps the Gauss-Newton iteration inside the fp32 noise
//     floor of the reference (>= 15 mantissa bits needed; bf16/tf32 single pass is not);
//   * only the hidden width x width layers are GEMM steps (14 per fwd+bwd tile of the 8 x 256 decoder): layer 0, with
//     its latent part folded into a per-object bias, is 3 FMAs per output while the first operand is built, and the
//     final Linear(width, 1) + tanh is a per-row dot product in the epilogue of the last hidden layer;
//   * J^T J / J^T r of the tile on the CUDA cores with packed fp32 FMAs (FFMA2), written as per-tile partials.
//
// Warp roles (320 threads): warps 0-3 / 4-7 = epilogue groups (thread = tile row; group g owns accumulator
// columns [128g, 128g+128)), warp 8 = MMA issuer (one elected lane), warp 9 = weight producer and, in the persistent
// kernels, the CTA's scheduler (pops the device work queue).
//
// Three schedules share this body (template SCHED): 0 = one launch per term and iteration (k_decoder_tc), 1 = persistent
// kernel with SDF tiles only (k_gn_persistent), 2 = persistent kernel with the render term: ray-sample tiles (only the run
// of samples inside the unit sphere of every ray: dspgn_solve.cuh, valid_sample_ranges), 64-ray scan items, band tiles,
// SDF tiles (k_gn_persistent_render).  The CTA that completes an object's last outstanding tile runs its solve
// (dspgn_solve.cuh), the next iteration's sample ranges, and queues the next iteration.
//
// Restates the same reference arithmetic as dspgn_simt.cuh (loss.py:22-43,143-150; loss_utils.py:51-103;
// deep_sdf_decoder.py:75-110; optimizer.py:161-167).
#pragma once
#include <cuda_fp16.h>
#include <string>
#include <vector>
#include "dspgn_simt.cuh"
#include "dspgn_solve.cuh"

namespace dspgn {

constexpr int kTcRows = 128;
constexpr int kTcThreads = 320;
constexpr int kTcEpiThreads = 256;
constexpr int kTcStages = 4;
constexpr int kTcStageBytes = 32768;

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tc_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::f16 (fp16 inputs, fp32 accumulate)
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// whole-warp variants: called converged by all 32 lanes, one elected lane issues.  Keeping the control flow
// warp-uniform lets the compiler hold descriptors in uniform registers instead of R2UR moves per operand.
__device__ __forceinline__ void tc_mma_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(smem_u32(bar)) : "memory");
}
#define DSPGN_R8(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
#define DSPGN_W8(v, o) "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7])

// 32 consecutive columns of this thread's TMEM lane (warp w%4 owns lanes 32(w%4)..+31)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : DSPGN_R8(v, 0), DSPGN_R8(v, 8), DSPGN_R8(v, 16), DSPGN_R8(v, 24)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : DSPGN_R8(v, 0), DSPGN_R8(v, 8)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ uint32_t tc_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      DSPGN_W8(v, 0), DSPGN_W8(v, 8), DSPGN_W8(v, 16), DSPGN_W8(v, 24)
      : "memory");
}

// debug timeline (CTA 0 only, when TermArgs.dbg_clk != nullptr): [tile][step][slot] = clock64
constexpr int kClkSlots = 8, kClkTiles = 4;
#define DSPGN_CLK(slot)                                                                                   \
  do {                                                                                                    \
    if (a.dbg_clk != nullptr && blockIdx.x == 0 && clk_tile < kClkTiles)                                  \
      a.dbg_clk[((size_t)clk_tile * kTcMaxSteps + s) * kClkSlots + (slot)] = clock64();                   \
  } while (0)

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// K-major, 128B-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm100):
// start>>4 [0,14) | LBO>>4 [16,30) = 1 | SBO>>4 [32,46) = 64 (8 rows x 128 B) | version [46,48) = 1 |
// layout [61,64) = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)64 << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=f16, K-major both, M=128
__host__ __device__ __forceinline__ uint32_t make_idesc(int n_mma) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n_mma >> 3) << 17) | ((128u >> 4) << 24);
}

// fp32 -> (fp16 hi, fp16 lo) for two consecutive K elements, packed low half = even element
__device__ __forceinline__ void split_pack(float a, float b, uint32_t& hi, uint32_t& lo) {
  __half2 h = __floats2half2_rn(a, b);
  float2 back = __half22float2(h);
  __half2 l = __floats2half2_rn(a - back.x, b - back.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up
// ------------------------------------------------------------------------------------------------
constexpr int kJpStride = 76;             // floats per row of the point-major Jacobian tile (72 + pad, 16B aligned)
struct TcSmemTail {
  float Jp[kTcRows * kJpStride];          // [row][72+4]: J row of each point; cols 0..66 double as latent_in skip gradient
  uint32_t maskw[8 * 8 * kTcRows];        // ReLU masks [layer][32-col word][row]
  float bias[9 * kHid];
  float wlast[kHid];
  float w0x[3 * kHid];                    // xyz rows of the layer-0 matrix (the latent rows are folded into ObjState.zb0)
  float zs[kMaxCode + 16];                // latent code of the tile's object (zero padded)
  float xr[3 * kTcRows];                  // object-frame point of every row
  float rr[kTcRows], rsc[kTcRows];
  int prefix[kMaxObjScan + 1];
  int warp_tmp[32];
  uint64_t w_full[kTcStages], w_empty[kTcStages];
  uint64_t acc_full[4];                   // accumulator quarter q (64 columns) of the current step is complete
  uint64_t a_ready[8];                    // 32-column unit u of the next A operand has been written
  uint32_t tmem_base;
  int cur_class;
  int fifo[4]; int fifo_pub; int epi_seq; int last_flag;   // persistent mode: CTA-local tile FIFO (scheduler = producer warp)
  TcPlan plans[DSPGN_MAX_CLASSES];        // step plans of every decoder class (read by all warp roles)
  // persistent mode: copies of the kernel arguments for the out-of-line solve step.  Passing references to the kernel
  // parameters themselves would make them address-taken: the compiler then parks all of them in local memory and the tile
  // loop reads its pointers with LDL instead of from the constant bank.
  MegaArgs ctx_q; SolveArgs ctx_sv; int ctx_D; const float* ctx_rays;
  int push_base, push_nF, push_nS;        // cooperative publication of an object's next-iteration tiles
  float ost[16]; int ost_rows;            // the tile's object: T_oc[12], dmin, dmax, dstep, dfar; rows of its term (counter)
};
constexpr size_t kTcSmemBytes = 1024 + (size_t)kTcStages * kTcStageBytes + sizeof(TcSmemTail);

__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      DSPGN_W8(v, 0), DSPGN_W8(v, 8)
      : "memory");
}

// A-operand layout inside a 256-column TMEM region: 32-column units, unit u holds K elements [32u, 32u+32):
// fp16 "hi" halves packed in columns [32u, 32u+16), "lo" halves in [32u+16, 32u+32).
// K-step t (16 elements) -> hi at column 32*(t>>1) + 8*(t&1), lo 16 columns further.
__device__ __forceinline__ uint32_t a_col_hi(int t) { return (uint32_t)(32 * (t >> 1) + 8 * (t & 1)); }

__device__ __forceinline__ void store_a_unit(uint32_t taddr, const float (&t)[32]) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int p = 0; p < 16; ++p) split_pack(t[2 * p], t[2 * p + 1], hi[p], lo[p]);
  tc_st16(taddr, hi);
  tc_st16(taddr + 16, lo);
}

// signal "unit u of the next A operand is in TMEM" (or simply "done with this unit")
// (one arrival per warp: 4 per unit instead of 128 -- the 128 individual arrivals on one mbarrier serialised for
//  several hundred cycles on the path that decides when the next layer's first MMA can issue)
__device__ __forceinline__ void unit_done(uint64_t* bar) {
  tc_wait_st();
  tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}

struct TileRef { int o, row0, slot, mode, tile; };

// pop one work item for this CTA (persistent mode); -1 = no more work anywhere
// A wait that outlives kMegaTimeoutNs (wall clock, so it also holds under compute-sanitizer / a debugger) raises the
// abort flag: every CTA drains and exits, the host reports DSPGN_E_CUDA.  No __trap: a trap would poison the CUDA
// context of the whole process (the detectors on the Tracking thread live in it too).
constexpr unsigned long long kMegaTimeoutNs = 30ull * 1000ull * 1000ull * 1000ull;
__device__ inline int mega_pop(const MegaArgs& q, int n_obj) {
  for (;;) {
    const int t = atomicAdd(q.q_head, 1);
    if (t >= q.q_cap) return -1;
    unsigned long long t0 = 0;
    int item = kItemNop;
    for (unsigned spins = 0;; ++spins) {
      const int v = ldv(q.q_flag + t);
      if (v != 0) { __threadfence(); item = v - 1; break; }
      if (ldv(q.done_objects) >= n_obj || ldv(q.abort_flag) != 0) return -1;
      __nanosleep(256);
      if ((spins & 1023u) == 1023u) {
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > kMegaTimeoutNs) { atomicExch(q.abort_flag, 1); return -1; }
      }
    }
    if (item != kItemNop) return item;       // filler of a reserved slot that was not needed: take the next ticket
  }
}

// tile number `seq` of this CTA: static round-robin over the launch's tiles, or the CTA-local FIFO
// SCHED: 0 = one launch per term (static tiles), 1 = persistent kernel, SDF tiles only (SDF-only joint runs, pose-only
// runs: the tile kind is a compile-time constant), 2 = persistent kernel with the render term (all item kinds)
template <int SCHED>
__device__ __forceinline__ bool tile_at(const TermArgs& a, TcSmemTail& S, int seq, int total_tiles, TileRef& t) {
  constexpr bool MEGA = SCHED != 0;
  if (!MEGA) {
    const int tile = blockIdx.x + seq * gridDim.x;
    if (tile >= total_tiles) return false;
    t.o = find_object(S.prefix, a.n_obj, tile);
    t.tile = tile - S.prefix[t.o];
    t.row0 = t.tile * kTcRows;
    t.slot = tile;
    t.mode = a.mode;
    return true;
  } else {
    volatile int* pub = &S.fifo_pub;
    while (*pub <= seq) __nanosleep(64);           // filled by this CTA's scheduler lane, which always terminates (mega_pop)
    const int item = reinterpret_cast<volatile int*>(S.fifo)[seq & 3];
    if (item < 0) return false;
    t.o = (item >> kItemObjShift) & kItemObjMask;
    t.mode = (SCHED == 1) ? MODE_SDF : (item >> kItemKindShift);
    const int j = item & kItemTileMask;
    t.tile = j;
    t.row0 = j * kTcRows;
    t.slot = (t.mode == MODE_BAND) ? a.tile_base_r[t.o] + j : (t.mode == MODE_SDF ? a.tile_base[t.o] + j : 0);
    return true;
  }
}

// rows of the term a tile belongs to (persistent kernel: the tile's own kind, counters written by other CTAs)
__device__ __forceinline__ int mega_rows(const TermArgs& a, const MegaArgs& q, const ObjMeta& M, int o, int mode) {
  if (mode == MODE_SDF) return M.n_pts;
  if (mode == MODE_BAND) return ldv(a.band_m + o);
  if (q.vpre != nullptr) return ldv(q.vpre + vpre_base(M, o) + M.n_rays) >> 7;   // valid-sample hulls only
  return M.n_rays * a.D;
}

// publish `n` queue items (kind, object, tile 0..n-1): reserve slots, fence (everything the items depend on, incl. the
// counters updated just before the call), one word per slot.  One thread.
__device__ inline void mega_push(const MegaArgs& q, int kind, int o, int n) {
  if (n <= 0) return;
  const int base = atomicAdd(q.q_tail, n);
  __threadfence();
  for (int j = 0; j < n; ++j) *reinterpret_cast<volatile int*>(q.q_flag + base + j) = make_item(kind, o, j) + 1;
}

// all terms of the object's current iteration are in: solve, update, queue the next iteration (or finish).  Called by
// the 256 epilogue threads of the CTA that completed the object's last outstanding tile.
template <bool MEGA>
__device__ __noinline__ void mega_solve_and_advance(TcSmemTail& S, int o, int tid) {
  const MegaArgs& q = S.ctx_q;
  const SolveArgs& sv = S.ctx_sv;
  __threadfence();
  SolveSmem& SM = *reinterpret_cast<SolveSmem*>(S.Jp);
  const int it = ldv(q.obj_iter + o);
  if (tid == 0) mega_event(q, EV_SOLVE_BEGIN, 0, o, it);
  const int fin = solve_object<true>(sv, o, tid, SM, it + 1 >= q.n_iters);
  epi_bar_sync();
  // ---- the next iteration's ray samples: only the run of samples inside the unit sphere of every ray (new pose and
  // depth range, written by the solve above).  `fin` is the same in every thread (shared-memory flags).
  int vh = -1;
  if (!fin && q.render && q.vpre != nullptr) {
    const ObjMeta M = sv.meta[o];
    if (M.n_rays > 0) vh = valid_sample_ranges<true>(M, sv.state[o], S.ctx_rays, S.ctx_D, q.vpre + vpre_base(M, o), tid, kTcEpiThreads, S.warp_tmp, q.vpre_exact != 0);
  }
  // ---- publish: finished, or the tiles of the next iteration.  All 256 threads write the queue slots (one thread
  // pushing 176 ray tiles + their flags one by one took ~3 us on the single-object critical path).
  if (tid == 0) {
    mega_event(q, EV_SOLVE_END, 0, o, it);
    int base = -1;
    if (fin) {
      __threadfence();                       // the result record before the object counts as done
      atomicAdd(q.done_objects, 1);
    } else {
      // (no fence in this branch: q_tail only reserves slots; state and counters are fenced below, before any slot is published)
      const ObjMeta M = sv.meta[o];
      const int ntS = (M.n_pts + kTcRows - 1) / kTcRows;
      const int ntF = q.render ? ((vh >= 0 ? vh : M.n_rays * S.ctx_D) + kTcRows - 1) / kTcRows : 0;
      *reinterpret_cast<volatile int*>(q.obj_iter + o) = it + 1;
      *reinterpret_cast<volatile int*>(q.pending + o) = ntS + (ntF > 0 ? 1 : 0);
      *reinterpret_cast<volatile int*>(q.ray_left + o) = ntF;
      base = atomicAdd(q.q_tail, ntF + ntS);  // the long chain (rays -> scan -> band -> solve) first, then the SDF tiles
      S.push_nF = ntF; S.push_nS = ntS;
    }
    S.push_base = base;
  }
  epi_bar_sync();
  const int base = S.push_base;
  if (base >= 0) {
    const int nF = S.push_nF, n = nF + S.push_nS;
    __threadfence();                          // this thread's share of the ray range words (thread 0: state, counters)
    epi_bar_sync();                           // ... of every thread, before the first slot is published
    for (int j = tid; j < n; j += kTcEpiThreads)
      *reinterpret_cast<volatile int*>(q.q_flag + base + j) = ((j < nF) ? make_item(MODE_RAYFWD, o, j) : make_item(MODE_SDF, o, j - nF)) + 1;
  }
}

template <int SCHED>
__device__ __forceinline__ void tc_body(const TermArgs& a, const MegaArgs& q, const SolveArgs& sv, const ScanArgs& sc_args) {
  constexpr bool MEGA = SCHED != 0;
  constexpr bool RENDER = SCHED == 2;
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* ring = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer
  TcSmemTail& S = *reinterpret_cast<TcSmemTail*>(ring + (size_t)kTcStages * kTcStageBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int total_tiles = MEGA ? 0 : build_tile_prefix(a, kTcRows, S.prefix, S.warp_tmp);
  {
    const int nwords = a.n_classes * (int)(sizeof(TcPlan) / 4);
    for (int i = tid; i < nwords; i += kTcThreads) {
      const int c = i / (int)(sizeof(TcPlan) / 4), w = i % (int)(sizeof(TcPlan) / 4);
      reinterpret_cast<int*>(&S.plans[c])[w] = reinterpret_cast<const int*>(&a.decs[c].tc_plan)[w];
    }
  }
  if (tid == 0) {
    for (int i = 0; i < kTcStages; ++i) { mbar_init(&S.w_full[i], 1); mbar_init(&S.w_empty[i], 1); }
    mbar_init(&S.acc_full[0], 1);
    for (int i = 0; i < 8; ++i) mbar_init(&S.a_ready[i], 4);
    S.cur_class = -1;
    S.fifo_pub = 0; S.epi_seq = 0; S.last_flag = 0;
    if (MEGA) { S.ctx_q = q; S.ctx_sv = sv; S.ctx_D = a.D; S.ctx_rays = a.rays; }
    fence_barrier_init();
  }
  if (warp == 8) tc_alloc(&S.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;

  if (warp == 9) {
    // ===================== weight producer ======================================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int seq = 0;; ++seq) {
        if (MEGA) {
          // scheduler: fetch this CTA's next tile into the local FIFO (at most 3 entries ahead of the epilogue)
          volatile int* es = &S.epi_seq;
          while (seq - *es >= 3) __nanosleep(64);   // the epilogue warps always make progress (bounded tile work)
          const int item = mega_pop(q, a.n_obj);
          if (item >= 0) mega_event(q, EV_POPPED, item >> kItemKindShift, (item >> kItemObjShift) & kItemObjMask, item & kItemTileMask);
          reinterpret_cast<volatile int*>(S.fifo)[seq & 3] = item;
          __threadfence_block();
          *reinterpret_cast<volatile int*>(&S.fifo_pub) = seq + 1;
        }
        TileRef tr;
        if (!tile_at<SCHED>(a, S, seq, total_tiles, tr)) break;
        const int o = tr.o;
        const int cls = a.meta[o].class_id;
        const TcPlan& plan = S.plans[cls];
        const unsigned char* blob = a.decs[cls].tc_blob;
        const bool fwd_only = (tr.mode == MODE_RAYFWD || tr.mode == MODE_PTSFWD);
        const int ns = (RENDER && tr.mode == kKindScan) ? 0 : (fwd_only ? plan.n_fwd : plan.n_steps);
        for (int s = 0; s < ns; ++s) {
          const TcStep st = plan.step[s];
          const uint32_t img = (uint32_t)st.n_mma * 128u;
          const int nch = (st.k_steps + 3) >> 2;
          const unsigned char* src = blob + st.w_off;
          for (int c = 0; c < 2 * nch; ++c) {           // hi image, lo image, hi, lo, ...
            mbar_wait(&S.w_empty[stage], phase ^ 1);
            mbar_expect_tx(&S.w_full[stage], img);
            bulk_g2s(ring + (size_t)stage * kTcStageBytes, src + (size_t)c * img, img, &S.w_full[stage]);
            if (++stage == kTcStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer ===========================================================
    // Full-width MMAs (N = the layer's padded output width): with the A operand in TMEM an MMA costs >= ~110
    // cycles whatever its N (measured), so N is never split.  Overlap with the epilogue comes from the
    // per-unit a_ready barriers: K chunk c of a step only needs operand units 2c and 2c+1.
    uint32_t stage = 0, phase = 0, ar_phase = 0;
    int clk_tile = -1;
    for (int seq = 0;; ++seq) {
      ++clk_tile;
      TileRef tr;
      if (!tile_at<SCHED>(a, S, seq, total_tiles, tr)) break;
      const int o = tr.o;
      const TcPlan& plan = S.plans[a.meta[o].class_id];
      const bool fwd_only = (tr.mode == MODE_RAYFWD || tr.mode == MODE_PTSFWD);
      const int ns = (RENDER && tr.mode == kKindScan) ? 0 : (fwd_only ? plan.n_fwd : plan.n_steps);
      for (int s = 0; s < ns; ++s) {
        const TcStep st = plan.step[s];
        const uint32_t d_t = tmem + (uint32_t)st.d_reg * 256u;
        const uint32_t a_t = tmem + (uint32_t)st.a_reg * 256u;
        const int nch = (st.k_steps + 3) >> 2;
        const uint32_t idesc = make_idesc(st.n_mma);
        for (int c = 0; c < nch; ++c) {
          const int nk = min(4, st.k_steps - 4 * c);
          const bool last = (c == nch - 1);
          // both 32-column units of this chunk must have been written by the epilogue warps; before the
          // LAST chunk (whose commit releases the accumulator) every unit of the previous step must be
          // finished, so no barrier phase can run ahead of a slow epilogue group
          mbar_wait(&S.a_ready[2 * c], ar_phase);
          mbar_wait(&S.a_ready[2 * c + 1], ar_phase);
          if (last)
            for (int u = 2 * c + 2; u < 8; ++u) mbar_wait(&S.a_ready[u], ar_phase);
          if (MEGA && c == 0 && s == 0 && lane == 0) mega_event(q, EV_FIRST_MMA, tr.mode, tr.o, tr.tile);
          if (c == 0 && lane == 0) {
            DSPGN_CLK(4);
            if (a.dbg_clk != nullptr && blockIdx.x == 0 && clk_tile < kClkTiles) {
              unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
              a.dbg_clk[((size_t)clk_tile * kTcMaxSteps + s) * kClkSlots + 6] = (long long)gt;
            }
          }
          // ---- W_hi image: A_hi*W_hi + A_lo*W_hi
          mbar_wait(&S.w_full[stage], phase);
          tc_fence_after();
          {
            const uint32_t b0 = smem_u32(ring + (size_t)stage * kTcStageBytes);
            for (int k = 0; k < nk; ++k) {
              const uint64_t bd = make_b_desc(b0 + 32u * k);
              const uint32_t ah = a_t + 64u * c + a_col_hi(k);
              tc_mma_ts_elect(d_t, ah, bd, idesc, (c | k) ? 1u : 0u);
              tc_mma_ts_elect(d_t, ah + 16u, bd, idesc, 1u);
            }
            tc_commit_elect(&S.w_empty[stage]);
          }
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
          // ---- W_lo image: A_hi*W_lo
          mbar_wait(&S.w_full[stage], phase);
          tc_fence_after();
          {
            const uint32_t b0 = smem_u32(ring + (size_t)stage * kTcStageBytes);
            for (int k = 0; k < nk; ++k) tc_mma_ts_elect(d_t, a_t + 64u * c + a_col_hi(k), make_b_desc(b0 + 32u * k), idesc, 1u);
            tc_commit_elect(&S.w_empty[stage]);
            if (last) { tc_commit_elect(&S.acc_full[0]); if (lane == 0) DSPGN_CLK(5); }
          }
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
        ar_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue groups ======================================================
    // group g (warps 4g..4g+3) converts operand units g, g+2, g+4, g+6 (32 columns each), in that order, so
    // that the two units of K chunk c are produced concurrently by the two groups; thread = tile row = TMEM lane.
    const int grp = warp >> 2;
    const int r = tid & 127;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t acc_phase = 0;
    int clk_tile = -1;
    for (int seq = 0;; ++seq) {
      ++clk_tile;
      TileRef tr;
      if (!tile_at<SCHED>(a, S, seq, total_tiles, tr)) break;
      if (MEGA && tid == 0) { *reinterpret_cast<volatile int*>(&S.epi_seq) = seq + 1; mega_event(q, EV_TILE_BEGIN, tr.mode, tr.o, tr.tile); }
      if (RENDER && tr.mode == kKindScan) {
        // ---- scan item: occupancy scan / rendered depth / band rows of 64 rays (loss.py:84-141); no GEMM steps ----------
        const int o = tr.o;
        scan_chunk(sc_args, q.seg_cnt, o, tr.tile, tid);
        __threadfence();
        epi_bar_sync();
        if (tid == 0) {
          mega_event(q, EV_TILE_END, tr.mode, o, tr.tile);
          *reinterpret_cast<volatile int*>(&S.last_flag) = (atomicSub(q.scan_left + o, 1) == 1) ? 1 : 0;
        }
        epi_bar_sync();
        int act = *reinterpret_cast<volatile int*>(&S.last_flag);
        if (act == 1) {
          // last chunk of the object: segment prefix -> band row count -> band tiles
          __threadfence();
          scan_prefix(sc_args, q.seg_cnt, q.seg_prefix, o, tid, S.warp_tmp);
          epi_bar_sync();
          if (tid == 0) {
            __threadfence();                         // prefix / band_m / band rows before the band tiles are published
            atomicAdd(q.valid_rows_total, (unsigned long long)ldv(sc_args.V_count + o));   // V of this iteration is complete (roofline accounting)
            const int m = ldv(sc_args.band_m + o);
            const int ntB = (m + kTcRows - 1) / kTcRows;
            atomicAdd(q.band_rows_total, m);
            // the render term's placeholder in `pending` becomes its ntB band tiles BEFORE they can be popped
            const int left = atomicAdd(q.pending + o, ntB - 1) + ntB - 1;
            mega_push(q, MODE_BAND, o, ntB);
            *reinterpret_cast<volatile int*>(&S.last_flag) = (left == 0) ? 2 : 0;
          }
          epi_bar_sync();
          act = *reinterpret_cast<volatile int*>(&S.last_flag);
        } else act = 0;
        if (act == 2) mega_solve_and_advance<MEGA>(S, o, tid);
        continue;
      }
      const int o = tr.o, row0 = tr.row0, tile = tr.slot, mode = tr.mode;
      const ObjMeta M = a.meta[o];
      const ObjState& ost = a.state[o];
      const DecoderDev& dec = a.decs[M.class_id];
      const TcPlan& plan = S.plans[M.class_id];
      const int L = dec.L, in0 = dec.in0, n_lin = dec.n_lin;
      const bool has_skip = dec.latent_in >= 0;
      const bool fwd_only = (mode == MODE_RAYFWD || mode == MODE_PTSFWD);
      const int ns = fwd_only ? plan.n_fwd : plan.n_steps;
      const float huber_b = (RENDER && mode == MODE_BAND) ? a.huber_b1 : a.huber_b;
      float* const part = (RENDER && mode == MODE_BAND) ? a.part_r : a.part;
      // ---- prologue, phase A: everything that comes from global memory, then ONE barrier ------------------------------
      // The pose / depth range of this object may have been rewritten by another CTA's solve: read (cache-bypassing) once
      // per tile by 16 threads and shared through smem.  As 12 + 3 loads in every thread they were ~130 requests per tile
      // for the same two L2 lines -- and on few-object batches every SM asks for them at the same moment.
      if (tid < 12) S.ost[tid] = ldv(&ost.T_oc[tid]);
      else if (tid < 16) S.ost[tid] = ldv(&ost.dmin + (tid - 12));          // dmin, dmax, dstep, dfar
      const bool pts_mode = (mode == MODE_SDF || mode == MODE_PTSFWD);
      if (MEGA && !pts_mode && tid == 16) S.ost_rows = mega_rows(a, q, M, o, mode);   // band / ray-sample rows: a counter

      // per-class constants in smem (bias, last row, xyz rows of layer 0), the tile's latent code
      if (S.cur_class != M.class_id) {
        for (int i = tid; i < n_lin * kHid; i += kTcEpiThreads) S.bias[i] = dec.bias[i / kHid][i % kHid];
        for (int i = tid; i < kHid; i += kTcEpiThreads) S.wlast[i] = dec.w_last[i];
        const float* __restrict__ w0g = dec.Wf[0] + (size_t)L * kHid;   // rows L..L+2 of the reduction-major layer-0 matrix
        for (int i = tid; i < 3 * kHid; i += kTcEpiThreads) S.w0x[i] = w0g[i];
      }
      if (tid < kMaxCode + 16) S.zs[tid] = (tid < L) ? ldv(&ost.z[tid]) : 0.f;
      // layer 0 with the latent part folded into a per-object bias (ObjState.zb0, refreshed by k_init / the solve step);
      // written after the per-class reload above, read after the barriers below
      S.bias[tid] = ldv(&ost.zb0[tid]);
      // pose-only inlier cut (optimizer.py:76-78): recorded while iteration `cut_iter` runs, applied afterwards
      const uint8_t* mask_in = a.pt_active;
      uint8_t* mask_out = a.pt_active_out;
      if (MEGA && a.cut_iter >= 0) {
        const int it_now = ldv(q.obj_iter + o);
        mask_in = (it_now > a.cut_iter) ? a.pt_active_out : nullptr;
        mask_out = (it_now == a.cut_iter) ? a.pt_active_out : nullptr;
      }
      const int* segp = nullptr;
      int nseg = 0;
      const bool compact = RENDER && mode == MODE_RAYFWD && q.vpre != nullptr;
      if (compact) {
        // the object's per-ray range words (<= 8193 ints) in the idle J tile: row -> (ray, sample) by binary search below
        int* sp = reinterpret_cast<int*>(S.Jp);
        const int* gp = q.vpre + vpre_base(M, o);
        for (int i = tid; i <= M.n_rays; i += kTcEpiThreads) sp[i] = __ldcg(gp + i);
        segp = sp;
      }
      if (RENDER && mode == MODE_BAND) {
        // band rows live compacted per ray segment: stage the object's segment prefix in the idle J tile
        nseg = (M.n_rays + kSegRays - 1) / kSegRays;
        int* sp = reinterpret_cast<int*>(S.Jp);
        const int* gp = q.seg_prefix + seg_base(M, o);
        for (int i = tid; i <= nseg; i += kTcEpiThreads) sp[i] = __ldcg(gp + i);
        segp = sp;
      }
      // surface points do not depend on anything above: fetch them before the barrier
      int nrows = 0;
      float p0 = 0.f, p1 = 0.f, p2 = 0.f, sc = 0.f;
      if (pts_mode) {
        nrows = min(kTcRows, (MEGA ? M.n_pts : term_rows(a, o)) - row0);
        if (r < nrows) {
          const float* pq = a.pts + 3 * (size_t)(M.pts_off + row0 + r);
          p0 = pq[0]; p1 = pq[1]; p2 = pq[2];
          sc = (mask_in == nullptr || ldv(mask_in + M.pts_off + row0 + r)) ? 1.f : 0.f;
        }
      }
      epi_bar_sync();      // (per-iteration schedule: Jp / rr of the previous tile are not written before the barrier further down;
                           //  persistent schedule: the previous tile ended with a barrier, its J tile is dead)
      // ---- phase B: this row's point in the object frame ----------------------------------------------------------------
      float Toc[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) Toc[i] = S.ost[i];
      if (!pts_mode) nrows = min(kTcRows, (MEGA ? S.ost_rows : term_rows(a, o)) - row0);
      float x0 = 0.f, x1 = 0.f, x2 = 0.f, res_in = 0.f;
      if (r < nrows) {
        const int rr_ = row0 + r;
        if (pts_mode) {
          xform_point(Toc, p0, p1, p2, x0, x1, x2);
        } else if (mode == MODE_BAND) {
          // band rows were written by the CTAs that ran this object's scan: L2 is the point of coherence
          size_t sidx = (size_t)M.smp_off + rr_;
          if (RENDER) {
            int lo = 0, hi = nseg;                     // largest segment with prefix <= row
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (segp[mid] <= rr_) lo = mid; else hi = mid; }
            sidx = (size_t)M.smp_off + (size_t)lo * kSegRays * a.D + (size_t)(rr_ - segp[lo]);
          }
          x0 = __ldcg(a.band_x + 3 * sidx); x1 = __ldcg(a.band_x + 3 * sidx + 1); x2 = __ldcg(a.band_x + 3 * sidx + 2);
          sc = __ldcg(a.band_s + sidx); res_in = __ldcg(a.band_r + sidx);
        } else {
          int ray = rr_ / a.D, j = rr_ - ray * a.D;
          if (compact) {
            int lo = 0, hi = M.n_rays;                 // largest ray whose hull starts at or before this row
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((segp[mid] >> 7) <= rr_) lo = mid; else hi = mid; }
            ray = lo; j = (segp[lo] & 127) + (rr_ - (segp[lo] >> 7));
          }
          const float* rq = a.rays + 3 * (size_t)(M.ray_off + ray);
          const float d = lin_depth(S.ost[12], S.ost[13], S.ost[14], j, a.D);
          xform_point(Toc, __fmul_rn(rq[0], d), __fmul_rn(rq[1], d), __fmul_rn(rq[2], d), x0, x1, x2);
          sc = inside_unit_sphere(x0, x1, x2) ? 1.f : 0.f;            // loss.py:68
        }
      }
      if (grp == 0) { S.xr[r] = x0; S.xr[kTcRows + r] = x1; S.xr[2 * kTcRows + r] = x2; }
      epi_bar_sync();                                // zs / xr / bias visible; previous tile fully drained
      if (tid == 0) S.cur_class = M.class_id;

      // decoder input element i of this row: [z | x | 0...]
      auto inp = [&](int i) -> float {
        const int j = i - L;
        return (j < 0) ? S.zs[i] : ((unsigned)j < 3u ? S.xr[j * kTcRows + r] : 0.f);
      };

      // ---- A operand of the first GEMM step (= layer 1): layer 0 on the CUDA cores.  With W0[:, :L] z folded into
      // zb0, layer 0 is 3 FMAs per output:  h0[j] = relu(zb0[j] + W0[j][L..L+2] . x)  (deep_sdf_decoder.py:91,103).
      // As a GEMM step (K = 80) it cost 4.8k cycles per tile, most of it the dependency bubble. --------------------
      {
        const TcStep s0 = plan.step[0];
        const uint32_t a_t = tmem + (uint32_t)s0.a_reg * 256u + lane_addr;
        const int kk = s0.k_steps * 16;
        const int n0out = dec.out_dim[0];
        // (the three weight rows come from shared memory: as 768 uniform __ldg per thread and tile they were 6k L1
        //  wavefronts per tile, most of the 8 us between a tile's begin and its first MMA)
        const float* w0x = S.w0x;
        const float px = x0, py = x1, pz = x2;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          const int u = grp + 2 * j, n0 = 32 * u;
          if (n0 < kk) {
            float t[32];
            uint32_t mw = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int c = n0 + i;
              float w = S.bias[c];
              w = fmaf(w0x[c], px, w);
              w = fmaf(w0x[kHid + c], py, w);
              w = fmaf(w0x[2 * kHid + c], pz, w);
              const bool on = (c < n0out) && (w > 0.f);
              mw |= (on ? 1u : 0u) << i;
              t[i] = on ? w : 0.f;
            }
            S.maskw[(0 * 8 + u) * kTcRows + r] = mw;
            store_a_unit(a_t + (uint32_t)n0, t);
          } else {
            S.maskw[(0 * 8 + u) * kTcRows + r] = 0u;
          }
          unit_done(&S.a_ready[u]);
        }
      }

      float yv = 0.f;
      for (int s = 0; s < ns; ++s) {
        const TcStep st = plan.step[s];
        const bool more = (s + 1 < ns);
        const int a_next = more ? plan.step[s + 1].a_reg : 0;
        const int k_next = more ? plan.step[s + 1].k_steps * 16 : 0;
        const uint32_t d_t = tmem + (uint32_t)st.d_reg * 256u + lane_addr;
        const uint32_t an_t = tmem + (uint32_t)a_next * 256u + lane_addr;
        mbar_wait(&S.acc_full[0], acc_phase);
        tc_fence_after();
        if (tid == 0) DSPGN_CLK(0);

        if (st.kind == TK_FWD_PENULT) {
          // ---- last hidden layer: bias + ReLU (mask saved), and the final Linear(width, 1) + tanh right here as a per-row
          // dot product on the CUDA cores while the values are in registers.  As an MMA step it was the worst one: N = 16
          // still costs the ~110-cycle floor per instruction (48 MMAs + the dependency bubble = 7.3k cycles for 256 MACs
          // per row) and needed its own TMEM operand.  deep_sdf_decoder.py:91,103,107-108.
          float part = 0.f;
#pragma unroll 1
          for (int j = 0; j < 4; ++j) {
            const int u = grp + 2 * j, n0 = 32 * u;
            uint32_t mw = 0;
            if (n0 < st.n_mma) {
              uint32_t v[32];
              tc_ld32(d_t + (uint32_t)n0, v);
              tc_wait_ld();
              const float* bb = S.bias + st.layer * kHid + n0;
              const float* wl = S.wlast + n0;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float w = __uint_as_float(v[i]) + bb[i];
                mw |= (w > 0.f ? 1u : 0u) << i;
                part = fmaf(fmaxf(w, 0.f), wl[i], part);
              }
            }
            S.maskw[(st.layer * 8 + u) * kTcRows + r] = mw;
          }
          (grp == 0 ? S.rr : S.rsc)[r] = part;             // the two column halves of the row, combined in a fixed order
          epi_bar_sync();                                  // (also: every accumulator read of this step is finished)
          yv = tanhf((S.rr[r] + S.rsc[r]) + S.bias[(st.layer + 1) * kHid]);      // deep_sdf_decoder.py:107-108
          if (fwd_only) {
            if (grp == 0 && r < nrows) {
              const size_t base = (mode == MODE_RAYFWD) ? (size_t)M.smp_off : (size_t)M.pts_off;
              a.sdf_out[base + row0 + r] = (sc != 0.f) ? yv : INFINITY;
            }
            if (!RENDER && mode == MODE_RAYFWD) {        // (persistent kernel: counted by the scan items, dspgn_solve.cuh)
              const unsigned b = __ballot_sync(0xffffffffu, grp == 0 && r < nrows && sc != 0.f);
              if (lane == 0 && b) atomicAdd(a.V_count + o, __popc(b));
            }
          }
          if (more) {
            // seed of the backward chain: g = (1 - y^2) W_last, masked by this layer's ReLU, written into the (dead) A
            // region of this step; the first backward GEMM accumulates into the region whose reads ended at the barrier
            const float gy = 1.f - yv * yv;
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
              const int u = grp + 2 * j, n0 = 32 * u;
              if (n0 < k_next) {
                const uint32_t mw = S.maskw[(st.layer * 8 + u) * kTcRows + r];
                float t[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) t[i] = ((mw >> i) & 1u) ? gy * S.wlast[n0 + i] : 0.f;
                store_a_unit(an_t + (uint32_t)n0, t);
              }
              unit_done(&S.a_ready[u]);
            }
          }
        } else if (st.kind == TK_FWD_HIDDEN) {
#pragma unroll 1
          for (int j = 0; j < 4; ++j) {
            const int u = grp + 2 * j, n0 = 32 * u;
            if (n0 < k_next) {
              float t[32];
              uint32_t mw = 0;
              if (n0 < st.n_mma) {
                uint32_t v[32];
                tc_ld32(d_t + (uint32_t)n0, v);
                tc_wait_ld();
                const float* bb = S.bias + st.layer * kHid + n0;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const float w = __uint_as_float(v[i]) + bb[i];
                  mw |= (w > 0.f ? 1u : 0u) << i;
                  t[i] = fmaxf(w, 0.f);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) t[i] = 0.f;
              }
              S.maskw[(st.layer * 8 + u) * kTcRows + r] = mw;
              if (st.cat_off >= 0 && n0 + 32 > st.cat_off) {      // deep_sdf_decoder.py:87-88: cat[x, input]
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (n0 + i >= st.cat_off) t[i] = inp(n0 + i - st.cat_off);
              }
              store_a_unit(an_t + (uint32_t)n0, t);
            }
            unit_done(&S.a_ready[u]);
          }
        } else if (st.kind == TK_BWD_MID) {
#pragma unroll 1
          for (int j = 0; j < 4; ++j) {
            const int u = grp + 2 * j, n0 = 32 * u;
            if (n0 < st.n_mma) {
              uint32_t v[32];
              tc_ld32(d_t + (uint32_t)n0, v);
              tc_wait_ld();
              const uint32_t mw = S.maskw[(st.mask_layer * 8 + u) * kTcRows + r];
              float t[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) t[i] = ((mw >> i) & 1u) ? __uint_as_float(v[i]) : 0.f;
              if (st.cat_off >= 0 && n0 + 32 > st.cat_off) {      // latent_in skip path -> d/d(input)
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const int ii = n0 + i - st.cat_off;
                  if (ii >= 0) {
                    if (ii < in0) S.Jp[r * kJpStride + ((ii < L) ? ii : (kMaxCode + ii - L))] = __uint_as_float(v[i]);
                    t[i] = 0.f;
                  }
                }
              }
              if (n0 < k_next) store_a_unit(an_t + (uint32_t)n0, t);
            }
            unit_done(&S.a_ready[u]);
          }
        } else {
          // ---- TK_BWD_FIRST: d/d(input) complete -> Jacobian row (loss.py:34-41 / :143-150) -------------
          float* jr = S.Jp + r * kJpStride;
#pragma unroll 1
          for (int j = 0; j < 4; ++j) {
            const int n0 = 32 * (grp + 2 * j);
            if (n0 < st.n_mma && n0 < in0) {
              uint32_t v[32];
              tc_ld32(d_t + (uint32_t)n0, v);         // columns beyond n_mma are never used below
              tc_wait_ld();
              if (L == kMaxCode && n0 + 32 <= kMaxCode) {
                // thread-per-row accesses as float4: with the 76-float row stride a quarter-warp covers all 32 banks
                // (the scalar form was a 4-way bank conflict)
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  float4* pj = reinterpret_cast<float4*>(jr + n0 + i);
                  float4 g = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
                  if (has_skip) { const float4 k = *pj; g.x += k.x; g.y += k.y; g.z += k.z; g.w += k.w; }
                  g.x *= sc; g.y *= sc; g.z *= sc; g.w *= sc;      // loss.py:145 (de_ds) / inactive rows
                  *pj = g;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const int ii = n0 + i;
                  if (ii < in0) {
                    const int jc = (ii < L) ? ii : (kMaxCode + ii - L);
                    float g = __uint_as_float(v[i]);
                    if (has_skip) g += jr[jc];
                    jr[jc] = g * sc;                               // loss.py:145 (de_ds) / inactive rows
                  }
                }
              }
            }
          }
        }
        if (tid == 0) DSPGN_CLK(3);
        acc_phase ^= 1;
      }
      if (!fwd_only) {
      // ---- pose columns, residual (thread = row; needs every d/d(input) column of the row) -----------
      epi_bar_sync();
      if (grp == 0) {
        float* jr = S.Jp + r * kJpStride;
        for (int i = L; i < kMaxCode; ++i) jr[i] = 0.f;
        const float g0 = jr[kMaxCode], g1 = jr[kMaxCode + 1], g2 = jr[kMaxCode + 2];
        // dsdf/dx . [I | -x^ | x] = [g, x cross g, g.x]   (loss_utils.py:166-185)
        jr[kMaxCode + 3] = x1 * g2 - x2 * g1;
        jr[kMaxCode + 4] = x2 * g0 - x0 * g2;
        jr[kMaxCode + 5] = x0 * g1 - x1 * g0;
        jr[kMaxCode + 6] = a.pose_only ? 0.f : (g0 * x0 + g1 * x1 + g2 * x2);
        jr[kMaxCode + 7] = 0.f;
        float res = (mode == MODE_SDF) ? yv : res_in;
        if (sc == 0.f && (mode == MODE_SDF || r >= nrows)) res = 0.f;
        if (mask_out != nullptr && mode == MODE_SDF && r < nrows)
          mask_out[M.pts_off + row0 + r] = (sc != 0.f && fabsf(res) <= 0.05f) ? 1 : 0;      // optimizer.py:76-78
        S.rr[r] = huber_weight(fabsf(res), huber_b) * res;
        S.rsc[r] = (mode == MODE_SDF) ? sc : (r < nrows ? 1.f : 0.f);
        if (a.dbg_J != nullptr && o == a.dbg_obj && mode == MODE_SDF && r < nrows) a.dbg_res[row0 + r] = res;
      }
      epi_bar_sync();
      if (a.dbg_J != nullptr && o == a.dbg_obj && mode == MODE_SDF) {
        const int P = a.dbg_P, npose = a.pose_only ? 6 : 7;
        for (int idx = tid; idx < nrows * P; idx += kTcEpiThreads) {
          const int p = idx / P, c = idx - p * P;
          const int ci = (c < npose) ? (kMaxCode + c) : (c - npose);
          a.dbg_J[(size_t)(row0 + p) * P + c] = S.Jp[p * kJpStride + ci];
        }
      }
      // ---- J^T J, J^T (rho r), loss over the 128 rows of the tile (optimizer.py:161-167) -------------
      float* accp = part + (size_t)tile * kAccStride;
      if (tid < 171) {
        int bi = 0, rem = tid;
        while (rem >= 18 - bi) { rem -= 18 - bi; ++bi; }
        const int bj = bi + rem;
        // Packed fp32 FMAs (fma.rn.f32x2 -> FFMA2, two FMAs per lane and instruction: the plain FFMA pipe issues one
        // warp instruction per two cycles, which made this loop 8k cycles on the two scheduler partitions that hold two of
        // the six warps).  h2[u][w] = (h[u][2w], h[u][2w+1]); every FMA is the same operation as in the scalar form.
        unsigned long long h2[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u) { h2[u][0] = 0ull; h2[u][1] = 0ull; }
        const float* pa = S.Jp + 4 * bi;
        const float* pb = S.Jp + 4 * bj;
#pragma unroll 4
        for (int p = 0; p < kTcRows; ++p) {
          const float4 A4 = *reinterpret_cast<const float4*>(pa + p * kJpStride);
          const ulonglong2 B2 = *reinterpret_cast<const ulonglong2*>(pb + p * kJpStride);     // (b0, b1), (b2, b3)
          const float av[4] = {A4.x, A4.y, A4.z, A4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            unsigned long long aa;
            asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(av[u]));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(h2[u][0]) : "l"(aa), "l"(B2.x));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(h2[u][1]) : "l"(aa), "l"(B2.y));
          }
        }
        float h[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          asm("mov.b64 {%0, %1}, %2;" : "=f"(h[u][0]), "=f"(h[u][1]) : "l"(h2[u][0]));
          asm("mov.b64 {%0, %1}, %2;" : "=f"(h[u][2]), "=f"(h[u][3]) : "l"(h2[u][1]));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int rI = 4 * bi + u, cI = 4 * bj + v;
            if (cI >= rI && cI < kMaxCode + 7) accp[tri_index(rI, cI)] = h[u][v];
          }
      } else if (tid < 171 + kMaxCode + 7) {
        const int c = tid - 171;
        float sacc = 0.f;
        for (int p = 0; p < kTcRows; ++p) sacc = fmaf(S.Jp[p * kJpStride + c], S.rr[p], sacc);
        accp[kAccB + c] = sacc;
      } else if (tid >= 248) {
        // loss and row count: 8 threads x 16 rows, fixed-order combine
        const int k = tid - 248;
        float sacc = 0.f, n = 0.f;
        for (int p = 16 * k; p < 16 * k + 16; ++p) { sacc = fmaf(S.rr[p], S.rr[p], sacc); n += S.rsc[p]; }
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
          sacc += __shfl_down_sync(0xff000000u, sacc, d);
          n += __shfl_down_sync(0xff000000u, n, d);
        }
        if (k == 0) { accp[kAccLoss] = sacc; accp[kAccLoss + 1] = n; }
      }
      }   // !fwd_only
      if (MEGA) {
        // ---- object pipeline.  Per object and iteration:  ray-sample tiles (forward only) -> [last one] per-ray scan +
        // band compaction -> band tiles (fwd+bwd) ;  SDF tiles (fwd+bwd) ;  [last SDF / band tile] solve, pose / code
        // update, tiles of the next iteration.  The CTA that finishes the last tile of a stage runs the serial step
        // with its 256 epilogue threads while every other SM keeps working on other objects.
        __threadfence();                             // this tile's partial sums / sdf values are visible device-wide
        epi_bar_sync();
        if (tid == 0) {
          mega_event(q, EV_TILE_END, mode, o, tr.tile);
          int act = 0;
          if (RENDER && mode == MODE_RAYFWD) { if (atomicSub(q.ray_left + o, 1) == 1) act = 1; }
          else if (atomicSub(q.pending + o, 1) == 1) act = 2;
          *reinterpret_cast<volatile int*>(&S.last_flag) = act;
        }
        epi_bar_sync();
        int act = *reinterpret_cast<volatile int*>(&S.last_flag);
        if (RENDER && act == 1) {
          // every ray sample of the object has its sdf value: the per-ray scan becomes 64-ray work items of its own
          if (tid == 0) {
            const int nch = (M.n_rays + kScanChunkRays - 1) / kScanChunkRays;
            *reinterpret_cast<volatile int*>(q.scan_left + o) = nch;
            __threadfence();
            mega_push(q, kKindScan, o, nch);
          }
        }
        if (act == 2) mega_solve_and_advance<MEGA>(S, o, tid);
      }
      // the next tile's prologue starts with epi_bar_sync(): Jp / rr are not rewritten before it
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tc_dealloc(tmem, 512);
}

__global__ void __launch_bounds__(kTcThreads, 1) k_decoder_tc(TermArgs a) {
  tc_body<0>(a, MegaArgs{}, SolveArgs{}, ScanArgs{});
}
// persistent object-pipelined variants: all GN iterations of all objects in ONE launch.
// k_gn_persistent: SDF tiles only (SDF-only joint runs, pose-only runs); k_gn_persistent_render: joint runs with the
// render term (ray-sample tiles, scan items, band tiles, SDF tiles)
__global__ void __launch_bounds__(kTcThreads, 1) k_gn_persistent(TermArgs a, MegaArgs q, SolveArgs sv) {
  tc_body<1>(a, q, sv, ScanArgs{});
}
__global__ void __launch_bounds__(kTcThreads, 1) k_gn_persistent_render(TermArgs a, MegaArgs q, SolveArgs sv, ScanArgs sc) {
  tc_body<2>(a, q, sv, sc);
}

// ------------------------------------------------------------------------------------------------
// self-test kernel: D[128 x n_mma] = A[128 x 16*k_steps] * B^T through exactly the same operand paths
// (TMEM A written by store_a_block, swizzled weight images, 3-pass split).  Used by tests only.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) k_tc_selftest(const float* __restrict__ A, int lda, const unsigned char* __restrict__ blob,
                                                        int n_mma, int k_steps, float* __restrict__ D) {
  extern __shared__ unsigned char st_raw[];
  unsigned char* ring = st_raw + ((1024u - (smem_u32(st_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar_w, bar_acc;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar_w, 1); mbar_init(&bar_acc, 1); fence_barrier_init(); }
  if (warp == 0) tc_alloc(&tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  const int nch = (k_steps + 3) >> 2;
  // A operand -> TMEM region 0
  for (int u = 0; u < 2 * nch; ++u) {
    float t[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int kk = u * 32 + i;
      t[i] = (kk < k_steps * 16) ? A[(size_t)tid * lda + kk] : 0.f;
    }
    store_a_unit(tmem + lane_addr + (uint32_t)u * 32u, t);
  }
  tc_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t img = (uint32_t)n_mma * 128u;
  const uint32_t idesc = make_idesc(n_mma);
  uint32_t wph = 0;
  for (int c = 0; c < nch; ++c) {
    const int nq = min(4, k_steps - 4 * c);
    if (tid == 0) {
      mbar_expect_tx(&bar_w, 2 * img);
      bulk_g2s(ring, blob + (size_t)(2 * c) * img, img, &bar_w);
      bulk_g2s(ring + kTcStageBytes, blob + (size_t)(2 * c + 1) * img, img, &bar_w);
    }
    mbar_wait(&bar_w, wph);
    wph ^= 1;
    tc_fence_after();
    if (tid == 0) {
      const uint32_t bh = smem_u32(ring), bl = smem_u32(ring + kTcStageBytes);
      const uint32_t a_blk = tmem + (uint32_t)c * 64u;
      for (int q = 0; q < nq; ++q) {
        tc_mma_ts(tmem + 256u, a_blk + a_col_hi(q), make_b_desc(bh + 32u * q), idesc, (c | q) ? 1u : 0u);
        tc_mma_ts(tmem + 256u, a_blk + a_col_hi(q) + 16u, make_b_desc(bh + 32u * q), idesc, 1u);
        tc_mma_ts(tmem + 256u, a_blk + a_col_hi(q), make_b_desc(bl + 32u * q), idesc, 1u);
      }
      tc_commit(&bar_acc);
    }
    mbar_wait(&bar_acc, (uint32_t)(c & 1));      // weights of this chunk consumed before the ring is reused
    tc_fence_after();
  }
  for (int n0 = 0; n0 < n_mma; n0 += 16) {
    uint32_t v[16];
    tc_ld16(tmem + 256u + lane_addr + (uint32_t)n0, v);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)tid * n_mma + n0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tc_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// host side: plan + weight images
// ------------------------------------------------------------------------------------------------
struct TcDecoderHost {
  bool ok = false;
  void* blob = nullptr;
  size_t blob_bytes = 0;
};

// image of B[n][kk] (n < n_mma, kk in [64c, 64c+64)) as fp16 hi / lo, K-major rows of 128 B, 128B swizzle
template <class F>
inline void tc_pack_images(std::vector<unsigned char>& out, int n_mma, int k_steps, F&& elem) {
  const int nch = (k_steps + 3) / 4;
  const size_t img = (size_t)n_mma * 128;
  const size_t base = out.size();
  out.resize(base + (size_t)nch * 2 * img, 0);
  for (int c = 0; c < nch; ++c) {
    unsigned char* hi = out.data() + base + (size_t)(2 * c) * img;
    unsigned char* lo = hi + img;
    for (int n = 0; n < n_mma; ++n)
      for (int e = 0; e < 64; ++e) {
        const int kk = c * 64 + e;
        const float w = (kk < k_steps * 16) ? elem(n, kk) : 0.f;
        const __half h = __float2half_rn(w);
        const __half l = __float2half_rn(w - __half2float(h));
        const size_t off = (size_t)n * 128 + (size_t)(((e >> 3) ^ (n & 7)) << 4) + (size_t)(e & 7) * 2;
        memcpy(hi + off, &h, 2);
        memcpy(lo + off, &l, 2);
      }
  }
}

inline int round16(int x) { return (x + 15) / 16 * 16; }

inline int tc_pack_decoder(const DspgnDecoderSpec& spec, const float* const* W, const float* const* b, TcDecoderHost& h,
                           DecoderDev* dv, std::string& err) {
  (void)b;
  h.ok = false;
  dv->tc_blob = nullptr;
  memset(&dv->tc_plan, 0, sizeof(TcPlan));
  const int nl = spec.num_linear, in0 = spec.latent_size + 3, li = dv->latent_in;
  if (nl != 9 && nl < 3) return 0;
  if (in0 > 80) return 0;
  TcPlan& P = dv->tc_plan;
  std::vector<unsigned char> blob;
  int ns = 0;
  // forward steps: layer k, A = activations (K = in_dim), B[n][kk] = W_k[n][kk].  The final Linear(width, 1) is not a
  // GEMM step: it is folded into the epilogue of the last hidden layer (TK_FWD_PENULT) as a dot product.
  // Layer 0 is no GEMM step either: with its latent part folded into a per-object bias (ObjState.zb0) it is 3 FMAs per
  // output and is evaluated while the first operand is built (tc_body prologue).
  if (nl < 4 || li == 1) return 0;          // needs a hidden GEMM layer after layer 0 and no concat at layer 1
  for (int k = 1; k < nl - 1; ++k) {
    TcStep& s = P.step[ns];
    const int nin = spec.in_dim[k], nout = spec.out_dim[k];
    s.kind = (k == nl - 2) ? TK_FWD_PENULT : TK_FWD_HIDDEN;
    s.n_mma = round16(nout);
    s.k_steps = round16(nin) / 16;
    s.a_reg = ns & 1; s.d_reg = (ns & 1) ^ 1;
    s.layer = k; s.n_real = nout;
    s.cat_off = (k + 1 == li) ? nout : -1;
    s.mask_layer = -1;
    s.w_off = (unsigned)blob.size();
    const float* Wk = W[k];
    tc_pack_images(blob, s.n_mma, s.k_steps, [&](int n, int kk) { return (n < nout && kk < nin) ? Wk[(size_t)n * nin + kk] : 0.f; });
    ++ns;
  }
  P.n_fwd = ns;
  // backward steps: layer k = nl-2 .. 0, A = masked gradient (K = out_dim), B[n][kk] = W_k[kk][n]
  int a_reg = P.step[ns - 1].a_reg;          // the seed overwrites the (dead) A operand of the last forward GEMM step
  for (int k = nl - 2; k >= 0; --k) {
    TcStep& s = P.step[ns];
    const int nin = spec.in_dim[k], nout = spec.out_dim[k];
    s.kind = (k == 0) ? TK_BWD_FIRST : TK_BWD_MID;
    s.n_mma = round16(nin);
    s.k_steps = round16(nout) / 16;
    s.a_reg = a_reg; s.d_reg = a_reg ^ 1;
    a_reg ^= 1;
    s.layer = k; s.n_real = nin;
    s.cat_off = (k == li) ? nin - in0 : -1;
    s.mask_layer = (k > 0) ? k - 1 : -1;
    s.w_off = (unsigned)blob.size();
    const float* Wk = W[k];
    tc_pack_images(blob, s.n_mma, s.k_steps, [&](int n, int kk) { return (n < nin && kk < nout) ? Wk[(size_t)kk * nin + n] : 0.f; });
    ++ns;
  }
  P.n_steps = ns;
  // the operand of step s+1 must have been produced for k_steps(s+1)*16 columns by step s
  void* d = nullptr;
  if (cudaMalloc(&d, blob.size()) != cudaSuccess) { cudaGetLastError(); err = "cudaMalloc(tc blob)"; return DSPGN_E_ALLOC; }
  if (cudaMemcpy(d, blob.data(), blob.size(), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); err = "cudaMemcpy(tc blob)"; return DSPGN_E_CUDA; }
  h.blob = d; h.blob_bytes = blob.size(); h.ok = true;
  dv->tc_blob = reinterpret_cast<const unsigned char*>(d);
  return 0;
}

inline void tc_free_decoder(TcDecoderHost& h) {
  if (h.blob) cudaFree(h.blob);
  h.blob = nullptr; h.ok = false;
}

inline int tc_setup_kernels(std::string& err) {
  if (cudaFuncSetAttribute(k_decoder_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes) != cudaSuccess) {
    err = std::string("cudaFuncSetAttribute(k_decoder_tc): ") + cudaGetErrorString(cudaGetLastError());
    return DSPGN_E_CUDA;
  }
  if (cudaFuncSetAttribute(k_gn_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes) != cudaSuccess ||
      cudaFuncSetAttribute(k_gn_persistent_render, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes) != cudaSuccess) {
    err = std::string("cudaFuncSetAttribute(k_gn_persistent): ") + cudaGetErrorString(cudaGetLastError());
    return DSPGN_E_CUDA;
  }
  if (cudaFuncSetAttribute(k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kTcStageBytes + 1024) != cudaSuccess) {
    err = std::string("cudaFuncSetAttribute(k_tc_selftest): ") + cudaGetErrorString(cudaGetLastError());
    return DSPGN_E_CUDA;
  }
  return 0;
}

inline bool tc_engine_default() { return true; }

inline int tc_launch_term(TermArgs& a, int num_sms, long long tiles_upper, cudaStream_t stream, std::string& err) {
  int grid = (int)std::min<long long>(tiles_upper, num_sms);
  if (grid < 1) grid = 1;
  k_decoder_tc<<<grid, kTcThreads, kTcSmemBytes, stream>>>(a);
  (void)err;
  return 0;
}

}  // namespace dspgn
