// Three-round 3D3V, degree-3, FP64 advection kernel (k_rounds_3d3v_k3) — included by kernel_fast6d.cu inside its
// anonymous namespace (it shares FastParams, the PTX wrappers, the row scheduler conventions and the halo senders).
//
// Why a second kernel: the two-role kernel above keeps two heavy compute warps per SM sub-partition (32 accumulators,
// ~1150 instructions per warp and cell, half of them not FP64); its time is the serial length of that instruction
// stream (profiles/r01h: FP64 pipe 42 %, issue 46 %, nothing saturated).  Here every sub-partition hosts THREE light
// compute warps that run one 160-DFMA task at a time (rounds6d_tasks.cuh), ~1300 instructions per sub-partition and
// cell instead of ~2300, and the face layers never pass through shared memory.
//
// One persistent CTA per SM, 512 threads = four warpgroups, mbarrier-only pipeline:
//   warpgroup r = 0,1,2   round r (directions 2r, 2r+1) of the cells this CTA works on, pipelined over consecutive cells.
//                         The partial sums travel through two shared-memory buffers P: round 0 writes, round 1 updates in
//                         place, round 2 computes its own part from u alone and adds P at the end of each task, then writes
//                         dst — or the fused LSRK update, time_integrators.templates.h:117-132.  So rounds 1 and 2 work on
//                         the same cell side by side, round 0 is up to two cells ahead, and nobody runs in lock-step (with
//                         round 2 *starting* from P the three buffers left no slack: 34 % of all warp time was barrier wait).
//                         The upwind trace values come from global memory (L2) one task ahead of their use; the
//                         direction-0 trace inside a row walk is the thread's own end layer of the previous cell and
//                         stays in registers.
//   warp 12               producer: takes rows of cells from the global counter (same work lists, row tiles and
//                         interior/boundary phases as the two-role kernel), computes the trace base offsets of every
//                         cell (lanes 0-5, one direction each; per row, then one multiply-add per cell) and TMA-loads the
//                         cell (one 32 KiB box, 128 B swizzle) into a 4-stage ring.  Its per-cell latency sets the pace of
//                         the ring (measured with the HD_R6_TRACE timeline: 2400 cycles per cell when it also issued the
//                         prefetches and indexed its coordinates dynamically = local memory; the kernel was producer-bound).
//   warp 13               prefetch warp: follows the producer through the cell info and asks L2 for the face layers the
//                         compute warps will read (cp.async.bulk.prefetch[.tensor]).
//   warps 14-15           only donate their registers (setmaxnreg: compute 152, producer warpgroup 56; 3*152+56 = 512).
// Shared memory: 4 x 32 KiB (cells) + 2 x 32 KiB (partial sums) + 4 x 8 KiB (direction-1 face layer of every cell stage)
// + 256 B (cell info) + barriers = 225.5 KiB.
// Per cell and SM: FP64 pipe 960 warp-DFMA per sub-partition (1920 cycles), shared memory 256 KiB of wavefronts (2048
// cycles), HBM 64 KiB algorithmic.

#ifndef HD_R6_REGS_COMPUTE
#define HD_R6_REGS_COMPUTE 152
#endif
#ifndef HD_R6_REGS_PRODUCER
#define HD_R6_REGS_PRODUCER 56
#endif
#ifndef HD_R6_FUSED_CS
#define HD_R6_FUSED_CS 1 // fused LSRK epilogue: `sol` read and `sol` / `Ti_next` written with .cs (evict-first) accesses, so that the
                         // streams do not push the face layers of src out of L2: DRAM reads 25.4 -> 23.9 GB, 8.58 -> 8.43 ms per stage
                         // (profiles/r02y_fused_stream_ab.txt, r02y_rounds_fused_cs_ncu_summary.json)
#endif
#ifndef HD_R6_UNROLL_TASKS
#define HD_R6_UNROLL_TASKS 1 // 1: both tasks unrolled (35 KB of hot loops, 2.04e9 instructions per apply, ~20 % of the stall samples are instruction fetch), 0: one rolled copy (17 KB, 2.42e9 instructions); A/B within one box: profiles/r02_rounds_ab.txt
#endif
static_assert(3 * HD_R6_REGS_COMPUTE + HD_R6_REGS_PRODUCER <= 512 && HD_R6_REGS_COMPUTE % 8 == 0 && HD_R6_REGS_PRODUCER % 8 == 0,
              "register split exceeds the launch allocation (512 threads x 128 registers)");

constexpr int R6_TRACE_CELLS = 512;
#ifdef HD_R6_TRACE
// event e of cell k as seen by warp `w` of CTA 0 (lane 0 only)
#define R6_TR(w, k, e)                                                                                    \
  do                                                                                                      \
    {                                                                                                     \
      if (p.r6_trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (k) < R6_TRACE_CELLS)               \
        p.r6_trace[((w)*R6_TRACE_CELLS + (k)) * 16 + (e)] = clock64();                                     \
    }                                                                                                     \
  while (0)
#else
#define R6_TR(w, k, e) ((void)0)
#endif
constexpr int R6_THREADS  = 512;
constexpr int R6_STAGES   = 4;
constexpr int R6_PBUFS    = 2;
constexpr int R6_P_OFF    = R6_STAGES * U_BYTES;            // 131072
constexpr int R6_F1_OFF   = R6_P_OFF + R6_PBUFS * U_BYTES;  // 196608: direction-1 face layer of every cell stage (8 KiB each)
constexpr int R6_INFO_OFF = R6_F1_OFF + R6_STAGES * F_BYTES; // 229376
constexpr int R6_BAR_OFF  = R6_INFO_OFF + R6_STAGES * 64;   // 229632
constexpr int R6_SMEM_BYTES = R6_BAR_OFF + 512 + 1024;      // + alignment slack = 231168 <= 232448

struct R6Bars
{
  uint32_t b;
  __device__ __forceinline__ uint32_t fullU(int s) const { return b + 8 * s; }        // cell stage s has landed (and its info is written)
  __device__ __forceinline__ uint32_t emptyU(int s) const { return b + 32 + 8 * s; }  // round 2 is done with it (rounds 0, 1 were before)
  __device__ __forceinline__ uint32_t infoFull(int s) const { return b + 320 + 8 * s; } // the cell info of stage s is written (prefetch warp)
  __device__ __forceinline__ uint32_t pFull1(int i) const { return b + 64 + 8 * i; }  // round 1 has updated P[i]
  __device__ __forceinline__ uint32_t pEmpty(int i) const { return b + 88 + 8 * i; }  // round 2 has read P[i]
  // round 0 -> round 1 is a warp-to-warp, task-to-task dependency: task j of warp w of round 1 reads exactly the rows
  // (i4 in {2 (w & 1), 2 (w & 1) + 1}, i5 = (w >> 1) + 2 j) that task j of warp w of round 0 writes — one barrier each
  __device__ __forceinline__ uint32_t pFull0(int i, int w, int j) const { return b + 128 + 8 * ((i * 4 + w) * 2 + j); }
};

// cell info in shared memory: fbase[6] (48 bytes), cell, flags (r6::CellInfo reordered so that the pair fbase[2r],
// fbase[2r+1] is one aligned 16-byte load)
struct R6Info
{
  int       cell, flags;
  long long fA, fB;
};
template <int R>
__device__ __forceinline__ R6Info
r6_read_info(uint32_t addr)
{
  R6Info i;
  asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(i.cell), "=r"(i.flags) : "r"(addr + 48));
  asm volatile("ld.shared.v2.s64 {%0, %1}, [%2];" : "=l"(i.fA), "=l"(i.fB) : "r"(addr + 16 * R));
  return i;
}

__device__ __forceinline__ void
r6_prefetch_tensor_3d(const CUtensorMap *map, int c0, int c1, int c2)
{
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void
r6_prefetch_bulk(const void *ptr, uint32_t bytes)
{
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
}

// mbarrier wait with a watchdog: a protocol error must end the launch with an error (trap), never hang the GPU.  Kept to
// a handful of instructions (the three rounds share a 32 KB instruction cache): every failed try_wait has already slept
// for the hardware's time slice, so a retry count stands in for a clock (2^26 retries are seconds).
__device__ __forceinline__ void
r6_wait(uint32_t bar, uint32_t parity)
{
  uint32_t done;
  for (uint32_t it = 0;; ++it)
    {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done)
                   : "r"(bar), "r"(parity)
                   : "memory");
      if (done)
        return;
      if (it > (1u << 26))
        asm volatile("trap;");
    }
}

// --------------------------------------------------------------------------------------------- compute warpgroups
template <int R, bool FUSED>
__device__ __forceinline__ void
r6_compute(const FastParams &p, const r6::Coef &cf, const uint32_t base, const R6Bars bars, const int t)
{
  const int  lane    = t & 31;
  const bool actA    = p.up_delta[2 * R] != 0, actB = p.up_delta[2 * R + 1] != 0;
  const bool descend = p.up_delta[0] > 0;
#ifdef HD_R6_STREAM_STORES
  constexpr bool stream = true; // dst written with streaming stores (evict-first in L2); measured: no gain (profiles/r02_rounds_ab.txt)
#else
  constexpr bool stream = false;
#endif
  auto       release = [&](uint32_t bar) {
    __syncwarp();
    if (lane == 0)
      mbar_arrive(bar);
  };
  r6::ThreadMap<R> tm;
  tm.init(t);

  // trace values of a task, requested from global memory (L2).  Round 0, direction 0: only the first cell of a row walk
  // reads its trace from memory; inside a row it is the thread's own end layer of the previous cell (registers).
  // The thread part of the addresses is computed once (r6::face_addr); strides and the task increment are compile-time
  // constants, so a request is one 64-bit add per side plus the loads with immediate offsets.
  int thrS[2], thrG[2];
  {
    int st;
    r6::face_addr<R, 0>(false, t, 0, thrS[0], st);
    r6::face_addr<R, 1>(false, t, 0, thrS[1], st);
    r6::face_addr<R, 0>(true, t, 0, thrG[0], st);
    r6::face_addr<R, 1>(true, t, 0, thrG[1], st);
  }
  auto load4 = [&](const double *q, auto stride_c, double(&f)[4]) {
    constexpr int stride = decltype(stride_c)::value;
    if (stride == 1)
      {
        r6_ldg256(q, f); // four contiguous doubles, 32-byte aligned: one 256-bit load
      }
    else
      {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          f[i] = r6_ldg(q + i * stride);
      }
  };
  constexpr int stepS = r6::trace_task_step(R, false), stepG = r6::trace_task_step(R, true);
  using SA = std::integral_constant<int, r6::trace_stride(R, 0, false)>;
  using GA = std::integral_constant<int, r6::trace_stride(R, 0, true)>;
  using SB = std::integral_constant<int, r6::trace_stride(R, 1, false)>;
  using GB = std::integral_constant<int, r6::trace_stride(R, 1, true)>;
  auto request = [&](const R6Info &inf, int j, double(&fa)[4], double(&fb)[4]) {
    const bool gA = (inf.flags >> (8 + 2 * R)) & 1, gB = (inf.flags >> (9 + 2 * R)) & 1;
    if (actA && (R != 0 || (inf.flags & 1)))
      {
        if (!gA)
          load4(p.src + inf.fA + thrS[0] + j * stepS, SA(), fa);
        else
          load4(p.ghost + inf.fA + thrG[0] + j * stepG, GA(), fa);
      }
    if (actB && R != 0) // (round 0: staged in shared memory, see load_tile)
      {
        if (!gB)
          load4(p.src + inf.fB + thrS[1] + j * stepS, SB(), fb);
        else
          load4(p.ghost + inf.fB + thrG[1] + j * stepG, GB(), fb);
      }
  };

  r6_wait(bars.fullU(0), 0u);
  R6Info cur = r6_read_info<R>(base + R6_INFO_OFF);
  if (cur.cell < 0)
    return;
  // traces of the task about to run; (round 0) eo = the thread's end layer that the task after it will need
  double fa[4] = {0.0, 0.0, 0.0, 0.0}, fb[4] = {0.0, 0.0, 0.0, 0.0}, eo[4] = {0.0, 0.0, 0.0, 0.0};
  request(cur, 0, fa, fb);
  double U[4][4]; // the u tile of the task about to run (loaded by the task before it)
  // round 0: the direction-1 traces of the task come from the face layer the producer staged with the cell (32-byte rows,
  // 32 B swizzle: the two 16-byte halves of row r are swapped when bit 2 of r is set) — two conflict-free LDS.128
  auto load_tile = [&](int stage, int j) {
    r6::load_u<R>(base + uint32_t(stage) * U_BYTES, tm, j, U);
    if (R == 0 && actB)
      {
        const uint32_t r32 = uint32_t(t) + 128u * uint32_t(j), fl = (r32 >> 2) & 1u;
        const uint32_t tb  = base + R6_F1_OFF + uint32_t(stage) * F_BYTES + r32 * 32u;
        const double2  v0 = r6_lds128(tb + ((0u ^ fl) << 4)), v1 = r6_lds128(tb + ((1u ^ fl) << 4));
        fb[0] = v0.x;
        fb[1] = v0.y;
        fb[2] = v1.x;
        fb[3] = v1.y;
      }
  };
  load_tile(0, 0);

  for (int k = 0;; ++k)
    {
      const int      s   = k & (R6_STAGES - 1);
      const int      pi  = k % R6_PBUFS;
      const uint32_t pph = uint32_t((k / R6_PBUFS) & 1);
      const uint32_t ub  = base + uint32_t(s) * U_BYTES;
      const uint32_t pb  = base + R6_P_OFF + uint32_t(pi) * U_BYTES;
      const int tw = R * 4 + (t >> 5); // (trace builds) this warp's row of the timeline
      (void)tw;
      R6_TR(tw, k, 0);
      if (R == 0)
        r6_wait(bars.pEmpty(pi), pph ^ 1u);
      R6_TR(tw, k, 1);
      // (round 1 waits per task for its own warp's rows, round 2 needs the partial sums only at the end of its first task)
      const long long g0 = (long long)cur.cell * CELL + (t & 15) + 16 * (t >> 4);
      R6Info          nxt;
      nxt.cell = 0;
      // both tasks of the cell (HD_R6_UNROLL_TASKS = 0: through ONE copy of the task code — the three rounds run side by
      // side and share the instruction cache)
#if HD_R6_UNROLL_TASKS
#pragma unroll
#else
#pragma unroll 1
#endif
      for (int j = 0; j < 2; ++j)
        {
          // called by the task once it has consumed (fa, fb): request the traces of the task that follows, straight into
          // the same registers (see rounds6d_tasks.cuh on why not earlier)
          double          sv[16]; // (round 2, fused LSRK) the `sol` values this task updates
          const long long g = g0 + 128 * j;
          auto after_traces = [&]() {
            R6_TR(tw, k, 8 + 4 * j); // trace terms done
            if (R == 2 && FUSED)
              {
                // requested here, not at the start of the task: a wait for the traces would wait for these loads as well
#pragma unroll
                for (int i = 0; i < 16; ++i)
#if HD_R6_FUSED_CS
                  asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(sv[i]) : "l"(p.sol + g + 256 * i)); // read once: evict-first, keeps the face layers of src in L2
#else
                  sv[i] = r6_ldg(p.sol + g + 256 * i);
#endif
              }
            if (R == 0)
              {
#pragma unroll
                for (int b = 0; b < 4; ++b)
                  fa[b] = eo[b];
              }
            if (j == 0)
              request(cur, 1, fa, fb);
            else
              {
                R6_TR(tw, k, 2);
                r6_wait(bars.fullU((k + 1) & (R6_STAGES - 1)), uint32_t(((k + 1) / R6_STAGES) & 1));
                R6_TR(tw, k, 3);
                nxt = r6_read_info<R>(base + R6_INFO_OFF + 64u * uint32_t((k + 1) & (R6_STAGES - 1)));
                if (nxt.cell >= 0)
                  request(nxt, 0, fa, fb);
              }
            R6_TR(tw, k, 9 + 4 * j); // next request issued
          };
          // called by the task once it has consumed U: the u tile of the task that follows (its cell has landed: after_traces
          // of task 1 waited for it)
          auto after_main = [&]() {
            R6_TR(tw, k, 10 + 4 * j); // main terms done
            if (j == 0)
              load_tile(s, 1);
            else if (nxt.cell >= 0)
              load_tile((k + 1) & (R6_STAGES - 1), 0);
          };
          if constexpr (R == 0)
            {
              double edge[4];
              r6::task_round0(cf, pb, tm, j, U, fa, fb, descend, edge, after_traces, after_main);
#pragma unroll
              for (int b = 0; b < 4; ++b)
                eo[b] = edge[b];
              release(bars.pFull0(pi, t >> 5, j));
            }
          else if constexpr (R == 1)
            {
              r6::task_round1(cf, pb, tm, j, U, fa, fb, after_traces, after_main, [&]() {
                R6_TR(tw, k, j == 0 ? 4 : 6);
                r6_wait(bars.pFull0(pi, t >> 5, j), pph);
                if (j == 0)
                  R6_TR(tw, k, 5);
              });
            }
          else
            {
              double q[4][4];
              r6::task_round2(cf, pb, tm, j, U, fa, fb, q, after_traces, after_main, [&]() {
                if (j == 0)
                  {
                    R6_TR(tw, k, 4);
                    r6_wait(bars.pFull1(pi), pph);
                    R6_TR(tw, k, 5);
                  }
                else
                  R6_TR(tw, k, 6);
              });
              if (j == 1)
                {
                  // shared memory of this cell is free again (the values are in registers)
                  __syncwarp();
                  if (lane == 0)
                    {
                      mbar_arrive(bars.pEmpty(pi));
                      mbar_arrive(bars.emptyU(s));
                    }
                }
#pragma unroll
              for (int i = 0; i < 16; ++i)
                {
                  const double kv = q[i >> 2][i & 3];
                  if (FUSED)
                    {
#if HD_R6_FUSED_CS
                      __stcs(p.sol + g + 256 * i, fma(p.fb, kv, sv[i]));
                      if (p.fa != 0.0)
                        __stcs(p.ti_next + g + 256 * i, fma(p.fa, kv, sv[i]));
#else
                      p.sol[g + 256 * i] = fma(p.fb, kv, sv[i]);
                      if (p.fa != 0.0)
                        p.ti_next[g + 256 * i] = fma(p.fa, kv, sv[i]);
#endif
                    }
                  else if (stream)
                    __stcs(p.dst + g + 256 * i, kv);
                  else
                    p.dst[g + 256 * i] = kv;
                }
            }
          R6_TR(tw, k, 11 + 4 * j); // task done (results stored)
        }
      if (R == 1)
        release(bars.pFull1(pi));
      R6_TR(tw, k, 7);
      if (nxt.cell < 0)
        break;
      cur = nxt;
    }
}

template <bool FUSED, bool HALO>
__global__ void __launch_bounds__(R6_THREADS, 1)
  k_rounds_3d3v_k3(const __grid_constant__ CUtensorMap mapU, const __grid_constant__ CUtensorMap mapT1, const __grid_constant__ CUtensorMap mapT2,
                   const __grid_constant__ CUtensorMap mapT3, const __grid_constant__ CUtensorMap mapT4, const __grid_constant__ CUtensorMap mapG1,
                   const __grid_constant__ FastParams p, const __grid_constant__ r6::Coef cf)
{
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw   = smem_u32(smem_raw);
  const uint32_t base  = (raw + 1023u) & ~1023u;
  unsigned char *gbase = smem_raw + (base - raw);
  const R6Bars   bars{base + R6_BAR_OFF};

  const int tid  = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0)
    {
      for (int s = 0; s < R6_STAGES; ++s)
        {
          mbar_init(bars.fullU(s), 1);
          mbar_init(bars.emptyU(s), 5); // the four warps of round 2 + the prefetch warp (it reads the cell info)
          mbar_init(bars.infoFull(s), 1);
        }
      for (int i = 0; i < R6_PBUFS; ++i)
        {
          for (int w = 0; w < 4; ++w)
            for (int j = 0; j < 2; ++j)
              mbar_init(bars.pFull0(i, w, j), 1);
          mbar_init(bars.pFull1(i), 4);
          mbar_init(bars.pEmpty(i), 4);
        }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  __syncthreads();

  if (HALO)
    {
      if (p.pass == 3 && p.n_sends > 0 && int(blockIdx.x) < p.n_sender_ctas) // (CTA-uniform)
        halo_send_cta<R6_THREADS>(p);
    }

  if (warp < 12)
    {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 " HD_STR(HD_R6_REGS_COMPUTE) ";");
      if (warp < 4)
        r6_compute<0, FUSED>(p, cf, base, bars, tid);
      else if (warp < 8)
        r6_compute<1, FUSED>(p, cf, base, bars, tid - 128);
      else
        r6_compute<2, FUSED>(p, cf, base, bars, tid - 256);
      return;
    }
  asm volatile("setmaxnreg.dec.sync.aligned.u32 " HD_STR(HD_R6_REGS_PRODUCER) ";");
  if (warp == 13)
    {
      // ===================================================================== prefetch warp
      // follows the producer through the cell info: asks L2 for the upwind face layers the compute warps will read one to
      // three cells later (lanes 1-5, one direction each; lanes 1-4 with ONE tensor-prefetch instruction) and, in the fused
      // LSRK variant, for the cell's `sol` values (lane 0).  Kept off the producer warp: that one sets the pace of the ring.
      if (lane == 0)
        {
          asm volatile("prefetch.tensormap [%0];" ::"l"(&mapT1));
          asm volatile("prefetch.tensormap [%0];" ::"l"(&mapT2));
          asm volatile("prefetch.tensormap [%0];" ::"l"(&mapT3));
          asm volatile("prefetch.tensormap [%0];" ::"l"(&mapT4));
        }
      const CUtensorMap *mymap = lane == 1 ? &mapT1 : (lane == 2 ? &mapT2 : (lane == 3 ? &mapT3 : &mapT4));
      const int          hi    = 1 << (2 * (5 - (lane < 6 ? lane : 5))); // face-layer rows per cell of this lane's map
      const int          ud    = (lane >= 1 && lane < 6) ? p.up_delta[lane] : 0;
      const bool         mine  = ud != 0 && lane >= 2 && ((p.r6_prefetch >> lane) & 1); // (direction 1 is staged by the producer)
      for (int k = 0;; ++k)
        {
          const int s = k & (R6_STAGES - 1);
          r6_wait(bars.infoFull(s), uint32_t((k / R6_STAGES) & 1));
          const uint32_t ia = base + R6_INFO_OFF + 64u * uint32_t(s);
          int            cellk, flags;
          long long      off = 0;
          asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(cellk), "=r"(flags) : "r"(ia + 48));
          if (lane < 6)
            asm volatile("ld.shared.s64 %0, [%1];" : "=l"(off) : "r"(ia + 8u * uint32_t(lane)));
          __syncwarp();
          if (lane == 0)
            mbar_arrive(bars.emptyU(s)); // the info is in registers
          if (cellk < 0)
            break;
          const bool ghost = (flags >> (8 + lane)) & 1;
          if (mine && !ghost)
            {
              if (lane < 5)
                r6_prefetch_tensor_3d(mymap, 0, ud < 0 ? 3 : 0, int(off >> 12) * hi);
              else
                r6_prefetch_bulk(p.src + off, 8192);
            }
          if (FUSED && lane == 0 && (p.r6_prefetch & 64))
            r6_prefetch_bulk(p.sol + (long long)cellk * CELL, U_BYTES);
        }
      return;
    }
  if (warp != 12)
    return;

  // ======================================================================= producer
  const int  n0      = p.ncell[0];
  const bool descend = p.up_delta[0] > 0; // upwind neighbour is the upper cell: walk downwards
  const bool ghost0  = p.up_delta[0] != 0 && p.up_kind[0] == HD_SIDE_GHOST;
  const bool act1    = p.up_delta[1] != 0;
  if (lane == 0)
    {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapT1));
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapU));
    }
  bool     halo_ready = p.pass != 3;
  unsigned seen       = 0; // (deferring pass 3) ghost sides whose arrival counter this producer has already seen at its target
  bool     deferred_phase = false;
  // interior rows: mixed-radix decode that leaves out the ghost layer of every cut direction (see FastParams)
  auto decode_interior = [&](int i, int(&cr)[6]) {
#pragma unroll
    for (int d = 1; d < 6; ++d)
      {
        const bool cut = p.cutg[d] >= 0;
        const int  r   = p.ncell[d] - (cut ? 1 : 0);
        const int  q   = i % r;
        i /= r;
        cr[d] = (cut && p.cutg[d] == 0) ? q + 1 : q;
      }
  };
  auto decode_boundary = [&](int i, int(&cr)[6]) {
    int dk = 0; // the cut direction whose ghost layer this row lies in (lower cut directions are not at theirs)
#pragma unroll
    for (int d = 1; d < 6; ++d)
      if (dk == 0)
        {
          if (i < p.bsize[d])
            dk = d;
          else
            i -= p.bsize[d];
        }
#pragma unroll
    for (int d = 1; d < 6; ++d)
      {
        const bool cut = p.cutg[d] >= 0;
        if (d == dk)
          cr[d] = p.cutg[d];
        else
          {
            const bool skip = cut && d < dk;
            const int  r    = p.ncell[d] - (skip ? 1 : 0);
            const int  q    = i % r;
            i /= r;
            cr[d] = (skip && p.cutg[d] == 0) ? q + 1 : q;
          }
      }
  };
  // work items come from a global counter; the next one is requested while the current row is being loaded
  int next_item = 0;
  if (lane == 0)
    next_item = atomicAdd(p.counters, 1);
  auto fetch_row = [&](int(&cr)[6], int &sb, int &se) -> bool {
    for (;;)
      {
        int item = 0;
        if (!deferred_phase)
          {
            if (lane == 0)
              {
                item = next_item;
                if (item < p.n_items)
                  next_item = atomicAdd(p.counters, 1);
              }
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= p.n_items)
              {
                if (!(p.pass == 3 && p.defer))
                  return false;
                // main list exhausted: on to the deferred rows
                deferred_phase = true;
                if (lane == 0)
                  {
                    __threadfence();
                    atomicAdd(p.counters + 5, 1);
                  }
              }
          }
        if (deferred_phase)
          {
            // take the next deferred row; the queue is complete once every CTA has exhausted the main list
            int ok = 0;
            if (lane == 0)
              {
                item = atomicAdd(p.counters + 4, 1);
                for (;;)
                  {
                    int produced, done;
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(produced) : "l"(p.counters + 3) : "memory");
                    if (item < produced)
                      {
                        ok = 1;
                        break;
                      }
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(done) : "l"(p.counters + 5) : "memory");
                    if (done >= int(gridDim.x))
                      {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(produced) : "l"(p.counters + 3) : "memory");
                        ok = item < produced;
                        break;
                      }
                    __nanosleep(200);
                  }
              }
            ok   = __shfl_sync(0xffffffffu, ok, 0);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (!ok)
              return false;
          }
        cr[0]    = 0;
        sb       = 0;
        se       = n0;
        int mode = p.pass; // 0: all cells of a lattice row, 1: interior list, 2: boundary list
        if (p.pass == 3 && p.defer)
          {
            // the usual tiled lattice order; ghost rows whose halo is not there yet go to the deferred queue
            const bool second = deferred_phase;
            if (second)
              {
                // item = index into the deferred queue (taken from counters[4] by next_deferred below)
                int row = 0;
                if (lane == 0)
                  {
                    volatile int *slot = p.defer_queue + item;
                    while ((row = *slot) == 0)
                      __nanosleep(100);
                    *slot = 0; // leave the queue clean for the next launch
                    row -= 1;
                  }
                item = __shfl_sync(0xffffffffu, row, 0);
              }
            int r = item;
#pragma unroll
            for (int d = 1; d < 6; ++d)
              {
                cr[d] = r % p.tile[d];
                r /= p.tile[d];
              }
#pragma unroll
            for (int d = 1; d < 6; ++d)
              {
                const int nt = p.ncell[d] / p.tile[d];
                cr[d] += (r % nt) * p.tile[d];
                r /= nt;
              }
            unsigned need = 0;
#pragma unroll
            for (int e = 1; e < 6; ++e)
              if (p.cutg[e] >= 0 && cr[e] == p.cutg[e])
                need |= 1u << (2 * e + (p.up_delta[e] < 0 ? 0 : 1));
            need &= ~seen;
            if (need)
              {
                unsigned got = 0;
                if (lane == 0)
                  {
                    unsigned long long t0, t1;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                    for (unsigned todo = need; todo;)
                      {
                        const int i = __ffs(todo) - 1;
                        int       v;
                        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p.halo_flag + i) : "memory");
                        if (v >= p.halo_target)
                          {
                            got |= 1u << i;
                            todo &= todo - 1;
                            continue;
                          }
                        if (!second)
                          break; // first pass over the list: do not wait, defer the row
                        __nanosleep(500);
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                        if (t1 - t0 > 4000000000ull)
                          {
                            atomicExch(p.counters + 2, 1);
                            break;
                          }
                      }
                    if (got)
                      asm volatile("fence.proxy.async;" ::: "memory"); // the TMA (async proxy) reads the direction-1 ghosts
                    if (!second && (need & ~got))
                      {
                        const int idx = atomicAdd(p.counters + 3, 1);
                        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.defer_queue + idx), "r"(item + 1) : "memory");
                      }
                  }
                got = __shfl_sync(0xffffffffu, got, 0);
                seen |= got;
                if (!second && (need & ~got))
                  continue; // deferred
              }
            return true;
          }
        if (p.pass == 3)
          {
            mode = item >= p.n_int ? 2 : 1;
            item -= item >= p.n_int ? p.n_int : 0;
          }
        if (mode == 0)
          {
            int r = item + p.row_begin;
#pragma unroll
            for (int d = 1; d < 6; ++d)
              {
                cr[d] = r % p.tile[d];
                r /= p.tile[d];
              }
#pragma unroll
            for (int d = 1; d < 6; ++d)
              {
                const int nt = p.ncell[d] / p.tile[d];
                cr[d] += (r % nt) * p.tile[d];
                r /= nt;
              }
          }
        else if (mode == 1)
          {
            decode_interior(item, cr);
            if (ghost0)
              sb = 1;
          }
        else if (item < p.n_bnd)
          decode_boundary(item, cr);
        else
          {
            decode_interior(item - p.n_bnd, cr); // the upwind-most cells of the interior rows (direction 0 cut)
            se = 1;
          }
        if (sb >= se)
          continue;
        if (mode == 2 && !halo_ready)
          {
            // the ghost faces are written by the neighbour GPUs while this kernel runs; give up after 4 s (error word)
            if (lane == 0)
              {
                unsigned long long t0, t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                for (unsigned todo = p.halo_mask; todo;)
                  {
                    const int i = __ffs(todo) - 1;
                    int       v;
                    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p.halo_flag + i) : "memory");
                    if (v >= p.halo_target)
                      {
                        todo &= todo - 1;
                        continue;
                      }
                    __nanosleep(500);
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if (t1 - t0 > 4000000000ull)
                      {
                        atomicExch(p.counters + 2, 1);
                        break;
                      }
                  }
                asm volatile("fence.proxy.async;" ::: "memory"); // the TMA (async proxy) reads the direction-1 ghosts
              }
            __syncwarp();
            halo_ready = true;
          }
        return true;
      }
  };
  int c[6], sb = 0, se = 0, k = 0;
  while (fetch_row(c, sb, se))
    {
      // per row: cell index of c_0 = 0 and, in lanes 1-5, the trace base of direction `lane` for c_0 = 0 — inside the row it
      // moves with c_0 (one cell = 4096 values in src, one face cell = 1024 values in the ghost buffer)
      c[0] = 0;
      const long long rowcell = cell_index(p, c);
      r6::FaceBase    rowfb;
      rowfb.off   = 0;
      rowfb.ghost = false;
      if (lane >= 1 && lane < 6)
        rowfb = r6::face_base(p, c, lane);
      for (int step = sb; step < se; ++step, ++k)
        {
          c[0]                 = descend ? n0 - 1 - step : step;
          const long long cell = rowcell + c[0];
          const int       s    = k & (R6_STAGES - 1);
          R6_TR(12, k, 0);
          r6_wait(bars.emptyU(s), uint32_t((k / R6_STAGES) & 1) ^ 1u);
          R6_TR(12, k, 1);
          unsigned char *info = gbase + R6_INFO_OFF + 64 * s;
          // lanes 0-5: trace base of one direction each
          r6::FaceBase fbv = rowfb;
          if (lane == 0)
            fbv = r6::face_base(p, c, 0);
          else if (lane < 6 && p.up_delta[lane] != 0)
            fbv.off += (long long)c[0] * (rowfb.ghost ? 1024 : CELL);
          if (lane < 6)
            reinterpret_cast<long long *>(info)[lane] = fbv.off;
          const unsigned  gmask  = __ballot_sync(0xffffffffu, fbv.ghost) & 0x3fu;
          const long long off1   = __shfl_sync(0xffffffffu, fbv.off, 1); // lane 1's direction-1 trace base
          const bool      ghost1 = (gmask >> 1) & 1u;
          __syncwarp();
          if (lane == 0)
            {
              reinterpret_cast<int *>(info)[12] = int(cell);
              reinterpret_cast<int *>(info)[13] = ((step == sb) ? 1 : 0) | int(gmask << 8);
              mbar_arrive(bars.infoFull(s)); // (release: the prefetch warp may read the info now)
#ifdef HD_R6_DEBUG_SKIP_F1 // timing experiment only (wrong results): what does fetching the direction-1 face layer cost?
              const bool f1 = false;
#else
              const bool f1 = act1;
#endif
              mbar_expect_tx(bars.fullU(s), U_BYTES + (f1 ? F_BYTES : 0));
              tma_load_2d(base + s * U_BYTES, &mapU, 0, int(cell * 256), bars.fullU(s)); // one box of 256 rows = the cell
              if (f1)
                {
                  // the upwind face layer of direction 1 (256 pieces of 32 bytes, one per row of the neighbour cell — the one
                  // trace no warp can read from global memory in a coalesced way) lands next to the cell
                  const uint32_t dstF = base + R6_F1_OFF + s * F_BYTES;
                  if (ghost1)
                    tma_load_2d(dstF, &mapG1, 0, int(off1 >> 2), bars.fullU(s)); // ghost segment viewed as rows of 4 doubles
                  else
                    tma_load_3d(dstF, &mapT1, 0, p.up_delta[1] < 0 ? 3 : 0, int(off1 >> 12) * 256, bars.fullU(s));
                }
              R6_TR(12, k, 2);
            }
        }
    }
  // end marker
  {
    const int s = k & (R6_STAGES - 1);
    r6_wait(bars.emptyU(s), uint32_t((k / R6_STAGES) & 1) ^ 1u);
    if (lane == 0)
      {
        reinterpret_cast<int *>(gbase + R6_INFO_OFF + 64 * s)[12] = -1;
        mbar_arrive(bars.infoFull(s));
        mbar_arrive(bars.fullU(s));
        // the last CTA to finish re-arms the row counter for the next launch
        __threadfence();
        const int done = atomicAdd(p.counters + 1, 1);
        if (done == int(gridDim.x) - 1)
          {
            p.counters[0] = 0;
            p.counters[1] = 0;
            p.counters[3] = 0;
            p.counters[4] = 0;
            p.counters[5] = 0;
            __threadfence();
          }
      }
  }
}
