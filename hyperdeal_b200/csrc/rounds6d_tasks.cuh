// Task bodies of the three-round 3D3V degree-3 FP64 kernel (kernel_fast6d.cu: k_rounds_3d3v_k3).
//
// The collapsed operator of one cell (basis.hpp; advection_operation.h:221-566 on Cartesian cells, constant velocity)
//     dst = sum_{d=0..5} (I x .. C_d .. x I) u + L_d(i_d) * trace_d(upwind neighbour)
// is evaluated in three rounds of two directions each, (0,1), (2,3), (4,5).  In round r a *task* is the 4x4 tile over
// the directions (A,B) = (2r, 2r+1) for one value of the four other indices: 16 accumulators, 16 values of u, 4 + 4
// trace values, 160 DFMA.  256 tasks per round and cell; a warpgroup of 128 threads runs two tasks per cell.  Round r
// adds to the partial sums of round r-1 in a shared-memory buffer P (round 0 starts it, round 2 writes dst), so the
// accumulators are *initialised* from P and nothing is added twice.
//
// Shared-memory layout of u and P: the cell as 256 rows (i2,i3,i4,i5) of 16 doubles (i0,i1), the 16-byte chunks of a
// row XOR-swizzled with (row & 7) — what a TMA box load with CU_TENSOR_MAP_SWIZZLE_128B produces.
//   round 0: thread t, task j -> row rho = t + 128 j; the tile is the row itself                 (8 x LDS.128)
//   round 1: c16 = t & 15 = (i0,i1), E = (t >> 4) + 8 j = i4 + 4 i5; tile (i2,i3) = rows a + 4 b + 16 E  (16 x LDS.64)
//   round 2: c16 = t & 15,           E = (t >> 4) + 8 j = i2 + 4 i3; tile (i4,i5) = rows E + 16 (a + 4 b) (16 x LDS.64)
// All three are conflict-free (a quarter/half warp reads whole 128-byte rows).
//
// Trace values come straight from global memory (L2): the upwind neighbour's end layer in `src`, or the ghost buffer.
// Element offset of trace value i of a task = fbase[d] + thr + i * stride with (thr, stride) from face_addr() below;
// fbase[d] is computed once per cell by the producer warp (FaceBase).
//
// The file compiles for the device and, with HD_R6_HOST_EMULATION, for the host (tests/rounds6d_emulation.cpp runs the
// task bodies thread by thread against the oracle before any GPU time is spent).
#pragma once
#include <cstdint>

#ifdef HD_R6_HOST_EMULATION
#include <cstring>
#define HD_R6_FN inline
namespace r6emu
{
  extern unsigned char *smem; // emulated shared memory
}
struct r6_double2
{
  double x, y;
};
HD_R6_FN r6_double2
r6_lds128(uint32_t a)
{
  r6_double2 v;
  std::memcpy(&v, r6emu::smem + a, 16);
  return v;
}
HD_R6_FN double
r6_lds64(uint32_t a)
{
  double v;
  std::memcpy(&v, r6emu::smem + a, 8);
  return v;
}
HD_R6_FN void
r6_sts128(uint32_t a, double x, double y)
{
  std::memcpy(r6emu::smem + a, &x, 8);
  std::memcpy(r6emu::smem + a + 8, &y, 8);
}
HD_R6_FN void
r6_sts64(uint32_t a, double x)
{
  std::memcpy(r6emu::smem + a, &x, 8);
}
HD_R6_FN double
r6_ldg(const double *p)
{
  return *p;
}
HD_R6_FN r6_double2
r6_ldg128(const double *p)
{
  return r6_double2{p[0], p[1]};
}
HD_R6_FN void
r6_ldg256(const double *p, double (&f)[4])
{
  for (int i = 0; i < 4; ++i)
    f[i] = p[i];
}
HD_R6_FN double
r6_fma(double a, double b, double c)
{
  return __builtin_fma(a, b, c);
}
#else
#define HD_R6_FN __device__ __forceinline__
typedef double2 r6_double2;
HD_R6_FN r6_double2
r6_lds128(uint32_t a)
{
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
HD_R6_FN double
r6_lds64(uint32_t a)
{
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
HD_R6_FN void
r6_sts128(uint32_t a, double x, double y)
{
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
HD_R6_FN void
r6_sts64(uint32_t a, double x)
{
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(x) : "memory");
}
// global loads that stay where they are written (volatile asm is not moved across the other volatile asm around it):
// the trace values of the NEXT task are requested before the arithmetic of the current one.  Plain (coherent) loads:
// ghost values are written by peer GPUs while the kernel runs.
#ifndef HD_R6_LDG_MOD
#define HD_R6_LDG_MOD "" // e.g. ".cg" (cache in L2 only): A/B knob for the trace loads
#endif
HD_R6_FN double
r6_ldg(const double *p)
{
  double v;
  asm volatile("ld.global" HD_R6_LDG_MOD ".f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
HD_R6_FN r6_double2
r6_ldg128(const double *p)
{
  double2 v;
  asm volatile("ld.global" HD_R6_LDG_MOD ".v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
// four contiguous doubles, 32-byte aligned, as ONE 256-bit load (LDG.E.ENL2.256 on sm_100a): the lanes of round 0 read
// 32 bytes each from 32 different 128-byte lines, and the L1TEX data pipe is charged per line and instruction
HD_R6_FN void
r6_ldg256(const double *p, double (&f)[4])
{
  asm volatile("ld.global" HD_R6_LDG_MOD ".v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(f[0]), "=d"(f[1]), "=d"(f[2]), "=d"(f[3]) : "l"(p));
}
HD_R6_FN double
r6_fma(double a, double b, double c)
{
  return fma(a, b, c);
}
#endif

namespace r6
{
  constexpr int CELL    = 4096;  // doubles per cell
  constexpr int U_BYTES = 32768; // one cell in shared memory

  struct Coef // matrices of round r: directions A = 2r and B = 2r + 1
  {
    double A[3][16], B[3][16]; // [i * 4 + j]
    double LA[3][4], LB[3][4]; // lifting vector of the upwind face (0 if a_d == 0)
  };

  // What the producer warp hands to the compute warps with every cell (64 bytes in shared memory).
  struct CellInfo
  {
    int       cell;  // -1: end of work
    int       flags; // bit 0: first cell of a row walk (the direction-0 trace comes from memory), bit 8 + d: the upwind
                     // neighbour in direction d lies behind a GHOST side (fbase[d] is relative to the ghost buffer)
    long long fbase[6];
    long long pad;
  };

  // element offset of the upwind trace of direction d for the cell with lattice coordinates c:
  //   inside the brick (periodic wrap included): neighbour cell * 4096 + layer * 4^d   (relative to src)
  //   behind a ghost side: ghost_off[d] + face_cell * 1024                            (relative to the ghost buffer)
  struct FaceBase
  {
    long long off;
    bool      ghost;
  };

  // P: anything with ncell[6], up_delta[6] (-1 / +1 / 0: upwind neighbour is the lower / upper cell / none), up_kind[6]
  // (HD_SIDE_* of the brick side the upwind neighbour may lie behind; 1 = HD_SIDE_GHOST) and ghost_off[6]
  // d may be a run-time value (the producer computes one direction per lane): everything that depends on it is picked
  // with compile-time indices, so c[] and the parameter arrays are never indexed dynamically (no local memory).
  template <class P>
  HD_R6_FN FaceBase
  face_base(const P &p, const int (&c)[6], int d)
  {
    FaceBase fbv;
    fbv.off   = 0;
    fbv.ghost = false;
    int       cd = 0, nd = 1, ud = 0, kd = 0;
    long long gd = 0;
#pragma unroll
    for (int e = 0; e < 6; ++e)
      if (e == d)
        {
          cd = c[e];
          nd = p.ncell[e];
          ud = p.up_delta[e];
          kd = p.up_kind[e];
          gd = p.ghost_off[e];
        }
    if (ud == 0)
      return fbv;
    const bool at_side = ud < 0 ? (cd == 0) : (cd == nd - 1);
    if (at_side && kd == 1)
      {
        long long fc = 0;
#pragma unroll
        for (int e = 5; e >= 0; --e)
          if (e != d)
            fc = fc * p.ncell[e] + c[e];
        fbv.off   = gd + fc * 1024;
        fbv.ghost = true;
        return fbv;
      }
    int n = cd + ud;
    if (n < 0)
      n = nd - 1;
    if (n >= nd)
      n = 0;
    long long idx = 0;
#pragma unroll
    for (int e = 5; e >= 0; --e)
      idx = idx * p.ncell[e] + (e == d ? n : c[e]);
    fbv.off = idx * CELL + (long long)(ud < 0 ? 3 : 0) * (1 << (2 * d));
    return fbv;
  }

  // (thr, stride) of the trace addressing, see the file header.  ROUND and SIDE (0 = A, 1 = B) are compile-time.
  template <int ROUND, int SIDE>
  HD_R6_FN void
  face_addr(bool ghost, int t, int j, int &thr, int &stride)
  {
    if (ROUND == 0)
      {
        const int rho = t + 128 * j;
        thr           = ghost ? 4 * rho : 16 * rho;
        stride        = (SIDE == 0 && !ghost) ? 4 : 1;
      }
    else
      {
        const int c16 = t & 15, E = (t >> 4) + 8 * j;
        if (ROUND == 1)
          {
            thr    = c16 + (ghost ? 64 : 256) * E;
            stride = (SIDE == 0 && !ghost) ? 64 : 16;
          }
        else
          {
            thr    = c16 + 16 * E;
            stride = (SIDE == 0 && !ghost) ? 1024 : 256;
          }
      }
  }

  // the same addressing as compile-time constants (the device loop precomputes the thread part once and gives the loads
  // immediate offsets): stride of the four values and thread-offset increment from task 0 to task 1
#ifdef HD_R6_HOST_EMULATION
#define HD_R6_CONSTEXPR constexpr
#else
#define HD_R6_CONSTEXPR __host__ __device__ constexpr
#endif
  HD_R6_CONSTEXPR int
  trace_stride(int round, int side, bool ghost)
  {
    return round == 0 ? ((side == 0 && !ghost) ? 4 : 1) : (round == 1 ? ((side == 0 && !ghost) ? 64 : 16) : ((side == 0 && !ghost) ? 1024 : 256));
  }
  HD_R6_CONSTEXPR int
  trace_task_step(int round, bool ghost)
  {
    return round == 0 ? (ghost ? 4 * 128 : 16 * 128) : (round == 1 ? (ghost ? 64 * 8 : 256 * 8) : 16 * 8);
  }

  // the 4 trace values of one side of a task
  template <int ROUND, int SIDE>
  HD_R6_FN void
  load_trace(const double *src, const double *ghosts, long long fbase, bool ghost, int t, int j, double (&f)[4])
  {
    int thr, stride;
    face_addr<ROUND, SIDE>(ghost, t, j, thr, stride);
    const double *p = (ghost ? ghosts : src) + fbase + thr;
    if (ROUND == 0 && (SIDE == 1 || ghost))
      {
        r6_ldg256(p, f); // four contiguous doubles, 32-byte aligned
      }
    else
      {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          f[i] = r6_ldg(p + (long long)i * stride);
      }
  }

  // per-thread shared-memory offsets (computed once)
  template <int ROUND>
  struct ThreadMap
  {
    uint32_t x[8]; // round 0: chunk offsets of the thread's row; round 1: row-parity dependent offsets; round 2: x[0] only
    HD_R6_FN void
    init(int t)
    {
      if (ROUND == 0)
        {
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            x[ch] = uint32_t(t) * 128u + ((uint32_t(ch) ^ uint32_t(t & 7)) << 4);
        }
      else if (ROUND == 1)
        {
          const uint32_t c16 = uint32_t(t & 15), e = uint32_t(t >> 4);
#pragma unroll
          for (int m = 0; m < 8; ++m)
            x[m] = e * 2048u + (c16 & 1u) * 8u + (((c16 >> 1) ^ uint32_t(m)) << 4);
        }
      else
        {
          const uint32_t c16 = uint32_t(t & 15), e = uint32_t(t >> 4);
          x[0]               = e * 128u + (c16 & 1u) * 8u + (((c16 >> 1) ^ e) << 4);
#pragma unroll
          for (int m = 1; m < 8; ++m)
            x[m] = 0;
        }
    }
    // byte offset of tile element (a, b) of task j inside a cell buffer
    HD_R6_FN uint32_t
    elem(int a, int b, int j) const
    {
      if (ROUND == 1)
        return x[(a + 4 * b) & 7] + uint32_t(a + 4 * b) * 128u + uint32_t(j) * 16384u;
      return x[0] + uint32_t(a + 4 * b) * 2048u + uint32_t(j) * 1024u; // ROUND == 2
    }
  };

  // The tile arithmetic is split in two so that the trace values of the NEXT task can be requested in between:
  //   trace_terms : q[b][a] (+)= LA[a] fa[b] + LB[b] fb[a]        (consumes the traces requested one task ago)
  //   main_terms  : q[b][a] += sum_j A[a][j] U[b][j] + sum_j B[b][j] U[j][a]
  // The order matters on the device: global loads are tracked by a handful of COUNTING scoreboards, so a wait for this
  // task's traces also waits for every younger load on the same scoreboard.  The next request is therefore issued only
  // after the trace terms — straight into the registers they have just freed — and has the whole main part (128 DFMA
  // plus the shared-memory traffic of the task) to arrive.
  template <int ROUND, bool INIT>
  HD_R6_FN void
  trace_terms(const Coef &cf, const double (&fa)[4], const double (&fb)[4], double (&q)[4][4])
  {
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int a = 0; a < 4; ++a)
        q[b][a] = INIT ? cf.LA[ROUND][a] * fa[b] : r6_fma(cf.LA[ROUND][a], fa[b], q[b][a]);
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int a = 0; a < 4; ++a)
        q[b][a] = r6_fma(cf.LB[ROUND][b], fb[a], q[b][a]);
  }

  // compiler fence for the order "trace terms, then the next request": the 16 partial sums are operands of a volatile
  // asm, so they are computed before it, and the (volatile asm) loads of the request stay behind it
  HD_R6_FN void
  pin_values(const double (&q)[4][4])
  {
#ifndef HD_R6_HOST_EMULATION
    asm volatile("" ::"d"(q[0][0]), "d"(q[0][1]), "d"(q[0][2]), "d"(q[0][3]), "d"(q[1][0]), "d"(q[1][1]), "d"(q[1][2]), "d"(q[1][3]), "d"(q[2][0]), "d"(q[2][1]),
                 "d"(q[2][2]), "d"(q[2][3]), "d"(q[3][0]), "d"(q[3][1]), "d"(q[3][2]), "d"(q[3][3])
                 : "memory");
#else
    (void)q;
#endif
  }

  template <int ROUND>
  HD_R6_FN void
  main_terms(const Coef &cf, const double (&U)[4][4], double (&q)[4][4])
  {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int a = 0; a < 4; ++a)
          q[b][a] = r6_fma(cf.A[ROUND][a * 4 + jj], U[b][jj], q[b][a]);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int a = 0; a < 4; ++a)
          q[b][a] = r6_fma(cf.B[ROUND][b * 4 + jj], U[jj][a], q[b][a]);
  }

  // The 16 values of u of a task (its 4x4 tile).  They are loaded by the PREVIOUS task, right after its main part (the
  // registers are free again, and the shared-memory latency disappears behind that task's epilogue): every task function
  // below takes U loaded and calls after_main() when it is done with it.
  template <int ROUND>
  HD_R6_FN void
  load_u(uint32_t ub, const ThreadMap<ROUND> &tm, int j, double (&U)[4][4])
  {
    if (ROUND == 0)
      {
        const uint32_t jo = uint32_t(j) * 16384u;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          {
            const r6_double2 v        = r6_lds128(ub + tm.x[ch] + jo);
            U[ch >> 1][(ch & 1) * 2]     = v.x;
            U[ch >> 1][(ch & 1) * 2 + 1] = v.y;
          }
      }
    else
      {
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int a = 0; a < 4; ++a)
            U[b][a] = r6_lds64(ub + tm.elem(a, b, j));
      }
  }

  // ---- round 0: directions (0,1); writes P.  `edge[b]` returns the thread's own end layer of direction 0 (the trace the
  // next cell of the row walk needs: i0 = 0 when the walk descends, else i0 = 3).  after_traces() is called once fa, fb
  // have been consumed (it may overwrite them: the request for the next task), after_main() once U has been (it loads the
  // next task's U).
  template <class F, class H>
  HD_R6_FN void
  task_round0(const Coef &cf, uint32_t pb, const ThreadMap<0> &tm, int j, double (&U)[4][4], const double (&fa)[4], const double (&fb)[4], bool descend,
              double (&edge)[4], F &&after_traces, H &&after_main)
  {
    double         q[4][4];
    const uint32_t jo = uint32_t(j) * 16384u;
    trace_terms<0, true>(cf, fa, fb, q);
    pin_values(q);
    after_traces();
    main_terms<0>(cf, U, q);
#pragma unroll
    for (int b = 0; b < 4; ++b)
      edge[b] = descend ? U[b][0] : U[b][3];
    pin_values(q);
    after_main();
#pragma unroll
    for (int ch = 0; ch < 8; ++ch)
      r6_sts128(pb + tm.x[ch] + jo, q[ch >> 1][(ch & 1) * 2], q[ch >> 1][(ch & 1) * 2 + 1]);
  }

  // ---- round 1: directions (2,3); P updated in place.  Like round 2 it computes its own contribution from u alone and meets
  // the partial sums only at the end (before_partial() = wait until round 0 has written this task's rows), so that all three
  // rounds can work on the same cell at the same time and no round ever waits at the START of a task.
  template <class F, class H, class G>
  HD_R6_FN void
  task_round1(const Coef &cf, uint32_t pb, const ThreadMap<1> &tm, int j, double (&U)[4][4], const double (&fa)[4], const double (&fb)[4], F &&after_traces,
              H &&after_main, G &&before_partial)
  {
    double q[4][4];
    trace_terms<1, true>(cf, fa, fb, q);
    pin_values(q);
    after_traces();
    main_terms<1>(cf, U, q);
    pin_values(q);
    after_main();
    before_partial();
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int a = 0; a < 4; ++a)
        q[b][a] += r6_lds64(pb + tm.elem(a, b, j));
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int a = 0; a < 4; ++a)
        r6_sts64(pb + tm.elem(a, b, j), q[b][a]);
  }

  // ---- round 2: directions (4,5); returns the finished values K[b][a] of dst index g0 + 256 a + 1024 b,
  // g0 = cell * 4096 + (t & 15) + 16 ((t >> 4) + 8 j).  Its own contribution needs only u, so it is computed first and the
  // partial sums of rounds 0 and 1 are added at the END (before_partial() = wait until round 1 is done with the cell):
  // rounds 1 and 2 work on the same cell side by side and round 0 may already be on the next one (two P buffers).
  template <class F, class H, class G>
  HD_R6_FN void
  task_round2(const Coef &cf, uint32_t pb, const ThreadMap<2> &tm, int j, double (&U)[4][4], const double (&fa)[4], const double (&fb)[4], double (&q)[4][4],
              F &&after_traces, H &&after_main, G &&before_partial)
  {
    trace_terms<2, true>(cf, fa, fb, q);
    pin_values(q);
    after_traces();
    main_terms<2>(cf, U, q);
    pin_values(q);
    after_main();
    before_partial();
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int a = 0; a < 4; ++a)
        q[b][a] += r6_lds64(pb + tm.elem(a, b, j));
  }
} // namespace r6
