// The two 3D3V, degree-3, FP64 advection kernels for sm_100a (the BASELINE.json headline case) and everything they share:
// tensor maps, row scheduler, work lists of the interior/boundary phases, halo sender CTAs, launch code.
//   * k_rounds_3d3v_k3 (kernel_rounds6d.cuh, included below; round 2; the default): three light compute warpgroups, one
//     round of two directions each, traces read through L2 — see the header of that file;
//   * k_advect_3d3v_k3 (this file; round 1; hd_advection_set_kernel(op, 2) or HD_FAST_VARIANT=pipe): two heavy compute
//     warpgroups, face layers staged by TMA — described in the rest of this comment.
//
// What it computes (same collapsed form as kernels_generic.cu, basis.hpp):
//   dst_cell = sum_{d=0..5} (I x .. C_d .. x I) u_cell + L_d(i_d) * trace_d(upwind neighbour)
// i.e. advection_operation.h:221-566 for Cartesian cells and a constant velocity, 30 DFMA per DoF.
//
// How it is mapped onto one SM (one persistent CTA per SM, three warpgroups of 128 threads, mbarrier-only pipeline):
//   warp 8      cell producer: takes rows of cells (along x_0) from a global atomic counter (the next item is requested
//               one item ahead) — so the 148 CTAs sweep the lattice as one compact window and neighbour faces are still
//               in L2 — and TMA-loads (cp.async.bulk.tensor, 128B swizzle) each cell (32 KiB) into a 3-deep ring, the
//               upwind face layers of directions 1 and 5 into a 2-deep ring, and gathers the direction-0 trace of the
//               first cell of a row with cp.async.  Rows are visited tile by tile (FastParams::tile) for L2 re-use.
//   warp 9      face producer for round 2: TMA-loads the upwind face layers of directions 2,3,4 (8 KiB each, from src
//               or from the ghost buffer) into a 2-slot ring.   (warps 10, 11 only donate their registers)
//   warps 0-3   round 1, directions (0,1 | 5): thread (i2,i3,i4; half h) owns two of the four output planes of a 4x4x4
//               sub-tensor over (i0,i1,i5): 32 FP64 accumulators, u streamed plane pair by plane pair from shared
//               memory.  The direction-0 neighbour trace never touches memory inside a row: cells are walked in upwind
//               order and the thread keeps the previous cell's end layer in 8 registers.  Result -> shared buffer.
//   warps 4-7   round 2, directions (2,3 | 4) on the same cell at the same time; adds round 1's sums in its epilogue,
//               then coalesced 128 B-per-half-warp stores to dst — or the fused LSRK update (sol += b dt K,
//               Ti' = sol_old + a dt K, time_integrators.templates.h:117-132) so K is never written.
// Registers: launched with 168 per thread; the producer warpgroup drops to 72 and the compute warpgroups rise to 216
// (setmaxnreg).  All (k+1)x(k+1) matrices sit in the kernel-parameter constant bank (uniform-register DFMA operands).
// The hot loops of both compute roles plus the producers must fit the 32 KB instruction cache: every variant that
// grew them lost 3-25 % (profiles/r01g_variants.txt), hence the rolled plane-pair loop.
//
// Shared memory: 3x32 (cells) + 2x16 (faces 1,5) + 32 (partial sums) + 2x24 (faces 2,3,4) + 8 (trace) KiB.
// Multi-GPU (pass 3, fused halo): the first n_sender_ctas CTAs begin by storing the brick's boundary layers into the
// neighbour GPUs' ghost buffers over NVLink, interior cells run meanwhile, boundary cells after the arrival counters.
// Algorithmic traffic 16 B/DoF (fused: 32 B/DoF); see DESIGN.md §4 for the roofline budget.
#include <cuda.h>

#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <type_traits>
#include <vector>

#include "hd_internal.h"
#include "rounds6d_tasks.cuh"

// Register budget.  An SM sub-partition holds 16384 registers = 512 per lane and hosts one warp of every warpgroup.
// The kernel is launched with 168 registers per thread (3 x 168 = 504 <= 512); the producer warpgroup then hands
// registers back (setmaxnreg.dec) and the two compute warpgroups take them (setmaxnreg.inc).  The pool is what the
// launch allocated, so 2 x HD_REGS_COMPUTE + HD_REGS_PRODUCER <= 504 — an inc beyond that waits forever.
#ifndef HD_REGS_COMPUTE
#define HD_REGS_COMPUTE 216
#endif
#ifndef HD_REGS_PRODUCER
#define HD_REGS_PRODUCER 72
#endif
static_assert(2 * HD_REGS_COMPUTE + HD_REGS_PRODUCER <= 504 && HD_REGS_COMPUTE % 8 == 0 && HD_REGS_PRODUCER % 8 == 0, "register split exceeds the launch allocation");
#define HD_STR2(x) #x
#define HD_STR(x) HD_STR2(x)
#ifndef HD_HINTS
#define HD_HINTS 0 // 1: compile the L2-hint code paths (hd_advection_set_l2_hints) in.  Off: no hint combination ever gained anything, and the 4.7 KB of code they add cost 3-5 % (instruction cache, profiles/r01g_variants.txt)
#endif

namespace
{
  constexpr int CELL       = 4096; // doubles per cell
#define HD_DEFAULT_ROW_TILE 0, 2, 2, 2, 0 // measured on 8^6 cells: DRAM reads 12.9 instead of 14.0 GB per apply, profiles/r01f_row_tile_sweep.txt
  constexpr int STAGES     = 3;
  constexpr int THREADS    = 384; // three warpgroups: warps 0-3 round 1, 4-7 round 2, 8 cell producer, 9 face producer (10, 11 idle)
  constexpr int U_BYTES    = 32768;
  constexpr int F_BYTES    = 8192;
  constexpr int R1F_OFF    = STAGES * U_BYTES;          // 98304: 2 slots x (direction 1, direction 5)
  constexpr int ACC_OFF    = R1F_OFF + 2 * 2 * F_BYTES; // 131072: round 1's partial sums (one cell)
  constexpr int R2F_OFF    = ACC_OFF + U_BYTES;         // 163840: 2 slots x (directions 2, 3, 4)
  constexpr int T0_OFF     = R2F_OFF + 2 * 3 * F_BYTES; // 212992
  constexpr int INFO_OFF   = T0_OFF + F_BYTES;          // 221184
  constexpr int BAR_OFF    = INFO_OFF + 128;
  constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;      // + alignment slack

  struct RoleCoef // matrices of one round: in-plane directions A, B and the plane direction C
  {
    double A[16], B[16], C[16]; // [i*4+j]
    double LA[4], LB[4], LC[4]; // lifting vector of the upwind face (0 if a_d == 0)
  };
  struct FastCoef
  {
    RoleCoef r[2]; // round 1: directions (0,1 | 5), round 2: (2,3 | 4)
  };

  // The (k+1)x(k+1) matrices of a launch travel as a __grid_constant__ kernel parameter: DFMA reads them straight
  // from the parameter constant bank (uniform operands), no upload, launches stay independent of each other.

  struct FastParams
  {
    const double *src;
    double *      dst;
    const double *ghost;
    int           ncell[6];
    int           up_delta[6];  // -1: upwind neighbour is the lower cell, +1: the upper one, 0: none
    int           up_kind[6];   // HD_SIDE_* of the brick side the upwind neighbour may lie behind
    long long     ghost_off[6]; // ghost segment of that side
    int           nrows;
    int           row_begin, row_end; // rows [row_begin, row_end) of the lattice are processed (default: all)
    // interior / boundary work lists (passes 1-3), enumerated directly: a row direction d (1..5) whose upwind side is a
    // GHOST side has its ghost-reading layer at coordinate cutg[d] (-1: direction not cut).  Interior rows = no cut
    // direction at its ghost layer (n_int of them); boundary rows = union over the cut directions k of
    // {c_k at the ghost layer, cut directions below k not} with bsize[k] rows each (n_bnd in total).
    int           cutg[6], bsize[6];
    int           n_int, n_bnd, n_items;
    int *         counters; // [0] next row, [1] finished CTAs (self-resetting)
    double *      sol;
    double *      ti_next;
    double        fb, fa;
    int           pass; // 0 all cells, 1 cells that need no ghost data, 2 cells that need ghost data, 3 = 1 then 2 in one launch
    // pass 3 (fused halo): warp 10 of every CTA stores its share of the brick's boundary layers into the neighbours'
    // ghost segments (peer-mapped pointers) and adds 1 to the neighbours' arrival counters; the boundary phase starts
    // once halo_flag[i] >= halo_target for all i in halo_mask
    const int *   halo_flag;
    int           halo_target;
    unsigned      halo_mask;
    int           n_sends;
    int           n_sender_ctas;
    int           send_dir[6], send_side[6];
    double *      send_dst[6];
    int *         send_flag[6];
    // L2 residency hints (bit mask, HD_L2_HINTS): 1 = the direction-4 outflow layer of every cell is loaded evict_last
    // (its downwind neighbour reads it ncell[1..3] rows later), 2 = the direction-4/5 face loads (last use resp.
    // streaming miss) are evict_first, 4 = streaming (.cs) stores/loads of dst, sol, Ti'
    int           hints;
    // Row order: rows are handed out tile by tile (tile[d] rows along direction d = 1..5, lexicographic inside a tile
    // and over the tiles) so that the upwind face layers of ALL five row directions are re-used from L2 within a tile:
    // inside a tile the re-use distance of direction d is tile[1]*..*tile[d-1] rows of 2 x 256 KiB traffic.
    // tile[d] == ncell[d] for every d is the plain lattice order.
    int           tile[6];
    // three-round kernel: bit d (1..5) = the producer asks for the upwind face layer of direction d of every cell to be in
    // L2 before the compute warps read it; bit 6 = same for the cell's `sol` values (fused LSRK)
    int           r6_prefetch;
    // three-round kernel, pass 3 (fused halo) without an interior/boundary split: ONE list of all rows in the usual tiled
    // lattice order; a row that reads a ghost side whose halo has not arrived yet is put on a device-wide deferred queue
    // and taken up again once the main list is exhausted.  Only the ghost rows met during the first ~ms are ever deferred,
    // everything else runs in lattice order with its L2 / TLB locality.
    // counters: [0] next row, [1] finished CTAs, [2] time-out flag, [3] deferred rows queued, [4] deferred rows taken,
    // [5] CTAs that have exhausted the main list
    int           defer;
    int *         defer_queue; // nrows entries, 0 = empty slot, else row item + 1 (self-cleaning)
    // HD_R6_TRACE builds only: timeline of CTA 0 (clock64 per event), long long [13 warps][R6_TRACE_CELLS][16 events]
    long long *   r6_trace;
  };

  struct CellInfo // 32 bytes, one per cell-ring stage
  {
    int cell; // -1: end of work
    int c[6];
    int first; // first cell of a row
  };

  // ---------------------------------------------------------------- PTX wrappers
  __device__ __forceinline__ uint32_t
  smem_u32(const void *p)
  {
    return (uint32_t)__cvta_generic_to_shared(p);
  }
  __device__ __forceinline__ void
  mbar_init(uint32_t bar, uint32_t count)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
  }
  __device__ __forceinline__ void
  mbar_expect_tx(uint32_t bar, uint32_t bytes)
  {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  }
  __device__ __forceinline__ void
  mbar_arrive(uint32_t bar)
  {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
  }
  __device__ __forceinline__ void
  mbar_wait(uint32_t bar, uint32_t parity)
  {
    uint32_t done;
    do
      {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
      }
    while (!done);
  }
  __device__ __forceinline__ void
  tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
  {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
  }
  __device__ __forceinline__ void
  tma_load_2d_hint(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar, uint64_t policy)
  {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar), "l"(policy)
                 : "memory");
  }
  __device__ __forceinline__ void
  tma_load_3d_hint(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar, uint64_t policy)
  {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(policy)
                 : "memory");
  }
  __device__ __forceinline__ uint64_t
  policy_evict_last()
  {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
  }
  __device__ __forceinline__ uint64_t
  policy_evict_first()
  {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
  }
  __device__ __forceinline__ void
  tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar)
  {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
  }
  __device__ __forceinline__ void
  bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
  {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
  }
  __device__ __forceinline__ void
  cp_async_8(uint32_t dst, const void *src)
  {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
  }
  __device__ __forceinline__ void
  cp_async_arrive_noinc(uint32_t bar)
  {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
  }
  __device__ __forceinline__ double2
  lds128(uint32_t addr)
  {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
  }
  __device__ __forceinline__ double
  lds64(uint32_t addr)
  {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
  }
  __device__ __forceinline__ int4
  lds_int4(uint32_t addr)
  {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
  }
  // global load that stays where it is written (asm volatile is not moved across the other volatile asm around it)
  __device__ __forceinline__ double
  ldg_f64_here(const double *p)
  {
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
  }
  __device__ __forceinline__ void
  sts128(uint32_t addr, double a, double b)
  {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
  }

  __device__ __forceinline__ long long
  cell_index(const FastParams &p, const int (&c)[6])
  {
    long long idx = 0;
#pragma unroll
    for (int d = 5; d >= 0; --d)
      idx = idx * p.ncell[d] + c[d];
    return idx;
  }

  // does the cell with coordinates c need ghost data in direction d (upwind neighbour behind a GHOST side)?
  __device__ __forceinline__ bool
  needs_ghost(const FastParams &p, const int (&c)[6], int d)
  {
    if (p.up_delta[d] == 0 || p.up_kind[d] != HD_SIDE_GHOST)
      return false;
    return p.up_delta[d] < 0 ? (c[d] == 0) : (c[d] == p.ncell[d] - 1);
  }

  // upwind neighbour cell (inside the brick, possibly wrapped); only valid if !needs_ghost
  __device__ __forceinline__ long long
  upwind_cell(const FastParams &p, const int (&c)[6], int d)
  {
    int cc[6];
#pragma unroll
    for (int e = 0; e < 6; ++e)
      cc[e] = c[e];
    int n = c[d] + p.up_delta[d];
    if (n < 0)
      n = p.ncell[d] - 1;
    if (n >= p.ncell[d])
      n = 0;
    cc[d] = n;
    return cell_index(p, cc);
  }

  // face-cell index (cell index with coordinate d removed)
  __device__ __forceinline__ long long
  face_cell(const FastParams &p, const int (&c)[6], int d)
  {
    long long fc = 0;
#pragma unroll
    for (int e = 5; e >= 0; --e)
      if (e != d)
        fc = fc * p.ncell[e] + c[e];
    return fc;
  }

  // mbarrier slots (8 bytes each, 256 bytes reserved at BAR_OFF)
  struct Bars
  {
    uint32_t b;
    __device__ __forceinline__ uint32_t fullU(int s) const { return b + 8 * s; }            // cell stage s has landed
    __device__ __forceinline__ uint32_t emptyU(int s) const { return b + 32 + 8 * s; }      // 8 compute warps + face producer are done with it
    __device__ __forceinline__ uint32_t r1fFull(int f) const { return b + 64 + 8 * f; }     // faces of directions 1,5
    __device__ __forceinline__ uint32_t r1fEmpty(int f) const { return b + 80 + 8 * f; }
    __device__ __forceinline__ uint32_t r2fFull(int f, int j) const { return b + 96 + 8 * (3 * f + j); } // face of direction 2+j
    __device__ __forceinline__ uint32_t r2fEmpty(int f, int j) const { return b + 144 + 8 * (3 * f + j); }
    __device__ __forceinline__ uint32_t accFull() const { return b + 192; }                 // round 1's partial sums are in shared memory
    __device__ __forceinline__ uint32_t accEmpty() const { return b + 200; }
    __device__ __forceinline__ uint32_t t0Full() const { return b + 208; }                  // direction-0 trace of a row start
    __device__ __forceinline__ uint32_t t0Empty() const { return b + 216; }
    __device__ __forceinline__ uint32_t infoFull(int s) const { return b + 224 + 8 * s; }   // CellInfo of stage s is written
  };

  // ---------------------------------------------------------------------------------------------
  // Compute warps (4 per round; every SM sub-partition hosts one warp of each round so that the FP64
  // pipe always has a second instruction stream to issue from).  Both rounds work on the SAME cell at the
  // same time (a ring stage lives for one cell time, so the 3-stage ring gives two cells of load lookahead).
  // Round ROLE works on directions (A,B | C) = (0,1 | 5) resp. (2,3 | 4).  A 4x4x4 tile [c][b][a] is shared
  // by two threads: thread half h owns the output planes c = 2h, 2h+1 (32 accumulators) and streams all four
  // source planes s:
  //     acc[c][b][a] += CC[c][s] P_s[b][a]                                                       (every s)
  //     acc[s][b][a] += sum_j CA[a][j] P_s[b][j] + sum_j CB[b][j] P_s[j][a] + LA[a] fa[b] + LB[b] fb[a]   (s in own half)
  //     acc[c][b][a] += LC[c] fc[b][a]                                                           (at the end)
  // Round 1 hands its sums to round 2 through one shared-memory buffer; round 2 adds them in its epilogue.
  // The plane loop is NOT unrolled: both rounds' FP64 cores stay resident in the instruction cache (the fully
  // unrolled form, 2 x 25 KiB, lost 30-45 % of its issue slots to instruction fetch: profiles/r01b).
  template <int ROLE, bool FUSED>
  __device__ __forceinline__ void
  compute_round(const FastParams &p, const FastCoef &cf, const uint32_t base, unsigned char *gbase, const Bars bars, const int tid_in_role)
  {
    const bool      act0 = p.up_delta[0] != 0, act1 = p.up_delta[1] != 0, act5 = p.up_delta[5] != 0;
    const bool      r1faces = act1 || act5;
    const bool      descend = p.up_delta[0] > 0;
    const bool      stream  = HD_HINTS && (p.hints & 4) != 0;
    constexpr int   role    = ROLE;
    const RoleCoef &rc      = cf.r[ROLE];
    const int       lane    = tid_in_role & 31;
    const int       h       = tid_in_role >> 6; // which half of the planes (warp-uniform)
    const int       t       = tid_in_role & 63;
    // one lane per warp signals the consumer-release barriers (after __syncwarp: all lanes' reads are done)
    auto release = [&](uint32_t bar) {
      __syncwarp();
      if (lane == 0)
        mbar_arrive(bar);
    };
    // round-1 addressing: thread (i2,i3,i4) = row t of each i5 block, 16 contiguous (swizzled) doubles
    const uint32_t sw   = uint32_t(t & 7);
    const uint32_t rowU = uint32_t(t) * 128u;
    // round-2 addressing: thread (i0,i1,i5) = column cc of the rows (i2,i3,i4) + 64 i5
    const int      cc      = t & 15;
    const int      i5      = t >> 4;
    const uint32_t col     = uint32_t(cc >> 1) << 4;
    const uint32_t rowbase = uint32_t(i5) * 8192u + uint32_t(cc & 1) * 8u;
    const bool     actA = role ? (p.up_delta[2] != 0) : act0;
    const bool     actB = role ? (p.up_delta[3] != 0) : act1;
    const bool     actC = role ? (p.up_delta[4] != 0) : act5;
    // partial-sum buffer (same swizzle as u): round 1 writes rows t, round 2 reads column cc
    const uint32_t ab = role == 0 ? base + ACC_OFF + rowU + uint32_t(h) * 16384u : base + ACC_OFF + rowbase + uint32_t(h) * 4096u;
    int            nrow_seq = 0;
    // round 1: the direction-0 trace a cell needs is the end layer of the previous cell of the row, i.e. values this
    // very thread held in its planes: keep them in registers (own planes only)
    double tr[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};

    for (int k = 0;; ++k)
      {
        const int      s  = k % STAGES;
        const int      f  = k & 1; // face-ring slot
        const uint32_t pf = uint32_t((k >> 1) & 1);
        const uint32_t pk = uint32_t(k & 1);
        const uint32_t ub = base + s * U_BYTES;
        mbar_wait(bars.fullU(s), uint32_t((k / STAGES) & 1));
        const int4 inf    = lds_int4(base + INFO_OFF + 32 * s + 16); // c[3], c[4], c[5], first
        const int  cellid = *reinterpret_cast<const volatile int *>(gbase + INFO_OFF + 32 * s);
        if (cellid < 0)
          break;
        // fused LSRK update: the 32 values of `sol` this thread updates in the epilogue are requested now (first half)
        // and after the plane loop (second half), so that their HBM latency is hidden behind the cell's arithmetic
        double          svA[16], svB[16];
        const long long gfu = (long long)cellid * CELL + cc + 1024 * i5 + 512 * h;
        if (FUSED && role == 1)
          {
#pragma unroll
            for (int q = 0; q < 16; ++q)
              svA[q] = ldg_f64_here(p.sol + gfu + q * 16);
          }

        double acc[2][4][4]; // [c - 2h][b][a]
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int z = 0; z < 4; ++z)
              acc[x][y][z] = 0.0;
        // ---- prologue
        const bool     from_t0 = (role == 0) && inf.w != 0;
        uint32_t       t0b = 0, fbuf = 0;
        const uint32_t fcol = base + R2F_OFF + uint32_t(f) * (3u * F_BYTES) + 8u * uint32_t(cc + 256 * i5);
        if (role == 0)
          {
            // direction-0 trace of the upwind neighbour at a row start: gathered by the producer
            if (from_t0 && act0)
              {
                mbar_wait(bars.t0Full(), uint32_t(nrow_seq & 1));
                ++nrow_seq;
              }
            t0b = base + T0_OFF + uint32_t(t) * 32u;
            if (r1faces)
              mbar_wait(bars.r1fFull(f), pf);
            fbuf = base + R1F_OFF + f * 2 * F_BYTES;
          }
        else
          {
            if (actA)
              mbar_wait(bars.r2fFull(f, 0), pf);
            if (actB)
              mbar_wait(bars.r2fFull(f, 1), pf);
          }
        const double lc0 = rc.LC[2 * h], lc1 = rc.LC[2 * h + 1];

        // ---- source planes (rolled)
        // (own planes first: the in-plane faces are released after two of the four iterations)
        auto load_plane = [&](double(&Q)[4][4], int spl) {
          if (role == 0)
            {
              const uint32_t up = ub + rowU + uint32_t(spl) * 8192u;
#pragma unroll
              for (int ch = 0; ch < 8; ++ch)
                {
                  const double2 v              = lds128(up + ((uint32_t(ch) ^ sw) << 4));
                  Q[ch >> 1][(ch & 1) * 2]     = v.x;
                  Q[ch >> 1][(ch & 1) * 2 + 1] = v.y;
                }
            }
          else
            {
              const uint32_t up = ub + rowbase + uint32_t(spl) * 2048u;
#pragma unroll
              for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int a = 0; a < 4; ++a)
                  Q[b][a] = lds64(up + uint32_t(a + 4 * b) * 128u + (col ^ (uint32_t((a + 4 * b) & 7) << 4)));
            }
        };
        // Two plane pairs (own plane it, other plane it + 2) per rolled iteration: both planes are requested from shared
        // memory at once (two exposed load latencies per cell instead of four), and the in-plane chains start from the
        // cross-plane term of the own plane instead of a separate multiply.
#pragma unroll 1
        for (int it = 0; it < 2; ++it)
          {
            const int sp = 2 * h + it, so = (sp + 2) & 3;
            double    P[4][4], Q[4][4]; // [b][a]
            load_plane(P, sp);
            load_plane(Q, so);
            const double q0 = rc.C[(2 * h) * 4 + so], q1 = rc.C[(2 * h + 1) * 4 + so];
            const double cown = rc.C[sp * 4 + sp], coth = rc.C[(2 * h + 1 - it) * 4 + sp];
            double       fa[4] = {0.0, 0.0, 0.0, 0.0}, fb[4] = {0.0, 0.0, 0.0, 0.0};
            if (role == 0)
              {
                if (actA)
                  {
                    if (from_t0)
                      {
                        const double2 v0 = lds128(t0b + sp * 2048), v1 = lds128(t0b + sp * 2048 + 16);
                        fa[0] = v0.x;
                        fa[1] = v0.y;
                        fa[2] = v1.x;
                        fa[3] = v1.y;
                      }
                    else if (it == 0)
                      {
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                          fa[b] = tr[0][b];
                      }
                    else
                      {
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                          fa[b] = tr[1][b];
                      }
                    if (it == 0)
                      {
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                          tr[0][b] = descend ? P[b][0] : P[b][3];
                      }
                    else
                      {
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                          tr[1][b] = descend ? P[b][0] : P[b][3];
                      }
                  }
                if (actB)
                  {
                    const uint32_t r32 = uint32_t(t) + 64u * uint32_t(sp);
                    const uint32_t fl  = (r32 >> 2) & 1u;
                    const uint32_t tb  = fbuf + r32 * 32u;
                    const double2  v0 = lds128(tb + ((0u ^ fl) << 4)), v1 = lds128(tb + ((1u ^ fl) << 4));
                    fb[0] = v0.x;
                    fb[1] = v0.y;
                    fb[2] = v1.x;
                    fb[3] = v1.y;
                  }
              }
            else
              {
                if (actA)
                  {
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                      fa[b] = lds64(fcol + 128u * uint32_t(b) + 512u * uint32_t(sp));
                  }
                if (actB)
                  {
#pragma unroll
                    for (int a = 0; a < 4; ++a)
                      fb[a] = lds64(fcol + F_BYTES + 128u * uint32_t(a) + 512u * uint32_t(sp));
                  }
              }
            // cross-plane sweep of the other plane into both output planes
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
              for (int a = 0; a < 4; ++a)
                {
                  acc[0][b][a] = fma(q0, Q[b][a], acc[0][b][a]);
                  acc[1][b][a] = fma(q1, Q[b][a], acc[1][b][a]);
                }
            // all 16 in-plane chains of the plane are independent instruction streams (the 216-register budget pays
            // for the 16 temporaries); one uniform branch per plane keeps the accumulator indices static
            double q[4][4];
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
              for (int a = 0; a < 4; ++a)
                q[b][a] = cown * P[b][a];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int a = 0; a < 4; ++a)
                  q[b][a] = fma(rc.A[a * 4 + j], P[b][j], q[b][a]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int a = 0; a < 4; ++a)
                  q[b][a] = fma(rc.B[b * 4 + j], P[j][a], q[b][a]);
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
              for (int a = 0; a < 4; ++a)
                q[b][a] = fma(rc.LB[b], fb[a], fma(rc.LA[a], fa[b], q[b][a]));
            if (it == 1)
              {
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                  for (int a = 0; a < 4; ++a)
                    {
                      acc[1][b][a] += q[b][a];
                      acc[0][b][a] = fma(coth, P[b][a], acc[0][b][a]);
                    }
              }
            else
              {
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                  for (int a = 0; a < 4; ++a)
                    {
                      acc[0][b][a] += q[b][a];
                      acc[1][b][a] = fma(coth, P[b][a], acc[1][b][a]);
                    }
              }
          }
        if (role != 0)
          {
            // in-plane faces of directions 2 and 3 are done: let the face producer refill this slot
            if (actA)
              release(bars.r2fEmpty(f, 0));
            if (actB)
              release(bars.r2fEmpty(f, 1));
          }

        if (FUSED && role == 1)
          {
#pragma unroll
            for (int q = 0; q < 16; ++q)
              svB[q] = ldg_f64_here(p.sol + gfu + (q + 16) * 16);
          }
        // ---- the cell stage is free (both rounds read it at the same time), then the face of direction C
        release(bars.emptyU(s));
        if (role == 0)
          {
            if (from_t0 && act0)
              release(bars.t0Empty());
          }
        else
          {
            if (actC)
              mbar_wait(bars.r2fFull(f, 2), pf);
          }
        if (actC)
          {
#pragma unroll
            for (int b = 0; b < 4; ++b)
              {
                double fc[4];
                if (role == 0)
                  {
                    const uint32_t tb = fbuf + F_BYTES + rowU;
                    const double2  v0 = lds128(tb + ((uint32_t(2 * b) ^ sw) << 4)), v1 = lds128(tb + ((uint32_t(2 * b + 1) ^ sw) << 4));
                    fc[0] = v0.x;
                    fc[1] = v0.y;
                    fc[2] = v1.x;
                    fc[3] = v1.y;
                  }
                else
                  {
#pragma unroll
                    for (int a = 0; a < 4; ++a)
                      fc[a] = lds64(fcol + 2 * F_BYTES + 128u * uint32_t(a + 4 * b));
                  }
#pragma unroll
                for (int a = 0; a < 4; ++a)
                  {
                    acc[0][b][a] = fma(lc0, fc[a], acc[0][b][a]);
                    acc[1][b][a] = fma(lc1, fc[a], acc[1][b][a]);
                  }
              }
          }

        // ---- epilogue
        if (role == 0)
          {
            if (r1faces)
              release(bars.r1fEmpty(f));
            // partial sums -> shared (same swizzle as u), once round 2 has taken the previous cell's
            mbar_wait(bars.accEmpty(), pk ^ 1u);
#pragma unroll
            for (int cl = 0; cl < 2; ++cl)
#pragma unroll
              for (int ch = 0; ch < 8; ++ch)
                sts128(ab + cl * 8192 + ((uint32_t(ch) ^ sw) << 4), acc[cl][ch >> 1][(ch & 1) * 2], acc[cl][ch >> 1][(ch & 1) * 2 + 1]);
            release(bars.accFull());
          }
        else
          {
            if (actC)
              release(bars.r2fEmpty(f, 2));
            // add round 1's sums
            mbar_wait(bars.accFull(), pk);
#pragma unroll
            for (int cl = 0; cl < 2; ++cl)
#pragma unroll
              for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int a = 0; a < 4; ++a)
                  acc[cl][b][a] += lds64(ab + uint32_t(a + 4 * b + 16 * cl) * 128u + (col ^ (uint32_t((a + 4 * b) & 7) << 4)));
            release(bars.accEmpty());
            // coalesced stores (a half-warp writes 128 contiguous bytes)
            const long long g = (long long)cellid * CELL + cc + 1024 * i5 + 512 * h;
            if (FUSED)
              {
                double *      solw = p.sol + g;
                double *      tiw  = p.ti_next + g;
#pragma unroll
                for (int cl = 0; cl < 2; ++cl)
                  {
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                      {
                        const double kv = acc[cl][q >> 2][q & 3];
                        const double sv = cl == 0 ? svA[q] : svB[q];
                        if (stream)
                          {
                            __stcs(solw + (q + 16 * cl) * 16, fma(p.fb, kv, sv));
                            if (p.fa != 0.0)
                              __stcs(tiw + (q + 16 * cl) * 16, fma(p.fa, kv, sv));
                          }
                        else
                          {
                            solw[(q + 16 * cl) * 16] = fma(p.fb, kv, sv);
                            if (p.fa != 0.0)
                              tiw[(q + 16 * cl) * 16] = fma(p.fa, kv, sv);
                          }
                      }
                  }
              }
            else
              {
                double *out = p.dst + g;
#pragma unroll
                for (int cl = 0; cl < 2; ++cl)
#pragma unroll
                  for (int b = 0; b < 4; ++b)
#pragma unroll
                    for (int a = 0; a < 4; ++a)
                      {
                        if (stream)
                          __stcs(out + (a + 4 * b + 16 * cl) * 16, acc[cl][b][a]);
                        else
                          out[(a + 4 * b + 16 * cl) * 16] = acc[cl][b][a];
                      }
              }
          }
      }
  }

  // ======================================================================= halo senders (fused-halo variant)
  // pack loop of export_to_ghosted_array_start (matrix_free/vector_partitioner.h:1443-1460) fused with the transport:
  // the nodal face layers go straight into the neighbour GPU's ghost segment over NVLink (peer-mapped pointers), then
  // the neighbour's arrival counter is bumped.  The first n_sender_ctas CTAs do this with ALL their warps before they
  // take their first row of cells (rows are handed out dynamically, so the late start balances itself): a warp moves
  // one face cell (8 KiB) per iteration with 16 x 16 B loads in flight per lane, i.e. 80 KiB in flight per sender CTA.
  // Kept out of line so that the compute warps' code is laid out and register-allocated exactly as in the plain variant.
  template <int NT>
  __device__ __noinline__ void
  halo_send_cta(const FastParams &p)
  {
    const int lane = threadIdx.x & 31;
    const int gw   = blockIdx.x * (NT / 32) + (threadIdx.x >> 5); // this warp among all sender warps
    const int nw   = p.n_sender_ctas * (NT / 32);
    for (int si = 0; si < p.n_sends; ++si)
      {
        const int d = p.send_dir[si], side = p.send_side[si];
        int       nfc = 1;
#pragma unroll
        for (int e = 0; e < 6; ++e)
          nfc *= (e == d) ? 1 : p.ncell[e];
        const int stride_d  = 1 << (2 * d);
        const int layer_off = (side ? 3 : 0) * stride_d;
        for (int fc = gw; fc < nfc; fc += nw)
          {
            long long cell = 0, m = 1;
            int       r    = fc;
#pragma unroll
            for (int e = 0; e < 6; ++e)
              {
                int ce;
                if (e == d)
                  ce = side ? p.ncell[e] - 1 : 0;
                else
                  {
                    ce = r % p.ncell[e];
                    r /= p.ncell[e];
                  }
                cell += ce * m;
                m *= p.ncell[e];
              }
            const double *sp = p.src + cell * CELL + layer_off;
            double *      op = p.send_dst[si] + (long long)fc * 1024;
            if (d == 0)
              {
                // layer of direction 0: single values, 32 bytes apart; lane takes values 2*lane, 2*lane+1 of 64-value groups
                double2 v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u)
                  {
                    const int i = 2 * (lane + 32 * u);
                    v[u].x      = __ldg(sp + 4 * i);
                    v[u].y      = __ldg(sp + 4 * i + 4);
                  }
#pragma unroll
                for (int u = 0; u < 16; ++u)
                  *reinterpret_cast<double2 *>(op + 2 * (lane + 32 * u)) = v[u];
              }
            else
              {
                // 16-byte chunks; the layer is contiguous over 4^d >= 4 values
                double2 v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u)
                  {
                    const int i = 2 * (lane + 32 * u), hi = i >> (2 * d), lo = i & (stride_d - 1);
                    v[u]        = __ldg(reinterpret_cast<const double2 *>(sp + hi * 4 * stride_d + lo));
                  }
#pragma unroll
                for (int u = 0; u < 16; ++u)
                  *reinterpret_cast<double2 *>(op + 2 * (lane + 32 * u)) = v[u];
              }
          }
      }
    __threadfence_system();
    __syncthreads();
    if (int(threadIdx.x) < p.n_sends)
      asm volatile("red.release.sys.global.add.s32 [%0], 1;" ::"l"(p.send_flag[threadIdx.x]) : "memory");
  }

  template <bool FUSED, bool HALO>
  __global__ void __launch_bounds__(THREADS, 1)
    k_advect_3d3v_k3(const __grid_constant__ CUtensorMap mapU, const __grid_constant__ CUtensorMap mapT1, const __grid_constant__ CUtensorMap mapT2,
                     const __grid_constant__ CUtensorMap mapT3, const __grid_constant__ CUtensorMap mapT4, const __grid_constant__ CUtensorMap mapG1,
                     const __grid_constant__ CUtensorMap mapG5, const __grid_constant__ CUtensorMap mapU16, const __grid_constant__ FastParams p,
                     const __grid_constant__ FastCoef cf)
  {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw  = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char *gbase = smem_raw + (base - raw);
    const Bars     bars{base + BAR_OFF};

    const int tid  = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0)
      {
        for (int s = 0; s < STAGES; ++s)
          {
            mbar_init(bars.fullU(s), 1);
            mbar_init(bars.emptyU(s), 9); // 8 compute warps + the face producer (it reads the CellInfo)
            mbar_init(bars.infoFull(s), 1);
          }
        for (int f = 0; f < 2; ++f)
          {
            mbar_init(bars.r1fFull(f), 1);
            mbar_init(bars.r1fEmpty(f), 4);
            for (int j = 0; j < 3; ++j)
              {
                mbar_init(bars.r2fFull(f, j), 1);
                mbar_init(bars.r2fEmpty(f, j), 4);
              }
          }
        mbar_init(bars.accFull(), 4);
        mbar_init(bars.accEmpty(), 4);
        mbar_init(bars.t0Full(), 32);
        mbar_init(bars.t0Empty(), 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
    __syncthreads();

    if (HALO)
      {
        if (p.pass == 3 && p.n_sends > 0 && int(blockIdx.x) < p.n_sender_ctas) // (CTA-uniform)
          halo_send_cta<THREADS>(p);
      }

    const int  n0      = p.ncell[0];
    const bool act0    = p.up_delta[0] != 0;
    const bool act1    = p.up_delta[1] != 0;
    const bool act5    = p.up_delta[5] != 0;
    const bool r1faces = act1 || act5;
    const bool descend = p.up_delta[0] > 0; // upwind neighbour is the upper cell: walk downwards

    if (warp < 8)
      {
        // compute warpgroups: take the registers the producer warpgroup hands back
        asm volatile("setmaxnreg.inc.sync.aligned.u32 " HD_STR(HD_REGS_COMPUTE) ";");
        if (warp < 4)
          compute_round<0, FUSED>(p, cf, base, gbase, bars, tid);
        else
          compute_round<1, FUSED>(p, cf, base, gbase, bars, tid - 128);
        return;
      }
    asm volatile("setmaxnreg.dec.sync.aligned.u32 " HD_STR(HD_REGS_PRODUCER) ";");
    if (warp == 8)
      {
        // ======================================================================= cell producer
        if (lane == 0)
          {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapU));
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapT1));
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapU16));
          }
        const uint32_t f_bytes  = (act1 ? F_BYTES : 0) + (act5 ? F_BYTES : 0);
        const bool     keep4    = HD_HINTS && (p.hints & 1) != 0 && p.up_delta[4] != 0;
        const int      keep_row = p.up_delta[4] < 0 ? 48 : 0; // rows (i2,i3,i4) of a piece with i4 = 3 resp. 0
        const bool     first5   = HD_HINTS && (p.hints & 2) != 0;
        const uint64_t pol_last = policy_evict_last(), pol_first = policy_evict_first();
        const bool     ghost0   = act0 && p.up_kind[0] == HD_SIDE_GHOST;
        int            k        = 0; // cell sequence number of this CTA
        int            nrow_seq = 0; // row sequence number of this CTA
        // Rows come from a global counter in lattice order, so the CTAs sweep the lattice as one compact window.
        // A row is walked in upwind order over the steps [sb, se).  With ghost faces the work is split into an
        // interior pass that needs no ghost data (overlapped with the halo exchange) and a boundary pass, the
        // reference's overlapping levels (matrix_free.templates.h:1516-1566): a row whose directions 1..5 need no
        // ghosts is interior, except for its upwind-most cell if direction 0 is cut.
        bool halo_ready = p.pass != 3;
        // interior rows: mixed-radix decode that leaves out the ghost layer of every cut direction
        auto decode_interior = [&](int i, int (&cr)[6]) {
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
        auto decode_boundary = [&](int i, int (&cr)[6]) {
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
        // The work items come from a global counter; the next one is requested while the current one is being loaded,
        // so that the atomic's round trip (microseconds on a busy memory system) never stalls the ring.
        int  next_item  = 0;
        if (lane == 0)
          next_item = atomicAdd(p.counters, 1);
        auto fetch_row  = [&](int (&cr)[6], int &sb, int &se) -> bool {
          for (;;)
            {
              int item = 0;
              if (lane == 0)
                {
                  item = next_item;
                  if (item < p.n_items)
                    next_item = atomicAdd(p.counters, 1);
                }
              item = __shfl_sync(0xffffffffu, item, 0);
              if (item >= p.n_items)
                return false;
              cr[0] = 0;
              sb    = 0;
              se    = n0;
              int mode = p.pass; // 0: all cells of a lattice row, 1: interior list, 2: boundary list
              if (p.pass == 3)
                {
                  // one launch, two phases: interior cells, then (once the halo has arrived) the rest
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
                  // a row whose directions 1..5 need no ghosts; its upwind-most cell is left out if direction 0 is cut
                  decode_interior(item, cr);
                  if (ghost0)
                    sb = 1;
                }
              else if (item < p.n_bnd)
                decode_boundary(item, cr);
              else
                {
                  // the upwind-most cells of the interior rows (direction 0 cut)
                  decode_interior(item - p.n_bnd, cr);
                  se = 1;
                }
              if (sb >= se)
                continue;
              if (mode == 2 && !halo_ready)
                {
                  // The ghost faces are written by the neighbour GPUs while this kernel runs; the host enqueues a
                  // flag write behind them.  Give up after 4 s (error word) rather than hang the GPU.
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
                      asm volatile("fence.proxy.async;" ::: "memory"); // the TMA (async proxy) reads of the ghosts come after
                    }
                  __syncwarp();
                  halo_ready = true;
                }
              return true;
            }
        };
        int c[6], sb = 0, se = 0;
        while (fetch_row(c, sb, se))
          {
            // direction-0 trace of the upwind neighbour of the first cell of the walk (asynchronous gather)
            if (act0)
              {
                c[0] = descend ? n0 - 1 - sb : sb;
                mbar_wait(bars.t0Empty(), uint32_t(nrow_seq & 1) ^ 1u);
                const uint32_t t0 = base + T0_OFF;
                if (needs_ghost(p, c, 0))
                  {
                    const double *g = p.ghost + p.ghost_off[0] + face_cell(p, c, 0) * 1024;
#pragma unroll 8
                    for (int j = 0; j < 32; ++j)
                      cp_async_8(t0 + 8u * (lane + 32 * j), g + lane + 32 * j);
                  }
                else
                  {
                    const long long nb = upwind_cell(p, c, 0);
                    const double *  g  = p.src + nb * CELL + (p.up_delta[0] < 0 ? 3 : 0);
#pragma unroll 8
                    for (int j = 0; j < 32; ++j)
                      cp_async_8(t0 + 8u * (lane + 32 * j), g + 4 * (lane + 32 * j));
                  }
                cp_async_arrive_noinc(bars.t0Full());
              }
            for (int step = sb; step < se; ++step, ++k)
              {
                c[0]                 = descend ? n0 - 1 - step : step;
                const long long cell = cell_index(p, c);
                const int       s    = k % STAGES;
                mbar_wait(bars.emptyU(s), uint32_t((k / STAGES) & 1) ^ 1u);
                if (lane == 0)
                  {
                    int *info = reinterpret_cast<int *>(gbase + INFO_OFF + 32 * s);
                    info[0]   = int(cell);
#pragma unroll
                    for (int d = 0; d < 6; ++d)
                      info[1 + d] = c[d];
                    info[7]             = (step == sb) ? 1 : 0;
                    mbar_arrive(bars.infoFull(s)); // (release: the face producer may start on this cell's faces now)
                    const uint32_t dstU = base + s * U_BYTES;
                    mbar_expect_tx(bars.fullU(s), U_BYTES);
                    if (keep4)
                      {
                        // the direction-4 outflow layer (16 of the 64 rows of every i5 piece) stays in L2 for the downwind
                        // neighbour, which comes ncell[1]*ncell[2]*ncell[3] rows later; everything else is normal
#pragma unroll
                        for (int piece = 0; piece < 4; ++piece)
                          {
                            const int r0 = int(cell * 256 + piece * 64);
                            tma_load_2d_hint(dstU + piece * 8192 + keep_row * 128, &mapU16, 0, r0 + keep_row, bars.fullU(s), pol_last);
#pragma unroll
                            for (int g = 0; g < 3; ++g)
                              {
                                const int rr = (keep_row == 0 ? 16 : 0) + 16 * g;
                                tma_load_2d(dstU + piece * 8192 + rr * 128, &mapU16, 0, r0 + rr, bars.fullU(s));
                              }
                          }
                      }
                    else
                      {
#pragma unroll
                        for (int piece = 0; piece < 4; ++piece)
                          tma_load_2d(dstU + piece * 8192, &mapU, 0, int(cell * 256 + piece * 64), bars.fullU(s));
                      }
                  }
                if (r1faces)
                  {
                    const int f = k & 1;
                    mbar_wait(bars.r1fEmpty(f), uint32_t((k >> 1) & 1) ^ 1u);
                    if (lane == 0)
                      {
                        const uint32_t dstF = base + R1F_OFF + f * 2 * F_BYTES;
                        mbar_expect_tx(bars.r1fFull(f), f_bytes);
                        if (act1)
                          {
                            if (needs_ghost(p, c, 1)) // ghost segment viewed as rows of 4 doubles (same 32 B swizzle)
                              tma_load_2d(dstF, &mapG1, 0, int((p.ghost_off[1] + face_cell(p, c, 1) * 1024) >> 2), bars.r1fFull(f));
                            else
                              tma_load_3d(dstF, &mapT1, 0, p.up_delta[1] < 0 ? 3 : 0, int(upwind_cell(p, c, 1) * 256), bars.r1fFull(f));
                          }
                        if (act5)
                          {
                            if (needs_ghost(p, c, 5)) // ghost segment viewed as rows of 16 doubles (128 B swizzle)
                              tma_load_2d(dstF + F_BYTES, &mapG5, 0, int((p.ghost_off[5] + face_cell(p, c, 5) * 1024) >> 4), bars.r1fFull(f));
                            else
                              {
                                const int r5 = int(upwind_cell(p, c, 5) * 256 + (p.up_delta[5] < 0 ? 192 : 0));
                                if (first5)
                                  tma_load_2d_hint(dstF + F_BYTES, &mapU, 0, r5, bars.r1fFull(f), pol_first);
                                else
                                  tma_load_2d(dstF + F_BYTES, &mapU, 0, r5, bars.r1fFull(f));
                              }
                          }
                      }
                  }
              }
            ++nrow_seq;
          }
        // end marker
        {
          const int s = k % STAGES;
          mbar_wait(bars.emptyU(s), uint32_t((k / STAGES) & 1) ^ 1u);
          if (lane == 0)
            {
              int *info = reinterpret_cast<int *>(gbase + INFO_OFF + 32 * s);
              info[0]   = -1;
              mbar_arrive(bars.infoFull(s));
              mbar_arrive(bars.fullU(s));
              // the last CTA to finish re-arms the row counter for the next launch
              __threadfence();
              const int done = atomicAdd(p.counters + 1, 1);
              if (done == int(gridDim.x) - 1)
                {
                  p.counters[0] = 0;
                  p.counters[1] = 0;
                  __threadfence();
                }
            }
        }
      }
    else if (warp == 9)
      {
        // ======================================================================= face producer for round 2
        if (lane == 0)
          {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapT2));
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapT3));
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapT4));
            const bool     first4    = HD_HINTS && (p.hints & 2) != 0;
            const uint64_t pol_first = policy_evict_first();
            for (int k = 0;; ++k)
              {
                const int s = k % STAGES;
                const int f = k & 1;
                mbar_wait(bars.infoFull(s), uint32_t((k / STAGES) & 1));
                const volatile int *info = reinterpret_cast<const volatile int *>(gbase + INFO_OFF + 32 * s);
                if (info[0] < 0)
                  break;
                int c[6];
#pragma unroll
                for (int d = 0; d < 6; ++d)
                  c[d] = info[1 + d];
                mbar_arrive(bars.emptyU(s)); // the CellInfo is in registers
                if (p.pass == 3)
                  asm volatile("fence.proxy.async;" ::: "memory"); // ghost data written by peers, ordered by the producer's acquire
#pragma unroll
                for (int j = 2; j >= 0; --j)
                  {
                    const int d = 2 + j;
                    if (p.up_delta[d] == 0)
                      continue;
                    mbar_wait(bars.r2fEmpty(f, j), uint32_t((k >> 1) & 1) ^ 1u);
                    const uint32_t dstF = base + R2F_OFF + (3 * f + j) * F_BYTES;
                    mbar_expect_tx(bars.r2fFull(f, j), F_BYTES);
                    if (needs_ghost(p, c, d))
                      bulk_load_1d(dstF, p.ghost + p.ghost_off[d] + face_cell(p, c, d) * 1024, F_BYTES, bars.r2fFull(f, j));
                    else
                      {
                        const long long nb    = upwind_cell(p, c, d);
                        const int       layer = p.up_delta[d] < 0 ? 3 : 0;
                        if (d == 2)
                          tma_load_3d(dstF, &mapT2, 0, layer, int(nb * 64), bars.r2fFull(f, j));
                        else if (d == 3)
                          tma_load_3d(dstF, &mapT3, 0, layer, int(nb * 16), bars.r2fFull(f, j));
                        else if (first4)
                          tma_load_3d_hint(dstF, &mapT4, 0, layer, int(nb * 4), bars.r2fFull(f, j), pol_first);
                        else
                          tma_load_3d(dstF, &mapT4, 0, layer, int(nb * 4), bars.r2fFull(f, j));
                      }
                  }
              }
          }
      }
  }

#include "kernel_rounds6d.cuh"

  // ------------------------------------------------------------------------------- host side
  typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

  struct Maps
  {
    CUtensorMap u, u16, u256, t1, t2, t3, t4;
  };
  struct GhostMaps
  {
    CUtensorMap g1, g5;
  };

  struct FastState
  {
    EncodeTiledFn                encode = nullptr;
    std::map<const void *, Maps> cache;
    std::map<const void *, GhostMaps> ghost_cache;
    bool                         attr_set[8] = {false, false, false, false, false, false, false, false};
    int *                        d_counters  = nullptr;
    int *                        d_defer     = nullptr; // deferred-row queue of the fused-halo pass (three-round kernel)
    int                          defer_rows  = 0;
  };

  int
  encode_face_map(FastState *st, CUtensorMap *m, const void *src, cuuint64_t ncells, int d)
  {
    // face layer of direction d: view src as [hi = 4^(5-d) * ncells][4][lo = 4^d], box = (lo, 1, 4^(5-d))
    const cuuint64_t lo = 1ull << (2 * d), hi = 1ull << (2 * (5 - d));
    cuuint64_t       gdim[3] = {lo, 4, hi * ncells};
    cuuint64_t       gstr[2] = {lo * 8, lo * 32};
    cuuint32_t       box[3]  = {(cuuint32_t)lo, 1, (cuuint32_t)hi};
    cuuint32_t       estr[3] = {1, 1, 1};
    CUresult         r = st->encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(src), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            d == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return hd::fail(HD_ERR_CUDA, "cuTensorMapEncodeTiled(face " + std::to_string(d) + ") failed with code " + std::to_string((int)r));
    return HD_OK;
  }

  int
  get_state(hd_advection *op, FastState **out)
  {
    FastState *st = static_cast<FastState *>(op->fast_state);
    if (!st)
      {
        st             = new FastState;
        op->fast_state = st;
      }
    if (!st->encode)
      {
        void *                          fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        HD_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess)
          return hd::fail(HD_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        st->encode = reinterpret_cast<EncodeTiledFn>(fn);
      }
    if (!st->d_counters)
      {
        HD_CUDA(cudaMalloc(&st->d_counters, 8 * sizeof(int)));
        HD_CUDA(cudaMemset(st->d_counters, 0, 8 * sizeof(int)));
      }
    *out = st;
    return HD_OK;
  }

  int
  get_maps(hd_advection *op, FastState *st, const void *src, Maps **out)
  {
    auto it = st->cache.find(src);
    if (it == st->cache.end())
      {
        if (st->cache.size() > 64)
          st->cache.clear();
        Maps             m;
        const hd_mesh *  mesh   = op->mesh;
        const cuuint64_t ncells = (cuuint64_t)mesh->ncells;
        {
          // cell data as rows of 16 doubles (one (i0,i1) plane), 128B swizzle, boxes of 64 rows
          cuuint64_t gdim[2] = {16, ncells * 256};
          cuuint64_t gstr[1] = {128};
          cuuint32_t box[2]  = {16, 64};
          cuuint32_t estr[2] = {1, 1};
          CUresult   r = st->encode(&m.u, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(src), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r != CUDA_SUCCESS)
            return hd::fail(HD_ERR_CUDA, "cuTensorMapEncodeTiled(u) failed with code " + std::to_string((int)r));
          // the same view in boxes of 16 rows (one i4 layer of an i5 piece), for loads with per-layer L2 hints
          box[1] = 16;
          r      = st->encode(&m.u16, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(src), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r != CUDA_SUCCESS)
            return hd::fail(HD_ERR_CUDA, "cuTensorMapEncodeTiled(u16) failed with code " + std::to_string((int)r));
          // and in boxes of 256 rows = one whole cell per TMA instruction (three-round kernel)
          box[1] = 256;
          r      = st->encode(&m.u256, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(src), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r != CUDA_SUCCESS)
            return hd::fail(HD_ERR_CUDA, "cuTensorMapEncodeTiled(u256) failed with code " + std::to_string((int)r));
        }
        int rc;
        if ((rc = encode_face_map(st, &m.t1, src, ncells, 1)) != HD_OK)
          return rc;
        if ((rc = encode_face_map(st, &m.t2, src, ncells, 2)) != HD_OK)
          return rc;
        if ((rc = encode_face_map(st, &m.t3, src, ncells, 3)) != HD_OK)
          return rc;
        if ((rc = encode_face_map(st, &m.t4, src, ncells, 4)) != HD_OK)
          return rc;
        it = st->cache.emplace(src, m).first;
      }
    *out = &it->second;
    return HD_OK;
  }

  // ghost buffer viewed as rows of 4 doubles (direction-1 faces, 32 B swizzle like mapT1) and as rows of 16 doubles
  // (direction-5 faces, 128 B swizzle like mapU); a ghost face of one cell is 1024 contiguous doubles
  int
  get_ghost_maps(hd_advection *op, FastState *st, const void *ghosts, GhostMaps **out)
  {
    auto it = st->ghost_cache.find(ghosts);
    if (it == st->ghost_cache.end())
      {
        if (st->ghost_cache.size() > 64)
          st->ghost_cache.clear();
        GhostMaps m;
        std::memset(&m, 0, sizeof(m));
        const cuuint64_t total = (cuuint64_t)op->mesh->ghost_total;
        if (ghosts && total > 0)
          {
            cuuint32_t estr[2] = {1, 1};
            {
              cuuint64_t gdim[2] = {4, total / 4};
              cuuint64_t gstr[1] = {32};
              cuuint32_t box[2]  = {4, 256};
              CUresult   r = st->encode(&m.g1, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(ghosts), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
              if (r != CUDA_SUCCESS)
                return hd::fail(HD_ERR_CUDA, "cuTensorMapEncodeTiled(ghost 1) failed with code " + std::to_string((int)r));
            }
            {
              cuuint64_t gdim[2] = {16, total / 16};
              cuuint64_t gstr[1] = {128};
              cuuint32_t box[2]  = {16, 64};
              CUresult   r = st->encode(&m.g5, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(ghosts), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
              if (r != CUDA_SUCCESS)
                return hd::fail(HD_ERR_CUDA, "cuTensorMapEncodeTiled(ghost 5) failed with code " + std::to_string((int)r));
            }
          }
        it = st->ghost_cache.emplace(ghosts, m).first;
      }
    *out = &it->second;
    return HD_OK;
  }
} // namespace

namespace hd
{
  // CTAs that pack and send the halo in the fused-halo variant (and signal the neighbours' arrival counters)
  int
  fast6d_halo_senders(const hd_advection *op)
  {
    static const int env = [] {
      const char *e = getenv("HD_HALO_SENDERS");
      return e ? atoi(e) : 32;
    }();
    const hd_mesh *m     = op->mesh;
    long long      nrows = m->ncells / m->d.n_cells[0];
    long long      grid  = nrows < m->ctx->sm_count ? nrows : m->ctx->sm_count;
    int            want  = op->halo_senders > 0 ? op->halo_senders : env;
    if (want < 1)
      want = 1;
    return (int)(want < grid ? want : grid);
  }

  // which of the two 3D3V degree-3 FP64 kernels runs: hd_advection_set_kernel(op, 2) = the two-role pipelined kernel,
  // 6 = the three-round kernel; automatic choice: HD_FAST_VARIANT=pipe|rounds, default rounds
  bool
  fast6d_use_rounds(const hd_advection *op)
  {
    static const int env = [] {
      const char *e = getenv("HD_FAST_VARIANT");
      return (e && std::strcmp(e, "pipe") == 0) ? 0 : 1;
    }();
    if (op->kernel_choice == 6)
      return true;
    if (op->kernel_choice == 2)
      return false;
    return env != 0;
  }

  bool
  fast6d_supported(const hd_advection *op)
  {
    const hd_mesh *m = op->mesh;
    if (m->d.dim_x != 3 || m->d.dim_v != 3 || m->n != 4 || m->d.number_type != HD_F64)
      return false;
    if (m->ncells * 256 >= (1ll << 31))
      return false;
    for (int d = 0; d < 6; ++d)
      for (int s = 0; s < 2; ++s)
        {
          const int kind = m->d.side_kind[d][s];
          if (kind == HD_SIDE_DIRICHLET || kind == HD_SIDE_DIRICHLET_HOM)
            return false;
        }
    return true;
  }

  int
  launch_fast6d(hd_advection *op, void *dst, const void *src, const void *ghosts, double, const FusedUpdate &fu, int part, const hd_halo_send *sends,
                int n_sends, const void *halo_flag, int halo_target, long long row_begin, long long row_end)
  {
    hd_mesh *  m = op->mesh;
    FastState *st;
    int        rc = get_state(op, &st);
    if (rc != HD_OK)
      return rc;
    Maps *maps;
    rc = get_maps(op, st, src, &maps);
    if (rc != HD_OK)
      return rc;
    GhostMaps *gmaps;
    if (ghosts && (reinterpret_cast<uintptr_t>(ghosts) & 127) != 0)
      return hd::fail(HD_ERR_INVALID, "the ghost buffer must be 128-byte aligned");
    rc = get_ghost_maps(op, st, ghosts, &gmaps);
    if (rc != HD_OK)
      return rc;
    FastCoef   cfh;
    FastParams p;
    p.src       = static_cast<const double *>(src);
    p.dst       = static_cast<double *>(dst);
    p.ghost     = static_cast<const double *>(ghosts);
    long long nrows = 1;
    for (int d = 0; d < 6; ++d)
      {
        p.ncell[d] = m->d.n_cells[d];
        if (d > 0)
          nrows *= p.ncell[d];
        const int role = (d >= 2 && d <= 4) ? 1 : 0;
        double *  Cd   = (d == 0 || d == 2) ? cfh.r[role].A : ((d == 1 || d == 3) ? cfh.r[role].B : cfh.r[role].C);
        double *  Ld   = (d == 0 || d == 2) ? cfh.r[role].LA : ((d == 1 || d == 3) ? cfh.r[role].LB : cfh.r[role].LC);
        for (int i = 0; i < 16; ++i)
          Cd[i] = op->hC[d][0][i];
        // upwind side: L0 (lower neighbour) is non-zero for a_d > 0, L1 (upper) for a_d < 0
        const bool lo = (op->nb_mask[d] & 1) != 0, hi = (op->nb_mask[d] & 2) != 0;
        p.up_delta[d]  = lo ? -1 : (hi ? +1 : 0);
        const int side = lo ? 0 : 1;
        p.up_kind[d]   = m->d.side_kind[d][side];
        p.ghost_off[d] = m->ghost_off[d][side];
        for (int i = 0; i < 4; ++i)
          Ld[i] = lo ? op->hL0[d][i] : (hi ? op->hL1[d][i] : 0.0);
      }
    p.nrows    = (int)nrows;
    if (row_end < 0)
      row_end = nrows;
    if (row_begin < 0 || row_begin > row_end || row_end > nrows || (part != 0 && (row_begin != 0 || row_end != nrows)))
      return hd::fail(HD_ERR_INVALID, "bad row range");
    p.row_begin = (int)row_begin;
    p.row_end   = (int)row_end;
    {
      // interior / boundary work lists (see FastParams)
      long long n_int = 1, n_bnd = 0;
      for (int d = 1; d < 6; ++d)
        {
          const bool cut = p.up_delta[d] != 0 && p.up_kind[d] == HD_SIDE_GHOST;
          p.cutg[d]      = cut ? (p.up_delta[d] < 0 ? 0 : p.ncell[d] - 1) : -1;
          n_int *= p.ncell[d] - (cut ? 1 : 0);
        }
      p.cutg[0] = p.bsize[0] = 0;
      for (int k = 1; k < 6; ++k)
        {
          long long sz = 0;
          if (p.cutg[k] >= 0)
            {
              sz = 1;
              for (int d = 1; d < 6; ++d)
                if (d != k)
                  sz *= p.ncell[d] - ((p.cutg[d] >= 0 && d < k) ? 1 : 0);
            }
          p.bsize[k] = (int)sz;
          n_bnd += sz;
        }
      const bool ghost0 = p.up_delta[0] != 0 && p.up_kind[0] == HD_SIDE_GHOST;
      p.n_int           = (int)n_int;
      p.n_bnd           = (int)n_bnd;
      const long long n_b_items = n_bnd + (ghost0 ? n_int : 0);
      const long long items     = part == 0 ? row_end - row_begin : (part == 1 ? n_int : (part == 2 ? n_b_items : n_int + n_b_items));
      p.n_items = (int)items;
      nrows     = items > 0 ? items : 1; // (grid size)
    }
    p.counters = st->d_counters;
    p.sol      = static_cast<double *>(fu.sol);
    p.ti_next  = static_cast<double *>(fu.ti_next);
    p.fb       = fu.fb;
    p.fa       = fu.fa;
    p.pass          = part; // 0 all cells, 1 interior (no ghost data needed), 2 boundary layer, 3 both with an in-kernel wait
    p.halo_flag     = static_cast<const int *>(halo_flag);
    p.halo_target   = halo_target;
    p.n_sends       = 0;
    for (int i = 0; i < 6; ++i)
      {
        p.send_dir[i] = p.send_side[i] = 0;
        p.send_dst[i]  = nullptr;
        p.send_flag[i] = nullptr;
      }
    if (part == 3)
      {
        if (n_sends < 0 || n_sends > 6 || (n_sends > 0 && !sends))
          return hd::fail(HD_ERR_INVALID, "at most 6 halo sends per operator application");
        for (int i = 0; i < n_sends; ++i)
          {
            const hd_halo_send &h = sends[i];
            if (h.dir < 0 || h.dir >= 6 || h.side < 0 || h.side > 1 || !h.dst || !h.arrival_counter || (reinterpret_cast<uintptr_t>(h.dst) & 15))
              return hd::fail(HD_ERR_INVALID, "bad hd_halo_send entry");
            p.send_dir[i]  = h.dir;
            p.send_side[i] = h.side;
            p.send_dst[i]  = static_cast<double *>(h.dst);
            p.send_flag[i] = static_cast<int *>(h.arrival_counter);
          }
        p.n_sends = n_sends;
      }
    {
      static const int env_hints = [] {
        const char *e = getenv("HD_L2_HINTS");
        return e ? atoi(e) : 0; // measured on 8^6 cells: no gain from any combination (profiles/r01e_l2_hints.txt)
      }();
      p.hints = op->l2_hints >= 0 ? op->l2_hints : env_hints;
      p.r6_prefetch = 0;
      p.r6_trace    = nullptr;
      p.defer       = 0;
      p.defer_queue = nullptr;
    }
    {
      // row tiles (see FastParams::tile); HD_ROW_TILE="t1,t2,t3,t4,t5" overrides the default, 0 = full extent.
      // Launches on a row range (the pipelined host path walks slabs of the slowest directions) keep the lattice order.
      static const std::array<int, 5> env_tile = [] {
        std::array<int, 5> t = {{HD_DEFAULT_ROW_TILE}};
        if (const char *e = getenv("HD_ROW_TILE"))
          {
            int v[5] = {0, 0, 0, 0, 0};
            sscanf(e, "%d,%d,%d,%d,%d", v, v + 1, v + 2, v + 3, v + 4);
            for (int i = 0; i < 5; ++i)
              t[i] = v[i];
          }
        return t;
      }();
      const bool full = p.row_begin == 0 && p.row_end == p.nrows;
      p.tile[0]       = 1;
      for (int d = 1; d < 6; ++d)
        {
          int want = op->row_tile[d - 1] >= 0 ? op->row_tile[d - 1] : env_tile[d - 1];
          int t    = p.ncell[d];
          if (full && want > 0 && want < t)
            for (t = want; p.ncell[d] % t != 0; --t) // largest divisor of ncell[d] not above the request
              ;
          p.tile[d] = t;
        }
    }
    p.halo_mask     = 0;
    for (int d = 0; d < 6; ++d)
      for (int sd = 0; sd < 2; ++sd)
        if (m->d.side_kind[d][sd] == HD_SIDE_GHOST && ((op->nb_mask[d] >> sd) & 1))
          p.halo_mask |= 1u << (2 * d + sd);
    const bool halo = (part == 3 && n_sends > 0) || getenv("HD_FORCE_HALO_VARIANT") != nullptr; // (env: A/B of the two kernel variants)
    long long  grid = nrows < m->ctx->sm_count ? nrows : m->ctx->sm_count;
    p.n_sender_ctas = hd::fast6d_halo_senders(op);
    if (hd::fast6d_use_rounds(op))
      {
        // three-round kernel (kernel_rounds6d.cuh)
        static const int env_pf = [] {
          const char *e = getenv("HD_R6_PREFETCH");
          return e ? atoi(e) : 0x7e;
        }();
        p.r6_prefetch = env_pf;
        {
          static const int env_defer = [] {
            const char *e = getenv("HD_R6_DEFER");
            return e ? atoi(e) : 0; // default: interior list, then boundary list (1-GPU self exchange: 4.95 ms against 5.08 ms with the queue, profiles/r02_fused_halo_selftest.txt)
          }();
          const bool ghost0 = p.up_delta[0] != 0 && p.up_kind[0] == HD_SIDE_GHOST;
          if (part == 3 && env_defer && !ghost0)
            {
              if (st->defer_rows < p.nrows)
                {
                  cudaFree(st->d_defer);
                  st->d_defer = nullptr;
                  HD_CUDA(cudaMalloc(&st->d_defer, sizeof(int) * (size_t)p.nrows));
                  HD_CUDA(cudaMemsetAsync(st->d_defer, 0, sizeof(int) * (size_t)p.nrows, m->ctx->stream));
                  st->defer_rows = p.nrows;
                }
              p.defer       = 1;
              p.defer_queue = st->d_defer;
              p.n_items     = p.nrows;
              grid          = p.nrows < m->ctx->sm_count ? p.nrows : m->ctx->sm_count;
              // (rows are decoded in the tiled lattice order of pass 0: the tile extents were set above for a full launch)
              p.row_begin = 0;
            }
        }
#ifdef HD_R6_TRACE
        {
          // debugging aid (tools/r6_timeline.py): HD_R6_TRACE_FILE=<path> dumps the last launch's timeline of CTA 0
          static long long *d_trace = nullptr;
          if (!d_trace)
            HD_CUDA(cudaMalloc(&d_trace, sizeof(long long) * 13 * R6_TRACE_CELLS * 16));
          HD_CUDA(cudaMemsetAsync(d_trace, 0, sizeof(long long) * 13 * R6_TRACE_CELLS * 16, m->ctx->stream));
          p.r6_trace = d_trace;
        }
#endif
        r6::Coef rc6;
        for (int d = 0; d < 6; ++d)
          {
            double *   Cd = (d & 1) ? rc6.B[d / 2] : rc6.A[d / 2];
            double *   Ld = (d & 1) ? rc6.LB[d / 2] : rc6.LA[d / 2];
            const bool lo = (op->nb_mask[d] & 1) != 0, hi = (op->nb_mask[d] & 2) != 0;
            for (int i = 0; i < 16; ++i)
              Cd[i] = op->hC[d][0][i];
            for (int i = 0; i < 4; ++i)
              Ld[i] = lo ? op->hL0[d][i] : (hi ? op->hL1[d][i] : 0.0);
          }
        const int ridx = 4 + (fu.enabled ? 1 : 0) + (halo ? 2 : 0);
        auto      rk   = fu.enabled ? (halo ? k_rounds_3d3v_k3<true, true> : k_rounds_3d3v_k3<true, false>) :
                                      (halo ? k_rounds_3d3v_k3<false, true> : k_rounds_3d3v_k3<false, false>);
        if (!st->attr_set[ridx])
          {
            HD_CUDA(cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, R6_SMEM_BYTES));
            st->attr_set[ridx] = true;
          }
        rk<<<(unsigned)grid, R6_THREADS, R6_SMEM_BYTES, m->ctx->stream>>>(maps->u256, maps->t1, maps->t2, maps->t3, maps->t4, gmaps->g1, p, rc6);
        HD_CUDA(cudaGetLastError());
#ifdef HD_R6_TRACE
        if (const char *tf = getenv("HD_R6_TRACE_FILE"))
          {
            std::vector<long long> h(13 * R6_TRACE_CELLS * 16);
            HD_CUDA(cudaStreamSynchronize(m->ctx->stream));
            HD_CUDA(cudaMemcpy(h.data(), p.r6_trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
            if (FILE *f = fopen(tf, "wb"))
              {
                fwrite(h.data(), sizeof(long long), h.size(), f);
                fclose(f);
              }
          }
#endif
        op->launches++;
        op->last_kernel = fu.enabled ? "rounds_3d3v_k3_fused_lsrk" : "rounds_3d3v_k3";
        return HD_OK;
      }
    const int  fidx = (fu.enabled ? 1 : 0) + (halo ? 2 : 0);
    auto       kern = fu.enabled ? (halo ? k_advect_3d3v_k3<true, true> : k_advect_3d3v_k3<true, false>) : (halo ? k_advect_3d3v_k3<false, true> : k_advect_3d3v_k3<false, false>);
    if (!st->attr_set[fidx])
      {
        HD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        st->attr_set[fidx] = true;
      }
    kern<<<(unsigned)grid, THREADS, SMEM_BYTES, m->ctx->stream>>>(maps->u, maps->t1, maps->t2, maps->t3, maps->t4, gmaps->g1, gmaps->g5, maps->u16, p, cfh);
    HD_CUDA(cudaGetLastError());
    op->launches++;
    op->last_kernel = fu.enabled ? "advect_3d3v_k3_fused_lsrk" : "advect_3d3v_k3";
    return HD_OK;
  }

  int
  fast6d_overlap_status(hd_advection *op, int *timed_out)
  {
    FastState *st = static_cast<FastState *>(op->fast_state);
    *timed_out    = 0;
    if (st && st->d_counters)
      {
        HD_CUDA(cudaMemcpy(timed_out, st->d_counters + 2, sizeof(int), cudaMemcpyDeviceToHost));
        HD_CUDA(cudaMemset(st->d_counters + 2, 0, sizeof(int)));
      }
    return HD_OK;
  }

  void
  fast6d_release(hd_advection *op)
  {
    FastState *st = static_cast<FastState *>(op->fast_state);
    if (st)
      {
        cudaFree(st->d_counters);
        cudaFree(st->d_defer);
      }
    delete st;
    op->fast_state = nullptr;
  }
} // namespace hd
