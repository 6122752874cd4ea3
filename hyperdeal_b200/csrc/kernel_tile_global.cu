// Tile kernel for cells that do not fit into shared memory twice — first of all 3D3V degree 5 in FP32 (BASELINE.json
// configs[2]: 6^6 = 46 656 values = 182 KiB per cell), where the generic kernel reaches 10 % of the HBM roofline.
//
// Same rounds as the shared-memory tile kernel (kernels_generic.cu: two directions per round, an N x N tile per thread in
// registers), but the cell is not staged: a CTA owns one cell at a time and reads its tiles straight from global memory (the
// cell is re-read in every round and stays in L1/L2); a block-wide barrier separates the rounds.  The partial sums of the
// rounds live
//   * in shared memory (k_apply_tile_global_sp, the shipped variant for degree 5: one persistent CTA per SM, 182 KiB;
//     dst is written once, DRAM traffic 6.7 GB per launch on the configs[2] lattice, 141 GDoF/s), or
//   * in `dst` itself (k_apply_tile_global: written in round 0, read-modified-written in the later rounds; 16.8 GB of DRAM
//     traffic, 95 GDoF/s — the partial sums of all resident CTAs do not stay in L2; HD_TG_SMEM_PARTIALS=0, and degree 3).
// The fused LSRK update is applied in the last round.  Periodic and ghost sides (Dirichlet lattices arrive here as ghost
// sides, capi.cu: apply_dirichlet_as_ghosts), even dim_x + dim_v.
//
// The body is host/device code: tests/vp_emulation_harness.cpp runs it on the CPU against the oracle
// (tests/test_tile_global_emulation.py, both variants); GPU parity in tests/tile_global_check.py and tests/test_apply_gpu.py.
// Automatic choice for degree 5; hd_advection_set_kernel(op, 5) elsewhere.
#ifdef HD_VP_HOST_EMULATION
#  ifndef HD_MAX_DIM
#    include <cmath>
#    include <cstddef>
#    include <vector>
#    define HD_MAX_DIM 6
#  endif
#  ifndef HD_SIDE_GHOST_DEFINED
#    define HD_SIDE_GHOST_DEFINED
enum
{
  HD_SIDE_PERIODIC_LOCAL = 0,
  HD_SIDE_GHOST          = 1
};
#  endif
#  define HD_TG_FN inline
#  define HD_TG_SYNC() ((void)0)
#  define HD_TG_LDG(p) (*(p))
#  define HD_TG_UNROLL
#else
#  include "hd_internal.h"
#  define HD_TG_FN __device__
#  define HD_TG_SYNC() __syncthreads()
#  define HD_TG_LDG(p) __ldg(p)
#  define HD_TG_UNROLL _Pragma("unroll")
#endif

namespace
{
  template <typename T, int N>
  struct TgCoef
  {
    T C[HD_MAX_DIM][N * N]; // [direction][out * N + in]
    T L0[HD_MAX_DIM][N];    // lifting of the lower neighbour's trace
    T L1[HD_MAX_DIM][N];    // ... upper neighbour's
  };

  template <typename T>
  struct TgParams
  {
    const T * src;
    T *       dst;
    const T * ghost;
    int       dim;
    int       ncell[HD_MAX_DIM];
    int       side_kind[HD_MAX_DIM][2];
    int       nb_mask[HD_MAX_DIM];
    long long ghost_off[HD_MAX_DIM][2];
    long long ncells;
    T *       sol;
    T *       ti_next;
    T         fb, fa;
    int       fused;
  };

  template <typename T, int N>
  HD_TG_FN void
  tg_cell(const TgParams<T> &p, const TgCoef<T, N> &cf, const long long cell, const int tid, const int nthr, T *part = nullptr)
  {
    // part != nullptr: the partial sums of the rounds live in a CTA-private buffer of nd values (shared memory) instead
    // of dst; only the last round touches dst (or sol / Ti_next)
    const int dim = p.dim;
    int       nd  = 1; // (indices inside a cell fit 32 bits: N^6 = 46656 for degree 5)
    for (int d = 0; d < dim; ++d)
      nd *= N;
    const int nt = nd / (N * N); // tiles per round
    int             c[HD_MAX_DIM];
    long long       cstr[HD_MAX_DIM];
    {
      long long r = cell, m = 1;
      for (int d = 0; d < dim; ++d)
        {
          c[d]    = int(r % p.ncell[d]);
          r /= p.ncell[d];
          cstr[d] = m;
          m *= p.ncell[d];
        }
    }
    const T *uc = p.src + cell * (long long)nd;
    T *      oc = p.dst + cell * (long long)nd;
    T *      pc = part ? part : oc; // where the partial sums of the earlier rounds are
    for (int r = 0; r < dim / 2; ++r)
      {
        const int dA = 2 * r, dB = 2 * r + 1;
        int sA = 1;
        for (int k = 0; k < dA; ++k)
          sA *= N;
        const int       sB   = sA * N;
        const bool      last = r == dim / 2 - 1;
        for (int tt = tid; tt < nt; tt += nthr)
          {
            const int base = (tt % sA) + (tt / sA) * (sA * N * N); // dof index of the tile's (a, b) = (0, 0) entry
            T               U[N][N], out[N][N];
            HD_TG_UNROLL
            for (int b = 0; b < N; ++b)
              HD_TG_UNROLL
              for (int a = 0; a < N; ++a)
                {
                  U[b][a]   = HD_TG_LDG(uc + base + a * sA + b * sB);
                  out[b][a] = r == 0 ? T(0) : pc[base + a * sA + b * sB];
                }
            HD_TG_UNROLL
            for (int b = 0; b < N; ++b)
              HD_TG_UNROLL
              for (int a = 0; a < N; ++a)
                {
                  T v = out[b][a];
                  HD_TG_UNROLL
                  for (int j = 0; j < N; ++j)
                    v += cf.C[dA][a * N + j] * U[b][j];
                  HD_TG_UNROLL
                  for (int j = 0; j < N; ++j)
                    v += cf.C[dB][b * N + j] * U[j][a];
                  out[b][a] = v;
                }
            for (int which = 0; which < 2; ++which)
              {
                const int d  = which ? dB : dA;
                const int sd = which ? sB : sA, so = which ? sA : sB;
                for (int side = 0; side < 2; ++side)
                  {
                    if (!((p.nb_mask[d] >> side) & 1))
                      continue;
                    const bool at_edge = side ? (c[d] == p.ncell[d] - 1) : (c[d] == 0);
                    const int  layer   = side ? 0 : N - 1; // neighbour's layer touching the shared face
                    T          tv[N];
                    if (at_edge && p.side_kind[d][side] == HD_SIDE_GHOST)
                      {
                        long long fc = 0, m = 1;
                        for (int e = 0; e < dim; ++e)
                          if (e != d)
                            {
                              fc += c[e] * m;
                              m *= p.ncell[e];
                            }
                        const T *g = p.ghost + p.ghost_off[d][side] + fc * (long long)(nd / N);
                        HD_TG_UNROLL
                        for (int x = 0; x < N; ++x)
                          {
                            const int o = base + x * so; // dof index with digit d = 0
                            tv[x]             = HD_TG_LDG(g + (o % sd) + (o / (sd * N)) * sd);
                          }
                      }
                    else
                      {
                        long long nb = cell + (side ? cstr[d] : -cstr[d]);
                        if (at_edge) // periodic inside the brick
                          nb = cell + (side ? -(long long)(p.ncell[d] - 1) * cstr[d] : (long long)(p.ncell[d] - 1) * cstr[d]);
                        const T *g = p.src + nb * (long long)nd + base + layer * sd;
                        HD_TG_UNROLL
                        for (int x = 0; x < N; ++x)
                          tv[x] = HD_TG_LDG(g + x * so);
                      }
                    const T *L = side ? cf.L1[d] : cf.L0[d];
                    HD_TG_UNROLL
                    for (int b = 0; b < N; ++b)
                      HD_TG_UNROLL
                      for (int a = 0; a < N; ++a)
                        out[b][a] += which ? L[b] * tv[a] : L[a] * tv[b];
                  }
              }
            HD_TG_UNROLL
            for (int b = 0; b < N; ++b)
              HD_TG_UNROLL
              for (int a = 0; a < N; ++a)
                {
                  const int i = base + a * sA + b * sB;
                  if (last && p.fused)
                    {
                      const long long g = cell * (long long)nd + i;
                      const T         s = p.sol[g];
                      p.sol[g]          = s + p.fb * out[b][a];
                      if (p.fa != T(0))
                        p.ti_next[g] = s + p.fa * out[b][a];
                    }
                  else
                    (last ? oc : pc)[i] = out[b][a];
                }
          }
        HD_TG_SYNC(); // the next round reads the partial sums other threads of this CTA have written
      }
  }
} // namespace

#ifndef HD_VP_HOST_EMULATION
namespace
{
  template <typename T, int N, int THREADS>
  __global__ void __launch_bounds__(THREADS) k_apply_tile_global(const __grid_constant__ TgParams<T> p, const __grid_constant__ TgCoef<T, N> cf)
  {
    tg_cell<T, N>(p, cf, blockIdx.x, threadIdx.x, blockDim.x);
  }

  // Variant with the partial sums in shared memory (one CTA per SM, the cell's nd values = 182 KiB for degree 5 in FP32): the
  // rounds re-read only src (L1/L2), dst is written once.  No padding needed: a tile of round 0 is 36 contiguous values, i.e.
  // consecutive threads are 144 B apart (conflict-free 16-byte accesses), the later rounds are contiguous across the lanes.
  template <typename T, int N, int THREADS>
  __global__ void __launch_bounds__(THREADS, 1) k_apply_tile_global_sp(const __grid_constant__ TgParams<T> p, const __grid_constant__ TgCoef<T, N> cf)
  {
    extern __shared__ __align__(16) unsigned char tg_smem[];
    for (long long cell = blockIdx.x; cell < p.ncells; cell += gridDim.x)
      {
        tg_cell<T, N>(p, cf, cell, threadIdx.x, blockDim.x, reinterpret_cast<T *>(tg_smem));
        __syncthreads(); // (the last round still reads `part` while the next cell's round 0 would overwrite it)
      }
  }

  // Experiment knobs (profiles/r02_tile_global_residency_ab.txt).  A resident CTA works on one cell = 2 x nd values (degree 5
  // FP32: 364 KiB), and ncu measured 16.8 GB of DRAM traffic per launch for 5.2 GB algorithmic
  // (profiles/r02x_tile_global_ncu_summary.json), which suggested that too many cells in flight thrash L2.  Capping the CTAs
  // per SM (HD_TG_CTAS_PER_SM = n: a dynamic shared-memory request nobody uses; HD_TG_THREADS = 256 | 512) does NOT help:
  // 1 CTA x 512 threads 10.1 ms, 2 x 512 7.3 ms, 3 or 4 x 256 6.85 ms, no cap 6.78 ms — the kernel needs the parallelism more
  // than the locality (long-scoreboard stalls 48 %).  Default: no cap.
  inline void
  tg_launch_shape(const hd_mesh *m, int *threads, size_t *smem)
  {
    static const int env_threads = [] {
      const char *e = getenv("HD_TG_THREADS");
      return e ? atoi(e) : 0;
    }();
    static const int env_ctas = [] {
      const char *e = getenv("HD_TG_CTAS_PER_SM");
      return e ? atoi(e) : -1;
    }();
    const int ctas = env_ctas > 0 ? env_ctas : 0;
    *threads = env_threads == 512 ? 512 : 256;
    // an SM has 228 KiB of shared memory and reserves 1 KiB per resident CTA: with this request exactly `ctas` CTAs fit
    *smem = (ctas > 0 && ctas < 8) ? (size_t)233472 / ctas - 1024 : 0;
    if (*smem > m->ctx->smem_optin)
      *smem = m->ctx->smem_optin;
  }

  template <typename T, int N>
  int
  launch_tg(hd_advection *op, void *dst, const void *src, const void *ghosts, const FusedUpdate &fu, void *scratch)
  {
    hd_mesh *    m = op->mesh;
    TgParams<T>  p;
    TgCoef<T, N> cf;
    p.src   = static_cast<const T *>(src);
    p.dst   = static_cast<T *>(fu.enabled ? scratch : dst); // fused update: the partial sums need a buffer of their own
    p.ghost = static_cast<const T *>(ghosts);
    p.dim   = m->dim;
    for (int d = 0; d < HD_MAX_DIM; ++d)
      {
        for (int i = 0; i < N * N; ++i)
          cf.C[d][i] = d < m->dim ? T(op->hC[d][0][i]) : T(0);
        for (int i = 0; i < N; ++i)
          {
            cf.L0[d][i] = d < m->dim ? T(op->hL0[d][i]) : T(0);
            cf.L1[d][i] = d < m->dim ? T(op->hL1[d][i]) : T(0);
          }
        p.ncell[d]        = d < m->dim ? m->d.n_cells[d] : 1;
        p.side_kind[d][0] = m->d.side_kind[d][0];
        p.side_kind[d][1] = m->d.side_kind[d][1];
        p.nb_mask[d]      = op->nb_mask[d];
        p.ghost_off[d][0] = m->ghost_off[d][0];
        p.ghost_off[d][1] = m->ghost_off[d][1];
      }
    p.ncells  = m->ncells;
    p.sol     = static_cast<T *>(fu.sol);
    p.ti_next = static_cast<T *>(fu.ti_next);
    p.fb      = T(fu.fb);
    p.fa      = T(fu.fa);
    p.fused   = fu.enabled;
    // partial sums in shared memory where a cell fits (HD_TG_SMEM_PARTIALS=0: in dst as before)
    static const int env_sp = [] {
      const char *e = getenv("HD_TG_SMEM_PARTIALS");
      return e ? atoi(e) : 1;
    }();
    const size_t part_bytes = (size_t)m->nd * sizeof(T);
    if (env_sp && N == 6 && part_bytes <= m->ctx->smem_optin)
      {
        // 1296 tiles per round: 448 threads (14 warps, 106 registers) take three tiles each; measured 448: 4.57 ms, 512: 4.62, 672 (two tiles
        // each, 96 registers with spills): 4.81
        // (HD_TG_SP_THREADS; profiles/r02tg_*.txt)
        static const int env_thr = [] {
          const char *e = getenv("HD_TG_SP_THREADS");
          return e ? atoi(e) : 448;
        }();
        const long long grid = m->ncells < m->ctx->sm_count ? m->ncells : m->ctx->sm_count;
        if (env_thr == 448)
          {
            HD_CUDA(cudaFuncSetAttribute(k_apply_tile_global_sp<T, N, 448>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_bytes));
            k_apply_tile_global_sp<T, N, 448><<<(unsigned)grid, 448, part_bytes, m->ctx->stream>>>(p, cf);
          }
        else if (env_thr == 512)
          {
            HD_CUDA(cudaFuncSetAttribute(k_apply_tile_global_sp<T, N, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_bytes));
            k_apply_tile_global_sp<T, N, 512><<<(unsigned)grid, 512, part_bytes, m->ctx->stream>>>(p, cf);
          }
        else
          {
            HD_CUDA(cudaFuncSetAttribute(k_apply_tile_global_sp<T, N, 672>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_bytes));
            k_apply_tile_global_sp<T, N, 672><<<(unsigned)grid, 672, part_bytes, m->ctx->stream>>>(p, cf);
          }
        HD_CUDA(cudaGetLastError());
        op->launches++;
        op->last_kernel = "tile_global"; // (same kernel family; HD_TG_SMEM_PARTIALS=0 selects the variant with the partial sums in dst)
        return HD_OK;
      }
    int    threads;
    size_t smem;
    tg_launch_shape(m, &threads, &smem);
    if (threads == 512)
      {
        if (smem > 48 * 1024)
          HD_CUDA(cudaFuncSetAttribute(k_apply_tile_global<T, N, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_apply_tile_global<T, N, 512><<<(unsigned)m->ncells, 512, smem, m->ctx->stream>>>(p, cf);
      }
    else
      {
        if (smem > 48 * 1024)
          HD_CUDA(cudaFuncSetAttribute(k_apply_tile_global<T, N, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_apply_tile_global<T, N, 256><<<(unsigned)m->ncells, 256, smem, m->ctx->stream>>>(p, cf);
      }
    HD_CUDA(cudaGetLastError());
    op->launches++;
    op->last_kernel = "tile_global";
    return HD_OK;
  }
} // namespace

namespace hd
{
  bool
  tile_global_supported(const hd_advection *op)
  {
    const hd_mesh *m = op->mesh;
    if ((m->n != 6 && m->n != 4) || m->dim % 2 != 0 || m->dim < 2)
      return false;
    for (int d = 0; d < m->dim; ++d)
      for (int s = 0; s < 2; ++s)
        if (m->d.side_kind[d][s] == HD_SIDE_DIRICHLET || m->d.side_kind[d][s] == HD_SIDE_DIRICHLET_HOM)
          return false;
    return true;
  }

  int
  launch_tile_global(hd_advection *op, void *dst, const void *src, const void *ghosts, double, const FusedUpdate &fu)
  {
    hd_mesh *m       = op->mesh;
    void *   scratch = nullptr;
    if (fu.enabled)
      {
        // rounds before the last one park their partial sums in the operator's staging vector
        const size_t bytes = (size_t)m->ndofs * m->elem_size;
        if (!op->d_stage_dst)
          {
            HD_CUDA(cudaMalloc(&op->d_stage_src, bytes));
            HD_CUDA(cudaMalloc(&op->d_stage_dst, bytes));
          }
        scratch = op->d_stage_dst;
      }
    const bool f64 = m->d.number_type == HD_F64;
    if (m->n == 6)
      return f64 ? launch_tg<double, 6>(op, dst, src, ghosts, fu, scratch) : launch_tg<float, 6>(op, dst, src, ghosts, fu, scratch);
    return f64 ? launch_tg<double, 4>(op, dst, src, ghosts, fu, scratch) : launch_tg<float, 4>(op, dst, src, ghosts, fu, scratch);
  }
} // namespace hd
#endif
