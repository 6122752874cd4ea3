// Generic advection kernel: any dim_x+dim_v in 2..6, degree 1..7, double or float.
//
// One CTA stages CPB cells in shared memory (coalesced 16-byte loads) and every thread
// produces whole lines of N outputs along direction 0:
//     dst_line = sum_d C_d u_line(d) + L0_d * trace_lower(d) + L1_d * trace_upper(d)
// (collapsed form of advection_operation.h:221-566, see basis.hpp).  Neighbour traces are the
// nodal face layers of the adjacent cells (the values FEFaceEvaluation::read_dof_values gathers
// through face_to_cell_index_nodal, matrix_free/read_write_operation.h:186-330), read through L2
// either from `src` itself or from the ghost-face buffer for bricks owned by another GPU.
// This kernel is the correctness workhorse and the fallback for every configuration; the
// 3D3V k=3 double case has its own pipelined kernel (kernel_fast6d.cu).
#include <cstdlib>

#include "hd_internal.h"

namespace
{
  template <typename T, int DIM>
  struct GenParams
  {
    const T *src;
    T *      dst;
    const T *ghost;
    const void *coef;
    int       ncell[DIM];
    int       side_kind[DIM][2];
    int       nb_mask[DIM];
    long long ghost_off[DIM][2];
    long long ncells;
    // fused LSRK epilogue
    T * sol;
    T * ti_next;
    T   fb, fa;
    int fused;
  };

  template <int N, int P>
  struct IPow
  {
    static constexpr long long value = N * IPow<N, P - 1>::value;
  };
  template <int N>
  struct IPow<N, 0>
  {
    static constexpr long long value = 1;
  };

  template <typename T, int N, int DIM, int CPB, int THREADS>
  __global__ void __launch_bounds__(THREADS) k_apply_generic(const GenParams<T, DIM> p)
  {
    constexpr int ND = IPow<N, DIM>::value;
    constexpr int NL = ND / N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *                  u    = reinterpret_cast<T *>(smem_raw);
    DirCoef<T, N> *      coef = reinterpret_cast<DirCoef<T, N> *>(u + (size_t)CPB * ND);
    const long long      cell0 = (long long)blockIdx.x * CPB;

    // stage coefficients
    {
      const int *     g = reinterpret_cast<const int *>(p.coef);
      int *           s = reinterpret_cast<int *>(coef);
      constexpr int   words = sizeof(DirCoef<T, N>) * DIM / 4;
      for (int i = threadIdx.x; i < words; i += THREADS)
        s[i] = g[i];
    }
    // stage cells (CPB consecutive cells are contiguous in memory)
    {
      long long n_valid = p.ncells - cell0;
      if (n_valid > CPB)
        n_valid = CPB;
      const long long total = n_valid * ND;
      const T *       g     = p.src + cell0 * ND;
      if ((ND * sizeof(T)) % 16 == 0)
        {
          const int4 *g4 = reinterpret_cast<const int4 *>(g);
          int4 *      s4 = reinterpret_cast<int4 *>(u);
          const int   n4 = int(total * sizeof(T) / 16);
          for (int i = threadIdx.x; i < n4; i += THREADS)
            s4[i] = __ldg(g4 + i);
        }
      else
        for (int i = threadIdx.x; i < total; i += THREADS)
          u[i] = g[i];
    }
    __syncthreads();

    for (int l = threadIdx.x; l < CPB * NL; l += THREADS)
      {
        const int       lc   = l / NL;
        const int       li   = l - lc * NL;
        const long long cell = cell0 + lc;
        if (cell >= p.ncells)
          break;
        // cell coordinates
        int c[DIM];
        {
          long long r = cell;
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            {
              c[d] = int(r % p.ncell[d]);
              r /= p.ncell[d];
            }
        }
        const T * uc     = u + (size_t)lc * ND;
        const int o_base = li * N;
        T         acc[N];
#pragma unroll
        for (int i = 0; i < N; ++i)
          acc[i] = 0;

        long long cstride = 1; // cell stride of direction d
        int       stride  = 1; // dof stride of direction d
        int       digits  = li; // remaining digits i_1, i_2, ...
#pragma unroll
        for (int d = 0; d < DIM; ++d)
          {
            const int i_d = (d == 0) ? 0 : (digits % N);
            if (d > 0)
              digits /= N;
            // variant: Dirichlet on lower / upper side of this cell
            const bool at_lo = (c[d] == 0), at_hi = (c[d] == p.ncell[d] - 1);
            const int  k_lo = p.side_kind[d][0], k_hi = p.side_kind[d][1];
            const bool dir_lo = at_lo && (k_lo >= HD_SIDE_DIRICHLET);
            const bool dir_hi = at_hi && (k_hi >= HD_SIDE_DIRICHLET);
            const T *  C      = coef[d].C[(dir_lo ? 1 : 0) + (dir_hi ? 2 : 0)];

            if (d == 0)
              {
                T r[N];
#pragma unroll
                for (int j = 0; j < N; ++j)
                  r[j] = uc[o_base + j];
#pragma unroll
                for (int i = 0; i < N; ++i)
#pragma unroll
                  for (int j = 0; j < N; ++j)
                    acc[i] += C[i * N + j] * r[j];
              }
            else
              {
                const int base = o_base - i_d * stride;
#pragma unroll
                for (int j = 0; j < N; ++j)
                  {
                    const T cj = C[i_d * N + j];
#pragma unroll
                    for (int i = 0; i < N; ++i)
                      acc[i] += cj * uc[base + j * stride + i];
                  }
              }

            // neighbour traces
#pragma unroll
            for (int side = 0; side < 2; ++side)
              {
                if (!((p.nb_mask[d] >> side) & 1))
                  continue;
                const bool at_edge = side ? at_hi : at_lo;
                const int  kind    = side ? k_hi : k_lo;
                if (at_edge && kind >= HD_SIDE_DIRICHLET)
                  continue;
                const int layer = side ? 0 : N - 1; // neighbour's layer touching the shared face
                const T * t;
                int       tstride; // stride between the N values of this thread's line
                if (at_edge && kind == HD_SIDE_GHOST)
                  {
                    // face-cell index: cell index with coordinate d removed
                    long long fc = 0, m = 1;
#pragma unroll
                    for (int e = 0; e < DIM; ++e)
                      if (e != d)
                        {
                          fc += c[e] * m;
                          m *= p.ncell[e];
                        }
                    // face dof index of (line, i0 = 0): dof index with digit d removed
                    const int o  = o_base;
                    const int fo = (d == 0) ? (o / N) : ((o % stride) + (o / (stride * N)) * stride);
                    t            = p.ghost + p.ghost_off[d][side] + fc * (ND / N) + fo;
                    tstride      = (d == 0) ? 0 : 1;
                  }
                else
                  {
                    long long nb = cell + (side ? cstride : -cstride);
                    if (at_edge) // periodic inside the brick
                      nb = cell + (side ? -(long long)(p.ncell[d] - 1) * cstride : (long long)(p.ncell[d] - 1) * cstride);
                    t       = p.src + nb * ND + (o_base - i_d * stride) + layer * stride;
                    tstride = (d == 0) ? 0 : 1;
                  }
                const T *L = side ? coef[d].L1 : coef[d].L0;
                if (d == 0)
                  {
                    const T tv = __ldg(t);
#pragma unroll
                    for (int i = 0; i < N; ++i)
                      acc[i] += L[i] * tv;
                  }
                else
                  {
                    const T li_d = L[i_d];
#pragma unroll
                    for (int i = 0; i < N; ++i)
                      acc[i] += li_d * __ldg(t + i * tstride);
                  }
              }
            cstride *= p.ncell[d];
            stride *= N;
          }

        const long long g = cell * ND + o_base;
        if (p.fused)
          {
#pragma unroll
            for (int i = 0; i < N; ++i)
              {
                const T s    = p.sol[g + i];
                p.sol[g + i] = s + p.fb * acc[i];
                if (p.fa != T(0))
                  p.ti_next[g + i] = s + p.fa * acc[i];
              }
          }
        else
          {
#pragma unroll
            for (int i = 0; i < N; ++i)
              p.dst[g + i] = acc[i];
          }
      }
  }

  // --------------------------------------------------------------------------------------
  // Dirichlet source:  K += 2 beta_f l_f(i_d) * ghat,   ghat = (x)_{e != d} Sinv  g(face quad pts)
  // One CTA per boundary face cell.  g is either uploaded by the host (hd_advection_set_
  // dirichlet_values) or a built-in analytic field evaluated here at the stage time
  // (advection_operation.h:490-519, boundary_descriptor.h:79-107, matrix_free/tools.h:31-50).
  template <typename T>
  struct DirParams
  {
    T *           dst;
    const double *g;        // [n_face_cells][nq^(dim-1)] or nullptr (built-in)
    const double *basis;    // nodes[n], xq[nq], w[nq], S[nq*n], Sinv[n*nq]
    const void *  lift;     // double[n]: 2*beta_f*l_f for this (dir, side)
    int           dim, n, nq, dir, side, fn_id;
    int           ncell[HD_MAX_DIM], cell_offset[HD_MAX_DIM];
    double        left[HD_MAX_DIM], h[HD_MAX_DIM];
    double        time;
    T *           sol;
    T *           ti_next;
    double        fb, fa;
    int           fused;
    // ghost mode (Dirichlet sides served by the ghost-side code of the fast kernels): instead of lifting 2 beta l ghat
    // into the cell, write the trace the upwind flux would see behind the face, -u_face + 2 ghat, into a ghost segment
    const T *src_for_ghost;
    T *      ghost_out; // start of this side's segment, [n_face_cells][n^(dim-1)]
    int      homogeneous;
  };

  __device__ double
  builtin_fn(int fn_id, int dim, const double *x, double t)
  {
    if (fn_id == HD_FN_HYPERRECTANGLE)
      {
        const double adv[6] = {1.0, 0.15, -0.05, 0.0, 0.0, 0.0};
        const double PI     = 3.14159265358979323846;
        double       r      = sin(2.0 * (x[0] - t * adv[0]) * PI);
        for (int d = 1; d < dim; ++d)
          r *= cos(2.0 * (x[d] - t * adv[d]) * PI);
        return r;
      }
    return 0.0;
  }

  template <typename T>
  __global__ void k_dirichlet_source(const DirParams<T> p)
  {
    extern __shared__ double sm[];
    const int dim = p.dim, n = p.n, nq = p.nq, fd = dim - 1;
    int       mx = n > nq ? n : nq;
    int       cap = 1;
    for (int e = 0; e < fd; ++e)
      cap *= mx;
    double *      A      = sm;
    double *      B      = sm + cap;
    const double *nodes  = p.basis;
    const double *xq     = nodes + n;
    const double *Sinv   = xq + nq + nq + nq * n;
    const double *lift   = reinterpret_cast<const double *>(p.lift);
    const long long fc   = blockIdx.x;
    // face cell -> cell coordinates
    int       c[HD_MAX_DIM];
    long long r = fc;
    for (int e = 0; e < dim; ++e)
      if (e != p.dir)
        {
          c[e] = int(r % p.ncell[e]);
          r /= p.ncell[e];
        }
    c[p.dir] = p.side ? p.ncell[p.dir] - 1 : 0;
    int nqf = 1, nf = 1;
    for (int e = 0; e < fd; ++e)
      {
        nqf *= nq;
        nf *= n;
      }
    // g at face quadrature points.  The built-in hyperrectangle solution is a product of one factor per direction
    // (cases/hyperrectangle.h:46-57): 1-D factor tables first, then products — dim * nq instead of dim * nq^(dim-1)
    // sin/cos evaluations per face cell, which dominated this kernel (3D3V: 6 x 1024 per face cell)
    double *   fac       = sm + 2 * cap; // [dim][nq]
    const bool separable = !p.homogeneous && !p.g && p.fn_id == HD_FN_HYPERRECTANGLE;
    if (separable)
      {
        for (int i = threadIdx.x; i < dim * nq; i += blockDim.x)
          {
            const int    e = i / nq, q = i % nq;
            const double adv[6] = {1.0, 0.15, -0.05, 0.0, 0.0, 0.0};
            const double PI     = 3.14159265358979323846;
            const double x = e == p.dir ? p.left[e] + p.h[e] * (c[e] + p.cell_offset[e] + (p.side ? 1.0 : 0.0)) : p.left[e] + p.h[e] * (c[e] + p.cell_offset[e] + xq[q]);
            fac[i]         = e == 0 ? sin(2.0 * (x - p.time * adv[e]) * PI) : cos(2.0 * (x - p.time * adv[e]) * PI);
          }
        __syncthreads();
        // The projection of a product of 1-D factors is the product of the 1-D projections: ghat[i] = prod_e (Sinv f_e)[i_e].
        // So the dim - 1 transverse sweeps over the whole face (with a block barrier each) reduce to dim - 1 tiny
        // matrix-vector products and one product per nodal face value.
        double *pf = fac + HD_MAX_DIM * mx; // [dim][n]
        for (int i = threadIdx.x; i < dim * n; i += blockDim.x)
          {
            const int e = i / n, r = i % n;
            double    acc = 0.0;
            if (e == p.dir)
              acc = fac[e * nq]; // the factor at the face coordinate (no projection along the face normal)
            else
              for (int k = 0; k < nq; ++k)
                acc += Sinv[r * nq + k] * fac[e * nq + k];
            pf[i] = acc;
          }
        __syncthreads();
        for (int i = threadIdx.x; i < nf; i += blockDim.x)
          {
            int    rr = i;
            double r  = 1.0;
            for (int e = 0; e < dim; ++e)
              {
                int ie = 0;
                if (e != p.dir)
                  {
                    ie = rr % n;
                    rr /= n;
                  }
                r = e == 0 ? pf[ie] : r * pf[e * n + ie];
              }
            A[i] = r;
          }
        __syncthreads();
      }
    for (int q = threadIdx.x; q < nqf && !separable; q += blockDim.x)
      {
        double val;
        if (p.homogeneous)
          val = 0.0;
        else if (p.g)
          val = p.g[fc * nqf + q];
        else if (separable)
          {
            // same order of the products as builtin_fn: r = f_0, then r *= f_e for e = 1..dim-1
            int    rr = q;
            double r  = 1.0;
            for (int e = 0; e < dim; ++e)
              {
                int qe = 0;
                if (e != p.dir)
                  {
                    qe = rr % nq;
                    rr /= nq;
                  }
                r = e == 0 ? fac[qe] : r * fac[e * nq + qe];
              }
            val = r;
          }
        else
          {
            double x[HD_MAX_DIM];
            int    rr = q;
            for (int e = 0; e < dim; ++e)
              {
                if (e == p.dir)
                  x[e] = p.left[e] + p.h[e] * (c[e] + p.cell_offset[e] + (p.side ? 1.0 : 0.0));
                else
                  {
                    x[e] = p.left[e] + p.h[e] * (c[e] + p.cell_offset[e] + xq[rr % nq]);
                    rr /= nq;
                  }
              }
            val = builtin_fn(p.fn_id, dim, x, p.time);
          }
        A[q] = val;
      }
    __syncthreads();
    // transverse Sinv sweeps, last face direction first (extents: nq for not-yet-swept dirs, n for swept)
    double *in = A, *out = B;
    for (int e = fd - 1; e >= 0 && !separable; --e)
      {
        int stride = 1;
        for (int k = 0; k < e; ++k)
          stride *= nq;
        int outer = 1;
        for (int k = e + 1; k < fd; ++k)
          outer *= n;
        const int total = outer * n * stride;
        for (int i = threadIdx.x; i < total; i += blockDim.x)
          {
            const int lo = i % stride, rest = i / stride, row = rest % n, o = rest / n;
            double    acc = 0;
            for (int k = 0; k < nq; ++k)
              acc += Sinv[row * nq + k] * in[(o * nq + k) * stride + lo];
            out[i] = acc;
          }
        __syncthreads();
        double *tmp = in;
        in          = out;
        out         = tmp;
      }
    // lift into the cell
    long long cell = 0, m = 1;
    for (int e = 0; e < dim; ++e)
      {
        cell += c[e] * m;
        m *= p.ncell[e];
      }
    int stride_d = 1;
    for (int e = 0; e < p.dir; ++e)
      stride_d *= n;
    const long long nd = (long long)nf * n;
    if (p.ghost_out)
      {
        // u+ = -u- + 2 g at a Dirichlet face (advection_operation.h:490-519) as the neighbour trace of the interior formula
        const int layer = p.side ? n - 1 : 0;
        for (int i = threadIdx.x; i < nf; i += blockDim.x)
          {
            const int    lo = i % stride_d, hi = i / stride_d;
            const double u  = double(p.src_for_ghost[cell * nd + ((long long)hi * n + layer) * stride_d + lo]);
            p.ghost_out[fc * nf + i] = T(2.0 * in[i] - u);
          }
        return;
      }
    for (int i = threadIdx.x; i < nd; i += blockDim.x)
      {
        const int    lo = i % stride_d, rest = i / stride_d, id = rest % n, hi = rest / n;
        const double v = lift[id] * in[hi * stride_d + lo];
        const long long g = cell * nd + i;
        if (p.fused)
          {
            p.sol[g] += T(p.fb * v);
            if (p.fa != 0.0)
              p.ti_next[g] += T(p.fa * v);
          }
        else
          p.dst[g] += T(v);
      }
  }

  template <typename T, int N, int DIM>
  int
  launch_t(hd_advection *op, void *dst, const void *src, const void *ghosts, const FusedUpdate &fu)
  {
    hd_mesh *         m  = op->mesh;
    constexpr int     ND = IPow<N, DIM>::value;
    constexpr int     CPB = (ND >= 2048) ? 1 : (2048 / ND);
    constexpr int     NL  = ND / N;
    constexpr int     THREADS = (CPB * NL >= 256) ? 256 : ((CPB * NL + 31) / 32 * 32);
    GenParams<T, DIM> p;
    p.src   = static_cast<const T *>(src);
    p.dst   = static_cast<T *>(dst);
    p.ghost = static_cast<const T *>(ghosts);
    p.coef  = op->d_coef;
    for (int d = 0; d < DIM; ++d)
      {
        p.ncell[d]        = m->d.n_cells[d];
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
    const size_t smem = (size_t)CPB * ND * sizeof(T) + sizeof(DirCoef<T, N>) * DIM;
    auto         kern = k_apply_generic<T, N, DIM, CPB, THREADS>;
    if (smem > 48 * 1024)
      {
        if (smem > m->ctx->smem_optin)
          return hd::fail(HD_ERR_UNSUPPORTED, "cell does not fit into shared memory");
        HD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      }
    const long long grid = (m->ncells + CPB - 1) / CPB;
    kern<<<(unsigned)grid, THREADS, smem, m->ctx->stream>>>(p);
    HD_CUDA(cudaGetLastError());
    op->launches++;
    op->last_kernel = "generic";
    return HD_OK;
  }

  template <typename T, int N>
  int
  launch_dim(hd_advection *op, void *dst, const void *src, const void *ghosts, const FusedUpdate &fu)
  {
    switch (op->mesh->dim)
      {
        case 2:
          return launch_t<T, N, 2>(op, dst, src, ghosts, fu);
        case 3:
          return launch_t<T, N, 3>(op, dst, src, ghosts, fu);
        case 4:
          return launch_t<T, N, 4>(op, dst, src, ghosts, fu);
        case 5:
          return launch_t<T, N, 5>(op, dst, src, ghosts, fu);
        case 6:
          return launch_t<T, N, 6>(op, dst, src, ghosts, fu);
      }
    return hd::fail(HD_ERR_UNSUPPORTED, "dim_x + dim_v must be in 2..6");
  }

  // --------------------------------------------------------------------------------------
  // Tile kernel: degree 3, dim_x + dim_v = 2, 4 or 6 (1D1V, 2D2V — BASELINE.json configs[0] —, 3D3V), periodic and ghost sides.
  //
  // The generic kernel above re-reads every value of a cell 13 times from shared memory (one line of outputs per thread,
  // all directions) and is shared-memory bound at ~20 % of the HBM roofline.  Here the directions are taken two at a
  // time ("rounds", like the pipelined 3D3V kernel): in round r a thread owns the 4x4 tile over directions (2r, 2r+1) of
  // one cell in registers — 16 loads for 16 x (4 + 4) FMAs — and hands its partial sums to the next round through a
  // second shared-memory buffer; the last round writes dst (or the fused LSRK update) coalesced.  A CTA of 256 threads
  // works on 256 / 4^(dim-2) consecutive cells (contiguous in memory, staged with 16-byte loads); tiles are padded by 16
  // bytes so that the 128-bit tile loads of round 0 are conflict-free.  The (k+1)x(k+1) matrices are a __grid_constant__
  // kernel parameter (FMA operands straight from the constant bank).  Neighbour traces: the direction-0 trace comes from
  // shared memory when the neighbour cell is in the same CTA (its values are 32 B apart in global memory), all others
  // are contiguous in global memory and read through L1/L2.
  // minimum CTAs per SM the register allocation of the tile kernel is held to: FP64 5 x 128 threads at 96 registers
  // (measured 270 GDoF/s; 132 registers / 3 CTAs: 207, 80 registers with spills / 6 CTAs: 169), FP32 8
  constexpr int
  tile_min_ctas(int elem_size, int threads)
  {
    const int want = (elem_size == 8 ? 640 : 1024) / threads; // resident threads per SM the registers must allow
    return want < 1 ? 1 : want;
  }
  template <typename T>
  struct TileCoef
  {
    T C[6][16]; // [direction][out * 4 + in]
    T L0[6][4]; // lifting of the lower neighbour's trace
    T L1[6][4]; // ... upper neighbour's
  };

  template <typename T, int DIM, int THREADS>
  __global__ void __launch_bounds__(THREADS, tile_min_ctas(sizeof(T), THREADS)) k_apply_tile(const __grid_constant__ GenParams<T, DIM> p, const __grid_constant__ TileCoef<T> cf)
  {
    constexpr int N   = 4;
    constexpr int ND  = IPow<N, DIM>::value;
    constexpr int NT  = ND / 16;          // tiles per cell and round
    constexpr int CPB = THREADS / NT;     // cells per CTA
    constexpr int PAD = 16 / sizeof(T);   // padding per tile (values)
    constexpr int TS  = 16 + PAD;         // padded tile size
    constexpr int NDP = NT * TS;          // padded cell size
    constexpr int V4  = 16 * sizeof(T) / 16, TS4 = TS * sizeof(T) / 16; // 16-byte words per tile, plain and padded
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *             u     = reinterpret_cast<T *>(smem_raw);
    T *             part  = u + (size_t)CPB * NDP;
    const long long cell0 = (long long)blockIdx.x * CPB;
    long long       n_valid = p.ncells - cell0;
    if (n_valid > CPB)
      n_valid = CPB;
    // cell coordinates first: the neighbour-trace loads below do not depend on the staged data
    const int       lc     = threadIdx.x / NT, tt = threadIdx.x % NT;
    const long long cell   = cell0 + lc;
    const bool      active = lc < n_valid;
    int             c[DIM];
    long long       cstr[DIM];
    {
      long long r = active ? cell : 0, m = 1;
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        {
          c[d]    = int(r % p.ncell[d]);
          r /= p.ncell[d];
          cstr[d] = m;
          m *= p.ncell[d];
        }
    }
    auto padded = [](int i) { return i + PAD * (i >> 4); };
    const T *uc = u + (size_t)lc * NDP;
    T *      pc = part + (size_t)lc * NDP;

    // Neighbour traces of the two directions of round r: tv[which][side][x] (which: 0 = direction 2r, 1 = 2r+1; x runs
    // along the tile's other direction).  Global / ghost sources are requested here; a direction-0 neighbour inside this
    // CTA's batch is taken from shared memory later (from_smem), once the batch has landed.
    auto request_traces = [&](const int r, T(&tv)[2][2][4], const T *(&from_smem)[2]) {
      const int dA = 2 * r, dB = 2 * r + 1;
      const int sA = 1 << (2 * dA), sB = 1 << (2 * dB);
      const int base = (tt % sA) + (tt / sA) * (sA * 16);
#pragma unroll
      for (int which = 0; which < 2; ++which)
        {
          const int d = which ? dB : dA, sd = which ? sB : sA, so = which ? sA : sB;
#pragma unroll
          for (int side = 0; side < 2; ++side)
            {
#pragma unroll
              for (int x = 0; x < 4; ++x)
                tv[which][side][x] = T(0);
              if (which == 0)
                from_smem[side] = nullptr;
              if (!active || !((p.nb_mask[d] >> side) & 1))
                continue;
              const bool at_edge = side ? (c[d] == p.ncell[d] - 1) : (c[d] == 0);
              const int  layer   = side ? 0 : N - 1; // neighbour's layer touching the shared face
              if (at_edge && p.side_kind[d][side] == HD_SIDE_GHOST)
                {
                  long long fc = 0, m = 1;
#pragma unroll
                  for (int e = 0; e < DIM; ++e)
                    if (e != d)
                      {
                        fc += c[e] * m;
                        m *= p.ncell[e];
                      }
                  const T *g = p.ghost + p.ghost_off[d][side] + fc * (ND / N);
#pragma unroll
                  for (int x = 0; x < 4; ++x)
                    {
                      const int o        = base + x * so; // dof index with digit d = 0
                      tv[which][side][x] = __ldg(g + (o % sd) + (o / (sd * N)) * sd);
                    }
                }
              else
                {
                  long long nb = cell + (side ? cstr[d] : -cstr[d]);
                  if (at_edge) // periodic inside the brick
                    nb = cell + (side ? -(long long)(p.ncell[d] - 1) * cstr[d] : (long long)(p.ncell[d] - 1) * cstr[d]);
                  const int o = base + layer * sd;
                  if (d == 0 && nb >= cell0 && nb < cell0 + n_valid)
                    from_smem[side] = u + (size_t)(nb - cell0) * NDP; // (which == 0 here)
                  else
                    {
                      const T *g = p.src + nb * ND + o;
#pragma unroll
                      for (int x = 0; x < 4; ++x)
                        tv[which][side][x] = __ldg(g + x * so);
                    }
                }
            }
        }
    };

    T        tv[2][2][4];
    const T *from_smem[2];
    request_traces(0, tv, from_smem);

    // stage the batch (16-byte loads, all of a thread's loads in flight before the first store)
    {
      const int4 *  g4  = reinterpret_cast<const int4 *>(p.src + cell0 * ND);
      int4 *        s4  = reinterpret_cast<int4 *>(u);
      constexpr int PER = CPB * ND * (int)sizeof(T) / 16 / THREADS;
      if (n_valid == CPB)
        {
          int4 w[PER];
#pragma unroll
          for (int k = 0; k < PER; ++k)
            w[k] = __ldg(g4 + threadIdx.x + k * THREADS);
#pragma unroll
          for (int k = 0; k < PER; ++k)
            {
              const int i                        = threadIdx.x + k * THREADS;
              s4[(i / V4) * TS4 + (i % V4)] = w[k];
            }
        }
      else
        {
          const int n4 = int(n_valid * ND * sizeof(T) / 16);
          for (int i = threadIdx.x; i < n4; i += THREADS)
            s4[(i / V4) * TS4 + (i % V4)] = __ldg(g4 + i);
        }
    }
    __syncthreads();

#pragma unroll
    for (int r = 0; r < DIM / 2; ++r)
      {
        const int dA = 2 * r, dB = 2 * r + 1;
        const int sA = 1 << (2 * dA), sB = 1 << (2 * dB);
        const int base = (tt % sA) + (tt / sA) * (sA * 16); // dof index of the tile's (a, b) = (0, 0) entry
        if (r > 0)
          request_traces(r, tv, from_smem);
        if (active)
          {
            T U[4][4], out[4][4]; // [b][a]
            if (r == 0)
              {
                // the tile is contiguous: 128-bit loads
                const int4 *q = reinterpret_cast<const int4 *>(uc + tt * TS);
                int4        w[V4];
#pragma unroll
                for (int i = 0; i < V4; ++i)
                  w[i] = q[i];
                const T *wv = reinterpret_cast<const T *>(w);
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                  for (int a = 0; a < 4; ++a)
                    {
                      U[b][a]   = wv[4 * b + a];
                      out[b][a] = T(0);
                    }
                // direction-0 neighbours inside the batch
#pragma unroll
                for (int side = 0; side < 2; ++side)
                  if (from_smem[side])
                    {
                      const int o = base + (side ? 0 : N - 1) * sA;
#pragma unroll
                      for (int x = 0; x < 4; ++x)
                        tv[0][side][x] = from_smem[side][padded(o + x * sB)];
                    }
              }
            else
              {
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                  for (int a = 0; a < 4; ++a)
                    {
                      const int i = padded(base + a * sA + b * sB);
                      U[b][a]     = uc[i];
                      out[b][a]   = pc[i];
                    }
              }
            // the two in-tile sweeps
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
              for (int a = 0; a < 4; ++a)
                {
                  T v = out[b][a];
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    v += cf.C[dA][a * 4 + j] * U[b][j];
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    v += cf.C[dB][b * 4 + j] * U[j][a];
                  out[b][a] = v;
                }
            // neighbour traces (zero where there is none)
#pragma unroll
            for (int side = 0; side < 2; ++side)
              {
                const T *LA = side ? cf.L1[dA] : cf.L0[dA];
                const T *LB = side ? cf.L1[dB] : cf.L0[dB];
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                  for (int a = 0; a < 4; ++a)
                    out[b][a] += LA[a] * tv[0][side][b] + LB[b] * tv[1][side][a];
              }
            if (r == DIM / 2 - 1)
              {
                const long long g = cell * ND + base;
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                  for (int a = 0; a < 4; ++a)
                    {
                      const long long i = g + a * sA + b * sB;
                      if (p.fused)
                        {
                          const T s = p.sol[i];
                          p.sol[i]  = s + p.fb * out[b][a];
                          if (p.fa != T(0))
                            p.ti_next[i] = s + p.fa * out[b][a];
                        }
                      else
                        p.dst[i] = out[b][a];
                    }
              }
            else
              {
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                  for (int a = 0; a < 4; ++a)
                    pc[padded(base + a * sA + b * sB)] = out[b][a];
              }
          }
        if (r < DIM / 2 - 1)
          __syncthreads();
      }
  }

  // --------------------------------------------------------------------------------------
  // Row-persistent tile kernel (3D3V, degree 3): the tile kernel's three rounds, but a CTA walks a whole row of cells
  // along x_0 in upwind order.  The previous cell stays in shared memory, so the direction-0 neighbour traces (values
  // 32 B apart in global memory — a full extra cell of sector traffic for the plain tile kernel) never leave the SM, and
  // the next cell is staged with cp.async while rounds 1 and 2 of the current one run.  One code path for all threads,
  // __syncthreads only.  Shared memory: 2 cell buffers + partial sums = 3 x 36 KiB (FP64), two CTAs per SM.
  __device__ __forceinline__ void
  cp_async_16(void *smem_dst, const void *gmem_src)
  {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
  }
  __device__ __forceinline__ void
  cp_async_wait_all()
  {
    asm volatile("cp.async.wait_all;" ::: "memory");
  }

  template <typename T>
  __global__ void __launch_bounds__(256, 2) k_apply_tile_row(const __grid_constant__ GenParams<T, 6> p, const __grid_constant__ TileCoef<T> cf)
  {
    constexpr int N = 4, DIM = 6, ND = 4096, NT = 256, THREADS = 256;
    constexpr int PAD = 16 / sizeof(T), TS = 16 + PAD, NDP = NT * TS;
    constexpr int V4 = 16 * sizeof(T) / 16, TS4 = TS * sizeof(T) / 16, PER = ND * (int)sizeof(T) / 16 / THREADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *       ub0  = reinterpret_cast<T *>(smem_raw);
    T *       ub1  = ub0 + NDP;
    T *       part = ub1 + NDP;
    const int tt   = threadIdx.x;
    const int n0   = p.ncell[0];
    const long long nrows = p.ncells / n0;
    const bool descend = (p.nb_mask[0] & 2) != 0; // upwind neighbour is the upper cell: walk downwards
    auto padded = [](int i) { return i + PAD * (i >> 4); };
    auto stage = [&](T *dst, long long cell) {
      const int4 *g4 = reinterpret_cast<const int4 *>(p.src + cell * ND);
      int4 *      s4 = reinterpret_cast<int4 *>(dst);
#pragma unroll
      for (int k = 0; k < PER; ++k)
        {
          const int i = tt + k * THREADS;
          cp_async_16(s4 + (i / V4) * TS4 + (i % V4), g4 + i);
        }
    };

    for (long long row = blockIdx.x; row < nrows; row += gridDim.x)
      {
        int       c[DIM];
        long long cstr[DIM];
        {
          long long r = row, m = n0;
          cstr[0]     = 1;
#pragma unroll
          for (int d = 1; d < DIM; ++d)
            {
              c[d]    = int(r % p.ncell[d]);
              r /= p.ncell[d];
              cstr[d] = m;
              m *= p.ncell[d];
            }
        }
        const long long row_cell0 = row * n0;
        stage(ub0, row_cell0 + (descend ? n0 - 1 : 0));
        cp_async_wait_all();
        __syncthreads();

        for (int step = 0; step < n0; ++step)
          {
            c[0]                 = descend ? n0 - 1 - step : step;
            const long long cell = row_cell0 + c[0];
            const T *       uc   = (step & 1) ? ub1 : ub0; // this cell
            T *             un   = (step & 1) ? ub0 : ub1; // the previous cell of the walk = upwind neighbour; then the next cell
#pragma unroll
            for (int r = 0; r < DIM / 2; ++r)
              {
                const int dA = 2 * r, dB = 2 * r + 1;
                const int sA = 1 << (2 * dA), sB = 1 << (2 * dB);
                const int base = (tt % sA) + (tt / sA) * (sA * 16);
                // ---- neighbour traces tv[which][side][x]
                T tv[2][2][4];
#pragma unroll
                for (int which = 0; which < 2; ++which)
                  {
                    const int d = which ? dB : dA, sd = which ? sB : sA, so = which ? sA : sB;
#pragma unroll
                    for (int side = 0; side < 2; ++side)
                      {
#pragma unroll
                        for (int x = 0; x < 4; ++x)
                          tv[which][side][x] = T(0);
                        if (!((p.nb_mask[d] >> side) & 1))
                          continue;
                        const bool at_edge = side ? (c[d] == p.ncell[d] - 1) : (c[d] == 0);
                        const int  layer   = side ? 0 : N - 1;
                        if (d == 0 && step > 0)
                          {
                            // the previous cell of the walk is still in shared memory
                            const int o = base + layer * sd;
#pragma unroll
                            for (int x = 0; x < 4; ++x)
                              tv[which][side][x] = un[padded(o + x * so)];
                          }
                        else if (at_edge && p.side_kind[d][side] == HD_SIDE_GHOST)
                          {
                            long long fc = 0, m = 1;
#pragma unroll
                            for (int e = 0; e < DIM; ++e)
                              if (e != d)
                                {
                                  fc += c[e] * m;
                                  m *= p.ncell[e];
                                }
                            const T *g = p.ghost + p.ghost_off[d][side] + fc * (ND / N);
#pragma unroll
                            for (int x = 0; x < 4; ++x)
                              {
                                const int o        = base + x * so;
                                tv[which][side][x] = __ldg(g + (o % sd) + (o / (sd * N)) * sd);
                              }
                          }
                        else
                          {
                            long long nb = cell + (side ? cstr[d] : -cstr[d]);
                            if (at_edge)
                              nb = cell + (side ? -(long long)(p.ncell[d] - 1) * cstr[d] : (long long)(p.ncell[d] - 1) * cstr[d]);
                            const T *g = p.src + nb * ND + base + layer * sd;
#pragma unroll
                            for (int x = 0; x < 4; ++x)
                              tv[which][side][x] = __ldg(g + x * so);
                          }
                      }
                  }
                // ---- the tile and the partial sums of the previous round
                T U[4][4], out[4][4];
                if (r == 0)
                  {
                    const int4 *q = reinterpret_cast<const int4 *>(uc + tt * TS);
                    int4        w[V4];
#pragma unroll
                    for (int i = 0; i < V4; ++i)
                      w[i] = q[i];
                    const T *wv = reinterpret_cast<const T *>(w);
#pragma unroll
                    for (int b = 0; b < 4; ++b)
#pragma unroll
                      for (int a = 0; a < 4; ++a)
                        {
                          U[b][a]   = wv[4 * b + a];
                          out[b][a] = T(0);
                        }
                  }
                else
                  {
#pragma unroll
                    for (int b = 0; b < 4; ++b)
#pragma unroll
                      for (int a = 0; a < 4; ++a)
                        {
                          const int i = padded(base + a * sA + b * sB);
                          U[b][a]     = uc[i];
                          out[b][a]   = part[i];
                        }
                  }
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                  for (int a = 0; a < 4; ++a)
                    {
                      T v = out[b][a];
#pragma unroll
                      for (int j = 0; j < 4; ++j)
                        v += cf.C[dA][a * 4 + j] * U[b][j];
#pragma unroll
                      for (int j = 0; j < 4; ++j)
                        v += cf.C[dB][b * 4 + j] * U[j][a];
                      out[b][a] = v;
                    }
#pragma unroll
                for (int side = 0; side < 2; ++side)
                  {
                    const T *LA = side ? cf.L1[dA] : cf.L0[dA];
                    const T *LB = side ? cf.L1[dB] : cf.L0[dB];
#pragma unroll
                    for (int b = 0; b < 4; ++b)
#pragma unroll
                      for (int a = 0; a < 4; ++a)
                        out[b][a] += LA[a] * tv[0][side][b] + LB[b] * tv[1][side][a];
                  }
                if (r == DIM / 2 - 1)
                  {
                    const long long g = cell * ND + base;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
#pragma unroll
                      for (int a = 0; a < 4; ++a)
                        {
                          const long long i = g + a * sA + b * sB;
                          if (p.fused)
                            {
                              const T s = p.sol[i];
                              p.sol[i]  = s + p.fb * out[b][a];
                              if (p.fa != T(0))
                                p.ti_next[i] = s + p.fa * out[b][a];
                            }
                          else
                            p.dst[i] = out[b][a];
                        }
                    // the next cell has landed, `part` and the old cell buffer are free again
                    cp_async_wait_all();
                    __syncthreads();
                  }
                else
                  {
#pragma unroll
                    for (int b = 0; b < 4; ++b)
#pragma unroll
                      for (int a = 0; a < 4; ++a)
                        part[padded(base + a * sA + b * sB)] = out[b][a];
                    __syncthreads();
                    // round 0 was the last reader of the previous cell: its buffer takes the next cell of the walk
                    if (r == 0 && step + 1 < n0)
                      stage(un, row_cell0 + (descend ? n0 - 2 - step : step + 1));
                  }
              }
          }
      }
  }

  template <typename T>
  int
  launch_tile_row_t(hd_advection *op, void *dst, const void *src, const void *ghosts, const FusedUpdate &fu)
  {
    hd_mesh *       m   = op->mesh;
    constexpr int   NDP = 256 * (16 + 16 / (int)sizeof(T));
    GenParams<T, 6> p;
    TileCoef<T>     cf;
    p.src   = static_cast<const T *>(src);
    p.dst   = static_cast<T *>(dst);
    p.ghost = static_cast<const T *>(ghosts);
    p.coef  = nullptr;
    for (int d = 0; d < 6; ++d)
      {
        for (int i = 0; i < 16; ++i)
          cf.C[d][i] = T(op->hC[d][0][i]);
        for (int i = 0; i < 4; ++i)
          {
            cf.L0[d][i] = T(op->hL0[d][i]);
            cf.L1[d][i] = T(op->hL1[d][i]);
          }
        p.ncell[d]        = m->d.n_cells[d];
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
    const size_t smem = (size_t)3 * NDP * sizeof(T);
    auto         kern = k_apply_tile_row<T>;
    HD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long nrows = m->ncells / m->d.n_cells[0];
    const long long slots = 2ll * m->ctx->sm_count;
    const long long grid  = nrows < slots ? nrows : slots;
    kern<<<(unsigned)grid, 256, smem, m->ctx->stream>>>(p, cf);
    HD_CUDA(cudaGetLastError());
    op->launches++;
    op->last_kernel = "tile_row";
    return HD_OK;
  }

  template <typename T, int DIM, int THREADS>
  int
  launch_tile_t(hd_advection *op, void *dst, const void *src, const void *ghosts, const FusedUpdate &fu)
  {
    hd_mesh *         m   = op->mesh;
    constexpr int     ND  = IPow<4, DIM>::value;
    constexpr int     CPB = THREADS / (ND / 16);
    constexpr int     NDP = (ND / 16) * (16 + 16 / (int)sizeof(T));
    GenParams<T, DIM> p;
    TileCoef<T>       cf;
    p.src   = static_cast<const T *>(src);
    p.dst   = static_cast<T *>(dst);
    p.ghost = static_cast<const T *>(ghosts);
    p.coef  = nullptr;
    for (int d = 0; d < 6; ++d)
      for (int i = 0; i < 16; ++i)
        {
          cf.C[d][i] = d < DIM ? T(op->hC[d][0][i]) : T(0);
          if (i < 4)
            {
              cf.L0[d][i] = d < DIM ? T(op->hL0[d][i]) : T(0);
              cf.L1[d][i] = d < DIM ? T(op->hL1[d][i]) : T(0);
            }
        }
    for (int d = 0; d < DIM; ++d)
      {
        p.ncell[d]        = m->d.n_cells[d];
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
    const size_t smem = (size_t)(DIM > 2 ? 2 : 1) * CPB * NDP * sizeof(T);
    auto         kern = k_apply_tile<T, DIM, THREADS>;
    if (smem > 48 * 1024)
      HD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = (m->ncells + CPB - 1) / CPB;
    kern<<<(unsigned)grid, THREADS, smem, m->ctx->stream>>>(p, cf);
    HD_CUDA(cudaGetLastError());
    op->launches++;
    op->last_kernel = "tile";
    return HD_OK;
  }

  template <typename T>
  int
  launch_n(hd_advection *op, void *dst, const void *src, const void *ghosts, const FusedUpdate &fu)
  {
    switch (op->mesh->n)
      {
        case 2:
          return launch_dim<T, 2>(op, dst, src, ghosts, fu);
        case 3:
          return launch_dim<T, 3>(op, dst, src, ghosts, fu);
        case 4:
          return launch_dim<T, 4>(op, dst, src, ghosts, fu);
        case 5:
          return launch_dim<T, 5>(op, dst, src, ghosts, fu);
        case 6:
          return launch_dim<T, 6>(op, dst, src, ghosts, fu);
      }
    return hd::fail(HD_ERR_UNSUPPORTED, "degree must be in 1..5");
  }
} // namespace

namespace hd
{
  // the tile kernel covers degree 3 in 1D1V, 2D2V and 3D3V without Dirichlet sides (those keep the generic kernel, whose
  // matrices come in four boundary variants)
  bool
  tile_preferred(const hd_advection *op)
  {
    // automatic choice: 2D2V, and the 3D3V cases the pipelined kernel does not take (FP32) — in 1D1V (16 values per cell)
    // the generic kernel is the faster one (170 vs 158 GDoF/s)
    return tile_supported(op) && op->mesh->dim >= 4;
  }

  bool
  tile_supported(const hd_advection *op)
  {
    const hd_mesh *m = op->mesh;
    if (m->n != 4 || (m->dim != 2 && m->dim != 4 && m->dim != 6))
      return false;
    for (int d = 0; d < m->dim; ++d)
      for (int s = 0; s < 2; ++s)
        if (m->d.side_kind[d][s] == HD_SIDE_DIRICHLET || m->d.side_kind[d][s] == HD_SIDE_DIRICHLET_HOM)
          return false;
    return true;
  }

  // the row-persistent variant: 3D3V, and the upwind neighbour along x_0 must be on one side only (it always is for the
  // upwind flux; a zero x_0 velocity needs no neighbour at all)
  bool
  tile_row_supported(const hd_advection *op)
  {
    return tile_supported(op) && op->mesh->dim == 6 && op->nb_mask[0] != 3;
  }

  int
  launch_tile_row(hd_advection *op, void *dst, const void *src, const void *ghosts, double, const FusedUpdate &fu)
  {
    return op->mesh->d.number_type == HD_F64 ? launch_tile_row_t<double>(op, dst, src, ghosts, fu) : launch_tile_row_t<float>(op, dst, src, ghosts, fu);
  }

  int
  launch_tile(hd_advection *op, void *dst, const void *src, const void *ghosts, double, const FusedUpdate &fu)
  {
    const bool f64 = op->mesh->d.number_type == HD_F64;
    if (op->mesh->dim == 2)
      return f64 ? launch_tile_t<double, 2, 256>(op, dst, src, ghosts, fu) : launch_tile_t<float, 2, 256>(op, dst, src, ghosts, fu);
    if (op->mesh->dim == 6) // one cell (256 tiles per round) per CTA, three rounds
      return f64 ? launch_tile_t<double, 6, 256>(op, dst, src, ghosts, fu) : launch_tile_t<float, 6, 256>(op, dst, src, ghosts, fu);
    // CTA size (cells per CTA = threads / 16): 128 threads = 8 cells, 36 KiB of shared memory, 6 CTAs per SM measured best
    // (228 vs 211 GDoF/s with 256, 149 with 512); HD_TILE_THREADS = 64 | 128 | 256 | 512 for experiments
    static const int threads = [] {
      const char *e = getenv("HD_TILE_THREADS");
      return e ? atoi(e) : 128;
    }();
    if (threads == 64)
      return f64 ? launch_tile_t<double, 4, 64>(op, dst, src, ghosts, fu) : launch_tile_t<float, 4, 64>(op, dst, src, ghosts, fu);
    if (threads == 128)
      return f64 ? launch_tile_t<double, 4, 128>(op, dst, src, ghosts, fu) : launch_tile_t<float, 4, 128>(op, dst, src, ghosts, fu);
    if (threads == 512)
      return f64 ? launch_tile_t<double, 4, 512>(op, dst, src, ghosts, fu) : launch_tile_t<float, 4, 512>(op, dst, src, ghosts, fu);
    return f64 ? launch_tile_t<double, 4, 256>(op, dst, src, ghosts, fu) : launch_tile_t<float, 4, 256>(op, dst, src, ghosts, fu);
  }

  int
  launch_generic(hd_advection *op, void *dst, const void *src, const void *ghosts, double, const FusedUpdate &fu)
  {
    if (op->mesh->d.number_type == HD_F64)
      return launch_n<double>(op, dst, src, ghosts, fu);
    return launch_n<float>(op, dst, src, ghosts, fu);
  }

  // ghost_mesh == nullptr: lift the boundary data into dst (generic-kernel path).  Otherwise: fill the ghost segments of
  // `ghost_mesh` (the same lattice with its inflow Dirichlet sides declared HD_SIDE_GHOST) in `ghosts` from src and g.
  template <typename T>
  static int
  dirichlet_t(hd_advection *op, void *dst, double time, const FusedUpdate &fu, const hd_mesh *ghost_mesh = nullptr, void *ghosts = nullptr, const void *src = nullptr)
  {
    hd_mesh *m = op->mesh;
    for (int d = 0; d < m->dim; ++d)
      for (int side = 0; side < 2; ++side)
        {
          const bool hom = m->d.side_kind[d][side] == HD_SIDE_DIRICHLET_HOM;
          if (m->d.side_kind[d][side] != HD_SIDE_DIRICHLET && !(ghost_mesh && hom))
            continue;
          if (!((op->nb_mask[d] >> side) & 1))
            continue; // outflow side: beta = 0
          DirParams<T> p;
          p.dst   = static_cast<T *>(dst);
          p.g     = op->d_g[d][side];
          if (!hom && !p.g && op->dirichlet_fn < 0)
            return hd::fail(HD_ERR_INVALID, "inhomogeneous Dirichlet side without boundary data");
          p.homogeneous   = hom ? 1 : 0;
          p.src_for_ghost = static_cast<const T *>(src);
          p.ghost_out     = ghost_mesh ? static_cast<T *>(ghosts) + ghost_mesh->ghost_off[d][side] : nullptr;
          p.basis = m->d_basis;
          p.lift  = reinterpret_cast<const char *>(op->d_coef) + op->coef_bytes + sizeof(double) * m->n * (2 * d + side);
          p.dim   = m->dim;
          p.n     = m->n;
          p.nq    = m->nq;
          p.dir   = d;
          p.side  = side;
          p.fn_id = op->dirichlet_fn;
          long long nfc = 1;
          for (int e = 0; e < HD_MAX_DIM; ++e)
            {
              p.ncell[e]       = e < m->dim ? m->d.n_cells[e] : 1;
              p.cell_offset[e] = e < m->dim ? m->d.cell_offset[e] : 0;
              p.left[e]        = m->d.left[e];
              p.h[e]           = m->h[e];
              if (e < m->dim && e != d)
                nfc *= m->d.n_cells[e];
            }
          p.time    = time;
          p.sol     = static_cast<T *>(fu.sol);
          p.ti_next = static_cast<T *>(fu.ti_next);
          p.fb      = fu.fb;
          p.fa      = fu.fa;
          p.fused   = fu.enabled;
          int mx = m->n > m->nq ? m->n : m->nq, cap = 1;
          for (int e = 0; e < m->dim - 1; ++e)
            cap *= mx;
          const size_t smem = 2 * sizeof(double) * cap + 2 * sizeof(double) * HD_MAX_DIM * mx;
          if (smem > 48 * 1024)
            {
              if (smem > m->ctx->smem_optin)
                return hd::fail(HD_ERR_UNSUPPORTED, "face does not fit into shared memory");
              HD_CUDA(cudaFuncSetAttribute(k_dirichlet_source<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            }
          // one CTA per boundary face cell; as many threads as the face has quadrature points / nodal values (64 in 2D2V,
          // 1024 in 3D3V), at most 256
          long long work = 1;
          for (int e = 0; e < m->dim - 1; ++e)
            work *= mx;
          const int threads = work >= 256 ? 256 : int((work + 31) / 32 * 32);
          k_dirichlet_source<T><<<(unsigned)nfc, threads, smem, m->ctx->stream>>>(p);
          HD_CUDA(cudaGetLastError());
          op->launches++;
        }
    return HD_OK;
  }

  int
  launch_dirichlet_source(hd_advection *op, void *dst, double time, const FusedUpdate &fu)
  {
    if (op->mesh->d.number_type == HD_F64)
      return dirichlet_t<double>(op, dst, time, fu);
    return dirichlet_t<float>(op, dst, time, fu);
  }

  int
  launch_dirichlet_ghosts(hd_advection *op, const hd_mesh *ghost_mesh, void *ghosts, const void *src, double time)
  {
    FusedUpdate fu;
    if (op->mesh->d.number_type == HD_F64)
      return dirichlet_t<double>(op, nullptr, time, fu, ghost_mesh, ghosts, src);
    return dirichlet_t<float>(op, nullptr, time, fu, ghost_mesh, ghosts, src);
  }
} // namespace hd
