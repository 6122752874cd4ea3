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
    // g at face quadrature points
    for (int q = threadIdx.x; q < nqf; q += blockDim.x)
      {
        double val;
        if (p.g)
          val = p.g[fc * nqf + q];
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
    for (int e = fd - 1; e >= 0; --e)
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
  int
  launch_generic(hd_advection *op, void *dst, const void *src, const void *ghosts, double, const FusedUpdate &fu)
  {
    if (op->mesh->d.number_type == HD_F64)
      return launch_n<double>(op, dst, src, ghosts, fu);
    return launch_n<float>(op, dst, src, ghosts, fu);
  }

  template <typename T>
  static int
  dirichlet_t(hd_advection *op, void *dst, double time, const FusedUpdate &fu)
  {
    hd_mesh *m = op->mesh;
    for (int d = 0; d < m->dim; ++d)
      for (int side = 0; side < 2; ++side)
        {
          if (m->d.side_kind[d][side] != HD_SIDE_DIRICHLET)
            continue;
          if (!((op->nb_mask[d] >> side) & 1))
            continue; // outflow side: beta = 0
          DirParams<T> p;
          p.dst   = static_cast<T *>(dst);
          p.g     = op->d_g[d][side];
          if (!p.g && op->dirichlet_fn < 0)
            return hd::fail(HD_ERR_INVALID, "inhomogeneous Dirichlet side without boundary data");
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
          const size_t smem = 2 * sizeof(double) * cap;
          if (smem > 48 * 1024)
            {
              if (smem > m->ctx->smem_optin)
                return hd::fail(HD_ERR_UNSUPPORTED, "face does not fit into shared memory");
              HD_CUDA(cudaFuncSetAttribute(k_dirichlet_source<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            }
          k_dirichlet_source<T><<<(unsigned)nfc, 256, smem, m->ctx->stream>>>(p);
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
} // namespace hd
