// x-space field solve of the Vlasov-Poisson right-hand side on a periodic Cartesian x-lattice:
//   rhs = -M (rho - mean) , mean removed again          examples/vlasov_poisson/include/application.h:529-565
//   K phi = rhs, K = symmetric-interior-penalty DG Laplacian   examples/vlasov_poisson/include/poisson.h:166-250
//           (cell term (grad u, grad v); faces -{d_n u}[v] - [u]{d_n v} + sigma [u][v], sigma = (1/h + 1/h) k (k+1), :226-231)
//   a_v = grad(phi) at the quadrature points of every x-cell   examples/vlasov_poisson/include/derivative_container.h:157-190
// solved by conjugate gradients on the device (the reference: CG with a Chebyshev or multigrid preconditioner, relative
// residual 1e-7, poisson.h:575-610).  On a Cartesian lattice every integrand is integrated exactly by the (k+1)-point Gauss
// rule, so K = sum_d (mass in the other directions) (x) (1-D SIP operator along d) with three (k+1)x(k+1) blocks per direction.
//
// The per-cell bodies are checked on the CPU against the oracle's dense operator (tests/test_poisson_emulation.py, via
// tests/vp_emulation_harness.cpp) and on the GPU through the Vlasov-Poisson right-hand side and the reference's Landau-damping
// golden (tests/test_zz_vp_device_gpu.py).  Reachable only through hd_poisson_*.
#ifdef HD_VP_HOST_EMULATION
#  ifndef HD_MAX_DIM
#    include <cmath>
#    include <cstddef>
#    include <stdexcept>
#    include <string>
#    include <vector>
#    define HD_MAX_DIM 6
#    include "basis.hpp"
#  endif
#  define HD_XS_FN inline
#  define HD_XS_SYNC() ((void)0)
#else
#  include "hd_internal.h"
#  define HD_XS_FN __device__
#  define HD_XS_SYNC() __syncthreads()
#endif

namespace
{
  struct XsParams
  {
    const double *coef;  // per x-direction: Kc[n*n], Klo[n*n], Khi[n*n], M1[n*n], Gq[nq*n] (d/dx of the nodal basis at the q-points)
    const double *basis; // nodes[n], xq[nq], w[nq], S[nq*n], Sinv[n*nq]
    int           dim_x, n, nq;
    int           ncell[3];
    long long     nd, ncells; // n^dim_x, number of x-cells
  };

  HD_XS_FN int
  xs_coef_block(const int n, const int nq)
  {
    return 4 * n * n + nq * n;
  }

  // out = sum_c Mat[r*cols + c] in(.., c, ..) along the direction with the given stride; in/out hold `total_in / cols * rows` values
  HD_XS_FN void
  xs_sweep(const double *in, double *out, const double *Mat, int rows, int cols, long long stride, long long n_outer, int tid, int nthr)
  {
    const long long total = n_outer * rows * stride;
    for (long long oi = tid; oi < total; oi += nthr)
      {
        const long long lo = oi % stride, rest = oi / stride;
        const int       r  = int(rest % rows);
        const long long hi = rest / rows;
        double          acc = 0.0;
        for (int c = 0; c < cols; ++c)
          acc += Mat[r * cols + c] * in[(hi * cols + c) * stride + lo];
        out[oi] = acc;
      }
  }

  // dst_cell = scale * (M1 (x) .. (x) M1) src_cell           (mass matrix of one x-cell, exact Gauss quadrature)
  HD_XS_FN void
  xs_mass_cell(const XsParams &p, double *sm, const double *src, double *dst, const double scale, const long long cell, const int tid, const int nthr)
  {
    const int n = p.n, blk = xs_coef_block(n, p.nq);
    double *  a = sm, *b = sm + p.nd;
    for (long long i = tid; i < p.nd; i += nthr)
      a[i] = src[cell * p.nd + i];
    HD_XS_SYNC();
    long long stride = 1;
    for (int d = 0; d < p.dim_x; ++d)
      {
        long long outer = 1;
        for (int k = d + 1; k < p.dim_x; ++k)
          outer *= n;
        xs_sweep(a, b, p.coef + (size_t)d * blk + 3 * n * n, n, n, stride, outer, tid, nthr);
        HD_XS_SYNC();
        double *t = a;
        a         = b;
        b         = t;
        stride *= n;
      }
    for (long long i = tid; i < p.nd; i += nthr)
      dst[cell * p.nd + i] = scale * a[i];
  }

  // dst_cell = (K src)_cell: per direction the three-block line operator (own cell, lower and upper neighbour, periodic), then the
  // mass matrix along the other directions
  HD_XS_FN void
  xs_laplace_cell(const XsParams &p, double *sm, const double *src, double *dst, const long long cell, const int tid, const int nthr)
  {
    const int n = p.n, blk = xs_coef_block(n, p.nq);
    double *  acc = sm, *a = sm + p.nd, *b = a + p.nd;
    int       c[3];
    long long cstr[3];
    {
      long long r = cell, m = 1;
      for (int d = 0; d < p.dim_x; ++d)
        {
          c[d]    = int(r % p.ncell[d]);
          r /= p.ncell[d];
          cstr[d] = m;
          m *= p.ncell[d];
        }
    }
    for (long long i = tid; i < p.nd; i += nthr)
      acc[i] = 0.0;
    HD_XS_SYNC();
    for (int d = 0; d < p.dim_x; ++d)
      {
        const double *Kc = p.coef + (size_t)d * blk, *Klo = Kc + n * n, *Khi = Klo + n * n;
        long long     stride = 1;
        for (int k = 0; k < d; ++k)
          stride *= n;
        const long long nb_lo = c[d] == 0 ? cell + (long long)(p.ncell[d] - 1) * cstr[d] : cell - cstr[d];
        const long long nb_hi = c[d] == p.ncell[d] - 1 ? cell - (long long)(p.ncell[d] - 1) * cstr[d] : cell + cstr[d];
        double *        cur   = a, *nxt = b;
        for (long long i = tid; i < p.nd; i += nthr)
          {
            const int       id   = int((i / stride) % n);
            const long long base = i - id * stride;
            double          s    = 0.0;
            for (int j = 0; j < n; ++j)
              s += Kc[id * n + j] * src[cell * p.nd + base + j * stride] + Klo[id * n + j] * src[nb_lo * p.nd + base + j * stride] +
                   Khi[id * n + j] * src[nb_hi * p.nd + base + j * stride];
            cur[i] = s;
          }
        HD_XS_SYNC();
        long long se = 1;
        for (int e = 0; e < p.dim_x; ++e)
          {
            if (e != d)
              {
                long long outer = 1;
                for (int k = e + 1; k < p.dim_x; ++k)
                  outer *= n;
                xs_sweep(cur, nxt, p.coef + (size_t)e * blk + 3 * n * n, n, n, se, outer, tid, nthr);
                HD_XS_SYNC();
                double *t = cur;
                cur       = nxt;
                nxt       = t;
              }
            se *= n;
          }
        for (long long i = tid; i < p.nd; i += nthr)
          acc[i] += cur[i];
        HD_XS_SYNC();
      }
    for (long long i = tid; i < p.nd; i += nthr)
      dst[cell * p.nd + i] = acc[i];
  }

  // a_v[cell][q][d] = d phi / d x_d at the quadrature points of the cell: Gq along d, S along the other directions
  HD_XS_FN void
  xs_gradient_cell(const XsParams &p, double *sm, const double *phi, double *a_v, const long long cell, const int tid, const int nthr)
  {
    const int     n = p.n, nq = p.nq, blk = xs_coef_block(n, nq);
    const double *S = p.basis + n + 2 * nq;
    long long     cap = 1, nqx = 1;
    for (int d = 0; d < p.dim_x; ++d)
      {
        cap *= n > nq ? n : nq;
        nqx *= nq;
      }
    double *u = sm, *a = sm + cap, *b = a + cap;
    for (long long i = tid; i < p.nd; i += nthr)
      u[i] = phi[cell * p.nd + i];
    HD_XS_SYNC();
    for (int d = 0; d < p.dim_x; ++d)
      {
        const double *cur = u;
        double *      nxt = a;
        long long     lo  = 1;
        for (int e = 0; e < p.dim_x; ++e)
          {
            long long outer = 1;
            for (int k = e + 1; k < p.dim_x; ++k)
              outer *= n;
            xs_sweep(cur, nxt, e == d ? p.coef + (size_t)d * blk + 4 * n * n : S, nq, n, lo, outer, tid, nthr);
            HD_XS_SYNC();
            cur = nxt;
            nxt = (nxt == a) ? b : a;
            lo *= nq;
          }
        for (long long q = tid; q < nqx; q += nthr)
          a_v[(cell * nqx + q) * p.dim_x + d] = cur[q];
        HD_XS_SYNC();
      }
  }

  // host: coefficient blocks of every x-direction (see XsParams::coef), from the product's own 1-D basis
  inline void
  xs_coefficients(const hd::Basis1D &b, const int dim_x, const double *h, std::vector<double> &out)
  {
    using hd::LD;
    const int n = b.n, nq = b.nq, blk = 4 * n * n + nq * n;
    out.assign((size_t)dim_x * blk, 0.0);
    // G[q][i] = derivative of nodal basis function i at quadrature point q (reference cell) = (D S)[q][i]
    std::vector<LD> G(nq * n, 0);
    for (int q = 0; q < nq; ++q)
      for (int i = 0; i < n; ++i)
        for (int r = 0; r < nq; ++r)
          G[q * n + i] += b.D[q * nq + r] * b.S[r * n + i];
    // derivative of the nodal basis at the two ends: interpolate the derivative from the quadrature points (exact)
    std::vector<LD> g0(n, 0), g1(n, 0);
    for (int i = 0; i < n; ++i)
      for (int q = 0; q < nq; ++q)
        {
          g0[i] += b.f0[q] * G[q * n + i];
          g1[i] += b.f1[q] * G[q * n + i];
        }
    for (int d = 0; d < dim_x; ++d)
      {
        const LD hd_ = h[d], sigma = (LD)2 / hd_ * (LD)(n - 1 > 1 ? n - 1 : 1) * (LD)n;
        double * o   = out.data() + (size_t)d * blk;
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j)
            {
              LD kc = 0, m1 = 0;
              for (int q = 0; q < nq; ++q)
                {
                  kc += G[q * n + i] * b.w[q] * G[q * n + j] / hd_;
                  m1 += b.S[q * n + i] * b.w[q] * b.S[q * n + j] * hd_;
                }
              // end-node indicator vectors of the Gauss-Lobatto nodal basis: v0 = e_0, v1 = e_{n-1}; end derivatives g / h
              const LD v0i = i == 0, v0j = j == 0, v1i = i == n - 1, v1j = j == n - 1;
              const LD d0i = g0[i] / hd_, d0j = g0[j] / hd_, d1i = g1[i] / hd_, d1j = g1[j] / hd_;
              // face at xi = 1 (this cell on the minus side) and face at xi = 0 (this cell on the plus side)
              kc += -v1i * (d1j / 2) - (d1i / 2) * v1j + sigma * v1i * v1j;
              kc += v0i * (d0j / 2) + (d0i / 2) * v0j + sigma * v0i * v0j;
              const LD khi = -v1i * (d0j / 2) + (d1i / 2) * v0j - sigma * v1i * v0j; // rows: this cell, columns: upper neighbour
              const LD klo = v0i * (d1j / 2) - (d0i / 2) * v1j - sigma * v0i * v1j;  // rows: this cell, columns: lower neighbour
              o[i * n + j]             = (double)kc;
              o[n * n + i * n + j]     = (double)klo;
              o[2 * n * n + i * n + j] = (double)khi;
              o[3 * n * n + i * n + j] = (double)m1;
            }
        for (int q = 0; q < nq; ++q)
          for (int i = 0; i < n; ++i)
            o[4 * n * n + q * n + i] = (double)(G[q * n + i] / hd_);
      }
  }
} // namespace

#ifndef HD_VP_HOST_EMULATION
namespace
{
  __global__ void __launch_bounds__(128) k_xs_mass(const XsParams p, const double *src, double *dst, double scale)
  {
    extern __shared__ double sm[];
    xs_mass_cell(p, sm, src, dst, scale, blockIdx.x, threadIdx.x, blockDim.x);
  }
  __global__ void __launch_bounds__(128) k_xs_laplace(const XsParams p, const double *src, double *dst)
  {
    extern __shared__ double sm[];
    xs_laplace_cell(p, sm, src, dst, blockIdx.x, threadIdx.x, blockDim.x);
  }
  __global__ void __launch_bounds__(128) k_xs_gradient(const XsParams p, const double *phi, double *a_v)
  {
    extern __shared__ double sm[];
    xs_gradient_cell(p, sm, phi, a_v, blockIdx.x, threadIdx.x, blockDim.x);
  }
  // small-vector helpers of the CG iteration (x-space vectors are tiny next to the phase-space ones)
  template <typename T>
  __global__ void k_xs_load(const T *src, double *dst, long long n)
  {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      dst[i] = double(src[i]);
  }
  __global__ void k_xs_dot(const double *a, const double *b, long long n, double *out)
  {
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      s += a[i] * (b ? b[i] : 1.0);
    for (int off = 16; off > 0; off >>= 1)
      s += __shfl_down_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0)
      atomicAdd(out, s);
  }
  // z = D^-1 r with the point-Jacobi diagonal of K (the same nd values in every cell of the uniform lattice)
  __global__ void k_xs_jacobi(const double *r, const double *dinv, double *z, long long n, long long nd)
  {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      z[i] = r[i] * dinv[i % nd];
  }
  // y = a * x + b * y ;  y += c (scalar shift)
  __global__ void k_xs_axpby(double a, const double *x, double b, double *y, double shift, long long n)
  {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      y[i] = a * (x ? x[i] : 0.0) + b * y[i] + shift;
  }
} // namespace

struct hd_poisson
{
  hd_mesh * mesh = nullptr;
  XsParams  p;
  double *  d_coef = nullptr, *d_phi = nullptr, *d_r = nullptr, *d_p = nullptr, *d_ap = nullptr, *d_b = nullptr, *d_scalar = nullptr;
  double *  d_z = nullptr, *d_dinv = nullptr; // preconditioned residual, inverse point-Jacobi diagonal (one cell's worth)
  long long n      = 0;
  size_t    smem_op = 0, smem_grad = 0;
  int       last_iterations = 0;
  double    last_rel_residual = 0.0;
};

namespace
{
  int
  xs_reduce(hd_poisson *ps, const double *a, const double *b, double *host_out)
  {
    cudaStream_t st = ps->mesh->ctx->stream;
    HD_CUDA(cudaMemsetAsync(ps->d_scalar, 0, sizeof(double), st));
    const int blocks = (int)((ps->n + 255) / 256 < 64 ? (ps->n + 255) / 256 : 64);
    k_xs_dot<<<blocks, 256, 0, st>>>(a, b, ps->n, ps->d_scalar);
    HD_CUDA(cudaGetLastError());
    HD_CUDA(cudaMemcpyAsync(host_out, ps->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, st));
    HD_CUDA(cudaStreamSynchronize(st));
    return HD_OK;
  }
  int
  xs_axpby(hd_poisson *ps, double a, const double *x, double b, double *y, double shift = 0.0)
  {
    const int blocks = (int)((ps->n + 255) / 256 < 256 ? (ps->n + 255) / 256 : 256);
    k_xs_axpby<<<blocks, 256, 0, ps->mesh->ctx->stream>>>(a, x, b, y, shift, ps->n);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
  }
} // namespace

extern "C" {

int
hd_poisson_create(hd_mesh *m, hd_poisson **out)
{
  HD_REQUIRE(m && out, "null argument");
  HD_REQUIRE(m->d.dim_x >= 1 && m->d.dim_x <= 3, "x-space must have 1..3 directions");
  for (int d = 0; d < m->d.dim_x; ++d)
    for (int s = 0; s < 2; ++s)
      if (m->d.side_kind[d][s] != HD_SIDE_PERIODIC_LOCAL)
        return hd::fail(HD_ERR_UNSUPPORTED, "the field solve covers periodic single-GPU x-lattices only");
  hd_poisson *ps = new (std::nothrow) hd_poisson;
  HD_REQUIRE(ps, "out of memory");
  ps->mesh     = m;
  XsParams &p  = ps->p;
  p.dim_x      = m->d.dim_x;
  p.n          = m->n;
  p.nq         = m->nq;
  p.nd         = 1;
  p.ncells     = 1;
  long long cap = 1;
  for (int d = 0; d < 3; ++d)
    {
      p.ncell[d] = d < p.dim_x ? m->d.n_cells[d] : 1;
      if (d < p.dim_x)
        {
          p.nd *= p.n;
          p.ncells *= p.ncell[d];
          cap *= p.n > p.nq ? p.n : p.nq;
        }
    }
  ps->n         = p.nd * p.ncells;
  ps->smem_op   = 3 * (size_t)p.nd * sizeof(double);
  ps->smem_grad = 3 * (size_t)cap * sizeof(double);
  std::vector<double> coef;
  xs_coefficients(m->basis, p.dim_x, m->h, coef);
  HD_CUDA(cudaSetDevice(m->ctx->device));
  HD_CUDA(cudaMalloc(&ps->d_coef, coef.size() * sizeof(double)));
  HD_CUDA(cudaMemcpy(ps->d_coef, coef.data(), coef.size() * sizeof(double), cudaMemcpyHostToDevice));
  {
    // point-Jacobi diagonal: K = sum_d (1-D SIP block along d) (x) (1-D mass along the others), poisson.h:581-590 takes the
    // same diagonal (compute_inverse_diagonal) as the smoother's preconditioner
    const int           n = p.n, blk = 4 * p.n * p.n + p.nq * p.n; // (= xs_coef_block, a device function)
    std::vector<double> dinv((size_t)p.nd);
    for (long long i = 0; i < p.nd; ++i)
      {
        int idx[3] = {0, 0, 0};
        long long r = i;
        for (int d = 0; d < p.dim_x; ++d)
          {
            idx[d] = int(r % n);
            r /= n;
          }
        double diag = 0.0;
        for (int d = 0; d < p.dim_x; ++d)
          {
            double t = coef[(size_t)d * blk + idx[d] * n + idx[d]];
            for (int e = 0; e < p.dim_x; ++e)
              if (e != d)
                t *= coef[(size_t)e * blk + 3 * n * n + idx[e] * n + idx[e]];
            diag += t;
          }
        dinv[(size_t)i] = 1.0 / diag;
      }
    HD_CUDA(cudaMalloc(&ps->d_dinv, dinv.size() * sizeof(double)));
    HD_CUDA(cudaMemcpy(ps->d_dinv, dinv.data(), dinv.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  double **vecs[] = {&ps->d_phi, &ps->d_r, &ps->d_p, &ps->d_ap, &ps->d_b, &ps->d_z};
  for (double **v : vecs)
    {
      HD_CUDA(cudaMalloc(v, (size_t)ps->n * sizeof(double)));
      HD_CUDA(cudaMemset(*v, 0, (size_t)ps->n * sizeof(double)));
    }
  HD_CUDA(cudaMalloc(&ps->d_scalar, sizeof(double)));
  p.coef  = ps->d_coef;
  p.basis = m->d_basis;
  *out    = ps;
  return HD_OK;
}

int
hd_poisson_destroy(hd_poisson *ps)
{
  if (!ps)
    return HD_OK;
  cudaFree(ps->d_coef);
  cudaFree(ps->d_phi);
  cudaFree(ps->d_r);
  cudaFree(ps->d_p);
  cudaFree(ps->d_ap);
  cudaFree(ps->d_b);
  cudaFree(ps->d_scalar);
  cudaFree(ps->d_z);
  cudaFree(ps->d_dinv);
  delete ps;
  return HD_OK;
}

int
hd_poisson_solve(hd_poisson *ps, const void *rho_x, double *a_v_device, double rel_tol, int max_iterations, int *iterations)
{
  HD_REQUIRE(ps && rho_x && a_v_device && rel_tol > 0 && max_iterations > 0, "bad argument");
  hd_mesh *    m  = ps->mesh;
  cudaStream_t st = m->ctx->stream;
  HD_CUDA(cudaSetDevice(m->ctx->device));
  const XsParams &p      = ps->p;
  const unsigned  ncells = (unsigned)p.ncells;
  const double    invn   = 1.0 / (double)ps->n;
  int             rc;
  double          s;
  // rho in double, minus its mean (application.h:531-533)
  const int blocks = (int)((ps->n + 255) / 256 < 256 ? (ps->n + 255) / 256 : 256);
  if (m->d.number_type == HD_F64)
    k_xs_load<double><<<blocks, 256, 0, st>>>(static_cast<const double *>(rho_x), ps->d_r, ps->n);
  else
    k_xs_load<float><<<blocks, 256, 0, st>>>(static_cast<const float *>(rho_x), ps->d_r, ps->n);
  HD_CUDA(cudaGetLastError());
  if ((rc = xs_reduce(ps, ps->d_r, nullptr, &s)) != HD_OK)
    return rc;
  if ((rc = xs_axpby(ps, 0.0, nullptr, 1.0, ps->d_r, -s * invn)) != HD_OK)
    return rc;
  // b = -M rho, mean removed again (:536-565)
  k_xs_mass<<<ncells, 128, ps->smem_op, st>>>(p, ps->d_r, ps->d_b, -1.0);
  HD_CUDA(cudaGetLastError());
  if ((rc = xs_reduce(ps, ps->d_b, nullptr, &s)) != HD_OK)
    return rc;
  if ((rc = xs_axpby(ps, 0.0, nullptr, 1.0, ps->d_b, -s * invn)) != HD_OK)
    return rc;
  // Jacobi-preconditioned CG on K phi = b, starting from the previous potential (poisson.h:575-603: SolverCG with a
  // Chebyshev smoother over the same diagonal as preconditioner, relative residual 1e-7, at most 2 * size iterations)
  k_xs_laplace<<<ncells, 128, ps->smem_op, st>>>(p, ps->d_phi, ps->d_ap);
  HD_CUDA(cudaGetLastError());
  HD_CUDA(cudaMemcpyAsync(ps->d_r, ps->d_b, (size_t)ps->n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if ((rc = xs_axpby(ps, -1.0, ps->d_ap, 1.0, ps->d_r)) != HD_OK)
    return rc;
  auto precondition = [&]() -> int {
    k_xs_jacobi<<<blocks, 256, 0, st>>>(ps->d_r, ps->d_dinv, ps->d_z, ps->n, p.nd);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
  };
  if ((rc = precondition()) != HD_OK)
    return rc;
  HD_CUDA(cudaMemcpyAsync(ps->d_p, ps->d_z, (size_t)ps->n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  double rr, bb, rz;
  if ((rc = xs_reduce(ps, ps->d_r, ps->d_r, &rr)) != HD_OK || (rc = xs_reduce(ps, ps->d_b, ps->d_b, &bb)) != HD_OK || (rc = xs_reduce(ps, ps->d_r, ps->d_z, &rz)) != HD_OK)
    return rc;
  const double target = rel_tol * rel_tol * (bb > 0 ? bb : 1.0);
  int          it     = 0;
  bool         broke  = false;
  while (rr > target && it < max_iterations)
    {
      k_xs_laplace<<<ncells, 128, ps->smem_op, st>>>(p, ps->d_p, ps->d_ap);
      HD_CUDA(cudaGetLastError());
      double pap;
      if ((rc = xs_reduce(ps, ps->d_p, ps->d_ap, &pap)) != HD_OK)
        return rc;
      if (!(pap > 0) || !(rz > 0))
        {
          broke = true; // breakdown or NaN
          break;
        }
      const double alpha = rz / pap;
      if ((rc = xs_axpby(ps, alpha, ps->d_p, 1.0, ps->d_phi)) != HD_OK || (rc = xs_axpby(ps, -alpha, ps->d_ap, 1.0, ps->d_r)) != HD_OK)
        return rc;
      if ((rc = precondition()) != HD_OK)
        return rc;
      double rz_new;
      if ((rc = xs_reduce(ps, ps->d_r, ps->d_r, &rr)) != HD_OK || (rc = xs_reduce(ps, ps->d_r, ps->d_z, &rz_new)) != HD_OK)
        return rc;
      if ((rc = xs_axpby(ps, 1.0, ps->d_z, rz_new / rz, ps->d_p)) != HD_OK)
        return rc;
      rz = rz_new;
      ++it;
    }
  ps->last_iterations   = it;
  ps->last_rel_residual = std::sqrt(rr / (bb > 0 ? bb : 1.0));
  if (iterations)
    *iterations = it;
  // the reference's SolverCG throws when the tolerance is not reached (dealii::SolverControl::NoConvergence): no field table is
  // written from a potential that did not converge
  if (broke || !(rr <= target))
    return hd::fail(HD_ERR_NO_CONVERGENCE, "hd_poisson_solve: CG " + std::string(broke ? "broke down" : "did not converge") + " after " + std::to_string(it) +
                                             " iterations, relative residual " + std::to_string(ps->last_rel_residual) + " (tolerance " + std::to_string(rel_tol) + ")");
  // negative electric field = grad(phi) at the quadrature points (derivative_container.h:157-190)
  k_xs_gradient<<<ncells, 128, ps->smem_grad, st>>>(p, ps->d_phi, a_v_device);
  HD_CUDA(cudaGetLastError());
  return HD_OK;
}

// iteration count and achieved relative residual |r| / |b| of the last solve
int
hd_poisson_last_solve(const hd_poisson *ps, int *iterations, double *rel_residual)
{
  HD_REQUIRE(ps, "null argument");
  if (iterations)
    *iterations = ps->last_iterations;
  if (rel_residual)
    *rel_residual = ps->last_rel_residual;
  return HD_OK;
}

// potential of the last solve (device, hd_mesh_n_dofs_x doubles, x-space layout)
const double *
hd_poisson_potential(const hd_poisson *ps)
{
  return ps ? ps->d_phi : nullptr;
}

// sum_q (d_d phi)^2 JxW per x-direction of a gradient table (diagnostics.h:88-143) is host arithmetic on a copy of the table:
// the table has n_cells_x * nq^dim_x * dim_x doubles, tiny next to the phase-space vectors.

} // extern "C"
#endif
