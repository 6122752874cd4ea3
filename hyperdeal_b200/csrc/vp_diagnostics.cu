// Diagnostics of the Vlasov-Poisson driver (examples/vlasov_poisson/include/diagnostics.h):
//   phase_space_diagnostics (:34-86)   mass, squared L2 norm, kinetic energy, momentum of f at the Gauss points
//   compute_electric_energy (:88-143)  sum_q (d_d phi)^2 JxW per x-direction, from the gradient table of hd_poisson_solve
//
// Bodies checked on the CPU through tests/vp_emulation_harness.cpp (tests/test_vp_diagnostics_emulation.py) and on the GPU
// against the oracle's diagnostics during the Landau-damping golden run (tests/vp_step_check.py).
#ifdef HD_VP_HOST_EMULATION
#  ifndef HD_MAX_DIM
#    include <cmath>
#    include <cstddef>
#    include <vector>
#    define HD_MAX_DIM 6
#  endif
#  define HD_DG_FN inline
#  define HD_DG_SYNC() ((void)0)
#else
#  include "hd_internal.h"
#  define HD_DG_FN __device__
#  define HD_DG_SYNC() __syncthreads()
#endif

namespace
{
  struct DiagParams
  {
    const double *basis; // nodes[n], xq[nq], w[nq], S[nq*n], Sinv[n*nq]
    int           dim_x, dim_v, n, nq;
    int           ncell[HD_MAX_DIM], cell_offset[HD_MAX_DIM];
    double        left[HD_MAX_DIM], h[HD_MAX_DIM];
    long long     nd, ncells;
    int           cap; // max(n, nq)^dim
  };

  // partial[0..5] of one cell: sum f JxW, sum f^2 JxW, sum |v|^2 f JxW, sum v_d f JxW (d < 3); sm: 2 * cap + 6 * nthr doubles
  template <typename T>
  HD_DG_FN void
  diag_cell(const DiagParams &p, double *sm, const T *f, double *partial, const long long cell, const int tid, const int nthr)
  {
    const int     dim = p.dim_x + p.dim_v, n = p.n, nq = p.nq;
    const double *xq = p.basis + n, *w = xq + nq, *S = w + nq;
    double *      a = sm, *b = sm + p.cap, *red = b + p.cap;
    int           c[HD_MAX_DIM];
    {
      long long r = cell;
      for (int d = 0; d < dim; ++d)
        {
          c[d] = int(r % p.ncell[d]);
          r /= p.ncell[d];
        }
    }
    for (long long i = tid; i < p.nd; i += nthr)
      a[i] = double(f[cell * p.nd + i]);
    HD_DG_SYNC();
    // S sweeps, direction 0 first: swept directions have extent nq, the others n
    long long lo = 1;
    for (int d = 0; d < dim; ++d)
      {
        long long outer = 1;
        for (int k = d + 1; k < dim; ++k)
          outer *= n;
        const long long total = outer * nq * lo;
        for (long long oi = tid; oi < total; oi += nthr)
          {
            const long long l = oi % lo, rest = oi / lo;
            const int       q = int(rest % nq);
            const long long hi = rest / nq;
            double          s = 0.0;
            for (int k = 0; k < n; ++k)
              s += S[q * n + k] * a[(hi * n + k) * lo + l];
            b[oi] = s;
          }
        HD_DG_SYNC();
        double *t = a;
        a         = b;
        b         = t;
        lo *= nq;
      }
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (long long i = tid; i < lo; i += nthr)
      {
        long long rr = i;
        double    jxw = 1.0, vv = 0.0, v[3] = {0, 0, 0};
        for (int d = 0; d < dim; ++d)
          {
            const int q = int(rr % nq);
            rr /= nq;
            jxw *= p.h[d] * w[q];
            if (d >= p.dim_x)
              {
                const double x = p.left[d] + p.h[d] * ((c[d] + p.cell_offset[d]) + xq[q]);
                v[d - p.dim_x] = x;
                vv += x * x;
              }
          }
        const double fq = a[i];
        s[0] += fq * jxw;
        s[1] += fq * fq * jxw;
        s[2] += vv * fq * jxw;
        for (int d = 0; d < p.dim_v; ++d)
          s[3 + d] += v[d] * fq * jxw;
      }
    for (int k = 0; k < 6; ++k)
      red[k * nthr + tid] = s[k];
    HD_DG_SYNC();
    if (tid == 0)
      for (int k = 0; k < 6; ++k)
        {
          double t = 0.0;
          for (int j = 0; j < nthr; ++j)
            t += red[k * nthr + j];
          partial[k] = t;
        }
  }

  // energy[d] += sum over the x-quadrature points of one x-cell of (a_v[cell][q][d])^2 JxW
  HD_DG_FN void
  field_energy_cell(const DiagParams &p, const double *a_v, double *partial, const long long cell)
  {
    const int     n = p.n, nq = p.nq;
    const double *w = p.basis + n + nq;
    long long     nqx = 1;
    for (int d = 0; d < p.dim_x; ++d)
      nqx *= nq;
    for (int comp = 0; comp < p.dim_x; ++comp)
      {
        double s = 0.0;
        for (long long q = 0; q < nqx; ++q)
          {
            long long rr = q;
            double    jxw = 1.0;
            for (int d = 0; d < p.dim_x; ++d)
              {
                jxw *= p.h[d] * w[rr % nq];
                rr /= nq;
              }
            const double g = a_v[(cell * nqx + q) * p.dim_x + comp];
            s += g * g * jxw;
          }
        partial[comp] = s;
      }
  }
} // namespace

#ifndef HD_VP_HOST_EMULATION
namespace
{
  template <typename T>
  __global__ void __launch_bounds__(128) k_phase_space_diagnostics(const DiagParams p, const T *f, double *out)
  {
    extern __shared__ double sm[];
    __shared__ double        partial[6];
    diag_cell<T>(p, sm, f, partial, blockIdx.x, threadIdx.x, blockDim.x);
    __syncthreads();
    if (threadIdx.x < 6)
      atomicAdd(out + threadIdx.x, partial[threadIdx.x]);
  }
  __global__ void k_field_energy(const DiagParams p, const double *a_v, double *out, long long n_cells_x)
  {
    const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= n_cells_x)
      return;
    double partial[3] = {0, 0, 0};
    field_energy_cell(p, a_v, partial, cell);
    for (int d = 0; d < p.dim_x; ++d)
      atomicAdd(out + d, partial[d]);
  }

  DiagParams
  diag_params(const hd_mesh *m)
  {
    DiagParams p;
    p.basis = m->d_basis;
    p.dim_x = m->d.dim_x;
    p.dim_v = m->d.dim_v;
    p.n     = m->n;
    p.nq    = m->nq;
    for (int d = 0; d < HD_MAX_DIM; ++d)
      {
        p.ncell[d]       = d < m->dim ? m->d.n_cells[d] : 1;
        p.cell_offset[d] = d < m->dim ? m->d.cell_offset[d] : 0;
        p.left[d]        = m->d.left[d];
        p.h[d]           = m->h[d];
      }
    p.nd     = m->nd;
    p.ncells = m->ncells;
    int       mx  = m->n > m->nq ? m->n : m->nq;
    long long cap = 1;
    for (int d = 0; d < m->dim; ++d)
      cap *= mx;
    p.cap = (int)cap;
    return p;
  }
} // namespace

extern "C" {

int
hd_phase_space_diagnostics(hd_mesh *m, const void *vec, double out[6])
{
  HD_REQUIRE(m && vec && out, "null argument");
  HD_REQUIRE(m->d.dim_v <= 3, "at most three velocity directions");
  HD_CUDA(cudaSetDevice(m->ctx->device));
  const DiagParams p    = diag_params(m);
  const size_t     smem = (2 * (size_t)p.cap + 6 * 128) * sizeof(double);
  if (smem > m->ctx->smem_optin)
    return hd::fail(HD_ERR_UNSUPPORTED, "phase_space_diagnostics: cell does not fit into shared memory");
  if (smem > 48 * 1024)
    {
      if (m->d.number_type == HD_F64)
        HD_CUDA(cudaFuncSetAttribute(k_phase_space_diagnostics<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      else
        HD_CUDA(cudaFuncSetAttribute(k_phase_space_diagnostics<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
  // (no early return while d_out is alive)
  double *d_out = nullptr;
  HD_CUDA(cudaMalloc(&d_out, 6 * sizeof(double)));
  cudaError_t e = cudaMemsetAsync(d_out, 0, 6 * sizeof(double), m->ctx->stream);
  if (e == cudaSuccess)
    {
      if (m->d.number_type == HD_F64)
        k_phase_space_diagnostics<double><<<(unsigned)m->ncells, 128, smem, m->ctx->stream>>>(p, static_cast<const double *>(vec), d_out);
      else
        k_phase_space_diagnostics<float><<<(unsigned)m->ncells, 128, smem, m->ctx->stream>>>(p, static_cast<const float *>(vec), d_out);
      e = cudaGetLastError();
    }
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(out, d_out, 6 * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(m->ctx->stream);
  cudaFree(d_out);
  if (e != cudaSuccess)
    return hd::fail(HD_ERR_CUDA, std::string("hd_phase_space_diagnostics: ") + cudaGetErrorString(e));
  return HD_OK;
}

int
hd_field_energy(hd_mesh *m, const double *a_v_device, double *out)
{
  HD_REQUIRE(m && a_v_device && out, "null argument");
  HD_REQUIRE(m->d.dim_x <= 3, "at most three space directions");
  HD_CUDA(cudaSetDevice(m->ctx->device));
  const DiagParams p = diag_params(m);
  long long        ncx = 1;
  for (int d = 0; d < m->d.dim_x; ++d)
    ncx *= m->d.n_cells[d];
  double *d_out = nullptr;
  HD_CUDA(cudaMalloc(&d_out, 3 * sizeof(double)));
  HD_CUDA(cudaMemsetAsync(d_out, 0, 3 * sizeof(double), m->ctx->stream));
  k_field_energy<<<(unsigned)((ncx + 127) / 128), 128, 0, m->ctx->stream>>>(p, a_v_device, d_out, ncx);
  double      h[3] = {0, 0, 0};
  cudaError_t e    = cudaGetLastError();
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(h, d_out, 3 * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(m->ctx->stream);
  cudaFree(d_out);
  if (e != cudaSuccess)
    return hd::fail(HD_ERR_CUDA, std::string("hd_field_energy: ") + cudaGetErrorString(e));
  for (int d = 0; d < m->d.dim_x; ++d)
    out[d] = h[d];
  return HD_OK;
}

} // extern "C"
#endif
