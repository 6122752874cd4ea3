// General-velocity advection kernel for the Vlasov-Poisson velocity field:
//   a_x = v at the v-space quadrature points, a_v = a table per (x-cell, x-quadrature point), typically grad(phi)
// (examples/vlasov_poisson/include/velocity_field_view.h:111-175 — PhaseSpaceVelocityFieldView).
//
// Two families behind hd_advection_set_phase_space_velocity: the register-tile kernels of kernel_vp_tile.cuh (degree 3 with 4
// quadrature points: 1D1V, 2D2V, 3D3V — the automatic choice there) and the generic kernel of this file (every other degree /
// quadrature, and the cross-check: hd_advection_set_kernel(op, 1)).  Parity of both with the literal oracle at round-off on
// B200: tests/test_vp_gpu.py (1D1V, 2D2V incl. over-integration, 3D3V, FP32); timings in profiles/r02*_vp_*.
//
// Collapsed form: with C = a C_a + |a| C_abs and L_f = a L_a,f + |a| L_abs,f the speed-independent parts of
// the constant-velocity matrices (basis.hpp), direction d contributes
//     M_a (x) [C_a u + sum_f L_a,f trace_f]  +  M_|a| (x) [C_abs u + sum_f L_abs,f trace_f],      M_g = Sinv diag(g(q)) S
// where M_g acts on the transverse node indices the coefficient depends on: the v-direction dim_x + d for an x-direction d
// (g = v-coordinate of the quadrature point), all x-directions for a v-direction (g = the table; applied as S sweeps ->
// pointwise product -> Sinv sweeps).  Both neighbours' traces are needed (the sign of a changes inside a cell).
//
// Correctness-first mapping, the counterpart of the generic kernel: one CTA per cell, everything in shared memory in
// double, block-wide barriers between the sweeps.  Periodic single-GPU lattices only.
//
// The per-cell body is written against (tid, nthreads) and a barrier macro, so that tests/vp_emulation_harness.cpp can
// include THIS file under g++ (-DHD_VP_HOST_EMULATION: one "thread", barriers are no-ops) and check the index logic and the
// coefficients against the oracle without a GPU.  The harness and its entry point live in tests/; the product has no CPU path.
#ifdef HD_VP_HOST_EMULATION
#  include <cmath>
#  include <cstddef>
#  include <stdexcept>
#  include <string>
#  include <vector>
#  define HD_MAX_DIM 6
#  include "basis.hpp"
#  define HD_VP_FN inline
#  define HD_VP_SYNC() ((void)0)
#else
#  include "hd_internal.h"
#  define HD_VP_FN __device__
#  define HD_VP_SYNC() __syncthreads()
#endif

namespace
{
  struct VpParams
  {
    const void *  src;
    void *        dst;
    const double *coef;  // per direction: Ca[n*n], Cabs[n*n], La0[n], La1[n], Labs0[n], Labs1[n]
    const double *basis; // nodes[n], xq[nq], w[nq], S[nq*n], Sinv[n*nq]
    const double *a_v;   // [n_cells_x][nq^dim_x][dim_v]
    int           dim_x, dim_v, n, nq;
    int           ncell[HD_MAX_DIM], cell_offset[HD_MAX_DIM];
    double        left[HD_MAX_DIM], h[HD_MAX_DIM];
    long long     nd, ncells;
    int           cap; // doubles per shared-memory buffer: max(n, nq)^dim
    void *        sol;
    void *        ti_next;
    double        fb, fa;
    int           fused;
  };

  // out[(hi * rows + r) * stride + lo] = sum_c Mat[r * cols + c] * in[(hi * cols + c) * stride + lo]: one sweep along the
  // direction whose current extent is `cols` and whose stride is `stride`; n_outer = product of the slower extents
  HD_VP_FN void
  sweep(const double *in, double *out, const double *Mat, int rows, int cols, long long stride, long long n_outer, int tid, int nthr)
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

  template <typename T>
  HD_VP_FN void
  vp_cell(const VpParams &p, double *sm, const long long cell, const int tid, const int nthr)
  {
    const int     dim = p.dim_x + p.dim_v, n = p.n, nq = p.nq;
    const double *xq = p.basis + n, *S = xq + 2 * nq, *Sinv = S + nq * n;
    double *      u = sm, *out = u + p.cap, *ta = out + p.cap, *tabs = ta + p.cap, *w1 = tabs + p.cap, *w2 = w1 + p.cap, *Mm = w2 + p.cap;
    const long long nd = p.nd;
    const T *       src  = static_cast<const T *>(p.src);
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
    long long cx = 0, nqx = 1;
    {
      long long m = 1;
      for (int d = 0; d < p.dim_x; ++d)
        {
          cx += c[d] * m;
          m *= p.ncell[d];
          nqx *= nq;
        }
    }
    for (long long i = tid; i < nd; i += nthr)
      {
        u[i]   = double(src[cell * nd + i]);
        out[i] = 0.0;
      }
    HD_VP_SYNC();

    const int blk = 2 * n * n + 4 * n;
    for (int d = 0; d < dim; ++d)
      {
        const double *Ca = p.coef + (size_t)d * blk, *Cabs = Ca + n * n, *La0 = Cabs + n * n, *La1 = La0 + n, *Labs0 = La1 + n, *Labs1 = Labs0 + n;
        long long     stride = 1;
        for (int k = 0; k < d; ++k)
          stride *= n;
        // neighbour cells along d (periodic inside the lattice)
        const long long nb_lo = c[d] == 0 ? cell + (long long)(p.ncell[d] - 1) * cstr[d] : cell - cstr[d];
        const long long nb_hi = c[d] == p.ncell[d] - 1 ? cell - (long long)(p.ncell[d] - 1) * cstr[d] : cell + cstr[d];
        // ---- line parts along d: own cell + the end nodes of both neighbours
        for (long long i = tid; i < nd; i += nthr)
          {
            const int       id   = int((i / stride) % n);
            const long long base = i - id * stride;
            double          sa = 0.0, sabs = 0.0;
            for (int j = 0; j < n; ++j)
              {
                const double uj = u[base + j * stride];
                sa += Ca[id * n + j] * uj;
                sabs += Cabs[id * n + j] * uj;
              }
            const double t0 = double(src[nb_lo * nd + base + (long long)(n - 1) * stride]);
            const double t1 = double(src[nb_hi * nd + base]);
            ta[i]   = sa + La0[id] * t0 + La1[id] * t1;
            tabs[i] = sabs + Labs0[id] * t0 + Labs1[id] * t1;
          }
        HD_VP_SYNC();
        if (d < p.dim_x)
          {
            // ---- a_x[d] = v-coordinate of the quadrature point in v-direction e: M = Sinv diag(g) S along e (n x n)
            const int e = p.dim_x + d;
            for (int t = tid; t < n * n; t += nthr)
              {
                const int i = t / n, j = t % n;
                double    ma = 0.0, mabs = 0.0;
                for (int q = 0; q < nq; ++q)
                  {
                    const double g = p.left[e] + p.h[e] * ((c[e] + p.cell_offset[e]) + xq[q]);
                    const double s = Sinv[i * nq + q] * S[q * n + j];
                    ma += s * g;
                    mabs += s * fabs(g);
                  }
                Mm[t]         = ma;
                Mm[n * n + t] = mabs;
              }
            HD_VP_SYNC();
            long long stride_e = 1;
            for (int k = 0; k < e; ++k)
              stride_e *= n;
            for (long long i = tid; i < nd; i += nthr)
              {
                const int       ie   = int((i / stride_e) % n);
                const long long base = i - ie * stride_e;
                double          acc  = 0.0;
                for (int j = 0; j < n; ++j)
                  acc += Mm[ie * n + j] * ta[base + j * stride_e] + Mm[n * n + ie * n + j] * tabs[base + j * stride_e];
                out[i] += acc;
              }
            HD_VP_SYNC();
          }
        else
          {
            // ---- a_v[d - dim_x] = table value at (x-cell, x-quadrature point): S sweeps over the x-directions,
            //      pointwise product, Sinv sweeps — once with g on the a-part, once with |g| on the |a|-part
            const int comp = d - p.dim_x;
            for (int pass = 0; pass < 2; ++pass)
              {
                const double *cur = pass ? tabs : ta;
                double *      nxt = w1;
                long long     lo  = 1; // product of the CURRENT extents of the x-directions below e
                for (int e = 0; e < p.dim_x; ++e)
                  {
                    long long outer = 1; // slower directions: x-directions above e still have n, all v-directions have n
                    for (int k = e + 1; k < dim; ++k)
                      outer *= n;
                    sweep(cur, nxt, S, nq, n, lo, outer, tid, nthr);
                    HD_VP_SYNC();
                    cur = nxt;
                    nxt = (nxt == w1) ? w2 : w1;
                    lo *= nq;
                  }
                long long nv = 1;
                for (int k = 0; k < p.dim_v; ++k)
                  nv *= n;
                double *curw = const_cast<double *>(cur); // w1 or w2
                for (long long i = tid; i < nqx * nv; i += nthr)
                  {
                    const double g = p.a_v[(cx * nqx + i % nqx) * p.dim_v + comp];
                    curw[i] *= pass ? fabs(g) : g;
                  }
                HD_VP_SYNC();
                lo = 1; // x-directions below e are back to n after their Sinv sweep
                for (int e = 0; e < p.dim_x; ++e)
                  {
                    long long outer = 1;
                    for (int k = e + 1; k < p.dim_x; ++k)
                      outer *= nq; // not yet swept back
                    outer *= nv;
                    sweep(cur, nxt, Sinv, n, nq, lo, outer, tid, nthr);
                    HD_VP_SYNC();
                    cur = nxt;
                    nxt = (nxt == w1) ? w2 : w1;
                    lo *= n;
                  }
                for (long long i = tid; i < nd; i += nthr)
                  out[i] += cur[i];
                HD_VP_SYNC();
              }
          }
      }
    // ---- store (or the fused LSRK update)
    T *dst = static_cast<T *>(p.dst), *sol = static_cast<T *>(p.sol), *tin = static_cast<T *>(p.ti_next);
    for (long long i = tid; i < nd; i += nthr)
      {
        const long long g = cell * nd + i;
        if (p.fused)
          {
            const double s = double(sol[g]);
            sol[g]         = T(s + p.fb * out[i]);
            if (p.fa != 0.0)
              tin[g] = T(s + p.fa * out[i]);
          }
        else
          dst[g] = T(out[i]);
      }
  }

  // speed-independent parts of the collapsed matrices of every direction, [dim][Ca n*n | Cabs n*n | La0 | La1 | Labs0 | Labs1]:
  // C(a) = a Ca + |a| Cabs  =>  Ca = (C(1) - C(-1)) / 2, Cabs = (C(1) + C(-1)) / 2, likewise the lifting vectors
  inline void
  vp_coefficients(hd::Basis1D &b, const int dim, const double *h, const double skew, std::vector<double> &out)
  {
    const int n = b.n, blk = 2 * n * n + 4 * n;
    out.assign((size_t)dim * blk, 0.0);
    b.set_skew((hd::LD)skew);
    for (int d = 0; d < dim; ++d)
      {
        std::vector<hd::LD> Cp[4], Cm[4], L0p, L1p, L0m, L1m;
        b.direction_matrices((hd::LD)1, (hd::LD)h[d], (hd::LD)skew, Cp, L0p, L1p);
        b.direction_matrices((hd::LD)-1, (hd::LD)h[d], (hd::LD)skew, Cm, L0m, L1m);
        double *o = out.data() + (size_t)d * blk;
        for (int i = 0; i < n * n; ++i)
          {
            o[i]         = (double)((Cp[0][i] - Cm[0][i]) / 2);
            o[n * n + i] = (double)((Cp[0][i] + Cm[0][i]) / 2);
          }
        for (int i = 0; i < n; ++i)
          {
            o[2 * n * n + i]         = (double)((L0p[i] - L0m[i]) / 2);
            o[2 * n * n + n + i]     = (double)((L1p[i] - L1m[i]) / 2);
            o[2 * n * n + 2 * n + i] = (double)((L0p[i] + L0m[i]) / 2);
            o[2 * n * n + 3 * n + i] = (double)((L1p[i] + L1m[i]) / 2);
          }
      }
  }

#include "kernel_vp_tile.cuh"

  // the tile kernels' coefficient block from the generic one ([dim][Ca | Cabs | La0 | La1 | Labs0 | Labs1]) and the basis
  inline void
  vp_tile_coefficients(const hd::Basis1D &b, const int dim, const std::vector<double> &coef, VpTileCoef &cf)
  {
    const int blk = 2 * 16 + 4 * 4;
    for (int d = 0; d < 4; ++d)
      for (int i = 0; i < 16; ++i)
        {
          cf.Ca[d][i]   = d < dim ? coef[(size_t)d * blk + i] : 0.0;
          cf.Cabs[d][i] = d < dim ? coef[(size_t)d * blk + 16 + i] : 0.0;
          if (i < 4)
            {
              cf.La0[d][i]   = d < dim ? coef[(size_t)d * blk + 32 + i] : 0.0;
              cf.La1[d][i]   = d < dim ? coef[(size_t)d * blk + 36 + i] : 0.0;
              cf.Labs0[d][i] = d < dim ? coef[(size_t)d * blk + 40 + i] : 0.0;
              cf.Labs1[d][i] = d < dim ? coef[(size_t)d * blk + 44 + i] : 0.0;
            }
        }
    for (int i = 0; i < 16; ++i)
      {
        cf.S[i]    = (double)b.S[i];
        cf.Sinv[i] = (double)b.Sinv[i];
      }
    for (int i = 0; i < 4; ++i)
      cf.xq[i] = (double)b.xq[i];
  }

  inline void
  vp_tile6_coefficients(const hd::Basis1D &b, const std::vector<double> &coef, VpTile6Coef &cf)
  {
    const int blk = 2 * 16 + 4 * 4;
    for (int d = 0; d < 6; ++d)
      for (int i = 0; i < 16; ++i)
        {
          cf.Ca[d][i]   = coef[(size_t)d * blk + i];
          cf.Cabs[d][i] = coef[(size_t)d * blk + 16 + i];
          if (i < 4)
            {
              cf.La0[d][i]   = coef[(size_t)d * blk + 32 + i];
              cf.La1[d][i]   = coef[(size_t)d * blk + 36 + i];
              cf.Labs0[d][i] = coef[(size_t)d * blk + 40 + i];
              cf.Labs1[d][i] = coef[(size_t)d * blk + 44 + i];
            }
        }
    for (int i = 0; i < 16; ++i)
      {
        cf.S[i]    = (double)b.S[i];
        cf.Sinv[i] = (double)b.Sinv[i];
      }
    for (int i = 0; i < 4; ++i)
      cf.xq[i] = (double)b.xq[i];
  }

#ifndef HD_VP_HOST_EMULATION
  template <typename T>
  __global__ void __launch_bounds__(256) k_apply_vp(const VpParams p)
  {
    extern __shared__ double sm[];
    vp_cell<T>(p, sm, blockIdx.x, threadIdx.x, blockDim.x);
  }
#endif
} // namespace

#ifndef HD_VP_HOST_EMULATION

namespace
{
  template <int WARPS, int MINB>
  int
  launch_vp_tile_2d2v(hd_mesh *m, const VpParams &p, const VpTileCoef &cf, const bool f64)
  {
    const size_t   smem = (size_t)WARPS * VPT_WARP * sizeof(double);
    const unsigned grid = (unsigned)((m->ncells + 2 * WARPS - 1) / (2 * WARPS));
    if (smem > 48 * 1024) // (per device, so not cached in a static)
      {
        if (f64)
          HD_CUDA(cudaFuncSetAttribute(k_vp_tile_2d2v<double, WARPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else
          HD_CUDA(cudaFuncSetAttribute(k_vp_tile_2d2v<float, WARPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      }
    if (f64)
      k_vp_tile_2d2v<double, WARPS, MINB><<<grid, WARPS * 32, smem, m->ctx->stream>>>(p, cf);
    else
      k_vp_tile_2d2v<float, WARPS, MINB><<<grid, WARPS * 32, smem, m->ctx->stream>>>(p, cf);
    return HD_OK;
  }
} // namespace

namespace hd
{
  bool
  vp_supported(const hd_advection *op, std::string *why)
  {
    const hd_mesh *m = op->mesh;
    if (m->d.dim_x != m->d.dim_v)
      {
        *why = "the phase-space velocity field needs dim_x == dim_v (a_x = v)";
        return false;
      }
    for (int d = 0; d < m->dim; ++d)
      for (int s = 0; s < 2; ++s)
        if (m->d.side_kind[d][s] != HD_SIDE_PERIODIC_LOCAL)
          {
            *why = "the general-velocity kernel covers periodic single-GPU lattices only";
            return false;
          }
    return true;
  }

  int
  vp_upload_coefficients(hd_advection *op)
  {
    hd_mesh *           m = op->mesh;
    std::vector<double> &h = op->h_vp_coef;
    vp_coefficients(m->basis, m->dim, m->h, op->skew, h);
    HD_CUDA(cudaSetDevice(m->ctx->device));
    if (!op->d_vp_coef)
      HD_CUDA(cudaMalloc(&op->d_vp_coef, h.size() * sizeof(double)));
    HD_CUDA(cudaMemcpy(op->d_vp_coef, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    return HD_OK;
  }

  int
  launch_vp(hd_advection *op, void *dst, const void *src, double, const FusedUpdate &fu)
  {
    hd_mesh *m = op->mesh;
    VpParams p;
    p.src   = src;
    p.dst   = dst;
    p.coef  = static_cast<const double *>(op->d_vp_coef);
    p.basis = m->d_basis;
    p.a_v   = static_cast<const double *>(op->d_av);
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
    int mx = m->n > m->nq ? m->n : m->nq;
    long long cap = 1;
    for (int d = 0; d < m->dim; ++d)
      cap *= mx;
    p.cap     = (int)cap;
    p.sol     = fu.sol;
    p.ti_next = fu.ti_next;
    p.fb      = fu.fb;
    p.fa      = fu.fa;
    p.fused   = fu.enabled;
    const bool f64 = m->d.number_type == HD_F64;
    // degree 3 with 4 quadrature points in 1D1V / 2D2V: the register-tile kernels (kernel_vp_tile.cuh); kernel choice 1
    // (hd_advection_set_kernel) keeps the generic one for A/B runs and as the cross-check of the tests
    if (m->n == 4 && m->nq == 4 && m->dim == 6 && op->kernel_choice != 1 && !op->h_vp_coef.empty() && m->ncells < (1ll << 31))
      {
        VpTile6Coef cf;
        vp_tile6_coefficients(m->basis, op->h_vp_coef, cf);
        const size_t smem = (size_t)VPT6_SMEM * sizeof(double);
        if (smem <= m->ctx->smem_optin)
          {
            if (f64)
              {
                HD_CUDA(cudaFuncSetAttribute(k_vp_tile_3d3v<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_vp_tile_3d3v<double><<<(unsigned)m->ncells, 256, smem, m->ctx->stream>>>(p, cf);
              }
            else
              {
                HD_CUDA(cudaFuncSetAttribute(k_vp_tile_3d3v<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_vp_tile_3d3v<float><<<(unsigned)m->ncells, 256, smem, m->ctx->stream>>>(p, cf);
              }
            HD_CUDA(cudaGetLastError());
            op->launches++;
            op->last_kernel = "vp_tile_3d3v";
            return HD_OK;
          }
      }
    if (m->n == 4 && m->nq == 4 && (m->dim == 2 || m->dim == 4) && op->kernel_choice != 1 && !op->h_vp_coef.empty() && m->ncells < (1ll << 31))
      {
        VpTileCoef cf;
        vp_tile_coefficients(m->basis, m->dim, op->h_vp_coef, cf);
        if (m->dim == 2)
          {
            const unsigned grid = (unsigned)((m->ncells + 127) / 128);
            if (f64)
              k_vp_tile_1d1v<double><<<grid, 128, 0, m->ctx->stream>>>(p, cf);
            else
              k_vp_tile_1d1v<float><<<grid, 128, 0, m->ctx->stream>>>(p, cf);
            op->last_kernel = "vp_tile_1d1v";
          }
        else
          {
            // CTA shape (warps, minimum CTAs per SM -> register cap), HD_VP_TILE_VARIANT.  Shared memory (21 KiB per warp)
            // would allow 10 warps per SM, but more than 8 warps (two per SM sub-partition) cap the registers at 168 and the
            // kernel spills; measured on 32^4 cells (profiles/r02_vp_tile_variants.txt): 0 (default) 2 warps x 4 CTAs, 254
            // registers: 94.0 GDoF/s; 1: 2 x 5 (168 + spills): 83.5; 2: 5 x 2: 76.4; 3: 3 x 3: 82.5; 4: 4 x 2: 93.4
            static const int variant = [] {
              const char *e = getenv("HD_VP_TILE_VARIANT");
              return e ? atoi(e) : 0;
            }();
            int rc;
            if (variant == 1)
              rc = launch_vp_tile_2d2v<2, 5>(m, p, cf, f64);
            else if (variant == 2)
              rc = launch_vp_tile_2d2v<5, 2>(m, p, cf, f64);
            else if (variant == 3)
              rc = launch_vp_tile_2d2v<3, 3>(m, p, cf, f64);
            else if (variant == 4)
              rc = launch_vp_tile_2d2v<4, 2>(m, p, cf, f64);
            else
              rc = launch_vp_tile_2d2v<2, 4>(m, p, cf, f64);
            if (rc != HD_OK)
              return rc;
            op->last_kernel = "vp_tile_2d2v";
          }
        HD_CUDA(cudaGetLastError());
        op->launches++;
        return HD_OK;
      }
    const size_t smem = (6 * (size_t)cap + 2 * (size_t)m->n * m->n) * sizeof(double);
    if (smem > m->ctx->smem_optin)
      return hd::fail(HD_ERR_UNSUPPORTED, "general-velocity kernel: the cell does not fit into shared memory six times");
    if (smem > 48 * 1024)
      {
        if (f64)
          HD_CUDA(cudaFuncSetAttribute(k_apply_vp<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else
          HD_CUDA(cudaFuncSetAttribute(k_apply_vp<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      }
    if (f64)
      k_apply_vp<double><<<(unsigned)m->ncells, 256, smem, m->ctx->stream>>>(p);
    else
      k_apply_vp<float><<<(unsigned)m->ncells, 256, smem, m->ctx->stream>>>(p);
    HD_CUDA(cudaGetLastError());
    op->launches++;
    op->last_kernel = "vp_generic";
    return HD_OK;
  }
} // namespace hd
#endif
