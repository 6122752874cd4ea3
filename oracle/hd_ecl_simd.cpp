// CPU BASELINE — TEST/BENCH INFRASTRUCTURE ONLY.  Not part of the product path.
//
// Performance-oriented CPU restatement of hyper.deal's element-centric (ECL) advection kernel
//   include/hyper.deal/operators/advection/advection_operation.h:221-566
// for the benchmark configuration (3D3V, degree 3, n_q = 4, FP64, Cartesian periodic lattice, constant velocity):
// the SAME literal algorithm as oracle/hd_oracle.cpp (S sweeps -> cell integrals with D / D^T sweeps -> all 12 faces with
// neighbour face gather, 5 S sweeps, upwind flux, face-normal interpolation -> JxW^-1 -> S^-1 sweeps), organised the way
// the reference organises it for speed:
//   * 8 cells per batch in SoA lanes, consecutive along x_0 — the reference's VectorizedArray<double> over x-cells
//     (matrix_free/vector_access_internal.h:28-98 vectorized_load_and_transpose; lane = x-cell);
//   * 1-D sweeps as unrolled 4x4 kernels on 512-bit vectors (GCC vector extensions; AVX-512 or 2 x AVX2);
//   * JxW hoisted into a table (the reference reads it from the low-dimensional mapping data, fe_evaluation_cell.h:202-209);
//   * one thread per core, pinned, static partition of the batches (the reference: one MPI rank per core).
// It is what bench.py times as `cpu_baseline` / `--impl reference` ("port": the hyper.deal binary itself needs deal.II + MPI,
// which this image does not have).  tests/test_ecl_simd_cpu.py pins it to the scalar oracle at 1e-12.
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <pthread.h>
#include <sched.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace
{
  constexpr int W = 8, ND = 4096, NF = 1024;
  typedef double v8 __attribute__((vector_size(64), aligned(64)));

  inline v8
  splat(double x)
  {
    return v8{x, x, x, x, x, x, x, x};
  }

  struct Op
  {
    int    ncell[6];
    double h[6], a[6], skew;
    double S[16], D[16], Sinv[16], w[4], f0[4], f1[4]; // row-major [q][i]
    std::vector<double> jxw, jxw_inv, jxw_face[6];
  };

  // y = (I x .. x M x .. x I) x along a direction with `stride` entries between consecutive indices; `total` entries.
  // TR: apply M^T.  ADD: accumulate.  In place is fine (a line is loaded before it is written).
  template <bool TR, bool ADD>
  inline void
  sweep4(const double *M, const v8 *x, v8 *y, int stride, int total)
  {
    for (int base = 0; base < total; base += 4 * stride)
      for (int lo = 0; lo < stride; ++lo)
        {
          const v8 *xp = x + base + lo;
          v8 *      yp = y + base + lo;
          const v8  x0 = xp[0], x1 = xp[stride], x2 = xp[2 * stride], x3 = xp[3 * stride];
#pragma GCC unroll 4
          for (int r = 0; r < 4; ++r)
            {
              const double m0 = TR ? M[0 * 4 + r] : M[r * 4 + 0], m1 = TR ? M[1 * 4 + r] : M[r * 4 + 1], m2 = TR ? M[2 * 4 + r] : M[r * 4 + 2],
                           m3 = TR ? M[3 * 4 + r] : M[r * 4 + 3];
              const v8 acc = m0 * x0 + m1 * x1 + m2 * x2 + m3 * x3;
              if (ADD)
                yp[r * stride] += acc;
              else
                yp[r * stride] = acc;
            }
        }
  }

  struct Scratch
  {
    v8 *u, *res, *fm, *fp, *fr;
    Scratch()
    {
      u   = static_cast<v8 *>(aligned_alloc(64, sizeof(v8) * ND));
      res = static_cast<v8 *>(aligned_alloc(64, sizeof(v8) * ND));
      fm  = static_cast<v8 *>(aligned_alloc(64, sizeof(v8) * NF));
      fp  = static_cast<v8 *>(aligned_alloc(64, sizeof(v8) * NF));
      fr  = static_cast<v8 *>(aligned_alloc(64, sizeof(v8) * NF));
    }
    ~Scratch()
    {
      free(u);
      free(res);
      free(fm);
      free(fp);
      free(fr);
    }
  };

  // one batch: the cells (c0 = b0 .. b0+7 clipped to the row, c1..c5 fixed)
  void
  apply_batch(const Op &op, const double *src, double *dst, const int *crow, int b0, Scratch &s)
  {
    const int n0    = op.ncell[0];
    const int valid = (n0 - b0) < W ? (n0 - b0) : W;
    int64_t   rowbase = 0;
    for (int d = 5; d >= 1; --d)
      rowbase = rowbase * op.ncell[d] + crow[d];
    rowbase *= n0;
    int64_t cell[W];
    for (int l = 0; l < W; ++l)
      cell[l] = rowbase + b0 + (l < valid ? l : 0);

    // --- 1) read_dof_values (transpose into lanes) + basis change to the quadrature points (:291-316)
    v8 *u = s.u, *res = s.res;
    for (int i = 0; i < ND; ++i)
      {
        v8 t;
        for (int l = 0; l < W; ++l)
          t[l] = src[cell[l] * ND + i];
        u[i] = t;
      }
    for (int d = 0, st = 1; d < 6; ++d, st *= 4)
      sweep4<false, false>(op.S, u, u, st, ND);

    // --- 2) cell integrals (:325-403): per line of direction d
    //        res_k += -skew c jxw_k (D u)_k + sum_p D[p][k] ((1-skew) c jxw_p u_p),  c = a_d / h_d
    for (int i = 0; i < ND; ++i)
      res[i] = splat(0.0);
    for (int d = 0, st = 1; d < 6; ++d, st *= 4)
      {
        const double c = op.a[d] / op.h[d], cs = -op.skew * c, cf = (1.0 - op.skew) * c;
        if (c == 0.0)
          continue;
        for (int base = 0; base < ND; base += 4 * st)
          for (int lo = 0; lo < st; ++lo)
            {
              const int     i0 = base + lo;
              const v8      u0 = u[i0], u1 = u[i0 + st], u2 = u[i0 + 2 * st], u3 = u[i0 + 3 * st];
              const double *j  = op.jxw.data();
              const double  j0 = j[i0], j1 = j[i0 + st], j2 = j[i0 + 2 * st], j3 = j[i0 + 3 * st];
              const v8      t0 = (cf * j0) * u0, t1 = (cf * j1) * u1, t2 = (cf * j2) * u2, t3 = (cf * j3) * u3;
              const double *D  = op.D;
#pragma GCC unroll 4
              for (int k = 0; k < 4; ++k)
                {
                  const double jk = j[i0 + k * st];
                  v8           r  = D[0 * 4 + k] * t0 + D[1 * 4 + k] * t1 + D[2 * 4 + k] * t2 + D[3 * 4 + k] * t3;
                  if (op.skew != 0.0)
                    r += (cs * jk) * (D[k * 4 + 0] * u0 + D[k * 4 + 1] * u1 + D[k * 4 + 2] * u2 + D[k * 4 + 3] * u3);
                  res[i0 + k * st] += r;
                }
            }
      }

    // --- 3) all 12 faces (:406-526)
    for (int d = 0, st = 1; d < 6; ++d, st *= 4)
      for (int side = 0; side < 2; ++side)
        {
          const double *fvec   = side ? op.f1 : op.f0;
          const double  normal = side ? +1.0 : -1.0;
          v8 *          um = s.fm, *up = s.fp, *fr = s.fr;
          // minus trace: face-normal interpolation of the cell's quadrature values (evaluation_kernels.h:64-106)
          for (int hi = 0, f = 0; hi < ND / (4 * st); ++hi)
            for (int lo = 0; lo < st; ++lo, ++f)
              {
                const v8 *p = u + hi * 4 * st + lo;
                um[f]       = fvec[0] * p[0] + fvec[1] * p[st] + fvec[2] * p[2 * st] + fvec[3] * p[3 * st];
              }
          // plus trace: nodal face layer of the neighbour (read_write_operation.h:186-330), then 5 S sweeps (:432-436)
          int64_t nb[W];
          for (int l = 0; l < W; ++l)
            {
              int cc[6] = {b0 + (l < valid ? l : 0), crow[1], crow[2], crow[3], crow[4], crow[5]};
              cc[d] += side ? 1 : -1;
              if (cc[d] < 0)
                cc[d] += op.ncell[d];
              if (cc[d] >= op.ncell[d])
                cc[d] -= op.ncell[d];
              int64_t idx = 0;
              for (int e = 5; e >= 0; --e)
                idx = idx * op.ncell[e] + cc[e];
              nb[l] = idx;
            }
          const int layer = side ? 0 : 3;
          for (int hi = 0, f = 0; hi < ND / (4 * st); ++hi)
            for (int lo = 0; lo < st; ++lo, ++f)
              {
                const int i = hi * 4 * st + layer * st + lo;
                v8        t;
                for (int l = 0; l < W; ++l)
                  t[l] = src[nb[l] * ND + i];
                up[f] = t;
              }
          for (int e = 0, fs = 1; e < 5; ++e, fs *= 4)
            sweep4<false, false>(op.S, up, up, fs, NF);
          // numerical flux and face submit_value (:455-520, fe_evaluation_face.h:384-400)
          const double  nts = op.a[d] * normal, ants = std::fabs(nts);
          const double *jf  = op.jxw_face[d].data();
          for (int f = 0; f < NF; ++f)
            {
              const v8 m = um[f], p = up[f];
              const v8 flux = 0.5 * ((m + p) * nts + ants * (m - p));
              fr[f]         = -(flux - (op.skew * nts) * m) * jf[f];
            }
          // back to the cell quadrature points (:523)
          for (int hi = 0, f = 0; hi < ND / (4 * st); ++hi)
            for (int lo = 0; lo < st; ++lo, ++f)
              {
                v8 *     r = res + hi * 4 * st + lo;
                const v8 v = fr[f];
                r[0] += fvec[0] * v;
                r[st] += fvec[1] * v;
                r[2 * st] += fvec[2] * v;
                r[3 * st] += fvec[3] * v;
              }
        }

    // --- 4) inverse mass (:529-559): JxW^-1, S^-1 sweeps from the last direction to the first; 5) set_dof_values
    for (int i = 0; i < ND; ++i)
      res[i] *= op.jxw_inv[i];
    for (int d = 5, st = 1024; d >= 0; --d, st /= 4)
      sweep4<false, false>(op.Sinv, res, res, st, ND);
    for (int l = 0; l < valid; ++l)
      {
        double *o = dst + cell[l] * ND;
        for (int i = 0; i < ND; ++i)
          o[i] = res[i][l];
      }
  }
} // namespace

extern "C" {

// dst = M^-1 A(src): 3D3V, degree 3, n_q = 4, periodic Cartesian lattice, constant velocity a[6].
// S, D, Sinv: 4x4 row-major; w, face0, face1: 4 values (same 1-D data as hdo_op of hd_oracle.cpp).
// Returns 0, or -1 for bad arguments.
int
hde_apply(const int *ncell, const double *left, const double *right, const double *a, double skew, const double *S, const double *D, const double *Sinv,
          const double *w, const double *face0, const double *face1, const double *src, double *dst, int nthreads, int pin)
{
  if (!ncell || !src || !dst || src == dst)
    return -1;
  Op op;
  op.skew = skew;
  int64_t nrows = 1;
  for (int d = 0; d < 6; ++d)
    {
      if (ncell[d] < 1)
        return -1;
      op.ncell[d] = ncell[d];
      op.h[d]     = (right[d] - left[d]) / ncell[d];
      op.a[d]     = a[d];
      if (d > 0)
        nrows *= ncell[d];
    }
  std::memcpy(op.S, S, sizeof(op.S));
  std::memcpy(op.D, D, sizeof(op.D));
  std::memcpy(op.Sinv, Sinv, sizeof(op.Sinv));
  std::memcpy(op.w, w, sizeof(op.w));
  std::memcpy(op.f0, face0, sizeof(op.f0));
  std::memcpy(op.f1, face1, sizeof(op.f1));
  op.jxw.resize(ND);
  op.jxw_inv.resize(ND);
  for (int i = 0; i < ND; ++i)
    {
      double j = 1.0;
      for (int e = 0, r = i; e < 6; ++e, r /= 4)
        j *= op.h[e] * op.w[r % 4];
      op.jxw[i]     = j;
      op.jxw_inv[i] = 1.0 / j;
    }
  for (int d = 0; d < 6; ++d)
    {
      op.jxw_face[d].resize(NF);
      for (int f = 0; f < NF; ++f)
        {
          double j = 1.0;
          int    r = f;
          for (int e = 0; e < 6; ++e)
            if (e != d)
              {
                j *= op.h[e] * op.w[r % 4];
                r /= 4;
              }
          op.jxw_face[d][f] = j;
        }
    }
  const int     nb_row  = (ncell[0] + W - 1) / W;
  const int64_t batches = nrows * nb_row;
  if (nthreads < 1)
    nthreads = 1;
  auto worker = [&](int tid) {
    if (pin)
      {
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(tid % (int)std::thread::hardware_concurrency(), &set);
        pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
      }
    Scratch       s;
    const int64_t chunk = (batches + nthreads - 1) / nthreads;
    const int64_t b = tid * chunk, e = (b + chunk < batches) ? b + chunk : batches;
    for (int64_t it = b; it < e; ++it)
      {
        int     crow[6] = {0, 0, 0, 0, 0, 0};
        int64_t r       = it / nb_row;
        for (int d = 1; d < 6; ++d)
          {
            crow[d] = int(r % ncell[d]);
            r /= ncell[d];
          }
        apply_batch(op, src, dst, crow, int(it % nb_row) * W, s);
      }
  };
  if (nthreads == 1)
    worker(0);
  else
    {
      std::vector<std::thread> th;
      for (int t = 0; t < nthreads; ++t)
        th.emplace_back(worker, t);
      for (auto &t : th)
        t.join();
    }
  return 0;
}
}
