// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of hyper.deal's element-centric (ECL) advection operator
//   dst = M^-1 A(src, t)
// following, step by step, the reference's per-cell kernel
//   include/hyper.deal/operators/advection/advection_operation.h:221-566
//   (local_apply_advect_and_inverse_mass_matrix)
// on Cartesian phase-space meshes.  The deal.II pieces the reference calls
// (EvaluatorTensorProduct sweeps, FEEvaluation geometry; deal.II itself is NOT
// in /root/reference, it is an external dependency pinned only as "deal.II
// master 2021-22", CMakeLists.txt:29, .github/workflows/tests.yml:23-26) are
// restated from their published definition:
//   values<d>     : y = (I x .. x S   x .. x I) x      S[q][i]   = l_i(x_q)   (GLL-nodal -> GL points)
//   gradients<d>  : y = (I x .. x D   x .. x I) x      D[q][p]   = l~_p'(x_q) (GL collocation derivative)
//   hessians<d> with inverse_shape_values : y = (I x .. x Sinv x .. x I) x
// The 1-D matrices are NOT computed here: the caller (oracle/oracle.py, numpy)
// passes them in, so this file holds only the loop structure of the operator.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference leg may load this library.  Parity status: PINNED — oracle.py's
// advection driver built on this file reproduces the reference's golden files
// examples/advection/tests/adv_{1D_1D,2D_2D}_k3*.out (see tests/test_oracle_golden.py).
//
// Layout conventions (shared with the product):
//   vector = cells back to back, cell c at offset c * n^dim
//   cell index lexicographic over the dim directions, direction 0 fastest
//     (matrix_free.templates.h:553-562: lid = lid_x + lid_v * n_cells_x)
//   DoF index inside a cell lexicographic, x_0 fastest ... v_last slowest
//     (matrix_free/shape_info.h:126-146)
//   face f = 2*d + side, side 0 = lower, side 1 = upper, x-directions first
//     (fe_evaluation_face.h:196)
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

extern "C" {

struct hdo_mesh
{
  int    dim_x, dim_v;
  int    n_cells[6];  // per direction
  double left[6], right[6];
  int    periodic[6]; // per direction (the reference has one flag per space)
};

struct hdo_op
{
  int           n;   // 1-D DoFs (degree + 1)
  int           nq;  // 1-D quadrature points
  const double *S;     // nq x n   row-major, values of nodal basis at quad points
  const double *D;     // nq x nq  derivative of the collocation basis at quad points
  const double *Sinv;  // n x nq   inverse basis change (n == nq: S^-1)
  const double *w;     // nq       quadrature weights on [0,1]
  const double *xq;    // nq       quadrature points on [0,1]
  const double *face0; // nq       collocation basis evaluated at xi = 0
  const double *face1; // nq       collocation basis evaluated at xi = 1
  double        skew;  // advection_operation_parameters.h:34-41
  int           velocity_kind; // 0: constant a[dim]; 1: separable tables
  const double *a_const;       // dim
  const double *a_x_table;     // [n_cells_v][nq^dim_v][dim_x]   a_x(v-cell, q_v)
  const double *a_v_table;     // [n_cells_x][nq^dim_x][dim_v]   a_v(x-cell, q_x)
  int           bc_kind; // boundary faces: 0 none expected, 1 Dirichlet g (fn_id), 2 homogeneous
  int           fn_id;   // 0: hyperrectangle ExactSolution (examples/advection/cases/hyperrectangle.h:29-66)
  int           eval_level; // AdvectionOperationEvaluationLevel (:37-42): 0 all, 1 cell (faces skipped, :407),
                            // 2 all_without_neighbor_load (:422-423: phi_p not read; taken as zero here)
};

// examples/advection/cases/hyperrectangle.h:46-57
double
hdo_function(int fn_id, int dim, const double *p, double t)
{
  (void)fn_id;
  const double adv[6] = {1.0, 0.15, -0.05, 0.0, 0.0, 0.0};
  const double PI     = 3.14159265358979323846;
  double       r      = std::sin(2.0 * (p[0] - t * adv[0]) * PI);
  for (int d = 1; d < dim; ++d)
    r *= std::cos(2.0 * (p[d] - t * adv[d]) * PI);
  return r;
}
}

namespace
{
  struct Ctx
  {
    int     dim, dim_x, dim_v, n, nq;
    int64_t nd;  // n^dim
    int64_t nqd; // nq^dim
    int64_t nf;  // n^(dim-1)
    int64_t nqf; // nq^(dim-1)
    double  h[6];
    int64_t ncells, ncells_x;
  };

  // y = (I x .. x M x .. x I) x along direction dir of a tensor whose extent is
  // ext_in in every direction < dir ... ; generic: extents given per direction.
  // M is rows x cols row-major; x has extent cols in direction dir, y has rows.
  void
  sweep(const double *M, int rows, int cols, const double *x, double *y, const int *ext_x, int dim, int dir, bool add)
  {
    int64_t stride = 1;
    for (int d = 0; d < dir; ++d)
      stride *= ext_x[d];
    int64_t outer = 1;
    for (int d = dir + 1; d < dim; ++d)
      outer *= ext_x[d];
    for (int64_t o = 0; o < outer; ++o)
      {
        const double *xo = x + o * stride * cols;
        double *      yo = y + o * stride * rows;
        for (int r = 0; r < rows; ++r)
          for (int64_t i = 0; i < stride; ++i)
            {
              double acc = 0.0;
              for (int c = 0; c < cols; ++c)
                acc += M[r * cols + c] * xo[c * stride + i];
              if (add)
                yo[r * stride + i] += acc;
              else
                yo[r * stride + i] = acc;
            }
      }
  }

  // y = (I x .. x M^T x .. x I) x : M is rows x cols; x has extent rows, y has cols
  void
  sweep_T(const double *M, int rows, int cols, const double *x, double *y, const int *ext_x, int dim, int dir, bool add)
  {
    int64_t stride = 1;
    for (int d = 0; d < dir; ++d)
      stride *= ext_x[d];
    int64_t outer = 1;
    for (int d = dir + 1; d < dim; ++d)
      outer *= ext_x[d];
    for (int64_t o = 0; o < outer; ++o)
      {
        const double *xo = x + o * stride * rows;
        double *      yo = y + o * stride * cols;
        for (int c = 0; c < cols; ++c)
          for (int64_t i = 0; i < stride; ++i)
            {
              double acc = 0.0;
              for (int r = 0; r < rows; ++r)
                acc += M[r * cols + c] * xo[r * stride + i];
              if (add)
                yo[c * stride + i] += acc;
              else
                yo[c * stride + i] = acc;
            }
      }
  }

  inline void
  cell_coords(const hdo_mesh &m, int dim, int64_t c, int *cc)
  {
    for (int d = 0; d < dim; ++d)
      {
        cc[d] = int(c % m.n_cells[d]);
        c /= m.n_cells[d];
      }
  }

  inline int64_t
  cell_index(const hdo_mesh &m, int dim, const int *cc)
  {
    int64_t c = 0;
    for (int d = dim - 1; d >= 0; --d)
      c = c * m.n_cells[d] + cc[d];
    return c;
  }

  struct Scratch
  {
    std::vector<double> a, b, buffer, res, tmp, fm, fm2, fp, fp2, fr;
  };

  // velocity component `comp` (0..dim-1) at cell quadrature point (multi-index
  // q[dim]) of cell cc.  ConstantVelocityFieldView
  // (operators/advection/velocity_field_view.h:107-145) or the phase-space
  // separable form a_x(q_v), a_v(cell_x, q_x) of
  // examples/vlasov_poisson/include/velocity_field_view.h:111-160.
  inline double
  velocity(const hdo_mesh &m, const hdo_op &op, const Ctx &c, const int *cc, const int *q, int comp)
  {
    if (op.velocity_kind == 0)
      return op.a_const[comp];
    if (comp < c.dim_x)
      {
        int64_t cv = 0, qv = 0;
        for (int d = c.dim - 1; d >= c.dim_x; --d)
          {
            cv = cv * m.n_cells[d] + cc[d];
            qv = qv * c.nq + q[d];
          }
        int64_t nqv = 1;
        for (int d = 0; d < c.dim_v; ++d)
          nqv *= c.nq;
        return op.a_x_table[(cv * nqv + qv) * c.dim_x + comp];
      }
    else
      {
        int64_t cx = 0, qx = 0;
        for (int d = c.dim_x - 1; d >= 0; --d)
          {
            cx = cx * m.n_cells[d] + cc[d];
            qx = qx * c.nq + q[d];
          }
        int64_t nqx = 1;
        for (int d = 0; d < c.dim_x; ++d)
          nqx *= c.nq;
        return op.a_v_table[(cx * nqx + qx) * c.dim_v + (comp - c.dim_x)];
      }
  }

  void
  apply_cell(const hdo_mesh &m, const hdo_op &op, const Ctx &c, const double *src, double *dst, double time, int64_t cell, Scratch &s)
  {
    const int dim = c.dim, n = c.n, nq = c.nq;
    int       cc[6];
    cell_coords(m, dim, cell, cc);

    int ext[6];

    // --- 1) read_dof_values + basis change GLL -> quadrature points
    //        (advection_operation.h:291-316)
    std::memcpy(s.a.data(), src + cell * c.nd, sizeof(double) * c.nd);
    for (int d = 0; d < dim; ++d)
      ext[d] = n;
    {
      double *      bufs[2] = {s.b.data(), s.a.data()};
      const double *in      = s.a.data();
      int           k       = 0;
      for (int d = 0; d < dim; ++d)
        {
          double *out = (d == dim - 1) ? s.buffer.data() : bufs[k];
          sweep(op.S, nq, n, in, out, ext, dim, d, false);
          ext[d] = nq;
          in     = out;
          k ^= 1;
        } // result = copy of quadrature values in `buffer` (:319-323)
    }
    const double *uq = s.buffer.data();
    for (int d = 0; d < dim; ++d)
      ext[d] = nq;

    // --- 2) cell integrals, x-space then v-space (:325-403); both groups have the
    //        same structure, so one loop over all directions is used here.
    double *res = s.res.data();
    std::fill(res, res + c.nqd, 0.0);
    double *tmp = s.tmp.data();
    for (int d = 0; d < dim; ++d)
      {
        const double inv_h = 1.0 / c.h[d]; // J^-1 = diag(1/h_d)
        if (op.skew != 0.0)
          {
            // gradients<d,true,false>(buffer, tempp) then
            // submit_value(-skew * (J^-T grad u . a) * JxW)   (:333-345, :375-387)
            sweep(op.D, nq, nq, uq, tmp, ext, dim, d, false);
            int q[6] = {0, 0, 0, 0, 0, 0};
            for (int64_t i = 0; i < c.nqd; ++i)
              {
                double jxw = 1.0;
                for (int e = 0; e < dim; ++e)
                  jxw *= c.h[e] * op.w[q[e]];
                res[i] += -op.skew * (inv_h * tmp[i] * velocity(m, op, c, cc, q, d)) * jxw;
                for (int e = 0; e < dim; ++e)
                  {
                    if (++q[e] < nq)
                      break;
                    q[e] = 0;
                  }
              }
          }
        if (op.skew != 1.0)
          {
            // grad_in[d] = (1-skew) * u * a_d ; submit_gradient (J^-1 ., . JxW);
            // gradients<d,false,true> = D^T sweep accumulating     (:347-363, :389-401)
            int q[6] = {0, 0, 0, 0, 0, 0};
            for (int64_t i = 0; i < c.nqd; ++i)
              {
                double jxw = 1.0;
                for (int e = 0; e < dim; ++e)
                  jxw *= c.h[e] * op.w[q[e]];
                tmp[i] = inv_h * ((1.0 - op.skew) * uq[i] * velocity(m, op, c, cc, q, d)) * jxw;
                for (int e = 0; e < dim; ++e)
                  {
                    if (++q[e] < nq)
                      break;
                    q[e] = 0;
                  }
              }
            sweep_T(op.D, nq, nq, tmp, res, ext, dim, d, true);
          }
      }

    // --- 3) faces (:406-526)
    for (int face = 0; face < 2 * dim && op.eval_level != 1; ++face)
      {
        const int     d        = face / 2;
        const int     side     = face % 2;
        const double  normal   = side ? +1.0 : -1.0;
        const double *fvec     = side ? op.face1 : op.face0;
        int64_t       stride_q = 1;
        for (int e = 0; e < d; ++e)
          stride_q *= nq;
        int64_t stride_n = 1;
        for (int e = 0; e < d; ++e)
          stride_n *= n;

        // neighbour / boundary classification (matrix_free.templates.h:252-330 builds
        // this from deal.II; on a Cartesian lattice it is index arithmetic)
        bool is_boundary = false;
        int  nb[6];
        for (int e = 0; e < dim; ++e)
          nb[e] = cc[e];
        nb[d] += side ? 1 : -1;
        if (nb[d] < 0 || nb[d] >= m.n_cells[d])
          {
            if (m.periodic[d])
              nb[d] = (nb[d] + m.n_cells[d]) % m.n_cells[d];
            else
              is_boundary = true;
          }

        double *um = s.fm.data(); // minus trace at face quadrature points
        double *up = s.fp.data(); // plus trace

        // minus side: interpolate_quadrature<true,false>(buffer -> face) (:428),
        // evaluation_kernels.h:64-106 (contract_onto_face)
        {
          int64_t outer = c.nqd / (stride_q * nq);
          for (int64_t o = 0; o < outer; ++o)
            for (int64_t i = 0; i < stride_q; ++i)
              {
                double acc = 0.0;
                for (int k = 0; k < nq; ++k)
                  acc += fvec[k] * uq[o * stride_q * nq + k * stride_q + i];
                um[o * stride_q + i] = acc;
              }
        }

        // face multi-index helper: directions e != d, lexicographic
        int fdirs[6], nfd = 0;
        for (int e = 0; e < dim; ++e)
          if (e != d)
            fdirs[nfd++] = e;

        if (!is_boundary)
          {
            // plus side: phi_p.read_dof_values(src) = nodal values of the neighbour on
            // the shared face (read_write_operation.h:186-330 through
            // face_to_cell_index_nodal), then dim-1 S sweeps (:432-436)
            const int64_t nbc   = cell_index(m, dim, nb);
            const double *unb   = src + nbc * c.nd;
            const int     layer = side ? 0 : n - 1;
            int64_t       outer = c.nd / (stride_n * n);
            double *      fn    = s.fp2.data();
            for (int64_t o = 0; o < outer; ++o)
              for (int64_t i = 0; i < stride_n; ++i)
                fn[o * stride_n + i] = unb[o * stride_n * n + layer * stride_n + i];
            int fext[6];
            for (int e = 0; e < dim - 1; ++e)
              fext[e] = n;
            if (dim == 1)
              up[0] = fn[0];
            double *      fb[2] = {s.fm2.data(), fn};
            const double *in    = fn;
            int           k     = 0;
            for (int e = 0; e < dim - 1; ++e)
              {
                double *out = (e == dim - 2) ? up : fb[k];
                sweep(op.S, nq, n, in, out, fext, dim - 1, e, false);
                fext[e] = nq;
                in      = out;
                k ^= 1;
              }
          }

        // flux (:455-520)
        double *fr = s.fr.data();
        {
          int qf[6] = {0, 0, 0, 0, 0, 0};
          for (int64_t i = 0; i < c.nqf; ++i)
            {
              int    q[6];
              double jxw_face = 1.0;
              for (int e = 0; e < nfd; ++e)
                {
                  q[fdirs[e]] = qf[e];
                  jxw_face *= c.h[fdirs[e]] * op.w[qf[e]];
                }
              q[d] = side ? nq - 1 : 0; // line-constant velocities: any q[d] gives the same a_d
              const double u_minus = um[i];
              double       u_plus;
              if (!is_boundary)
                u_plus = op.eval_level == 2 ? 0.0 : up[i];
              else if (op.bc_kind == 2)
                u_plus = -u_minus; // DirichletHomogenous (:494-495)
              else
                {
                  double p[6];
                  for (int e = 0; e < dim; ++e)
                    p[e] = m.left[e] + c.h[e] * (cc[e] + (e == d ? double(side) : op.xq[q[e]]));
                  u_plus = -u_minus + 2.0 * hdo_function(op.fn_id, dim, p, time); // :496
                }
              const double nts  = velocity(m, op, c, cc, q, d) * normal;
              const double flux = 0.5 * ((u_minus + u_plus) * nts + std::abs(nts) * (u_minus - u_plus));
              fr[i]             = -(flux - op.skew * u_minus * nts) * jxw_face; // face submit_value, fe_evaluation_face.h:384-400
              for (int e = 0; e < nfd; ++e)
                {
                  if (++qf[e] < nq)
                    break;
                  qf[e] = 0;
                }
            }
        }

        // interpolate_quadrature<false,true>: face -> cell quadrature residual (:523)
        {
          int64_t outer = c.nqd / (stride_q * nq);
          for (int64_t o = 0; o < outer; ++o)
            for (int k = 0; k < nq; ++k)
              for (int64_t i = 0; i < stride_q; ++i)
                res[o * stride_q * nq + k * stride_q + i] += fvec[k] * fr[o * stride_q + i];
        }
      }

    // --- 4) inverse mass (:529-559): submit_inv (divide by JxW) and Sinv sweeps,
    //        directions from last to first
    {
      int q[6] = {0, 0, 0, 0, 0, 0};
      for (int64_t i = 0; i < c.nqd; ++i)
        {
          double jxw = 1.0;
          for (int e = 0; e < dim; ++e)
            jxw *= c.h[e] * op.w[q[e]];
          res[i] /= jxw;
          for (int e = 0; e < dim; ++e)
            {
              if (++q[e] < nq)
                break;
              q[e] = 0;
            }
        }
      for (int e = 0; e < dim; ++e)
        ext[e] = nq;
      double *      bufs[2] = {s.a.data(), s.b.data()};
      const double *in      = res;
      int           k       = 0;
      for (int d = dim - 1; d >= 0; --d)
        {
          double *out = bufs[k];
          sweep(op.Sinv, n, nq, in, out, ext, dim, d, false);
          ext[d] = n;
          in     = out;
          k ^= 1;
        }
      // --- 5) set_dof_values (:562): overwrite
      std::memcpy(dst + cell * c.nd, in, sizeof(double) * c.nd);
    }
  }
} // namespace

extern "C" {

// dst[cell_begin..cell_end) = (M^-1 A(src, time))[cells]; other cells untouched.
void
hdo_apply(const hdo_mesh *m, const hdo_op *op, const double *src, double *dst, double time, int nthreads, int64_t cell_begin, int64_t cell_end)
{
  Ctx c;
  c.dim_x = m->dim_x;
  c.dim_v = m->dim_v;
  c.dim   = m->dim_x + m->dim_v;
  c.n     = op->n;
  c.nq    = op->nq;
  c.nd = c.nqd = c.nf = c.nqf = 1;
  c.ncells = c.ncells_x = 1;
  for (int d = 0; d < c.dim; ++d)
    {
      c.nd *= c.n;
      c.nqd *= c.nq;
      c.h[d] = (m->right[d] - m->left[d]) / m->n_cells[d];
      c.ncells *= m->n_cells[d];
      if (d < c.dim_x)
        c.ncells_x *= m->n_cells[d];
    }
  c.nf  = c.nd / c.n;
  c.nqf = c.nqd / c.nq;
  if (cell_end < 0 || cell_end > c.ncells)
    cell_end = c.ncells;
  if (nthreads < 1)
    nthreads = 1;

  auto worker = [&](int tid) {
    Scratch s;
    int64_t big = 1;
    for (int d = 0; d < c.dim; ++d)
      big *= (c.nq > c.n ? c.nq : c.n);
    for (auto *v : {&s.a, &s.b, &s.buffer, &s.res, &s.tmp, &s.fm, &s.fm2, &s.fp, &s.fp2, &s.fr})
      v->resize(big);
    const int64_t ntot  = cell_end - cell_begin;
    const int64_t chunk = (ntot + nthreads - 1) / nthreads;
    const int64_t b     = cell_begin + tid * chunk;
    const int64_t e     = (b + chunk < cell_end) ? b + chunk : cell_end;
    for (int64_t cell = b; cell < e; ++cell)
      apply_cell(*m, *op, c, src, dst, time, cell, s);
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
}

int
hdo_version()
{
  return 1;
}
}
