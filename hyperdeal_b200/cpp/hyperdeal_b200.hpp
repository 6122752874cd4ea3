// hyperdeal_b200.hpp — header-only C++ host layer over the C ABI of libhdgpu.so
// (include/hyperdeal_b200.h).
//
// It re-creates, for the advection hot path, the class and method names a hyper.deal driver
// uses, so that examples/advection and performance/operators_advection_* read
// the same after switching the include.  Every class cites the reference declaration it mirrors
// (paths relative to the hyper.deal source tree).  Nothing is computed on the host: each method
// forwards to one C-ABI call; failures of the library surface as hyperdeal::ExcMessage
// (the reference's AssertThrow -> exception convention, drivers catch in main and return 1).
//
// What cannot be mirrored: user-supplied cell/face lambdas (MatrixFree::cell_loop / loop with
// FEEvaluation objects) cannot run on the device; those members exist and throw
// ExcNotImplemented, like the reference does for unsupported options
// (matrix_free/matrix_free.templates.h:870-875).
#ifndef HYPERDEAL_B200_HPP
#define HYPERDEAL_B200_HPP

#include <hyperdeal_b200.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace hyperdeal
{
  // ---- errors -------------------------------------------------------------------------------
  struct ExcMessage : std::runtime_error
  {
    explicit ExcMessage(const std::string &m)
      : std::runtime_error(m)
    {}
  };
  struct ExcNotImplemented : ExcMessage
  {
    explicit ExcNotImplemented(const std::string &what = "")
      : ExcMessage("ExcNotImplemented: " + what)
    {}
  };

  namespace internal
  {
    inline void
    check(const int rc, const char *call)
    {
      if (rc < 0)
        throw ExcMessage(std::string(call) + ": " + hd_last_error() + " (code " + std::to_string(rc) + ")");
    }
    template <typename Number>
    constexpr int
    number_type()
    {
      static_assert(std::is_same<Number, double>::value || std::is_same<Number, float>::value, "Number must be double or float");
      return std::is_same<Number, double>::value ? HD_F64 : HD_F32;
    }
  } // namespace internal
#define HD_CALL(expr) ::hyperdeal::internal::check((expr), #expr)

  // ---- the few deal.II vocabulary types the signatures mention --------------------------------
  namespace dealii_compat
  {
    namespace types
    {
      using boundary_id = unsigned int;
    }
    namespace numbers
    {
      constexpr types::boundary_id internal_face_boundary_id = static_cast<types::boundary_id>(-1);
    }

    template <int rank, int dim, typename Number = double>
    class Tensor;
    template <int dim, typename Number>
    class Tensor<1, dim, Number>
    {
    public:
      Tensor() { v.fill(Number(0)); }
      Number &operator[](const unsigned int i) { return v[i]; }
      const Number &operator[](const unsigned int i) const { return v[i]; }

    protected:
      std::array<Number, dim> v;
    };

    template <int dim, typename Number = double>
    class Point : public Tensor<1, dim, Number>
    {
    public:
      Point() = default;
      explicit Point(const Number x) { this->v[0] = x; }
      Point(const Number x, const Number y)
      {
        this->v[0] = x;
        this->v[1] = y;
      }
      Point(const Number x, const Number y, const Number z)
      {
        this->v[0] = x;
        this->v[1] = y;
        this->v[2] = z;
      }
    };

    // dealii::Function: value(p) + the time the drivers set before evaluating it.
    template <int dim, typename Number = double>
    class Function
    {
    public:
      virtual ~Function() = default;
      virtual Number value(const Point<dim, Number> &, const unsigned int = 0) const { return Number(0); }
      void   set_time(const Number t) { time = t; }
      Number get_time() const { return time; }
      // id of the device-side implementation of this field (HD_FN_*), or -1 if it only exists on the host
      virtual int device_function_id() const { return -1; }

    private:
      Number time = Number(0);
    };
    namespace Functions
    {
      template <int dim, typename Number = double>
      class ZeroFunction : public Function<dim, Number>
      {
      public:
        int device_function_id() const override { return HD_FN_ZERO; }
      };
    } // namespace Functions
  }   // namespace dealii_compat

  // ---- communicator stand-in: one process = one GPU ---------------------------------------------
  // Takes the place of the (comm, comm_sm) pair of hyperdeal::MatrixFree's constructor (matrix_free/matrix_free.h:108).
  class DeviceCommunicator
  {
  public:
    explicit DeviceCommunicator(const int device = 0)
    {
      HD_CALL(hd_context_create(device, &ctx));
    }
    ~DeviceCommunicator()
    {
      if (ctx)
        hd_context_destroy(ctx);
    }
    DeviceCommunicator(const DeviceCommunicator &) = delete;
    DeviceCommunicator &operator=(const DeviceCommunicator &) = delete;
    hd_context *context() const { return ctx; }
    void        set_stream(void *cuda_stream) const { HD_CALL(hd_context_set_stream(ctx, cuda_stream)); }
    void        synchronize() const { HD_CALL(hd_context_synchronize(ctx)); }

  private:
    hd_context *ctx = nullptr;
  };

  // ---- internal::MatrixFreeFunctions::ShapeInfo (matrix_free/shape_info.h:26-231) ------------------------------------
  // Host tables of the phase-space element: which nodal values of a cell lie on each of its 2 * dim faces, in the order
  // the reference's face evaluators enumerate them.  The device kernels do not read these tables (on a Cartesian lattice a
  // face layer is "the cell index with one digit fixed"); they are here for callers of get_shape_info().
  // Ordering, restated from its definition: inside a dim_s-dimensional subspace the face of direction e lists the other
  // digits in natural order, except for the middle direction of a 3-D subspace, whose local coordinate system is (z, x):
  // the index along z runs fastest.  An x-face then takes all v-digits as the slow part, a v-face all x-digits as the fast part.
  namespace internal
  {
    namespace MatrixFreeFunctions
    {
      template <typename Number>
      struct ShapeInfo
      {
        template <int dim_x, int dim_v>
        void
        reinit(const unsigned int degree)
        {
          const unsigned int n = degree + 1, dim = dim_x + dim_v;
          const auto         ipow = [](unsigned int b, unsigned int e) {
            unsigned int r = 1;
            while (e--)
              r *= b;
            return r;
          };
          dofs_per_cell = ipow(n, dim);
          dofs_per_face = ipow(n, dim - 1);
          // subspace table: cell index (inside the subspace) of face entry l
          const auto sub = [&](const int dim_s, const unsigned int f, const unsigned int l) {
            const unsigned int e = f / 2, layer = (f % 2) * (n - 1);
            unsigned int       digit[3] = {0, 0, 0};
            if (dim_s == 3 && e == 1)
              {
                digit[2] = l % n; // z fastest
                digit[0] = l / n;
              }
            else
              {
                unsigned int r = l;
                for (int k = 0; k < dim_s; ++k)
                  if (k != int(e))
                    {
                      digit[k] = r % n;
                      r /= n;
                    }
              }
            digit[e] = layer;
            unsigned int idx = 0;
            for (int k = dim_s - 1; k >= 0; --k)
              idx = idx * n + digit[k];
            return idx;
          };
          const unsigned int nx = ipow(n, dim_x), nfx = ipow(n, dim_x - 1), nfv = ipow(n, dim_v - 1);
          face_to_cell_index_nodal.assign(2 * dim, std::vector<unsigned int>(dofs_per_face));
          for (unsigned int f = 0; f < 2 * dim; ++f)
            for (unsigned int k = 0; k < dofs_per_face; ++k)
              {
                if (f < 2 * (unsigned int)dim_x)
                  face_to_cell_index_nodal[f][k] = sub(dim_x, f, k % nfx) + nx * (k / nfx);
                else
                  face_to_cell_index_nodal[f][k] = (k % nx) + nx * sub(dim_v, f - 2 * dim_x, (k / nx) % nfv);
              }
          // the 2 x 8 orientations of a quadrilateral face of a 3-D subspace (x-faces: tables 0-7, v-faces: 8-15).  Entry c
          // of a table = where face value c goes; (orientation, flip, rotation) swap and mirror the two in-face indices.
          // A Cartesian lattice only ever uses the identity (tables 0 and 8).
          if (dim_x == 3 || dim_v == 3)
            {
              face_orientations.assign(16, std::vector<unsigned int>(dofs_per_face));
              // {first index: 0 = take k, 1 = take j; mirrored?}, {second index ...} for tables 0..7
              static const int pick[8][4] = {{0, 0, 1, 0}, {1, 0, 0, 0}, {0, 1, 1, 1}, {1, 1, 0, 1}, {1, 0, 0, 1}, {0, 0, 1, 1}, {1, 1, 0, 0}, {0, 1, 1, 0}};
              const auto     entry = [&](const int t, const unsigned int j, const unsigned int k) {
                const unsigned int a = pick[t][0] ? j : k, b = pick[t][2] ? j : k;
                return (pick[t][1] ? n - 1 - a : a) + (pick[t][3] ? n - 1 - b : b) * n;
              };
              for (int t = 0; t < 8; ++t)
                for (unsigned int c = 0; c < dofs_per_face; ++c)
                  {
                    // x-face: c = i * n^2 + j * n + k (i over the v-digits); v-face: c = (j * n + k) * n^dim_x + i
                    if (dim_x == 3)
                      face_orientations[t][c] = entry(t, (c / n) % n, c % n) + (c / (n * n)) * n * n;
                    else
                      face_orientations[t][c] = c;
                    if (dim_v == 3)
                      face_orientations[8 + t][c] = entry(t, (c / nx) / n, (c / nx) % n) * nx + c % nx;
                    else
                      face_orientations[8 + t][c] = c;
                  }
            }
        }
        std::size_t
        memory_consumption() const
        {
          std::size_t b = 0;
          for (const auto &v : face_to_cell_index_nodal)
            b += v.size() * sizeof(unsigned int);
          return b;
        }
        unsigned int                           dofs_per_cell = 0, dofs_per_face = 0;
        std::vector<std::vector<unsigned int>> face_to_cell_index_nodal;
        std::vector<std::vector<unsigned int>> face_orientations;
      };
    } // namespace MatrixFreeFunctions

    // boundary id of face `face` (2 * direction + side) of local cell `cell`, or internal_face_boundary_id
    // (MatrixFree::get_faces_by_cells_boundary_id, matrix_free.h:279): only the outer layer of a non-periodic direction
    // has boundary faces; 1-D subspaces number their two end points 0 and 1 (see get_boundary_id)
    inline dealii_compat::types::boundary_id
    face_boundary_id(const hd_mesh_desc &d, std::int64_t cell, const unsigned int face)
    {
      const int dim = d.dim_x + d.dim_v, dir = int(face / 2), side = int(face % 2);
      int       c   = 0;
      for (int e = 0; e < dim; ++e)
        {
          if (e == dir)
            c = int(cell % d.n_cells[e]);
          cell /= d.n_cells[e];
        }
      const bool outer = side ? (c + d.cell_offset[dir] == d.n_cells_global[dir] - 1) : (c + d.cell_offset[dir] == 0);
      const int  kind  = d.side_kind[dir][side];
      if (!outer || (kind != HD_SIDE_DIRICHLET && kind != HD_SIDE_DIRICHLET_HOM))
        return dealii_compat::numbers::internal_face_boundary_id;
      const bool one_d = dir < d.dim_x ? d.dim_x == 1 : d.dim_v == 1;
      return one_d ? side : 0;
    }
  } // namespace internal

  // ---- low-dimensional lattice: what the two dealii::MatrixFree arguments describe on this path ----
  // (subdivided_hyper_rectangle, grid/grid_generator.cc:235, with FE_DGQ(degree) and QGauss(n_points) or the
  // collocation quadrature, tests/tests_mf.h:162-202)
  template <int dim>
  struct CartesianLattice
  {
    std::array<double, dim>       left{}, right{};
    std::array<unsigned int, dim> n_cells{};
    bool                          periodic    = true;
    unsigned int                  degree      = 3;
    unsigned int                  n_points    = 4;
    bool                          collocation = false;
  };

  template <typename Number>
  class DeviceVector;

  // ---- hyperdeal::MatrixFree (matrix_free/matrix_free.h:39) ------------------------------------------
  template <int dim_x, int dim_v, typename Number = double>
  class MatrixFree
  {
  public:
    static const int dim = dim_x + dim_v;

    struct AdditionalData // matrix_free.h:63-102
    {
      AdditionalData()
        : do_ghost_faces(true)
        , do_buffering(false)
        , use_ecl(true)
        , overlapping_level(0)
      {}
      bool         do_ghost_faces;
      bool         do_buffering;
      bool         use_ecl;
      unsigned int overlapping_level;
    };

    MatrixFree(const DeviceCommunicator &comm, const CartesianLattice<dim_x> &matrix_free_x, const CartesianLattice<dim_v> &matrix_free_v)
      : comm(comm)
      , lattice_x(matrix_free_x)
      , lattice_v(matrix_free_v)
    {}
    ~MatrixFree()
    {
      if (mesh)
        hd_mesh_destroy(mesh);
    }
    MatrixFree(const MatrixFree &) = delete;
    MatrixFree &operator=(const MatrixFree &) = delete;

    // matrix_free.templates.h:862: builds the device-side lattice description.  FCL (use_ecl = false) and buffering give
    // the same results as ECL and are served by the same kernels; ghost cells are refused as in the reference (:870).
    void
    reinit(const AdditionalData &ad = AdditionalData())
    {
      if (!ad.do_ghost_faces)
        throw ExcNotImplemented("ghost cells (do_ghost_faces = false)");
      if (lattice_x.degree != lattice_v.degree || lattice_x.n_points != lattice_v.n_points || lattice_x.collocation != lattice_v.collocation)
        throw ExcNotImplemented("different degree / quadrature in x and v");
      additional_data = ad;
      hd_mesh_desc d{};
      d.dim_x       = dim_x;
      d.dim_v       = dim_v;
      d.degree      = lattice_x.degree;
      d.n_points    = lattice_x.n_points;
      d.collocation = lattice_x.collocation;
      d.number_type = internal::number_type<Number>();
      for (int i = 0; i < HD_MAX_DIM; ++i)
        {
          d.left[i]           = 0.0;
          d.right[i]          = 1.0;
          d.n_cells_global[i] = d.n_cells[i] = 1;
          d.cell_offset[i]                   = 0;
          d.side_kind[i][0] = d.side_kind[i][1] = HD_SIDE_PERIODIC_LOCAL;
        }
      for (int i = 0; i < dim; ++i)
        {
          const bool in_x     = i < dim_x;
          const int  j        = in_x ? i : i - dim_x;
          d.left[i]           = in_x ? lattice_x.left[j] : lattice_v.left[j];
          d.right[i]          = in_x ? lattice_x.right[j] : lattice_v.right[j];
          d.n_cells_global[i] = d.n_cells[i] = in_x ? lattice_x.n_cells[j] : lattice_v.n_cells[j];
          const bool periodic                = in_x ? lattice_x.periodic : lattice_v.periodic;
          d.side_kind[i][0] = d.side_kind[i][1] = periodic ? HD_SIDE_PERIODIC_LOCAL : HD_SIDE_DIRICHLET;
        }
      if (mesh)
        hd_mesh_destroy(mesh);
      mesh = nullptr;
      HD_CALL(hd_mesh_create(comm.context(), &d, &mesh));
      desc = d;
    }

    // matrix_free.templates.h:1369
    void
    initialize_dof_vector(DeviceVector<Number> &vec, const unsigned int dof_handler_index = 0, const bool do_ghosts = true, const bool zero_out = true) const
    {
      if (dof_handler_index != 0)
        throw ExcNotImplemented("dof_handler_index != 0");
      vec.reinit(mesh, do_ghosts);
      (void)zero_out; // vectors are always zero-initialised (hd_vector_alloc)
    }

    // x-space vector (dealii::MatrixFree<dim_x>::initialize_dof_vector of the Vlasov-Poisson driver)
    void
    initialize_dof_vector_x(DeviceVector<Number> &vec) const
    {
      vec.reinit_x(mesh);
    }

    // user-supplied host lambdas cannot run on the device (see the header comment)
    template <typename OutVector, typename InVector, typename Fn>
    void
    cell_loop(const Fn &, OutVector &, const InVector &) const
    {
      throw ExcNotImplemented("MatrixFree::cell_loop with a host lambda; use advection::AdvectionOperation / VectorTools");
    }
    template <typename OutVector, typename InVector, typename Fn>
    void
    loop_cell_centric(const Fn &, OutVector &, const InVector &) const
    {
      throw ExcNotImplemented("MatrixFree::loop_cell_centric with a host lambda");
    }

    const DeviceCommunicator &get_communicator() const { return comm; }
    bool                      is_ecl_supported() const { return true; }
    bool                      are_ghost_faces_supported() const { return true; }
    const CartesianLattice<dim_x> &get_matrix_free_x() const { return lattice_x; }
    const CartesianLattice<dim_v> &get_matrix_free_v() const { return lattice_v; }
    const AdditionalData &         get_additional_data() const { return additional_data; }
    hd_mesh *                      get_mesh() const { return mesh; }
    const hd_mesh_desc &           get_mesh_desc() const { return desc; }
    std::int64_t                   n_dofs() const { return hd_mesh_n_dofs(mesh); }
    std::int64_t                   n_cells() const { return hd_mesh_n_cells(mesh); }
    // boundary id of side (direction, side) as subdivided_hyper_rectangle leaves it (grid/grid_generator.cc:235, no
    // colorize): 0 everywhere, except that a 1-D triangulation numbers its two end points 0 and 1 — the reason for the
    // "hack for 1D" in examples/advection/cases/hyperrectangle.h:190-199
    dealii_compat::types::boundary_id
    get_boundary_id(const unsigned int direction, const unsigned int side) const
    {
      const bool one_d = direction < (unsigned int)dim_x ? dim_x == 1 : dim_v == 1;
      return one_d ? side : 0;
    }
    // matrix_free.h:279: boundary id of a face of a local cell, numbers::internal_face_boundary_id for interior faces
    dealii_compat::types::boundary_id
    get_faces_by_cells_boundary_id(const std::int64_t cell, const unsigned int face) const
    {
      return internal::face_boundary_id(desc, cell, face);
    }
    // matrix_free.h:308: the face tables of the phase-space element (host side)
    const internal::MatrixFreeFunctions::ShapeInfo<Number> &
    get_shape_info() const
    {
      if (shape_info.dofs_per_cell == 0)
        shape_info.template reinit<dim_x, dim_v>(lattice_x.degree);
      return shape_info;
    }
    std::size_t
    memory_consumption() const
    {
      return sizeof(*this) + shape_info.memory_consumption();
    }

  private:
    mutable internal::MatrixFreeFunctions::ShapeInfo<Number> shape_info;
    const DeviceCommunicator &comm;
    CartesianLattice<dim_x>   lattice_x;
    CartesianLattice<dim_v>   lattice_v;
    AdditionalData            additional_data;
    hd_mesh *                 mesh = nullptr;
    hd_mesh_desc              desc{};
  };

  // ---- VectorType: dealii::LinearAlgebra::distributed::Vector<Number> on the device ------------------------
  // Members the drivers use: begin(), operator=(0), zero_out_ghost_values, has_ghost_elements, size,
  // memory_consumption; plus explicit host transfers.
  template <typename Number>
  class DeviceVector
  {
  public:
    using value_type = Number;
    DeviceVector()   = default;
    ~DeviceVector() { clear(); }
    DeviceVector(const DeviceVector &) = delete;
    DeviceVector &operator=(const DeviceVector &) = delete;

    void
    reinit(hd_mesh *m, const bool do_ghosts)
    {
      clear();
      mesh    = m;
      ghosted = do_ghosts;
      HD_CALL(hd_vector_alloc(mesh, do_ghosts ? 1 : 0, &ptr));
      n = hd_mesh_n_dofs(mesh);
    }
    // x-space vector (dealii::MatrixFree<dim_x>::initialize_dof_vector in examples/vlasov_poisson/include/application.h)
    void
    reinit_x(hd_mesh *m)
    {
      clear();
      mesh    = m;
      ghosted = false;
      HD_CALL(hd_vector_alloc_x(mesh, &ptr));
      n = hd_mesh_n_dofs_x(mesh);
    }
    void
    clear()
    {
      if (ptr)
        hd_vector_free(mesh, ptr);
      ptr = nullptr;
    }
    Number *      begin() { return static_cast<Number *>(ptr); }
    const Number *begin() const { return static_cast<const Number *>(ptr); }
    std::int64_t  size() const { return n; }
    std::int64_t  locally_owned_size() const { return n; }
    bool          has_ghost_elements() const { return ghosted; }
    void          zero_out_ghost_values() const {}
    std::size_t   memory_consumption() const { return std::size_t(n) * sizeof(Number); }
    hd_mesh *     get_mesh() const { return mesh; }
    // element access from the host (one device -> host copy per call: for inspection, not for loops)
    Number
    operator()(const std::int64_t i) const
    {
      if (i < 0 || i >= n)
        throw ExcMessage("DeviceVector: index out of range");
      Number v;
      HD_CALL(hd_vector_copy_out(mesh, static_cast<const Number *>(ptr) + i, &v, 1));
      return v;
    }
    // the vectors of all ranks of the shared-memory domain (LinearAlgebra::SharedMPI::Vector): one process = one GPU here
    std::vector<Number *> shared_vector_data() { return std::vector<Number *>(1, begin()); }
    DeviceVector &
    operator=(const Number s)
    {
      if (s != Number(0))
        throw ExcNotImplemented("vector = s with s != 0");
      HD_CALL(hd_vector_zero_n(mesh, ptr, n)); // n = the size this vector was allocated with (phase space or x-space)
      return *this;
    }
    void
    copy_locally_owned_data_from(const DeviceVector &src)
    {
      if (src.n != n)
        throw ExcMessage("copy_locally_owned_data_from: vectors of different size");
      HD_CALL(hd_vector_copy_n(mesh, ptr, src.ptr, n));
    }
    void
    copy_from_host(const std::vector<Number> &h)
    {
      if (std::int64_t(h.size()) > n)
        throw ExcMessage("copy_from_host: the host vector is larger than the device vector");
      HD_CALL(hd_vector_copy_in(mesh, ptr, h.data(), std::int64_t(h.size())));
    }
    void
    copy_to_host(std::vector<Number> &h) const
    {
      h.resize(n);
      HD_CALL(hd_vector_copy_out(mesh, ptr, h.data(), n));
    }

  private:
    hd_mesh *    mesh    = nullptr;
    void *       ptr     = nullptr;
    std::int64_t n       = 0;
    bool         ghosted = false;
  };

  // ---- timers (base/timers.h:36): device-event timing of a section on the context's stream -------------------
  class Timers
  {
  public:
    explicit Timers(const bool = false) {}
    struct Timer
    {
      double       accumulated_us = 0;
      unsigned int counter        = 0;
      double       get_accumulated_time() const { return accumulated_us; }
      unsigned int get_counter() const { return counter; }
    };
    Timer &operator[](const std::string &label) { return timers[label]; }
    void   reset() { timers.clear(); }

  private:
    std::map<std::string, Timer> timers;
  };
  class DynamicConvergenceTable // base/dynamic_convergence_table.h: label -> value rows
  {
  public:
    void set(const std::string &label, const double v) { values[label] = v; }
    void
    print() const
    {
      for (const auto &kv : values)
        std::printf("%-40s %.10g\n", kv.first.c_str(), kv.second);
    }
    std::map<std::string, double> values;
  };

  namespace advection
  {
    // operators/advection/advection_operation_parameters.h:29-41 (sic: "Paramters")
    struct AdvectionOperationParamters
    {
      double factor_skew = 0.0;
    };

    enum class BoundaryType // boundary_descriptor.h:31-37
    {
      Undefined,
      DirichletInhomogenous,
      DirichletHomogenous,
    };

    // boundary_descriptor.h:42-107
    template <int dim, typename Number>
    struct BoundaryDescriptor
    {
      std::map<dealii_compat::types::boundary_id, std::shared_ptr<dealii_compat::Function<dim, Number>>> dirichlet_bc;
      std::set<dealii_compat::types::boundary_id>                                                        homogeneous_dirichlet_bc;

      std::pair<BoundaryType, std::shared_ptr<dealii_compat::Function<dim, Number>>>
      get_boundary(const dealii_compat::types::boundary_id &boundary_id) const
      {
        const auto it = dirichlet_bc.find(boundary_id);
        if (it != dirichlet_bc.end())
          return {BoundaryType::DirichletInhomogenous, it->second};
        if (homogeneous_dirichlet_bc.count(boundary_id))
          return {BoundaryType::DirichletHomogenous, std::make_shared<dealii_compat::Functions::ZeroFunction<dim, Number>>()};
        throw ExcMessage("Boundary type of face is invalid or not implemented.");
      }
      void
      set_time(const Number time)
      {
        for (auto &bc : dirichlet_bc)
          bc.second->set_time(time);
      }
    };

    // operators/advection/velocity_field_view.h:69-166
    template <int dim, typename Number>
    class ConstantVelocityFieldView
    {
    public:
      explicit ConstantVelocityFieldView(const dealii_compat::Tensor<1, dim, Number> &transport_direction)
        : transport_direction(transport_direction)
      {}
      const dealii_compat::Tensor<1, dim, Number> &get_transport_direction() const { return transport_direction; }
      const double *                               phase_space_table() const { return nullptr; }

    private:
      dealii_compat::Tensor<1, dim, Number> transport_direction;
    };

    enum class AdvectionOperationEvaluationLevel // advection_operation.h:37-42 (hd_advection_set_evaluation_level)
    {
      cell,
      all_without_neighbor_load,
      all
    };

    // operators/advection/advection_operation.h:56
    template <int dim_x, int dim_v, int degree, int n_points, typename Number, typename VectorType, typename VelocityField>
    class AdvectionOperation
    {
    public:
      static const int dim = dim_x + dim_v;

      AdvectionOperation(const MatrixFree<dim_x, dim_v, Number> &data, DynamicConvergenceTable &table)
        : data(data)
        , table(table)
      {}
      ~AdvectionOperation()
      {
        if (op)
          hd_advection_destroy(op);
      }
      AdvectionOperation(const AdvectionOperation &) = delete;
      AdvectionOperation &operator=(const AdvectionOperation &) = delete;

      // advection_operation.h:98
      void
      reinit(std::shared_ptr<BoundaryDescriptor<dim, Number>> boundary_descriptor, std::shared_ptr<VelocityField> velocity_field, const AdvectionOperationParamters additional_data)
      {
        const auto &d = data.get_mesh_desc();
        if (d.degree != degree || d.n_points != n_points)
          throw ExcMessage("Degrees " + std::to_string(d.degree) + " and " + std::to_string(degree) + " do not match!");
        this->boundary_descriptor = boundary_descriptor;
        this->velocity_field      = velocity_field;
        double a[HD_MAX_DIM]      = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < dim; ++i)
          a[i] = velocity_field->get_transport_direction()[i];
        if (op)
          hd_advection_destroy(op);
        op = nullptr;
        HD_CALL(hd_advection_create(data.get_mesh(), additional_data.factor_skew, a, &op));
        // a PhaseSpaceVelocityFieldView hands over its gradient table: a_x = v, a_v = table (velocity_field_view.h:111-175)
        if (velocity_field->phase_space_table() != nullptr)
          HD_CALL(hd_advection_set_phase_space_velocity(op, velocity_field->phase_space_table()));
        // Dirichlet sides: u+ = -u- + 2 g; g comes from the descriptor (advection_operation.h:484-520)
        host_sampled_bc = false;
        for (int dir = 0; dir < dim; ++dir)
          for (int side = 0; side < 2; ++side)
            if (d.side_kind[dir][side] == HD_SIDE_DIRICHLET)
              {
                const auto bc = boundary_descriptor->get_boundary(data.get_boundary_id(dir, side));
                const int  id = bc.second->device_function_id();
                if (id >= 0)
                  HD_CALL(hd_advection_set_dirichlet_builtin(op, id));
                else
                  host_sampled_bc = true;
              }
      }

      // advection_operation.h:137: dst = M^-1 A(src, time)
      template <AdvectionOperationEvaluationLevel eval_level = AdvectionOperationEvaluationLevel::all>
      void
      apply(VectorType &dst, const VectorType &src, const Number time, Timers * = nullptr)
      {
        HD_CALL(hd_advection_set_evaluation_level(op, eval_level == AdvectionOperationEvaluationLevel::all  ? HD_EVAL_ALL :
                                                      eval_level == AdvectionOperationEvaluationLevel::cell ? HD_EVAL_CELL :
                                                                                                               HD_EVAL_ALL_WITHOUT_NEIGHBOR_LOAD));
        if (host_sampled_bc)
          upload_dirichlet_values(time);
        HD_CALL(hd_advection_apply(op, dst.begin(), src.begin(), nullptr, double(time)));
      }

      hd_advection *handle() const { return op; }
      bool          needs_host_boundary_data() const { return host_sampled_bc; }
      const char *  kernel_name() const { return hd_advection_kernel_name(op); }
      const MatrixFree<dim_x, dim_v, Number> &get_matrix_free() const { return data; }

    private:
      // boundary data that only exists as a host dealii::Function: sample g at the face quadrature points of every
      // boundary face for the stage time (MatrixFreeTools::evaluate_scalar_function, matrix_free/tools.h:31) and upload
      void
      upload_dirichlet_values(const Number time)
      {
        const auto &        d  = data.get_mesh_desc();
        const int           nq = d.n_points;
        std::vector<double> xq(nq);
        hd_mesh_basis(data.get_mesh(), 1, xq.data());
        boundary_descriptor->set_time(time);
        for (int dir = 0; dir < dim; ++dir)
          for (int side = 0; side < 2; ++side)
            {
              if (d.side_kind[dir][side] != HD_SIDE_DIRICHLET)
                continue;
              const auto bc = boundary_descriptor->get_boundary(data.get_boundary_id(dir, side));
              if (bc.second->device_function_id() >= 0)
                continue;
              std::int64_t n_face_cells = 1, n_face_q = 1;
              for (int e = 0; e < dim; ++e)
                if (e != dir)
                  {
                    n_face_cells *= d.n_cells[e];
                    n_face_q *= nq;
                  }
              std::vector<double>                g(n_face_cells * n_face_q);
              dealii_compat::Point<dim, Number> p;
              for (std::int64_t fc = 0; fc < n_face_cells; ++fc)
                for (std::int64_t q = 0; q < n_face_q; ++q)
                  {
                    std::int64_t c = fc, qq = q;
                    for (int e = 0; e < dim; ++e)
                      {
                        const double h = (d.right[e] - d.left[e]) / d.n_cells_global[e];
                        if (e == dir)
                          {
                            p[e] = side == 0 ? d.left[e] : d.right[e];
                            continue;
                          }
                        const int ce = c % d.n_cells[e];
                        c /= d.n_cells[e];
                        const int qe = qq % nq;
                        qq /= nq;
                        p[e] = d.left[e] + (d.cell_offset[e] + ce + xq[qe]) * h;
                      }
                    g[fc * n_face_q + q] = bc.second->value(p);
                  }
              HD_CALL(hd_advection_set_dirichlet_values(op, dir, side, g.data(), std::int64_t(g.size())));
            }
      }

      const MatrixFree<dim_x, dim_v, Number> &         data;
      DynamicConvergenceTable &                        table;
      std::shared_ptr<BoundaryDescriptor<dim, Number>> boundary_descriptor;
      std::shared_ptr<VelocityField>                   velocity_field;
      hd_advection *                                   op              = nullptr;
      bool                                             host_sampled_bc = false;
    };

    // operators/advection/cfl.h:58-124 on a Cartesian lattice: the inverse Jacobian is diag(1/h_d), so the critical
    // step of one space is 1 / max_d |u_d / h_d|, and the phase-space value is the minimum over x and v.
    template <int dim_x, int dim_v, typename Number>
    Number
    compute_critical_time_step(const MatrixFree<dim_x, dim_v, Number> &matrix_free, const dealii_compat::Tensor<1, dim_x + dim_v, Number> &u)
    {
      const auto &d        = matrix_free.get_mesh_desc();
      Number      crit[2]  = {std::numeric_limits<Number>::infinity(), std::numeric_limits<Number>::infinity()};
      Number      v_max[2] = {0, 0};
      for (int i = 0; i < dim_x + dim_v; ++i)
        {
          const Number h = (d.right[i] - d.left[i]) / d.n_cells_global[i];
          v_max[i < dim_x ? 0 : 1] = std::max(v_max[i < dim_x ? 0 : 1], std::abs(u[i] / h));
        }
      for (int s = 0; s < 2; ++s)
        crit[s] = Number(1.0) / v_max[s];
      return std::min(crit[0], crit[1]);
    }
  } // namespace advection

  // ---- Vlasov-Poisson pieces (examples/vlasov_poisson/include) ---------------------------------------------------------------
  namespace vp
  {
    // DerivativeContainer (derivative_container.h:30-257): grad(phi) at the quadrature points of every x-cell, on the device
    template <int dim_x, int dim_v, typename Number>
    class DerivativeContainer
    {
    public:
      explicit DerivativeContainer(const MatrixFree<dim_x, dim_v, Number> &mf)
        : mf(mf)
      {
        const auto & d = mf.get_mesh_desc();
        std::int64_t n = dim_x;
        for (int e = 0; e < dim_x; ++e)
          n *= std::int64_t(d.n_cells[e]) * d.n_points;
        HD_CALL(hd_device_malloc(mf.get_communicator().context(), std::size_t(n) * sizeof(double), &table));
      }
      ~DerivativeContainer()
      {
        if (table)
          hd_device_free(mf.get_communicator().context(), table);
      }
      DerivativeContainer(const DerivativeContainer &) = delete;
      DerivativeContainer &operator=(const DerivativeContainer &) = delete;
      double *device_table() const { return static_cast<double *>(table); }

    private:
      const MatrixFree<dim_x, dim_v, Number> &mf;
      void *                                  table = nullptr;
    };

    // LaplaceOperator + PoissonSolver (poisson.h:57-610) and the right-hand side of application.h:529-565 in one object:
    // solve(negative_electric_field, particle_density) fills the gradient table
    template <int dim_x, int dim_v, typename Number>
    class PoissonSolver
    {
    public:
      explicit PoissonSolver(const MatrixFree<dim_x, dim_v, Number> &mf) { HD_CALL(hd_poisson_create(mf.get_mesh(), &ps)); }
      ~PoissonSolver()
      {
        if (ps)
          hd_poisson_destroy(ps);
      }
      PoissonSolver(const PoissonSolver &) = delete;
      PoissonSolver &operator=(const PoissonSolver &) = delete;
      unsigned int
      // poisson.h:593-603: ReductionControl(2 * size, 1e-20, 1e-7); a solve that misses the tolerance throws (HD_CALL turns
      // HD_ERR_NO_CONVERGENCE into hyperdeal::ExcMessage), as dealii::SolverCG does
      solve(DerivativeContainer<dim_x, dim_v, Number> &negative_electric_field, const DeviceVector<Number> &particle_density, const double rel_tol = 1e-7)
      {
        int it = 0;
        const std::int64_t max_it = 2 * particle_density.size();
        HD_CALL(hd_poisson_solve(ps, particle_density.begin(), negative_electric_field.device_table(), rel_tol, int(max_it < 100 ? 100 : (max_it > 100000 ? 100000 : max_it)), &it));
        return it;
      }
      double
      last_relative_residual() const
      {
        double r = 0.0;
        HD_CALL(hd_poisson_last_solve(ps, nullptr, &r));
        return r;
      }

    private:
      hd_poisson *ps = nullptr;
    };

    // diagnostics.h:34-143
    template <int dim_x, int dim_v, typename Number>
    std::array<Number, 6>
    phase_space_diagnostics(const MatrixFree<dim_x, dim_v, Number> &matrix_free, const DeviceVector<Number> &src)
    {
      double out[6];
      HD_CALL(hd_phase_space_diagnostics(matrix_free.get_mesh(), src.begin(), out));
      out[1] = std::sqrt(out[1]);
      return {{Number(out[0]), Number(out[1]), Number(out[2]), Number(out[3]), Number(out[4]), Number(out[5])}};
    }
    template <int dim_x, int dim_v, typename Number>
    std::array<Number, dim_x>
    compute_electric_energy(const MatrixFree<dim_x, dim_v, Number> &matrix_free, const DerivativeContainer<dim_x, dim_v, Number> &negative_electric_field)
    {
      double out[3] = {0, 0, 0};
      HD_CALL(hd_field_energy(matrix_free.get_mesh(), negative_electric_field.device_table(), out));
      std::array<Number, dim_x> r;
      for (int d = 0; d < dim_x; ++d)
        r[d] = Number(out[d]);
      return r;
    }
  } // namespace vp

  namespace advection
  {
    // examples/vlasov_poisson/include/velocity_field_view.h:34-213: a_x = v, a_v = negative electric field
    template <int dim_x, int dim_v, typename Number>
    class PhaseSpaceVelocityFieldView
    {
    public:
      PhaseSpaceVelocityFieldView(const MatrixFree<dim_x, dim_v, Number> &, const vp::DerivativeContainer<dim_x, dim_v, Number> &negative_electric_field)
        : negative_electric_field(negative_electric_field)
      {}
      dealii_compat::Tensor<1, dim_x + dim_v, Number> get_transport_direction() const { return dealii_compat::Tensor<1, dim_x + dim_v, Number>(); }
      const double *                                  phase_space_table() const { return negative_electric_field.device_table(); }

    private:
      const vp::DerivativeContainer<dim_x, dim_v, Number> &negative_electric_field;
    };
  } // namespace advection

  // ---- base/time_integrators.h:48 ---------------------------------------------------------------------------
  template <typename Number, typename VectorType>
  class LowStorageRungeKuttaIntegrator
  {
  public:
    LowStorageRungeKuttaIntegrator(VectorType &vec_Ki, VectorType &vec_Ti, const std::string type, const bool only_Ti_is_ghosted = true)
      : vec_Ki(vec_Ki)
      , vec_Ti(vec_Ti)
      , only_Ti_is_ghosted(only_Ti_is_ghosted)
    {
      HD_CALL(hd_lsrk_create(vec_Ti.get_mesh(), type.c_str(), &rk));
      const int s = hd_lsrk_n_stages(rk);
      bi.resize(s);
      ai.resize(std::max(s - 1, 1));
      hd_lsrk_coefficients(rk, 0, bi.data());
      hd_lsrk_coefficients(rk, 1, ai.data());
      ai.resize(s - 1);
    }
    ~LowStorageRungeKuttaIntegrator()
    {
      if (rk)
        hd_lsrk_destroy(rk);
    }

    // the reference signature (time_integrators.h:66-72; note the (src, dst) order of `op`): unfused — `op` is any
    // callable that fills dst on the device, the stage update is one streaming kernel (hd_lsrk_stage_update)
    void
    perform_time_step(VectorType &solution, const Number &current_time, const Number &time_step, const std::function<void(const VectorType &, VectorType &, const Number)> &op)
    {
      vec_Ti.copy_locally_owned_data_from(solution); // time_integrators.templates.h:146-163 (only_Ti_is_ghosted branch)
      double sum_previous_bi = 0.0;
      for (unsigned int stage = 0; stage < bi.size(); ++stage)
        {
          double c_i = 0.0;
          if (stage > 0)
            {
              c_i = sum_previous_bi + ai[stage - 1];
              sum_previous_bi += bi[stage - 1];
            }
          op(vec_Ti, vec_Ki, Number(current_time + c_i * time_step));
          const double fa = stage + 1 == bi.size() ? 0.0 : ai[stage] * time_step;
          HD_CALL(hd_lsrk_stage_update(vec_Ti.get_mesh(), solution.begin(), vec_Ti.begin(), vec_Ki.begin(), bi[stage] * time_step, fa));
        }
    }

    // fused device path: the operator applies the stage update in its epilogue (32 instead of 48 B/DoF per stage)
    template <int dim_x, int dim_v, int degree, int n_points, typename VelocityField>
    void
    perform_time_step(VectorType &solution, const Number &current_time, const Number &time_step,
                      advection::AdvectionOperation<dim_x, dim_v, degree, n_points, Number, VectorType, VelocityField> &op)
    {
      if (op.needs_host_boundary_data())
        {
          // boundary data that lives in a host dealii::Function must be re-sampled at every stage time: stage by stage
          perform_time_step(solution, current_time, time_step, [&op](const VectorType &src, VectorType &dst, const Number t) { op.apply(dst, src, t); });
          return;
        }
      HD_CALL(hd_lsrk_step(rk, op.handle(), solution.begin(), vec_Ki.begin(), vec_Ti.begin(), double(current_time), double(time_step)));
    }

    // staged fused path for a right-hand side that depends on the stage vector (Vlasov-Poisson, application.h:516-600):
    // `prepare(stage_vector, stage_time)` refreshes the operator's velocity field (rho -> Poisson -> grad(phi)), then ONE
    // kernel applies the operator and the stage update (hd_lsrk_stage_fused); the stage vector alternates between vec_Ti
    // and vec_Ki.  Same numbers as the std::function path, one streaming kernel less per stage.
    template <int dim_x, int dim_v, int degree, int n_points, typename VelocityField>
    void
    perform_time_step(VectorType &solution, const Number &current_time, const Number &time_step,
                      advection::AdvectionOperation<dim_x, dim_v, degree, n_points, Number, VectorType, VelocityField> &op,
                      const std::function<void(const VectorType &, const Number)> &                                       prepare)
    {
      vec_Ti.copy_locally_owned_data_from(solution);
      VectorType *cur = &vec_Ti, *nxt = &vec_Ki;
      double      sum_previous_bi = 0.0;
      for (unsigned int stage = 0; stage < bi.size(); ++stage)
        {
          double c_i = 0.0;
          if (stage > 0)
            {
              c_i = sum_previous_bi + ai[stage - 1];
              sum_previous_bi += bi[stage - 1];
            }
          prepare(*cur, Number(current_time + c_i * time_step));
          HD_CALL(hd_lsrk_stage_fused(rk, op.handle(), int(stage), solution.begin(), cur->begin(), nxt->begin(), nullptr, double(current_time), double(time_step)));
          std::swap(cur, nxt);
        }
    }

    unsigned int n_stages() const { return bi.size(); }

  private:
    VectorType &        vec_Ki;
    VectorType &        vec_Ti;
    const bool          only_Ti_is_ghosted;
    std::vector<double> ai, bi;
    hd_lsrk *           rk = nullptr;
  };

  // ---- base/time_loop_parameters.h:27-49, base/time_loop.h:30-80 ------------------------------------------------
  template <typename Number>
  struct TimeLoopParamters
  {
    Number       time_step            = 0.1;
    Number       start_time           = 0.0;
    Number       final_time           = 20.0;
    unsigned int max_time_step_number = 100000000;
  };

  template <typename Number, typename VectorType>
  class TimeLoop
  {
  public:
    void
    reinit(const TimeLoopParamters<Number> &p)
    {
      time_step            = p.time_step;
      start_time           = p.start_time;
      final_time           = p.final_time;
      max_time_step_number = p.max_time_step_number;
    }
    // source/base/time_loop.cc:35-67
    int
    loop(VectorType &                                                                                                                                         solution,
         const std::function<void(VectorType &, const Number, const Number, const std::function<void(const VectorType &, VectorType &, const Number)> &)> &time_integrator,
         const std::function<void(const VectorType &, VectorType &, const Number)> &                                                                      runnable,
         const std::function<void(const Number)> &                                                                                                        diagnostics)
    {
      unsigned int time_step_number = 1;
      diagnostics(start_time);
      for (Number time = start_time + time_step; time <= final_time * (1.0000000000001) && time_step_number <= max_time_step_number; time += time_step, ++time_step_number)
        {
          time_integrator(solution, time - time_step, time_step, runnable);
          diagnostics(time);
        }
      return time_step_number - 1;
    }
    Number       time_step            = 0.1;
    Number       start_time           = 0.0;
    Number       final_time           = 20.0;
    unsigned int max_time_step_number = 100000000;
  };

  // ---- numerics/vector_tools.h:88, :151 --------------------------------------------------------------------------
  namespace VectorTools
  {
    // nodal interpolation at the Gauss-Lobatto points.  A function with a device implementation is evaluated on the
    // GPU; any other dealii::Function is sampled by the host at the nodes and uploaded (input preparation only).
    template <int degree, int n_points, int dim_x, int dim_v, typename Number, typename VectorType>
    void
    interpolate(const std::shared_ptr<dealii_compat::Function<dim_x + dim_v, Number>> analytical_solution, const MatrixFree<dim_x, dim_v, Number> &matrix_free, VectorType &dst,
                const unsigned int = 0, const unsigned int = 0, const unsigned int = 2, const unsigned int = 2)
    {
      const int id = analytical_solution->device_function_id();
      if (id >= 0)
        {
          HD_CALL(hd_interpolate_builtin(matrix_free.get_mesh(), dst.begin(), id, double(analytical_solution->get_time())));
          return;
        }
      constexpr int       dim = dim_x + dim_v;
      const auto &        d   = matrix_free.get_mesh_desc();
      const int           n   = d.degree + 1;
      std::vector<double> nodes(n);
      hd_mesh_basis(matrix_free.get_mesh(), 0, nodes.data());
      std::vector<Number>               h(matrix_free.n_dofs());
      dealii_compat::Point<dim, Number> p;
      std::int64_t                      idx = 0;
      const std::int64_t                n_cells = matrix_free.n_cells();
      std::int64_t                      dofs_per_cell = 1;
      for (int e = 0; e < dim; ++e)
        dofs_per_cell *= n;
      for (std::int64_t c = 0; c < n_cells; ++c)
        for (std::int64_t i = 0; i < dofs_per_cell; ++i, ++idx)
          {
            std::int64_t cc = c, ii = i;
            for (int e = 0; e < dim; ++e)
              {
                const double he = (d.right[e] - d.left[e]) / d.n_cells_global[e];
                const int    ce = cc % d.n_cells[e];
                cc /= d.n_cells[e];
                const int ie = ii % n;
                ii /= n;
                p[e] = d.left[e] + (d.cell_offset[e] + ce + nodes[ie]) * he;
              }
            h[idx] = analytical_solution->value(p);
          }
      dst.copy_from_host(h);
    }

    // numerics/vector_tools.h:238-315: particle density at the x-space nodes, rho = int f dv with the Gauss-Lobatto rule
    // (quad_no_v = 2, the only rule the drivers use); dst is an x-space vector (DeviceVector::reinit_x)
    template <int degree, int n_points, int dim_x, int dim_v, typename Number, typename Vector_Out, typename Vector_In>
    void
    velocity_space_integration(const MatrixFree<dim_x, dim_v, Number> &data, Vector_Out &dst, const Vector_In &src, const unsigned int = 0, const unsigned int = 0,
                               const unsigned int quad_no_v = 2)
    {
      if (quad_no_v != 2)
        throw ExcNotImplemented("velocity_space_integration with a quadrature other than Gauss-Lobatto (quad_no_v = 2)");
      HD_CALL(hd_velocity_space_integration(data.get_mesh(), dst.begin(), src.begin()));
    }

    // {L2 norm of u_h, L2 norm of u_h - f} at the quadrature points; device reduction, device-side f only
    template <int degree, int n_points, int dim_x, int dim_v, typename Number, typename VectorType>
    std::array<Number, 2>
    norm_and_error(const std::shared_ptr<dealii_compat::Function<dim_x + dim_v, Number>> analytical_solution, const MatrixFree<dim_x, dim_v, Number> &matrix_free, const VectorType &src,
                   const unsigned int = 0, const unsigned int = 0, const unsigned int = 0, const unsigned int = 0)
    {
      const int id = analytical_solution->device_function_id();
      if (id < 0)
        throw ExcNotImplemented("norm_and_error against a host-only Function");
      double out[2];
      HD_CALL(hd_norm_and_error_builtin(matrix_free.get_mesh(), src.begin(), id, double(analytical_solution->get_time()), out));
      return {{Number(std::sqrt(out[0])), Number(std::sqrt(out[1]))}};
    }
  } // namespace VectorTools

  // ---- more than one GPU in ONE process (hd_multi_*) ----------------------------------------------------------------
  // The reference runs PartitionX x PartitionV MPI ranks (examples/advection/include/application.h:150-176; the process
  // grid of base/mpi.h create_rectangular_comm) and partitions the x- and the v-triangulation among the rows and columns
  // of that grid.  Here the same two numbers cut the Cartesian lattice into bricks, one per GPU of this process; ghost
  // faces travel over NVLink peer stores inside the library.  The classes below carry the same member names as their
  // single-GPU counterparts so a driver's set-up / time loop reads the same.
  namespace multi
  {
    // cut `parts` ways: the slowest direction first, each direction as far as its cell count allows
    template <int dim>
    std::array<int, dim>
    brick_grid(const CartesianLattice<dim> &lattice, int parts)
    {
      std::array<int, dim> g;
      g.fill(1);
      for (int d = dim - 1; d >= 0 && parts > 1; --d)
        {
          int a = parts, b = int(lattice.n_cells[d]);
          while (b)
            {
              const int t = a % b;
              a           = b;
              b           = t;
            }
          g[d] = a; // gcd(parts, n_cells[d])
          parts /= a;
        }
      if (parts != 1)
        throw ExcMessage("the lattice cannot be cut into the requested number of bricks");
      return g;
    }

    template <typename Number>
    class DistributedDeviceVector;

    template <int dim_x, int dim_v, typename Number = double>
    class MatrixFree
    {
    public:
      static const int dim = dim_x + dim_v;
      MatrixFree(const CartesianLattice<dim_x> &matrix_free_x, const CartesianLattice<dim_v> &matrix_free_v, const int partition_x, const int partition_v)
        : lattice_x(matrix_free_x)
        , lattice_v(matrix_free_v)
        , size_x(partition_x)
        , size_v(partition_v)
      {}
      ~MatrixFree()
      {
        if (mm)
          hd_multi_destroy(mm);
      }
      MatrixFree(const MatrixFree &) = delete;
      MatrixFree &operator=(const MatrixFree &) = delete;

      void
      reinit()
      {
        if (lattice_x.degree != lattice_v.degree || lattice_x.n_points != lattice_v.n_points || lattice_x.collocation != lattice_v.collocation)
          throw ExcNotImplemented("different degree / quadrature in x and v");
        hd_mesh_desc d{};
        d.dim_x       = dim_x;
        d.dim_v       = dim_v;
        d.degree      = lattice_x.degree;
        d.n_points    = lattice_x.n_points;
        d.collocation = lattice_x.collocation;
        d.number_type = internal::number_type<Number>();
        int grid[HD_MAX_DIM];
        for (int i = 0; i < HD_MAX_DIM; ++i)
          {
            d.left[i]           = 0.0;
            d.right[i]          = 1.0;
            d.n_cells_global[i] = d.n_cells[i] = 1;
            d.cell_offset[i]                   = 0;
            d.side_kind[i][0] = d.side_kind[i][1] = HD_SIDE_PERIODIC_LOCAL;
            grid[i]                               = 1;
          }
        const auto gx = brick_grid<dim_x>(lattice_x, size_x);
        const auto gv = brick_grid<dim_v>(lattice_v, size_v);
        for (int i = 0; i < dim; ++i)
          {
            const bool in_x     = i < dim_x;
            const int  j        = in_x ? i : i - dim_x;
            d.left[i]           = in_x ? lattice_x.left[j] : lattice_v.left[j];
            d.right[i]          = in_x ? lattice_x.right[j] : lattice_v.right[j];
            d.n_cells_global[i] = d.n_cells[i] = in_x ? lattice_x.n_cells[j] : lattice_v.n_cells[j];
            const bool periodic                = in_x ? lattice_x.periodic : lattice_v.periodic;
            d.side_kind[i][0] = d.side_kind[i][1] = periodic ? HD_SIDE_PERIODIC_LOCAL : HD_SIDE_DIRICHLET;
            grid[i]                               = in_x ? gx[j] : gv[j];
          }
        if (mm)
          hd_multi_destroy(mm);
        mm = nullptr;
        HD_CALL(hd_multi_create(size_x * size_v, nullptr, &d, grid, &mm));
        desc = d;
      }
      void
      initialize_dof_vector(DistributedDeviceVector<Number> &vec, const unsigned int = 0, const bool = true, const bool = true) const
      {
        vec.reinit(mm);
      }
      hd_multi *          get_multi() const { return mm; }
      const hd_mesh_desc &get_mesh_desc() const { return desc; } // the GLOBAL lattice
      std::int64_t        n_dofs() const { return hd_multi_n_dofs(mm); }
      int                 n_bricks() const { return hd_multi_n_gpus(mm); }

    private:
      CartesianLattice<dim_x> lattice_x;
      CartesianLattice<dim_v> lattice_v;
      const int               size_x, size_v;
      hd_multi *              mm = nullptr;
      hd_mesh_desc            desc{};
    };

    // one device pointer per brick
    template <typename Number>
    class DistributedDeviceVector
    {
    public:
      DistributedDeviceVector() = default;
      ~DistributedDeviceVector() { clear(); }
      DistributedDeviceVector(const DistributedDeviceVector &) = delete;
      DistributedDeviceVector &operator=(const DistributedDeviceVector &) = delete;
      void
      reinit(hd_multi *m)
      {
        clear();
        mm = m;
        ptrs.assign(hd_multi_n_gpus(mm), nullptr);
        HD_CALL(hd_multi_vector_alloc(mm, ptrs.data()));
      }
      void
      clear()
      {
        if (!ptrs.empty())
          hd_multi_vector_free(mm, ptrs.data());
        ptrs.clear();
      }
      void *const *bricks() const { return ptrs.data(); }
      std::int64_t size() const { return hd_multi_n_dofs(mm); }
      // host transfers in the ordering of the single-GPU lattice
      void
      copy_from_host(const std::vector<Number> &h)
      {
        if (std::int64_t(h.size()) != size())
          throw ExcMessage("copy_from_host: size mismatch");
        HD_CALL(hd_multi_vector_copy_in(mm, ptrs.data(), h.data()));
      }
      void
      copy_to_host(std::vector<Number> &h) const
      {
        h.resize(size());
        HD_CALL(hd_multi_vector_copy_out(mm, ptrs.data(), h.data()));
      }

    private:
      hd_multi *          mm = nullptr;
      std::vector<void *> ptrs;
    };

    // advection::AdvectionOperation on all bricks (constant velocity; Dirichlet data from a device-side Function)
    template <int dim_x, int dim_v, typename Number>
    class AdvectionOperation
    {
    public:
      static const int dim = dim_x + dim_v;
      explicit AdvectionOperation(const MatrixFree<dim_x, dim_v, Number> &data)
        : data(data)
      {}
      ~AdvectionOperation()
      {
        if (op)
          hd_multi_advection_destroy(op);
      }
      void
      reinit(const std::shared_ptr<advection::BoundaryDescriptor<dim, Number>> boundary_descriptor, const dealii_compat::Tensor<1, dim, Number> &transport_direction,
             const advection::AdvectionOperationParamters &                   additional_data)
      {
        double velocity[HD_MAX_DIM] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < dim; ++i)
          velocity[i] = transport_direction[i];
        if (op)
          hd_multi_advection_destroy(op);
        op = nullptr;
        HD_CALL(hd_multi_advection_create(data.get_multi(), additional_data.factor_skew, velocity, &op));
        const auto &d = data.get_mesh_desc();
        bool        dirichlet = false;
        for (int i = 0; i < dim; ++i)
          dirichlet |= d.side_kind[i][0] == HD_SIDE_DIRICHLET;
        if (dirichlet)
          {
            int id = -1;
            for (const auto &bc : boundary_descriptor->dirichlet_bc)
              id = bc.second->device_function_id();
            if (id < 0)
              throw ExcNotImplemented("multi-GPU Dirichlet data from a host-only Function");
            HD_CALL(hd_multi_advection_set_dirichlet_builtin(op, id));
          }
      }
      void
      apply(DistributedDeviceVector<Number> &dst, const DistributedDeviceVector<Number> &src, const Number current_time)
      {
        HD_CALL(hd_multi_advection_apply(op, dst.bricks(), src.bricks(), double(current_time)));
      }
      hd_multi_advection *handle() const { return op; }
      std::string         kernel_name() const { return hd_multi_advection_kernel_name(op); }

    private:
      const MatrixFree<dim_x, dim_v, Number> &data;
      hd_multi_advection *                    op = nullptr;
    };

    template <typename Number>
    class LowStorageRungeKuttaIntegrator
    {
    public:
      using VectorType = DistributedDeviceVector<Number>;
      LowStorageRungeKuttaIntegrator(hd_multi *mm, VectorType &vec_Ki, VectorType &vec_Ti, const std::string type)
        : vec_Ki(vec_Ki)
        , vec_Ti(vec_Ti)
      {
        HD_CALL(hd_multi_lsrk_create(mm, type.c_str(), &rk));
      }
      ~LowStorageRungeKuttaIntegrator()
      {
        if (rk)
          hd_multi_lsrk_destroy(rk);
      }
      template <int dim_x, int dim_v>
      void
      perform_time_step(VectorType &solution, const Number &current_time, const Number &time_step, AdvectionOperation<dim_x, dim_v, Number> &op)
      {
        HD_CALL(hd_multi_lsrk_step(rk, op.handle(), solution.bricks(), vec_Ki.bricks(), vec_Ti.bricks(), double(current_time), double(time_step)));
      }

    private:
      VectorType &   vec_Ki;
      VectorType &   vec_Ti;
      hd_multi_lsrk *rk = nullptr;
    };

    template <int dim_x, int dim_v, typename Number>
    void
    interpolate(const std::shared_ptr<dealii_compat::Function<dim_x + dim_v, Number>> f, const MatrixFree<dim_x, dim_v, Number> &matrix_free, DistributedDeviceVector<Number> &dst)
    {
      const int id = f->device_function_id();
      if (id < 0)
        throw ExcNotImplemented("multi-GPU interpolation of a host-only Function");
      HD_CALL(hd_multi_interpolate_builtin(matrix_free.get_multi(), dst.bricks(), id, double(f->get_time())));
    }

    template <int dim_x, int dim_v, typename Number>
    std::array<Number, 2>
    norm_and_error(const std::shared_ptr<dealii_compat::Function<dim_x + dim_v, Number>> f, const MatrixFree<dim_x, dim_v, Number> &matrix_free, const DistributedDeviceVector<Number> &src)
    {
      const int id = f->device_function_id();
      if (id < 0)
        throw ExcNotImplemented("norm_and_error against a host-only Function");
      double out[2];
      HD_CALL(hd_multi_norm_and_error_builtin(matrix_free.get_multi(), src.bricks(), id, double(f->get_time()), out));
      return {{Number(std::sqrt(out[0])), Number(std::sqrt(out[1]))}};
    }
  } // namespace multi
} // namespace hyperdeal

#ifndef HYPERDEAL_B200_NO_DEALII_ALIAS
namespace dealii = hyperdeal::dealii_compat;
#endif

#endif
