// examples/advection re-hosted on the B200 library: the same parameter files (examples/advection/tests/*.json),
// the same set-up sequence (examples/advection/include/application.h:97-420), time loop (:423-560) and output format
// ("   Time:%10.3e, norm: %17.10e, error: %17.10e") as the reference driver, with every vector and the whole
// right-hand side on the GPU through hyperdeal_b200.hpp -> libhdgpu.so.
//
//   advection <file.json> [more.json ...]      (DIM_X, DIM_V, DEGREE, N_POINTS from <name>.configuration beside the
//                                               json, or --config "DIM_X=2 DIM_V=2 DEGREE=3 N_POINTS=4")
// Environment: HD_DRIVER_UNFUSED=1 uses the std::function time-integrator path (operator and stage update as separate
// kernels, exactly the reference's call structure); HD_DRIVER_HOST_FUNCTIONS=1 gives the operator host-only
// dealii::Function objects for the initial and boundary data (sampled on the host, uploaded per stage).
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "hyperdeal_b200.hpp"
#include "json_parameters.hpp"

namespace hyperdeal
{
  namespace advection
  {
    namespace hyperrectangle
    {
      // examples/advection/cases/hyperrectangle.h:29-66
      template <int DIM, typename Number = double>
      class ExactSolution : public dealii::Function<DIM, Number>
      {
      public:
        explicit ExactSolution(const bool on_device = true)
          : on_device(on_device)
        {
          adv[0] = 1.;
          if (DIM > 1)
            adv[1] = 0.15;
          if (DIM > 2)
            adv[2] = -0.05;
        }
        Number
        value(const dealii::Point<DIM, Number> &p, const unsigned int = 0) const override
        {
          const double pi = 3.14159265358979323846264338327950288;
          const double t  = this->get_time();
          double       r  = std::sin(2. * (p[0] - t * adv[0]) * pi);
          for (int d = 1; d < DIM; ++d)
            r *= std::cos(2. * (p[d] - t * adv[d]) * pi);
          return r;
        }
        int device_function_id() const override { return on_device ? HD_FN_HYPERRECTANGLE : -1; }
        dealii::Tensor<1, DIM, Number> get_transport_direction() const { return adv; }

      private:
        dealii::Tensor<1, DIM, Number> adv;
        const bool                     on_device;
      };
    } // namespace hyperrectangle

    struct Parameters // examples/advection/include/parameters.h:30-148
    {
      explicit Parameters(const JsonParameters &prm)
      {
        do_collocation                          = prm.get_bool("SpatialDiscretization/DoCollocation", false);
        time_loop_parameters.time_step          = prm.get_double("TemporalDiscretization/TimeStep", 0.1);
        time_loop_parameters.start_time         = prm.get_double("TemporalDiscretization/StartTime", 0.0);
        time_loop_parameters.final_time         = prm.get_double("TemporalDiscretization/FinalTime", 20.0);
        time_loop_parameters.max_time_step_number = prm.get_int("TemporalDiscretization/MaxTimeStepNumber", 100000000);
        rk_type                                 = prm.get("TemporalDiscretization/RKType", "rk45");
        cfl_number                              = prm.get_double("TemporalDiscretization/CFLNumber", 0.3);
        dignostics_enabled                      = prm.get_bool("TemporalDiscretization/DiagnosticsEnabled", true);
        dignostics_tick                         = prm.get_double("TemporalDiscretization/DiagnosticsTick", 0.1);
        advection_operation_parameters.factor_skew = prm.get_double("AdvectionOperation/SkewFactor", 0.0);
        do_ghost_faces                          = prm.get_bool("Matrixfree/GhostFaces", true);
        do_buffering                            = prm.get_bool("Matrixfree/DoBuffering", false);
        use_ecl                                 = prm.get_bool("Matrixfree/UseECL", true);
        n_refinements_x                         = prm.get_int("Case/NRefinementsX", 0);
        n_refinements_v                         = prm.get_int("Case/NRefinementsV", 0);
        periodic_x                              = prm.get_bool("Case/PeriodicX", true);
        periodic_v                              = prm.get_bool("Case/PeriodicV", true);
        const char *xyz[3]                      = {"X", "Y", "Z"};
        for (int d = 0; d < 3; ++d)
          {
            n_subdivisions_x[d] = prm.get_int(std::string("Case/NSubdivisionsX/") + xyz[d], 4);
            n_subdivisions_v[d] = prm.get_int(std::string("Case/NSubdivisionsV/") + xyz[d], 4);
          }
        case_name = prm.get("General/Case", "hyperrectangle");
        dim_x     = prm.get_int("General/DimX", 0);
        dim_v     = prm.get_int("General/DimV", 0);
        degree    = prm.get_int("General/DegreeX", 0);
        // process grid of the reference run (parameters.h:62-63); here: bricks = GPUs of this process.  The environment
        // overrides the file the way `mpirun -np N` does for the reference's tests (tests/*.mpirun=N.out)
        partition_x = prm.get_int("General/PartitionX", 1);
        partition_v = prm.get_int("General/PartitionV", 1);
        if (const char *e = std::getenv("HD_PARTITION_X"))
          partition_x = std::atoi(e);
        if (const char *e = std::getenv("HD_PARTITION_V"))
          partition_v = std::atoi(e);
      }
      bool                        do_collocation;
      TimeLoopParamters<double>   time_loop_parameters;
      std::string                 rk_type;
      double                      cfl_number;
      bool                        dignostics_enabled;
      double                      dignostics_tick;
      AdvectionOperationParamters advection_operation_parameters;
      bool                        do_ghost_faces, do_buffering, use_ecl;
      unsigned int                n_refinements_x, n_refinements_v;
      bool                        periodic_x, periodic_v;
      unsigned int                n_subdivisions_x[3], n_subdivisions_v[3];
      std::string                 case_name;
      int                         dim_x, dim_v, degree;
      int                         partition_x = 1, partition_v = 1;
    };

    // examples/advection/include/application.h:60-640
    template <int dim_x, int dim_v, int degree, int n_points, typename Number>
    class Application
    {
    public:
      static const int dim = dim_x + dim_v;
      using VectorType     = DeviceVector<Number>;
      using VelocityField  = ConstantVelocityFieldView<dim, Number>;

      Application(const DeviceCommunicator &comm, DynamicConvergenceTable &table)
        : comm(comm)
        , table(table)
      {}

      void
      reinit(Parameters &param)
      {
        this->param = &param;
        if (param.case_name != "hyperrectangle")
          throw ExcNotImplemented("case " + param.case_name + " (Cartesian lattices only)");
        CartesianLattice<dim_x> lx;
        CartesianLattice<dim_v> lv;
        for (int d = 0; d < dim_x; ++d)
          {
            lx.left[d]    = -1.0;
            lx.right[d]   = +1.0;
            lx.n_cells[d] = param.n_subdivisions_x[d] << param.n_refinements_x;
          }
        for (int d = 0; d < dim_v; ++d)
          {
            lv.left[d]    = -1.0;
            lv.right[d]   = +1.0;
            lv.n_cells[d] = param.n_subdivisions_v[d] << param.n_refinements_v;
          }
        lx.periodic = param.periodic_x;
        lv.periodic = param.periodic_v;
        lx.degree = lv.degree = degree;
        lx.n_points = lv.n_points = n_points;
        lx.collocation = lv.collocation = param.do_collocation;
        matrix_free.reset(new MatrixFree<dim_x, dim_v, Number>(comm, lx, lv));
        typename MatrixFree<dim_x, dim_v, Number>::AdditionalData ad;
        ad.do_ghost_faces = param.do_ghost_faces;
        ad.do_buffering   = param.do_buffering;
        ad.use_ecl        = param.use_ecl;
        matrix_free->reinit(ad);
        matrix_free->initialize_dof_vector(vct_Ki, 0, !param.use_ecl, true);
        matrix_free->initialize_dof_vector(vct_Ti, 0, true, true);
        matrix_free->initialize_dof_vector(vct_solution, 0, !param.use_ecl, true);

        const bool host_functions = std::getenv("HD_DRIVER_HOST_FUNCTIONS") != nullptr;
        analytical_solution.reset(new hyperrectangle::ExactSolution<dim, Number>());
        {
          std::shared_ptr<dealii::Function<dim, Number>> initial(new hyperrectangle::ExactSolution<dim, Number>(!host_functions));
          VectorTools::interpolate<degree, degree + 1>(initial, *matrix_free, vct_solution, 0, 0, 2, 2);
        }
        boundary_descriptor.reset(new BoundaryDescriptor<dim, Number>());
        boundary_descriptor->dirichlet_bc[0].reset(new hyperrectangle::ExactSolution<dim, Number>(!host_functions));
        boundary_descriptor->dirichlet_bc[1].reset(new hyperrectangle::ExactSolution<dim, Number>(!host_functions));
        const auto transport_direction = hyperrectangle::ExactSolution<dim, Number>().get_transport_direction();
        velocity_field                 = std::make_shared<VelocityField>(transport_direction);
        advection_operation.reset(new AdvectionOperation<dim_x, dim_v, degree, n_points, Number, VectorType, VelocityField>(*matrix_free, table));
        advection_operation->reinit(boundary_descriptor, velocity_field, param.advection_operation_parameters);

        // application.h:369-392
        auto &       tl                 = param.time_loop_parameters;
        const Number critical_time_step = compute_critical_time_step(*matrix_free, transport_direction);
        const Number dt                 = std::min<Number>(tl.time_step, param.cfl_number * critical_time_step / std::pow(degree, 1.5));
        tl.time_step                    = (tl.final_time - tl.start_time) / std::ceil((tl.final_time - tl.start_time) / dt);
        time_loop.reinit(tl);
      }

      void
      solve()
      {
        LowStorageRungeKuttaIntegrator<Number, VectorType> time_integrator(vct_Ki, vct_Ti, param->rk_type, param->use_ecl);
        const bool                                         unfused = std::getenv("HD_DRIVER_UNFUSED") != nullptr;
        std::array<Number, 2>                              error;
        const auto &                                       tl = param->time_loop_parameters;

        const unsigned int time_steps = time_loop.loop(
          vct_solution,
          [&](auto &solution, const auto cur_time, const auto time_step, const auto &runnable) {
            if (unfused)
              time_integrator.perform_time_step(solution, cur_time, time_step, runnable);
            else
              time_integrator.perform_time_step(solution, cur_time, time_step, *advection_operation);
          },
          [&](const VectorType &src, VectorType &dst, const Number cur_time) { advection_operation->apply(dst, src, cur_time); },
          [&](const Number cur_time) {
            if (!param->dignostics_enabled ||
                (cur_time != tl.start_time && static_cast<int>((cur_time + 0.00000000001 - tl.start_time) / param->dignostics_tick) ==
                                                static_cast<int>((cur_time + 0.00000000001 - tl.start_time - tl.time_step) / param->dignostics_tick)))
              return;
            analytical_solution->set_time(cur_time);
            error = VectorTools::norm_and_error<degree, n_points>(analytical_solution, *matrix_free, vct_solution, 0, 0, 0, 0);
            printf("   Time:%10.3e, norm: %17.10e, error: %17.10e\n", cur_time, error[0], error[1]);
          });
        table.set("info->time_steps", time_steps);
        table.set("info->n_dofs", double(matrix_free->n_dofs()));
      }

    private:
      const DeviceCommunicator &                                                                         comm;
      DynamicConvergenceTable &                                                                          table;
      Parameters *                                                                                       param = nullptr;
      std::unique_ptr<MatrixFree<dim_x, dim_v, Number>>                                                  matrix_free;
      VectorType                                                                                         vct_Ki, vct_Ti, vct_solution;
      std::shared_ptr<dealii::Function<dim, Number>>                                                     analytical_solution;
      std::shared_ptr<BoundaryDescriptor<dim, Number>>                                                   boundary_descriptor;
      std::shared_ptr<VelocityField>                                                                     velocity_field;
      std::unique_ptr<AdvectionOperation<dim_x, dim_v, degree, n_points, Number, VectorType, VelocityField>> advection_operation;
      TimeLoop<Number, VectorType>                                                                       time_loop;
    };

    // the same application on PartitionX x PartitionV GPUs of this process (hyperdeal::multi, hd_multi_*): identical
    // set-up sequence, time step and diagnostics; every stage = one ghost exchange over NVLink + one fused kernel per GPU
    template <int dim_x, int dim_v, int degree, int n_points, typename Number>
    class MultiApplication
    {
    public:
      static const int dim = dim_x + dim_v;
      using VectorType     = multi::DistributedDeviceVector<Number>;

      explicit MultiApplication(DynamicConvergenceTable &table)
        : table(table)
      {}

      void
      reinit(Parameters &param)
      {
        this->param = &param;
        if (param.case_name != "hyperrectangle")
          throw ExcNotImplemented("case " + param.case_name + " (Cartesian lattices only)");
        if (std::getenv("HD_DRIVER_UNFUSED") || std::getenv("HD_DRIVER_HOST_FUNCTIONS"))
          throw ExcNotImplemented("HD_DRIVER_UNFUSED / HD_DRIVER_HOST_FUNCTIONS on more than one GPU");
        CartesianLattice<dim_x> lx;
        CartesianLattice<dim_v> lv;
        for (int d = 0; d < dim_x; ++d)
          {
            lx.left[d]    = -1.0;
            lx.right[d]   = +1.0;
            lx.n_cells[d] = param.n_subdivisions_x[d] << param.n_refinements_x;
          }
        for (int d = 0; d < dim_v; ++d)
          {
            lv.left[d]    = -1.0;
            lv.right[d]   = +1.0;
            lv.n_cells[d] = param.n_subdivisions_v[d] << param.n_refinements_v;
          }
        lx.periodic = param.periodic_x;
        lv.periodic = param.periodic_v;
        lx.degree = lv.degree = degree;
        lx.n_points = lv.n_points = n_points;
        lx.collocation = lv.collocation = param.do_collocation;
        matrix_free.reset(new multi::MatrixFree<dim_x, dim_v, Number>(lx, lv, param.partition_x, param.partition_v));
        matrix_free->reinit();
        matrix_free->initialize_dof_vector(vct_Ki);
        matrix_free->initialize_dof_vector(vct_Ti);
        matrix_free->initialize_dof_vector(vct_solution);
        analytical_solution.reset(new hyperrectangle::ExactSolution<dim, Number>());
        multi::interpolate<dim_x, dim_v, Number>(analytical_solution, *matrix_free, vct_solution);
        boundary_descriptor.reset(new BoundaryDescriptor<dim, Number>());
        boundary_descriptor->dirichlet_bc[0].reset(new hyperrectangle::ExactSolution<dim, Number>());
        boundary_descriptor->dirichlet_bc[1].reset(new hyperrectangle::ExactSolution<dim, Number>());
        const auto transport_direction = hyperrectangle::ExactSolution<dim, Number>().get_transport_direction();
        advection_operation.reset(new multi::AdvectionOperation<dim_x, dim_v, Number>(*matrix_free));
        advection_operation->reinit(boundary_descriptor, transport_direction, param.advection_operation_parameters);

        // application.h:369-392 (the critical time step depends on the global lattice only)
        auto &       tl                 = param.time_loop_parameters;
        const auto & d                  = matrix_free->get_mesh_desc();
        Number       v_max[2]           = {0, 0};
        for (int i = 0; i < dim; ++i)
          v_max[i < dim_x ? 0 : 1] = std::max<Number>(v_max[i < dim_x ? 0 : 1], std::abs(transport_direction[i] / ((d.right[i] - d.left[i]) / d.n_cells_global[i])));
        const Number critical_time_step = std::min(Number(1.0) / v_max[0], Number(1.0) / v_max[1]);
        const Number dt                 = std::min<Number>(tl.time_step, param.cfl_number * critical_time_step / std::pow(degree, 1.5));
        tl.time_step                    = (tl.final_time - tl.start_time) / std::ceil((tl.final_time - tl.start_time) / dt);
        time_loop.reinit(tl);
      }

      void
      solve()
      {
        multi::LowStorageRungeKuttaIntegrator<Number> time_integrator(matrix_free->get_multi(), vct_Ki, vct_Ti, param->rk_type);
        std::array<Number, 2>                         error;
        const auto &                                  tl = param->time_loop_parameters;
        const unsigned int time_steps = time_loop.loop(
          vct_solution,
          [&](auto &solution, const auto cur_time, const auto time_step, const auto &) { time_integrator.perform_time_step(solution, cur_time, time_step, *advection_operation); },
          [&](const VectorType &src, VectorType &dst, const Number cur_time) { advection_operation->apply(dst, src, cur_time); },
          [&](const Number cur_time) {
            if (!param->dignostics_enabled ||
                (cur_time != tl.start_time && static_cast<int>((cur_time + 0.00000000001 - tl.start_time) / param->dignostics_tick) ==
                                                static_cast<int>((cur_time + 0.00000000001 - tl.start_time - tl.time_step) / param->dignostics_tick)))
              return;
            analytical_solution->set_time(cur_time);
            error = multi::norm_and_error<dim_x, dim_v, Number>(analytical_solution, *matrix_free, vct_solution);
            printf("   Time:%10.3e, norm: %17.10e, error: %17.10e\n", cur_time, error[0], error[1]);
          });
        table.set("info->time_steps", time_steps);
        table.set("info->n_dofs", double(matrix_free->n_dofs()));
        table.set("info->n_gpus", double(matrix_free->n_bricks()));
        printf("   bricks: %d, kernel: %s\n", matrix_free->n_bricks(), advection_operation->kernel_name().c_str());
      }

    private:
      DynamicConvergenceTable &                                     table;
      Parameters *                                                  param = nullptr;
      std::unique_ptr<multi::MatrixFree<dim_x, dim_v, Number>>      matrix_free;
      VectorType                                                    vct_Ki, vct_Ti, vct_solution;
      std::shared_ptr<dealii::Function<dim, Number>>                analytical_solution;
      std::shared_ptr<BoundaryDescriptor<dim, Number>>              boundary_descriptor;
      std::unique_ptr<multi::AdvectionOperation<dim_x, dim_v, Number>> advection_operation;
      TimeLoop<Number, VectorType>                                  time_loop;
    };
  } // namespace advection
} // namespace hyperdeal

namespace
{
  struct Configuration
  {
    int dim_x = 0, dim_v = 0, degree = 0, n_points = 0;
  };

  bool
  parse_configuration(const std::string &text, Configuration &c)
  {
    const auto field = [&](const char *key, int &out) {
      const auto p = text.find(key);
      if (p != std::string::npos)
        out = std::atoi(text.c_str() + p + std::strlen(key));
    };
    field("DIM_X=", c.dim_x);
    field("DIM_V=", c.dim_v);
    field("DEGREE=", c.degree);
    field("N_POINTS=", c.n_points);
    return c.dim_x > 0 && c.dim_v > 0 && c.degree > 0 && c.n_points > 0;
  }

  template <int dim_x, int dim_v, int degree, int n_points>
  void
  run_application(const hyperdeal::DeviceCommunicator &comm, hyperdeal::DynamicConvergenceTable &table, hyperdeal::advection::Parameters &param)
  {
    if (param.partition_x * param.partition_v > 1)
      {
        hyperdeal::advection::MultiApplication<dim_x, dim_v, degree, n_points, double> app(table);
        app.reinit(param);
        app.solve();
        return;
      }
    hyperdeal::advection::Application<dim_x, dim_v, degree, n_points, double> app(comm, table);
    app.reinit(param);
    app.solve();
  }

  // the reference is compiled once per (DIM_X, DIM_V, DEGREE, N_POINTS) (examples/advection/tests/*.configuration);
  // this driver carries the instantiations its tests use and dispatches at run time
  template <int dim_x, int dim_v>
  void
  dispatch_degree(const Configuration &c, const hyperdeal::DeviceCommunicator &comm, hyperdeal::DynamicConvergenceTable &table, hyperdeal::advection::Parameters &param)
  {
#define HD_CASE(K, Q)                          \
  if (c.degree == K && c.n_points == Q)        \
    {                                          \
      run_application<dim_x, dim_v, K, Q>(comm, table, param); \
      return;                                  \
    }
    HD_CASE(2, 3)
    HD_CASE(3, 4)
    HD_CASE(3, 5)
    HD_CASE(4, 5)
    HD_CASE(5, 6)
#undef HD_CASE
    throw hyperdeal::ExcNotImplemented("DEGREE=" + std::to_string(c.degree) + " N_POINTS=" + std::to_string(c.n_points));
  }

  void
  run(const std::string &file_name, const std::string &config_override, const hyperdeal::DeviceCommunicator &comm, hyperdeal::DynamicConvergenceTable &table)
  {
    hyperdeal::JsonParameters       prm(file_name);
    hyperdeal::advection::Parameters param(prm);
    Configuration                   c;
    c.dim_x  = param.dim_x;
    c.dim_v  = param.dim_v;
    c.degree = param.degree;
    if (!config_override.empty())
      parse_configuration(config_override, c);
    else
      {
        // <dir>/<name>.<case>.json -> <dir>/<name>.configuration
        const auto  slash = file_name.find_last_of('/');
        const auto  dot   = file_name.find('.', slash == std::string::npos ? 0 : slash);
        std::ifstream in(file_name.substr(0, dot) + ".configuration");
        if (in)
          {
            std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
            parse_configuration(text, c);
          }
      }
    if (c.n_points == 0)
      c.n_points = c.degree + 1;
    if (param.dim_x != 0 && (param.dim_x != c.dim_x || param.dim_v != c.dim_v || param.degree != c.degree))
      throw hyperdeal::ExcMessage("Degrees/dimensions of the parameter file and of the configuration do not match!");
    if (c.dim_x == 1 && c.dim_v == 1)
      dispatch_degree<1, 1>(c, comm, table, param);
    else if (c.dim_x == 2 && c.dim_v == 2)
      dispatch_degree<2, 2>(c, comm, table, param);
    else if (c.dim_x == 3 && c.dim_v == 3)
      dispatch_degree<3, 3>(c, comm, table, param);
    else
      throw hyperdeal::ExcNotImplemented("DIM_X=" + std::to_string(c.dim_x) + " DIM_V=" + std::to_string(c.dim_v));
  }
} // namespace

int
main(int argc, char **argv)
{
  try
    {
      if (argc == 1)
        {
          printf("ERROR: No .json parameter files has been provided!\n");
          return 1;
        }
      std::string config;
      int         first = 1;
      if (argc >= 4 && std::string(argv[1]) == "--config")
        {
          config = argv[2];
          first  = 3;
        }
      hyperdeal::DeviceCommunicator      comm(std::getenv("HD_DEVICE") ? std::atoi(std::getenv("HD_DEVICE")) : 0);
      hyperdeal::DynamicConvergenceTable table;
      for (int i = first; i < argc; ++i)
        {
          std::cout << std::string(argv[i]) << std::endl;
          run(argv[i], config, comm, table);
        }
      table.print();
    }
  catch (std::exception &exc)
    {
      std::cerr << std::endl
                << std::endl
                << "----------------------------------------------------" << std::endl;
      std::cerr << "Exception on processing: " << std::endl
                << exc.what() << std::endl
                << "Aborting!" << std::endl
                << "----------------------------------------------------" << std::endl;
      return 1;
    }
  return 0;
}
