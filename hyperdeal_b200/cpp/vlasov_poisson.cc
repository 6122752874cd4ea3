// examples/vlasov_poisson re-hosted on the B200 library: the same parameter file (examples/vlasov_poisson/tests/*.json), set-up
// (include/application.h:97-470, cases/hyperrectangle.h), right-hand side (application.h:516-600: density, field solve, electric
// field, phase-space advection) and diagnostics output (:619-660, "time  energy components  mass  l2norm  kinetic energy
// momentum") as the reference driver, with every step on the GPU through hyperdeal_b200.hpp -> libhdgpu.so.
//
//   vlasov_poisson <file.json>      (DIM_X, DIM_V, DEGREE, N_POINTS from <name>.configuration beside the json)
//
// Reproduces examples/vlasov_poisson/tests/vp_2D_2D_k3.hyperrectangle_01.out on the GPU (tests/test_zz_vp_device_gpu.py), with
// fused stages (field refresh, then ONE kernel for operator + stage update) or, HD_DRIVER_UNFUSED=1, the reference's call structure.
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>

#include "hyperdeal_b200.hpp"
#include "json_parameters.hpp"

namespace hyperdeal
{
  namespace vp
  {
    namespace hyperrectangle
    {
      // cases/hyperrectangle.h:29-76: (1 + 0.01 cos(x_0 / 2)) prod_d exp(-v_d^2 / 2) / sqrt(2 pi); host-only field (sampled at the nodes)
      template <int dim_x, int dim_v, typename Number = double>
      class ExactSolution : public dealii::Function<dim_x + dim_v, Number>
      {
      public:
        Number
        value(const dealii::Point<dim_x + dim_v, Number> &p, const unsigned int = 0) const override
        {
          double result = 1.0 + 0.01 * std::cos(0.5 * p[0]);
          for (int d = dim_x; d < dim_x + dim_v; ++d)
            result = result * std::exp(-0.5 * p[d] * p[d]) / std::sqrt(2.0 * 3.14159265358979323846264338327950288);
          return result;
        }
        dealii::Tensor<1, dim_x + dim_v, Number>
        get_transport_direction() const
        {
          dealii::Tensor<1, dim_x + dim_v, Number> a;
          for (int d = 0; d < dim_x; ++d)
            a[d] = 1.0;
          for (int d = 0; d < dim_v; ++d)
            a[d + dim_v] = 6.0;
          return a;
        }
      };
    } // namespace hyperrectangle

    template <int dim_x, int dim_v, int degree, int n_points, typename Number>
    class Application
    {
    public:
      static const int dim = dim_x + dim_v;
      using VectorType     = DeviceVector<Number>;
      using VelocityField  = advection::PhaseSpaceVelocityFieldView<dim_x, dim_v, Number>;

      Application(const DeviceCommunicator &comm, DynamicConvergenceTable &table)
        : comm(comm)
        , table(table)
      {}

      void
      reinit(const JsonParameters &prm)
      {
        const char *            xyz[3] = {"X", "Y", "Z"};
        CartesianLattice<dim_x> lx;
        CartesianLattice<dim_v> lv;
        for (int d = 0; d < dim_x; ++d)
          {
            lx.left[d]    = 0.0; // cases/hyperrectangle.h:150-165
            lx.right[d]   = 4.0 * 3.14159265358979323846264338327950288;
            lx.n_cells[d] = prm.get_int(std::string("Case/NSubdivisionsX/") + xyz[d], 4) << prm.get_int("Case/NRefinementsX", 0);
          }
        for (int d = 0; d < dim_v; ++d)
          {
            lv.left[d]    = -6.0;
            lv.right[d]   = 6.0;
            lv.n_cells[d] = prm.get_int(std::string("Case/NSubdivisionsV/") + xyz[d], 4) << prm.get_int("Case/NRefinementsV", 0);
          }
        lx.periodic = prm.get_bool("Case/PeriodicX", true);
        lv.periodic = prm.get_bool("Case/PeriodicV", true);
        lx.degree = lv.degree = degree;
        lx.n_points = lv.n_points = n_points;
        lx.collocation = lv.collocation = prm.get_bool("SpatialDiscretization/DoCollocation", false);
        matrix_free.reset(new MatrixFree<dim_x, dim_v, Number>(comm, lx, lv));
        matrix_free->reinit();
        matrix_free->initialize_dof_vector(vct_Ki, 0, false, true);
        matrix_free->initialize_dof_vector(vct_Ti, 0, true, true);
        matrix_free->initialize_dof_vector(vct_solution, 0, false, true);
        matrix_free->initialize_dof_vector_x(particle_density);

        std::shared_ptr<dealii::Function<dim, Number>> initial(new hyperrectangle::ExactSolution<dim_x, dim_v, Number>());
        VectorTools::interpolate<degree, degree + 1>(initial, *matrix_free, vct_solution, 0, 0, 2, 2);

        negative_electric_field.reset(new DerivativeContainer<dim_x, dim_v, Number>(*matrix_free));
        poisson_solver.reset(new PoissonSolver<dim_x, dim_v, Number>(*matrix_free));
        boundary_descriptor.reset(new advection::BoundaryDescriptor<dim, Number>());
        velocity_field = std::make_shared<VelocityField>(*matrix_free, *negative_electric_field);
        advection_operation.reset(new advection::AdvectionOperation<dim_x, dim_v, degree, n_points, Number, VectorType, VelocityField>(*matrix_free, table));
        advection::AdvectionOperationParamters op_param;
        op_param.factor_skew = prm.get_double("AdvectionOperation/SkewFactor", 0.0);
        advection_operation->reinit(boundary_descriptor, velocity_field, op_param);

        // application.h:421-446
        tl.time_step            = prm.get_double("TemporalDiscretization/TimeStep", 0.1);
        tl.start_time           = prm.get_double("TemporalDiscretization/StartTime", 0.0);
        tl.final_time           = prm.get_double("TemporalDiscretization/FinalTime", 20.0);
        tl.max_time_step_number = prm.get_int("TemporalDiscretization/MaxTimeStepNumber", 100000000);
        rk_type                 = prm.get("TemporalDiscretization/RKType", "rk45");
        dignostics_tick         = prm.get_double("TemporalDiscretization/DiagnosticsTick", 0.1);
        diag_file               = prm.get("TemporalDiscretization/DiagnosticsFileName", "time_history_diagnostic.out");
        const Number cfl        = prm.get_double("TemporalDiscretization/CFLNumber", 0.3);
        const Number critical_time_step =
          advection::compute_critical_time_step(*matrix_free, hyperrectangle::ExactSolution<dim_x, dim_v, Number>().get_transport_direction());
        const Number dt = std::min<Number>(tl.time_step, cfl * critical_time_step / std::pow(degree, 1.5));
        tl.time_step    = (tl.final_time - tl.start_time) / std::ceil((tl.final_time - tl.start_time) / dt);
        time_loop.reinit(tl);
      }

      void
      solve()
      {
        LowStorageRungeKuttaIntegrator<Number, VectorType> time_integrator(vct_Ki, vct_Ti, rk_type, true);
        bool                                               clear_diag_file = true;
        unsigned int                                       n_total_poisson_iterations = 0;
        // HD_DRIVER_UNFUSED=1: the reference's call structure (std::function right-hand side, stage update as its own kernel)
        const bool unfused = std::getenv("HD_DRIVER_UNFUSED") != nullptr;

        const unsigned int time_steps = time_loop.loop(
          vct_solution,
          [&](auto &solution, const auto cur_time, const auto time_step, const auto &runnable) {
            if (unfused)
              time_integrator.perform_time_step(solution, cur_time, time_step, runnable);
            else
              // per stage: refresh the field from the stage vector, then one kernel for operator + stage update
              time_integrator.perform_time_step(solution, cur_time, time_step, *advection_operation, [&](const VectorType &src, const Number) {
                VectorTools::velocity_space_integration<degree, n_points>(*matrix_free, particle_density, src, 0, 0, 2);
                n_total_poisson_iterations += poisson_solver->solve(*negative_electric_field, particle_density);
              });
          },
          [&](const VectorType &src, VectorType &dst, const Number cur_time) {
            // steps 1-5 of application.h:516-600
            VectorTools::velocity_space_integration<degree, n_points>(*matrix_free, particle_density, src, 0, 0, 2);
            n_total_poisson_iterations += poisson_solver->solve(*negative_electric_field, particle_density);
            advection_operation->apply(dst, src, cur_time);
          },
          [&](const Number cur_time) {
            if (cur_time != tl.start_time && static_cast<int>((cur_time + 0.00000000001 - tl.start_time) / dignostics_tick) ==
                                               static_cast<int>((cur_time + 0.00000000001 - tl.start_time - tl.time_step) / dignostics_tick))
              return;
            // the field energy is the one of the last right-hand side (application.h:619-622: "[TODO] recompute potential!")
            const auto en       = compute_electric_energy(*matrix_free, *negative_electric_field);
            const auto diag_val = phase_space_diagnostics(*matrix_free, vct_solution);
            std::printf("   Time:%10.3e \n", cur_time);
            std::ofstream out(diag_file, clear_diag_file ? std::ios::trunc : std::ios::app);
            if (clear_diag_file)
              out << "# time    energy components  mass  l2norm  kinetic energy  momentum " << std::endl;
            out << std::setw(8) << std::fixed << std::setprecision(3) << cur_time << "  " << std::setprecision(16) << std::scientific;
            for (int d = 0; d < dim_x; ++d)
              out << en[d] << " ";
            out << diag_val[0] << " " << diag_val[1] << " " << diag_val[2] << " " << diag_val[3] << " " << diag_val[4] << " " << diag_val[5] << std::endl;
            clear_diag_file = false;
          });
        table.set("info->time_steps", time_steps);
        table.set("info->n_dofs", double(matrix_free->n_dofs()));
        table.set("info->poisson_iterations", n_total_poisson_iterations);
      }

    private:
      const DeviceCommunicator &                                                                                    comm;
      DynamicConvergenceTable &                                                                                     table;
      std::unique_ptr<MatrixFree<dim_x, dim_v, Number>>                                                             matrix_free;
      VectorType                                                                                                    vct_Ki, vct_Ti, vct_solution, particle_density;
      std::unique_ptr<DerivativeContainer<dim_x, dim_v, Number>>                                                    negative_electric_field;
      std::unique_ptr<PoissonSolver<dim_x, dim_v, Number>>                                                          poisson_solver;
      std::shared_ptr<advection::BoundaryDescriptor<dim, Number>>                                                   boundary_descriptor;
      std::shared_ptr<VelocityField>                                                                                velocity_field;
      std::unique_ptr<advection::AdvectionOperation<dim_x, dim_v, degree, n_points, Number, VectorType, VelocityField>> advection_operation;
      TimeLoopParamters<Number>                                                                                     tl;
      TimeLoop<Number, VectorType>                                                                                  time_loop;
      std::string                                                                                                   rk_type, diag_file;
      Number                                                                                                        dignostics_tick = 0.1;
    };
  } // namespace vp
} // namespace hyperdeal

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
      hyperdeal::JsonParameters prm(argv[1]);
      int                       dim_x = prm.get_int("General/DimX", 0), dim_v = prm.get_int("General/DimV", 0), degree = prm.get_int("General/DegreeX", 0), n_points = 0;
      {
        const std::string file_name = argv[1];
        const auto        slash     = file_name.find_last_of('/');
        const auto        dot       = file_name.find('.', slash == std::string::npos ? 0 : slash);
        std::ifstream     in(file_name.substr(0, dot) + ".configuration");
        if (in)
          {
            std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
            const auto  p = text.find("N_POINTS=");
            if (p != std::string::npos)
              n_points = std::atoi(text.c_str() + p + 9);
          }
      }
      if (n_points == 0)
        n_points = degree + 1;
      hyperdeal::DeviceCommunicator      comm(std::getenv("HD_DEVICE") ? std::atoi(std::getenv("HD_DEVICE")) : 0);
      hyperdeal::DynamicConvergenceTable table;
      std::cout << std::string(argv[1]) << std::endl;
#define HD_CASE(DX, K, Q)                                                                   \
  if (dim_x == DX && dim_v == DX && degree == K && n_points == Q)                           \
    {                                                                                       \
      hyperdeal::vp::Application<DX, DX, K, Q, double> app(comm, table);                    \
      app.reinit(prm);                                                                      \
      app.solve();                                                                          \
    }                                                                                       \
  else
      HD_CASE(1, 3, 4)
      HD_CASE(2, 3, 4)
      HD_CASE(3, 3, 4)
      HD_CASE(2, 2, 3)
      throw hyperdeal::ExcNotImplemented("DIM_X=" + std::to_string(dim_x) + " DIM_V=" + std::to_string(dim_v) + " DEGREE=" + std::to_string(degree));
#undef HD_CASE
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
