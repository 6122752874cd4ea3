// performance/operators_advection_01 (and _02's Cartesian variant) re-hosted on the B200 library: same parameter
// file keys (performance/operators_advection_01/node_level_basic.json), same protocol (warm-up applies, then timed
// applies of dst = M^-1 A(src, 0), performance/operators_advection_01.likwid.cc:205-239) and the same reported
// quantity, "throughput [GDoFs/s]" = n_dofs * n_iterations / t (:341-348).  Timing is by CUDA events on the stream.
//
//   operators_advection <file.json> [--float]        (General.PartitionX x PartitionV > 1, or HD_PARTITION_X / _V: that many GPUs)
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "hyperdeal_b200.hpp"
#include "json_parameters.hpp"

namespace
{
  template <int dim_x, int dim_v, int degree, typename Number>
  void
  run(const hyperdeal::JsonParameters &prm, hyperdeal::DynamicConvergenceTable &table)
  {
    using namespace hyperdeal;
    constexpr int      dim      = dim_x + dim_v;
    constexpr int      n_points = degree + 1;
    DeviceCommunicator comm(std::getenv("HD_DEVICE") ? std::atoi(std::getenv("HD_DEVICE")) : 0);

    CartesianLattice<dim_x> lx;
    CartesianLattice<dim_v> lv;
    const char *            xyz[3] = {"X", "Y", "Z"};
    for (int d = 0; d < dim_x; ++d)
      {
        lx.left[d]    = 0.0;
        lx.right[d]   = 1.0;
        lx.n_cells[d] = prm.get_int(std::string("Case/NSubdivisionsX/") + xyz[d], 1) << prm.get_int("Case/NRefinementsX", 0);
      }
    for (int d = 0; d < dim_v; ++d)
      {
        lv.left[d]    = 0.0;
        lv.right[d]   = 1.0;
        lv.n_cells[d] = prm.get_int(std::string("Case/NSubdivisionsV/") + xyz[d], 1) << prm.get_int("Case/NRefinementsV", 0);
      }
    lx.degree = lv.degree = degree;
    lx.n_points = lv.n_points = n_points;
    lx.collocation = lv.collocation = prm.get_bool("SpatialDiscretization/DoCollocation", false);

    dealii::Tensor<1, dim, Number> a; // the reference benchmark uses a = 0 (:188-189); HD_BENCH_VELOCITY=1 switches all six face terms on
    if (std::getenv("HD_BENCH_VELOCITY"))
      {
        const double v[6] = {1.0, 0.15, -0.05, 0.1, -0.15, 0.5};
        for (int d = 0; d < dim; ++d)
          a[d] = v[d];
      }
    advection::AdvectionOperationParamters op_param;
    op_param.factor_skew                   = prm.get_double("AdvectionOperation/SkewFactor", 0.0);
    const unsigned int n_iterations_warmup = prm.get_int("Performance/IterationsWarmup", 5);
    const unsigned int n_iterations        = prm.get_int("Performance/Iterations", 10);

    // the reference's process grid (performance/util/driver.h:133-161, keys General.PartitionX / PartitionV; the environment
    // overrides the file like `mpirun -np N` does): more than one rank = that many GPUs of this process (hyperdeal::multi)
    int partition_x = prm.get_int("General/PartitionX", 1), partition_v = prm.get_int("General/PartitionV", 1);
    if (const char *e = std::getenv("HD_PARTITION_X"))
      partition_x = std::atoi(e);
    if (const char *e = std::getenv("HD_PARTITION_V"))
      partition_v = std::atoi(e);
    if (partition_x * partition_v > 1)
      {
        multi::MatrixFree<dim_x, dim_v, Number> matrix_free(lx, lv, partition_x, partition_v);
        matrix_free.reinit();
        multi::AdvectionOperation<dim_x, dim_v, Number> advection_operation(matrix_free);
        auto boundary_descriptor = std::make_shared<advection::BoundaryDescriptor<dim, Number>>();
        advection_operation.reinit(boundary_descriptor, a, op_param);
        multi::DistributedDeviceVector<Number> vec_src, vec_dst;
        matrix_free.initialize_dof_vector(vec_src);
        matrix_free.initialize_dof_vector(vec_dst);
        for (unsigned int i = 0; i < n_iterations_warmup; i++)
          advection_operation.apply(vec_dst, vec_src, 0.0);
        HD_CALL(hd_multi_synchronize(matrix_free.get_multi()));
        // every brick has its own stream: wall clock around a synchronised batch
        const auto t0 = std::chrono::steady_clock::now();
        for (unsigned int i = 0; i < n_iterations; i++)
          advection_operation.apply(vec_dst, vec_src, 0.0);
        HD_CALL(hd_multi_synchronize(matrix_free.get_multi()));
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        table.set("info->size [DoFs]", double(matrix_free.n_dofs()));
        table.set("info->dim_x", dim_x);
        table.set("info->dim_v", dim_v);
        table.set("info->degree", degree);
        table.set("info->procs", matrix_free.n_bricks());
        table.set("apply:total [ms]", ms);
        table.set("throughput [GDoFs/s]", double(matrix_free.n_dofs()) * n_iterations / (ms * 1e3) / 1000);
        std::printf("kernel: %s\n", advection_operation.kernel_name().c_str());
        return;
      }

    MatrixFree<dim_x, dim_v, Number>                          matrix_free(comm, lx, lv);
    typename MatrixFree<dim_x, dim_v, Number>::AdditionalData ad;
    ad.do_buffering = prm.get_bool("MatrixFree/DoBuffering", false);
    ad.use_ecl      = prm.get_bool("MatrixFree/UseECL", true);
    matrix_free.reinit(ad);

    using VectorType    = DeviceVector<Number>;
    using VelocityField = advection::ConstantVelocityFieldView<dim, Number>;
    advection::AdvectionOperation<dim_x, dim_v, degree, n_points, Number, VectorType, VelocityField> advection_operation(matrix_free, table);

    auto boundary_descriptor = std::make_shared<advection::BoundaryDescriptor<dim, Number>>();
    auto velocity_field      = std::make_shared<VelocityField>(a);
    advection_operation.reinit(boundary_descriptor, velocity_field, op_param);

    VectorType vec_src, vec_dst;
    matrix_free.initialize_dof_vector(vec_src, 0, true, true);
    matrix_free.initialize_dof_vector(vec_dst, 0, !ad.use_ecl, true);

    for (unsigned int i = 0; i < n_iterations_warmup; i++)
      advection_operation.apply(vec_dst, vec_src, 0.0);
    double ms = 0.0;
    HD_CALL(hd_timer_start(comm.context()));
    for (unsigned int i = 0; i < n_iterations; i++)
      advection_operation.apply(vec_dst, vec_src, 0.0);
    HD_CALL(hd_timer_stop(comm.context(), &ms));

    table.set("info->size [DoFs]", double(matrix_free.n_dofs()));
    table.set("info->dim_x", dim_x);
    table.set("info->dim_v", dim_v);
    table.set("info->degree", degree);
    table.set("info->procs", 1);
    table.set("apply:total [ms]", ms);
    table.set("throughput [GDoFs/s]", double(matrix_free.n_dofs()) * n_iterations / (ms * 1e3) / 1000);
    std::printf("kernel: %s\n", advection_operation.kernel_name());
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
      const bool                         use_float = argc > 2 && std::string(argv[2]) == "--float";
      hyperdeal::JsonParameters          prm(argv[1]);
      hyperdeal::DynamicConvergenceTable table;
      const int                          dim = prm.get_int("General/Dim", 6), degree = prm.get_int("General/Degree", 3);
#define HD_CASE(DX, DV, K)                                          \
  if (dim == DX + DV && degree == K)                                \
    {                                                               \
      if (use_float)                                                \
        run<DX, DV, K, float>(prm, table);                          \
      else                                                          \
        run<DX, DV, K, double>(prm, table);                         \
    }                                                               \
  else
      HD_CASE(1, 1, 3)
      HD_CASE(2, 2, 3)
      HD_CASE(3, 3, 3)
      HD_CASE(2, 2, 5)
      HD_CASE(3, 3, 5)
      HD_CASE(3, 3, 2)
      HD_CASE(3, 3, 4)
      throw hyperdeal::ExcNotImplemented("Dim=" + std::to_string(dim) + " Degree=" + std::to_string(degree));
#undef HD_CASE
      table.print();
    }
  catch (std::exception &exc)
    {
      std::cerr << "Exception on processing: " << std::endl << exc.what() << std::endl << "Aborting!" << std::endl;
      return 1;
    }
  return 0;
}
