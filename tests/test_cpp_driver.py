"""The C++ host layer (hyperdeal_b200/cpp/): the re-hosted examples/advection driver must reproduce the
reference's golden files (examples/advection/tests/*.out) when run on the GPU through the C ABI, and must fail
loudly (exit code 1, like the reference's main) without one."""
import os
import subprocess

import pytest

from conftest import has_gpu

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def drivers():
    from hyperdeal_b200 import build, build_cpp

    build.build()
    return {os.path.basename(p): p for p in build_cpp.build()}


def _parse_stdout(text):
    out = []
    for line in text.splitlines():
        line = line.strip()
        if line.startswith("Time:"):
            parts = line.replace(",", " ").split()
            out.append((float(parts[0].split(":")[1]) if parts[0] != "Time:" else float(parts[1]), float(parts[-3]), float(parts[-1])))
    return out


def test_drivers_build(drivers):
    assert set(drivers) == {"advection", "operators_advection", "vlasov_poisson"}
    for p in drivers.values():
        assert os.access(p, os.X_OK)


def test_driver_without_arguments_returns_1(drivers):
    r = subprocess.run([drivers["advection"]], capture_output=True, text=True)
    assert r.returncode == 1 and "No .json parameter files" in r.stdout


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_driver_fails_loudly_without_gpu(drivers, golden_dir):
    r = subprocess.run([drivers["advection"], os.path.join(golden_dir, "adv_2D_2D_k3.hyperrectangle_01.json")], capture_output=True, text=True)
    assert r.returncode == 1
    assert "Exception on processing" in r.stderr and "hd_context_create" in r.stderr


def test_shim_header_is_self_contained(tmp_path):
    """the shim compiles on its own and its names match the reference API surface"""
    src = tmp_path / "t.cc"
    src.write_text(
        '#include "hyperdeal_b200.hpp"\n'
        "using namespace hyperdeal;\n"
        "using VT = DeviceVector<double>;\n"
        "using VF = advection::ConstantVelocityFieldView<4, double>;\n"
        "template class MatrixFree<2, 2, double>;\n"
        "template class advection::AdvectionOperation<2, 2, 3, 4, double, VT, VF>;\n"
        "template class LowStorageRungeKuttaIntegrator<double, VT>;\n"
        "template class TimeLoop<double, VT>;\n"
        "static_assert(sizeof(MatrixFree<2, 2, double>::AdditionalData) > 0, \"\");\n"
        "int main() { advection::AdvectionOperationParamters p; dealii::Tensor<1, 4, double> a; (void)a; return p.factor_skew == 0.0 ? 0 : 1; }\n")
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "hyperdeal_b200", "cpp"), str(src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


# ..._04, _08 and _q5..._02 run the reference with UseECL = false / DoBuffering = true (its face-centric loop,
# advection_operation.h:574-1130) — served by the same ECL-style kernels here, and their .out files equal the ECL twins'.
GOLDEN = ["adv_2D_2D_k3.hyperrectangle_01", "adv_2D_2D_k3.hyperrectangle_03", "adv_2D_2D_k3.hyperrectangle_05", "adv_2D_2D_k3.hyperrectangle_07",
          "adv_2D_2D_k3_q5.hyperrectangle_01", "adv_1D_1D_k3.hyperrectangle_01_rk47",
          "adv_2D_2D_k3.hyperrectangle_02", "adv_2D_2D_k3.hyperrectangle_04", "adv_2D_2D_k3.hyperrectangle_08", "adv_2D_2D_k3_q5.hyperrectangle_02"]


def _check(lines, gold, name):
    assert len(lines) == len(gold), name
    for (t1, n1, e1), (t2, n2, e2) in zip(lines, gold):
        assert abs(t1 - t2) <= 1e-3 * max(abs(t2), 1e-3)
        # both sides are printed with 11 significant digits (numdiff -r 1e-8 in the reference)
        assert abs(n1 - n2) <= 3e-10 * abs(n2), (name, t1, n1, n2)
        if e2 > 1e-12:
            assert abs(e1 - e2) <= 3e-10 * abs(e2), (name, t1, e1, e2)
        else:
            assert e1 < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN)
def test_cpp_advection_driver_reproduces_golden(drivers, golden_dir, name):
    r = subprocess.run([drivers["advection"], os.path.join(golden_dir, name + ".json")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    _check(_parse_stdout(r.stdout), O.parse_golden(os.path.join(golden_dir, name + ".out")), name)


def _n_gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


# the reference's own multi-rank tests are the same parameter files run under mpirun -np N with the same expected output
# (examples/advection/tests/*.mpirun=N.out style); here PartitionX x PartitionV GPUs of one process (hd_multi_*)
@pytest.mark.gpu
@pytest.mark.parametrize("px,pv", [(2, 1), (1, 2), (2, 2), (4, 1), (4, 2)])
@pytest.mark.parametrize("name", ["adv_2D_2D_k3.hyperrectangle_01", "adv_2D_2D_k3.hyperrectangle_03", "adv_2D_2D_k3_q5.hyperrectangle_01", "adv_1D_1D_k3.hyperrectangle_01_rk47"])
def test_cpp_advection_driver_multi_gpu_golden(drivers, golden_dir, name, px, pv):
    if _n_gpus() < px * pv:
        pytest.skip("needs %d GPUs" % (px * pv))
    env = dict(os.environ, HD_PARTITION_X=str(px), HD_PARTITION_V=str(pv))
    r = subprocess.run([drivers["advection"], os.path.join(golden_dir, name + ".json")], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr
    assert "bricks: %d" % (px * pv) in r.stdout
    _check(_parse_stdout(r.stdout), O.parse_golden(os.path.join(golden_dir, name + ".out")), name)


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_multi_gpu_driver_fails_loudly_without_gpu(drivers, golden_dir):
    r = subprocess.run([drivers["advection"], os.path.join(golden_dir, "adv_2D_2D_k3.hyperrectangle_01.json")], capture_output=True, text=True,
                       env=dict(os.environ, HD_PARTITION_X="2"))
    assert r.returncode == 1 and "Exception on processing" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"HD_DRIVER_UNFUSED": "1"}, {"HD_DRIVER_HOST_FUNCTIONS": "1"}, {"HD_DRIVER_UNFUSED": "1", "HD_DRIVER_HOST_FUNCTIONS": "1"}])
def test_cpp_driver_reference_call_structure(drivers, golden_dir, env):
    """std::function integrator path (operator and stage update as separate kernels) and host-side dealii::Function
    objects for initial/boundary data: the Dirichlet golden must come out the same"""
    name = "adv_2D_2D_k3.hyperrectangle_03"
    r = subprocess.run([drivers["advection"], os.path.join(golden_dir, name + ".json")], capture_output=True, text=True, timeout=600, env=dict(os.environ, **env))
    assert r.returncode == 0, r.stderr
    _check(_parse_stdout(r.stdout), O.parse_golden(os.path.join(golden_dir, name + ".out")), name)


@pytest.mark.gpu
def test_cpp_operators_advection_driver(drivers, golden_dir):
    r = subprocess.run([drivers["operators_advection"], os.path.join(golden_dir, "operators_advection_small.json")], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, HD_BENCH_VELOCITY="1"))
    assert r.returncode == 0, r.stderr
    assert "kernel: advect_3d3v_k3" in r.stdout or "kernel: rounds_3d3v_k3" in r.stdout
    rows = dict(line.rsplit(None, 1) for line in r.stdout.splitlines() if line.startswith(("info", "throughput")))
    assert float(rows["info->size [DoFs]"]) == 2 * 4 * 2 * 2 * 2 * 4 * 4096
    assert float(rows["throughput [GDoFs/s]"]) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("px,pv,velocity", [(2, 1, True), (1, 2, True), (2, 1, False), (2, 2, True)])
def test_cpp_operators_advection_driver_multi_gpu(drivers, golden_dir, px, pv, velocity):
    """the reference's performance driver on PartitionX x PartitionV GPUs of one process (with and without transport velocity:
    the reference benchmark uses a = 0, i.e. no face term reads a neighbour)"""
    if _n_gpus() < px * pv:
        pytest.skip("needs %d GPUs" % (px * pv))
    env = dict(os.environ, HD_PARTITION_X=str(px), HD_PARTITION_V=str(pv))
    if velocity:
        env["HD_BENCH_VELOCITY"] = "1"
    r = subprocess.run([drivers["operators_advection"], os.path.join(golden_dir, "operators_advection_small.json")], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr
    rows = dict(line.rsplit(None, 1) for line in r.stdout.splitlines() if line.startswith(("info", "throughput")))
    assert float(rows["info->size [DoFs]"]) == 2 * 4 * 2 * 2 * 2 * 4 * 4096
    assert float(rows["info->procs"]) == px * pv
    assert float(rows["throughput [GDoFs/s]"]) > 0


def test_json_parameter_reader(tmp_path, golden_dir):
    """the ParameterHandler-style JSON reader of the drivers: nested sections, quoted and bare values, booleans"""
    src = tmp_path / "j.cc"
    src.write_text(
        '#include "json_parameters.hpp"\n#include <cstdio>\n'
        "int main(int argc, char **argv) {\n"
        "  hyperdeal::JsonParameters a(argv[1]), b(argv[2]);\n"
        '  std::printf("%ld %g %d %s %ld|", a.get_int("General/DimX", 0), a.get_double("TemporalDiscretization/CFLNumber", 0), (int)a.get_bool("Case/PeriodicX", true),\n'
        '              a.get("TemporalDiscretization/RKType", "?").c_str(), a.get_int("Case/NSubdivisionsV/Y", 0));\n'
        '  std::printf("%ld %g %d %ld %d\\n", b.get_int("General/Dim", 0), b.get_double("AdvectionOperation/SkewFactor", 0), (int)b.get_bool("MatrixFree/UseECL", false),\n'
        '              b.get_int("Case/NSubdivisionsX/Y", 0), (int)b.has("Nope/Key"));\n'
        "  try { hyperdeal::JsonParameters c(argv[3]); } catch (const std::exception &e) { std::printf(\"error: %s\\n\", e.what()); }\n"
        "  return 0;\n}\n")
    bad = tmp_path / "bad.json"
    bad.write_text('{"A": {"B": 1, }')
    exe = tmp_path / "j"
    r = subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "hyperdeal_b200", "cpp"), str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), os.path.join(golden_dir, "adv_2D_2D_k3.hyperrectangle_03.json"), os.path.join(golden_dir, "operators_advection_small.json"), str(bad)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[0] == "2 0.3 0 rk45 4|6 0.5 1 2 0"
    assert lines[1].startswith("error: parameter file:")

