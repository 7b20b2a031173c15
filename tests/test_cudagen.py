"""Rate partition of the model DAG (cudagen): every per-point value is computed exactly once, at
the slowest rate it changes; frontiers are complete; the emitted CUDA compiles for sm_100a."""
import pytest

import cases
import oracle
from inflatox_b200 import cexpr, cudagen
from inflatox_b200.compiler import nvrtc_compile


@pytest.fixture(scope="module", params=["hyper", "angular", "egno", "d5"])
def program(request):
    unit = cexpr.parse_c_unit(oracle.golden_c_text(request.param))
    return request.param, cudagen.ModelProgram(unit)


def test_partition_is_a_partition(program):
    model, prog = program
    for g, gp in prog.groups.items():
        ops = [i for i in gp.grid_nodes if gp.is_op(i)]
        by_class = {c: gp.nodes_of(c, gp.grid_nodes) for c in "PRCM"}
        assert sorted(sum(by_class.values(), [])) == sorted(ops)
        for c, nodes in by_class.items():
            for i in nodes:
                deps = {gp.klass(o) for o in gp.operands(i)} - {"K"}
                allowed = {"P": {"P"}, "R": {"P", "R"}, "C": {"P", "C"}, "M": {"P", "R", "C", "M"}}[c]
                assert deps <= allowed, (model, g, i, c, deps)
        # frontier completeness: a faster class only reads slower-class values through a slot
        for i in by_class["M"] + by_class["C"]:
            for o in gp.operands(i):
                if gp.klass(o) == "P":
                    assert o in gp.p_slot
                if gp.klass(o) == "R":
                    assert o in gp.r_slot
        assert len(gp.p_frontier) <= cudagen.PC_CAPACITY


def test_hoisting_removes_most_per_point_work(program):
    model, prog = program
    st = prog.groups["cmp"].stats()
    if model in ("egno", "d5"):
        assert st["flops_per_point_executed"] < 0.4 * st["flops_model"]
        # most per-point quotients reuse a hoisted reciprocal
        assert st["reciprocals_per_point"] < st["divisions_per_point"]
    if model == "hyper":  # no mixed node at all: the model is a function of x[0] only
        assert st["class_ops"]["M"] == 0


def test_emitted_cuda_compiles_for_sm_100a(program):
    model, prog = program
    src = prog.groups["cmp"].cuda_source(model)
    assert "inflx_grid_complete_analysis_sweep" in src and "inflx_slow_roots" in src
    cubin = nvrtc_compile(src, f"{model}_cmp.cu", ["--gpu-architecture=sm_100a", "--std=c++17",
                                                   "--fmad=false", "-lineinfo"])  # fmt: skip
    assert cubin[:4] == b"\x7fELF"


def test_literal_powers_use_the_double_double_chains():
    unit = cexpr.parse_c_unit(oracle.golden_c_text("egno"))
    src = cudagen.ModelProgram(unit).groups["cmp"].cuda_source("egno")
    assert "inflx_powi<3>(" in src and "inflx_powh<1>(" in src and "inflx_powh_neg<0>(" in src
    assert "pow(" in src  # the symbolic exponent -3*alpha stays a libdevice pow
