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


def test_libm_flavours_of_the_hoisted_calls():
    """EGNO: pow(x, 3), pow(x, 3/2), pow(x, -1/2) and the symbolic exponent -3*alpha, all per row.
    Default flavour: the reference host's libm bit for bit (inflx_gl_pow, literal exponents
    included: glibc's pow is not always the correctly rounded value the dd chains return);
    "cr": correctly rounded (dd chains for the literal powers); "device": libdevice."""
    unit = cexpr.parse_c_unit(oracle.golden_c_text("egno"))
    src = cudagen.ModelProgram(unit).groups["cmp"].cuda_source("egno")
    gen = src[src.index("// ===== generated"):]
    assert "inflx_gl_pow(" in gen and "inflx_cr_" not in src
    rows = gen[gen.index("void inflx_rows("):gen.index("inflx_slow_roots")]
    assert "inflx_powi<" not in rows and "inflx_powh" not in rows  # no dd chain in a hoisted class
    src = cudagen.ModelProgram(unit, libm="cr").groups["cmp"].cuda_source("egno")
    assert "inflx_powi<3>(" in src and "inflx_powh<1>(" in src and "inflx_powh_neg<0>(" in src
    assert "inflx_cr_pow(" in src and "inflx_gl_pow" not in src and "INFLX_EXACT_ATAN_TAN 1" not in src
    src = cudagen.ModelProgram(unit, libm="device").groups["cmp"].cuda_source("egno")
    assert "inflx_cr_" not in src and "inflx_gl_pow" not in src and " pow(" in src
    # "glibc-all": per-point powers and the epilogue's atan / tan through the restated glibc too
    src = cudagen.ModelProgram(unit, libm="glibc-all").groups["cmp"].cuda_source("egno")
    assert src.startswith("#define INFLX_GROUP_MIN_BLOCKS") and "#define INFLX_EXACT_ATAN_TAN 1" in src
    assert "double inflx_gl_atan(double x) {" in src and "double inflx_gl_tan(double x) {" in src
    with pytest.raises(Exception, match="unknown libm flavour"):
        cudagen.ModelProgram(unit, libm="musl")


def test_hoisted_libm_calls_use_the_reference_libm_and_expensive_columns_get_a_prepass():
    """log / exp / pow / sin / cos / tanh of classes P, R, C go through csrc/inflx_glibcmath.cuh; a
    column block that holds one is evaluated by the `inflx_cols` pre-pass (d5: sin, cos; angular:
    pow(x1, n)), a pure-arithmetic column block stays in the grid kernel's prologue (EGNO)."""
    import re

    progs = {
        m: cudagen.ModelProgram(cexpr.parse_c_unit(oracle.golden_c_text(m)))
        for m in ("hyper", "angular", "egno", "d5")
    }
    for m, prog in progs.items():
        for g, gp in prog.groups.items():
            col_calls = [
                i for i in gp.grid_nodes
                if gp.node(i)[0] == "f" and gp.klass(i) == "C" and gp._hoisted_libm(i)
            ]
            assert gp.cols_prepass == bool(col_calls), (m, g)
            src = gp.cuda_source(m)
            assert ("void inflx_cols(" in src) == gp.cols_prepass
            assert f"#define INFLX_NCF {len(gp.c_frontier)}\n" in src
            for i in gp.all_nodes:
                n = gp.node(i)
                if n[0] == "f" and n[1] in cudagen.GL_FUNCTIONS:
                    if gp.klass(i) == "M":
                        # the only per-point calls of the test models: pow(., 3/2), pow(., -1/2)
                        # (EGNO, d5) - dd chains by default, inflx_gl_pow_m in flavour "glibc-all"
                        assert n[1] == "pow" and gp.dag.cval(n[3]) in (1.5, -0.5), (m, g, n)
                        assert gp._hoisted_libm(i) is None
                    else:
                        assert gp._hoisted_libm(i) == "inflx_gl_"
            # the generated part (after the headers) spells no plain libm call of that set
            gen = src[src.index("// ===== generated"):]
            assert not re.search(r"(?<![A-Za-z_])(log|exp|sin|cos|tanh|pow)\(", gen), (m, g)
            if m == "egno" and g == "cmp":
                all_src = cudagen.ModelProgram(prog.unit, libm="glibc-all").groups[g].cuda_source(m)
                loop = all_src[all_src.index("#pragma unroll 1"):]
                assert "inflx_gl_pow_m(" in loop[: loop.index("inflx_grid_complete_analysis_sweep")]
    assert progs["d5"].groups["cmp"].cols_prepass and not progs["egno"].groups["cmp"].cols_prepass
    gp = progs["d5"].groups["cmp"]
    src = gp.cuda_source("d5")
    cols = src[src.index("void inflx_cols("):src.index("inflx_slow_roots")]
    assert "inflx_gl_cos(x1)" in cols and "inflx_gl_sin(x1)" in cols
    grid = src[src.index("void __launch_bounds__(INFLX_BLOCK, INFLX_MIN_BLOCKS) inflx_grid_complete_analysis("):]
    grid = grid[: grid.index("#pragma unroll 1")]
    assert grid.count("__ldg(cc + ") == len(gp.c_frontier) and "inflx_gl_" not in grid
    # frontier completeness for the column pre-pass: M nodes read C values through a slot
    for i in gp.nodes_of("M", gp.grid_nodes):
        for o in gp.operands(i):
            if gp.klass(o) == "C" and gp.is_op(o):
                assert o in gp.c_slot
    assert "inflx_gl_tanh(" in progs["hyper"].groups["cmp"].cuda_source("hyper")
