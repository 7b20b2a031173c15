"""The GENERATED CUDA translation units, executed on the host (tests/native/cuda_host_shim.h +
emulate.cpp): parameter block, row / column pre-passes, grid kernels (plain and sweep), smem
staging, frontier slots, slow path - the same source text NVRTC compiles, one emulated thread per
CTA, launched in the engine's order.  This checks the generator (rate partition, frontiers,
hoisted reciprocals, hoisted libm calls with the reference libm's bits, epilogues, stores) against the oracle
without a GPU; the GPU tests then only have to establish that the device executes that text the
way the host does (tests/test_gpu_*.py).

Not emulated exactly: the two hardware seeds (MUFU.RCP64H / RSQ64H) are IEEE values here.  The
refined quotient / root is the correctly rounded one either way for operands the fast path
accepts; correction terms that use a bare seed (half-integer powers, atan for y > 1) may differ
in the last bit in ~1e-7 of the calls.  libm calls outside the restated glibc set (sinh, atan, ...)
and per-point calls resolve to the HOST's libm here and to libdevice on the GPU.  On 512^2 grids
the emulation reproduced the GPU's round-1 parity statistics against the oracle digit for digit
(profiles/parity_r1.json)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import cases
import oracle
from inflatox_b200 import cexpr, cudagen
from raw_units import N_PAR, RawOracle, make_unit

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVE = os.path.join(ROOT, "tests", "native")
_DP = ctypes.POINTER(ctypes.c_double)
_PER_POINT = {"complete_analysis": 6, "hesse": 4}


class Emulated:
    """One generated translation unit + one grid op, compiled for the host."""

    def __init__(self, program, name: str, group: str, op: str, workdir, rpt_max: int = 16):
        cu = os.path.join(str(workdir), f"{name}_{group}.cu")
        if not os.path.exists(cu):
            with open(cu, "w") as fh:
                fh.write(program.groups[group].cuda_source(name))
        so = os.path.join(str(workdir), f"emu_{name}_{group}_{op}.so")
        subprocess.run(
            ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC",
             "-DINFLX_BLOCK=1", f"-DINFLX_RPT={rpt_max}", f'-DINFLX_GENERATED_CU="{cu}"',
             f"-DINFLX_EMU_KERNEL=inflx_grid_{op}", f"-DINFLX_EMU_KERNEL_SWEEP=inflx_grid_{op}_sweep",
             f"-I{NATIVE}", os.path.join(NATIVE, "emulate.cpp"), "-o", so],
            check=True,
        )  # fmt: skip
        self.lib = ctypes.CDLL(so)
        self.lib.emu_grid.argtypes = [
            _DP, ctypes.c_uint, _DP, ctypes.c_ulonglong, ctypes.c_uint, _DP, ctypes.c_ulonglong,
            ctypes.c_ulonglong, ctypes.c_uint, ctypes.c_double, ctypes.c_int, ctypes.c_uint,
            ctypes.c_uint,
        ]  # fmt: skip
        self.op, self.per = op, _PER_POINT.get(op, 1)

    def grid(self, p, n0, n1, ext, rows=None, rpt=16, aux=0.0, fused=True, n_big=None, rpt_tail=0):
        p2 = np.ascontiguousarray(np.atleast_2d(np.asarray(p, dtype=np.float64)))
        s = p2.shape[0]
        r0, r1 = rows if rows is not None else (0, n0)
        if self.op == "hesse":
            out = np.full((4, s, r1 - r0, n1), -7.0)
        else:
            out = np.full((s, r1 - r0, n1, self.per), -7.0)  # every element must be overwritten
        ss = np.ascontiguousarray(ext, dtype=np.float64)
        rc = self.lib.emu_grid(p2.ctypes.data_as(_DP), s, out.ctypes.data_as(_DP), n0, n1,
                               ss.ctypes.data_as(_DP), r0, r1, rpt, aux, int(fused),
                               0xFFFFFFFF if n_big is None else n_big, rpt_tail)  # fmt: skip
        assert rc == 0
        if self.op == "hesse":
            return out[:, 0] if np.ndim(p) == 1 else out
        out = out[..., 0] if self.per == 1 else out
        return out[0] if np.ndim(p) == 1 else out


def _program(model):
    return cudagen.ModelProgram(cexpr.parse_c_unit(oracle.golden_c_text(model)))


def _same_bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b)) | ((a == 0) & (b == 0))


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    return tmp_path_factory.mktemp("emu")


@pytest.mark.parametrize("model", cases.MODELS)
def test_generated_kernels_reproduce_the_oracle(model, workdir):
    """complete_analysis on a ragged 237 x 331 grid: NaN / inf masks of the oracle and EVERY finite
    point within 1e-10 of it (north_star's bar) - the hoisted libm calls return the reference
    host's bits (csrc/inflx_glibcmath.cuh); eps_V and the pure-arithmetic planes are bit-identical
    wherever the model has no per-point libm call."""
    emu = Emulated(_program(model), model, "cmp", "complete_analysis", workdir)
    p, ext, n0, n1 = cases.params(model), cases.EXTENT[model], 237, 331
    got = emu.grid(p, n0, n1, ext)
    assert not (got == -7.0).any()
    ref = oracle.Oracle(model).complete_analysis(p, n0, n1, ext)
    err, fin, nan_mm, inf_mm = cases.rel_err(got, ref)
    assert nan_mm == 0 and inf_mm == 0
    assert (err[fin] <= 1e-10).all(), float(err[fin].max())
    for k in (0, 1, 2, 5):  # consistency, eps_V, eps_H, omega: IEEE arithmetic on top of the model
        same = _same_bits(got[..., k], ref[..., k])
        assert same.mean() >= 0.999, (model, k, float(same.mean()))


@pytest.mark.parametrize("model", cases.MODELS)
def test_flavour_glibc_all_reproduces_every_bit_of_the_oracle(model, workdir):
    """libm="glibc-all": per-point half-integer powers through the restated glibc pow AND the
    epilogue's delta / tan(delta) through the restated glibc atan / tan.  No operation of the path
    then differs from the reference's: all six planes of every model equal the oracle's bit for bit
    (NaNs at the same places)."""
    prog = cudagen.ModelProgram(cexpr.parse_c_unit(oracle.golden_c_text(model)), libm="glibc-all")
    emu = Emulated(prog, model + "_glall", "cmp", "complete_analysis", workdir)
    p, ext, n0, n1 = cases.params(model), cases.EXTENT[model], 237, 331
    got = emu.grid(p, n0, n1, ext)
    ref = oracle.Oracle(model).complete_analysis(p, n0, n1, ext)
    for k, name in enumerate(["consistency", "eps_V", "eps_H", "eta", "delta", "omega"]):
        same = _same_bits(got[..., k], ref[..., k])
        assert same.all(), (model, name, int((~same).sum()))


@pytest.mark.parametrize("model", ["angular", "egno"])
def test_correctly_rounded_flavour_differs_from_the_oracle_only_where_glibc_misrounds(model, workdir):
    """The round-1 flavour (libm="cr") stays available: every finite point within 1e-10 of the
    oracle variant that links the correctly rounded libm."""
    prog = cudagen.ModelProgram(cexpr.parse_c_unit(oracle.golden_c_text(model)), libm="cr")
    emu = Emulated(prog, model + "_cr", "cmp", "complete_analysis", workdir)
    p, ext, n0, n1 = cases.params(model), cases.EXTENT[model], 61, 83
    got = emu.grid(p, n0, n1, ext)
    ref_cr = oracle.Oracle(model, libm="cr").complete_analysis(p, n0, n1, ext)
    err, fin, nan_mm, inf_mm = cases.rel_err(got, ref_cr)
    assert nan_mm == 0 and inf_mm == 0
    assert (err[fin] <= 1e-10).all(), float(err[fin].max())


@pytest.mark.parametrize("model", ["angular", "d5"])  # d5: column pre-pass
def test_launch_geometry_does_not_change_a_bit(model, workdir):
    emu = Emulated(_program(model), model, "cmp", "complete_analysis", workdir)
    p, ext, n0, n1 = cases.params(model), cases.EXTENT[model], 45, 37
    base = emu.grid(p, n0, n1, ext, rpt=16)
    for rpt in (1, 2, 5, 8):  # rows per CTA: the engine's launch argument
        assert _same_bits(emu.grid(p, n0, n1, ext, rpt=rpt), base).all(), rpt
    # two tile heights: n_big full tiles, then short ones (the engine's tail policy)
    for rpt, n_big, rpt_tail in ((16, 1, 4), (16, 2, 2), (8, 0, 3), (8, 5, 1), (16, 0, 16)):
        got = emu.grid(p, n0, n1, ext, rpt=rpt, n_big=n_big, rpt_tail=rpt_tail)
        assert _same_bits(got, base).all(), (rpt, n_big, rpt_tail)
    # a row shard uses global coordinates
    part = emu.grid(p, n0, n1, ext, rows=(13, 40), rpt=4)
    assert _same_bits(part, base[13:40]).all()


@pytest.mark.parametrize("model", cases.MODELS)
def test_fused_prologue_equals_the_three_step_prologue(model, workdir):
    """One parameter vector: `inflx_prologue` (parameters by value, parameter block recomputed by
    every row thread) against inflx_params -> bank -> inflx_rows: not a bit of the output moves."""
    emu = Emulated(_program(model), model, "cmp", "complete_analysis", workdir)
    p, ext, n0, n1 = cases.params(model), cases.EXTENT[model], 150, 41  # > 128 rows: two row CTAs
    a = emu.grid(p, n0, n1, ext, fused=True)
    b = emu.grid(p, n0, n1, ext, fused=False)
    assert _same_bits(a, b).all()
    part = emu.grid(p, n0, n1, ext, rows=(129, 150), fused=True)
    assert _same_bits(part, b[129:150]).all()


@pytest.mark.parametrize("model", ["hyper", "d5"])
def test_sweep_kernel_equals_one_launch_per_vector(model, workdir):
    emu = Emulated(_program(model), model, "cmp", "complete_analysis", workdir)
    rng = np.random.default_rng(3)
    p0 = cases.params(model)
    ps = np.stack([p0, p0 * (1 + 0.1 * rng.random(p0.size)), p0 * (1 - 0.1 * rng.random(p0.size))])
    ext, n0, n1 = cases.EXTENT[model], 33, 29
    fused = emu.grid(ps, n0, n1, ext, rpt=8)
    assert fused.shape == (3, n0, n1, 6)
    for k in range(3):
        assert _same_bits(fused[k], emu.grid(ps[k], n0, n1, ext, rpt=8)).all(), k


def test_single_plane_and_array_ops(workdir):
    prog, model = _program("angular"), "angular"
    p, ext, n0, n1 = cases.params(model), cases.EXTENT[model], 40, 52
    orc = oracle.Oracle(model)
    for group, op, ref in (
        ("con", "consistency_only", orc.consistency_only(p, n0, n1, ext)),
        ("con", "consistency_rapidturn_only", orc.consistency_rapidturn_only(p, n0, n1, ext)),
        ("eps", "epsilon_v_only", orc.epsilon_v_only(p, n0, n1, ext)),
        ("pot", "potential", orc.potential_array(p, n0, n1, ext)),
        ("hes", "hesse", orc.hesse_array(p, n0, n1, ext).reshape(4, n0, n1)),
    ):
        got = Emulated(prog, model, group, op, workdir).grid(p, n0, n1, ext)
        err, fin, nan_mm, inf_mm = cases.rel_err(got, ref)
        assert nan_mm == 0 and inf_mm == 0, op
        assert (err[fin] <= 1e-10).all(), (op, float(err[fin].max()))


@pytest.mark.parametrize("seed,conditionals", [(0, False), (1, False), (2, False), (100, True), (101, True)])
def test_random_arithmetic_models_are_bit_identical_on_the_host(seed, conditionals, tmp_path):
    """The CPU twin of tests/test_gpu_random_models.py: random units made of + - * / sqrt fabs
    pow(., 2) (and comparisons / ?: when `conditionals`), gcc against the generated kernels."""
    c_text = make_unit(seed, conditionals)
    orc = RawOracle(c_text, str(tmp_path))
    prog = cudagen.ModelProgram(cexpr.parse_c_unit(c_text))
    rng = np.random.default_rng(seed)
    p = rng.uniform(0.5, 3.0, N_PAR)
    ext, n0, n1 = (-2.0, 3.0, -1.5, 2.5), 37, 59
    v = Emulated(prog, "rnd", "pot", "potential", tmp_path).grid(p, n0, n1, ext)
    assert _same_bits(v, orc.potential_array(p, n0, n1, ext)).all()
    h = Emulated(prog, "rnd", "hes", "hesse", tmp_path).grid(p, n0, n1, ext)
    assert _same_bits(h, orc.hesse_array(p, n0, n1, ext).reshape(4, n0, n1)).all()
    out = Emulated(prog, "rnd", "cmp", "complete_analysis", tmp_path).grid(p, n0, n1, ext)
    ref = orc.complete_analysis(p, n0, n1, ext)
    for k in (0, 1, 2, 5):  # every plane that is pure IEEE arithmetic
        same = _same_bits(out[..., k], ref[..., k])
        assert same.all(), (seed, k, int((~same).sum()))
    for k in (3, 4):
        assert (np.isnan(out[..., k]) == np.isnan(ref[..., k])).all()


@pytest.mark.parametrize(
    "params",
    [[1.0, 2.0, 3.0], [1e308, 0.0, 1e-320], [np.inf, -0.0, np.nan], [3e-310, 1e300, 5e-324]],
)
def test_irregular_operands_take_the_exact_path_on_the_host(params, tmp_path):
    """CPU twin of the GPU test of the same name: zero, infinite, subnormal, huge and NaN operands
    in every class; the generated validity tests must send those points to the IEEE slow path
    (`inflx_slow_roots`) - results equal gcc's bit for bit, signs of infinities included."""
    from raw_units import SPECIAL_UNIT

    orc = RawOracle(SPECIAL_UNIT, str(tmp_path))
    prog = cudagen.ModelProgram(cexpr.parse_c_unit(SPECIAL_UNIT))
    p = np.array(params, dtype=np.float64)
    ext, n0, n1 = (-1.0, 3.0, -2.0, 2.0), 16, 32  # contains x0 = 0, +-1, x1 = 0, x0 = x1 exactly
    v = Emulated(prog, "special", "pot", "potential", tmp_path).grid(p, n0, n1, ext)
    assert _same_bits(v, orc.potential_array(p, n0, n1, ext)).all()
    h = Emulated(prog, "special", "hes", "hesse", tmp_path).grid(p, n0, n1, ext)
    assert _same_bits(h, orc.hesse_array(p, n0, n1, ext).reshape(4, n0, n1)).all()
    out = Emulated(prog, "special", "cmp", "complete_analysis", tmp_path).grid(p, n0, n1, ext)
    ref = orc.complete_analysis(p, n0, n1, ext)
    for k in (0, 1, 2, 5):
        ok = _same_bits(out[..., k], ref[..., k])
        assert ok.all(), (k, out[..., k][~ok][:4], ref[..., k][~ok][:4])
    assert (np.isnan(out[..., 3:5]) == np.isnan(ref[..., 3:5])).all()
    assert (np.isinf(out[..., 3:5]) == np.isinf(ref[..., 3:5])).all()


@pytest.mark.parametrize(
    "params",
    [[1.0, 2.0, 3.0], [1e308, 0.0, 1e-320], [np.inf, -0.0, np.nan], [3e-310, 1e300, 5e-324]],
)
def test_constant_zero_v10_takes_the_special_epilogue_bit_for_bit(params, tmp_path):
    """v10 == 0 (diagonal projected Hesse matrix): the generator emits inflx_op_complete_v10z_s,
    which never divides by the zero.  Ordinary and irregular operands (zero, inf, NaN, subnormal,
    huge in v, v00, v11, |grad V|^2): every plane that is IEEE arithmetic equals gcc's bits,
    consistency is NaN everywhere, delta is +0 or NaN exactly where gcc says."""
    from raw_units import ZERO_V10_UNIT

    orc = RawOracle(ZERO_V10_UNIT, str(tmp_path))
    prog = cudagen.ModelProgram(cexpr.parse_c_unit(ZERO_V10_UNIT))
    src = prog.groups["cmp"].cuda_source("zero_v10")
    grid = src[src.index("void __launch_bounds__(INFLX_BLOCK, INFLX_MIN_BLOCKS) inflx_grid_complete_analysis("):]
    assert "inflx_op_complete_v10z_s(" in grid[: grid.index("if (bad.any())")]
    p = np.array(params, dtype=np.float64)
    ext, n0, n1 = (-1.0, 3.0, -2.0, 2.0), 16, 32
    out = Emulated(prog, "zero_v10", "cmp", "complete_analysis", tmp_path).grid(p, n0, n1, ext)
    ref = orc.complete_analysis(p, n0, n1, ext)
    assert np.isnan(out[..., 0]).all() and np.isnan(ref[..., 0]).all()
    for k in (1, 2, 3, 4, 5):  # no atan / tan evaluation is left: eta and delta are exact too
        ok = _same_bits(out[..., k], ref[..., k])
        assert ok.all(), (k, out[..., k][~ok][:4], ref[..., k][~ok][:4])
    assert (np.isnan(out) == np.isnan(ref)).all() and (np.isinf(out) == np.isinf(ref)).all()
    c1 = Emulated(prog, "zero_v10", "con", "consistency_only", tmp_path).grid(p, n0, n1, ext)
    assert np.isnan(c1).all() and np.isnan(orc.consistency_only(p, n0, n1, ext)).all()
    rt = Emulated(prog, "zero_v10", "con", "consistency_rapidturn_only", tmp_path).grid(p, n0, n1, ext)
    ref_rt = orc.consistency_rapidturn_only(p, n0, n1, ext)
    assert _same_bits(rt, ref_rt).all(), (rt[~_same_bits(rt, ref_rt)][:4], ref_rt[~_same_bits(rt, ref_rt)][:4])
    if params == [1.0, 2.0, 3.0]:
        assert (rt == 1.0).any() and np.isnan(rt).any()


TRANSCENDENTAL_UNIT = """
double V(const double x[], const double args[]){
    return exp(-x[0]*x[1])*cos(x[0] + x[1]) + args[0]*log(2 + x[0]*x[0]*x[1]*x[1]) + pow(1 + x[0]*x[0], args[1]);
}
double v00(const double x[], const double args[]){
    return pow(2 + x[0]*x[1]*x[1], args[2]) + sin(x[1])*tanh(x[0]);
}
double v01(const double x[], const double args[]){
    return atan(x[0]*x[1]) + sinh(0.25*x[0]);
}
double v10(const double x[], const double args[]){
    return exp(0.5*x[1]) - log(3 + x[0]) + 0.125;
}
double v11(const double x[], const double args[]){
    return cos(x[0]*x[1]) + pow(args[0], x[1]) + 2;
}
double grad_norm_squared(const double x[], const double args[]){
    return pow(x[0]*x[0] + x[1]*x[1] + 1, 1.5) + exp(args[1]);
}
double inner_prod(const double x[], const double args[], const double v1[], const double v2[]){
    const double g00 = 1;
    const double g11 = 1;
    return 0.0 + (g00 * v1[0] * v2[0]) + (g11 * v1[1] * v2[1]);
}
void v(const double x[], const double args[], double v_out[]){
    v_out[0] = x[0];
    v_out[1] = x[1];
    return;
}
void w1(const double x[], const double args[], double v_out[]){
    v_out[0] = x[1];
    v_out[1] = x[0];
    return;
}
"""


def test_libm_calls_in_every_class(tmp_path):
    """exp / log / pow / sin / cos / tanh / sinh / atan of parameters, rows, columns and of both
    coordinates: hoisted calls of the glibc set go through inflx_gl_*, per-point (class M) ones
    stay plain libm calls (libdevice on the GPU), and the unit evaluates to the oracle's values
    (the emulation's libm is the oracle's and the hoisted calls return its bits)."""
    from raw_units import PREAMBLE

    c_text = PREAMBLE % (N_PAR, 7) + TRANSCENDENTAL_UNIT
    orc = RawOracle(c_text, str(tmp_path))
    prog = cudagen.ModelProgram(cexpr.parse_c_unit(c_text))
    gp = prog.groups["cmp"]
    classes = {}
    for i in gp.grid_nodes:
        n = gp.node(i)
        if n[0] == "f" and n[1] not in ("sqrt", "fabs"):
            classes.setdefault(gp.klass(i), set()).add(n[1])
    assert {"exp", "cos", "log", "pow"} <= classes["M"]
    assert "pow" in classes["R"] and "exp" in classes["P"] and {"sin", "exp"} <= classes["C"]
    assert gp.cols_prepass  # sin(x1), exp(x1/2) are correctly rounded calls of the column block
    src = gp.cuda_source("transc")
    grid = src[src.index("inflx_grid_complete_analysis("):]
    loop = grid[grid.index("#pragma unroll 1"): grid.index("inflx_grid_complete_analysis_sweep")]
    assert " exp(" in loop and " cos(" in loop and "inflx_gl_" not in loop
    p = np.array([1.5, 0.75, -1.25])
    ext, n0, n1 = (0.1, 2.0, -1.0, 1.5), 41, 53
    out = Emulated(prog, "transc", "cmp", "complete_analysis", tmp_path).grid(p, n0, n1, ext)
    ref = orc.complete_analysis(p, n0, n1, ext)
    err, fin, nan_mm, inf_mm = cases.rel_err(out, ref)
    assert nan_mm == 0 and inf_mm == 0 and fin.any()
    assert (err[fin] <= 1e-13).all(), float(err[fin].max())
    h = Emulated(prog, "transc", "hes", "hesse", tmp_path).grid(p, n0, n1, ext)
    err, fin, nan_mm, inf_mm = cases.rel_err(h, orc.hesse_array(p, n0, n1, ext).reshape(4, n0, n1))
    assert nan_mm == 0 and (err[fin] <= 1e-14).all(), float(err[fin].max())
