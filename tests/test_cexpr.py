"""The C-text parser gives every expression the meaning gcc gives it: a Python interpreter of the
DAG (IEEE double ops, the platform libm through `math`) must reproduce the gcc-compiled
reference-generated C bit for bit."""
import math

import numpy as np
import pytest

import cases
import oracle
from inflatox_b200 import cexpr


def evaluate(dag, roots, x, p):
    vals = {}
    for i in dag.reachable(roots):
        n = dag.nodes[i]
        k = n[0]
        if k == "c":
            v = n[1]
        elif k == "i":
            v = float(n[1])
        elif k == "x":
            v = x[n[1]]
        elif k == "p":
            v = p[n[1]]
        elif k == "neg":
            v = -vals[n[1]]
        elif k in "+-*/":
            a, b = vals[n[1]], vals[n[2]]
            if k == "+":
                v = a + b
            elif k == "-":
                v = a - b
            elif k == "*":
                v = a * b
            else:
                v = cexpr._ieee_binop("/", a, b)
        elif k == "cmp":
            a, b = vals[n[2]], vals[n[3]]
            v = 1.0 if {"<": a < b, ">": a > b, "<=": a <= b, ">=": a >= b, "==": a == b,
                        "!=": a != b}[n[1]] else 0.0  # fmt: skip
        elif k == "and":
            v = 1.0 if (vals[n[1]] != 0.0 and vals[n[2]] != 0.0) else 0.0
        elif k == "or":
            v = 1.0 if (vals[n[1]] != 0.0 or vals[n[2]] != 0.0) else 0.0
        elif k == "not":
            v = 1.0 if vals[n[1]] == 0.0 else 0.0
        elif k == "sel":
            v = vals[n[2]] if vals[n[1]] != 0.0 else vals[n[3]]
        elif k == "f":
            args = [vals[a] for a in n[2:]]
            extra = {
                "fmax": lambda a, b: b if math.isnan(a) else a if math.isnan(b) else max(a, b),
                "fmin": lambda a, b: b if math.isnan(a) else a if math.isnan(b) else min(a, b),
            }
            try:
                v = (extra.get(n[1]) or getattr(math, n[1]))(*args)
            except (ValueError, OverflowError):
                v = math.nan
        else:
            raise AssertionError(n)
        vals[i] = v
    return [vals[r] for r in roots]


@pytest.mark.parametrize("model", cases.MODELS)
def test_dag_matches_gcc_bitwise(model):
    unit = cexpr.parse_c_unit(oracle.golden_c_text(model))
    f = unit.functions
    roots = [f[n].result for n in ("V", "v00", "v10", "v11", "grad_norm_squared")]
    orc = oracle.Oracle(model)
    p, ext = cases.params(model), cases.EXTENT[model]
    n0, n1 = 6, 5
    ref = orc.model_functions(p, n0, n1, ext)
    dx0, dx1 = (ext[1] - ext[0]) / n0, (ext[3] - ext[2]) / n1
    same = total = 0
    for i in range(n0):
        for j in range(n1):
            x = [i * dx0 + ext[0], j * dx1 + ext[2]]
            got = evaluate(unit.dag, roots, x, list(p))
            for a, b in zip(got, ref[i, j]):
                total += 1
                same += (a == b) or (math.isnan(a) and math.isnan(b))
    assert same == total, f"{model}: {same}/{total} values bit-identical"


def test_metadata_and_constants():
    unit = cexpr.parse_c_unit(oracle.golden_c_text("d5"))
    assert unit.version == (5, 0, 0) and unit.dim == 2 and unit.n_parameters == 10
    assert unit.model_name == "d5"
    # -std=c17 hides glibc's M_PI: the short fall-back of the preamble is what gets evaluated
    assert cexpr.REFERENCE_STRICT_C17_CONSTANTS["M_PI"] == "3.14159265359"


def test_c_semantics():
    d = cexpr.Dag()
    fn = cexpr.ParsedFunction("f", "double", ["x", "args"])
    parse = lambda t: cexpr._ExprParser(d, t, {}, fn, cexpr.REFERENCE_STRICT_C17_CONSTANTS).parse()
    # integer division truncates, literals fold, unary minus binds tighter than * /
    assert d.nodes[d.to_double(parse("7/2"))] == ("c", 3.0)
    assert d.nodes[d.to_double(parse("1.0/2.0"))] == ("c", 0.5)
    assert set(d.nodes[parse("-1.0/2.0*x[0]")][1:]) == {d.const(-0.5), d.leaf("x", 0)}
    # pow(x,2) -> x*x, x/4 -> x*0.25, left associativity
    assert d.nodes[parse("pow(x[0], 2)")] == ("*", d.leaf("x", 0), d.leaf("x", 0))
    assert set(d.nodes[parse("x[0]/4")][1:]) == {d.leaf("x", 0), d.const(0.25)}
    # commutative operators are canonicalised (bit-identical in IEEE arithmetic) ...
    assert parse("x[0]*x[1]") == parse("x[1]*x[0]") and parse("x[0]+args[0]") == parse("args[0]+x[0]")
    # ... the others are not
    assert parse("x[0]-x[1]") != parse("x[1]-x[0]") and parse("x[0]/x[1]") != parse("x[1]/x[0]")
    a = parse("x[0] - x[1] - args[0]")
    assert d.nodes[a][0] == "-" and d.nodes[d.nodes[a][1]][0] == "-"
    # identical sub-trees are one node
    assert parse("sin(x[0])*x[1]") == parse("sin(x[0])*x[1]")


def test_unsupported_functions_are_rejected():
    text = "double V(const double x[], const double args[]){\n    return gsl_sf_bessel_J0(x[0]);\n}\n"
    with pytest.raises(cexpr.UnsupportedFunctionError, match="no fp64 device implementation"):
        cexpr.parse_c_unit(text)


@pytest.mark.parametrize("seed", range(200, 212))
def test_random_units_with_conditionals_match_gcc_bitwise(seed, tmp_path):
    """Random C units with comparisons, && || !, ?: (what sympy prints for Piecewise / sign /
    Heaviside, multi-line statements included): the DAG means what gcc makes of the text."""
    import numpy as np

    import raw_units

    text = raw_units.make_unit(seed, conditionals=True)
    unit = cexpr.parse_c_unit(text)
    f = unit.functions
    roots = [f[n].result for n in ("V", "v00", "v10", "v11", "grad_norm_squared")]
    orc = raw_units.RawOracle(text, str(tmp_path))
    p = [1.5, 0.75, 2.25]
    ext = (-2.0, 3.0, -1.5, 2.5)
    n0, n1 = 9, 11
    ref = orc.model_functions(np.array(p), n0, n1, ext)
    dx0, dx1 = (ext[1] - ext[0]) / n0, (ext[3] - ext[2]) / n1
    with np.errstate(all="ignore"):
        for i in range(n0):
            for j in range(n1):
                x = [i * dx0 + ext[0], j * dx1 + ext[2]]
                got = evaluate(unit.dag, roots, x, p)
                for a, b in zip(got, ref[i, j]):
                    assert a == b or (math.isnan(a) and math.isnan(b)), (seed, i, j, a, b)


def test_sympy_conditionals_parse():
    import sympy
    from sympy.printing.c import C99CodePrinter

    x, y, a = sympy.symbols("Q0 Q1 Q2")
    pr = C99CodePrinter()
    exprs = [
        sympy.Piecewise((x**2, x > 0), (y, True)),
        sympy.sign(x) * y,
        sympy.Heaviside(x),
        sympy.Max(x, y, a),
        sympy.Piecewise((x, sympy.And(x > 0, y < 1)), (a, x <= -1), (0, True)),
        sympy.Piecewise((1, sympy.Or(x > 1, ~(y > 0))), (0, True)),
    ]
    for e in exprs:
        text = pr.doprint(e).replace("Q0", "x[0]").replace("Q1", "x[1]").replace("Q2", "args[0]")
        unit = cexpr.parse_c_unit(
            "double V(const double x[], const double args[]){\n    return " + text + ";\n}\n"
        )
        root = unit.functions["V"].result
        for xv, yv, av in [(1.5, 0.5, 2.0), (-2.0, 3.0, 0.25), (0.0, -1.0, 1.0)]:
            want = float(e.subs({x: xv, y: yv, a: av}))
            (got,) = evaluate(unit.dag, [root], [xv, yv], [av])
            assert got == want, (e, xv, yv, av, got, want)
