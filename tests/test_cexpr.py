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
        elif k == "f":
            args = [vals[a] for a in n[2:]]
            try:
                v = getattr(math, n[1])(*args)
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
