"""Parser for the C99 model-artefact source (the reference's L0 format) into one hash-consed DAG.

The reference hands its sympy expressions to a C compiler as C99 text (reference
python/inflatox/compiler.py:474-566) and the *meaning* of a model — the order in which every
`+ - * /` is rounded — is therefore whatever a C compiler makes of that text: left-associative
binary operators, unary minus binding tighter than `* /`, integer literals promoted to double, one
rounding per operation (no re-association without -ffast-math).  The CUDA back-end has to evaluate
the same operations in the same order to stay bit-identical with the reference's CPU path, so it
does not walk the sympy trees a second time; it reads the C text back with C semantics.

All functions of one translation unit are interned into ONE DAG (`Dag`).  Structurally identical
sub-expressions — across `V`, `v00..v11`, `grad_norm_squared`, the basis vectors — collapse to one
node.  That is the "joint common-subexpression reuse" of the CUDA back-end, and it is value
preserving by construction: a node is shared only if it is the same operation on the same
operands (up to the operand order of the commutative `+` and `*`, which IEEE arithmetic ignores).

Constant folding follows what gcc/clang do at -O3 *without* -ffast-math but with the reference's
`-fno-math-errno -fno-signed-zeros` (compiler.py:299-310; checked against gcc 13.3 assembly):
const∘const is folded with IEEE round-to-nearest, `pow(x,2)→x*x`, `pow(x,1)→x`, `pow(x,-1)→1/x`,
`pow(x,0.5)→sqrt(x)`, `x*1→x`, `x+0→x`, `x/2^k→x*2^-k`; nothing else is rewritten.
"""
from __future__ import annotations

import math
import re
import struct

# --------------------------------------------------------------------------------------------
# The math.h constants the generated C can name.  The reference compiles with -std=c17
# (compiler.py:307), under which glibc's <math.h> does NOT define M_PI & co, so the preamble's
# own low-precision fallbacks (compiler.py:74-88) are what the reference's CPU path really uses.
# Reproduced here digit for digit (a semantic quirk to keep, like the ones in SURVEY.md H6).
# `M_SQRT_1_2` is the reference preamble's spelling; sympy prints `M_SQRT1_2`, which under
# -std=c17 is undefined (the reference would fail to compile) - we reject it the same way.
# --------------------------------------------------------------------------------------------
REFERENCE_STRICT_C17_CONSTANTS = {
    "M_E": "2.71828182846",
    "M_LOG2E": "1.44269504089",
    "M_LOG10E": "0.4342944819",
    "M_LN2": "0.69314718056",
    "M_LN10": "2.30258509299",
    "M_PI": "3.14159265359",
    "M_PI_2": "1.57079632679",
    "M_PI_4": "0.78539816339",
    "M_1_PI": "0.31830988618",
    "M_2_PI": "0.63661977236",
    "M_2_SQRTPI": "1.1283791671",
    "M_SQRT2": "1.41421356237",
    "M_SQRT_1_2": "0.70710678118",
}

# libm functions a model may call -> arity.  Everything here has an fp64 implementation on the
# device (libdevice); anything else is rejected at compile time (north_star: no CPU/GSL fall-back).
LIBM_FUNCTIONS = {
    "pow": 2, "sqrt": 1, "cbrt": 1, "exp": 1, "exp2": 1, "expm1": 1, "log": 1, "log2": 1,
    "log10": 1, "log1p": 1, "sin": 1, "cos": 1, "tan": 1, "asin": 1, "acos": 1, "atan": 1,
    "atan2": 2, "sinh": 1, "cosh": 1, "tanh": 1, "asinh": 1, "acosh": 1, "atanh": 1,
    "fabs": 1, "hypot": 2, "erf": 1, "erfc": 1, "tgamma": 1, "lgamma": 1, "floor": 1,
    "ceil": 1, "fmin": 2, "fmax": 2, "fmod": 2,
}  # fmt: skip


class UnsupportedFunctionError(Exception):
    """Raised at Compiler.compile() time for a special function with no device implementation."""


class CParseError(Exception):
    pass


_TOKEN_RE = re.compile(
    r"\s*(?:(?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?)"
    r"|(?P<id>[A-Za-z_][A-Za-z_0-9]*)"
    r"|(?P<op><=|>=|==|!=|&&|\|\||[-+*/()\[\],=;<>!?:]))"
)


def _f2key(v: float) -> str:
    """Exact, sign-of-zero preserving key of a double."""
    return struct.pack("<d", v).hex()


class Dag:
    """Hash-consed expression DAG.  Node i is `self.nodes[i]`, a tuple:

    ("c", float)            double constant
    ("i", int)              C integer constant (only alive until it meets a double)
    ("x", k) / ("p", k)     field coordinate x[k] / model parameter args[k]
    ("xd", k)               field velocity xdot[k] (equations of motion only)
    ("v1", k) / ("v2", k)   inner_prod's vector arguments
    ("+", a, b) ("-", a, b) ("*", a, b) ("/", a, b) ("neg", a)
    ("f", name, a[, b])     libm call
    ("cmp", op, a, b)       C comparison (< > <= >= == !=): 1.0 if it holds, else 0.0
    ("and", a, b) ("or", a, b) ("not", a)   C logical operators on "non-zero is true"
    ("sel", c, a, b)        c ? a : b (c non-zero selects a); both arms are pure, so evaluating
                            both and selecting is what the C code means
    """

    def __init__(self):
        self.nodes: list[tuple] = []
        self._index: dict[tuple, int] = {}

    def _intern(self, key: tuple, node: tuple) -> int:
        i = self._index.get(key)
        if i is None:
            i = len(self.nodes)
            self.nodes.append(node)
            self._index[key] = i
        return i

    # -- leaves ------------------------------------------------------------------------------
    def const(self, v: float) -> int:
        v = float(v)
        return self._intern(("c", _f2key(v)), ("c", v))

    def iconst(self, v: int) -> int:
        return self._intern(("i", int(v)), ("i", int(v)))

    def leaf(self, kind: str, k: int) -> int:
        return self._intern((kind, k), (kind, k))

    # -- helpers -----------------------------------------------------------------------------
    def is_const(self, i: int) -> bool:
        return self.nodes[i][0] == "c"

    def is_int(self, i: int) -> bool:
        return self.nodes[i][0] == "i"

    def to_double(self, i: int) -> int:
        n = self.nodes[i]
        return self.const(float(n[1])) if n[0] == "i" else i

    def cval(self, i: int) -> float:
        return self.nodes[i][1]

    # -- operations --------------------------------------------------------------------------
    def neg(self, a: int) -> int:
        if self.is_int(a):
            return self.iconst(-self.nodes[a][1])
        if self.is_const(a):
            return self.const(-self.cval(a))
        if self.nodes[a][0] == "neg":
            return self.nodes[a][1]
        return self._intern(("neg", a), ("neg", a))

    def binop(self, op: str, a: int, b: int) -> int:
        if self.is_int(a) and self.is_int(b):
            x, y = self.nodes[a][1], self.nodes[b][1]
            if op == "+":
                return self.iconst(x + y)
            if op == "-":
                return self.iconst(x - y)
            if op == "*":
                return self.iconst(x * y)
            if y == 0:
                raise CParseError("integer division by zero in generated C")
            q = abs(x) // abs(y)  # C truncates towards zero
            return self.iconst(q if (x >= 0) == (y >= 0) else -q)
        a, b = self.to_double(a), self.to_double(b)
        ca, cb = self.is_const(a), self.is_const(b)
        if ca and cb:
            return self.const(_ieee_binop(op, self.cval(a), self.cval(b)))
        # value-preserving identities a C compiler applies under -fno-signed-zeros
        if op == "*":
            if ca and self.cval(a) == 1.0:
                return b
            if cb and self.cval(b) == 1.0:
                return a
        elif op == "+":
            if ca and self.cval(a) == 0.0:
                return b
            if cb and self.cval(b) == 0.0:
                return a
        elif op == "-":
            if cb and self.cval(b) == 0.0:
                return a
        elif op == "/":
            if cb and self.cval(b) == 1.0:
                return a
            if cb and _is_pow2(self.cval(b)):
                # x / 2^k == x * 2^-k exactly (gcc does this too); keeps a divide off the fp64 pipe
                return self.binop("*", a, self.const(1.0 / self.cval(b)))
        if op in "+*" and a > b:
            # IEEE addition and multiplication are commutative bit for bit: one canonical operand
            # order lets `a*b` and `b*a` (sympy emits both across functions) share a node
            a, b = b, a
        return self._intern((op, a, b), (op, a, b))

    # -- conditionals (sympy prints Piecewise / sign / Heaviside with these) ----------------------
    def cmp(self, op: str, a: int, b: int) -> int:
        a, b = self.to_double(a), self.to_double(b)
        if self.is_const(a) and self.is_const(b):
            x, y = self.cval(a), self.cval(b)
            r = {"<": x < y, ">": x > y, "<=": x <= y, ">=": x >= y, "==": x == y, "!=": x != y}[op]
            return self.const(1.0 if r else 0.0)
        return self._intern(("cmp", op, a, b), ("cmp", op, a, b))

    def logical(self, op: str, a: int, b: int | None = None) -> int:
        a = self.to_double(a)
        if op == "not":
            if self.is_const(a):
                return self.const(0.0 if self.cval(a) != 0.0 else 1.0)
            return self._intern(("not", a), ("not", a))
        b = self.to_double(b)
        if self.is_const(a) and self.is_const(b):
            x, y = self.cval(a) != 0.0, self.cval(b) != 0.0
            return self.const(1.0 if ((x and y) if op == "and" else (x or y)) else 0.0)
        return self._intern((op, a, b), (op, a, b))

    def select(self, c: int, a: int, b: int) -> int:
        c, a, b = self.to_double(c), self.to_double(a), self.to_double(b)
        if self.is_const(c):
            return a if self.cval(c) != 0.0 else b
        if a == b:
            return a
        return self._intern(("sel", c, a, b), ("sel", c, a, b))

    def call(self, name: str, args: list[int]) -> int:
        if name not in LIBM_FUNCTIONS:
            raise UnsupportedFunctionError(
                f'function "{name}" has no fp64 device implementation: the CUDA back-end only '
                f"accepts libm-class functions ({', '.join(sorted(LIBM_FUNCTIONS))}). GSL special "
                "functions (Bessel/hypergeometric, Compiler(link_gsl=True)) are rejected instead "
                "of silently falling back to the CPU."
            )
        if len(args) != LIBM_FUNCTIONS[name]:
            raise CParseError(f"{name} expects {LIBM_FUNCTIONS[name]} argument(s), got {len(args)}")
        args = [self.to_double(a) for a in args]
        if name == "pow":
            x, y = args
            if self.is_const(y):
                e = self.cval(y)
                if e == 2.0:
                    return self.binop("*", x, x)
                if e == 1.0:
                    return x
                if e == -1.0:
                    return self.binop("/", self.const(1.0), x)
                if e == 0.5:
                    return self.call("sqrt", [x])
                if e == 0.0:
                    return self.const(1.0)
        if all(self.is_const(a) for a in args):
            folded = _fold_call(name, [self.cval(a) for a in args])
            if folded is not None:
                return self.const(folded)
        if name == "fabs" and self.nodes[args[0]][0] == "neg":
            args = [self.nodes[args[0]][1]]
        key = ("f", name, *args)
        return self._intern(key, key)

    # -- analysis ----------------------------------------------------------------------------
    def operands(self, i: int) -> tuple[int, ...]:
        n = self.nodes[i]
        k = n[0]
        if k in ("+", "-", "*", "/"):
            return (n[1], n[2])
        if k in ("neg", "not"):
            return (n[1],)
        if k == "f":
            return tuple(n[2:])
        if k == "cmp":
            return (n[2], n[3])
        if k in ("and", "or"):
            return (n[1], n[2])
        if k == "sel":
            return (n[1], n[2], n[3])
        return ()

    def reachable(self, roots) -> list[int]:
        """Node ids reachable from `roots`, in topological (= creation) order."""
        seen = set()
        stack = [r for r in roots]
        while stack:
            i = stack.pop()
            if i in seen:
                continue
            seen.add(i)
            stack.extend(self.operands(i))
        return sorted(seen)  # operands are always created before their users


def _is_pow2(v: float) -> bool:
    if v == 0.0 or math.isinf(v) or math.isnan(v):
        return False
    m, _ = math.frexp(abs(v))
    return m == 0.5 and 1e-300 < abs(v) < 1e300


def _ieee_binop(op: str, x: float, y: float) -> float:
    try:
        if op == "+":
            return x + y
        if op == "-":
            return x - y
        if op == "*":
            return x * y
        return x / y
    except ZeroDivisionError:
        if x == 0.0 or math.isnan(x):
            return math.nan
        return math.copysign(math.inf, x) * math.copysign(1.0, y)
    except OverflowError:  # pragma: no cover - python floats do not raise on + - * overflow
        return math.inf


def _fold_call(name: str, vals: list[float]):
    """Correctly rounded compile-time evaluation (what gcc's MPFR folding yields)."""
    import mpmath

    fns = {
        "sqrt": mpmath.sqrt, "log": mpmath.log, "exp": mpmath.exp, "sin": mpmath.sin,
        "cos": mpmath.cos, "tan": mpmath.tan, "sinh": mpmath.sinh, "cosh": mpmath.cosh,
        "tanh": mpmath.tanh, "atan": mpmath.atan, "pow": mpmath.power, "cbrt": mpmath.cbrt,
        "log2": lambda v: mpmath.log(v, 2), "log10": mpmath.log10,
    }  # fmt: skip
    if name == "fabs":
        return abs(vals[0])
    fn = fns.get(name)
    if fn is None or any(math.isnan(v) or math.isinf(v) for v in vals):
        return None
    with mpmath.workprec(200):
        try:
            r = fn(*[mpmath.mpf(v) for v in vals])
        except Exception:
            return None
        if isinstance(r, mpmath.mpc):
            return None
        return float(r)


# --------------------------------------------------------------------------------------------
# C translation unit -> functions
# --------------------------------------------------------------------------------------------
_FUNC_RE = re.compile(
    r"^(?P<ret>double|void)\s+(?P<name>[A-Za-z_][A-Za-z_0-9]*)\s*\((?P<args>[^)]*)\)\s*\{\s*$"
)


class ParsedFunction:
    def __init__(self, name: str, ret: str, params: list[str]):
        self.name = name
        self.ret = ret
        self.params = params
        self.result: int | None = None  # scalar functions
        self.outputs: dict[int, int] = {}  # vector functions: v_out[i] -> node
        self.locals: dict[str, int] = {}


class ParsedUnit:
    """All functions and the metadata globals of one generated C file."""

    def __init__(self):
        self.dag = Dag()
        self.functions: dict[str, ParsedFunction] = {}
        self.version: tuple[int, int, int] | None = None
        self.dim: int | None = None
        self.n_parameters: int | None = None
        self.model_name: str | None = None
        self.use_gsl: int = 0


class _ExprParser:
    """Precedence-climbing parser for one C expression with C's arithmetic semantics."""

    def __init__(self, dag: Dag, text: str, env: dict[str, int], fn: ParsedFunction, constants):
        self.dag = dag
        self.env = env
        self.fn = fn
        self.constants = constants
        self.toks: list[tuple[str, str]] = []
        pos = 0
        n = len(text)
        while pos < n:
            m = _TOKEN_RE.match(text, pos)
            if m is None:
                if text[pos:].strip() == "":
                    break
                raise CParseError(f"cannot tokenise C expression near: {text[pos:pos + 40]!r}")
            pos = m.end()
            kind = m.lastgroup
            self.toks.append((kind, m.group(kind)))
        self.i = 0

    def _peek(self):
        return self.toks[self.i] if self.i < len(self.toks) else (None, None)

    def _next(self):
        t = self._peek()
        self.i += 1
        return t

    def _expect(self, val: str):
        k, v = self._next()
        if v != val:
            raise CParseError(f"expected {val!r}, found {v!r} in function {self.fn.name}")

    def parse(self) -> int:
        e = self._conditional()
        if self.i != len(self.toks):
            raise CParseError(f"trailing tokens in expression of {self.fn.name}: {self._peek()}")
        return e

    # C precedence, lowest first: ?:  ||  &&  == !=  < > <= >=  + -  * /  unary
    def _conditional(self) -> int:
        c = self._logical_or()
        if self._peek()[1] == "?":
            self._next()
            a = self._conditional()
            self._expect(":")
            b = self._conditional()
            return self.dag.select(c, a, b)
        return c

    def _logical_or(self) -> int:
        lhs = self._logical_and()
        while self._peek()[1] == "||":
            self._next()
            lhs = self.dag.logical("or", lhs, self._logical_and())
        return lhs

    def _logical_and(self) -> int:
        lhs = self._equality()
        while self._peek()[1] == "&&":
            self._next()
            lhs = self.dag.logical("and", lhs, self._equality())
        return lhs

    def _equality(self) -> int:
        lhs = self._relational()
        while self._peek()[1] in ("==", "!="):
            op = self._next()[1]
            lhs = self.dag.cmp(op, lhs, self._relational())
        return lhs

    def _relational(self) -> int:
        lhs = self._additive()
        while self._peek()[1] in ("<", ">", "<=", ">="):
            op = self._next()[1]
            lhs = self.dag.cmp(op, lhs, self._additive())
        return lhs

    def _additive(self) -> int:
        lhs = self._multiplicative()
        while self._peek()[1] in ("+", "-"):
            op = self._next()[1]
            rhs = self._multiplicative()
            lhs = self.dag.binop(op, lhs, rhs)
        return lhs

    def _multiplicative(self) -> int:
        lhs = self._unary()
        while self._peek()[1] in ("*", "/"):
            op = self._next()[1]
            rhs = self._unary()
            lhs = self.dag.binop(op, lhs, rhs)
        return lhs

    def _unary(self) -> int:
        v = self._peek()[1]
        if v == "-":
            self._next()
            return self.dag.neg(self._unary())
        if v == "+":
            self._next()
            return self._unary()
        if v == "!":
            self._next()
            return self.dag.logical("not", self._unary())
        return self._primary()

    def _index(self) -> int:
        self._expect("[")
        k, v = self._next()
        if k != "num" or not v.isdigit():
            raise CParseError(f"non-literal array index in {self.fn.name}")
        self._expect("]")
        return int(v)

    def _primary(self) -> int:
        k, v = self._next()
        if k == "num":
            if v.isdigit():
                return self.dag.iconst(int(v))
            return self.dag.const(float(v))
        if v == "(":
            e = self._conditional()
            self._expect(")")
            return e
        if k == "id":
            nxt = self._peek()[1]
            if nxt == "[":
                idx = self._index()
                kind = {"x": "x", "args": "p", "xdot": "xd", "v1": "v1", "v2": "v2"}.get(v)
                if kind is None or v not in self.fn.params:
                    raise CParseError(f"unknown array {v!r} in function {self.fn.name}")
                return self.dag.leaf(kind, idx)
            if nxt == "(":
                self._next()
                args = []
                if self._peek()[1] != ")":
                    args.append(self._conditional())
                    while self._peek()[1] == ",":
                        self._next()
                        args.append(self._conditional())
                self._expect(")")
                return self.dag.call(v, args)
            if v in self.env:
                return self.env[v]
            if v in self.constants:
                return self.dag.const(float(self.constants[v]))
            raise CParseError(f"unknown identifier {v!r} in function {self.fn.name}")
        raise CParseError(f"unexpected token {v!r} in function {self.fn.name}")


_LOCAL_RE = re.compile(r"^const\s+double\s+([A-Za-z_][A-Za-z_0-9]*)\s*=\s*(.*);$", re.S)
_VOUT_RE = re.compile(r"^v_out\[(\d+)\]\s*=\s*(.*);$", re.S)
_RETURN_RE = re.compile(r"^return\s*(.*);$", re.S)


def parse_c_unit(text: str, constants: dict[str, str] | None = None) -> ParsedUnit:
    """Parse a generated model source (reference compiler.py:474-566 layout) into a ParsedUnit."""
    constants = REFERENCE_STRICT_C17_CONSTANTS if constants is None else constants
    unit = ParsedUnit()
    m = re.search(r"VERSION\[3\]\s*=\s*\{(\d+),(\d+),(\d+)\}", text)
    if m:
        unit.version = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    m = re.search(r"uint32_t\s+DIM\s*=\s*(\d+)", text)
    if m:
        unit.dim = int(m.group(1))
    m = re.search(r"uint32_t\s+N_PARAMETERS\s*=\s*(\d+)", text)
    if m:
        unit.n_parameters = int(m.group(1))
    m = re.search(r'MODEL_NAME\s*=\s*"([^"]*)"', text)
    if m:
        unit.model_name = m.group(1)
    m = re.search(r"char\s+USE_GSL\s*=\s*(\d+)", text)
    if m:
        unit.use_gsl = int(m.group(1))

    cur: ParsedFunction | None = None
    pending = ""
    for raw in text.split("\n"):
        line = raw.strip()
        if cur is not None and line != "}" and not line.startswith("//"):
            # a statement may span several lines (sympy prints Piecewise that way): join up to ';'
            pending = (pending + " " + line).strip() if pending else line
            if pending and not pending.endswith(";"):
                continue
            line, pending = pending, ""
        if cur is None:
            fm = _FUNC_RE.match(line)
            if fm:
                params = re.findall(r"([A-Za-z_][A-Za-z_0-9]*)\s*\[\]", fm.group("args"))
                cur = ParsedFunction(fm.group("name"), fm.group("ret"), params)
            continue
        if line == "}":
            unit.functions[cur.name] = cur
            cur = None
            continue
        if not line or line.startswith("//"):
            continue
        lm = _LOCAL_RE.match(line)
        if lm:
            cur.locals[lm.group(1)] = _ExprParser(
                unit.dag, lm.group(2), cur.locals, cur, constants
            ).parse()
            continue
        vm = _VOUT_RE.match(line)
        if vm:
            cur.outputs[int(vm.group(1))] = unit.dag.to_double(
                _ExprParser(unit.dag, vm.group(2), cur.locals, cur, constants).parse()
            )
            continue
        rm = _RETURN_RE.match(line)
        if rm:
            if rm.group(1).strip():
                cur.result = unit.dag.to_double(
                    _ExprParser(unit.dag, rm.group(1), cur.locals, cur, constants).parse()
                )
            continue
        raise CParseError(f"cannot parse statement in {cur.name}: {line[:80]!r}")
    return unit
