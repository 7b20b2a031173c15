#!/usr/bin/env python3
"""Regenerates the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py [model ...]

What it does, per model (doc, hyper, angular, egno, d5):
  1. imports the reference's Python half from /root/reference/python through three shims
     placed in a temp dir (a SIGALRM `interruptingcow.timeout`, a no-op `inflatox.libinflx_rs`,
     and dist-info metadata so `importlib.metadata.version("inflatox")` resolves); the shims are
     on PYTHONPATH because the reference's joblib workers re-import the package
     (reference python/inflatox/symbolic.py:376);
  2. builds the model exactly like the reference's tests do (tests/test_doc.py:27-36,
     README.md:59-66, tests/test_angular.py:39-60, tests/test_egno.py:39-77,
     tests/test_d5.py:40-141) with the reference `InflationModelBuilder`;
  3. pickles the resulting symbolic expressions (plain dict of sympy objects, no reference
     classes inside) to tests/golden/models/<name>.pkl.gz — this is the INPUT fixture of
     `inflatox_b200.Compiler`;
  4. lets the reference `Compiler._generate_c_file()` (compiler.py:474-566) emit the C99 model
     artefact source with the CSE setting the reference's test uses, and stores it as
     tests/golden/c/<name>.c.gz together with the symbol dictionary — this is the INPUT of the
     oracle (oracle/), i.e. the reference's own generated code.
Nothing from the reference's *source* is copied: both fixtures are outputs of running it.
"""
import gzip
import json
import os
import pickle
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _install_shims() -> str:
    site = tempfile.mkdtemp(prefix="inflx_ref_site_")
    pkg = os.path.join(site, "inflatox")
    os.makedirs(pkg)
    for f in os.listdir(f"{REF}/python/inflatox"):
        if f.endswith(".py"):
            os.symlink(f"{REF}/python/inflatox/{f}", os.path.join(pkg, f))
    with open(os.path.join(pkg, "libinflx_rs.py"), "w") as fh:
        fh.write(
            "import sys\n"
            "def log_info(m): print('[info]', m, file=sys.stderr)\n"
            "def log_warn(m): print('[warn]', m, file=sys.stderr)\n"
            "class InflatoxPyDyLib: pass\n"
            "def _noop(*a, **k): return None\n"
            "for _n in ['open_inflx_dylib','flag_quantum_dif_py','consistency_only',"
            "'consistency_rapidturn_only','epsilon_v_only','complete_analysis',"
            "'complete_analysis_on_trajectory','consistency_only_on_trajectory',"
            "'consistency_rapidturn_only_on_trajectory','epsilon_v_only_on_trajectory',"
            "'solve_eom_rk4','solve_eom_rkf']:\n"
            "    globals()[_n] = _noop\n"
        )
    os.makedirs(os.path.join(site, "inflatox-0.10.0.dist-info"))
    with open(os.path.join(site, "inflatox-0.10.0.dist-info", "METADATA"), "w") as fh:
        fh.write("Metadata-Version: 2.1\nName: inflatox\nVersion: 0.10.0\n")
    with open(os.path.join(site, "interruptingcow.py"), "w") as fh:
        fh.write(
            "import signal, contextlib\n"
            "@contextlib.contextmanager\n"
            "def timeout(seconds, exception=RuntimeError):\n"
            "    def handler(signum, frame):\n"
            "        raise exception\n"
            "    old = signal.signal(signal.SIGALRM, handler)\n"
            "    signal.setitimer(signal.ITIMER_REAL, seconds)\n"
            "    try:\n"
            "        yield\n"
            "    finally:\n"
            "        signal.setitimer(signal.ITIMER_REAL, 0)\n"
            "        signal.signal(signal.SIGALRM, old)\n"
        )
    sys.path.insert(0, site)
    os.environ["PYTHONPATH"] = site + os.pathsep + os.environ.get("PYTHONPATH", "")
    return site


# ------------------------------------------------------------------------------------------
# Model definitions (user-level inputs; same construction steps as the reference's tests so
# that sympy arrives at the same expression trees)
# ------------------------------------------------------------------------------------------
def build_doc(inflatox, sympy):
    r, th, m = sympy.symbols("r θ m")
    V = (1 / 2 * m**2 * (th**2 - 2 / (3 * r**2))).nsimplify()
    g = [[0.5, 0], [0, 0.5 * r**2]]
    return inflatox.InflationModelBuilder.new([r, th], g, V, silent=True).build(), dict(cse=False)


def build_hyper(inflatox, sympy):
    phi, th, L, m, phi0 = sympy.symbols("φ θ L m φ0")
    V = (1 / 2 * m**2 * (phi - phi0) ** 2).nsimplify()
    g = [[1, 0], [0, L**2 * sympy.sinh(phi / L) ** 2]]
    model = inflatox.InflationModelBuilder.new(
        [phi, th], g, V, model_name="hyperinflation", silent=True
    ).build()
    return model, dict(cse=False)


def build_angular(inflatox, sympy):
    p, x = sympy.symbols("phi chi")
    mp, mx, a = sympy.symbols("m_phi m_chi alpha")
    potential = a / 2 * ((mp * p) ** 2 + (mx * x) ** 2).nsimplify()
    diag = 6 * a / (1 - p**2 - x**2) ** 2
    metric = [[diag, 0], [0, diag]]
    model = inflatox.InflationModelBuilder.new(
        [p, x], metric, potential, model_name="angular", silent=True
    ).build()
    return model, dict(cse=True)


def build_egno(inflatox, sympy):
    alpha, m, p, c, a = sympy.symbols("alpha m p c a")
    r, th = sympy.symbols("r θ")
    Phi, Phi_Bar, S, S_Bar = sympy.symbols("Phi Phi_B S S_B")
    K = (
        -3 * alpha * sympy.ln(Phi + Phi_Bar - c * (Phi + Phi_Bar - 1) ** 4)
        + (S * S_Bar) / (Phi + Phi_Bar) ** 3
    ).nsimplify()
    sf, sfc = [Phi, S], [Phi_Bar, S_Bar]
    metric = [[sympy.diff(sympy.diff(K, sf[b]), sfc[a_]) for a_ in range(2)] for b in range(2)]
    metric = [
        [g.subs({Phi: r + 1j * th, Phi_Bar: r - 1j * th}).nsimplify().simplify() for g in gb]
        for gb in metric
    ]
    metric = [[g.subs({S: 0, S_Bar: 0}).simplify() for g in gb] for gb in metric]
    real_metric = [[metric[0][0], 0], [0, metric[0][0]]]
    potential = (
        (6 * m**2 * r**3 * ((a - r) ** 2 + th**2))
        / (a**2 * (2 * r - c * (1 - 2 * r) ** 4) ** (3 * alpha))
    ).nsimplify()
    model = inflatox.InflationModelBuilder.new(
        [r, th],
        real_metric,
        potential,
        model_name="egno",
        silent=True,
        simplify=False,
        assertions=False,
    ).build([[0, 1]])
    return model, dict(cse=True)


def build_d5(inflatox, sympy):
    from sympy.simplify.radsimp import collect_sqrt

    r, th = sympy.symbols("r θ2")
    gs, ls, N = sympy.symbols("g_s l_s N")
    mu5 = 1 / ((2 * sympy.pi) ** 5 * ls**6)
    T5 = mu5 / gs
    u = sympy.symbols("u")
    rho = r / (3 * u)
    H = (
        ((sympy.pi * N * gs * ls**4) / (12 * u**4) * (2 / rho**2 - 2 * sympy.ln(1 / rho**2 + 1)))
        .nsimplify()
        .collect([u, r])
        .expand()
        .powsimp(force=True)
    )
    p, q = sympy.symbols("p q")
    F = (
        (H / 9 * (r**2 + 3 * u**2) ** 2 + (sympy.pi * q * ls**2) ** 2)
        .nsimplify()
        .collect([r, u])
        .expand()
        .powsimp()
    )
    gamma = 4 * sympy.pi**2 * ls**2 * p * q * T5 * gs
    sqrtF = sympy.sqrt(F)
    g00 = (
        collect_sqrt(
            4 * sympy.pi * p * T5 * sqrtF * ((r**2 + 6 * u**2) / (r**2 + p * u**2)),
            evaluate=True,
        )
        .expand()
        .powsimp()
    )
    g11 = (
        collect_sqrt((4 / 6) * sympy.pi * p * T5 * sqrtF * (r**2 + 6 * u**2), evaluate=True)
        .nsimplify()
        .collect([r, u])
        .expand()
        .powsimp()
    )
    metric = [[g00, 0], [0, g11]]
    Phi_min = (
        (
            (5 / 72)
            * (
                81 * (9 * rho**2 - 2) * rho**2
                + 162 * sympy.ln(9 * (rho**2 + 1))
                + -9
                + -160 * sympy.ln(10)
            )
        )
        .nsimplify()
        .collect([u])
        .expand()
        .powsimp()
    )
    a0, a1, b1 = sympy.symbols("a0 a1 b1")
    Phi_h = (
        (
            a0 * (2 / rho**2 - 2 * sympy.ln(1 / rho**2 + 1))
            + 2
            * a1
            * (6 + 1 / rho**2 - 2 * (2 + 3 * rho**2) * sympy.ln(1 + 1 / rho**2))
            * sympy.cos(th)
            + (b1 / 2) * (2 + 3 * rho**2) * sympy.cos(th)
        )
        .nsimplify()
        .collect([u, r])
        .expand()
        .powsimp()
    )
    V0 = sympy.symbols("V0")
    potential = (
        V0
        + (4 * sympy.pi * p * T5 / H) * (sympy.sqrt(F) - (ls**2) * sympy.pi * q * gs)
        + gamma * (Phi_min + Phi_h)
    )
    potential = potential.nsimplify().collect([ls, gs]).expand().powsimp()
    model = inflatox.InflationModelBuilder.new(
        [r, th],
        metric,
        potential,
        model_name="d5",
        assertions=False,
        silent=True,
        simplify=False,
    ).build([[1, 0]])
    return model, dict(cse=False)


BUILDERS = {
    "doc": build_doc,
    "hyper": build_hyper,
    "angular": build_angular,
    "egno": build_egno,
    "d5": build_d5,
}


def main(names):
    _install_shims()
    import contextlib
    import io

    import inflatox  # the reference, through the shims
    import sympy

    os.makedirs(os.path.join(HERE, "models"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "c"), exist_ok=True)
    for name in names:
        t0 = time.time()
        model, opts = BUILDERS[name](inflatox, sympy)
        t1 = time.time()
        fields = dict(
            model_name=model.model_name,
            coordinates=model.coordinates,
            tangents=model.coordinate_tangents,
            basis=model.basis,
            eom_fields=model.eom_fields,
            eom_h=model.eom_h,
            eom_hdot=model.eom_hdot,
            potential=model.potential,
            metric=model.metric,
            gradient_square=model.gradient_square,
            hesse_cmp=model.hesse_cmp,
        )
        with gzip.GzipFile(os.path.join(HERE, "models", f"{name}.pkl.gz"), "wb", mtime=0) as fh:
            pickle.dump(fields, fh, protocol=4)

        c_path = os.path.join(tempfile.gettempdir(), f"inflx_golden_{name}.c")
        comp = inflatox.Compiler(model, output_path=c_path, cleanup=False, silent=True, **opts)
        with contextlib.redirect_stdout(io.StringIO()):  # compiler.py:509 prints unconditionally
            comp._generate_c_file()
        with open(c_path) as fh:
            c_text = fh.read()
        os.remove(c_path)
        # drop the timestamp / interpreter lines of the preamble so the fixture is reproducible
        c_lines = [
            ln
            for ln in c_text.split("\n")
            if not (ln.startswith("// Model:") or ln.startswith("// System info:"))
        ]
        c_text = "\n".join(c_lines)
        with gzip.GzipFile(os.path.join(HERE, "c", f"{name}.c.gz"), "wb", mtime=0) as fh:
            fh.write(c_text.encode())
        meta = dict(
            name=name,
            cse=opts["cse"],
            symbol_dictionary=comp.symbol_dict,
            n_fields=model.dim,
            n_parameters=len(comp.symbol_dict) - model.dim,
            sympy=sympy.__version__,
            symbolic_seconds=round(t1 - t0, 1),
            c_bytes=len(c_text),
        )
        with open(os.path.join(HERE, "c", f"{name}.json"), "w") as fh:
            json.dump(meta, fh, indent=1, ensure_ascii=False)
        print(f"{name}: symbolic {t1 - t0:.1f}s, C {len(c_text)} bytes, params {comp.symbol_dict}")


if __name__ == "__main__":
    main(sys.argv[1:] or list(BUILDERS))
