"""Parity of the CUDA path (through the C ABI / the drop-in module) with the CPU oracle.

Bar (BASELINE.json north_star): NaN / +-inf masks bit-identical; <= 1e-10 relative error per
finite point - asserted here for EVERY finite point of every model, plane and entry point.
+,-,*,/,sqrt are evaluated bit-identically by construction, and the model's hoisted libm calls
(pow, log, exp, sin, cos, tanh per parameter vector / row / column) return the bits of the
reference host's glibc (csrc/inflx_glibcmath.cuh): round 1's residue (EGNO 99.0-99.7 %, angular
99.87-100 % within 1e-10) was glibc's own misrounding of a row-level pow, amplified by the
models' conditioning.  What may still differ in the last bits in the default build: the epilogue's
atan / tan (delta, eta) and per-point literal half-integer powers (correctly rounded dd chains);
libm flavour "glibc-all" removes both, and every output bit is then the oracle's.
"""
import ctypes
import math

import numpy as np
import pytest

import cases
import oracle
from inflatox_b200 import libinflx_rs as rs

pytestmark = pytest.mark.gpu

N0, N1 = 203, 157  # ragged on purpose: neither a multiple of the CTA width nor of rows/thread


@pytest.fixture(scope="module", params=cases.MODELS)
def setup(request):
    m = request.param
    lib = rs.open_inflx_dylib(cases.artifact(m).shared_object_path, False)
    lib.set_devices([0])
    return m, lib, oracle.Oracle(m), cases.params(m), cases.EXTENT[m]


def ss_of(ext):
    return np.array([[ext[0], ext[1]], [ext[2], ext[3]]])


def check(model, got, ref, what=""):
    """north_star's bar, no allowance: masks identical, every finite point within 1e-10."""
    err, fin, nan_mm, inf_mm = cases.rel_err(got, ref)
    assert nan_mm == 0, f"{model} {what}: {nan_mm} NaN-mask mismatches"
    assert inf_mm == 0, f"{model} {what}: {inf_mm} inf-mask mismatches"
    n_bad = int((err[fin] > 1e-10).sum())
    worst = float(err[fin].max()) if fin.any() else 0.0
    assert n_bad == 0, f"{model} {what}: {n_bad} of {int(fin.sum())} finite points beyond 1e-10 (max {worst:.3g})"


def test_complete_analysis(setup):
    m, lib, orc, p, ext = setup
    out = np.zeros((N0, N1, 6))
    rs.complete_analysis(lib, p, out, ss_of(ext), False, 0)
    ref = orc.complete_analysis(p, N0, N1, ext)
    for k, name in enumerate(["consistency", "eps_V", "eps_H", "eta", "delta", "omega"]):
        check(m, out[..., k], ref[..., k], what=name)


@pytest.mark.parametrize("model", cases.MODELS)
def test_flavour_glibc_all_reproduces_every_bit_of_the_oracle(model):
    """`INFLATOX_LIBM=glibc-all`: the per-point literal half-integer powers of EGNO and d5 go through
    the restated glibc `pow` (default: correctly rounded dd chains, which differ from glibc's in
    the last bit on ~1e-3 of the calls) and the epilogue's delta / tan(delta) through the restated
    glibc `atan` / `tan`.  Every operation of the path is then the reference's own - IEEE
    +,-,*,/,sqrt and glibc's libm - so all six planes, the single-plane ops and the on-trajectory
    op must equal the oracle's BIT FOR BIT."""
    lib = rs.open_inflx_dylib(cases.artifact(model, libm="glibc-all").shared_object_path, False)
    lib.set_devices([0])
    orc, p, ext = oracle.Oracle(model), cases.params(model), cases.EXTENT[model]

    def same(a, b):
        return (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b)) | ((a == 0) & (b == 0))

    out = np.zeros((N0, N1, 6))
    rs.complete_analysis(lib, p, out, ss_of(ext), False, 0)
    ref = orc.complete_analysis(p, N0, N1, ext)
    for k, name in enumerate(["consistency", "eps_V", "eps_H", "eta", "delta", "omega"]):
        ok = same(out[..., k], ref[..., k])
        assert ok.all(), f"{model} {name}: {int((~ok).sum())} points differ in the last bits"
    one = np.zeros((N0, N1))
    rs.consistency_only(lib, p, one, ss_of(ext), False, 0)
    assert same(one, orc.consistency_only(p, N0, N1, ext)).all()
    if model in ("angular", "egno", "d5"):
        xs = np.ascontiguousarray(cases.trajectory(model)[:2000])
    else:
        rng = np.random.default_rng(11)
        xs = np.ascontiguousarray(np.stack([rng.uniform(ext[0], ext[1], 2000),
                                            rng.uniform(ext[2], ext[3], 2000)], axis=1))
    traj = np.zeros((xs.shape[0], 6))
    rs.complete_analysis_on_trajectory(lib, p, xs, traj, False, 0)
    ok = same(traj, orc.complete_analysis_on_trajectory(p, xs))
    assert ok.all(), f"{model} on trajectory: {int((~ok).sum())} values differ"


@pytest.mark.parametrize(
    "fn", ["consistency_only", "consistency_rapidturn_only", "epsilon_v_only"]
)
def test_single_plane_ops(setup, fn):
    m, lib, orc, p, ext = setup
    out = np.zeros((N0, N1))
    getattr(rs, fn)(lib, p, out, ss_of(ext), False, 0)
    check(m, out, getattr(orc, fn)(p, N0, N1, ext), what=fn)


def test_flag_quantum_dif(setup):
    m, lib, orc, p, ext = setup
    # thresholds inside and outside the range of the basis-vector components; a byte output is
    # bit-exact in this tier: not one flag may differ
    seen = set()
    for acc in (0.5, 0.0, -0.3, 1.5):
        x = np.zeros((N0, N1), dtype=bool)
        rs.flag_quantum_dif_py(lib, p, x, ss_of(ext), False, acc)
        ref = orc.flag_quantum_dif(p, N0, N1, ext, acc)
        assert (x != ref).sum() == 0, f"{m} accuracy {acc}: {(x != ref).sum()} flags differ"
        seen |= set(np.unique(x).tolist())
    assert seen == {False, True}, (m, seen)


def test_potential_and_hesse_arrays(setup):
    m, lib, orc, p, ext = setup
    v = np.zeros((N0, N1))
    lib.potential_array(v, p, ss_of(ext))
    check(m, v, orc.potential_array(p, N0, N1, ext), what="potential_array")
    h = lib.hesse_array(np.array([N0, N1]), p, ss_of(ext))
    assert h.shape == (2, 2, N0, N1)
    ref = orc.hesse_array(p, N0, N1, ext)
    for a in range(2):
        for b in range(2):
            check(m, h[a, b], ref[a, b], what=f"hesse_array[{a}{b}]")


def test_scalar_entry_points(setup):
    m, lib, orc, p, ext = setup
    x = np.array([0.3 * ext[0] + 0.7 * ext[1], 0.6 * ext[2] + 0.4 * ext[3]])
    v, vr = lib.potential(x, p), orc.potential(x, p)
    assert v == vr or abs(v - vr) <= 1e-10 * abs(vr)
    h, hr = lib.hesse(x, p), orc.hesse(x, p)
    assert h.shape == (2, 2)
    assert np.allclose(h, hr, rtol=1e-8, atol=0, equal_nan=True)


def test_row_shard_uses_global_coordinates(setup):
    m, lib, orc, p, ext = setup
    full = np.zeros((N0, N1, 6))
    rs.grid_eval(lib, "complete_analysis", p, full, N0, N1, ext)
    part = np.zeros((64, N1, 6))
    rs.grid_eval(lib, "complete_analysis", p, part, N0, N1, ext, rows=(37, 101))
    assert np.array_equal(full[37:101], part, equal_nan=True)


def test_chunked_pipeline_matches_single_launch(setup, monkeypatch):
    """Row chunking + double-buffered D2H (forced by a 1 MiB chunk target) changes no bit."""
    m, lib, orc, p, ext = setup
    a = np.zeros((N0, N1, 6))
    rs.grid_eval(lib, "complete_analysis", p, a, N0, N1, ext)
    monkeypatch.setenv("INFLATOX_CHUNK_MB", "1")
    b = np.zeros((N0, N1, 6))
    rep = rs.grid_eval(lib, "complete_analysis", p, b, N0, N1, ext)
    assert rep["launches"] > 3
    assert np.array_equal(a, b, equal_nan=True)
    c = rs.pinned_empty((N0, N1, 6))  # direct DMA path
    rs.grid_eval(lib, "complete_analysis", p, c, N0, N1, ext)
    assert np.array_equal(a, c, equal_nan=True)


def test_fused_sweep_equals_separate_calls(setup):
    m, lib, orc, p, ext = setup
    rng = np.random.default_rng(5)
    S, n0, n1 = 5, 40, 70
    ps = p[None, :] * (1.0 + 0.05 * rng.standard_normal((S, p.size)))
    fused = np.zeros((S, n0, n1, 6))
    rs.sweep(lib, "complete_analysis", ps, fused, ext)
    for s in range(S):
        one = np.zeros((n0, n1, 6))
        rs.complete_analysis(lib, np.ascontiguousarray(ps[s]), one, ss_of(ext), False, 0)
        assert np.array_equal(fused[s], one, equal_nan=True), (m, s)


def test_c5_parameter_vectors_against_the_oracle():
    """BASELINE C5's own inputs: the 1024 default_rng(0) vectors (L ~ U(0.05, 2), m ~ 10^U(-3, 1),
    phi0 ~ U(-1, 1); tanh(x / L) reaches |x / L| = 20), every one against the oracle on a 64 x 64
    grid through the fused sweep."""
    import bench

    model, op, _, _, S, ext, ps = bench.workload("C5")
    assert model == "hyper" and S == 1024 and ps.shape == (1024, 3)
    lib = rs.open_inflx_dylib(cases.artifact(model).shared_object_path, False)
    lib.set_devices([0])
    orc = oracle.Oracle(model)
    n = 64
    fused = np.zeros((S, n, n, 6))
    rs.sweep(lib, "complete_analysis", ps, fused, ext)
    for k in range(S):
        ref = orc.complete_analysis(np.ascontiguousarray(ps[k]), n, n, ext)
        for c in range(6):
            check(model, fused[k, ..., c], ref[..., c], what=f"C5 vector {k} plane {c}")


def test_device_resident_output_equals_host_output(setup):
    import torch

    m, lib, orc, p, ext = setup
    host = np.zeros((N0, N1, 6))
    rs.grid_eval(lib, "complete_analysis", p, host, N0, N1, ext)
    d = torch.empty(N0 * N1 * 6, dtype=torch.float64, device="cuda:0")
    rep = rs.grid_eval(lib, "complete_analysis", p, None, N0, N1, ext, device=0,
                       out_device_ptr=d.data_ptr())  # fmt: skip
    assert rep["d2h_bytes"] == 0 and rep["grid_ms"] > 0
    assert np.array_equal(d.cpu().numpy().reshape(N0, N1, 6), host, equal_nan=True)


def test_tail_tiles_do_not_change_a_bit(setup, monkeypatch):
    """One launch with two tile heights (the engine's default: the last ~1.5 waves of rows on
    quarter-height tiles) against uniform tiles and against other forced heights, on a ragged
    8190 x 1000 grid in ONE launch (device-resident), where both kinds of tile occur."""
    import torch

    m, lib, orc, p, ext = setup
    n0, n1 = 8190, 1000
    guard = 1 << 16  # doubles on either side of the output: no tile may write outside it
    whole = torch.empty(n0 * n1 * 6 + 2 * guard, dtype=torch.float64, device="cuda:0")
    d = whole[guard:guard + n0 * n1 * 6]

    def run():
        whole.fill_(-7.0)
        rep = rs.grid_eval(lib, "complete_analysis", p, None, n0, n1, ext, device=0,
                           out_device_ptr=d.data_ptr())  # fmt: skip
        assert rep["launches"] <= 3
        assert bool((whole[:guard] == -7.0).all()) and bool((whole[-guard:] == -7.0).all())
        got = d.cpu().numpy()
        assert not (got == -7.0).any()  # every record written
        return got.view(np.uint64).copy()

    monkeypatch.setenv("INFLATOX_RPT_TAIL", "0")
    uniform = run()
    monkeypatch.delenv("INFLATOX_RPT_TAIL")
    assert np.array_equal(run(), uniform)
    for rpt, tail in ((16, 4), (16, 1), (8, 3), (4, 2)):
        monkeypatch.setenv("INFLATOX_RPT", str(rpt))
        monkeypatch.setenv("INFLATOX_RPT_TAIL", str(tail))
        assert np.array_equal(run(), uniform), (rpt, tail)


@pytest.mark.parametrize("model", ["angular", "egno", "d5"])
def test_on_trajectory(model):
    """The trajectories the reference's own tests evaluate (tests/test_angular.py:79-83,
    test_egno.py:98-102, test_d5.py:168-170)."""
    lib = rs.open_inflx_dylib(cases.artifact(model).shared_object_path, False)
    lib.set_devices([0])
    orc, p = oracle.Oracle(model), cases.params(model)
    xs = cases.trajectory(model)
    out = np.zeros((xs.shape[0], 6))
    rs.complete_analysis_on_trajectory(lib, p, xs, out, False, 1)
    ref = orc.complete_analysis_on_trajectory(p, xs)
    for k in range(6):
        check(model, out[:, k], ref[:, k], what=f"ot[{k}]")
    for fn in ("consistency_only", "consistency_rapidturn_only", "epsilon_v_only"):
        o1 = np.zeros(xs.shape[0])
        getattr(rs, fn + "_on_trajectory")(lib, p, xs, o1, False, 1)
        check(model, o1, getattr(orc, fn + "_on_trajectory")(p, xs), what=fn)


def test_long_point_list_equals_short_batches():
    """2.3 M points (the chunked, double-buffered path of inflx_points_eval) against the same points
    sent in batches of 50 000 (one launch, plain copies): not a bit may differ, ragged tail included."""
    lib = rs.open_inflx_dylib(cases.artifact("angular").shared_object_path, False)
    lib.set_devices([0])
    p, ext = cases.params("angular"), cases.EXTENT["angular"]
    rng = np.random.default_rng(11)
    n = 2_300_017
    xs = np.ascontiguousarray(
        np.stack([rng.uniform(ext[0], ext[1], n), rng.uniform(ext[2], ext[3], n)], axis=1)
    )
    long = np.zeros((n, 6))
    rs.complete_analysis_on_trajectory(lib, p, xs, long, False, 1)
    for b in range(0, n, 500_000):  # spot-check five batches incl. the chunk boundaries
        for lo in (b, max(0, b + (1 << 20) - 25_000) if b + (1 << 20) < n else b):
            hi = min(n, lo + 50_000)
            short = np.zeros((hi - lo, 6))
            rs.complete_analysis_on_trajectory(lib, p, np.ascontiguousarray(xs[lo:hi]), short, False, 1)
            assert np.array_equal(long[lo:hi], short, equal_nan=True), (lo, hi)
    tail = np.zeros((17, 6))
    rs.complete_analysis_on_trajectory(lib, p, np.ascontiguousarray(xs[-17:]), tail, False, 1)
    assert np.array_equal(long[-17:], tail, equal_nan=True)
    one = np.zeros(n)
    rs.epsilon_v_only_on_trajectory(lib, p, xs, one, False, 1)
    ref = oracle.Oracle("angular").epsilon_v_only_on_trajectory(p, xs[:100_000])
    check("angular", one[:100_000], ref, what="long trajectory eps_V")


def test_two_streams_on_one_device_do_not_corrupt_each_other():
    """ADVICE (round 1): a device-resident call on a CALLER's stream returns without synchronising
    while its kernels still read the device's shared scratch and the module's __constant__ bank.
    Interleave calls with DIFFERENT parameter vectors on two streams and on the internal stream:
    every output must equal the one a synchronous call produces."""
    import torch

    lib = rs.open_inflx_dylib(cases.artifact("egno").shared_object_path, False)
    lib.set_devices([0])
    p0, ext = cases.params("egno"), cases.EXTENT["egno"]
    n0, n1 = 1536, 1024
    ps = [p0, p0 * np.array([1.5, 1.0, 1.0, 1.0]), p0 * np.array([1.0, 1.1, 1.0, 1.0])]
    want = []
    for p in ps:
        h = np.zeros((n0, n1, 6))
        rs.grid_eval(lib, "complete_analysis", p, h, n0, n1, ext)
        want.append(h)
    s1, s2 = torch.cuda.Stream(device=0), torch.cuda.Stream(device=0)
    outs = [torch.empty(n0 * n1 * 6, dtype=torch.float64, device="cuda:0") for _ in range(6)]
    order = [(0, s1), (1, s2), (2, s1), (0, s2), (1, None), (2, s2)]
    for (k, st), d in zip(order, outs):
        rs.grid_eval(lib, "complete_analysis", ps[k], None, n0, n1, ext, device=0,
                     out_device_ptr=d.data_ptr(), stream=None if st is None else st.cuda_stream)  # fmt: skip
    torch.cuda.synchronize()
    for (k, _), d in zip(order, outs):
        assert np.array_equal(d.cpu().numpy().reshape(n0, n1, 6), want[k], equal_nan=True), k
    del lib  # inflx_close after asynchronous work: must wait for it, not unload under it


def test_edge_shapes():
    lib = rs.open_inflx_dylib(cases.artifact("doc").shared_object_path, False)
    lib.set_devices([0])
    orc, p, ext = oracle.Oracle("doc"), cases.params("doc"), cases.EXTENT["doc"]
    for n0, n1 in [(1, 1), (1, 300), (300, 1), (5, 129), (0, 7), (7, 0)]:
        out = np.full((n0, n1, 6), 7.0)
        rs.complete_analysis(lib, p, out, ss_of(ext), False, 0)
        if n0 and n1:
            ref = orc.complete_analysis(p, n0, n1, ext)
            check("doc", out, ref, what=f"{n0}x{n1}")
    out = np.zeros((0, 6))
    rs.complete_analysis_on_trajectory(lib, p, np.zeros((0, 2)), out, False, 1)


@pytest.mark.parametrize("shape", [(3, 9_000_001), (2_100_001, 5)])
def test_extreme_aspect_ratios(shape):
    """Launch-grid limits: 70 313 column tiles in gridDim.x; 2.1 M rows = more row tiles than
    gridDim.y holds, so the engine must split the shard into several launches - device-resident
    (one request) and through the chunked host path.  Ragged in both directions."""
    import torch

    n0, n1 = shape
    lib = rs.open_inflx_dylib(cases.artifact("doc").shared_object_path, False)
    lib.set_devices([0])
    orc, p, ext = oracle.Oracle("doc"), cases.params("doc"), cases.EXTENT["doc"]
    ref = orc.consistency_only(p, n0, n1, ext)
    out = np.full((n0, n1), -7.0)
    rs.consistency_only(lib, p, out, ss_of(ext), False, 0)
    check("doc", out, ref, what=f"{n0}x{n1} host")
    d = torch.full((n0 * n1,), -7.0, dtype=torch.float64, device="cuda:0")
    rs.grid_eval(lib, "consistency_only", p, None, n0, n1, ext, device=0, out_device_ptr=d.data_ptr())
    dev = d.cpu().numpy().reshape(n0, n1)
    assert np.array_equal(dev.view(np.uint64), out.view(np.uint64))


def test_full_size_rows_of_baseline_grids():
    """At BASELINE.json's full sizes the oracle cannot sweep the grid in seconds, but any ROW of
    the full grid can be checked: coordinates come from global indices, so rows evaluated as
    shards of the 16384^2 (C3, C4) and 4096^2 (C2) grids must match the oracle's same rows."""
    for model, op, n in [("egno", "complete_analysis", 16384), ("d5", "complete_analysis", 16384),
                         ("angular", "consistency_only", 4096)]:  # fmt: skip
        lib = rs.open_inflx_dylib(cases.artifact(model).shared_object_path, False)
        lib.set_devices([0])
        orc, p, ext = oracle.Oracle(model), cases.params(model), cases.EXTENT[model]
        per = 6 if op == "complete_analysis" else 1
        for r in (0, n // 3 + 1, n - 2):
            got = np.zeros((2, n, per) if per > 1 else (2, n))
            rs.grid_eval(lib, op, p, got, n, n, ext, rows=(r, r + 2))
            ref = getattr(orc, op)(p, n, n, ext, rows=(r, r + 2))
            check(model, got, ref, what=f"{op} rows {r}..{r + 2} of {n}^2")


def test_full_size_c1_grid():
    """BASELINE C1 in full (hyperinflation, 1000 x 1000): every finite point within 1e-10."""
    lib = rs.open_inflx_dylib(cases.artifact("hyper").shared_object_path, False)
    lib.set_devices([0])
    p, ext = cases.params("hyper"), cases.EXTENT["hyper"]
    out = np.zeros((1000, 1000, 6))
    rs.complete_analysis(lib, p, out, ss_of(ext), False, 0)
    ref = oracle.Oracle("hyper").complete_analysis(p, 1000, 1000, ext)
    for k in range(6):
        check("hyper", out[..., k], ref[..., k], what=f"C1 plane {k}")


def test_facade_matches_reference_doc_test():
    """reference tests/test_doc.py:38-58 through the unchanged public API."""
    import inflatox_b200 as inflatox
    from inflatox_b200.consistency_conditions import GeneralisedAL

    anguelova = GeneralisedAL(cases.artifact("doc"))
    x, args = np.array([2.0, -2.0]), np.array([1.0])
    assert anguelova.calc_V(x, args) == 1.9166666666666667
    assert np.allclose(
        anguelova.calc_H(x, args), [[0.41206897, -1.05517241], [-1.05517241, -0.07873563]]
    )
    extent = (0.0, 2.5, 0.0, np.pi)
    consistency, ev, eh, eta, delta, omega = anguelova.complete_analysis(args, *extent)
    assert consistency.shape == (1000, 1000) and consistency.strides == (48000, 48)
    assert np.nanmax(consistency) <= 1
    assert isinstance(inflatox.__version__, str)


def test_basis_validation():
    import inflatox_b200 as ix

    # a sound basis passes (possibly with out-of-domain warnings, as upstream)
    cases.load_checked(lambda: rs.open_inflx_dylib(cases.artifact("angular").shared_object_path, True))
    # a basis whose first vector is not normalised must be refused (reference src/lib.rs:171-173)
    m = cases.load_model("doc")
    m.basis[0] = [2 * c for c in m.basis[0]]
    art = ix.Compiler(m, silent=True).compile()
    with pytest.raises(Exception, match="Expected basis vector 0 to be normalised"):
        rs.open_inflx_dylib(art.shared_object_path, True)


@pytest.mark.parametrize("model", ["angular", "d5"])  # d5: column pre-pass on every device
def test_multi_device_row_sharding_in_one_process(model):
    from inflatox_b200 import _native

    n_dev = _native.lib().inflx_device_count()
    if n_dev < 2:
        pytest.skip("needs >= 2 GPUs")
    lib = rs.open_inflx_dylib(cases.artifact(model).shared_object_path, False)
    p, ext = cases.params(model), cases.EXTENT[model]
    lib.set_devices([0])
    one = np.zeros((N0, N1, 6))
    rs.grid_eval(lib, "complete_analysis", p, one, N0, N1, ext)
    lib.set_devices(list(range(n_dev)))
    many = np.zeros((N0, N1, 6))
    rep = rs.grid_eval(lib, "complete_analysis", p, many, N0, N1, ext)
    assert rep["n_devices"] == n_dev
    assert np.array_equal(one, many, equal_nan=True)


@pytest.mark.parametrize("model", ["angular", "egno", "d5"])
def test_reference_test_flows_through_the_facade(model):
    """The call sequences of the reference's integration tests (tests/test_angular.py:62-86,
    test_egno.py:79-105, test_d5.py:143-173) through the unchanged public API: same arguments,
    same return shapes; plus the oracle on what they return."""
    from inflatox_b200.consistency_conditions import GeneralisedAL

    anguelova = cases.load_checked(lambda: GeneralisedAL(cases.artifact(model)))
    args, extent = cases.params(model), cases.EXTENT[model]
    N = 100
    orc = oracle.Oracle(model)
    v = anguelova.calc_V_array(args, [extent[0], extent[2]], [extent[1], extent[3]], [N, N])
    assert v.shape == (N, N)
    check(model, v, orc.potential_array(args, N, N, extent), what="calc_V_array")
    out = anguelova.complete_analysis(args, *extent, *[N, N])
    assert len(out) == 6 and all(o.shape == (N, N) for o in out)
    ref = orc.complete_analysis(args, N, N, extent)
    for k in range(6):
        check(model, out[k], ref[..., k], what=f"facade plane {k}")
    traj = cases.trajectory(model)
    ot = anguelova.complete_analysis_ot(args, traj)
    assert len(ot) == 6 and ot[0].shape == (traj.shape[0], 1)
    rt = anguelova.consistency_rapidturn(args, *extent, *[N, N])
    check(model, rt, orc.consistency_rapidturn_only(args, N, N, extent), what="rapidturn")
    assert anguelova.consistency(args, *extent, N, N).shape == (N, N)
    assert anguelova.epsilon_v(args, *extent, N, N).shape == (N, N)
    assert anguelova.consistency_ot(args, traj).shape == (traj.shape[0],)
    assert anguelova.epsilon_v_ot(args, traj).shape == (traj.shape[0],)
    assert anguelova.consistency_rapidturn_ot(args, traj).shape == (traj.shape[0],)
    flags = anguelova.flag_quantum_dif(args, *extent, N, N, accuracy=0.5)
    assert flags.dtype == bool and flags.shape == (N, N)
    h = anguelova.calc_H_array(args, [extent[0], extent[2]], [extent[1], extent[3]], [N, N])
    assert h.shape == (2, 2, N, N)
    # the reference's own signature (consistency_conditions.py:119-127): scalars per axis, N last
    h_ref_sig = anguelova.calc_H_array(args, extent[0], extent[1], extent[2], extent[3], [N, N])
    assert np.array_equal(h, h_ref_sig, equal_nan=True)
    with pytest.raises(TypeError):
        anguelova.calc_H_array(args, extent[0], extent[1])
    if model == "d5":
        # the reference walks the axes from `stop` outwards (src/lib.rs:250-258); for d5 that leaves
        # the model's domain, where w1 degenerates into v - the reference-generated C says the
        # same (oracle: <v, w1> = 1 at [58.5, 0]), so the reference raises BasisOth here as well
        with pytest.raises(Exception, match="to be orthogonal everywhere"):
            anguelova.validate_basis_on_domain(args, [extent[0], extent[2]], [extent[1], extent[3]], N=8)
        x = np.array([58.5, 0.0])
        v, w = orc.basis(0, x, args), orc.basis(1, x, args)
        assert abs(orc.inner_prod(x, args, v, w) - 1.0) < 1e-12
    else:
        anguelova.validate_basis_on_domain(args, [extent[0], extent[2]], [extent[1], extent[3]], N=8)
    sw = anguelova.sweep_complete_analysis(np.stack([args, args * 1.01]), *extent, 24, 40)
    assert sw.shape == (2, 24, 40, 6)
    assert np.array_equal(sw[0], np.stack(anguelova.complete_analysis(args, *extent, 24, 40), axis=-1),
                          equal_nan=True)
