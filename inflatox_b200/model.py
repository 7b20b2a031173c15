"""`InflationModel`: the symbolic description a `Compiler` consumes.

This is the data format on the *input* side of the hot path.  It mirrors the attribute names of
the reference's container (reference python/inflatox/symbolic.py:30-88) so that an object built by
the reference's unchanged `InflationModelBuilder` can be handed to `inflatox_b200.Compiler`
directly (duck typing: only the attributes below are read).  The symbolic derivation itself
(Christoffels, covariant Hesse, vielbein projection; symbolic.py:287-726) is upstream of the hot
path and out of scope: models are produced by the reference package (the
`inflatox` overlay `__graft_entry__.build()` assembles runs the reference's unmodified
`symbolic.py` on top of this back-end).  The test fixtures under tests/golden/models are such
models, pickled; their loader lives in tests/cases.py, not here.
"""
from __future__ import annotations

import sympy


class InflationModel:
    def __init__(
        self,
        model_name: str,
        coordinates: list[sympy.Symbol],
        tangents: list[sympy.Symbol],
        basis: list[list[sympy.Expr]],
        eom_fields: list[sympy.Expr],
        eom_h: sympy.Expr,
        eom_hdot: sympy.Expr,
        potential: sympy.Expr,
        metric: list[list[sympy.Expr]],
        gradient_square: sympy.Expr,
        hesse_cmp: list[list[sympy.Expr]],
    ):
        self.model_name = model_name
        self.coordinates = coordinates
        self.coordinate_tangents = tangents
        self.dim = len(coordinates)
        self.basis = basis
        self.eom_fields = eom_fields
        self.eom_h = eom_h
        self.eom_hdot = eom_hdot
        self.potential = potential
        self.metric = metric
        self.gradient_square = gradient_square
        self.hesse_cmp = hesse_cmp
        # same consistency checks (and failure mode: plain Exception) as symbolic.py:71-88
        if any(len(row) != len(hesse_cmp) for row in hesse_cmp):
            raise Exception("The Hesse matrix is square; the provided list was not")
        if any(len(row) != len(metric) for row in metric):
            raise Exception("The metric tensor is square; the provided list was not")
        if len(hesse_cmp) != len(basis[0]):
            raise Exception("The provided Hesse Matrix and basis are of different dimensionality")
        if len(basis) != self.dim:
            raise Exception("The dimension of the provided basis does not match the number of fields.")
        if len(tangents) != self.dim:
            raise Exception(
                "The number of coordinate symbols does not match the number of tangent symbols."
            )

    # -- fixtures ----------------------------------------------------------------------------
    FIELDS = (
        "model_name", "coordinates", "tangents", "basis", "eom_fields", "eom_h", "eom_hdot",
        "potential", "metric", "gradient_square", "hesse_cmp",
    )  # fmt: skip

    def __str__(self):
        return (
            f"[Inflatox Inflation Model]\nmodel name: {self.model_name}\n"
            f"dimensionality: {self.dim} field(s)\ncoordinates: {list(self.coordinates)}\n"
            f"potential: {self.potential}\n"
        )
