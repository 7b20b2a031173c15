"""inflatox_b200 - B200-native back-end for inflatox's grid-evaluation hot path.

`import inflatox_b200 as inflatox` gives the names the reference package exports for this path
(reference python/inflatox/__init__.py:20-40): `Compiler`, `CompilationArtifact`,
`InflationModel`, `consistency_conditions` (`InflationCondition`, `GeneralisedAL`), `log_info`,
`log_warn`.  The symbolic model builder (`InflationModelBuilder`, reference symbolic.py) is
upstream of the hot path and is NOT re-implemented: build models with the reference package and
hand them to `inflatox_b200.Compiler` (the `inflatox` overlay assembled by
`__graft_entry__.build()` does exactly that under the reference's own package name).
"""
from .version import __abi_version__, __version__
from .model import InflationModel
from .compiler import CompilationArtifact, Compiler, UnsupportedFunctionError
from . import consistency_conditions
from .libinflx_rs import log_info, log_warn

__all__ = [
    "CompilationArtifact",
    "Compiler",
    "InflationModel",
    "UnsupportedFunctionError",
    "consistency_conditions",
    "log_info",
    "log_warn",
    "__version__",
    "__abi_version__",
]


def __getattr__(name):
    # `InflationModelBuilder` (reference python/inflatox/symbolic.py) is upstream of the hot path
    # and not re-implemented; hand out the reference's when that package is installed.
    if name == "InflationModelBuilder":
        try:
            from inflatox import InflationModelBuilder  # type: ignore
        except ImportError as e:
            raise ImportError(
                "InflationModelBuilder is the reference package's symbolic front-end "
                "(`pip install inflatox`); inflatox_b200 accelerates the numerical path and takes "
                "its InflationModel objects as input"
            ) from e
        return InflationModelBuilder
    raise AttributeError(name)
