"""Version constants.

`__abi_version__` is the model-artefact ABI of the reference this package is a drop-in for
(reference python/inflatox/version.py:22, src/lib.rs:50 `V_INFLX_ABI = 5.0.0`).  Artefacts carry it
in their `VERSION` global and `inflx_open` refuses an artefact whose major.minor differ
(reference src/inflatox_version.rs:48-53 compares major and minor only).
"""

__version__ = "0.1.0"
__abi_version__ = "5.0.0"
# Version of the *container* this package wraps around the cubin (see compiler.py / inflx_b200.h)
__container_version__ = 2  # 2: grid kernels take the rows-per-CTA count as a launch argument
