"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's attention-sampling arithmetic (plain C in ``*_oracle.c``, driven through
ctypes/numpy) plus ``_ref/`` = the reference's own CUDA kernel translation units compiled unmodified (GPU-side A/B).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this package — and there only as the checker / the baseline being reported, never as the thing shipped. The product
package ``bevformer_tensorrt_b200`` must not import it (tests/test_host_logic.py enforces that).
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(_HERE, "_build", "liboracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libref_kernels.so")


def build(ref: bool = True) -> None:
    """Compile liboracle.so (gcc) and, when /root/reference is present, _ref/libref_kernels.so (nvcc)."""
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


_oracle_lib = None


def lib() -> ctypes.CDLL:
    global _oracle_lib
    if _oracle_lib is None:
        srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith("_oracle.c")]
        stale = (not os.path.exists(ORACLE_LIB)) or any(
            os.path.getmtime(s) > os.path.getmtime(ORACLE_LIB) for s in srcs
        )
        if stale:
            build(ref=False)
        _oracle_lib = ctypes.CDLL(ORACLE_LIB)
    return _oracle_lib
