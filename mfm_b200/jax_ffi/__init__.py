"""jax.ffi registration of the C ABI (include/mfm_b200.h) — the binding `north_star` names for a JAX host.

This image has neither JAX nor the XLA FFI headers, so nothing here runs in the repo's tests: the module is
shipped as SOURCE for a maintainer of the JAX reference.  With JAX installed, `register()` compiles `mfm_jax_ffi.cc` against
`jax.ffi.include_dir()`, loads it and registers one custom-call target per handler; `mala_step` below shows the call the
reference's `jax.vmap(kernel)` (exe_flow_matching.py:313) turns into.  The ctypes-over-torch binding in `mfm_b200/_lib.py`
is what the tests and the benchmark exercise; both bind the same symbols of the same library."""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_SO = os.path.join(_HERE, "libmfm_jax_ffi.so")
TARGETS = {"mfm_init": "MfmInit", "mfm_mala_step": "MfmMalaStep", "mfm_ode_flow": "MfmOdeFlow", "mfm_flow_step": "MfmFlowStep",
           "mfm_fm_loss_grad": "MfmFmLossGrad", "mfm_adamw": "MfmAdamW"}

try:
    import jax
    import jax.numpy as jnp
    HAVE_JAX = True
except ImportError:          # this image: documented, not an error
    jax = jnp = None
    HAVE_JAX = False


def build() -> str:
    """g++ the handler file against jaxlib's FFI headers and libmfm_b200.so."""
    if not HAVE_JAX:
        raise RuntimeError("jax is not installed: the jax.ffi shim cannot be built here (use the ctypes binding, mfm_b200._lib)")
    inc = jax.ffi.include_dir()
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", inc, "-I", os.path.join(_ROOT, "include"),
           "-I", "/usr/local/cuda/include", os.path.join(_HERE, "mfm_jax_ffi.cc"), "-L", os.path.dirname(_HERE), "-lmfm_b200",
           "-Wl,-rpath," + os.path.dirname(_HERE), "-o", _SO]
    subprocess.check_call(cmd)
    return _SO


def register():
    """Register every handler as a CUDA custom-call target; returns the loaded library."""
    if not os.path.exists(_SO):
        build()
    lib = ctypes.CDLL(_SO)
    for name, sym in TARGETS.items():
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, sym)), platform="CUDA")
    return lib


def mala_step(rng_key, state, target_blob, workspace, step_size, chain_offset=0, n_total=0):
    """Replacement of `jax.vmap(lambda k, s: kernel(k, s, logprob, step_size))(split(key, N), states)` (exe_flow_matching.py:303,313):
    `rng_key` is the UNSPLIT uint32[2] key (rows of split(key, n_total) are derived on the device) or the uint32[N,2] keys.
    `target_blob`: uint8 array holding the packed mfm_target_t (mfm_b200.distributions.Distribution._desc, bytes())."""
    n, d = state.position.shape
    f32, out = jnp.float32, jax.ShapeDtypeStruct
    out_types = (out((n, d), f32), out((n,), f32), out((n, d), f32), out((n,), f32), out((n,), jnp.uint8), out((n, d), f32), out((n,), f32))
    x, l, g, acc, flag, prop, w = jax.ffi.ffi_call("mfm_mala_step", out_types)(
        target_blob, rng_key, state.position, state.logdensity, state.logdensity_grad, workspace,
        step_size=jnp.float32(step_size), chain_offset=jnp.int32(chain_offset), n_total=jnp.int32(n_total))
    return type(state)(x, l, g), (acc, flag.astype(bool), prop, w)
