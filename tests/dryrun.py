"""Host-logic dry run (test infrastructure): replaces the ctypes call with an argument validator so
the whole Python control flow (shapes, dtypes, autograd wiring, argument counts and ctypes
convertibility) can be exercised on a machine without a GPU.  No arithmetic happens, outputs are
uninitialised memory; this is NOT a CPU fallback and lives only under tests/."""
import contextlib
import ctypes as C

import torch

import vit3d_b200
from vit3d_b200 import _lib, functional as F


@contextlib.contextmanager
def dry_run():
    calls = []

    def fake_call(name, *args):
        res, argtypes = _lib.SIGNATURES[name]
        assert len(args) == len(argtypes), f"{name}: {len(args)} args, signature has {len(argtypes)}"
        for i, (a, t) in enumerate(zip(args, argtypes)):
            try:
                t.from_param(a) if a is not None else None
            except Exception as e:  # pragma: no cover
                raise AssertionError(f"{name}: arg {i} = {a!r} not convertible to {t}: {e}")
            if a is None:
                assert t is C.c_void_p, f"{name}: arg {i} is None but not a pointer"
        calls.append(name)

    def fake_ptr(t):
        return None if t is None else t.data_ptr()

    class FakeLib:
        def __getattr__(self, n):
            if n == "vit3d_patch_embed_ws_bytes":
                return lambda B, X, Y, Z, p0, p1, p2, H, prec: 4 * (B * (X // p0) * (Y // p1) * (Z // p2) * (p0 * p1 * p2 + H)) + 256
            if n in ("vit3d_tc_supported", "vit3d_attn_padded_supported", "vit3d_train_supported", "vit3d_mlp_bwd_supported"):
                return lambda *a: 0
            raise AttributeError(n)

    saved = (_lib.call, _lib.ptr, _lib.stream, _lib.lib, F.call, F.ptr, F.stream, F._need_cuda)
    _lib.call, _lib.ptr, _lib.stream, _lib.lib = fake_call, fake_ptr, (lambda: None), (lambda: FakeLib())
    F.call, F.ptr, F.stream, F._need_cuda = fake_call, fake_ptr, (lambda: None), (lambda *a: None)
    try:
        yield calls
    finally:
        (_lib.call, _lib.ptr, _lib.stream, _lib.lib, F.call, F.ptr, F.stream, F._need_cuda) = saved
