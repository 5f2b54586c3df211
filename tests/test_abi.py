"""The C-ABI library loads on a box without a GPU and exports exactly what include/brdae.h
declares; struct layouts of the ctypes mirror match the header (checked with the C compiler);
calls that need no device behave (lookup, defaults, argument validation)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

import dae_scipy_b200 as br
from dae_scipy_b200 import _abi
from _util import ROOT

HEADER = os.path.join(ROOT, "include", "brdae.h")


def _lib():
    if not os.path.exists(_abi.LIB_PATH):
        from dae_scipy_b200 import build
        build.build()
    return _abi.load()


def test_library_exports_every_declared_symbol():
    lib = _lib()
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(brdae_[a-z0-9_]+)\s*\(", text))
    assert declared == set(_abi.EXPORTED_SYMBOLS), declared ^ set(_abi.EXPORTED_SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert lib.brdae_abi_version() == 4


def test_ctypes_struct_layout_matches_header():
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "brdae.h"
    int main(void) {
        printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(brdae_options), offsetof(brdae_options, max_newton_ite),
               offsetof(brdae_options, nmax_step), sizeof(brdae_batch), offsetof(brdae_batch, y0),
               offsetof(brdae_batch, rtol), offsetof(brdae_batch, y_eval), offsetof(brdae_batch, counters),
               offsetof(brdae_batch, dense_q), offsetof(brdae_batch, n_events), offsetof(brdae_batch, event_coef),
               offsetof(brdae_batch, event_y));
        return 0;
    }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o",
                        os.path.join(d, "t")], check=True)
        got = [int(x) for x in subprocess.run([os.path.join(d, "t")], capture_output=True, text=True).stdout.split()]
    O, Bt = _abi.BrdaeOptions, _abi.BrdaeBatch
    want = [C.sizeof(O), O.max_newton_ite.offset, O.nmax_step.offset, C.sizeof(Bt), Bt.y0.offset, Bt.rtol.offset,
            Bt.y_eval.offset, Bt.counters.offset, Bt.dense_q.offset, Bt.n_events.offset, Bt.event_coef.offset,
            Bt.event_y.offset]
    assert got == want


def test_problem_lookup_and_defaults():
    lib = _lib()
    for name, prob in br.PROBLEMS.items():
        if prob.lib_path:                  # a user problem registered by another test: lives in its own plugin build
            continue
        pid, n, npar = C.c_int32(), C.c_int32(), C.c_int32()
        assert lib.brdae_problem_lookup(name.encode(), C.byref(pid), C.byref(n), C.byref(npar)) == 0
        assert (pid.value, n.value, npar.value) == (prob.problem_id, prob.n, prob.n_params)
    assert lib.brdae_problem_lookup(b"nope", None, None, None) != 0
    o = _abi.BrdaeOptions()
    lib.brdae_default_options(C.byref(o))
    d = _abi.make_options()
    for f, _ in _abi.BrdaeOptions._fields_:
        assert getattr(o, f) == getattr(d, f), f


def test_bad_arguments_are_rejected_without_a_device():
    lib = _lib()
    ens = br.ensembles.pendulum_ensemble(4)
    hb = br.prepare(ens["problem"], ens["t_span"], ens["y0"].shape, t_eval=ens["t_eval"], **ens["options"])
    b = hb.struct({})                      # all batch pointers NULL
    assert lib.brdae_solve(C.byref(b), C.byref(hb.options), None, 0, None) == 1   # BRDAE_ERR_BAD_ARGUMENT
    b.problem = 99
    assert lib.brdae_solve(C.byref(b), C.byref(hb.options), None, 0, None) == 2   # unknown problem
    assert lib.brdae_strerror(2) == b"unknown problem id"
    b = hb.struct({})
    b.n_systems = 0                        # empty ensemble: nothing to do, no device needed
    assert lib.brdae_solve(C.byref(b), C.byref(hb.options), None, 0, None) == 0


def test_product_path_has_no_cpu_fallback(monkeypatch):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ens = br.ensembles.pendulum_ensemble(4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        br.solve_ivp_batched(ens["problem"], ens["t_span"], ens["y0"], ens["params"], t_eval=ens["t_eval"], **ens["options"])
    monkeypatch.setattr(_abi, "LIB_PATH", "/nonexistent/libbrdae.so")
    monkeypatch.setattr(_abi, "_lib", None)
    with pytest.raises(RuntimeError, match="missing"):
        _abi.load()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dae_scipy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "from oracle" not in text and "import oracle" not in text and "oracle/" not in text.replace("oracle/ ", ""), f


def test_integration_md_stub_matches_the_abi():
    """The ctypes stub INTEGRATION.md tells a maintainer of the reference to add is executed here (definitions only):
    its structs must have the layout of include/brdae.h, i.e. of the package's own mirror."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = text.split("```python", 1)[1].split("```", 1)[0]
    code = code.replace('C.CDLL("libbrdae.so")', f'C.CDLL({_abi.LIB_PATH!r})')
    _lib()
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    Options, Batch = ns["Options"], ns["Batch"]
    assert C.sizeof(Options) == C.sizeof(_abi.BrdaeOptions) and C.sizeof(Batch) == C.sizeof(_abi.BrdaeBatch)
    assert [f[0] for f in Options._fields_] == [f[0] for f in _abi.BrdaeOptions._fields_]
    assert [(f[0], getattr(Batch, f[0]).offset) for f in Batch._fields_] == \
           [(f[0], getattr(_abi.BrdaeBatch, f[0]).offset) for f in _abi.BrdaeBatch._fields_]
    assert callable(ns["solve_ivp_ensemble"]) and ns["_lib"].brdae_abi_version() == _abi.ABI_VERSION


def test_error_codes_of_the_header_match_the_python_side():
    """include/brdae.h is the contract: the codes the host side interprets are the header's."""
    import re
    text = open(os.path.join(ROOT, "include", "brdae.h")).read()
    codes = {m.group(1): int(m.group(2)) for m in re.finditer(r"(BRDAE_(?:OK|ERR_[A-Z0-9_]+))\s*=\s*(\d+)", text)}
    assert codes["BRDAE_OK"] == 0 and codes["BRDAE_ERR_NONFINITE_Y0"] == _abi.ERR_NONFINITE_Y0 == 6
    lib = _abi.load()
    lib.brdae_strerror.restype = C.c_char_p
    for name, code in codes.items():
        assert lib.brdae_strerror(code).decode() not in ("", "unknown error"), name
