"""CPU tests of the drop-in boundary: libmvd.so loads, exports every symbol include/mvd.h
declares, its struct layouts match the ctypes mirror, and -- with no GPU -- refuses to compute
(there is no CPU fallback).  No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mvd.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mvd_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    names = declared_functions()
    for must in ("mvd_create", "mvd_destroy", "mvd_set_code", "mvd_set_states", "mvd_set_loglik", "mvd_learn_counts",
                 "mvd_detect", "mvd_trace", "mvd_last_error"):
        assert must in names
    text = open(HEADER).read()
    # every entry point cites the reference interface it replaces
    for cite in ("Pd_plotter.py:210-223", "Pd_plotter.py:158-163", "viterbi_markov.py:166-195", "viterbi_markov.py:118-132",
                 "Pd_plotter.py:149,212,219"):
        assert cite in text


def test_library_exports_every_declared_symbol():
    from mvd import _capi
    lib = _capi.load()
    names = declared_functions()
    assert set(names) == set(_capi.EXPORTS), "ctypes binding and header disagree"
    for name in names:
        assert getattr(lib, name) is not None
    out = subprocess.check_output(["nm", "-D", "--defined-only", _capi.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (mvd_[a-z0-9_]+)", out))
    assert set(names) <= exported
    header_abi = int(re.search(r"#define MVD_ABI_VERSION (\d+)", open(HEADER).read()).group(1))
    assert lib.mvd_abi_version() == header_abi == _capi.ABI_VERSION


def test_stale_library_is_rejected(monkeypatch):
    """A libmvd.so whose ABI version differs from the ctypes mirror's must not load (struct layouts would be wrong)."""
    from mvd import _capi
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "ABI_VERSION", _capi.ABI_VERSION + 1)
    with pytest.raises(ImportError) as ei:
        _capi.load()
    assert "ABI version" in str(ei.value)


def test_option_constants_match_header():
    """The ctypes mirror's MVD_OPT_* values are the header's (mvd_set_option)."""
    from mvd import _capi
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    opts = dict(re.findall(r"\bMVD_(OPT_[A-Z0-9_]+)\s*=\s*(\d+)", text))
    assert set(opts) == {"OPT_FORCE_GENERIC", "OPT_NO_PAIR", "OPT_LEARN_WARM", "OPT_NO_FSM1", "OPT_SPLIT", "OPT_NO_ANTIPODAL", "OPT_ASYNC_DETECT", "OPT_SPLIT_SEQUENTIAL", "OPT_SPLIT_CHUNK"}
    for name, value in opts.items():
        assert getattr(_capi, name) == int(value), name


def test_library_is_built_for_sm_100a():
    from mvd import _capi
    out = subprocess.run(["cuobjdump", "-lelf", _capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout


def test_library_does_not_link_the_oracle():
    from mvd import _capi
    out = subprocess.check_output(["ldd", _capi.LIB_PATH], text=True)
    assert "oracle" not in out
    syms = subprocess.check_output(["nm", "-D", _capi.LIB_PATH], text=True)
    assert "mvdo_" not in syms


def test_struct_layouts_match_header():
    """Compile a tiny C program against include/mvd.h and compare sizeof/offsetof with ctypes."""
    import tempfile
    from mvd import _capi
    fields_seg = ["N", "threshold", "stream", "table", "enc_taps", "decide", "random_input", "trial_begin", "trial_end",
                  "bits_offset"]
    fields_src = ["mode", "bits_on_device", "seed", "bits", "bits_words"]
    fields_par = ["N", "m", "n", "threshold", "stream", "decide", "enc_taps", "tmpl", "gamma", "trial_begin", "trial_end",
                  "bits_offset"]
    fields_bfs = ["S", "frontier", "iterations", "launches", "candidates", "closed", "max_metric", "ms", "reserved"]
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "mvd.h"', "int main(void){",
            'printf("%zu %zu\\n", sizeof(mvd_segment), sizeof(mvd_src));']
    for f in fields_seg:
        prog.append(f'printf("%zu\\n", offsetof(mvd_segment, {f}));')
    for f in fields_src:
        prog.append(f'printf("%zu\\n", offsetof(mvd_src, {f}));')
    prog.append('printf("%zu %zu\\n", sizeof(mvd_parity_segment), sizeof(mvd_bfs_stats));')
    for f in fields_par:
        prog.append(f'printf("%zu\\n", offsetof(mvd_parity_segment, {f}));')
    for f in fields_bfs:
        prog.append(f'printf("%zu\\n", offsetof(mvd_bfs_stats, {f}));')
    prog.append("return 0;}")
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        open(src, "w").write("\n".join(prog))
        exe = os.path.join(td, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, src])
        out = subprocess.check_output([exe], text=True).split()
    vals = list(map(int, out))
    assert vals[0] == C.sizeof(_capi.Segment) and vals[1] == C.sizeof(_capi.Src)
    got_seg = [getattr(_capi.Segment, f).offset for f in fields_seg]
    got_src = [getattr(_capi.Src, f).offset for f in fields_src]
    assert vals[2:2 + len(fields_seg)] == got_seg
    base = 2 + len(fields_seg) + len(fields_src)
    assert vals[2 + len(fields_seg):base] == got_src
    assert vals[base] == C.sizeof(_capi.ParitySegment) and vals[base + 1] == C.sizeof(_capi.BfsStats)
    assert vals[base + 2:base + 2 + len(fields_par)] == [getattr(_capi.ParitySegment, f).offset for f in fields_par]
    assert vals[base + 2 + len(fields_par):] == [getattr(_capi.BfsStats, f).offset for f in fields_bfs]


def test_create_fails_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mvd import _capi
    lib = _capi.load()
    ctx = C.c_void_p()
    rc = lib.mvd_create(C.byref(ctx), 0)
    assert rc == -2 and not ctx.value                      # MVD_E_CUDA
    msg = lib.mvd_last_error(None).decode()
    assert "no CPU fallback" in msg
    # NULL contexts are rejected, not dereferenced
    assert lib.mvd_destroy(None) == 0
    assert lib.mvd_synchronize(None) == -1
    f = C.c_float()
    assert lib.mvd_last_kernel_ms(None, C.byref(f)) == -1


def test_missing_library_is_an_import_error(monkeypatch):
    from mvd import _capi
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", "/nonexistent/libmvd.so")
    with pytest.raises(ImportError) as ei:
        _capi.load()
    assert "no CPU fallback" in str(ei.value)


def test_numpy_segment_records_match_the_ctypes_struct():
    """engine.SEG_DTYPE (column-wise filling of a sweep's segments) has mvd_segment's layout, and
    Detector.segment_array fills the same bytes as the per-field ctypes path."""
    import numpy as np
    from mvd import _capi, engine
    assert engine.SEG_DTYPE.itemsize == C.sizeof(_capi.Segment)
    for name in engine.SEG_DTYPE.names:
        assert engine.SEG_DTYPE.fields[name][1] == getattr(_capi.Segment, name).offset, name
    fake = engine.Detector.__new__(engine.Detector)          # no device needed for the two packers
    fake.dec_taps = [7, 5]
    segs = [engine.Seg(N=500 + i, threshold=1000 * i + 7, stream=2 * i + 1, table=i % 3, enc_taps=[7 - i % 2, 5], decide=i % 2,
                       random_input=True, trial_begin=10, trial_end=10 + 3 * i, bits_offset=0) for i in range(6)]
    a = fake._segments(segs)
    b = fake.segment_array(N=[s.N for s in segs], threshold=[s.threshold for s in segs], stream=[s.stream for s in segs],
                           table=[s.table for s in segs], enc_taps=[list(s.enc_taps) for s in segs], decide=[s.decide for s in segs],
                           trial_begin=10, trial_end=[s.trial_end for s in segs])
    assert bytes(a) == b.tobytes()
    fake.ctx = None                                          # __del__ must not touch a context that was never created
