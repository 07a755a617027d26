"""CPU-side checks: the C-ABI library loads and exports every symbol include/mfit_cuda.h declares, fails loudly
without a GPU (no CPU fallback), and the host-side mirror normalises inputs like walker-create does."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from lisp_mcmc_b200 import _lib, families, shard_chains, shard_datasets
from lisp_mcmc_b200.walker import clean_data, clean_data_error, force_list, get_depth, plist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "mfit_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mfit_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _header_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), "libmfit_cuda.so does not export %s" % n
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree: %s" % (set(names) ^ set(_lib.SIGNATURES))
    assert L.mfit_abi_version() == 1


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.AdaptCfg) == 40 and C.sizeof(_lib.ChainStats) == 64


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = _lib.lib().mfit_ctx_create(0, None, C.byref(h))
    assert rc == _lib.ERR_CUDA and not h.value
    assert b"no CPU fallback" in _lib.lib().mfit_last_error(None)
    from lisp_mcmc_b200 import walker_create, builtin
    with pytest.raises(_lib.MfitError):
        walker_create(function=builtin("poly", degree=1), data=[[0.0, 1.0], [0.0, 1.0]], params=[("c0", 0.0), ("c1", 1.0)])


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "lisp_mcmc_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".lisp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liborc" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_lisp_cffi_bindings_match_the_header():
    """lisp/mfit-cuda.lisp cannot be executed in this image (no Common Lisp): at least every defcfun must name a
    symbol the header declares, with the declared number of arguments and 64-bit integers where the ABI has them."""
    txt = open(os.path.join(ROOT, "lisp_mcmc_b200", "lisp", "mfit-cuda.lisp")).read()
    txt = re.sub(r";[^\n]*", "", txt)
    forms = re.findall(r'\(cffi:defcfun \("(mfit_[a-z0-9_]+)" [^)]*\) (:[a-z0-9]+)((?:\s*\([a-z0-9-]+ :[a-z0-9]+\))*)\)', txt)
    assert len(forms) >= 25
    lisp_t = {C.c_int: ":int", C.c_int64: ":int64", C.c_uint64: ":uint64", C.c_double: ":double", C.c_char_p: ":string"}
    for name, ret, args in forms:
        assert name in _lib.SIGNATURES, name
        res, argtypes = _lib.SIGNATURES[name]
        got = re.findall(r"\([a-z0-9-]+ (:[a-z0-9]+)\)", args)
        assert len(got) == len(argtypes), "%s: %d lisp args, header has %d" % (name, len(got), len(argtypes))
        for g, a in zip(got, argtypes):
            ok = (":string", ":pointer") if a is C.c_char_p else (lisp_t.get(a, ":pointer"),)   # char* out-buffers
            assert g in ok, "%s: %s vs %s" % (name, g, a)
        assert ret == {None: ":void", C.c_char_p: ":string", C.c_int64: ":int64"}.get(res, ":int"), name
    # the two structs passed by pointer
    for struct, cls in (("adapt-cfg", _lib.AdaptCfg), ("chain-stats", _lib.ChainStats)):
        body = re.search(r"\(cffi:defcstruct %s(.*?)\)\)" % struct, txt, flags=re.S).group(1)
        types = re.findall(r"\([a-z0-9-]+ (:[a-z0-9]+)\)", body + ")")
        want = [{C.c_int64: ":int64", C.c_int32: ":int32", C.c_double: ":double"}[t] for _, t in cls._fields_]
        assert types == want, struct


def test_get_depth_and_force_list_docstring_examples():
    # get-depth :763-764, force-list :755-759
    assert get_depth([1, 2, 3]) == 1 and get_depth([[1, 2], [3, 4]]) == 2 and get_depth([[[1]]]) == 3
    assert force_list(3) == [3] and force_list([3, 4]) == [3, 4]


def test_clean_data_shapes_and_errors():
    d = clean_data([[1, 2, 3], [4, 5, 6]], 1)
    assert len(d) == 1 and d[0][0].dtype == np.float64 and list(d[0][1]) == [4.0, 5.0, 6.0]
    d2 = clean_data([[[1, 2], [3, 4]], [[5, 6, 7], [8, 9, 10]]], 2)
    assert len(d2) == 2 and len(d2[1][0]) == 3
    with pytest.raises(ValueError):
        clean_data([[1, 2, 3], [4, 5, 6]], 2)        # one data set for two functions (:819)
    with pytest.raises(ValueError):
        clean_data([[1, 2, 3], [4, 5]], 1)
    with pytest.raises(ValueError):
        clean_data([1, 2, 3], 1)


def test_clean_data_keeps_sequences_as_x():
    """multi-dimensional inputs (:1137): each x is itself a sequence; clean-data :807-825 takes such a tree as it is when
    its length equals the number of functions (a single (x y) set of depth 3 is therefore an error for one function --
    the reference asks for (list (list xs ys)))"""
    xs = [[0.0, 1.0], [0.5, 2.0], [1.0, 4.0]]
    d = clean_data([[xs, [1, 2, 3]]], 1)
    assert d[0][0].shape == (3, 2) and d[0][1].shape == (3,)
    with pytest.raises(ValueError):
        clean_data([xs, [1, 2, 3]], 1)
    with pytest.raises(ValueError):
        clean_data([[xs, [1, 2]]], 1)


def test_clean_data_error_broadcasts_first_value():
    data = clean_data([[[1, 2, 3], [4, 5, 6]], [[1, 2], [3, 4]]], 2)
    e = clean_data_error(0.2, data)                    # scalar -> every point (:776-777)
    assert [len(v) for v in e] == [3, 2] and np.all(e[0] == 0.2) and np.all(e[1] == 0.2)
    e = clean_data_error(None, data)                   # nil -> 1 (:1144)
    assert np.all(e[0] == 1.0)
    e = clean_data_error([0.1, 0.3], data)             # flat list: fits the 2-point set, first value elsewhere
    assert np.all(e[0] == 0.1) and list(e[1]) == [0.1, 0.3]
    e = clean_data_error([[0.1], [0.4]], data)         # docstring example :776-777: structure differs -> first value
    assert np.all(e[0] == 0.1) and np.all(e[1] == 0.1)
    e = clean_data_error([[0.1, 0.2, 0.3], [0.5, 0.6]], data)   # full per-point lists are kept
    assert list(e[0]) == [0.1, 0.2, 0.3] and list(e[1]) == [0.5, 0.6]
    one = clean_data([[1, 2, 3], [4, 5, 6]], 1)
    assert np.all(clean_data_error([0.7, 0.9], one)[0] == 0.7)   # short list: the FIRST value is used (:794-805)


def test_plist_forms():
    assert plist([":b", -1, ":m", 2])[0] == ["b", "m"]
    k, v = plist({"B": -1, "m": 2})
    assert k == ["b", "m"] and v.dtype == np.float64 and list(v) == [-1.0, 2.0]
    with pytest.raises(ValueError):
        plist([":a", 1, ":a", 2])
    with pytest.raises(ValueError):
        plist([":a", 1, ":b"])


def test_shard_plans():
    parts = [shard_chains(4099, r, 8) for r in range(8)]
    assert parts[0][0] == 0 and sum(c for _, c in parts) == 4099
    assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(7))
    assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    assert shard_datasets(8, 8) == list(range(8)) and shard_datasets(5, 2) == [0, 1, 0, 1, 0]


def test_family_texts_parse_and_lower():
    for name, deg in [("poly", 0), ("poly", 7), ("expo", 0), ("sinus", 0), ("lorentz", 0), ("gauss_lin", 0), ("nv_multi", 3)]:
        keys = families.slots(name, deg)
        code = _lib.lower_sexpr(families.lisp_text(name, deg), keys)
        assert "struct MfitUser" in code and "eval(double x, G& gd)" in code


def test_cl_reader_numbers_and_file_reader():
    """the Common Lisp reader quirk the input side must reproduce: `2.003E+3` is a single-float, `2.003d3` a double,
    integers and ratios are exact; read-file->data / file->file-specs / create-walker-data on a TSV with a header line
    and CRLF line ends (an excerpt of the reference's example-data.xls)"""
    from lisp_mcmc_b200 import io as mio
    assert mio.cl_read_number("2.003E+3") == 2003.0
    assert mio.cl_read_number("-4.172E-7") == float(np.float32(-4.172e-7)) != -4.172e-7
    assert mio.cl_read_number("-4.172d-7") == -4.172e-7
    assert mio.cl_read_number("1e0") == 1.0 and mio.cl_read_number("7") == 7.0 and mio.cl_read_number("3/4") == 0.75
    assert mio.cl_read_number("0.1") == float(np.float32(0.1)) and mio.cl_read_number("0.1d0") == 0.1
    assert mio.cl_read_number("Probe") is None and mio.cl_read_number("") is None
    path = os.path.join(ROOT, "tests", "golden", "example_data_excerpt.tsv")
    specs = mio.file_specs(path)
    assert specs["header-lines"] == 1 and specs["data-length"] == 9 and specs["data-rows"] == 6 and specs["num-pages"] == 1
    data = mio.read_file_to_data(path)
    assert len(data) == 9 and all(len(c) == 6 for c in data)
    assert data[0][0] == 2000.0 and data[0][1] == 2003.0
    assert data[2][0] == float(np.float32(-4.172e-7))
    x, y = mio.create_walker_data(data, 0, 3)
    assert np.array_equal(x, data[0]) and np.array_equal(y, data[3])
    rows = mio.read_file_to_data(path, transpose=False)
    assert len(rows) == 6 and len(rows[0]) == 9 and rows[1][0] == 2003.0


def test_file_reader_pages(tmp_path):
    """pages = blocks of data rows separated by an empty line (read-file->data :pages, :1443-1450)"""
    from lisp_mcmc_b200 import io as mio
    p = tmp_path / "pages.csv"
    p.write_text("x;y\n1;1.5\n2;2.5\n3;3.5\n\n4;4.5\n5;5.5\n6;6.5\n")
    specs = mio.file_specs(str(p), delim=";")
    assert specs["header-lines"] == 1 and specs["data-rows"] == 3 and specs["num-pages"] == 2
    pages = mio.read_file_to_data(str(p), delim=";", pages=True)
    assert len(pages) == 2 and np.array_equal(pages[1][0], [4.0, 5.0, 6.0]) and np.array_equal(pages[0][1], [1.5, 2.5, 3.5])


def test_reference_function_names_are_exported():
    """the host mirror keeps the reference's exported names for the path (mcmc-fitting.lisp `export` forms) -- including
    the walker-set forms :1003-1036, which a walker SET (W chains in one object) serves directly"""
    import lisp_mcmc_b200 as m
    for name in ["walker_create", "walker_adaptive_steps", "walker_adaptive_steps_full", "walker_many_steps",
                 "walker_sample_region", "walker_force_take_step", "walker_get", "walker_save", "walker_load", "mcmc_fit",
                 "walker_set_create", "walker_set_get", "walker_set_save", "walker_set_load", "walker_set_delete", "linspace"]:
        assert callable(getattr(m, name)), name
    for meth in ["data_and_fit", "modify", "keep_walks", "restore", "analytics", "make_step", "trace", "stats"]:
        assert callable(getattr(m.Walker, meth)), meth


def test_lisp_host_file_is_balanced():
    """no Common Lisp in the image: at least the reader must be able to get through lisp/mfit-cuda.lisp -- parentheses
    balance outside strings, comments and character literals, and every top-level form starts at depth 0"""
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lisp_mcmc_b200", "lisp", "mfit-cuda.lisp")
    s = open(path).read()
    depth, i, n, forms = 0, 0, len(s), 0
    while i < n:
        c = s[i]
        if c == ";":
            while i < n and s[i] != "\n":
                i += 1
            continue
        if c == "#" and s[i + 1:i + 2] == "|":
            i = s.index("|#", i + 2) + 2
            continue
        if c == '"':
            i += 1
            while s[i] != '"':
                i += 2 if s[i] == "\\" else 1
            i += 1
            continue
        if c == "#" and s[i + 1:i + 2] == "\\":
            i += 3
            continue
        if c == "(":
            forms += depth == 0
            depth += 1
        elif c == ")":
            depth -= 1
            assert depth >= 0, "unbalanced ) near offset %d" % i
        i += 1
    assert depth == 0 and forms > 50


def test_adaptive_defaults_match_the_reference_lambda_list():
    """walker-adaptive-steps-full (mcmc-fitting.lisp:862): (n 100000) (temperature 1d3) (auto (or :prob-settle
    :slope-settle nil)) -> :prob-settle.  The Python mirror and the Lisp host must agree with it and with each other."""
    import inspect
    from lisp_mcmc_b200.walker import Walker
    sig = inspect.signature(Walker.adaptive_steps_full).parameters
    assert sig["n"].default == 100000 and sig["temperature"].default == 1e3 and sig["auto"].default == "prob-settle"
    txt = open(os.path.join(ROOT, "lisp_mcmc_b200", "lisp", "mfit-cuda.lisp")).read()
    m = re.search(r"\(defun walker-adaptive-steps-full \(walker &key \(n (\d+)\) \(temperature ([0-9d.]+)\) \(auto (:[a-z-]+)\)", txt)
    assert m and m.group(1) == "100000" and m.group(2) == "1d3" and m.group(3) == ":prob-settle"
    sig = inspect.signature(Walker.adaptive_steps).parameters
    assert sig["n"].default == 30000          # walker-adaptive-steps :946
