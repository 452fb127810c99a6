"""CPU: the C-ABI library loads, exports every symbol include/topo_b200.h declares, agrees with the
ctypes binding, and validates arguments before touching CUDA."""
import ctypes as C
import os
import re

import pytest
import torch

from tests.helpers import ROOT

HEADER = os.path.join(ROOT, "include", "topo_b200.h")


def declared_functions(debug=False):
    """entry points the header declares for the product library, or (debug=True) only inside #ifdef TOPO_DEBUG_KERNELS"""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    block = re.compile(r"#ifdef TOPO_DEBUG_KERNELS(.*?)#endif", re.S)
    text = "\n".join(block.findall(text)) if debug else block.sub("", text)
    return sorted(set(re.findall(r"\b(topo_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_the_binding_binds():
    from topo_audio_autoencoder_b200 import _lib
    declared = declared_functions()
    assert len(declared) >= 30
    assert sorted(_lib.SIGNATURES) == declared, "include/topo_b200.h and _lib.SIGNATURES list different entry points"


def test_library_exports_every_declared_symbol():
    from topo_audio_autoencoder_b200 import _lib
    raw = C.CDLL(_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(raw, name), f"{name} is declared in the header but not exported by libtopo_b200.so"
    assert raw.topo_version() >= 100


def test_debug_entry_points_are_not_in_the_product_library():
    from topo_audio_autoencoder_b200 import _lib
    debug = declared_functions(debug=True)
    assert debug == sorted(_lib.DEBUG_SIGNATURES) and all(n.startswith("topo_debug_") for n in debug)
    assert not any(n.startswith("topo_debug_") for n in declared_functions())
    raw = C.CDLL(_lib.LIB_PATH)
    for name in debug:
        assert not hasattr(raw, name), f"{name} leaked into the product library"
    twin = C.CDLL(_lib.DEBUG_LIB_PATH)
    for name in debug + declared_functions():
        assert hasattr(twin, name), f"{name} missing from libtopo_b200_debug.so"


def test_no_torch_types_cross_the_boundary():
    text = open(HEADER).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)          # declarations only, comments stripped
    assert "torch" not in code.lower() and "at::" not in code and "#include <cuda" not in code
    assert 'extern "C"' in code


def test_arguments_are_validated_before_cuda():
    from topo_audio_autoencoder_b200._lib import lib, ERR_INVALID, TopoError, check
    assert lib.topo_rectify_fwd(None, None, 1e-10, 1, None, None) == ERR_INVALID
    assert b"null" in lib.topo_last_error()
    assert lib.topo_spmm_csr(-1, None, None, None, None, 64, None, None) == ERR_INVALID
    with pytest.raises(TopoError):
        check(lib.topo_tables_create_ex(0, 0, C.byref(C.c_void_p())))
    seg = (C.c_int64 * 5)(129150, 128763, 128757, 129129, 130065)
    dp = lib.topo_distance_padded_size(seg, 5)
    assert dp % 64 == 0 and 645864 <= dp < 645864 + 5 * 64
    assert lib.topo_distance_image_bytes(129, seg, 5) == 2 * (dp // 64) * 3 * 128 * 128
    assert lib.topo_distance_workspace_floats(3, 7, 5) == 2 * 5 * 3 * 7


def test_cpu_tensors_are_refused():
    from topo_audio_autoencoder_b200._lib import ptr, TopoError
    with pytest.raises(TopoError, match="no CPU fallback"):
        ptr(torch.zeros(4))


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 9, 20, 33])
def test_host_tables_match_itertools_order(n):
    """The closed-form rank/unrank tables equal the reference's itertools.combinations construction
    (rectifier.py:28-55) -- host side only, no device."""
    from oracle import rectifier_oracle as ro
    from topo_audio_autoencoder_b200.rectifier import _Tables
    t, orc = _Tables(n, upload=False), ro.make_tables(n)
    assert t.counts == list(orc.sizes)
    assert torch.equal(t.simplex_vertices(1), orc.edges)
    assert torch.equal(t.simplex_vertices(2), orc.triangles)
    assert torch.equal(t.simplex_vertices(3), orc.tetra)
    assert torch.equal(t.faces(1).long(), orc.edges)
    assert torch.equal(t.faces(2).long(), orc.tri_edges)
    assert torch.equal(t.faces(3).long(), orc.tet_tris)
    for r, faces in ((0, orc.edges), (1, orc.tri_edges), (2, orc.tet_tris)):
        cof = t.cofaces(r)
        if cof.numel() == 0:
            continue
        # coface lists are the ascending inverse of the face lists
        inv = [[] for _ in range(t.counts[r])]
        for s, fs in enumerate(faces.tolist()):
            for f in fs:
                inv[f].append(s)
        assert cof.tolist() == inv
