"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol the
header declares (no compute calls here)."""
import os
import re

from multistgraph_b200 import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "matgcn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(matgcn_\w+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    build.build()
    assert os.path.exists(_cabi.LIB_PATH)
    lib = _cabi.lib()
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(_cabi.EXPORTED_SYMBOLS) == names
    assert lib.matgcn_abi_version() == _cabi.ABI_VERSION


def test_workspace_queries_are_pure_host_functions():
    lib = _cabi.lib()
    dims = (24, 403, 64, 64, 64, 5)
    n = lib.matgcn_encoder_layer_fwd_ws_bytes(*dims)
    assert n > 0 and n % 4 == 0
    assert lib.matgcn_encoder_layer_bwd_ws_bytes(*dims, 1) > lib.matgcn_encoder_layer_bwd_ws_bytes(*dims, 0)
    y_off = lib.matgcn_encoder_layer_y_offset(*dims)
    ph = lib.matgcn_encoder_layer_slot_offset(b"PH", *dims)
    assert y_off == ph + 5 * 403 * 64 * 64
    assert lib.matgcn_encoder_layer_y_tstride(*dims) == 5 * 403 * 64 * 64
    assert lib.matgcn_encoder_layer_slot_offset(b"nope", *dims) == 2 ** 64 - 1


def test_bad_arguments_return_error_codes():
    lib = _cabi.lib()
    rc = lib.matgcn_adaptive_adj_fwd(None, None, 4, 2, None, 8, None)
    assert rc != 0
    assert b"null" in lib.matgcn_last_error()
