"""CPU: the C-ABI library loads and exports every symbol include/stk.h declares; argument
validation returns error codes (no compute is launched without a GPU)."""
import ctypes
import os
import re

from stonkgs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_in_header():
    src = open(os.path.join(ROOT, "include", "stk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|long long|int64_t)\s+(stk_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_in_header() == _lib.declared_symbols()


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_in_header():
        assert hasattr(lib, name), name
    assert _lib.load().stk_version() == _lib.STK_VERSION


def test_error_codes_and_text():
    lib = _lib.load()
    epi = _lib.GemmEpilogue()
    rc = lib.stk_gemm(0, None, 0, 0, None, 8, None, 8, 128, 256, 64, 0, None, 8, ctypes.byref(epi), 1)
    assert rc == -1 and "null operand" in _lib.last_error()
    rc = lib.stk_gemm(0, None, 0, 0, None, 8, None, 8, 0, 256, 64, 0, None, 8, ctypes.byref(epi), 1)
    assert rc == -1 and "empty problem" in _lib.last_error()
    rc = lib.stk_attn_fwd(0, None, ctypes.c_void_p(16), None, 1, 100, ctypes.c_void_p(16), None)
    assert rc == -1 and "S must be" in _lib.last_error()
    assert lib.stk_launch_count() == 0


def test_epilogue_enum_matches_binding():
    """The epilogue codes of include/stk.h and of the ctypes binding are the same numbers."""
    src = open(os.path.join(ROOT, "include", "stk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    header = {k: int(v) for k, v in re.findall(r"\bSTK_(EPI_\w+)\s*=\s*(\d+)", src)}
    assert len(header) >= 13 and sorted(header.values()) == list(range(len(header)))
    for name, code in header.items():
        assert getattr(_lib, name) == code, name


def test_new_entry_points_validate_arguments():
    lib = _lib.load()
    p16 = ctypes.c_void_p(16)
    # joint embedding stage: text part must be shorter than the sequence, padding not below the sequence
    rc = lib.stk_embed_joint_ln_fwd_shape(0, None, p16, None, 1, 256, 200, 384, p16, p16, 10, p16, p16, p16, p16, p16, None, None,
                                          None, None)
    assert rc == -1 and "bad joint shape" in _lib.last_error()
    rc = lib.stk_dropout_fwd(0, None, p16, 8, 1, 2, 200, p16)          # threshold is round(128 p) < 128
    assert rc == -1 and "stk_dropout_fwd" in _lib.last_error()
    rc = lib.stk_layernorm_bwd_fused(0, None, p16, p16, 8, p16, p16, p16, p16, p16, p16, None, None, 1, 2, 13)
    assert rc == -1 and "dxm" in _lib.last_error()                     # dropout needs the masked output
    assert lib.stk_launch_count() == 0


def test_workspace_query_and_head_entry_points():
    """stk_query_workspace is host arithmetic (no GPU): sizes of the calls that take a workspace; the single-call heads
    validate their arguments before anything is launched."""
    lib = _lib.load()
    B, S, R, V = 64, 512, 2432, 175003
    assert lib.stk_query_workspace(_lib.WS_ATTN_BWD, B, S) == 4 * (B * S * 768 + B * 12 * S)
    pitch = 2 * ((V + 255) // 256)
    fwd = lib.stk_query_workspace(_lib.WS_LINEAR_CE_FWD, R, V)
    assert fwd >= 4 * (R * pitch * 2 + R) and fwd % 256 == 0
    bwd = lib.stk_query_workspace(_lib.WS_LINEAR_CE_BWD, R, V)
    assert 0 < bwd <= (48 << 20) + 256 and bwd % (2 * R) % 1 == 0
    assert lib.stk_query_workspace(_lib.WS_LINEAR_CE_BWD, 8, 1000) == 2 * 8 * 1024      # one chunk: vocabulary rounded to 256
    assert lib.stk_query_workspace(99, 1, 1) == -1 and "unknown op" in _lib.last_error()
    assert lib.stk_query_workspace(_lib.WS_ATTN_BWD, 0, 512) == -1
    p16 = ctypes.c_void_p(16)
    rc = lib.stk_linear_ce_fwd(0, None, p16, p16, R, V, p16, p16, 1024, p16, p16, None)
    assert rc == -1 and "workspace too small" in _lib.last_error()
    rc = lib.stk_linear_ce_bwd(0, None, p16, p16, R, V, p16, p16, p16, p16, 1024, p16, p16)
    assert rc == -1 and "workspace too small" in _lib.last_error()
    rc = lib.stk_compact_labels(0, None, p16, 4, 256, 512, 0, 28996, 0, p16, p16, p16, None)
    assert rc == -1 and "stk_compact_labels" in _lib.last_error()
    epi = _lib.GemmEpilogue()
    epi.resid, epi.ldr, epi.ln_gamma, epi.ln_beta = 16, 768, 16, 16
    rc = lib.stk_gemm(0, None, 0, 0, p16, 768, p16, 768, 256, 768, 768, _lib.EPI_BIAS_DROP_RESID_LN, p16, 768, ctypes.byref(epi), 1)
    assert rc == -1 and "drop_thr" in _lib.last_error()
    for q_rows in (100, 640, -128):                                    # multiples of 128 in (0, S] only
        rc = lib.stk_attn_fwd_qrows(0, None, p16, None, 2, 512, q_rows, p16, None)
        assert rc == -1 and "q_rows" in _lib.last_error()
    assert lib.stk_launch_count() == 0
