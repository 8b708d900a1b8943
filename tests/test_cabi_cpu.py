"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/synchformer_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from synchformer_b200 import _lib, build

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(REPO, 'include', 'synchformer_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(sfb_[a-z0-9_]+)\s*\(', text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert sorted(_lib.SIGNATURES) == declared, 'ctypes signatures and header disagree'


def test_host_only_entry_points():
    lib = _lib.load()
    assert lib.sfb_abi_version() == 1
    assert isinstance(lib.sfb_last_error(), bytes)


def test_argument_validation_happens_before_any_launch():
    lib = _lib.load()
    # null pointers / bad shapes are rejected on the host with SFB_E_INVALID (-1) and an error text
    rc = lib.sfb_gemm_bf16(None, 8, None, None, None, 0, None, 8, 1, 8, 8, 0, 0, None)
    assert rc == -1 and b'null' in lib.sfb_last_error()
    rc = lib.sfb_mel_frontend(None, None, 0, None)
    assert rc == -1
    d = _lib.AttnDesc()
    assert lib.sfb_attention(ctypes.byref(d), None) == -1


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG (B200_PROFILING.md 'What proves a Blackwell-native kernel')."""
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([cuobjdump, '-sass', build.LIB], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'LDTM', 'UTMALDG', 'HMMA'):
        assert mnemonic in sass, mnemonic
