"""The C-ABI shared library loads on a box without a GPU and exports every symbol include/ntf_b200.h declares
(no compute calls here)."""
import ctypes, os, re, subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, 'include', 'ntf_b200.h')


@pytest.fixture(scope='module')
def built():
    from opentf_b200.csrc import build
    return build.build()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ntf_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_the_documented_surface():
    names = declared_symbols()
    for must in ('ntf_csr_bag_fwd', 'ntf_csr_bag_bwd', 'ntf_neg_sample', 'ntf_out_train', 'ntf_adam_step', 'ntf_infer_scores',
                 'ntf_topk_select', 'ntf_topk_merge', 'ntf_last_error', 'ntf_version', 'ntf_out_train_workspace_bytes'):
        assert must in names


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    for name in declared_symbols():
        assert hasattr(lib, name), f'{name} declared in ntf_b200.h but not exported'
    assert lib.ntf_version() == 2


def test_python_binding_covers_the_header(built):
    from opentf_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    _lib.lib()  # types every function


def test_built_for_sm_100a_only(built):
    out = subprocess.run(['cuobjdump', '--list-elf', built], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs


def test_no_cpu_fallback_without_a_gpu(built):
    import torch
    if torch.cuda.is_available(): pytest.skip('this box has a GPU')
    from opentf_b200 import _lib
    from opentf_b200.engine import Engine
    with pytest.raises(_lib.NtfError):
        Engine(10, [8], 13, 'cuda:0')
    h = ctypes.c_void_p()
    assert _lib.lib().ntf_create(0, ctypes.byref(h)) != 0  # and the C side refuses too
    assert _lib.last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'opentf_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), f
