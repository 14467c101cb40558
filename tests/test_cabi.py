"""CPU tier: the C-ABI library builds for sm_100a, loads, and exports every symbol include/mpdo_b200.h declares
(no compute calls - there is no GPU here). Also: the product refuses to run without a CUDA device."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'mpdo_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(mpdo_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as entry
    path = entry.build()
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), f'{name} declared in include/mpdo_b200.h but not exported'
    from MPDOSimulator._engine import lib as binding
    assert set(binding.SYMBOLS) == set(names)
    lib.mpdo_version.restype = ctypes.c_int
    assert lib.mpdo_version() >= 100


def test_library_targets_sm_100a():
    import shutil
    import subprocess
    import __graft_entry__ as entry
    if shutil.which('cuobjdump') is None:
        pytest.skip('cuobjdump not available')
    out = subprocess.run(['cuobjdump', '--list-elf', entry.build()], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


@pytest.mark.skipif(torch.cuda.is_available(), reason='only meaningful without a GPU')
def test_no_cpu_fallback():
    from MPDOSimulator._engine.prims import CudaPrims
    with pytest.raises(RuntimeError):
        CudaPrims()
    import MPDOSimulator as Simulator
    c = Simulator.TensorCircuit(qn=2, ideal=True, dtype=torch.complex64, device='cpu')
    c.h(0)
    with pytest.raises(RuntimeError):
        c.evolve(Simulator.Tools.create_ket0Series(2))
