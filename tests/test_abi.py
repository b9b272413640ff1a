"""CPU: the C-ABI library loads and exports every symbol include/lidal_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest
import torch

from lidal_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lidal_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from lidal_b200.build import build_library
    build_library()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/lidal_b200.h but not exported"


def test_python_binding_covers_header():
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert _lib.lib().lb_abi_version() == 3


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of every field of the two ABI structs, taken from include/lidal_b200.h by the C compiler, equal the
    ctypes mirrors in lidal_b200/_lib.py (field names differ only for `in`, a Python keyword)."""
    import subprocess
    structs = {"lb_conv_args": (_lib.ConvArgs, {"inp": "in"}), "lb_frame_ref": (_lib.FrameRef, {})}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "lidal_b200.h"', 'int main(void) {']
    for cname, (ct, rename) in structs.items():
        lines.append(f'  printf("{cname} sizeof %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {rename.get(fname, fname)}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = {}
    for line in subprocess.check_output([str(exe)], text=True).splitlines():
        cname, fname, val = line.split()
        got[(cname, fname)] = int(val)
    for cname, (ct, _) in structs.items():
        assert got[(cname, "sizeof")] == ctypes.sizeof(ct), cname
        for fname, _t in ct._fields_:
            assert got[(cname, fname)] == getattr(ct, fname).offset, (cname, fname)
    assert got[("lb_conv_args", "sizeof")] == 176 and got[("lb_frame_ref", "sizeof")] == 32


def test_no_cpu_fallback():
    import lidal_b200.compat as ts
    with pytest.raises(_lib.LidalError):
        ts.nn.functional.sphash(torch.zeros((4, 4), dtype=torch.int))
    with pytest.raises(_lib.LidalError):
        ts.nn.functional.spcount(torch.zeros(4, dtype=torch.int), 4)


def test_argument_validation_without_gpu():
    L = _lib.lib()
    assert L.lb_hash(None, -1, None, None) == -1
    assert b"n < 0" in L.lb_last_error()
    assert L.lb_hashtable_bytes(1000) == 2048 * 12
    a = _lib.ConvArgs()
    a.n_out = 5
    assert L.lb_conv_fwd(ctypes.byref(a), None) == -1


@pytest.mark.skipif(not os.path.isdir("/root/reference/network"), reason="reference checkout only exists in the build container")
def test_install_lets_the_unmodified_reference_networks_import():
    """Drop-in check for boundary B1: after compat.install() the reference's own network/minkunet.py and network/spvcnn.py
    import `torchsparse` from this package, construct, and expose exactly the golden state_dict keys / shapes."""
    import importlib
    import subprocess
    import sys
    code = (
        "import sys, numpy as np; sys.path[:0] = [%r, '/root/reference'];"
        "import lidal_b200.compat as c; c.install();"
        "import torchsparse, torchsparse.nn as spnn; assert torchsparse is c and spnn.Conv3d is c.nn.Conv3d;"
        "from network.minkunet import MinkUNet; from network.spvcnn import SPVCNN;"
        "g = np.load(%r, allow_pickle=True);"
        "m = MinkUNet(19); s = SPVCNN(16);"
        "assert [f'{k}:{tuple(v.shape)}' for k, v in m.state_dict().items()] == list(g['minkunet_keys']);"
        "assert [f'{k}:{tuple(v.shape)}' for k, v in s.state_dict().items()] == list(g['spvcnn_keys']);"
        "import torch.nn as nn; assert isinstance(m.stem[1], nn.BatchNorm1d); print('ok')"
    ) % (ROOT, os.path.join(ROOT, "tests", "golden", "nets.npz"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
