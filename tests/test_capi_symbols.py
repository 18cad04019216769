"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol include/covidabs.h declares.
No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "covidabs.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cabs_[a-z_0-9]+)\s*\(", text)))


def test_header_functions_all_exported(native_lib):
    names = declared_functions()
    assert len(names) >= 13
    for n in names:
        assert hasattr(native_lib, n), n


def test_binding_list_matches_header():
    from covid19_agentbasedsimulation_b200 import _native
    assert sorted(_native.EXPORTED) == declared_functions()


def test_struct_sizes_match_header(native_lib, tmp_path):
    """ctypes mirrors of the ABI structs have the sizes gcc gives the header's structs."""
    import subprocess
    from covid19_agentbasedsimulation_b200 import _native
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "covidabs.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(cabs_config),sizeof(cabs_params),sizeof(cabs_trigger),sizeof(cabs_soa_host),'
                   'sizeof(cabs_info),sizeof(cabs_kernel_time));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    mirrors = [_native.Config, _native.Params, _native.Trigger, _native.SoaHost, _native.Info, _native.KernelTime]
    assert sizes == [ctypes.sizeof(m) for m in mirrors]


def test_library_is_sm100a_with_lineinfo():
    import subprocess
    from covid19_agentbasedsimulation_b200 import _native
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], stdout=subprocess.PIPE, universal_newlines=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_no_gpu_fails_loudly():
    """Without a CUDA device the product path raises; it never falls back to a CPU implementation."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    from covid19_agentbasedsimulation_b200 import _native
    from covid19_agentbasedsimulation_b200.abs import Simulation
    sim = Simulation(population_size=10)
    with pytest.raises(_native.NativeError) as e:
        sim.initialize()
    assert e.value.code == _native.ERR_NOGPU


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "covid19_agentbasedsimulation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "oracle/" not in text or f.endswith(".md"), f
