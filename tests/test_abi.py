"""The C-ABI library loads and exports every symbol include/mipb200.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    hdr = open(os.path.join(ROOT, "include", "mipb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mipb200_[a-z_0-9]+)\s*\(", hdr)))


def test_header_and_python_binding_agree(mip):
    assert _declared_functions() == sorted(mip.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(mip):
    lib = ctypes.CDLL(mip.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} is declared in include/mipb200.h but not exported by libmipb200.so"


def test_header_compiles_as_c_and_cxx(tmp_path):
    import subprocess
    for comp, ext in (("gcc", "c"), ("g++", "cpp")):
        src = tmp_path / f"t.{ext}"
        src.write_text('#include "mipb200.h"\nint main(void){ mipb200_config c; c.width = 0; return sizeof(mipb200_result) > 0 ? c.width : 1; }\n')
        subprocess.run([comp, "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / f"t_{ext}.o")], check=True)


def test_struct_layout_matches_ctypes(mip, tmp_path):
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mipb200.h"\nint main(void){ printf("%zu %zu %zu %zu %zu ", sizeof(mipb200_config), sizeof(mipb200_result), offsetof(mipb200_result, cost), offsetof(mipb200_result, gpu_ms), offsetof(mipb200_config, emit)); printf("%zu %zu %zu %zu\\n", offsetof(mipb200_config, bit_depth), offsetof(mipb200_config, top_k), offsetof(mipb200_result, top_k), offsetof(mipb200_result, topk_cost)); return 0; }\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert [int(v) for v in out] == [ctypes.sizeof(mip.Config), ctypes.sizeof(mip.Result), mip.Result.cost.offset, mip.Result.gpu_ms.offset, mip.Config.emit.offset,
                                     mip.Config.bit_depth.offset, mip.Config.top_k.offset, mip.Result.top_k.offset, mip.Result.topk_cost.offset]


def test_no_gpu_means_loud_failure_not_fallback(mip):
    """Without a CUDA device the engine must refuse to exist (there is no CPU path)."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    import pytest
    with pytest.raises(mip.MipError) as ei:
        mip.Engine(256, 128)
    assert ei.value.code in (-3, -2)
    assert mip.lib().mipb200_num_ctus(1920, 1080) == 135 and mip.lib().mipb200_num_ctus(3840, 2160) == 510


def test_product_never_touches_the_oracle():
    """Nothing shipped under vvc-mip-gpu_b200/ or include/ may import, link or execute oracle/."""
    bad = []
    for base in ("vvc-mip-gpu_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            if os.sep + "lib" in dp or os.sep + "bin" in dp or "__pycache__" in dp:
                continue
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cpp", ".h", ".c", "Makefile")):
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    if re.search(r"oracle[/.]|import oracle|mip_oracle|libmip_oracle", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_create_validates_the_configuration_before_touching_cuda(mip):
    """Bad geometry / filter / shortlist / depth are MIPB200_EINVAL (-1) with a message, also on a box without a GPU; a
    valid configuration on such a box is MIPB200_ENODEV (-3): there is no CPU fallback."""
    import pytest
    for kw in (dict(width=100, height=128), dict(width=128, height=130), dict(width=128, height=128, filter_type=9),
               dict(width=128, height=128, filter_type=5, kernel_idx=3), dict(width=128, height=128, slots=0),
               dict(width=128, height=128, emit=0), dict(width=128, height=128, top_k=3), dict(width=128, height=128, bit_depth=9),
               dict(width=128 * 300, height=128 * 200)):
        with pytest.raises(mip.MipError) as ei:
            mip.Engine(**kw)
        assert ei.value.code == -1, (kw, str(ei.value))
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(mip.MipError) as ei:
            mip.Engine(128, 128)
        assert ei.value.code == -3 and "no CPU fallback" in str(ei.value)


def _build_example(tmp_path):
    import subprocess
    exe = tmp_path / "minimal"
    libdir = os.path.join(ROOT, "vvc-mip-gpu_b200", "lib")
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "minimal.c"),
                    "-L", libdir, "-lmipb200", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    return exe


def test_plain_c_example_builds_and_refuses_to_run_without_a_gpu(mip, tmp_path):
    """examples/minimal.c is C (not C++): it compiles against the header, links the library and -- on a box without a GPU --
    reports the missing device instead of computing anything on the CPU."""
    import subprocess
    import torch
    exe = _build_example(tmp_path)
    if torch.cuda.is_available():
        return
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 2 and "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_plain_c_example_runs(mip, tmp_path):
    import subprocess
    exe = _build_example(tmp_path)
    r = subprocess.run([str(exe), "384", "264"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("frame ") == 8 and "8 kernel launches" in r.stdout
