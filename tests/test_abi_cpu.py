"""CPU-side checks of the C-ABI library: it builds, loads, exports every symbol include/fjgpu.h declares,
its ctypes struct mirror has the header's layout, and — with no GPU — it refuses to create a context
instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import pytest

import scenekit as sk

REPO = sk.REPO


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build_fjgpu()
    sk.pkg()
    from fujiyama_renderer_b200 import abi
    return abi.load_fjgpu()


def header_functions():
    with open(os.path.join(REPO, "include", "fjgpu.h")) as f:
        txt = f.read()
    return sorted(set(re.findall(r"\b(fjgpu_[a-z_]+)\s*\(", txt)))


def test_exports_every_declared_symbol(lib):
    from fujiyama_renderer_b200 import abi
    names = header_functions()
    assert sorted(abi.FJGPU_SYMBOLS) == names
    for n in names:
        assert getattr(lib, n) is not None


def test_struct_layout_matches_header(tmp_path):
    """sizeof/offsetof of every struct, printed by a C program compiled against include/fjgpu.h."""
    from fujiyama_renderer_b200 import abi
    structs = {"fjgpu_instance": abi.Instance, "fjgpu_shader": abi.Shader, "fjgpu_light": abi.Light,
               "fjgpu_camera": abi.Camera, "fjgpu_render_params": abi.RenderParams, "fjgpu_tile": abi.Tile,
               "fjgpu_stats": abi.Stats, "fjgpu_scene_info": abi.SceneInfo}
    lines = ['#include "fjgpu.h"', "#include <stdio.h>", "#include <stddef.h>", "int main(void){"]
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).strip().split("\n"))
    for cname, cls in structs.items():
        assert int(out[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = C.c_void_p()
    rc = lib.fjgpu_create(0, C.byref(ctx))
    assert rc == -3 and not ctx.value                  # FJGPU_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.fjgpu_last_error(None)
    assert lib.fjgpu_api_version() == 2
    from fujiyama_renderer_b200 import device
    with pytest.raises(device.FjGpuError):
        device.Device(0)


def test_product_does_not_link_the_oracle():
    """The oracle is test infrastructure: no product source mentions it and libfjgpu has no dependency on it."""
    pkg = os.path.join(REPO, "fujiyama-renderer_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cc", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "libfjoracle" not in txt and "fjo_" not in txt and "import scenekit" not in txt, f
    out = subprocess.check_output(["ldd", os.path.join(pkg, "csrc", "libfjgpu.so")], text=True)
    assert "oracle" not in out
