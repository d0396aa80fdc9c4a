"""The C-ABI library loads and exports every symbol include/posecodec.h declares;
the ctypes mirrors have the header's struct sizes.  No compute calls (no GPU)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "posecodec.h")


@pytest.fixture(scope="module")
def lib():
    from mindpose_b200.csrc import build

    build.build()
    from mindpose_b200 import _lib

    return _lib.load()


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_something():
    names = _declared_functions()
    assert "pc_topdown_decode" in names and "pc_warp_affine_u8" in names and len(names) >= 14


def test_every_declared_symbol_is_exported(lib):
    from mindpose_b200 import _lib

    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in posecodec.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"


def test_version_and_error_string(lib):
    import re

    with open(HEADER) as f:
        declared = int(re.search(r"#define\s+PC_VERSION\s+(\d+)", f.read()).group(1))
    assert lib.pc_version() == declared
    assert isinstance(lib.pc_last_error(), bytes)


def test_struct_sizes_match_header(tmp_path):
    from mindpose_b200 import _lib

    structs = {
        "pc_box_params": _lib.BoxParams,
        "pc_affine_params": _lib.AffineParams,
        "pc_warp_params": _lib.WarpParams,
        "pc_encode_params": _lib.EncodeParams,
        "pc_warp_norm_params": _lib.WarpNormParams,
        "pc_topdown_decode_params": _lib.TopDownDecodeParams,
        "pc_bottomup_decode_params": _lib.BottomUpDecodeParams,
        "pc_bottomup_encode_params": _lib.BottomUpEncodeParams,
        "pc_group_params": _lib.GroupParams,
        "pc_affine_host_params": _lib.AffineHostParams,
    }
    src = tmp_path / "sizes.c"
    body = "\n".join(f'  printf("{n} %zu\\n", sizeof({n}));' for n in structs)
    src.write_text(f'#include <stdio.h>\n#include "posecodec.h"\nint main(void) {{\n{body}\n  return 0;\n}}\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True)
    for line in out.strip().splitlines():
        name, size = line.split()
        assert ctypes.sizeof(structs[name]) == int(size), name


def test_struct_field_offsets_match_header(tmp_path):
    """Every field of every ctypes mirror sits at the offset the header gives the field of the
    same name (two swapped fields of one type would keep the size)."""
    from mindpose_b200 import _lib

    structs = {
        "pc_box_params": _lib.BoxParams,
        "pc_affine_params": _lib.AffineParams,
        "pc_warp_params": _lib.WarpParams,
        "pc_encode_params": _lib.EncodeParams,
        "pc_warp_norm_params": _lib.WarpNormParams,
        "pc_topdown_decode_params": _lib.TopDownDecodeParams,
        "pc_bottomup_decode_params": _lib.BottomUpDecodeParams,
        "pc_bottomup_encode_params": _lib.BottomUpEncodeParams,
        "pc_group_params": _lib.GroupParams,
        "pc_refine_params": _lib.RefineParams,
        "pc_oks_nms_params": _lib.OksNmsParams,
        "pc_gather_target": _lib.GatherTarget,
        "pc_affine_host_params": _lib.AffineHostParams,
    }
    lines = []
    for cname, cls in structs.items():
        for field in cls._fields_:
            lines.append(f'  printf("{cname} {field[0]} %zu\\n", offsetof({cname}, {field[0]}));')
    src = tmp_path / "offsets.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "posecodec.h"\n'
                   'int main(void) {\n' + "\n".join(lines) + "\n  return 0;\n}\n")
    exe = tmp_path / "offsets"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True)
    for line in out.strip().splitlines():
        cname, field, off = line.split()
        assert getattr(structs[cname], field).offset == int(off), (cname, field)


def test_argument_errors_raise_value_error_without_gpu(lib):
    """Validation happens before any CUDA call, so it is checkable on CPU."""
    from mindpose_b200 import _lib

    p = _lib.TopDownDecodeParams()
    p.num_joints, p.height, p.width = 17, 64, 48
    p.shift_coordinate = p.dark_udp_refine = 1
    p.kernel_size = 11
    with pytest.raises(ValueError, match="cannot be `true` in the same time"):
        _lib.call("pc_topdown_decode", 0, 0, 0, 0, 0, 0, 0, ctypes.byref(p), 4, 0)
    p.shift_coordinate = 0
    p.num_joints = 1000
    with pytest.raises(ValueError):
        _lib.call("pc_topdown_decode", 0, 0, 0, 0, 0, 0, 0, ctypes.byref(p), 4, 0)
    e = _lib.EncodeParams()
    e.num_joints, e.image_w, e.image_h, e.heatmap_w, e.heatmap_h = 17, 192, 256, 48, 64
    e.sigma = 1.5  # 3 * sigma not an integer
    with pytest.raises(ValueError):
        _lib.call("pc_topdown_encode", 0, 0, 0, ctypes.byref(e), 1, 0)
    b = _lib.BottomUpEncodeParams()
    b.num_joints, b.num_scales, b.num_people, b.max_num = 17, 2, 31, 30
    b.heatmap_w[0] = b.heatmap_h[0] = 128
    b.heatmap_w[1] = b.heatmap_h[1] = 256
    b.sigma, b.tag_per_joint = 2.0, 1
    with pytest.raises(ValueError, match="exeeds the maximum num"):  # the reference's message
        _lib.call("pc_bottomup_encode", 0, 0, 0, ctypes.byref(b), 1, 0)
    # host-buffer front end: argument checks come before any CUDA call
    with pytest.raises(ValueError, match="ctx is NULL"):
        _lib.call("pc_ctx_last_transfer_bytes", None, None, None)
    a = _lib.AffineHostParams(480, 640, 3, 192, 256, 200.0, 1.25, 0, _lib.UPLOAD_ROI)
    with pytest.raises(ValueError, match="NULL pointer"):
        _lib.call("pc_crop_source_rect", None, ctypes.c_float(0.0), ctypes.byref(a), None)
    with pytest.raises(ValueError, match="ctx / params is NULL"):
        _lib.call("pc_topdown_affine_host", None, 0, 0, 0, 0, 0, 0, ctypes.byref(a), 1)


def test_missing_library_fails_loudly(monkeypatch):
    from mindpose_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libposecodec.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_host_tensor_is_rejected():
    import torch

    from mindpose_b200 import codec

    with pytest.raises(ValueError, match="CUDA"):
        codec.topdown_decode(torch.zeros(1, 17, 64, 48), torch.zeros(1, 2), torch.ones(1, 2),
                             torch.zeros(1))


def test_graft_entry_build_runs():
    """The driver's "does it build" entry point (the library is already built by the fixture,
    so this only re-checks the stamp, the import and the version the header declares)."""
    import __graft_entry__ as entry

    entry.build()


def _prototypes():
    """name -> list of parameter declarations, parsed from the header's prototypes."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|const char\s*\*)\s+(pc_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        params = [p.strip() for p in m.group(2).replace("\n", " ").split(",")]
        out[m.group(1)] = [] if params == ["void"] else params
    return out


def test_ctypes_signatures_match_the_prototypes():
    """Same number of arguments, and the same kind (pointer / 32-bit / 64-bit / float) in every
    position, as the header declares -- an argument added on one side only would otherwise
    shift everything after it silently."""
    from mindpose_b200 import _lib

    protos = _prototypes()
    assert len(protos) >= 20
    kinds = {ctypes.c_int32: "i32", ctypes.c_int: "i32", ctypes.c_uint32: "i32",
             ctypes.c_int64: "i64", ctypes.c_float: "f32", ctypes.c_double: "f64"}
    for name, params in protos.items():
        assert name in _lib.SIGNATURES, name
        _, argtypes = _lib.SIGNATURES[name]
        assert len(argtypes) == len(params), (name, len(argtypes), params)
        for decl, at in zip(params, argtypes):
            if "*" in decl:
                want = "ptr"
            elif re.search(r"\b(int64_t|long long)\b", decl):
                want = "i64"
            elif re.search(r"\bfloat\b", decl):
                want = "f32"
            elif re.search(r"\bdouble\b", decl):
                want = "f64"
            elif re.search(r"\b(int32_t|uint32_t|int)\b", decl):
                want = "i32"
            else:
                raise AssertionError(f"{name}: cannot classify `{decl}`")
            got = kinds.get(at, "ptr")
            assert got == want, (name, decl, at)
