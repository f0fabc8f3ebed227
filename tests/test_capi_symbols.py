"""The C-ABI library loads and exports every symbol include/vrt.h declares (no compute calls)."""
import os
import re

from conftest import ROOT


def declared_in_header():
    text = open(os.path.join(ROOT, "include", "vrt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vrt_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(vrt):
    assert declared_in_header() == vrt.capi.declared_symbols()


def test_library_exports_every_symbol(vrt):
    lib = vrt.capi.lib()
    for name in declared_in_header():
        assert getattr(lib, name) is not None
    assert lib.vrt_abi_version() == 3
    assert b"sm_100a" in lib.vrt_build_info()


def test_struct_sizes(vrt, tmp_path):
    """The ctypes mirrors have the layout of the C structs in include/vrt.h (sizes and the offsets of the last fields, taken
    from a C program compiled against the header)."""
    import ctypes as C
    import subprocess
    src = tmp_path / "sizes.c"
    src.write_text('''#include <stdio.h>
#include <stddef.h>
#include <vrt.h>
int main(void) {
    printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(vrt_hit), sizeof(vrt_lnode), sizeof(vrt_camera), sizeof(vrt_render_params),
           offsetof(vrt_render_params, mirror_y1), offsetof(vrt_render_params, autofocus), sizeof(vrt_present_params), sizeof(vrt_render_stats),
           sizeof(vrt_shade_job), sizeof(vrt_shade_result), offsetof(vrt_shade_job, direction));
    return 0;
}
''')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    RP = vrt.capi.RenderParams
    want = [vrt.HIT.itemsize, vrt.LNODE.itemsize, C.sizeof(vrt.capi.Camera), C.sizeof(RP), RP.mirror_y1.offset, RP.autofocus.offset,
            C.sizeof(vrt.capi.PresentParams), C.sizeof(vrt.capi.RenderStats), vrt.capi.SHADE_JOB.itemsize, vrt.capi.SHADE_RESULT.itemsize,
            vrt.capi.SHADE_JOB.fields["direction"][1]]
    assert got == want
    assert vrt.HIT.itemsize == 64 and vrt.LNODE.itemsize == 8 and C.sizeof(RP) == 23 * 4


def test_no_cpu_fallback_without_device(vrt):
    """On a box without a GPU the compute entry points must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    with pytest.raises(vrt.VrtError) as e:
        vrt.Context(0)
    assert e.value.code == -2   # VRT_ERR_CUDA


def test_bad_arguments_are_errors(vrt):
    import ctypes as C
    import pytest
    lib = vrt.capi.lib()
    n = C.c_uint64(0)
    assert lib.vrt_host_build_terrain_lsvo(5, None, None, 0, C.byref(n)) == -1
    assert b"depth" in lib.vrt_last_error()
    with pytest.raises(vrt.VrtError):
        vrt.host_build_lsvo_from_voxels(4, [[16, 0, 0]])       # out of range: UB in the reference, an error here


def test_new_entry_points_validate_before_touching_the_device(vrt):
    """Presentation and dynamic-scene calls: NULL / malformed arguments are VRT_ERR_INVALID on any box."""
    import ctypes as C
    lib = vrt.capi.lib()
    p = vrt.capi.PresentParams(16, 16, 3, 0.1)
    buf = (C.c_uint8 * (16 * 16 * 4))()
    assert lib.vrt_present(None, buf, buf, C.byref(p)) == -1
    assert lib.vrt_present_device(None, buf, buf, C.byref(p)) == -1
    h = C.c_void_p()
    assert lib.vrt_lsvo_create_heightfield(None, 9, None, 0, C.byref(h)) == -1
    assert lib.vrt_scene_edit_heights(None, 0, 0, 1, 1, buf) == -1
    assert lib.vrt_scene_download_heights(None, buf) == -1
    assert lib.vrt_lsvo_create_from_voxels(None, 5, None, 0, 0, C.byref(h)) == -1
    assert lib.vrt_scene_set_cells(None, None, 0, 1) == -1
    n = C.c_uint64(0)
    assert lib.vrt_scene_voxel_count(None, C.byref(n)) == -1
    assert b"NULL" in lib.vrt_last_error()


def test_node_array_validation_is_host_side(vrt):
    """vrt_lsvo_create rejects malformed arrays before touching the device (the traversal trusts child_offset)."""
    import ctypes as C
    import numpy as np
    lib = vrt.capi.lib()
    bad = np.zeros(9, vrt.LNODE)
    bad["child_mask"][0] = 1
    bad["child_offset"][0] = 5            # child block would end at slot 13 > 9
    h = C.c_void_p()
    fake_ctx = C.c_void_p(1)              # never dereferenced: validation comes first
    assert lib.vrt_lsvo_create(fake_ctx, bad.ctypes.data_as(C.c_void_p), 9, 3, 0, C.byref(h)) == -1
    assert b"malformed" in lib.vrt_last_error()
    bad["child_offset"][0] = 1
    bad["leaf_mask"][0] = 2               # leaf bit without child bit
    assert lib.vrt_lsvo_create(fake_ctx, bad.ctypes.data_as(C.c_void_p), 9, 3, 0, C.byref(h)) == -1
