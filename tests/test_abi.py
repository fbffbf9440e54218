"""CPU-side checks of the drop-in boundary: the library builds/loads without a GPU and exports
exactly the symbols include/b2d_glue.h declares, with the arity the Python binding assumes."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b2d_glue.h")


def _header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|size_t|uint64_t|const char\*)\s+(b2d_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


@pytest.fixture(scope="module")
def built_lib():
    from faster_rcnn_pytorch_multimodal_b200 import build
    return build.build()


def test_header_declares_the_bound_functions():
    from faster_rcnn_pytorch_multimodal_b200 import _lib
    decl = _header_functions()
    assert set(decl) == set(_lib.PROTOTYPES), set(decl) ^ set(_lib.PROTOTYPES)
    for name, n_args in decl.items():
        assert len(_lib.PROTOTYPES[name][1]) == n_args, name


def test_library_exports_every_symbol(built_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = set(_header_functions()) - exported
    assert not missing, missing


def test_library_loads_and_answers_without_a_gpu(built_lib):
    from faster_rcnn_pytorch_multimodal_b200 import _lib
    L = _lib.lib()
    assert L.b2d_abi_version() == 1
    assert L.b2d_status_string(-2) == b"workspace missing or too small"
    assert L.b2d_max_pre_nms() == 16384
    # workspace queries are pure host arithmetic
    small = L.b2d_proposal_workspace_bytes(1, 24 * 78, 25, 6000, 300)
    big = L.b2d_proposal_workspace_bytes(4, 80 * 120, 25, 12000, 2000)
    assert 0 < small < big
    assert L.b2d_proposal_workspace_bytes(0, 10, 1, 1, 1) == 0
    assert L.b2d_anchor_target_workspace_bytes(1, 240000, 32) > 240000 * 17
    assert L.b2d_pipeline_device_bytes(2, 9600, 25, 1024, 80, 120, 6000, 300, 7) > 2 * 300 * 1024 * 49 * 4


def test_sass_is_sm100a_with_bulk_tma(built_lib):
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass, "plane-resident RoIAlign must load with 1-D bulk TMA"
    assert "UTMALDG" in sass, "the rows RoIAlign kernel must fetch feature rows with tiled TMA"
    assert "SYNCS" in sass and "CREDUX" in sass


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "faster_rcnn_pytorch_multimodal_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in text and "from oracle" not in text, fn
                assert "torchvision" not in text or fn.endswith((".py", ".cu", ".cuh")) and "import torchvision" not in text, fn


def test_cpu_tensor_is_rejected():
    import torch
    from faster_rcnn_pytorch_multimodal_b200 import ops, _lib
    with pytest.raises(_lib.B2DError):
        ops.nms(torch.zeros(4, 4), torch.zeros(4), 0.5)
