"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/physx_b200.h declares,
and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from physx_b200 import engine, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "physx_b200.h")).read()
    return sorted(set(re.findall(r"PXB_API\s+[\w\s\*]+?\b(pxb_\w+)\s*\(", src)))


def test_library_is_built_in_tree():
    assert os.path.exists(engine.lib_path()), "run __graft_entry__.build()"
    assert os.path.dirname(engine.lib_path()) == os.path.join(ROOT, "physx_b200")


def test_exports_every_declared_symbol():
    lib = engine.load_library()
    decl = declared_symbols()
    assert len(decl) >= 30
    for s in decl:
        assert hasattr(lib, s), f"{s} declared in physx_b200.h but not exported"
    assert sorted(engine.EXPORTS) == decl


def test_scene_desc_layout_matches_header():
    # 3 floats + u32 + 4 floats + 3 floats + 2 floats + 2 u32 + 2 u32 + i32 + 8 u32 = 26 words
    assert ctypes.sizeof(engine.SceneDesc) == 26 * 4
    assert scenes.ACTOR_DTYPE.itemsize == 128 and scenes.HEADER_DTYPE.itemsize == 96


def test_no_cpu_fallback():
    lib = engine.load_library()
    if lib.pxb_device_count() > 0:
        pytest.skip("CUDA device present")
    sc = scenes.box_stacks(n_stacks=1, height=2)
    with pytest.raises(engine.PhysxB200Error) as e:
        engine.Scene(sc)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_reference_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "physx_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for line in txt.splitlines():
                    if re.search(r"#include.*oracle|import oracle|oracle_lib|libpxb_oracle|oracle/_ref", line):
                        raise AssertionError(f"{f}: {line}")
