"""Host-side checks that need no GPU: the C-ABI library loads, exports every symbol the header
declares, agrees with the Python mirror of the config defaults, and fails loudly without a GPU."""
import ctypes as C
import subprocess

import pytest
import torch

from swarm_robotics_pheromones_b200 import _ffi, abi, build


def test_library_exports_every_declared_symbol():
    L = _ffi.lib()
    declared = _ffi.header_symbols()
    assert len(declared) >= 30
    assert set(declared) == set(_ffi.SIGNATURES), "ctypes table out of sync with include/pherofield.h"
    for name in declared:
        assert hasattr(L, name), f"libpherofield.so does not export {name}"
    assert L.pf_abi_version() == abi.PF_ABI_VERSION


def test_config_default_matches_python_mirror():
    L = _ffi.lib()
    for (H, W, Cn) in [(1024, 1024, 1), (32768, 32768, 2), (497, 495, 1)]:
        c = abi.PFConfig()
        assert L.pf_config_default(C.byref(c), H, W, Cn) == 0
        py = abi.default_config(H, W, Cn)
        assert bytes(c) == bytes(py), {k: (v, py.as_dict()[k]) for k, v in c.as_dict().items() if v != py.as_dict()[k]}
    bad = abi.PFConfig()
    assert L.pf_config_default(C.byref(bad), 0, 8, 1) == abi.PF_EINVAL
    assert L.pf_config_default(C.byref(bad), 8, 8, 9) == abi.PF_EINVAL


def test_reference_constants():
    c = abi.default_config(16, 16)
    assert (c.intensity_explore, c.intensity_high, c.intensity_homing) == (1.0, 20.0, 5.0)  # sup:232-234
    assert c.high_every == 3 and abs(c.move_threshold - 0.0015) < 1e-9                      # sup:316, 260
    assert abs(c.max_speed - 3.14) < 1e-6 and c.cruise_speed == 1.0                         # sup:74, 197
    bl, br = abi.braitenberg_sums()
    assert abs(bl - 1.572) < 1e-12 and abs(br - 1.572) < 1e-12
    bl60, _ = abi.braitenberg_sums((60.0,) * 8)
    assert abs(1.0 + bl60 - 2.38778125) < 1e-12
    assert abs(c.dt - 0.064) < 1e-7 and abs(c.wheel_radius - 0.02) < 1e-9


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    L = _ffi.lib()
    cfg = abi.default_config(64, 64)
    h = C.c_void_p()
    rc = L.pf_create(C.byref(cfg), C.byref(h))
    assert rc == abi.PF_ECUDA
    assert b"no CPU fallback" in L.pf_last_error(None)
    from swarm_robotics_pheromones_b200.field import PheroField
    with pytest.raises(RuntimeError):
        PheroField(cfg)


def test_invalid_configs_are_rejected_before_touching_the_gpu():
    L = _ffi.lib()
    h = C.c_void_p()
    cfg = abi.default_config(64, 64)
    cfg.stencil = 7
    assert L.pf_create(C.byref(cfg), C.byref(h)) == abi.PF_EINVAL
    cfg = abi.default_config(64, 64)
    cfg.row0, cfg.row1 = 10, 5
    assert L.pf_create(C.byref(cfg), C.byref(h)) == abi.PF_EINVAL
    cfg = abi.default_config(64, 64)
    cfg.row1 = 32
    cfg.sense_dist = 0.05  # 5 cells: deeper than the one ghost row of a strip
    assert L.pf_create(C.byref(cfg), C.byref(h)) == abi.PF_EINVAL
    assert b"sense_dist" in L.pf_last_error(None)


def test_product_path_does_not_import_the_oracle():
    """the product package must never import, include, link or load anything under oracle/
    (test infrastructure); comments may cite it"""
    import pathlib
    import re
    pkg = pathlib.Path(build.PKG)
    py_import = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)", re.M)
    loads = re.compile(r"(libpf_oracle|oracle[/\\][\w.]+\.(so|py)|#\s*include\s*[\"<][^\">]*oracle)")
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        text = p.read_text()
        assert not py_import.search(text), f"{p} imports the oracle"
        code = "\n".join(ln for ln in text.splitlines() if not ln.strip().startswith(("//", "#", "*", "/*")))
        assert not loads.search(code), f"{p} loads or includes the oracle"
    assert "oracle" not in " ".join(build.NVCC_FLAGS) and all("oracle" not in str(x) for x in build.sources())


def test_sass_has_tma_bulk_copy_and_128bit_accesses():
    out = subprocess.run(["cuobjdump", "-sass", str(build.LIB)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "UBLKCP" in out.stdout          # cp.async.bulk -> TMA engine
    assert "SYNCS" in out.stdout           # mbarrier
    assert "LDG.E.128" in out.stdout and "STG.E.128" in out.stdout
