"""Error behaviour of the C ABI on a live device: bad arguments come back as PF_E* codes with a
message, never as a crash; a sharded handle refuses whole-tick calls."""
import ctypes as C

import numpy as np
import pytest
import torch

from swarm_robotics_pheromones_b200 import _ffi, abi
from swarm_robotics_pheromones_b200.field import PheroField, Swarm

pytestmark = pytest.mark.gpu


def test_bad_arguments_return_codes():
    L = _ffi.lib()
    cfg = abi.default_config(64, 128, 1)
    g = PheroField(cfg)
    h = g._h
    assert L.pf_deposit(h, 10, None, None, None, None, None) == abi.PF_EINVAL
    assert L.pf_step_field(h, -1, None) == abi.PF_EINVAL
    assert L.pf_field_fill(h, 5, 1.0, None) == abi.PF_EINVAL
    assert L.pf_set_stencil_variant(h, 99, 0) == abi.PF_EINVAL
    assert L.pf_set_option(h, 1234, 1) == abi.PF_EINVAL
    assert b"unknown option" in L.pf_last_error(h)
    p = C.c_void_p()
    assert L.pf_field_ptr(h, 0, 7, C.byref(p), None) == abi.PF_EINVAL
    bad = abi.PFAgents()
    bad.n = 5  # null arrays
    assert L.pf_agents_deposit(h, C.byref(bad), None) == abi.PF_EINVAL
    assert b"null arrays" in L.pf_last_error(h)
    cfg.sense_mode = abi.SENSE_HEADING
    x = torch.zeros(4, device="cuda")
    with pytest.raises(_ffi.PFError):
        g.sense(x, x, None)  # heading-relative sensing needs theta
    g.close()


def test_sharded_handle_refuses_whole_tick_and_multi_step():
    cfg = abi.default_config(64, 128, 1)
    cfg.stencil = 5
    cfg.row0, cfg.row1 = 16, 48
    g = PheroField(cfg)
    sw = Swarm(np.array([0.3], np.float32), np.array([0.5], np.float32), np.zeros(1, np.float32))
    with pytest.raises(_ffi.PFError) as e:
        g.step(sw, 1)
    assert e.value.code == abi.PF_ESTATE
    with pytest.raises(_ffi.PFError):
        g.step_field(2)
    g.step_field(1)  # one step per halo exchange is fine
    g.close()


def test_trail_and_heatmap_argument_checks():
    L = _ffi.lib()
    h = C.c_void_p()
    assert L.pf_trail_create(0, 10, 65, C.byref(h)) == abi.PF_EINVAL   # capacity > 64
    assert L.pf_trail_create(0, 0, 32, C.byref(h)) == abi.PF_EINVAL
    w = np.ones(3)
    assert L.pf_heatmap_create(0, 8, 8, 100, w.ctypes.data, C.byref(h)) == abi.PF_EINVAL
    assert L.pf_heatmap_create(0, 0, 8, 2, w.ctypes.data, C.byref(h)) == abi.PF_EINVAL
