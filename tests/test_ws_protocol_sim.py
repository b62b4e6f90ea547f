"""Barrier protocol of the experimental warp-specialised pass A (variant 3), checked on CPU with a discrete-event model
(tools/sim_pass_a_ws.py): random interleavings of feeder, sampler warps and stencil warps never deadlock, never read a stale plane
and never overwrite a buffer a reader still needs; breaking the ring-depth rules the kernel static_asserts is detected."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def model():
    spec = importlib.util.spec_from_file_location("sim_pass_a_ws", os.path.join(ROOT, "tools", "sim_pass_a_ws.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_protocol_is_deadlock_free_and_ordered():
    m = model()
    src = open(os.path.join(ROOT, "sobfu_b200", "csrc", "solver_tiled.cu")).read()
    assert "constexpr int NSTAGE = %d, AHEAD = %d, NWB = %d" % (m.NSTAGE, m.AHEAD, m.NWB) in src      # the model has the kernel's ring depths
    assert "#define PAW_NSAMP %d" % m.NSAMP in src
    for seed in range(40):
        m.run(seed, 40)


def test_model_detects_broken_ring_depths():
    m = model()
    m.AHEAD = m.NSTAGE - 2          # the feeder would wait for planes the stencil can only release after the feeder's own warp moves on
    with pytest.raises(AssertionError):
        for seed in range(20):
            m.run(seed, 30)
