"""Host output pool of the façade (libinflx_rs.host_output): CPU-side behaviour.

The reference allocates every output with np.zeros (consistency_conditions.py:290, 354, 407, 464,
515, 582) and the callee overwrites every element; the pool must hand out arrays that behave the
same way for the caller (shape, dtype, C order, writable, independent lifetimes)."""
import gc
import time

import numpy as np
import pytest

from inflatox_b200 import consistency_conditions as cc
from inflatox_b200 import libinflx_rs as rs


@pytest.fixture(autouse=True)
def _eager_pin_jobs(monkeypatch):
    # the default mode parks the pin job until an engine call has completed; there is none here
    monkeypatch.setenv("INFLATOX_PIN_MODE", "eager")


def _wait_jobs():
    for job in list(rs._pin_jobs.values()):
        job.join(timeout=30)


def test_host_output_is_a_plain_writable_c_array():
    a = rs.host_output((300, 500, 6))
    assert a.shape == (300, 500, 6) and a.dtype == np.float64
    assert a.flags.c_contiguous and a.flags.writeable and a.ctypes.data % 64 == 0
    a[:] = 3.0
    b = rs.host_output((300, 500, 6))  # a second live array must not alias the first
    b[:] = 4.0
    assert a[0, 0, 0] == 3.0 and a[-1, -1, -1] == 3.0 and b[-1, -1, -1] == 4.0
    _wait_jobs()


def test_blocks_return_to_the_pool_and_are_reused():
    a = rs.host_output((1 << 20,))
    a[:] = 7.0
    addr = a.ctypes.data
    nbytes = rs._round_block((1 << 20,), np.float64)[2]
    del a
    gc.collect()
    _wait_jobs()
    pin_key = (nbytes, rs._placement)  # pinned blocks are pooled per (size, NUMA placement)
    assert rs._pin_pool.get(pin_key) or rs._page_pool.get(nbytes), "the block was not returned"
    b = rs.host_output((1 << 20,))
    # without a GPU the pageable block comes back; with one, the freshly pinned block may be used
    assert b.ctypes.data == addr or rs._pin_pool.get(pin_key) is not None
    views = [rs.host_output((1 << 20,)) for _ in range(rs._POOL_DEPTH + 2)]
    del views, b
    gc.collect()
    assert len(rs._page_pool.get(nbytes, [])) <= rs._POOL_DEPTH
    _wait_jobs()


def test_a_view_keeps_the_block_alive():
    a = rs.host_output((1000, 6))
    a[:] = np.arange(6)
    v = a[:, 3]
    del a
    gc.collect()
    c = rs.host_output((1000, 6))  # must not be handed the block `v` still looks at
    c[:] = -1.0
    assert (v == 3.0).all()
    _wait_jobs()


def test_facade_small_outputs_stay_np_zeros(monkeypatch):
    out = cc._new_output((10, 10))
    assert isinstance(out, np.ndarray) and out.base is None and not out.any()
    monkeypatch.setenv("INFLATOX_PINNED", "0")
    big = cc._new_output((1024, 1024, 6))
    assert big.base is None and not big.any()


def test_pin_budget_env(monkeypatch):
    monkeypatch.setenv("INFLATOX_PINNED_MAX_GB", "0.5")
    assert rs._pin_budget() == 1 << 29
    monkeypatch.delenv("INFLATOX_PINNED_MAX_GB")
    assert rs._pin_budget() > 0


def test_deferred_pin_job_waits_for_an_engine_call(monkeypatch):
    monkeypatch.setenv("INFLATOX_PIN_MODE", "deferred")
    a = rs.host_output((1 << 19, 3))
    job = rs._pin_jobs[(rs._round_block((1 << 19, 3), np.float64)[2], rs._placement)]
    time.sleep(0.2)
    assert job.is_alive(), "the job must wait until the call the array is for has returned"
    with rs._gate:  # stands in for the engine call
        a[:] = 1.0
    job.join(timeout=30)
    assert not job.is_alive()


def test_pin_modes(monkeypatch):
    monkeypatch.setenv("INFLATOX_PIN_MODE", "off")
    before = dict(rs._pin_jobs)
    a = rs.host_output((1 << 18, 5))
    assert rs._pin_jobs == before and a.flags.writeable
