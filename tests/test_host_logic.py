"""Host-side logic added in round 2 that needs no GPU: the pinned staging pool of the image preparation (slot ownership,
ADVICE r1) and the aborted-loop report of the persistent decode loops."""
import numpy as np
import pytest


def test_staging_pool_slot_ownership(pkg):
    P = pkg.preprocess
    pool = P._PinnedPool()
    i0, b0 = pool.get(1000)
    i1, b1 = pool.get(2000)
    i2, b2 = pool.get(500)                       # both slots held by "plans" that have not run: a private buffer
    assert {i0, i1} == {0, 1} and i2 == -1
    assert len({b0.data_ptr(), b1.data_ptr(), b2.data_ptr()}) == 3
    pool.release(i0)                             # a plan dropped without running gives its slot back
    i3, b3 = pool.get(800)
    assert i3 == i0 and b3.data_ptr() == b0.data_ptr()
    pool.mark(i1, None)                          # run(): copy queued (no event on a CPU-only box)
    i4, _ = pool.get(100)
    assert i4 == i1
    pool.mark(-1, None); pool.release(-1)        # private buffers are not tracked


def test_unrun_plans_keep_their_pixels(pkg):
    """three ResizePlans alive at once (host side only: the plan builder runs on the CPU)"""
    P = pkg.preprocess
    g = np.random.default_rng(3)
    imgs = [[g.integers(0, 256, (30 + k, 90 + 7 * k), dtype=np.uint8)] for k in range(3)]
    plans = [P.ResizePlan(b, 64, 320) for b in imgs]
    assert len({p.host.data_ptr() for p in plans}) == 3
    for p, b in zip(plans, imgs):
        assert np.array_equal(p.host[: b[0].size].numpy().reshape(b[0].shape), b[0])


def test_aborted_loop_is_an_exception(pkg):
    from hmer_img2latex_b200.model.seq2seq import _checked_steps
    assert _checked_steps(0) == 0 and _checked_steps(150) == 150
    with pytest.raises(RuntimeError, match="aborted"):
        _checked_steps(-1)
