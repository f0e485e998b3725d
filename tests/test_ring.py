"""The page-locked buffer ring against a model of the reference's AtomicCircularBuffer
(src/AtomicAbstractSDRs.jl:67-190).  CPU only: pinned=False, no device call."""
import threading

import numpy as np
import pytest

import tempestsdr_b200 as tsdr


class RefRing:
    """circ_put! / circ_take! as the reference writes them (single-threaded restatement)"""

    def __init__(self, nEch, depth):
        self.buf = np.zeros((depth, nEch), np.complex64)   # zeros(T, nEch*depth), :51
        self.depth, self.w, self.r, self.t_new = depth, 0, 0, 0

    def put(self, data):
        self.buf[self.w] = data                              # atomic_write, :112-117
        self.w = (self.w + 1) % self.depth                   # atomic_update, :103-107
        self.t_new = min(self.t_new + 1, self.depth)         # atomic_prodData, :122-126

    def take(self):
        assert self.t_new > 0                                # wait_consData, :142-150
        out = self.buf[self.r].copy()
        self.r = (self.r + 1) % self.depth
        self.t_new = max(self.t_new - 1, 0)                  # atomic_consData, :130-134
        return out


@pytest.mark.parametrize("depth", [1, 2, 5])
def test_ring_follows_reference_model(depth):
    nEch = 37
    rng = np.random.default_rng(depth)
    ring = tsdr.AtomicCircularBuffer(nEch, depth, pinned=False)
    ref = RefRing(nEch, depth)
    lost = 0
    for step in range(400):
        if ref.t_new == 0 or rng.random() < 0.6:
            d = (rng.standard_normal(nEch) + 1j * rng.standard_normal(nEch)).astype(np.complex64)
            lost += ref.t_new == depth
            ref.put(d)
            tsdr.circ_put(ring, d)
        else:
            got = tsdr.circ_take(np.empty(nEch, np.complex64), ring, timeout_ms=1000)
            assert np.array_equal(got, ref.take())
        st = ring.stats()
        assert st["available"] == ref.t_new and st["overwritten"] == lost
    ring.close()


def test_ring_errors_and_timeout():
    ring = tsdr.AtomicCircularBuffer(16, 3, pinned=False)
    with pytest.raises(ValueError):
        ring.put(np.zeros(15, np.complex64))                 # the reference asserts equal lengths (:113)
    with pytest.raises(tsdr.TempestError) as e:
        ring.take(timeout_ms=20)                             # nothing produced yet
    assert e.value.status == -5
    ring.put(np.arange(16, dtype=np.complex64))
    assert np.array_equal(ring.take(timeout_ms=20), np.arange(16, dtype=np.complex64))
    ring.close()
    with pytest.raises(tsdr.TempestError):
        tsdr.AtomicCircularBuffer(16, 0, pinned=False)


def test_ring_int16_slots():
    ring = tsdr.AtomicCircularBuffer(8, 2, dtype=np.int16, pinned=False)
    d = np.arange(16, dtype=np.int16).reshape(8, 2)
    ring.put(d)
    assert np.array_equal(ring.take(timeout_ms=100), d.reshape(-1))
    ring.close()


def test_ring_two_threads_no_torn_buffers():
    # a fast producer laps a slow consumer: buffers may be lost or arrive out of order (as in the reference),
    # but every buffer taken is one whole put, and the counters add up
    nEch, depth, n_put = 4096, 4, 600
    ring = tsdr.AtomicCircularBuffer(nEch, depth, pinned=False)
    taken = []

    def producer():
        for k in range(1, n_put + 1):
            ring.put(np.full(nEch, k, np.complex64))

    def consumer():
        out = np.empty(nEch, np.complex64)
        while True:
            try:
                ring.take(out, timeout_ms=300)
            except tsdr.TempestError:
                return                                        # producer finished and the ring is drained
            assert (out == out[0]).all()
            taken.append(int(out[0].real))

    tc, tp = threading.Thread(target=consumer), threading.Thread(target=producer)
    tc.start(); tp.start(); tp.join(); tc.join()
    st = ring.stats()
    assert st["produced"] == n_put and st["consumed"] == len(taken) and st["available"] == 0
    assert len(taken) + st["overwritten"] == n_put
    assert all(1 <= k <= n_put for k in taken) and n_put in taken[-depth:]
    ring.close()


def test_ring_zero_copy_slots():
    import ctypes as C
    from tempestsdr_b200 import _lib
    L = _lib.load()
    ring = tsdr.AtomicCircularBuffer(4, 2, pinned=False)
    slot = C.c_void_p()
    _lib.check(L.tsdr_ring_acquire_write(ring._h, C.byref(slot)))
    assert L.tsdr_ring_acquire_write(ring._h, C.byref(slot)) == -1          # one write slot at a time
    np.ctypeslib.as_array(C.cast(slot, C.POINTER(C.c_float)), (8,))[:] = np.arange(8)
    _lib.check(L.tsdr_ring_commit(ring._h))
    assert L.tsdr_ring_commit(ring._h) == -1
    rd = C.c_void_p()
    _lib.check(L.tsdr_ring_acquire_read(ring._h, C.byref(rd), 100))
    assert np.array_equal(np.ctypeslib.as_array(C.cast(rd, C.POINTER(C.c_float)), (8,)), np.arange(8))
    assert ring.stats()["available"] == 1                                    # still counted until released
    _lib.check(L.tsdr_ring_release_read(ring._h))
    assert ring.stats()["available"] == 0 and L.tsdr_ring_slot_bytes(ring._h) == 32
    ring.close()
