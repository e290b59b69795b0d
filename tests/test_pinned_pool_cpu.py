"""The sub-allocator behind the page-locked result arrays (magphase_b200/_lib.py:_PinnedPool), with a fake backing
allocator: no CUDA needed.  Page-locking must not recur in a steady-state loop, whatever the order of takes and returns."""
import ctypes
import gc

import numpy as np

from magphase_b200._lib import _PinnedPool


def make_pool(arena_mb=4, max_mb=16):
    keep = []

    def alloc(size):
        b = ctypes.create_string_buffer(size)
        keep.append(b)
        return ctypes.addressof(b)
    p = _PinnedPool(alloc=alloc)
    p.ARENA, p.MAX_TOTAL = arena_mb << 20, max_mb << 20
    p._keep = keep
    return p


def test_blocks_are_reused_and_coalesced():
    p = make_pool()
    a = p.empty((1000, 60), np.float32)
    b = p.empty(300000, np.float32)
    a[:] = 1.0
    b[:] = 2.0
    assert p.stats['allocs'] == 1 and a.flags['C_CONTIGUOUS'] and float(a.sum()) == 60000.0 and float(b[-1]) == 2.0
    addr_a = a.ctypes.data
    del a, b
    gc.collect()
    assert p.arenas[0][2] == [[0, p.ARENA]]                      # everything returned and merged into one block
    c = p.empty((1000, 60), np.float32)
    assert c.ctypes.data == addr_a and p.stats['allocs'] == 1


def test_steady_state_never_page_locks_again():
    p = make_pool(arena_mb=4, max_mb=64)
    rng = np.random.Generator(np.random.PCG64(3))
    live = []
    for it in range(400):
        if live and (len(live) >= 6 or rng.random() < 0.4):
            live.pop(int(rng.integers(0, len(live))))            # out-of-order returns, like several worker threads
        n = int(rng.integers(1, 900000))
        live.append(p.empty(n, np.uint8))
        if it == 100:
            warm = p.stats['allocs']
    assert p.stats['allocs'] == warm and p.stats['pageable'] == 0
    views = [x[::2] for x in live]                               # a view keeps its block alive
    del live
    gc.collect()
    assert sum(len(a[2]) for a in p.arenas) >= 1 and any(a[2] != [[0, a[1]]] for a in p.arenas)
    del views
    gc.collect()
    assert all(a[2] == [[0, a[1]]] for a in p.arenas)


def test_budget_spent_falls_back_to_pageable_memory():
    p = make_pool(arena_mb=1, max_mb=2)
    a = p.empty(900000, np.uint8)
    b = p.empty(900000, np.uint8)
    c = p.empty(900000, np.uint8)                                 # third arena would exceed the budget
    assert p.stats['allocs'] == 2 and p.stats['pageable'] == 1 and c.size == 900000
    big = p.empty(3 << 20, np.uint8)                              # larger than the whole budget
    assert p.stats['pageable'] == 2 and big.size == 3 << 20
    assert a.size == b.size


def test_ensure_reserves_whole_arenas_up_front():
    p = make_pool(arena_mb=1, max_mb=8)
    p.ensure(int(2.5 * (1 << 20)))
    assert p.stats['allocs'] == 3 and p.total == 3 << 20
    p.ensure(1 << 20)                                             # already there
    assert p.stats['allocs'] == 3
    a = [p.empty(900000, np.uint8) for _ in range(3)]
    assert p.stats['allocs'] == 3 and p.stats['reuses'] == 3
    p.ensure(100 << 20)                                           # capped by the budget
    assert p.total == 8 << 20
    del a
