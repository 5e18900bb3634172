"""CPU checks of the tile scheduler of the block-task kernel (feddlib_b200/csrc/tasks.cuh): the same function the
device kernel runs with one thread per tile.  Invariants: every (element, column) pair of a tile appears in exactly one
task, a task's pairs all belong to its position, the tasks of a position sit on adjacent lanes of one pass with
rem = tasks that follow, exactly one head per position, positions without pairs get an empty head task."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "cpp", "_build", "libtask_sched.so")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "task_sched_c.cpp")
    hdr = os.path.join(ROOT, "feddlib_b200", "csrc", "tasks.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-x", "c++", src, "-o", SO])
    L = C.CDLL(SO)
    L.fb_schedule_tile.restype = C.c_int
    L.fb_schedule_tile.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    return L


def make_tile(rng, n_nodes, max_tets=32, ncol=10, holes=False):
    """Random star tile: node i has ninc[i] elements; the ncol canonical columns of an element hit distinct positions."""
    left = max_tets
    lens, nincs, recs = [], [], []
    for i in range(n_nodes):
        ninc = int(rng.integers(1, max(1, min(left - (n_nodes - 1 - i), 32)) + 1))
        left -= ninc
        ln = int(rng.integers(ncol, min(256, ncol + 3 * ninc + 5) + 1))
        r = np.zeros((ninc, 8), dtype=np.uint32)
        for t in range(ninc):
            hi = ln - 1 if holes and ln > ncol else ln      # with holes: the last position never receives a pair
            pos = rng.choice(hi, size=ncol, replace=False)
            pos[0] = 0                                       # the row node itself: every element hits position 0
            if 0 in pos[1:]:
                pos[1:][pos[1:] == 0] = [x for x in range(1, hi) if x not in pos][0]
            for jc in range(ncol):
                r[t, jc >> 1] |= np.uint32(int(pos[jc]) << (16 * (jc & 1)))
            r[t, 5] = rng.integers(0, 1 << 30)
        lens.append(ln); nincs.append(ninc); recs.append(r)
    return lens, nincs, recs


def run(lib, lens, nincs, recs, ncol=10):
    n = len(lens)
    # scatter the nodes' records over a larger array (records of consecutive bucket rows are not adjacent)
    order = np.random.default_rng(1).permutation(n)
    k0 = np.zeros(n, dtype=np.int64)
    blocks, at = [], 0
    for i in order:
        k0[i] = at + 1
        blocks += [np.full((1, 8), 0xdeadbeef, dtype=np.uint32), recs[i]]
        at += 1 + recs[i].shape[0]
    rec = np.ascontiguousarray(np.concatenate(blocks, axis=0))
    ln, ni = np.asarray(lens, dtype=np.int32), np.asarray(nincs, dtype=np.int32)
    npass = lib.fb_schedule_tile(n, ln.ctypes.data, ni.ctypes.data, k0.ctypes.data, rec.ctypes.data, 8, ncol, None)
    assert npass >= 1
    out = np.full(32 * npass + 4, 0xffffffffffffffff, dtype=np.uint64)
    assert lib.fb_schedule_tile(n, ln.ctypes.data, ni.ctypes.data, k0.ctypes.data, rec.ctypes.data, 8, ncol, out.ctypes.data) == npass
    assert np.all(out[32 * npass:] == 0xffffffffffffffff)            # nothing written past the end
    return npass, out[: 32 * npass]


def check(lens, nincs, recs, npass, words, ncol=10):
    m0 = np.concatenate([[0], np.cumsum(nincs)])
    seen = set()
    heads = {}
    w = [int(x) for x in words]
    for lane_abs, t in enumerate(w):
        npairs, pos, slot, rem, head = (t >> 36) & 7, (t >> 39) & 255, (t >> 47) & 7, (t >> 50) & 7, (t >> 53) & 1
        if t == 0:
            continue
        assert npairs <= 4 and slot < len(lens) and pos < lens[slot]
        for s in range(npairs):
            pr = (t >> (9 * s)) & 0x1ff
            m, jc = pr & 31, pr >> 5
            assert m0[slot] <= m < m0[slot + 1] and jc < ncol
            tt = m - m0[slot]
            p = (int(recs[slot][tt, jc >> 1]) >> (16 * (jc & 1))) & 0xffff
            assert p == pos
            assert (m, jc) not in seen
            seen.add((m, jc))
        if head:
            assert (slot, pos) not in heads
            heads[(slot, pos)] = lane_abs
            # the group: rem following lanes in the same pass, same position, rem counting down, not heads
            assert (lane_abs & 31) + rem < 32
            for d in range(1, rem + 1):
                u = w[lane_abs + d]
                assert (u >> 39) & 255 == pos and (u >> 47) & 7 == slot and (u >> 50) & 7 == rem - d and (u >> 53) & 1 == 0
        else:
            assert lane_abs > 0
            u = w[lane_abs - 1]                               # a non-head task follows a task of the same group
            assert (u >> 39) & 255 == pos and (u >> 47) & 7 == slot and (u >> 50) & 7 == rem + 1
    assert len(seen) == ncol * int(np.sum(nincs))
    for i, ln in enumerate(lens):
        for p in range(ln):
            assert (i, p) in heads                            # every position is written exactly once (holes: zeros)
    # balance: the tasks of a pass differ by at most one pair from the pass maximum, except where sizes run out
    sizes = np.array([(t >> 36) & 7 for t in w]).reshape(npass, 32)
    assert np.all(np.diff(sizes.max(axis=1)) <= 0)            # decreasing task size from pass to pass


@pytest.mark.parametrize("seed", range(12))
def test_random_tiles(lib, seed):
    rng = np.random.default_rng(seed)
    n_nodes = int(rng.integers(1, 9))
    lens, nincs, recs = make_tile(rng, n_nodes, holes=bool(seed & 1))
    npass, words = run(lib, lens, nincs, recs)
    check(lens, nincs, recs, npass, words)


def test_interior_vertex_star_shape(lib):
    """Multiplicities of an interior vertex row of the Kuhn cube: 24 elements, 65 positions with 24 / 6 / 4 / 2 pairs."""
    rng = np.random.default_rng(7)
    counts = [24] + [6] * 16 + [4] * 12 + [2] * 36
    # deal the 240 pair slots to 24 elements x 10 columns so that an element never hits a position twice: every
    # position goes to the elements with the fewest columns filled so far
    cols = [[] for _ in range(24)]
    for p, c in enumerate(counts):
        order = sorted(range(24), key=lambda t: (len(cols[t]), rng.random()))
        for t in order[:c]:
            cols[t].append(p)
    assert all(len(c) == 10 for c in cols)
    arr = np.array(cols)
    rec = np.zeros((24, 8), dtype=np.uint32)
    for t in range(24):
        for jc in range(10):
            rec[t, jc >> 1] |= np.uint32(int(arr[t, jc]) << (16 * (jc & 1)))
    npass, words = run(lib, [65], [24], [rec])
    check([65], [24], [rec], npass, words)
    sizes = np.array([(int(t) >> 36) & 7 for t in words]).reshape(npass, 32)
    assert npass == 3 and int(sizes.max(axis=1).sum()) <= 9   # 9 pair steps for 240 pairs (ideal 7.5)


def test_limits(lib):
    rng = np.random.default_rng(3)
    lens, nincs, recs = make_tile(rng, 2)
    ln, ni = np.asarray([300, lens[1]], dtype=np.int32), np.asarray(nincs, dtype=np.int32)
    k0 = np.zeros(2, dtype=np.int64)
    rec = np.concatenate(recs)
    assert lib.fb_schedule_tile(2, ln.ctypes.data, ni.ctypes.data, k0.ctypes.data, rec.ctypes.data, 8, 10, None) == -1
    assert lib.fb_schedule_tile(9, ln.ctypes.data, ni.ctypes.data, k0.ctypes.data, rec.ctypes.data, 8, 10, None) == -1
