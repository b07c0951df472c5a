"""CPU tests of the host-side schedule logic: the left-looking schedule's bulk-lane plan (csrc/lu.cu: plan_left) and
the ctypes mirrors of the interface structs."""
import ctypes

import pytest


def plan(mplu, n, nb, eager):
    lib = mplu.load_library()
    lib.mplu_debug_plan_left.argtypes = [ctypes.c_int] * 3 + [ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    cnt = lib.mplu_debug_plan_left(n, nb, eager, None, 0)
    assert cnt >= 0
    buf = (ctypes.c_int * (5 * max(cnt, 1)))()
    assert lib.mplu_debug_plan_left(n, nb, eager, buf, cnt) == cnt
    return [tuple(buf[5 * i:5 * i + 5]) for i in range(cnt)]


@pytest.mark.parametrize("n,nb", [(32768, 2048), (131072, 2048), (2304, 512), (1000, 256), (4096, 1024), (640, 512),
                                  (512, 512), (100, 128), (9000, 1152)])
@pytest.mark.parametrize("eager", [0, 1])
def test_left_plan_applies_every_update_once_in_order_and_in_time(mplu, n, nb, eager):
    npad = -(-n // 128) * 128
    nbe = min(nb, npad)
    nt = -(-npad // nbe)
    ops = plan(mplu, n, nb, eager)
    seen = {m: [] for m in range(nt)}
    last_step = 0
    for step, k, m0, m1, mandatory in ops:
        assert 1 <= step <= nt - 2 and step >= last_step  # steps in order; the last step has no bulk-lane updates
        last_step = step
        assert 0 <= k <= step - 1          # L panel k is complete from step k+1 on
        assert step + 1 <= m0 < m1 <= nt   # only columns that have not had their turn
        if mandatory:
            assert (m0, m1) == (step + 1, step + 2)
        for m in range(m0, m1):
            assert k <= m - 2              # update m-1 belongs to the chain lane / the final-update launch
            seen[m].append((step, k))
    for m in range(nt):
        ks = [k for _, k in seen[m]]
        assert ks == list(range(max(m - 1, 0))), (m, ks)     # every update k = 0..m-2 exactly once, increasing
        assert all(step <= m - 1 for step, _ in seen[m])      # ... before the column's own step m
    if not eager:
        assert all(mand for *_, mand in ops)                  # strictly left-looking: nothing ahead of need


def bounds(mplu, n, nb, edge):
    lib = mplu.load_library()
    lib.mplu_debug_tile_bounds.argtypes = [ctypes.c_int] * 3 + [ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    cnt = lib.mplu_debug_tile_bounds(n, nb, edge, None, 0)
    buf = (ctypes.c_int * cnt)()
    assert lib.mplu_debug_tile_bounds(n, nb, edge, buf, cnt) == cnt
    return list(buf)


@pytest.mark.parametrize("n,nb,edge", [(32768, 2048, 1024), (32768, 2048, 512), (9000, 1152, 384), (4096, 512, 128), (8192, 1024, 512),
                                       (2304, 512, 256), (32768, 2048, 0), (1024, 512, 256), (33000, 2048, 1024)])
@pytest.mark.parametrize("eager", [0, 1, 2, 3])  # bit 1: paired updates (opts.update_pair), reported as their two updates
def test_left_plan_with_narrow_edge_tiles(mplu, n, nb, edge, eager):
    """opts.edge_nb: boundaries stay multiples of 128, no block column is wider than nb, the first and last are `edge` wide
    whenever the matrix has at least four nb-wide block columns, and the plan on those boundaries keeps its invariants."""
    npad = -(-n // 128) * 128
    tb = bounds(mplu, n, nb, edge)
    assert tb[0] == 0 and tb[-1] == npad and all(b % 128 == 0 for b in tb)
    w = [b - a for a, b in zip(tb, tb[1:])]
    assert all(0 < x <= nb for x in w)
    if edge and npad >= 4 * nb:
        assert w[0] == edge and w[-1] == edge
    else:
        assert w[:-1] == [min(nb, npad)] * (len(w) - 1)
    nt = len(w)
    lib = mplu.load_library()
    lib.mplu_debug_plan_left_edge.argtypes = [ctypes.c_int] * 4 + [ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    cnt = lib.mplu_debug_plan_left_edge(n, nb, edge, eager, None, 0)
    buf = (ctypes.c_int * (5 * max(cnt, 1)))()
    assert lib.mplu_debug_plan_left_edge(n, nb, edge, eager, buf, cnt) == cnt
    seen = {m: [] for m in range(nt)}
    for i in range(cnt):
        step, k, m0, m1, mandatory = buf[5 * i:5 * i + 5]
        assert 1 <= step <= nt - 2 and 0 <= k <= step - 1 and step + 1 <= m0 < m1 <= nt
        for m in range(m0, m1):
            seen[m].append((step, k))
    for m in range(nt):
        assert [k for _, k in seen[m]] == list(range(max(m - 1, 0))), m
        assert all(step <= m - 1 for step, _ in seen[m])


def test_eager_plan_balances_the_steps(mplu):
    """n=32768, nb=2048: no step carries more than ~1.5x the average update work (plain left-looking: the last step
    carries 14 updates, the first one)."""
    npad, nb = 32768, 2048

    def work(k, m0, m1):
        return (2.0 * (npad - (k + 1) * nb) + nb) * (m1 - m0)

    for eager, bound in ((1, 1.5), (0, 3.0)):
        per = {}
        for step, k, m0, m1, _ in plan(mplu, npad, nb, eager):
            per[step] = per.get(step, 0.0) + work(k, m0, m1)
        avg = sum(per.values()) / len(per)
        ratio = max(per.values()) / avg
        assert (ratio <= bound) if eager else (ratio > 1.5), (eager, ratio)


def test_ctypes_mirrors_match_the_library(mplu):
    lib = mplu.load_library()
    assert lib.mplu_sizeof_options() == ctypes.sizeof(mplu.Options)
    assert lib.mplu_sizeof_stats() == ctypes.sizeof(mplu.Stats)
    o = mplu.default_options()
    # the last fields of the struct read back the library's defaults: the mirror's field order is right to the end
    assert (o.stream_host, o.schedule, o.side_sms_left, o.eager, o.stream_c, o.early_scale) == (1, 1, 16, 1, 1, 0)
    assert (o.fuse_w, o.fuse_ctas, o.lazy_touch) == (512, 8, 1)
    assert (o.flow_w, o.flow_ctas, o.flow_merge_ctas) == (2048, 16, -1)
    assert (o.edge_nb, o.pair_ts, o.update_pair) == (0, 0, 0)
