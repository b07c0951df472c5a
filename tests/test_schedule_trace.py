"""CPU test of the factorization schedules' cross-lane synchronisation.  mplu_debug_trace() dry-runs the schedule code of
csrc/lu.cu (the very functions that enqueue the real launches) and returns every launch with the lane it runs on and the
array regions it reads / writes, plus every event record / wait.  Replaying that with vector clocks proves that each
access happens-after the accesses it conflicts with -- a missing, mis-ordered or mis-issued event (a wait issued before
its record is not captured into the CUDA graph) fails here instead of as a timing-dependent race on a GPU."""
import ctypes

import pytest

KINDS = {0: "gemm", 1: "leaf", 2: "record", 3: "wait", 4: "cast", 5: "memset"}
ARRAYS = ["W", "Wh", "Fh", "Linv16", "Uinv16", "Tb1", "Tb2"]


def trace(mplu, n, nb, **kw):
    lib = mplu.load_library()
    lib.mplu_debug_trace.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    opts = mplu.default_options(**kw)
    need = lib.mplu_debug_trace(n, nb, ctypes.byref(opts), None, 0)
    assert need > 0, need
    buf = (ctypes.c_int * need)()
    assert lib.mplu_debug_trace(n, nb, ctypes.byref(opts), buf, need) == need
    ops, i = [], 0
    while i < need:
        kind, stream, ev, group, nreg = buf[i:i + 5]
        i += 5
        regs = [tuple(buf[i + 6 * k:i + 6 * k + 6]) for k in range(nreg)]
        i += 6 * nreg
        ops.append((kind, stream, ev, group, regs))
    return ops


def overlap(a, b):
    return a[0] == b[0] and a[1] < b[2] and b[1] < a[2] and a[3] < b[4] and b[3] < a[4]


def check(ops):
    """Vector clocks over the two lanes.  Returns (number of launches, number of cross-lane conflicts that were ordered)."""
    clock = [[0, 0], [0, 0]]          # clock[s] = what stream s has seen of (stream 0, stream 1)
    events = {}
    history = []                      # (stream, time, group, region)
    launches = ordered = 0
    for idx, (kind, s, ev, group, regs) in enumerate(ops):
        if kind == 2:
            events[ev] = list(clock[s])
            continue
        if kind == 3:
            assert ev in events, f"op {idx}: lane {s} waits for event {ev} before it was recorded (issue order)"
            clock[s] = [max(a, b) for a, b in zip(clock[s], events[ev])]
            continue
        launches += 1
        clock[s][s] += 1
        now = clock[s][s]
        for reg in regs:
            for (ps, pt, pg, preg) in history:
                if not (reg[5] or preg[5]) or not overlap(reg, preg):
                    continue
                if ps == s:
                    continue              # stream order
                assert clock[s][ps] >= pt, (
                    f"op {idx} ({KINDS[kind]}, lane {s}) {'writes' if reg[5] else 'reads'} {ARRAYS[reg[0]]}{reg[1:5]} "
                    f"without waiting for lane {ps}'s {'write' if preg[5] else 'read'} of {preg[1:5]} (its launch {pt}, seen {clock[s][ps]})")
                ordered += 1
        for reg in regs:
            history.append((s, now, group, reg))
    return launches, ordered


CASES = [(32768, 2048), (2304, 512), (1000, 256), (1536, 512), (640, 512), (9000, 1152)]


@pytest.mark.parametrize("n,nb", CASES)
@pytest.mark.parametrize("kw", [dict(schedule=1), dict(schedule=1, eager=0), dict(schedule=1, lookahead=0), dict(schedule=0),
                                dict(schedule=0, lookahead=0), dict(schedule=0, group=0), dict(schedule=1, group=0),
                                dict(schedule=1, pair_ts=1), dict(schedule=1, pair_ts=1, eager=0),
                                dict(schedule=1, update_pair=1), dict(schedule=1, update_pair=1, eager=0, pair_ts=1)])
def test_schedule_is_race_free(mplu, n, nb, kw):
    ops = trace(mplu, n, nb, **kw)
    launches, ordered = check(merge_groups(ops))
    assert launches > 0
    npad = -(-n // 128) * 128
    two_lanes = kw.get("lookahead", 1) and -(-npad // min(nb, npad)) > 2
    if two_lanes:
        assert ordered > 0      # the lanes really do hand data to each other, and every hand-off is covered by an event
    # every leaf of the matrix is factored exactly once
    leaves = sorted(r[0][1] for k, _, _, _, r in ops if k == 1)
    assert leaves == list(range(0, npad, 128))


@pytest.mark.parametrize("n,nb,edge", [(32768, 2048, 1024), (9000, 1152, 384), (4096, 512, 128), (8192, 1024, 512), (2304, 512, 256)])
@pytest.mark.parametrize("kw", [dict(), dict(eager=0), dict(group=0), dict(pair_ts=1), dict(update_pair=1)])
def test_schedule_with_narrow_edge_tiles_is_race_free(mplu, n, nb, edge, kw):
    """opts.edge_nb: the first and last block column narrower than nb (non-uniform boundaries through the whole left-looking
    schedule): same vector-clock check, and every 128-leaf still factored exactly once."""
    ops = trace(mplu, n, nb, schedule=1, edge_nb=edge, **kw)
    launches, ordered = check(merge_groups(ops))
    npad = -(-n // 128) * 128
    assert launches > 0 and ordered > 0
    leaves = sorted(r[0][1] for k, _, _, _, r in ops if k == 1)
    assert leaves == list(range(0, npad, 128))


def merge_groups(ops):
    """Launches that carry several problems (same group id) are ONE launch: merge their regions into one op."""
    out = []
    for op in ops:
        kind, s, ev, group, regs = op
        if kind == 0 and out and out[-1][0] == 0 and out[-1][3] == group and out[-1][1] == s:
            # conflicts inside one launch: a problem may not write what a sibling reads or writes
            for reg in regs:
                for preg in out[-1][4]:
                    assert not ((reg[5] or preg[5]) and overlap(reg, preg)), \
                        f"problems of one launch conflict on {ARRAYS[reg[0]]}{reg[1:5]} / {preg[1:5]}"
            out[-1] = (kind, s, ev, group, out[-1][4] + regs)
        else:
            out.append(op)
    return out


def test_checker_catches_a_missing_wait(mplu):
    """Sanity of the checker itself: drop one cross-lane wait from a correct trace and it must complain."""
    ops = merge_groups(trace(mplu, 2304, 512, schedule=1))
    check(ops)
    waits = [i for i, op in enumerate(ops) if op[0] == 3 and op[2] < 900000]
    caught = 0
    for i in waits:
        try:
            check(ops[:i] + ops[i + 1:])
        except AssertionError:
            caught += 1
    assert caught >= len(waits) // 2, (caught, len(waits))   # redundant waits exist; most are load-bearing
