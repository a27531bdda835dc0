"""Host-side plans of the two PCIe schedules of cytvdn_denoise, checked WITHOUT a GPU through the C ABI
(`cytvdn_pipeline_schedule`, `cytvdn_stream_plan` -- the functions the device code path itself runs).

The checks replay the plans on a tiny model of the state (version numbers instead of arrays): every read must see
the iterate it expects, every voxel must end at the final iterate."""
import ctypes as C

import numpy as np
import pytest

from cytvdn_b200 import _lib


def schedule(nbox, n_iter):
    lib = _lib.load()
    cnt = C.c_int64(0)
    assert lib.cytvdn_pipeline_schedule(nbox, n_iter, None, None, 0, C.byref(cnt)) == 0
    box = (C.c_int32 * max(cnt.value, 1))()
    it = (C.c_int32 * max(cnt.value, 1))()
    assert lib.cytvdn_pipeline_schedule(nbox, n_iter, box, it, cnt.value, C.byref(cnt)) == 0
    return list(zip(box[:cnt.value], it[:cnt.value]))


@pytest.mark.parametrize("nbox,n_iter", [(16, 100), (16, 34), (16, 35), (16, 10), (16, 1), (2, 7), (3, 50), (5, 12),
                                         (1, 4), (16, 0)])
def test_pipeline_schedule_respects_the_dependence_cone(nbox, n_iter):
    """Replay on version numbers.  State set q of box c holds `ver[q][c]` = the iterate stored there.  Iteration m
    of box c reads set m%2 of boxes c-1, c, c+1 (all must hold iterate m) and writes iterate m+1 into set (m+1)%2."""
    order = schedule(nbox, n_iter)
    ver = [[0] * nbox, [None] * nbox]                      # set 0 holds the input (iterate 0), set 1 nothing yet
    runs = {}
    first_launch_of_box, last_launch_of_box = {}, {}
    for pos, (c, m) in enumerate(order):
        boxes = range(nbox) if c < 0 else [c]
        for b in boxes:
            for nb in (b - 1, b, b + 1):
                if 0 <= nb < nbox:
                    assert ver[m % 2][nb] == m, f"launch {pos}: box {b} iteration {m} reads box {nb} at {ver[m % 2][nb]}"
        for b in boxes:
            ver[(m + 1) % 2][b] = m + 1
            runs[(b, m)] = runs.get((b, m), 0) + 1
            first_launch_of_box.setdefault(b, pos)
            last_launch_of_box[b] = pos
    assert runs == {(b, m): 1 for b in range(nbox) for m in range(n_iter)}      # everything exactly once
    if n_iter:
        assert all(ver[n_iter % 2][b] == n_iter for b in range(nbox))
    if nbox > 1 and n_iter > 0:
        # what the pipeline is for: box c starts before box c+1 (its input arrives earlier) and finishes before it
        # (its result leaves earlier); box 0 needs only boxes 0 and 1 to be uploaded
        assert all(first_launch_of_box[b] < first_launch_of_box[b + 1] for b in range(nbox - 1))
        assert all(last_launch_of_box[b] < last_launch_of_box[b + 1] for b in range(nbox - 1))
        assert order[0] == (0, 0)
        for pos, (c, m) in enumerate(order):               # at its first iteration a box needs uploads <= c+1 only
            if m == 0 and c >= 0:
                assert max(b for b, _ in order[:pos + 1]) <= c
    if n_iter > 2 * nbox + 2:                              # long runs: whole-array sweeps between two wavefronts
        assert sum(1 for c, _ in order if c < 0) == n_iter - 2 * nbox
    else:
        assert all(c >= 0 for c, _ in order)


def test_pipeline_schedule_argument_checks():
    lib = _lib.load()
    cnt = C.c_int64(0)
    assert lib.cytvdn_pipeline_schedule(0, 5, None, None, 0, C.byref(cnt)) == 1
    assert lib.cytvdn_pipeline_schedule(4, 5, None, None, 0, None) == 1
    box, it = (C.c_int32 * 3)(), (C.c_int32 * 3)()
    assert lib.cytvdn_pipeline_schedule(4, 5, box, it, 3, C.byref(cnt)) == 1
    assert b"capacity" in lib.cytvdn_last_error()
    assert cnt.value == 20


def stream_plan(shape, dtype, n_fista, n_plain, budget):
    lib = _lib.load()
    P = _lib.DenoiseParams()
    P.ndim, P.dtype = len(shape), 0 if dtype == "float32" else 1
    for k, n in enumerate(shape):
        P.shape[k] = n
    P.iters_fista, P.iters_plain, P.bc_mode = n_fista, n_plain, 2
    out = (C.c_int64 * 8)()
    rc = lib.cytvdn_stream_plan(C.byref(P), int(budget), out)
    return rc, dict(zip(["P", "K", "core", "tiles", "passes", "arrays", "plane_bytes", "host_bytes"], out[:]))


@pytest.mark.parametrize("shape,dtype,nF,nU,planes", [
    ((41, 5, 6, 16), "float32", 23, 0, 12), ((41, 5, 6, 16), "float32", 7, 6, 16), ((30, 4, 5, 13), "float32", 0, 9, 8),
    ((26, 6, 22), "float64", 12, 0, 9), ((10, 3, 4, 8), "float64", 5, 0, 100), ((64, 2, 3, 8), "float32", 4, 0, 40),
    ((1024, 1024, 128, 128), "float32", 100, 0, 70), ((256, 256, 128, 128), "float32", 100, 0, 5),
])
def test_stream_plan_geometry_and_replay(shape, dtype, nF, nU, planes):
    """Replay the out-of-core schedule on version numbers per axis-0 plane: the host state is updated in place, the
    planes two tiles share travel through the carry buffer, a plane can be advanced only when both neighbours hold
    the previous iterate."""
    nd, n0, M = len(shape), shape[0], nF + nU
    elem = 4 if dtype == "float32" else 8
    n3p = -(-shape[-1] // (16 // elem)) * (16 // elem)
    plane_b = int(np.prod(shape[1:-1])) * n3p * elem
    arrays = 2 + nd * (2 if nF else 1)
    budget = 2 * arrays * plane_b * planes + 1000
    rc, g = stream_plan(shape, dtype, nF, nU, budget)
    assert rc == 0
    assert g["arrays"] == arrays and g["plane_bytes"] == plane_b
    P, K, core = g["P"], g["K"], g["core"]
    # the whole axis when two slots hold it, else two slots + the carry buffer (2K <= P/2 planes) within the budget
    assert P == (n0 if planes >= n0 else budget // (5 * arrays * plane_b // 2))
    if P >= n0:
        assert (K, core, g["tiles"], g["passes"], g["host_bytes"]) == (M, n0, 1, 1, 0)
    else:
        assert K == min(M, max(1, P // 4)) and core == P - 2 * K and core >= 2 * K >= K
        assert g["tiles"] == -(-n0 // core) and g["passes"] == -(-M // K)
        per_voxel_state = nd * (2 if (nF > K) else 1) if g["passes"] > 1 else 0
        assert g["host_bytes"] == per_voxel_state * n0 * plane_b
    if n0 > 200:            # the replay below is per plane; geometry checks suffice for the big shapes
        return
    host = [0] * n0                                        # iterate held by the host state, per plane
    m0 = 0
    while m0 < M:
        Kp = min(K, M - m0)
        tiles = []
        for t in range(g["tiles"]):
            c0, c1 = t * core, min(n0, (t + 1) * core)
            tiles.append((max(0, c0 - Kp), min(n0, c1 + Kp), c0, c1))
        assert all(e1 - e0 <= P for e0, e1, _, _ in tiles)             # a tile fits its slot
        loaded, carry = {}, {}

        def upload(t):
            """Planes shared with tile t-1 come from the carry buffer, the rest from the host state."""
            e0, e1, _, _ = tiles[t]
            have = min(tiles[t - 1][1], e1) - e0 if t > 0 else 0
            assert 0 <= have <= 2 * K
            st = {g_: carry[g_] for g_ in range(e0, e0 + have)}
            for g_ in range(e0 + have, e1):
                assert host[g_] == m0, f"tile {t} reads plane {g_} already advanced on the host"
                st[g_] = host[g_]
            assert all(v == m0 for v in st.values())
            loaded[t] = st
            carry.clear()
            if t + 1 < len(tiles):                         # saved before tile t iterates
                for g_ in range(tiles[t + 1][0], min(e1, tiles[t + 1][1])):
                    carry[g_] = st[g_]
                assert len(carry) <= 2 * K

        upload(0)
        for t, (e0, e1, c0, c1) in enumerate(tiles):
            if t + 1 < len(tiles):
                upload(t + 1)                              # runs ahead of the iterations of tile t
            st = loaded.pop(t)
            for k in range(Kp):
                lo = e0 + (k + 1 if e0 > 0 else 0)
                hi = e1 - (k + 1 if e1 < n0 else 0)
                new = dict(st)
                for g_ in range(lo, hi):
                    for nb in (g_ - 1, g_, g_ + 1):
                        if 0 <= nb < n0:
                            assert st[nb] == m0 + k, f"plane {g_} iteration {m0 + k} reads plane {nb} at {st.get(nb)}"
                    new[g_] = m0 + k + 1
                st = new
            for g_ in range(c0, c1):
                assert st[g_] == m0 + Kp
                host[g_] = m0 + Kp
        m0 += Kp
    assert host == [M] * n0


def test_stream_plan_errors():
    rc, _ = stream_plan((41, 5, 6, 16), "float32", 10, 0, 100)          # not even two tiles of 4 planes + carry
    assert rc == 3 and b"do not fit" in _lib.load().cytvdn_last_error()
    rc, _ = stream_plan((41, 5, 6, 16), "float32", 0, 0, 1 << 30)
    assert rc == 1


@pytest.mark.parametrize("shape,nF,nU,planes,ndev", [
    ((41, 5, 6, 16), 23, 0, 12, 2), ((41, 5, 6, 16), 7, 6, 12, 3), ((64, 2, 3, 8), 9, 0, 16, 4), ((30, 4, 5, 13), 0, 9, 8, 2),
    ((9, 4, 6, 8), 7, 0, 8, 4), ((100, 2, 3, 8), 40, 0, 24, 3),
])
@pytest.mark.parametrize("order", ["forward", "reverse", "last_first"])
def test_sharded_stream_plan_replay(shape, nF, nU, planes, ndev, order):
    """Replay `cytvdn_denoise_sharded_streamed`'s schedule for several devices on version numbers, with the devices'
    tile loops interleaved adversarially between the two barriers of a pass: whatever the order, every plane a tile
    copies in -- from the shared host state, the carry buffer or the edge snapshot -- holds the iterate of the pass's
    start, and every plane ends at the final iterate."""
    lib = _lib.load()
    nd, n0, M = len(shape), shape[0], nF + nU
    n3p = -(-shape[-1] // 4) * 4
    plane_b = int(np.prod(shape[1:-1])) * n3p * 4
    arrays = 2 + nd * (2 if nF else 1)
    budget = (11 * arrays * plane_b // 4) * planes + 1000
    P_ = _lib.DenoiseParams()
    P_.ndim, P_.dtype = nd, 0
    for k, n in enumerate(shape):
        P_.shape[k] = n
    P_.iters_fista, P_.iters_plain, P_.bc_mode = nF, nU, 2
    out = (C.c_int64 * 8)()
    assert lib.cytvdn_stream_plan_sharded(C.byref(P_), int(budget), ndev, out) == 0
    P, K, core, nt, passes = out[0], out[1], out[2], out[3], out[4]
    assert P == planes and K == min(M, max(1, P // 4)) and core <= P - 2 * K and nt == -(-n0 // core)
    per = -(-nt // ndev)
    ranges = [(min(nt, r * per), min(nt, (r + 1) * per)) for r in range(ndev)]
    assert sum(hi - lo for lo, hi in ranges) == nt
    host = [0] * n0
    m0 = 0
    while m0 < M:
        Kp = min(K, M - m0)
        first = m0 == 0
        tiles = [(max(0, t * core - Kp), min(n0, min(n0, (t + 1) * core) + Kp), t * core, min(n0, (t + 1) * core)) for t in range(nt)]
        assert all(e1 - e0 <= P for e0, e1, _, _ in tiles)

        class Dev:
            def __init__(self, lo, hi):
                self.lo, self.hi, self.loaded, self.carry, self.edge = lo, hi, {}, {}, {}

            def read_host(self, g_):
                # pass 0 builds its state from the (never written) input; later passes read the shared host state
                if not first:
                    assert host[g_] == m0, f"plane {g_} read from the host at iterate {host[g_]}, pass starts at {m0}"
                return m0

            def upload(self, t):
                e0, e1, _, _ = tiles[t]
                have = min(tiles[t - 1][1], e1) - e0 if t > self.lo else 0
                edge_g0 = min(n0, self.hi * core)
                tail = (e1 - edge_g0) if (t == self.hi - 1 and self.hi < nt and not first) else 0
                st = {g_: self.carry[g_] for g_ in range(e0, e0 + have)}
                for g_ in range(e0 + have, e1 - tail):
                    st[g_] = self.read_host(g_)
                for g_ in range(e1 - tail, e1):
                    st[g_] = self.edge[g_]
                assert all(v == m0 for v in st.values())
                self.loaded[t] = st
                self.carry = {}
                if t + 1 < self.hi:
                    for g_ in range(tiles[t + 1][0], min(e1, tiles[t + 1][1])):
                        self.carry[g_] = st[g_]

            def snapshot(self):
                if self.hi < nt and self.hi > self.lo and not first:
                    g0 = min(n0, self.hi * core)
                    self.edge = {g_: self.read_host(g_) for g_ in range(g0, min(n0, g0 + Kp))}

            def run_tile(self, t):
                e0, e1, c0, c1 = tiles[t]
                if t + 1 < self.hi:
                    self.upload(t + 1)
                st = self.loaded.pop(t)
                for k in range(Kp):
                    lo = e0 + (k + 1 if e0 > 0 else 0)
                    hi = e1 - (k + 1 if e1 < n0 else 0)
                    new = dict(st)
                    for g_ in range(lo, hi):
                        for nb in (g_ - 1, g_, g_ + 1):
                            if 0 <= nb < n0:
                                assert st[nb] == m0 + k
                        new[g_] = m0 + k + 1
                    st = new
                for g_ in range(c0, c1):
                    assert st[g_] == m0 + Kp
                    host[g_] = m0 + Kp                      # written back to the shared host state

        devs = [Dev(lo, hi) for lo, hi in ranges]
        for d in devs:                                      # phase 1 (before the barrier): first tile + edge snapshot
            if d.hi > d.lo:
                d.snapshot()                                # (same copy stream, this order: one-tile devices use it at once)
                d.upload(d.lo)
        seq = list(range(ndev))                             # phase 2: tile loops in an adversarial device order
        if order == "reverse":
            seq = seq[::-1]
        elif order == "last_first":
            seq = seq[-1:] + seq[:-1]
        for r in seq:                                       # one device runs its whole pass before the next one starts
            for t in range(devs[r].lo, devs[r].hi):
                devs[r].run_tile(t)
        m0 += Kp
    assert host == [M] * n0
