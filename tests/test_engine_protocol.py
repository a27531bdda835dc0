"""Model check of the shard engine's exchange protocol (csrc/cytvdn_shard.cu) on CPU.

The GPU tests show bit-identical results on the machines they ran on; they cannot show that the counters and events
are SUFFICIENT under every timing.  This test replays the protocol -- per rank an in-order compute queue, an in-order
copy queue and an in-order upload queue holding the operations `enqueue_step` / `cytvdn_shard_run_host` issue, counters
raised by the neighbours' copy queues -- with version numbers in place of arrays and a randomised scheduler that runs
any runnable queue head, and asserts
  * at every read, that the buffer holds exactly the iterate the reader expects,
  * at every write, that no operation that is enqueued but has not run yet still expects the value being overwritten,
  * that the run never dead-locks and ends with every box at the final iterate.
Both modes: whole-array iterations (`cytvdn_shard_iterate`) and the host-pipelined wavefront over boxes of scan axis 1
(`cytvdn_shard_run_host`; the order comes from the library's own `cytvdn_pipeline_schedule`).

Modelled (names as in cytvdn_shard.cu): state sets in = it & 1 / out; the owned planes' rows of box c per set; the lower
/ upper overlap plane's rows of box c per set, written by pushes only (flags bit 1: owned-only stores); the counters
whole / box[c] per side; ev_halo; ev_pushed[box][parity] + pushed_it; ev_up per box; the wait rules "boxes c-1 .. c+1"
for a box step (its halo planes read one row of each neighbouring box) and "all boxes" for a whole sweep.
"""
import ctypes as C
import random

import pytest

from cytvdn_b200 import _lib


def pipeline_order(nbox, n_it):
    lib = _lib.load()
    cnt = C.c_int64(0)
    assert lib.cytvdn_pipeline_schedule(nbox, n_it, None, None, 0, C.byref(cnt)) == 0
    box, it = (C.c_int32 * cnt.value)(), (C.c_int32 * cnt.value)()
    assert lib.cytvdn_pipeline_schedule(nbox, n_it, box, it, cnt.value, C.byref(cnt)) == 0
    return list(zip(box[:], it[:]))


class Op:
    def __init__(self, ready, reads=(), action=None, label=""):
        self.ready, self.reads, self.action, self.label = ready, list(reads), action, label


class Rank:
    def __init__(self, r, world, nbox):
        self.r, self.nbox = r, nbox
        self.has_lo, self.has_hi = r > 0, r < world - 1
        self.buf = {}                                   # buffer key -> version held (-1: garbage)
        self.whole = {"lo": 0, "hi": 0}                 # counters in this rank's header, raised by the neighbours
        self.box = {"lo": [0] * nbox, "hi": [0] * nbox}
        self.q = {"comp": [], "copy": [], "up": []}
        self.ev = set()
        self.pushed_it = {}

    def get(self, key):
        return self.buf.get(key, -1)

    def write(self, key, new, who):
        old = self.get(key)
        if old != new:
            for queue in self.q.values():               # everything still enqueued on the rank that owns the buffer
                for op in queue:
                    assert (key, old) not in op.reads or old == -1, \
                        f"{who} overwrites {key} (iterate {old}) on rank {self.r} while '{op.label}' has not read it yet"
        self.buf[key] = new


def build(world, nbox, n_it, pipelined, base):
    ranks = [Rank(r, world, nbox) for r in range(world)]
    order = pipeline_order(nbox, n_it) if pipelined else [(-1, m) for m in range(n_it)]
    for R in ranks:
        R.whole = {"lo": base, "hi": base}              # a re-load: the previous run's last pushes left the counters at `base`
        for c in range(nbox):
            if pipelined:                               # uploads box by box on their own stream
                def up(R=R, c=c):
                    R.write(("orig", c), 0, f"upload of box {c}")
                    R.ev.add(("up", c))
                R.q["up"].append(Op(lambda: True, (), up, f"upload {c}"))
            else:
                R.buf[("orig", c)] = 0
        for (c, m) in order:
            it = base + m
            inn, out = it & 1, (it & 1) ^ 1
            boxes = list(range(nbox)) if c < 0 else [c]
            nb = list(range(nbox)) if c < 0 else [b for b in (c - 1, c, c + 1) if 0 <= b < nbox]
            key = nbox if c < 0 else c                  # index into ev_pushed
            # ---------- compute stream ----------
            conds = []
            if pipelined and m == 0:
                need = nbox - 1 if c < 0 else min(c + 1, nbox - 1)
                conds.append(lambda R=R, need=need: ("up", need) in R.ev)
            if it > 0:                                  # counters: the neighbours' planes of iteration it-1 have landed
                for side, has in (("lo", R.has_lo), ("hi", R.has_hi)):
                    if has:
                        conds.append(lambda R=R, side=side, nb=nb, it=it: all(R.whole[side] >= it or R.box[side][b] >= it for b in nb))
            if R.has_lo or R.has_hi:                    # ev_pushed: the copy engine's reads of iteration it-2 (same set)
                for q in (range(nbox + 1) if c < 0 else (c, nbox)):
                    if R.pushed_it.get((q, it & 1)) == it - 2:
                        conds.append(lambda R=R, q=q, it=it: ("pushed", q, it - 2) in R.ev)
            reads = []
            for b in nb:
                reads.append((("orig", b), 0) if m == 0 else (("own", inn, b), m))
                if m > 0 and R.has_lo:
                    reads.append((("ovlo", inn, b), m))
                if m > 0 and R.has_hi:
                    reads.append((("ovhi", inn, b), m))

            def sweep(R=R, c=c, m=m, out=out, boxes=boxes, reads=tuple(reads)):
                for k, v in reads:
                    assert R.get(k) == v, f"rank {R.r} step ({c},{m}) reads {k} at iterate {R.get(k)}, expects {v}"
                for b in boxes:
                    R.write(("own", out, b), m + 1, f"rank {R.r} step ({c},{m})")
                R.ev.add(("halo", c, m))
            R.q["comp"].append(Op(lambda conds=tuple(conds): all(f() for f in conds), reads, sweep, f"step ({c},{m})"))
            # ---------- copy stream: push the boxes' rows of the new planes, raise the neighbours' counters ----------
            if R.has_lo or R.has_hi:
                preads = [(("own", out, b), m + 1) for b in boxes]

                def push(R=R, c=c, m=m, it=it, out=out, boxes=boxes, key=key, preads=tuple(preads)):
                    for k, v in preads:
                        assert R.get(k) == v, f"rank {R.r} push ({c},{m}) reads {k} at {R.get(k)}"
                    for side, has, nbr, tgt, cside in (("lo", R.has_lo, R.r - 1, "ovhi", "hi"), ("hi", R.has_hi, R.r + 1, "ovlo", "lo")):
                        if not has:
                            continue
                        N = ranks[nbr]
                        for b in boxes:
                            N.write((tgt, out, b), m + 1, f"rank {R.r} push ({c},{m})")
                        if c < 0:
                            N.whole[cside] = it + 1
                        else:
                            N.box[cside][c] = it + 1
                    R.ev.add(("pushed", key, it))
                R.q["copy"].append(Op(lambda R=R, c=c, m=m: ("halo", c, m) in R.ev, preads, push, f"push ({c},{m})"))
                R.pushed_it[(key, it & 1)] = it
    return ranks


def run(world, nbox, n_it, pipelined, seed, base=0):
    ranks = build(world, nbox, n_it, pipelined, base)
    rng = random.Random(seed)
    n_ops = 0
    while any(q for R in ranks for q in R.q.values()):
        heads = [(R, name) for R in ranks for name, q in R.q.items() if q]
        rng.shuffle(heads)
        for R, name in heads:
            op = R.q[name][0]
            if op.ready():
                R.q[name].pop(0)                        # (leaves the queue before it runs: it is not its own pending reader)
                op.action()
                n_ops += 1
                break
        else:
            raise AssertionError(f"dead-lock after {n_ops} operations: " +
                                 "; ".join(f"rank {R.r} {n}: {q[0].label}" for R in ranks for n, q in R.q.items() if q))
    final = (base + n_it) & 1
    for R in ranks:
        assert [R.get(("own", final, b)) for b in range(nbox)] == [n_it] * nbox, (R.r, R.buf)


@pytest.mark.parametrize("world", [1, 2, 3, 5])
@pytest.mark.parametrize("nbox,n_it,pipelined", [(1, 7, False), (1, 2, False), (4, 3, True), (4, 9, True), (4, 14, True),
                                                 (6, 40, True), (16, 70, True), (3, 1, True)])
def test_engine_protocol_random_schedules(world, nbox, n_it, pipelined):
    for seed in range(5):
        run(world, nbox, n_it, pipelined, seed)
    run(world, nbox, n_it, pipelined, 99, base=5)       # after a re-load: counters keep counting, the sets start at odd parity


def test_engine_protocol_model_has_teeth():
    """Weaken the protocol in the model and it must fail: (1) a box step that waits for its own box's counter only
    (the halo planes read one row of each neighbouring box); (2) no counter waits at all."""
    src = open(__file__).read()
    for old, new, args in (
            ("all(R.whole[side] >= it or R.box[side][b] >= it for b in nb)",
             "all(R.whole[side] >= it or R.box[side][b] >= it for b in ([nb[len(nb) // 2]] if len(nb) == 3 else nb))", (3, 4, 9, True)),
            ("if it > 0:                                  # counters", "if False:                                   # counters", (2, 1, 7, False))):
        assert old in src
        ns = {"__file__": __file__}
        exec(compile(src.replace(old, new), "mutated_model", "exec"), ns)
        failures = 0
        for seed in range(8):
            try:
                ns["run"](*args, seed)
            except AssertionError:
                failures += 1
        assert failures == 8, (new, failures)
