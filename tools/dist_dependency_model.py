"""Dependency model of the PIPELINED distributed Cholesky schedule (fit_dist_impl, dist_sched == 2 in csrc/engine.cu).

Every rank enqueues the same program on three in-order streams (s = main / panel chain, s2 = rest updates, scm = panel
broadcasts) with event edges between them.  This model replays that enqueue order op by op -- which memory objects each op
reads / writes, which events it waits for and records -- builds the happens-before relation (stream order + event edges +
"a receive completes after the owner's send was enqueued behind its pack"), and reports every pair of ops ON THE SAME RANK
that touch the same object, at least one writing, without being ordered.  It is the design check that would have caught the
round-2 race (side-stream rest update of step k-1 vs main-stream block-column update of step k on the next owner):
`check(R, nto, split_first=False)` reports it, the shipped schedule (`split_first=True`) is clean.

Objects:  ("col", j)   local block column j (the rank that owns j)
          ("P", b, g)  column piece g of packed-panel buffer b = k % 3
          ("ws", b)    slice buffer b = k & 1
Pure Python, no GPU.  tests/test_dist_dependency_model.py runs it for R in {2, 3, 4, 8}.
"""
from __future__ import annotations

import itertools
from collections import defaultdict


class Rank:
    def __init__(self, me, R, nto, G, split_first, defer, use_oz=True):
        self.me, self.R, self.nto, self.G, self.split_first, self.defer = me, R, nto, G, split_first, defer
        self.use_oz = use_oz   # False: DMMA / FFMA trailing updates read the packed panel itself, nothing is sliced
        self.ops = []            # (name, stream, reads, writes, waits[event ids], record event id or None)
        self.next_event = 0

    def ev(self):
        self.next_event += 1
        return (self.me, self.next_event)

    def op(self, name, stream, reads=(), writes=(), waits=(), record=None):
        self.ops.append(dict(name=name, stream=stream, reads=set(reads), writes=set(writes), waits=list(waits), record=record))

    def local_cols_after(self, kk):
        return [j for j in range(kk + 1, self.nto) if j % self.R == self.me]

    def program(self):
        me, R, nto, G = self.me, self.R, self.nto, self.G
        e_rest, e_first = {}, {}
        deferred = None
        pending_waits = defaultdict(list)   # stream -> events to attach to the next op on that stream

        def wait(stream, e):
            if e is not None:
                pending_waits[stream].append(e)

        def emit(name, stream, reads=(), writes=(), record=None):
            w = pending_waits.pop(stream, [])
            self.op(name, stream, reads, writes, w, record)

        def operand(kk):  # what a trailing update of step kk reads
            return [("ws", kk & 1)] if self.use_oz else [("P", kk % 3, g) for g in range(G)]

        def issue_rest(kk, cols):
            e_rest[kk] = self.ev()
            if self.split_first and kk + 2 < nto and (kk + 2) % R == me and cols:
                e_first[kk] = self.ev()
                emit("rest%d:first" % kk, "s2", reads=operand(kk), writes=[("col", cols[0])], record=e_first[kk])
                emit("rest%d" % kk, "s2", reads=operand(kk), writes=[("col", j) for j in cols[1:]], record=e_rest[kk])
            else:
                emit("rest%d" % kk, "s2", reads=operand(kk), writes=[("col", j) for j in cols], record=e_rest[kk])

        e_start = self.ev()
        emit("gram", "s", writes=[("col", j) for j in range(nto) if j % R == me], record=e_start)
        wait("scm", e_start)
        wait("s2", e_start)
        for kk in range(nto):
            owner = kk % R
            b3 = kk % 3
            if kk >= 3:
                wait("scm", e_rest.get(kk - 3))
            if owner == me:
                if kk >= 3:
                    wait("s", e_rest.get(kk - 3))
                e_col = None
                for g in range(G):
                    emit("factor%d.%d" % (kk, g), "s", reads=[("col", kk)], writes=[("col", kk)])
                    e_col = self.ev()
                    emit("pack%d.%d" % (kk, g), "s", reads=[("col", kk)], writes=[("P", b3, g)], record=e_col)
                    wait("scm", e_col)
                    emit("send%d.%d" % (kk, g), "scm", reads=[("P", b3, g)])
                if deferred is not None:
                    e_fact = self.ev()
                    emit("mark_fact%d" % kk, "s", record=e_fact)
                    wait("s2", e_fact)
                    issue_rest(*deferred)
                    deferred = None
            else:
                for g in range(G):
                    emit("recv%d.%d" % (kk, g), "scm", writes=[("P", b3, g)])
            if kk == nto - 1:
                break
            e_recv = self.ev()
            emit("mark_recv%d" % kk, "scm", record=e_recv)
            wait("s", e_recv)
            if kk >= 2:
                wait("s", e_rest.get(kk - 2))
            e_prep = self.ev()
            if self.use_oz:
                emit("prepare%d" % kk, "s", reads=[("P", b3, g) for g in range(G)], writes=[("ws", kk & 1)], record=e_prep)
            else:
                emit("mark_prep%d" % kk, "s", record=e_prep)
            cols = self.local_cols_after(kk)
            next_is_mine = (kk + 1) % R == me
            if next_is_mine:
                if kk >= 1:
                    wait("s", e_first.get(kk - 1) if self.split_first else None)
                emit("colupd%d" % kk, "s", reads=operand(kk), writes=[("col", kk + 1)])
                cols = cols[1:]
            if next_is_mine and self.defer:
                deferred = (kk, cols)
            else:
                wait("s2", e_prep)
                issue_rest(kk, cols)
        e_s2, e_cm = self.ev(), self.ev()
        emit("mark_s2", "s2", record=e_s2)
        emit("mark_cm", "scm", record=e_cm)
        wait("s", e_s2)
        wait("s", e_cm)
        emit("join", "s")
        return self.ops


def check(R, nto, G=4, split_first=True, defer=True, use_oz=True):
    """returns the list of unordered conflicting op pairs (rank, op a, op b, object); empty = schedule is race-free"""
    ranks = [Rank(me, R, nto, G, split_first, defer, use_oz) for me in range(R)]
    progs = [r.program() for r in ranks]
    # node ids: (rank, index); edges: stream order, event record -> waiter, owner's send -> every peer's matching recv
    nodes = [(r, i) for r in range(R) for i in range(len(progs[r]))]
    succ = defaultdict(list)
    recorder = {}
    for r in range(R):
        last = {}
        for i, o in enumerate(progs[r]):
            if o["stream"] in last:
                succ[(r, last[o["stream"]])].append((r, i))
            last[o["stream"]] = i
            if o["record"] is not None:
                recorder[o["record"]] = (r, i)
    sends = {}
    for r in range(R):
        for i, o in enumerate(progs[r]):
            for e in o["waits"]:
                succ[recorder[e]].append((r, i))
            if o["name"].startswith("send"):
                sends[o["name"][4:]] = (r, i)
    for r in range(R):
        for i, o in enumerate(progs[r]):
            if o["name"].startswith("recv"):
                succ[sends[o["name"][4:]]].append((r, i))
    # reachability per rank-local pairs: ancestors via DFS from each node restricted to what we need (small graphs)
    order = {n: k for k, n in enumerate(nodes)}
    reach = {}

    def reachable(a):
        if a in reach:
            return reach[a]
        seen, stack = set(), [a]
        while stack:
            n = stack.pop()
            for m in succ.get(n, ()):
                if m not in seen:
                    seen.add(m)
                    stack.append(m)
        reach[a] = seen
        return seen

    bad = []
    for r in range(R):
        by_obj = defaultdict(list)
        for i, o in enumerate(progs[r]):
            for x in o["reads"]:
                by_obj[x].append((i, False))
            for x in o["writes"]:
                by_obj[x].append((i, True))
        for obj, acc in by_obj.items():
            for (i, wi), (j, wj) in itertools.combinations(acc, 2):
                if i == j or not (wi or wj):
                    continue
                a, b = (r, i), (r, j)
                if b not in reachable(a) and a not in reachable(b):
                    bad.append((r, progs[r][i]["name"], progs[r][j]["name"], obj))
    return bad


if __name__ == "__main__":
    for R, nto in ((2, 9), (4, 11), (8, 16), (8, 20)):
        for split in (False, True):
            b = check(R, nto, split_first=split)
            print("R=%d nto=%d split_first=%s: %d unordered conflicts%s" % (R, nto, split, len(b), (" e.g. %s" % (b[0],)) if b else ""))
