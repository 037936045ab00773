"""Developer tool: discrete-event model of the mbarrier protocol of body2_umma_kernel (csrc/body2_umma.cuh) for ONE CTA.

Every role of the kernel (TMA warp, the two MMA issuers, the SE warp, 8 epilogue warps) is a coroutine that follows the
kernel's waits / arrivals / commits literally (same tables, same parities); TMA landings, MMA completions, peer flags and
per-step delays are random and heavy-tailed.  The model checks, at every MMA issue and at every TMA issue, what the
hardware cannot: that each ring slot a tile reads holds exactly the box the tile expects (landed, not being overwritten
until the tile's MMAs complete), that the weights in shared memory are the layer's, and that accumulators are handed over
in order.  A violation means the barrier protocol itself admits a wrong tile (phase aliasing, early recycling);
    python tools/body2_protocol_sim.py [seeds] [layers]
"""
import heapq, random, sys

kTileM, kPitch, kMaxShift, kBoxRows, kBoxPx = 128, 66, 134, 2, 132
kSlots, kAcc, kPre = 7, 7, 4
SE_SELF = True   # the SE warp issues its own MMAs and is a party of the weight hand-back (FEN_B2_SE_SELF)
TURN = True     # the issuers take turns, one whole tile each (FEN_B2_TURN)
ROTATE = False   # the issuers' tile shares rotate per pass (FEN_B2_ROTATE)
FIX = "both"   # "both": two alternating mbarriers per ring slot + a shared count of requested boxes (the kernel); "two" / "count": either alone; "nofix": the protocol as it was (reproduces the aliasing)
H, TPS = 64, 33
import os
C, SET_B = 148, int(os.environ.get("SET_B", "32"))   # images per set (batch 64 = 2 x 32)
T = SET_B * TPS
C = min(C, T)                 # (the host launches one CTA per tile when there are fewer tiles than SMs)
RQ, RR = T // C, T % C


def run_begin(c, s):
    return c * RQ + min(c, RR) if s == 0 else c * RQ + max(0, c - (C - RR))


def tables(c, s):
    g, g_end = run_begin(c, s), run_begin(c + 1, s)
    tiles, boxes, b_cum, i = [], [], 0, 0
    while g < g_end:
        n = g // TPS; t0 = g - n * TPS; t1 = min(TPS, t0 + (g_end - g))
        ra = (kTileM * t0) // kPitch
        rb = min((kTileM * t1 + kMaxShift - 1) // kPitch, H + 1)
        nb = (rb - ra) // kBoxRows + 1
        for j in range(nb):
            boxes.append(dict(mirror=j > 0, last_tile=i))
        for t in range(t0, t1):
            base = kTileM * t - kPitch * ra
            fb = b_cum + base // kBoxPx
            wu = b_cum + min((base + kTileM + kMaxShift - 1) // kBoxPx, nb - 1) + 1
            tiles.append(dict(first_box=fb, wait_upto=wu))
            for b in range(fb, wu):
                boxes[b]["last_tile"] = i
            i += 1
        b_cum += nb
        g += t1 - t0
    return tiles, boxes


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.ph, self.tx = count, count, 0, 0

    def test(self, parity):
        return (self.ph & 1) != parity

    def arrive(self, n=1):
        self.pending -= n
        self._check()

    def expect_tx(self, nbytes):      # arrive + expect
        self.tx += nbytes
        self.pending -= 1
        self._check()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._check()

    def _check(self):
        assert self.pending >= 0, "more arrivals than the barrier expects in one phase"
        if self.pending == 0 and self.tx == 0:
            self.ph += 1
            self.pending = self.count


class Violation(Exception):
    pass


class Sim:
    def __init__(self, cta, n_layers, seed, slow):
        self.rng = random.Random(seed)
        self.now = 0.0
        self.q = []
        self.seq = 0
        self.cta, self.NL, self.slow = cta, n_layers, slow
        self.tab = [tables(cta, 0), tables(cta, 1)]
        self.bar_w = [Bar(1) for _ in range(9)]
        self.n_issuers = min(2, max(len(self.tab[0][0]), len(self.tab[1][0]), 1))
        self.bar_wfree = Bar(self.n_issuers + (1 if SE_SELF else 0))
        self.bar_full = [Bar(1) for _ in range(2 * kSlots)]
        self.nb = 2 * kSlots if FIX in ("two", "both") else kSlots   # barriers in use
        self.bar_acc_full = [Bar(1) for _ in range(kAcc)]
        self.bar_acc_empty = [Bar(8) for _ in range(kAcc)]
        self.bar_done = Bar(8)
        self.bar_s_ready, self.bar_s_free, self.bar_se_full, self.bar_se_empty = Bar(1), Bar(1), Bar(1), Bar(1)
        self.bar_scale = [Bar(1), Bar(1)]
        self.bar_turn = [Bar(1) for _ in range(4)]
        self.s_hist = [-1] * kSlots
        self.turn_next = 0
        self.issued = 0                          # boxes issued so far (the fix: s_issued in shared memory)
        # ground truth
        self.slot_box = [None] * kSlots          # global box id whose data is in the slot
        self.slot_inflight = [False] * kSlots
        self.readers = [set() for _ in range(kSlots)]   # incomplete tiles reading the slot
        self.w_layer = [None] * 9
        self.w_inflight = [False] * 9
        self.w_readers = set()
        self.acc_tile = [None] * kAcc            # tile whose result is in the accumulator, (G, complete?, reads left)
        self.own_flag = [0, 0]                   # this CTA's flags per set
        self.last_done = [0.0, 0.0]              # completion time of the last MMA per issuer (in-order completion per thread)

    def conv2(self, L):
        return L % 2 == 1

    # ---- event loop
    def at(self, t, fn):
        self.seq += 1
        heapq.heappush(self.q, (t, self.seq, fn))

    def spawn(self, gen):
        def step():
            try:
                op = next(gen)
            except StopIteration:
                return
            if op[0] == "delay":
                self.at(self.now + op[1], step)
            else:                       # ("wait", predicate)
                def poll():
                    if op[1]():
                        self.at(self.now + self.jit(5), step)
                    else:
                        self.at(self.now + 20 + self.jit(20), poll)
                poll()
        self.at(self.now, step)

    def jit(self, scale):
        return self.rng.random() * scale

    def tail(self, base, p=0.05, mult=40):
        r = self.rng.random()
        return base * (1 + self.rng.random()) * (mult * self.rng.random() if r < p * self.slow else 1)

    def run(self):
        self.spawn(self.tma())
        self.spawn(self.issuer(0))
        if self.n_issuers == 2:
            self.spawn(self.issuer(1))
        self.spawn(self.se_warp())
        for w in range(8):
            self.spawn(self.epilogue(w))
        n = 0
        while self.q:
            t, _, fn = heapq.heappop(self.q)
            self.now = t
            fn()
            n += 1
            if n > 5_000_000:
                raise Violation("event budget exceeded (deadlock / livelock?)")
        if not self.finished:
            raise Violation("deadlock: event queue drained before the last layer completed")

    finished = False

    # ---- roles
    def wait_flags(self, L, s):
        if L == 0:
            return
        # peers finish layer L-1 of set s some time after this CTA could; own flag must be there too
        yield ("wait", lambda: self.own_flag[s] >= L)
        yield ("delay", self.tail(200, p=0.3, mult=60))

    def tma(self):
        slot, gbase, frontier = 0, 0, 0
        gbox = 0
        for L in range(self.NL):
            def issue_boxes(s, b0, b1, gbase_pass, gbox_pass):
                nonlocal slot, frontier
                for b in range(b0, b1):
                    e = self.tab[s][1][b]
                    need = self.s_hist[slot]
                    while frontier <= need:
                        f = frontier
                        yield ("wait", lambda f=f: self.bar_acc_full[f % kAcc].test((f // kAcc) & 1))
                        frontier += 1
                    if self.readers[slot]:
                        raise Violation(f"L{L} set {s}: box {b} overwrites slot {slot} still read by tiles {sorted(self.readers[slot])}")
                    self.s_hist[slot] = gbase_pass + e["last_tile"]
                    g = gbox_pass + b
                    assert g % kSlots == slot
                    self.bar_full[g % self.nb].expect_tx(1)
                    self.issued = g + 1
                    self.slot_inflight[slot] = True
                    k = slot
                    def land(k=k, g=g):
                        self.slot_box[k] = g
                        self.slot_inflight[k] = False
                        self.bar_full[g % self.nb].complete_tx(1)
                    self.at(self.now + self.tail(900, p=0.1, mult=30), land)
                    yield ("delay", 30 + self.jit(30))
                    slot = (slot + 1) % kSlots
            pre = min(kPre, len(self.tab[0][1]))
            yield from self.wait_flags(L, 0)
            yield from issue_boxes(0, 0, pre, gbase, gbox)
            for tap in range(9):
                if L > 0 and tap == 0:
                    yield ("wait", lambda: self.bar_wfree.test((L - 1) & 1))
                if self.w_readers:
                    raise Violation(f"L{L}: weights of tap {tap} overwritten while tiles {sorted(self.w_readers)} still read them")
                self.bar_w[tap].expect_tx(1)
                self.w_inflight[tap] = True
                def wland(tap=tap, L=L):
                    self.w_layer[tap] = L
                    self.w_inflight[tap] = False
                    self.bar_w[tap].complete_tx(1)
                self.at(self.now + self.tail(700, p=0.1, mult=10), wland)
                yield ("delay", 30)
            for s in range(2):
                if s > 0:
                    yield from self.wait_flags(L, s)
                yield from issue_boxes(s, pre if s == 0 else 0, len(self.tab[s][1]), gbase, gbox)
                gbase += len(self.tab[s][0])
                gbox += len(self.tab[s][1])

    def issuer(self, wi):
        gbase = gbox = se_n = 0
        for L in range(self.NL):
            conv2 = self.conv2(L)
            w_seen = False
            se_done = 0
            for s in range(2):
                tiles, boxes = self.tab[s]
                n_tiles, n_boxes = len(tiles), len(boxes)
                rot = (1 if conv2 else s & 1) if (ROTATE and self.n_issuers == 2 and n_tiles >= 2) else 0      # FEN_B2_ROTATE
                w0 = (wi + rot) % self.n_issuers
                ni = self.n_issuers
                last_own = w0 + ni * ((n_tiles - 1 - w0) // ni) if n_tiles - 1 - w0 >= 0 else -1
                gb0, gbase_pass = gbox, gbase
                gbox += n_boxes; gbase += n_tiles
                last_pass = s == 1
                se_layer = conv2 and wi == 0 and not SE_SELF
                if s == 0:
                    se_done = 0
                waited = 0

                def issue_se():
                    nonlocal w_seen, se_n, se_done
                    if se_n >= 1:
                        n0 = se_n
                        yield ("wait", lambda: self.bar_se_empty.test((n0 - 1) & 1))
                    if not w_seen:
                        for tap in range(9):
                            yield ("wait", lambda tap=tap: self.bar_w[tap].test(L & 1))
                    if any(self.w_layer[t] != L or self.w_inflight[t] for t in range(9)):
                        raise Violation(f"L{L}: SE batch reads weights {self.w_layer}")
                    yield ("delay", 36 * 60)
                    done = max(self.now + 400, self.last_done[wi]) + self.jit(200)
                    self.last_done[wi] = done
                    tag = ("se", L, s, se_n)
                    self.w_readers.add(tag)
                    def fin():
                        self.w_readers.discard(tag)
                        self.bar_se_full.arrive(); self.bar_s_free.arrive()
                    self.at(done, fin)
                    w_seen = True
                    se_n += 1; se_done += 1

                if last_own < 0:          # no tile for this issuer in this pass: only the weight hand-back, in the last pass
                    if last_pass:
                        self.at(max(self.now, self.last_done[wi]), self.bar_wfree.arrive)
                    continue
                for i in range(w0, n_tiles, ni):
                    e = tiles[i]
                    G = gbase_pass + i; acc = G % kAcc; aph = (G // kAcc) & 1
                    if se_layer and se_done <= s:
                        if i == last_own:
                            n0 = se_n
                            yield ("wait", lambda: self.bar_s_ready.test(n0 & 1))
                            yield from issue_se()
                        else:
                            while True:
                                yield ("delay", 10)
                                if w_seen and self.bar_s_ready.test(se_n & 1):
                                    yield from issue_se(); break
                                if self.bar_acc_empty[acc].test(aph ^ 1):
                                    break
                    if se_layer and se_done > s and se_done < 2 and w_seen and self.bar_s_ready.test(se_n & 1):
                        yield from issue_se()
                    yield ("wait", lambda: self.bar_acc_empty[acc].test(aph ^ 1))
                    waited = max(waited, e["first_box"])
                    while waited < e["wait_upto"]:
                        g = gb0 + waited
                        if FIX in ("count", "both"):
                            yield ("wait", lambda g=g: self.issued > g)
                        yield ("wait", lambda g=g: self.bar_full[g % self.nb].test((g // self.nb) & 1))
                        waited += 1
                    if TURN and G > 0:
                        yield ("wait", lambda G=G: self.bar_turn[(G - 1) & 3].test(((G - 1) >> 2) & 1))
                    if not w_seen:
                        for tap in range(9):
                            yield ("wait", lambda tap=tap: self.bar_w[tap].test(L & 1))
                    # ---- ground-truth checks at issue
                    for b in range(e["first_box"], e["wait_upto"]):
                        g = gb0 + b; k = g % kSlots
                        if self.slot_box[k] != g or self.slot_inflight[k]:
                            raise Violation(f"L{L} set {s} tile {i} (issuer {wi}): box {b} (global {g}) expected in slot {k}, "
                                            f"slot holds {self.slot_box[k]} inflight={self.slot_inflight[k]}")
                    if any(self.w_layer[t] != L or self.w_inflight[t] for t in range(9)):
                        raise Violation(f"L{L} set {s} tile {i} (issuer {wi}): weights in smem are {self.w_layer}")
                    if self.acc_tile[acc] is not None:
                        raise Violation(f"L{L} set {s} tile {i}: accumulator {acc} still holds tile {self.acc_tile[acc]}")
                    self.acc_tile[acc] = [G, False, 8]
                    tag = (L, s, i)
                    for b in range(e["first_box"], e["wait_upto"]):
                        self.readers[(gb0 + b) % kSlots].add(tag)
                    self.w_readers.add(tag)
                    yield ("delay", 36 * (50 + self.jit(40)))
                    if TURN:
                        if self.turn_next != G:
                            raise Violation(f"L{L} set {s} tile {i}: issued out of turn (expected global tile {self.turn_next}, got {G})")
                        self.turn_next = G + 1
                        self.bar_turn[G & 3].arrive()
                    done = max(self.now + self.tail(300, p=0.05, mult=20), self.last_done[wi] + 50)
                    self.last_done[wi] = done
                    w_rel = last_pass and i == last_own
                    def fin(tag=tag, e=e, gb0=gb0, acc=acc, G=G, w_rel=w_rel):
                        for b in range(e["first_box"], e["wait_upto"]):
                            self.readers[(gb0 + b) % kSlots].discard(tag)
                        self.w_readers.discard(tag)
                        assert self.acc_tile[acc][0] == G
                        self.acc_tile[acc][1] = True
                        if w_rel:
                            self.bar_wfree.arrive()
                        self.bar_acc_full[acc].arrive()
                    self.at(done, fin)
                    w_seen = True

    def se_warp(self):
        se_n = 0
        for L in range(self.NL):
            if not self.conv2(L):
                if SE_SELF:
                    if L > 0:
                        yield ("wait", lambda: self.bar_wfree.test((L - 1) & 1))
                    self.bar_wfree.arrive()
                continue
            for s in range(2):
                yield from self.wait_flags(L, s)
                yield ("delay", self.tail(600, p=0.3, mult=30))
                if se_n >= 1 and not SE_SELF:
                    n0 = se_n
                    yield ("wait", lambda: self.bar_s_free.test((n0 - 1) & 1))
                yield ("delay", 300)
                if SE_SELF:
                    for tap in range(9):
                        yield ("wait", lambda tap=tap: self.bar_w[tap].test(L & 1))
                    if any(self.w_layer[t] != L or self.w_inflight[t] for t in range(9)):
                        raise Violation(f"L{L}: SE batch reads weights {self.w_layer}")
                    yield ("delay", 36 * 55)
                    tag = ("se", L, s)
                    self.w_readers.add(tag)
                    def fin(tag=tag, last=(s == 1)):
                        self.w_readers.discard(tag)
                        self.bar_se_full.arrive()
                        if last:
                            self.bar_wfree.arrive()
                    self.at(self.now + 400 + self.tail(300, p=0.1, mult=20), fin)
                else:
                    self.bar_s_ready.arrive()
                n0 = se_n
                yield ("wait", lambda: self.bar_se_full.test(n0 & 1))
                yield ("delay", 200)
                self.bar_se_empty.arrive()
                yield ("delay", 1500 + self.jit(500))
                self.bar_scale[s].arrive()
                se_n += 1

    def epilogue(self, ew):
        G = P = m_cnt = 0
        for L in range(self.NL):
            for s in range(2):
                n_tiles = len(self.tab[s][0])
                if self.conv2(L):
                    m0 = m_cnt
                    yield ("wait", lambda: self.bar_scale[s].test(m0 & 1))
                for i in range(n_tiles):
                    acc = G % kAcc; aph = (G // kAcc) & 1
                    yield ("wait", lambda: self.bar_acc_full[acc].test(aph))
                    a = self.acc_tile[acc]
                    if a is None or a[0] != G or not a[1]:
                        raise Violation(f"L{L} set {s} tile {i}: epilogue warp {ew} reads accumulator {acc} holding {a}, expected tile {G}")
                    yield ("delay", 150 + self.jit(100))
                    a[2] -= 1
                    if a[2] == 0:
                        self.acc_tile[acc] = None
                    self.bar_acc_empty[acc].arrive()
                    yield ("delay", self.tail(900, p=0.05, mult=10))
                    G += 1
                if P > 0:
                    p0 = P
                    yield ("wait", lambda: self.bar_done.test((p0 - 1) & 1))
                self.bar_done.arrive()
                if ew == 0:
                    p0 = P
                    yield ("wait", lambda: self.bar_done.test(p0 & 1))
                    self.own_flag[s] = L + 1
                    if L == self.NL - 1 and s == 1:
                        self.finished = True
                P += 1
            if self.conv2(L):
                m_cnt += 1


def sweep(ctas, seeds, layers, mode, regimes=(1, 4, 16), turn=True, rotate=False, se_self=True):
    """Runs the model for every (CTA, seed, timing regime); returns the violation messages."""
    global FIX, TURN, ROTATE, SE_SELF
    FIX, TURN, ROTATE, SE_SELF = mode, turn, rotate, se_self
    out = []
    for cta in ctas:
        for seed in range(seeds):
            for slow in regimes:
                try:
                    Sim(cta, layers, seed * 7919 + cta, slow).run()
                except Violation as v:
                    out.append(f"CTA {cta} seed {seed} slow {slow}: {v}")
    return out


def main():
    seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    layers = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    mode = sys.argv[3] if len(sys.argv) > 3 else FIX
    ctas = [c for c in [0, 1, 19, 20, 21, 70, 127, 128, 129, 147] + list(range(2, 148, 9)) if c < C]
    bad = sweep(ctas, seeds, layers, mode, se_self=bool(int(os.environ.get("SE_SELF", "1"))))
    for line in bad[:12]:
        print(line)
    print(f"{mode}: {len(ctas)} CTAs x {seeds} seeds x 3 timing regimes x {layers} layers: {len(bad)} violations")


if __name__ == "__main__":
    main()
