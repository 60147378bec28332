"""CPU model of k_alloc's tournament-tree argmin (odr_audioenc_b200/csrc/mp2_kernels.cu, "The argmin as a tournament
tree"): the tree must return, after every update, exactly what the reference's scan returns -- the FIRST entry in scan
order holding the smallest value (ref: libtoolame-dab/encode_new.c:1066-1075, `if (small > mnr[..])` is strict), and
"nothing left" when every value is >= 999999.0.  Same bit-field layout as the kernel: 16 x 1, 8 x 2, 4 x 3, 2 x 4 bits."""
import math
import random

INF = math.inf


class Tree:
    def __init__(self, vals):
        self.v = list(vals) + [INF] * (32 - len(vals))
        self.n = len(vals)
        self.w = [0, 0, 0, 0]  # w4, w3, w2, w1 of the kernel
        self.best, self.small = 0, INF
        for e in range(0, 32, 2):  # the kernel's build: one climb per pair of leaves, left to right
            self.best, self.small = self.climb(e, self.val(e))

    def val(self, e):
        return self.v[e] if e < self.n else INF

    def climb(self, e, v):
        cur = e

        def meet(sidx, cur, v):
            sv = self.val(sidx)
            if sv < v or (sv == v and sidx < cur):
                return sidx, sv
            return cur, v

        cur, v = meet(e ^ 1, cur, v)
        for lvl, (bits, shift) in enumerate(((1, 1), (2, 2), (3, 3), (4, 4))):
            node = e >> shift
            mask = (1 << bits) - 1
            self.w[lvl] = (self.w[lvl] & ~(mask << (bits * node))) | ((cur & mask) << (bits * node))
            sib = node ^ 1
            cur, v = meet((sib << shift) + ((self.w[lvl] >> (bits * sib)) & mask), cur, v)
        return cur, v

    def update(self, e, v):
        self.v[e] = v
        self.best, self.small = self.climb(e, v)


def scan(vals):
    small, best = 999999.0, -1
    for i, v in enumerate(vals):
        if small > v:
            small, best = v, i
    return best


def test_tree_equals_first_strictly_smaller_scan():
    rnd = random.Random(7)
    for n in (1, 2, 13, 15, 27, 30, 32):
        for trial in range(60):
            # few distinct values: ties everywhere
            vals = [rnd.choice([-3.5, 0.0, 0.0, 7.0, 11.0, 2e6]) for _ in range(n)]
            t = Tree(vals)
            for step in range(200):
                want = scan(t.v[:n])
                got = t.best if t.small < 999999.0 else -1
                assert got == want, (n, trial, step, t.v[:n])
                if want < 0:
                    break
                # what a round does: the winner's value rises (a grant) or the entry is closed (+inf)
                e = want if rnd.random() < 0.8 else rnd.randrange(n)
                t.update(e, INF if rnd.random() < 0.25 else t.v[e] + rnd.choice([0.0, 4.0, 5.0, 6.16]))
