#!/usr/bin/env python3
"""Generate gmat_b200/csrc/median5_nets.inc: the min/max blocks of the streaming 5x5 median
(median5_stream.cuh), as straight-line SSA code on u16x2 lanes with 3-input min/max fusion.

  med5_merge55 : two sorted columns of 5  -> sorted 10            (pruned Batcher odd-even merge)
  med5_select6 : two sorted 10s (the 4 columns two neighbouring outputs share) -> ranks 7..12 of the 20, sorted
  med5_final   : those 6 candidates + the sorted 5th column of an output -> its median of 25

Why ranks 7..12: an element of rank r among the 20 shared samples has rank r..r+5 among an output's 25, so only
r = 7..12 can be the median (rank 12); dropping 7 below and 7 above leaves the median of 11 = rank 5 of
(6 candidates + 5 own samples), and the k-th of two sorted lists is min over splits of max(prefix ends).
Every block is verified here exhaustively on 0-1 inputs (the 0-1 principle holds for min/max circuits with
sorted-input preconditions) and on random bytes."""
import itertools, random, sys


def batcher(np2):
    out = []
    p = 1
    while p < np2:
        k = p
        while k >= 1:
            j = k % p
            while j + k < np2:
                for i in range(k):
                    if i + j + k < np2 and (i + j) // (2 * p) == (i + j + k) // (2 * p):
                        out.append((p, i + j, i + j + k))
                j += 2 * k
            k //= 2
        p *= 2
    return out


class Net:
    """SSA min/max program."""
    def __init__(self):
        self.ops = []          # (dst, kind, srcs)
        self.n = 0

    def new(self, kind, *srcs):
        d = f"t{self.n}"; self.n += 1
        self.ops.append((d, kind, list(srcs)))
        return d

    def optimise(self, outs):
        # dead code elimination
        need = set(outs); keep = []
        for d, k, s in reversed(self.ops):
            if d in need:
                keep.append((d, k, s)); need.update(s)
        keep.reverse()
        # 3-input fusion: a single-use min feeding a 2-input min (same for max)
        changed = True
        while changed:
            changed = False
            uses = {}
            for d, k, s in keep:
                for x in s: uses[x] = uses.get(x, 0) + 1
            for o in outs: uses[o] = uses.get(o, 0) + 1
            defs = {d: (k, s) for d, k, s in keep}
            for idx, (d, k, s) in enumerate(keep):
                if len(s) != 2: continue
                for pos in (0, 1):
                    x = s[pos]
                    if x in defs and defs[x][0] == k and len(defs[x][1]) == 2 and uses[x] == 1:
                        keep[idx] = (d, k, defs[x][1] + [s[1 - pos]])
                        keep = [o for o in keep if o[0] != x]
                        changed = True
                        break
                if changed: break
        self.ops = keep
        return self

    def run(self, env):
        env = dict(env)
        for d, k, s in self.ops:
            env[d] = (min if k == "min" else max)(env[x] for x in s)
        return env

    def emit(self):
        f = {("min", 2): "mmin2", ("max", 2): "mmax2", ("min", 3): "mmin3", ("max", 3): "mmax3"}
        return [f"const unsigned {d} = {f[(k, len(s))]}({', '.join(s)});" for d, k, s in self.ops]


def merge_net(na, nb, np2, want):
    """last phase of Batcher's sort on np2 wires: A on wires 0.., B on wires np2/2.., pads are +inf."""
    net = Net()
    half = np2 // 2
    wire = {i: f"a[{i}]" for i in range(na)}
    wire.update({half + i: f"b[{i}]" for i in range(nb)})
    for p, i, j in batcher(np2):
        if p != half: continue
        ri, rj = i in wire, j in wire
        if ri and rj:
            lo = net.new("min", wire[i], wire[j]); hi = net.new("max", wire[i], wire[j])
            wire[i], wire[j] = lo, hi
        elif rj and not ri:
            wire[i] = wire.pop(j)
    outs = [wire[k] for k in want]
    net.optimise(outs)
    return net, outs


def final_net():
    net = Net()
    t = [net.new("max", f"c[{i}]", f"e[{4 - i}]") for i in range(5)] + ["c[5]"]
    m = net.new("min", net.new("min", net.new("min", t[0], t[1]), t[2]), net.new("min", net.new("min", t[3], t[4]), t[5]))
    net.optimise([m])
    return net, [m]


def sorted01(n):
    return [[0] * (n - k) + [1] * k for k in range(n + 1)]


def check(name, net, outs, na, nb, an, bn, ref):
    for A in sorted01(na):
        for B in sorted01(nb):
            env = {f"{an}[{i}]": A[i] for i in range(na)}; env.update({f"{bn}[{i}]": B[i] for i in range(nb)})
            r = net.run(env)
            assert [r.get(o, env.get(o)) for o in outs] == ref(A, B), (name, A, B)
    for _ in range(2000):
        A = sorted(random.randrange(256) for _ in range(na)); B = sorted(random.randrange(256) for _ in range(nb))
        env = {f"{an}[{i}]": A[i] for i in range(na)}; env.update({f"{bn}[{i}]": B[i] for i in range(nb)})
        r = net.run(env)
        assert [r.get(o, env.get(o)) for o in outs] == ref(A, B), (name, A, B)


def main(path):
    m55, o55 = merge_net(5, 5, 16, range(10))
    check("merge55", m55, o55, 5, 5, "a", "b", lambda A, B: sorted(A + B))
    s6, o6 = merge_net(10, 10, 32, range(7, 13))
    check("select6", s6, o6, 10, 10, "a", "b", lambda A, B: sorted(A + B)[7:13])
    fn, of = final_net()
    check("final", fn, of, 6, 5, "c", "e", lambda C, E: [sorted(C + E)[5]])
    # end to end: median of 25 through the pair structure
    for _ in range(3000):
        cols = [sorted(random.choice([random.randrange(256), random.randrange(4)]) for _ in range(5)) for _ in range(6)]
        def ev(net, outs, an, A, bn, B):
            env = {f"{an}[{i}]": A[i] for i in range(len(A))}; env.update({f"{bn}[{i}]": B[i] for i in range(len(B))})
            r = net.run(env); return [r.get(o, env.get(o)) for o in outs]
        P = ev(m55, o55, "a", cols[1], "b", cols[2]); Q = ev(m55, o55, "a", cols[3], "b", cols[4])
        C = ev(s6, o6, "a", P, "b", Q)
        ml = ev(fn, of, "c", C, "e", cols[0])[0]; mr = ev(fn, of, "c", C, "e", cols[5])[0]
        assert ml == sorted(sum(cols[0:5], []))[12] and mr == sorted(sum(cols[1:6], []))[12]
    with open(path, "w") as f:
        f.write("// generated by tools/gen_median5_stream.py -- do not edit\n")
        for name, net, outs, sig, outname in (
                ("med5_merge55", m55, o55, "const unsigned (&a)[5], const unsigned (&b)[5], unsigned (&o)[10]", "o"),
                ("med5_select6", s6, o6, "const unsigned (&a)[10], const unsigned (&b)[10], unsigned (&o)[6]", "o"),
                ("med5_final", fn, of, "const unsigned (&c)[6], const unsigned (&e)[5], unsigned (&o)[1]", "o")):
            f.write(f"// {name}: {len(net.ops)} min/max instructions\n")
            f.write(f"__device__ __forceinline__ void {name}({sig}) {{\n")
            for l in net.emit(): f.write("    " + l + "\n")
            for i, o in enumerate(outs): f.write(f"    {outname}[{i}] = {o};\n")
            f.write("}\n\n")
            print(name, len(net.ops), "instructions")


if __name__ == "__main__":
    main(sys.argv[1])
