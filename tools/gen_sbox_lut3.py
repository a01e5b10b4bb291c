#!/usr/bin/env python3
"""gen_sbox_lut3.py -- derive the bitsliced AES S-box used by the ALU co-runner of the CTR kernel.

Input: the public Boyar-Peralta depth-16 straight-line program for the AES S-box (115 two-input
XOR/AND/XNOR gates; J. Boyar, R. Peralta, "A small depth-16 circuit for the AES S-box", 2012),
restated below and checked exhaustively against an S-box computed from the FIPS-197 definition.
Step 2 maps the gate network onto 3-input look-up tables (one LOP3.LUT instruction each on
sm_100a): 3-feasible cut enumeration, area-flow start, annealed local search on the cut choice.
Step 3 simulates the mapped network on all 256 inputs and writes
micro-aes_b200/csrc/uaes_sbox_lut3.cuh.   Run:  python tools/gen_sbox_lut3.py [seeds [output]]
(the committed header is the output of the default, 4 seeds; tests/test_bitslice_host.py regenerates it)
"""
import itertools, math, os, random, sys

# Boyar-Peralta style S-box circuit, U0 = MSB. Verify vs computed AES S-box.
def sbox_table():
    def gmul(a,b):
        r=0
        for i in range(8):
            if b&1: r^=a
            a=((a<<1)^(0x11b if a&0x80 else 0))&0xff
            b>>=1
        return r
    inv=[0]*256
    for a in range(1,256):
        for b in range(1,256):
            if gmul(a,b)==1: inv[a]=b;break
    S=[]
    for x in range(256):
        q=inv[x]; y=q
        for s in (1,2,3,4): y^=((q<<s)|(q>>(8-s)))&0xff
        S.append(y^0x63)
    return S
TOP="""
y14 = U3 ^ U5
y13 = U0 ^ U6
y9 = U0 ^ U3
y8 = U0 ^ U5
t0 = U1 ^ U2
y1 = t0 ^ U7
y4 = y1 ^ U3
y12 = y13 ^ y14
y2 = y1 ^ U0
y5 = y1 ^ U6
y3 = y5 ^ y8
t1 = U4 ^ y12
y15 = t1 ^ U5
y20 = t1 ^ U1
y6 = y15 ^ U7
y10 = y15 ^ t0
y11 = y20 ^ y9
y7 = U7 ^ y11
y17 = y10 ^ y11
y19 = y10 ^ y8
y16 = t0 ^ y11
y21 = y13 ^ y16
y18 = U0 ^ y16
t2 = y12 & y15
t3 = y3 & y6
t4 = t3 ^ t2
t5 = y4 & U7
t6 = t5 ^ t2
t7 = y13 & y16
t8 = y5 & y1
t9 = t8 ^ t7
t10 = y2 & y7
t11 = t10 ^ t7
t12 = y9 & y11
t13 = y14 & y17
t14 = t13 ^ t12
t15 = y8 & y10
t16 = t15 ^ t12
t17 = t4 ^ t14
t18 = t6 ^ t16
t19 = t9 ^ t14
t20 = t11 ^ t16
t21 = t17 ^ y20
t22 = t18 ^ y19
t23 = t19 ^ y21
t24 = t20 ^ y18
t25 = t21 ^ t22
t26 = t21 & t23
t27 = t24 ^ t26
t28 = t25 & t27
t29 = t28 ^ t22
t30 = t23 ^ t24
t31 = t22 ^ t26
t32 = t31 & t30
t33 = t32 ^ t24
t34 = t23 ^ t33
t35 = t27 ^ t33
t36 = t24 & t35
t37 = t36 ^ t34
t38 = t27 ^ t36
t39 = t29 & t38
t40 = t25 ^ t39
t41 = t40 ^ t37
t42 = t29 ^ t33
t43 = t29 ^ t40
t44 = t33 ^ t37
t45 = t42 ^ t41
z0 = t44 & y15
z1 = t37 & y6
z2 = t33 & U7
z3 = t43 & y16
z4 = t40 & y1
z5 = t29 & y7
z6 = t42 & y11
z7 = t45 & y17
z8 = t41 & y10
z9 = t44 & y12
z10 = t37 & y3
z11 = t33 & y4
z12 = t43 & y13
z13 = t40 & y5
z14 = t29 & y2
z15 = t42 & y9
z16 = t45 & y14
z17 = t41 & y8
t46 = z15 ^ z16
t47 = z10 ^ z11
t48 = z5 ^ z13
t49 = z9 ^ z10
t50 = z2 ^ z12
t51 = z2 ^ z5
t52 = z7 ^ z8
t53 = z0 ^ z3
t54 = z6 ^ z7
t55 = z16 ^ z17
t56 = z12 ^ t48
t57 = t50 ^ t53
t58 = z4 ^ t46
t59 = z3 ^ t54
t60 = t46 ^ t57
t61 = z14 ^ t57
t62 = t52 ^ t58
t63 = t49 ^ t58
t64 = z4 ^ t59
t65 = t61 ^ t62
t66 = z1 ^ t63
S0 = t59 ^ t63
S6 = t56 # t62
S7 = t48 # t60
t67 = t64 ^ t65
S3 = t53 ^ t66
S4 = t51 ^ t66
S5 = t47 ^ t65
S1 = t64 # S3
S2 = t55 # t67
"""
def parse():
    g=[]
    for l in TOP.strip().splitlines():
        d,_,a,op,b=l.split()
        g.append((d,op,a,b))
    return g
def evalc(g,x):
    v={f"U{i}":(x>>(7-i))&1 for i in range(8)}
    for d,op,a,b in g:
        A,B=v[a],v[b]
        v[d]= A^B if op=='^' else A&B if op=='&' else 1^A if op=='~' else 1^A^B
    return sum(v[f"S{i}"]<<(7-i) for i in range(8))

K = 3
def build(g):
    nodes={}  # name -> (op,a,b) ; inputs are leaves
    order=[]
    for d,op,a,b in g:
        nodes[d]=(op,a,b); order.append(d)
    return nodes,order
def enum_cuts(nodes,order):
    cuts={}
    def get(n):
        if n not in nodes: return [frozenset([n])]
        return cuts[n]
    for n in order:
        op,a,b=nodes[n]
        cs=set()
        for c1 in get(a):
            for c2 in get(b):
                u=c1|c2
                if len(u)<=K: cs.add(u)
        cs=sorted(cs, key=lambda c: sorted(c))       # deterministic order (string hashes vary per process)
        # drop dominated cuts (superset of another cut)
        cs=[c for c in cs if not any(o<c for o in cs)]
        cuts[n]=cs+[frozenset([n])]
    return cuts
def mapping_area(nodes,outs,choice):
    # choice: node -> cut ; compute set of LUT roots needed
    need=set(); st=list(outs)
    while st:
        n=st.pop()
        if n in need or n not in nodes: continue
        need.add(n)
        for l in choice[n]: st.append(l)
    return need
def optimize(nodes,order,cuts,outs,iters=20000,seed=0):
    rnd=random.Random(seed)
    nontriv={n:[c for c in cuts[n] if c!=frozenset([n])] for n in order}
    # initial: area-flow
    fan={n:0 for n in order}
    for n in order:
        op,a,b=nodes[n]
        for x in (a,b):
            if x in fan: fan[x]+=1
    for o in outs: fan[o]+=1
    af={}
    choice={}
    for n in order:
        best=None
        for c in nontriv[n]:
            v=1+sum(af.get(l,0) for l in c)
            if best is None or v<best[0]: best=(v,c)
        choice[n]=best[1]; af[n]=best[0]/max(1,fan[n])
    cur=len(mapping_area(nodes,outs,choice))
    best=(cur,dict(choice))
    # local search: change a random node's cut, accept if not worse (plateau moves)
    T=0.3
    import math
    for it in range(iters):
        need=mapping_area(nodes,outs,choice)
        n=rnd.choice(sorted(need))
        c=rnd.choice(nontriv[n])
        if c==choice[n]: continue
        old=choice[n]; choice[n]=c
        a=len(mapping_area(nodes,outs,choice))
        if a<=cur or rnd.random()<math.exp((cur-a)/T):
            cur=a
            if cur<best[0]: best=(cur,dict(choice))
        else: choice[n]=old
        T=max(0.02,T*0.9997)
    return best

def lut_immediates(nodes, order, choice, outs):
    need = mapping_area(nodes, outs, choice)
    def ev(n, env):
        if n in env: return env[n]
        op, a, b = nodes[n]; A = ev(a, env); B = ev(b, env)
        return A ^ B if op == '^' else A & B if op == '&' else 1 ^ A if op == '~' else 1 ^ A ^ B
    luts = []
    for n in order:
        if n not in need: continue
        leaves = sorted(choice[n])
        while len(leaves) < 3: leaves.append(leaves[-1])
        imm = 0
        for idx in range(8):
            env = {}
            for l, bv in zip(leaves, [(idx >> 2) & 1, (idx >> 1) & 1, idx & 1]): env.setdefault(l, bv)
            if ev(n, dict(env)): imm |= 1 << idx
        luts.append((n, leaves, imm))
    return luts

def verify(luts, S=None):
    S = S or sbox_table()
    for x in range(256):
        v = {f"U{i}": (x >> (7 - i)) & 1 for i in range(8)}
        for n, l, imm in luts:
            v[n] = (imm >> ((v[l[0]] << 2) | (v[l[1]] << 1) | v[l[2]])) & 1
        assert sum(v[f"S{i}"] << (7 - i) for i in range(8)) == S[x], x

def emit_function(L, luts, fname, comment):
    name = {f"U{i}": f"x[{7 - i}]" for i in range(8)}
    L.append(comment)
    L.append(f"UAES_HD void {fname}(uint32_t x[8])")
    L.append("{")
    for n, l, imm in luts:
        args = ", ".join(name[z] for z in l)
        v = f"o{n[1:]}" if n.startswith("S") else n
        name[n] = v
        L.append(f"    const uint32_t {v} = lut3<0x{imm:02x}>({args});")
    for i in range(8):
        L.append(f"    x[{7 - i}] = o{i};")
    L.append("}")


# ---------------------------------------------------------------- the inverse S-box, derived
#
# The forward circuit is  S(u) = Bottom(Middle(Top u)) ^ 0x63  with linear Top (8 -> 22 signals) and
# Bottom (18 products -> 8 bits) around the shared non-linear Middle (the GF(2^8) inversion in a tower
# basis).  With the affine map A of FIPS-197 5.1.1 (S = A Inv ^ 0x63):
#     Inv(u)   = A^-1 (S(u) ^ 0x63)           = (A^-1 Bottom) Middle(Top u)
#     S^-1(x)  = Inv(A^-1 (x ^ 0x63))          = (A^-1 Bottom) Middle(Top A^-1 x  ^  Top A^-1 0x63)
# so the inverse S-box (micro_aes.c:53-65, 268-275) re-uses Middle with a new top layer Top A^-1 (plus
# constant complements) and a new bottom layer A^-1 Bottom, both obtained numerically below and
# re-synthesised as XOR networks by greedy common-subexpression elimination (Paar).
def paar(rows, in_names, prefix):
    """rows: list of sets of input names; returns (gates, out_signal_per_row)"""
    rows = [set(r) for r in rows]
    gates, k = [], 0
    while True:
        best, cnt = None, 1
        sigs = sorted(set().union(*rows))
        for i, a in enumerate(sigs):
            for b in sigs[i + 1:]:
                c = sum(1 for r in rows if a in r and b in r)
                if c > cnt: best, cnt = (a, b), c
        if best is None: break
        n = f"{prefix}{k}"; k += 1
        gates.append((n, '^', best[0], best[1]))
        for r in rows:
            if best[0] in r and best[1] in r:
                r -= {best[0], best[1]}; r.add(n)
    outs = []
    for r in rows:                      # what is left shares nothing: chain it
        r = sorted(r)
        acc = r[0]
        for x in r[1:]:
            n = f"{prefix}{k}"; k += 1
            gates.append((n, '^', acc, x)); acc = n
        outs.append(acc)
    return gates, outs


def inverse_network(g):
    first_mid = next(i for i, x in enumerate(g) if x[0] == 't2')
    first_bot = next(i for i, x in enumerate(g) if x[0] == 't46')
    TOP, MID, BOT = g[:first_mid], g[first_mid:first_bot], g[first_bot:]
    mid_sigs = [f"y{i}" for i in range(1, 22)] + ["U7"]
    zs = [f"z{i}" for i in range(18)]

    def run(gates, env):
        v = dict(env)
        for d, op, a, b in gates:
            A, B = v[a], v[b]
            v[d] = A ^ B if op == '^' else A & B if op == '&' else 1 ^ A if op == '~' else 1 ^ A ^ B
        return v
    bits = lambda x: [(x >> (7 - i)) & 1 for i in range(8)]
    frombits = lambda b: sum(v << (7 - i) for i, v in enumerate(b))
    top_of = lambda u: [run(TOP, {f"U{i}": (u >> (7 - i)) & 1 for i in range(8)})[s] for s in mid_sigs]
    bot_of = lambda z: [run(BOT, dict(zip(zs, z)))[f"S{i}"] for i in range(8)]
    unit = lambda n, j: [1 if i == j else 0 for i in range(n)]
    c0 = bot_of([0] * 18)
    assert frombits(c0) == 0x63
    Bm = [frombits([a ^ b for a, b in zip(bot_of(unit(18, j)), c0)]) for j in range(18)]
    rotl = lambda q, s: ((q << s) | (q >> (8 - s))) & 0xff
    Ainv = {q ^ rotl(q, 1) ^ rotl(q, 2) ^ rotl(q, 3) ^ rotl(q, 4): q for q in range(256)}
    newtop = lambda x: top_of(Ainv[x ^ 0x63])
    t0 = newtop(0)
    Tp = [[a ^ b for a, b in zip(newtop(1 << (7 - j)), t0)] for j in range(8)]
    Bp = [bits(Ainv[Bm[j]]) for j in range(18)]
    # XOR networks.  New top: signal k = XOR of inputs X_j with Tp[j][k], complemented when t0[k].
    tg, touts = paar([{f"U{j}" for j in range(8) if Tp[j][k]} for k in range(22)], None, "a")
    gates = list(tg)
    ren = {}
    for k, sname in enumerate(mid_sigs):
        src = touts[k]
        if t0[k]:
            gates.append((f"n{k}", '~', src, src)); src = f"n{k}"
        ren[sname] = src
    for d, op, a, b in MID:                                  # the shared middle, inputs renamed
        gates.append((d, op, ren.get(a, a), ren.get(b, b)))
    bg, bouts = paar([{zs[j] for j in range(18) if Bp[j][i]} for i in range(8)], None, "b")
    gates += bg
    out_name = {bouts[i]: f"S{i}" for i in range(8)}         # the rows' final signals are the outputs
    assert len(out_name) == 8
    return [(out_name.get(d, d), op, out_name.get(a, a), out_name.get(b, b)) for d, op, a, b in gates]


def emit(luts, luts_inv, path):
    L = []
    L.append("// uaes_sbox_lut3.cuh -- GENERATED by tools/gen_sbox_lut3.py, do not edit.")
    L.append("// Bitsliced AES S-box (SubBytes, micro_aes.c:187-194, on 32 blocks at once): x[b] holds bit b")
    L.append(f"// (b = 0 least significant) of one state byte for 32 blocks.  {len(luts)} LOP3 instructions, mapped from")
    L.append("// the Boyar-Peralta 115-gate circuit and verified on all 256 inputs by the generator.")
    L.append("#pragma once")
    L.append("#include <stdint.h>")
    L.append("namespace uaes {")
    L.append("template <int IMM> UAES_HD uint32_t lut3(uint32_t a, uint32_t b, uint32_t c)")
    L.append("{")
    L.append("#ifdef __CUDA_ARCH__")
    L.append("    uint32_t r;")
    L.append('    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(IMM));')
    L.append("    return r;")
    L.append("#else")
    L.append("    uint32_t r = 0;")
    L.append("    for (int i = 0; i < 8; ++i)")
    L.append("        if ((IMM >> i) & 1) r |= ((i & 4) ? a : ~a) & ((i & 2) ? b : ~b) & ((i & 1) ? c : ~c);")
    L.append("    return r;")
    L.append("#endif")
    L.append("}")
    L.append(f"constexpr int kSboxLut3Count = {len(luts)};")
    L.append(f"constexpr int kInvSboxLut3Count = {len(luts_inv)};")
    emit_function(L, luts, "sbox_bitsliced", "// SubBytes, micro_aes.c:187-194")
    emit_function(L, luts_inv, "sbox_inv_bitsliced",
                  "// InvSubBytes, micro_aes.c:268-275: the forward circuit's non-linear middle between a top and a bottom\n"
                  "// linear layer derived by the generator (Top A^-1 and A^-1 Bottom), verified on all 256 inputs")
    L.append("}  // namespace uaes")
    open(path, "w").write("\n".join(L) + "\n")

if __name__ == "__main__":
    g = parse()
    S = sbox_table()
    assert all(evalc(g, x) == S[x] for x in range(256)), "gate network is not the AES S-box"
    nodes, order = build(g); cuts = enum_cuts(nodes, order)
    outs = [f"S{i}" for i in range(8)]
    res = None
    seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    for seed in range(seeds):
        b = optimize(nodes, order, cuts, outs, iters=30000, seed=seed)
        if res is None or b[0] < res[0]: res = b
    luts = lut_immediates(nodes, order, res[1], outs)
    verify(luts)
    # inverse S-box
    gi = inverse_network(g)
    Sinv = [0] * 256
    for i, v in enumerate(S): Sinv[v] = i
    assert all(evalc(gi, x) == Sinv[x] for x in range(256)), "derived network is not the inverse S-box"
    nodes_i, order_i = build(gi); cuts_i = enum_cuts(nodes_i, order_i)
    res_i = None
    for seed in range(seeds):
        b = optimize(nodes_i, order_i, cuts_i, outs, iters=30000, seed=seed)
        if res_i is None or b[0] < res_i[0]: res_i = b
    luts_i = lut_immediates(nodes_i, order_i, res_i[1], outs)
    verify(luts_i, Sinv)
    here = os.path.dirname(os.path.abspath(__file__))
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(here, "..", "micro-aes_b200", "csrc", "uaes_sbox_lut3.cuh")
    emit(luts, luts_i, out)
    print(f"S-box: {len(g)} gates -> {len(luts)} LOP3; inverse S-box: {len(gi)} gates -> {len(luts_i)} LOP3; "
          f"both verified on 256 inputs; wrote {os.path.normpath(out)}")
