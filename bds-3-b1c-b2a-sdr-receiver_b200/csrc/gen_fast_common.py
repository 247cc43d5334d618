"""Shared part of the chip-body generators (gen_fast_wb.py, gen_fast_b2a.py).

A thread integrates one *unit* of the code (one B1C chip = ten B2a chips = 97.14 samples at 99.375 MHz) that splits into
`nseg` segments on which every replica is constant.  Boundary k (k = 1..nseg) falls at sample R[k] or one later,
decided per thread by a jitter bit (FAST*_SELU).  `emit_body` writes the straight-line per-word code: each re-aligned
32-bit word of four int8 samples is AND-masked to the bytes of a segment and fed to IDP.2A against the Q15 carrier
table, into the accumulator `acc_name(k)` of that segment's class.
"""


def emit_body(R, nseg, nwords, acc_name, prefix="FAST"):
    """-> list of C statements (the `t_k` mask declarations first, then one block per word)"""
    lines = []
    e = lines.append
    tdone = set()

    def bytemask(samples, i):
        m = 0
        for s in samples:
            if s >> 2 == i:
                m |= 0xFF << (8 * (s & 3))
        return m

    def need_t(k):
        if k not in tdone:
            jb = 0xFF << (8 * (R[k] & 3))
            e("const unsigned t%d = %s_SELU(%d, 0x%08xu);" % (k, prefix, k, jb))
            tdone.add(k)

    for i in range(nwords):
        body = []
        for k in range(nseg):
            core = list(range(R[k] + 1, R[k + 1]))
            if k == 0:
                core = [0] + core                      # sample 0 always belongs to segment 0
            sc = bytemask(core, i)
            sj = (k >= 1 and R[k] >> 2 == i)            # start jitter sample in this word
            ej = (R[k + 1] >> 2 == i)                   # end jitter sample in this word
            jbs = (0xFF << (8 * (R[k] & 3))) if sj else 0
            jbe = (0xFF << (8 * (R[k + 1] & 3))) if ej else 0
            pot = sc | jbs | jbe
            if pot == 0:
                continue
            if sj:
                need_t(k)
            if ej:
                need_t(k + 1)
            if sj and ej:
                expr = "X & ((0x%08xu ^ t%d) | t%d)" % (sc | jbs, k, k + 1)
            elif sj:
                expr = "X & (0x%08xu ^ t%d)" % (sc | jbs, k)
            elif ej:
                expr = "X & (0x%08xu | t%d)" % (sc, k + 1)
            elif sc == 0xFFFFFFFF:
                expr = "X"
            else:
                expr = "X & 0x%08xu" % sc
            n = acc_name(k)
            s = "{ const unsigned M = %s; " % expr
            if pot & 0x0000FFFF:
                s += "%sr = %s_DP_LO(T.x, M, %sr); %si = %s_DP_LO(T.z, M, %si); " % (n, prefix, n, n, prefix, n)
            if pot & 0xFFFF0000:
                s += "%sr = %s_DP_HI(T.y, M, %sr); %si = %s_DP_HI(T.w, M, %si); " % (n, prefix, n, n, prefix, n)
            s += "}"
            body.append(s)
        e("{ const unsigned X = %s_FSH(%s_RAW(%d), %s_RAW(%d)); const int4 T = %s_WTAB(%d);" % (prefix, prefix, i, prefix, i + 1, prefix, i))
        for b in body:
            e(b)
        e("}")
    # the t_k declarations must precede the word blocks that use them: hoist them to the front
    tdecl = [ln for ln in lines if ln.startswith("const unsigned t")]
    rest = [ln for ln in lines if not ln.startswith("const unsigned t")]
    return tdecl + rest
