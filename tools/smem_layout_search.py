#!/usr/bin/env python
"""Bank-conflict model of the three shared-memory access patterns of ax_hex3d_t_kernel (layouts C, A, B) and a
search over the row stride LD, slab stride SS and element stride ESS for every order.

Model: 32 banks of 4 bytes; a warp-wide access costs max over banks of the number of distinct 32-bit words
requested from that bank (same-word accesses broadcast).  Ideal = ceil(distinct words / 32).
usage: python tools/smem_layout_search.py            (prints the current and the best layout per Nq)
"""
import sys

EPB = {2: 16, 3: 7, 4: 4, 5: 5, 6: 5, 7: 3, 8: 1, 9: 1}


def wavefronts(word_lists):
    """word_lists: per active lane, list of 32-bit word addresses touched by one instruction"""
    banks = {}
    for words in word_lists:
        for w in words:
            banks.setdefault(w % 32, set()).add(w)
    return max((len(v) for v in banks.values()), default=0)


def cost(Nq, LD, SS, ESS, epb=None, perm8=True):
    epb = EPB[Nq] if epb is None else epb
    Nq2 = Nq * Nq
    work = epb * Nq2
    threads = (work + 31) // 32 * 32
    tot = ideal = 0

    def acc(instr_words, weight):
        nonlocal tot, ideal
        lanes = [w for w in instr_words if w is not None]
        if not lanes:
            return
        tot += weight * wavefronts(lanes)
        distinct = len({x for ws in lanes for x in ws})
        ideal += weight * ((distinct + 31) // 32)

    for w0 in range(0, threads, 32):
        lanes = []
        for t in range(w0, w0 + 32):
            if t >= work:
                lanes.append(None)
                continue
            es, ij = divmod(t, Nq2)
            b, a = divmod(ij, Nq)
            jc = (4 * (b & 1) + (b >> 1)) if (Nq == 8 and perm8) else b
            lanes.append((es * ESS + jc * LD + a, es * ESS + b * SS + a * LD, es * ESS + b * SS + a))
        d = lambda off: [2 * off, 2 * off + 1]                      # one double = two words
        # layout C: scalar accesses at slab k (same pattern for every k): ~7 per node
        acc([None if l is None else d(l[0]) for l in lanes], 7 * Nq)
        # layout A: 128-bit row pieces (4 passes over the element), + the odd tail
        for c in range(Nq // 2):
            acc([None if l is None else d(l[1] + 2 * c) + d(l[1] + 2 * c + 1) for l in lanes], 4)
        if Nq & 1:
            acc([None if l is None else d(l[1] + Nq - 1) for l in lanes], 4)
        # layout B: column accesses (4 passes)
        for m in range(Nq):
            acc([None if l is None else d(l[2] + m * LD) for l in lanes], 4)
    return tot, ideal, 3 * epb * 0 + 3 * 8 * ((epb - 1) * ESS + Nq * SS)


def current(Nq):
    """strides compiled into AxT<Nq> (csrc/ax_hex3d.cu); the rule before this search was LD = 2/6/10 with LD/2 odd and
    SS = 8 mod 16 for every order"""
    LD = {2: 2, 3: 4, 4: 4, 5: 6, 6: 6, 7: 10, 8: 10, 9: 10}[Nq]
    SS = {2: 4, 3: 12, 4: 18, 5: 30, 6: 38, 7: 70, 8: 88, 9: 90}[Nq]
    ESS = {2: 10, 3: 42, 4: 72, 5: 150, 6: 228, 7: 496}.get(Nq, Nq * SS)
    return LD, SS, ESS


def main():
    for Nq in range(2, 10):
        LD0, SS0, ESS0 = current(Nq)
        c0 = cost(Nq, LD0, SS0, ESS0)
        best = None
        ldmin = Nq + (Nq & 1)
        for LD in range(ldmin, ldmin + 9, 2):
            for SS in range(Nq * LD, Nq * LD + 33, 2):
                for pad in (range(0, 33, 2) if EPB[Nq] > 1 else [0]):
                    ESS = Nq * SS + pad
                    c = cost(Nq, LD, SS, ESS)
                    if c[2] > 48 * 1024:
                        continue
                    key = (c[0], c[2])
                    if best is None or key < best[0]:
                        best = (key, LD, SS, ESS, c)
        _, LD, SS, ESS, c = best
        print(f"Nq={Nq} EPB={EPB[Nq]}: current LD={LD0} SS={SS0} ESS={ESS0} wavefronts={c0[0]} (ideal {c0[1]}, x{c0[0] / c0[1]:.2f}) "
              f"smem={c0[2]}  |  best LD={LD} SS={SS} ESS={ESS} wavefronts={c[0]} (x{c[0] / c[1]:.2f}) smem={c[2]}")


if __name__ == "__main__":
    sys.exit(main())
