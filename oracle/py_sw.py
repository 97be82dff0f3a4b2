"""Second, independent restatement of rules P1-P5 in plain Python.

TEST INFRASTRUCTURE ONLY.  Deliberately structured differently from fade_oracle.c: full H/E/F
matrices filled row-major, and a traceback that re-derives every decision from the matrices
instead of reading stored trace bits.  tests/ cross-check the two on random inputs.
Reference call site: source/analysis.d:67 (p.sw_striped(q_seq, ref_seq)).
"""
from __future__ import annotations

_IDX = {"A": 0, "C": 1, "T": 2, "G": 3, "N": 4}
NEG = -10 ** 9


def _code(ch: str) -> int:
    return _IDX.get(ch.upper(), 5)


def sw_trace_py(q: str, t: str, gap_open=10, gap_extend=2, match=2, mismatch=-3):
    """Returns dict(score, end_query, end_ref, beg_query, beg_ref, cigar) -- cigar with S padding."""
    n, m = len(q), len(t)
    qc = [_code(c) for c in q]
    tc = [_code(c) for c in t]

    def sub(a, b):
        if a == 5 or b == 5:
            return 0
        return match if a == b else mismatch

    H = [[0] * (m + 1) for _ in range(n + 1)]   # 1-based with zero borders
    E = [[NEG] * (m + 1) for _ in range(n + 1)]  # horizontal gap (consumes target, 'D')
    F = [[NEG] * (m + 1) for _ in range(n + 1)]  # vertical gap (consumes query, 'I')
    for i in range(1, n + 1):
        Hi, Him, Ei, Fi, Fim = H[i], H[i - 1], E[i], F[i], F[i - 1]
        for j in range(1, m + 1):
            Ei[j] = max(Hi[j - 1] - gap_open, Ei[j - 1] - gap_extend)
            Fi[j] = max(Him[j] - gap_open, Fim[j] - gap_extend)
            Hi[j] = max(0, Him[j - 1] + sub(qc[i - 1], tc[j - 1]), Ei[j], Fi[j])
    score = max(max(row) for row in H)
    if score <= 0:
        return dict(score=0, end_query=0, end_ref=0, beg_query=0, beg_ref=0, cigar="")
    # P3: first column containing the score, then first row in that column
    end_ref = min(j for j in range(1, m + 1) if any(H[i][j] == score for i in range(1, n + 1)))
    end_query = min(i for i in range(1, n + 1) if H[i][end_ref] == score)
    i, j, state = end_query, end_ref, "H"
    ops = []
    while i >= 1 and j >= 1:
        if state == "H":
            h = H[i][j]
            hd = max(0, H[i - 1][j - 1] + sub(qc[i - 1], tc[j - 1]))
            if h == hd:
                if h == 0:
                    break
                ops.append("=" if q[i - 1].upper() == t[j - 1].upper() else "X")
                i -= 1
                j -= 1
            elif h == F[i][j]:
                state = "F"
            else:
                state = "E"
        elif state == "F":
            ops.append("I")
            opened = H[i - 1][j] - gap_open > F[i - 1][j] - gap_extend
            i -= 1
            if opened:
                state = "H"
        else:
            ops.append("D")
            opened = H[i][j - 1] - gap_open > E[i][j - 1] - gap_extend
            j -= 1
            if opened:
                state = "H"
    beg_query, beg_ref = i, j  # 0-based index of the first aligned cell == 1-based i, j after loop
    ops.reverse()
    cig = []
    if beg_query > 0:
        cig.append([beg_query, "S"])
    for op in ops:
        if cig and cig[-1][1] == op and not (cig[-1][1] == "S"):
            cig[-1][0] += 1
        else:
            cig.append([1, op])
    if n - end_query > 0:
        cig.append([n - end_query, "S"])
    return dict(score=score, end_query=end_query - 1, end_ref=end_ref - 1, beg_query=beg_query,
                beg_ref=beg_ref, cigar="".join(f"{l}{o}" for l, o in cig))
