"""Shape functions of the 4-node and 9-node Lagrange quads and their (r, s) derivatives.

Drop-in for /root/reference/tools/interpFunctions.py (``funH4`` :3-15, ``funH9`` :17-37,
``funHrs4`` :39-53, ``funHrs9`` :55-84): same names, same ``np.matrix`` return shapes
(1 x n for values, 2 x n for derivatives) and -- because the device tables uploaded to
``__constant__`` memory must match the reference bit for bit -- the same floating-point
expression order.  Local node order: 1(+,+) 2(-,+) 3(-,-) 4(+,-) 5(0,+) 6(-,0) 7(0,-)
8(+,0) 9(0,0).

These run once per problem on the host (``FemProblem.setMatrix``); the per-element work
that consumes the tables is the CUDA path in ``csrc/``.  The C twin used by the library
itself is ``fe_tables`` (include/feprogram_b200.h).
"""
import numpy as np

# corner signs (sign of r, sign of s) for local nodes 1..4
_CORNER = ((+1, +1), (-1, +1), (-1, -1), (+1, -1))


def _edge(sign, t):
    """``1 + t`` for sign > 0, ``1 - t`` for sign < 0 (the reference's r_plus / r_minus)."""
    return 1 + t if sign > 0 else 1 - t


def funH4(r, s):
    vals = [0.25 * _edge(sr, r) * _edge(ss, s) for sr, ss in _CORNER]
    return np.matrix(vals)


def funH9(r, s):
    bub_r = 1 - r ** 2
    bub_s = 1 - s ** 2
    h = [0.0] * 9
    h[8] = bub_r * bub_s
    # mid-side nodes 5..8: (0,+) (-,0) (0,-) (+,0)
    h[4] = 0.5 * bub_r * _edge(+1, s) - 0.5 * h[8]
    h[5] = 0.5 * bub_s * _edge(-1, r) - 0.5 * h[8]
    h[6] = 0.5 * bub_r * _edge(-1, s) - 0.5 * h[8]
    h[7] = 0.5 * bub_s * _edge(+1, r) - 0.5 * h[8]
    # corners: bilinear part minus half of the two adjacent mid-sides minus a quarter of the bubble.
    # Adjacent mid-sides in the reference's subtraction order.
    adj = ((4, 7), (4, 5), (5, 6), (6, 7))
    for a, (sr, ss) in enumerate(_CORNER):
        m0, m1 = adj[a]
        h[a] = 0.25 * _edge(sr, r) * _edge(ss, s) - 0.5 * h[m0] - 0.5 * h[m1] - 0.25 * h[8]
    return np.matrix(h)


def _corner_slopes(r, s):
    """Un-scaled bilinear derivative numerators, reference literal forms."""
    dr = [1 + s, -1 - s, -1 + s, 1 - s]
    ds = [1 + r, 1 - r, -1 + r, -1 - r]
    return dr, ds


def funHrs4(r, s):
    dr, ds = _corner_slopes(r, s)
    out = np.matrix(np.zeros((2, 4)))
    out[0] = [0.25 * v for v in dr]
    out[1] = [0.25 * v for v in ds]
    return out


def funHrs9(r, s):
    dr4, ds4 = _corner_slopes(r, s)
    dr_mid = [-2 * r * (1 + s), -1 * (1 - s ** 2), -2 * r * (1 - s), (1 - s ** 2)]
    ds_mid = [(1 - r ** 2), -2 * s * (1 - r), -1 * (1 - r ** 2), -2 * s * (1 + r)]
    dr_c = -2 * r * (1 - s ** 2)
    ds_c = -2 * s * (1 - r ** 2)

    hr = [0] * 9
    hs = [0] * 9
    hr[8], hs[8] = dr_c, ds_c
    for k in range(4):
        hr[4 + k] = 0.5 * dr_mid[k] - 0.5 * dr_c
        hs[4 + k] = 0.5 * ds_mid[k] - 0.5 * ds_c
    # corner a subtracts mid-sides (a+4) and (a+3), corner 0 wraps to (4, 7)
    for a in range(4):
        m0, m1 = (4, 7) if a == 0 else (a + 4, a + 3)
        hr[a] = 0.25 * dr4[a] - 0.5 * hr[m0] - 0.5 * hr[m1] - 0.25 * dr_c
        hs[a] = 0.25 * ds4[a] - 0.5 * hs[m0] - 0.5 * hs[m1] - 0.25 * ds_c

    out = np.matrix(np.zeros((2, 9)))
    out[0] = hr
    out[1] = hs
    return out
