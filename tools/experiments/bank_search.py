"""Offline search for a slab layout (row stride, block-slot permutation) that minimises shared-memory bank
conflicts of phase 2a on the structured Quad4 tile (8 x 16 elements, row-major elements inside the tile)."""
import itertools, sys
import numpy as np

N = 4
def tri(lo, hi): return lo * N - lo * (lo - 1) // 2 + (hi - lo)

PW, PH = 16, 8           # patch width/height in elements
W = 4097                 # mesh row pitch in nodes (only for ordering)
# element rows in slab: own elements row-major; halo: top row (PW+1 elements incl. corner), right column
def erow(ex, ey):
    if ex < PW and ey < PH: return ey * PW + ex
    if ey == PH: return PW * PH + ex            # top halo row, ex in 0..PW
    return PW * PH + PW + 1 + ey                # right halo column (ex == PW), ey in 0..PH-1
# owned nodes: (ix,iy) ix in 1..PW, iy in 1..PH ; local numbering of element (ex,ey): 0=(ex,ey) 1=(ex+1,ey) 2=(ex+1,ey+1) 3=(ex,ey+1)
def local_index(ex, ey, ix, iy):
    return {(0, 0): 0, (1, 0): 1, (1, 1): 2, (0, 1): 3}[(ix - ex, iy - ey)]
items = []   # per work item: list of sources (row, lo, hi, swap) ; self -> None
for iy in range(1, PH + 1):
    for ix in range(1, PW + 1):
        nbrs = [(ix + dx, iy + dy) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]   # sorted by global id
        for (mx, my) in nbrs:
            if (mx, my) == (ix, iy):
                items.append(None); continue
            srcs = []
            for ey in (iy - 1, iy):
                for ex in (ix - 1, ix):          # ascending element id: row-major
                    if ex <= mx <= ex + 1 and ey <= my <= ey + 1:
                        a = local_index(ex, ey, ix, iy); b = local_index(ex, ey, mx, my)
                        srcs.append((erow(ex, ey), min(a, b), max(a, b), a > b))
            items.append(srcs)
print('work items', len(items), 'per node', len(items) / (PW * PH))

def wavefronts(stride, perm, verbose=False):
    tot = 0
    for w0 in range(0, len(items), 32):
        chunk = items[w0:w0 + 32]
        for half in (chunk[:16], chunk[16:]):
            for si in (0, 1):
                for k in (0, 1, 2):
                    addrs = set()
                    for it in half:
                        if it is None: continue
                        if si < len(it):
                            row, lo, hi, sw = it[si]
                            kk = k if k == 0 else ((1 + sw) if k == 1 else (2 - sw))
                            addrs.add(row * stride + 3 * perm[tri(lo, hi)] + kk)
                        else:
                            addrs.add(-1000 + k)   # zero block
                    if not addrs: continue
                    banks = {}
                    for a in addrs: banks[a % 16] = banks.get(a % 16, 0) + 1
                    tot += max(banks.values())
    return tot

base = wavefronts(31, list(range(10)))
ideal = sum(1 for w0 in range(0, len(items), 32) for _ in range(12))
print('current layout (stride 31, identity):', base, 'conflict-free would be ~', ideal)
rng = np.random.default_rng(0)
best = (base, 31, list(range(10)))
for stride in (31, 33, 35, 37, 39, 41, 43, 45):
    for trial in range(400):
        perm = list(rng.permutation(10)) if trial else list(range(10))
        wv = wavefronts(stride, perm)
        if wv < best[0]:
            best = (wv, stride, perm); print('better', best, flush=True)
print('best', best)
