"""Host emulation of map_rays_brick_kernel's walk (csrc/map.cu: ray_setup_brick / ray_jump / the step loop), checked
against the plain step-by-step walk of the oracle (oracle/icp_oracle.c: orc_map_integrate_rays) on random rays, random
occupancy and random z-slabs.  Pure integers; run on the CPU before spending GPU time on the kernel.

  python tools/checks/brick_walk_emulation.py [n_cases]
"""
import random
import sys

import numpy as np

B = 8     # fine bricks (kBrick)
B2 = 32   # coarse bricks (kBrick2)


def plain_visits(o, e, z_lo, z_hi):
    nx, ny, nz = (abs(e[k] - o[k]) for k in range(3))
    sx, sy, sz = ((e[k] > o[k]) - (e[k] < o[k]) for k in range(3))
    mx, my, mz = max(nx, 1), max(ny, 1), max(nz, 1)
    INF = 1 << 62
    ex, ey, ez = (my * mz if nx else INF), (mx * mz if ny else INF), (mx * my if nz else INF)
    dx, dy, dz = 2 * my * mz, 2 * mx * mz, 2 * mx * my
    x, y, z = o
    cx = cy = cz = 0
    out = []
    for _ in range(nx + ny + nz - 1):
        if ex <= ey and ex <= ez:
            x += sx; cx += 1; ex = ex + dx if cx < nx else INF
        elif ey <= ez:
            y += sy; cy += 1; ey = ey + dy if cy < ny else INF
        else:
            z += sz; cz += 1; ez = ez + dz if cz < nz else INF
        if z_lo <= z < z_hi:
            out.append((x, y, z))
    return out


def brick_visits(o, e, z_lo, z_hi, occupied):
    """Voxels the brick walk READS (those in occupied bricks), in order."""
    zs = z_hi - z_lo
    ox, oy, oz = o
    nx, ny, nz = (abs(e[k] - o[k]) for k in range(3))
    sx, sy, sz = ((e[k] > o[k]) - (e[k] < o[k]) for k in range(3))
    mx, my, mz = max(nx, 1), max(ny, 1), max(nz, 1)
    P3 = 3 * mx * my * mz
    ex, ey, ez = (my * mz if nx else P3), (mx * mz if ny else P3), (mx * my if nz else P3)
    dx, dy, dz = 2 * my * mz, 2 * mx * mz, 2 * mx * my
    steps = nx + ny + nz
    cx = cy = cz = 0
    live, entered = True, False
    if oz < z_lo or oz >= z_hi:
        k = 0
        if sz > 0 and oz < z_lo:
            k = z_lo - oz
        elif sz < 0 and oz >= z_hi:
            k = oz - (z_hi - 1)
        if k <= 0 or k > nz:
            live = False
        else:
            cz = k
            cx = min(nx, ((2 * k - 1) * nx + nz) // (2 * nz)) if nx else 0
            cy = min(ny, ((2 * k - 1) * ny + nz) // (2 * nz)) if ny else 0
            if nx: ex = (2 * cx + 1) * my * mz
            if ny: ey = (2 * cy + 1) * mx * mz
            ez = (2 * cz + 1) * mx * my
            entered = True
    x, y, zr = ox + sx * cx, oy + sy * cy, oz + sz * cz - z_lo
    rx, ry, rz = nx - cx, ny - cy, nz - cz
    done = cx + cy + cz
    rem = max(steps - 1 - done, 0) if live else 0
    out = []
    if not live:
        return out
    empty = empty2 = False

    def walls_before(e_, d_, r_, T, first):
        # the kernel's float estimate of the quotient settled by the exact remainder
        lim = T - e_ - (0 if first else 1)
        if lim < 0:
            return 0
        c = int(np.float32(lim) * (np.float32(1.0) / np.float32(d_)))
        rmd = lim - c * d_
        c += (1 if rmd >= d_ else 0) - (1 if rmd < 0 else 0)
        assert c == lim // d_, (lim, d_, c)
        return min(c + 1, r_)

    occupied2 = {(i * B // B2, j * B // B2, k * B // B2) for (i, j, k) in occupied}

    def lookup(coarse_too):
        nonlocal empty, empty2
        if coarse_too:
            empty2 = (x // B2, y // B2, zr // B2) not in occupied2
        empty = empty2 or (x // B, y // B, zr // B) not in occupied

    lookup(True)
    if entered and done <= steps - 1 and not empty:
        out.append((x, y, zr + z_lo))
    BIG = 1 << 62
    while rem > 0:
        if empty:
            BB = B2 if empty2 else B
            lx, ly, lz = x % BB, y % BB, zr % BB
            kx = BB - lx if sx > 0 else lx + 1
            ky = BB - ly if sy > 0 else ly + 1
            kz = min(BB - lz, zs - zr) if sz > 0 else lz + 1
            vx, vy, vz = kx <= rx, ky <= ry, kz <= rz
            Tx = ex + (kx - 1) * dx if vx else BIG
            Ty = ey + (ky - 1) * dy if vy else BIG
            Tz = ez + (kz - 1) * dz if vz else BIG
            if not (vx or vy or vz):
                rem = 0
                break
            bx = Tx <= Ty and Tx <= Tz
            by = (not bx) and Ty <= Tz
            T = Tx if bx else (Ty if by else Tz)
            jx = kx - 1 if bx else walls_before(ex, dx, rx, T, True)
            jy = ky - 1 if by else walls_before(ey, dy, ry, T, not bx)
            jz = kz - 1 if (not bx and not by) else walls_before(ez, dz, rz, T, False)
            skip = jx + jy + jz
            if skip >= rem:
                rem = 0
                break
            rem -= skip
            x += sx * jx; y += sy * jy; zr += sz * jz
            rx -= jx; ry -= jy; rz -= jz
            ex += jx * dx; ey += jy * dy; ez += jz * dz
        # one ordinary step
        px = ex <= ey and ex <= ez
        py = (not px) and ey <= ez
        pz = not px and not py
        if px: x += sx; rx -= 1; ex += dx
        elif py: y += sy; ry -= 1; ey += dy
        else: zr += sz; rz -= 1; ez += dz
        rem -= 1
        if zr < 0 or zr >= zs:
            rem = 0
            break
        c, s = (x, sx) if px else ((y, sy) if py else (zr, sz))
        if (c % B) == (0 if s > 0 else B - 1):
            lookup((c % B2) == (0 if s > 0 else B2 - 1))
        if not empty:
            out.append((x, y, zr + z_lo))
    return out


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    rng = random.Random(1234)
    bad = 0
    for case in range(n_cases):
        dims = (rng.randint(1, 100), rng.randint(1, 100), rng.randint(1, 100))
        z_lo = rng.randint(0, dims[2] - 1)
        z_hi = rng.randint(z_lo + 1, dims[2])
        if rng.random() < 0.5:
            z_lo, z_hi = 0, dims[2]
        o = tuple(rng.randint(0, d - 1) for d in dims)
        e = tuple(rng.randint(0, d - 1) for d in dims)
        if rng.random() < 0.2:   # axis-aligned / diagonal rays: many ties
            k = rng.randint(0, 2)
            e = tuple(o[j] if j != k else e[j] for j in range(3))
        if rng.random() < 0.2:
            dlt = rng.randint(-min(dims), min(dims))
            e = tuple(min(max(o[j] + dlt * rng.choice((-1, 1)), 0), dims[j] - 1) for j in range(3))
        nb = [(d + B - 1) // B for d in (dims[0], dims[1], z_hi - z_lo)]
        p_occ = rng.choice((0.0, 0.01, 0.1, 0.5, 1.0))
        occupied = {(i, j, k) for i in range(nb[0]) for j in range(nb[1]) for k in range(nb[2]) if rng.random() < p_occ}
        want = [v for v in plain_visits(o, e, z_lo, z_hi) if (v[0] // B, v[1] // B, (v[2] - z_lo) // B) in occupied]
        got = brick_visits(o, e, z_lo, z_hi, occupied)
        if want != got:
            bad += 1
            if bad < 5:
                print("MISMATCH", dims, (z_lo, z_hi), o, e, sorted(occupied)[:6], want[:8], got[:8])
    print(f"{n_cases} cases, {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
