"""The credit arithmetic of tick_graph_kernel (fse_tick_rows.cuh), restated in Python and run as an event simulation: every chunk
visit must become ready exactly once, only after all 8 neighbours (and therefore, transitively, the chunk itself) have finished
their visits of earlier phases, and every visit of the tick must run.  Formulas mirrored from the kernel: colour of a chunk,
the phase a finishing visit credits at each neighbour, and the credit count `n0 + n * nnb` that makes a visit ready."""
import random

import pytest


def colour(ci, cj):  # graph_colour(): phase index of a chunk's colour, world.cpp:1059-1060 order (0,1),(1,1),(0,0),(1,0)
    return (ci & 1) + 2 * (1 - (cj & 1))


def neighbours(ci, cj, nx, ny):
    for q in range(9):
        if q == 4:
            continue
        ni, nj = ci + q % 3 - 1, cj + q // 3 - 1
        if 0 <= ni < nx and 0 <= nj < ny:
            yield ni, nj


@pytest.mark.parametrize("nx,ny,iters", [(2, 2, 1), (3, 2, 3), (5, 7, 3), (8, 8, 4), (62, 6, 2)])
def test_every_visit_becomes_ready_once_and_in_dependency_order(nx, ny, iters):
    n_phases = 4 * iters
    credits = {(i, j): 0 for i in range(nx) for j in range(ny)}
    done = {(i, j): 0 for i in range(nx) for j in range(ny)}  # visits finished per chunk
    ready = [((i, j), 0) for j in range(ny) for i in range(nx) if colour(i, j) == 0]  # graph_init_kernel: phase 0 has no dependencies
    seen = set()
    rng = random.Random(nx * 100 + ny)
    finished = 0
    while ready:
        (ci, cj), p = ready.pop(rng.randrange(len(ready)))  # any order the CTAs might pop and finish in
        assert ((ci, cj), p) not in seen
        seen.add(((ci, cj), p))
        tk = p & 3
        assert colour(ci, cj) == tk and done[(ci, cj)] == p >> 2  # its own earlier visits are over
        for ni, nj in neighbours(ci, cj, nx, ny):  # what the kernel's former spin-wait checked: need = visits of the neighbour before p
            tkn = colour(ni, nj)
            need = 0 if p <= tkn else (p - tkn + 3) >> 2
            assert done[(ni, nj)] == need, ((ci, cj), p, (ni, nj), done[(ni, nj)], need)
        done[(ci, cj)] += 1
        finished += 1
        for ni, nj in neighbours(ci, cj, nx, ny):  # credit the next visit of each neighbour
            tkc = colour(ni, nj)
            pc = p + ((tkc - tk + 4) & 3)
            assert pc > p and (pc & 3) == tkc
            if pc >= n_phases:
                continue
            nb = list(neighbours(ni, nj, nx, ny))
            n0 = sum(1 for e in nb if colour(*e) < tkc)
            target = n0 + (pc >> 2) * len(nb)
            credits[(ni, nj)] += 1
            if credits[(ni, nj)] == target:
                ready.append(((ni, nj), pc))
    assert finished == nx * ny * iters == len(seen)
    assert all(v == iters for v in done.values())
