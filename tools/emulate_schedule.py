"""Checks (on CPU) the distance-class schedule of csrc/eigen_jacobi.cu: for every nb, the pairs of step (d, par) are
mutually disjoint and all steps together cover every pair (I, J), I < J, exactly once."""
for nb in (2, 4, 8, 14, 34, 68, 126, 252):
    seen = set()
    for d in range(1, nb):
        for par in range(2):
            used = set()
            ncand = ((nb + 2 * d - 1) // (2 * d)) * d
            for c in range(ncand):
                I = (c // d) * 2 * d + par * d + (c % d)
                J = I + d
                if J < nb:
                    assert (I // d) & 1 == par
                    assert I not in used and J not in used, (nb, d, par, I, J)
                    used.update((I, J))
                    assert (I, J) not in seen
                    seen.add((I, J))
    assert len(seen) == nb * (nb - 1) // 2, (nb, len(seen))
print("distance-class schedule ok")
