"""Bank-conflict check of the tile-major log-likelihood layout proposed for the next round (DESIGN.md
section 8, item 1a): a tile is [K rows][32 nodes] doubles at row stride 32, node n of state row k stored at
column n ^ (4 * (k & 7)).  Checks, by enumeration of the shared-memory banks (32 banks x 4 B; a 64-bit
access is served per half-warp):
  * the DMMA A-operand load of estep_bulk.cuh (lane g = lane/4, t = lane%4 reads state 8*kt+g, node 4*ns+t),
  * the producers' column access (lane = node, one state row at a time),
  * that four consecutive nodes stay contiguous (the emission kernel's 32-byte stores),
are conflict-free / contiguous for every state tile and k-step.  Pure CPU; prints OK or the first conflict."""
import itertools


def col(k, n):
    return n ^ (4 * (k & 7))


def half_warp_conflicts(addresses_doubles):
    """addresses of the 16 lanes of a half-warp in doubles -> True if two distinct addresses share a bank pair"""
    seen = {}
    for a in addresses_doubles:
        b = a % 16
        if b in seen and seen[b] != a:
            return True
        seen[b] = a
    return False


def main():
    for kt, ns in itertools.product(range(5), range(8)):
        for half in (0, 1):
            lanes = range(16 * half, 16 * half + 16)
            addr = [(8 * kt + (l >> 2)) * 32 + col(8 * kt + (l >> 2), 4 * ns + (l & 3)) for l in lanes]
            assert not half_warp_conflicts(addr), ("A-operand", kt, ns, half)
    for k in range(40):
        for half in (0, 1):
            addr = [k * 32 + col(k, n) for n in range(16 * half, 16 * half + 16)]
            assert not half_warp_conflicts(addr), ("column", k, half)
        assert sorted(col(k, n) for n in range(32)) == list(range(32))
        for n0 in range(0, 32, 4):
            c = [col(k, n0 + u) for u in range(4)]
            assert c == list(range(c[0], c[0] + 4)) and c[0] % 4 == 0, ("contiguity", k, n0)
    print("OK: stride-32 rows with column n ^ 4*(k & 7) are conflict-free for the DMMA operand loads and the "
          "per-node column accesses, and keep 4-node groups contiguous")


if __name__ == "__main__":
    main()
