"""Deterministic key / offset derivation used wherever the reference draws from a time-seeded rng.

The reference seeds `rand.New(rand.NewSource(time.Now().UnixNano()))` for master keys (pianopir/pir.go:132,208)
and for replacement offsets (pir.go:305-306, 345-349); Go's math/rand stream cannot be reproduced without Go.
For testable parity every draw here is a counter-based splitmix64 hash of (seed, counter); DESIGN.md states the
scheme.  Pure integer arithmetic on the host: this is key management, not part of the accelerated path.
"""
MASK = (1 << 64) - 1


def mix64(seed, ctr):
    z = (seed + (ctr + 1) * 0x9E3779B97F4A7C15) & MASK
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
    return z ^ (z >> 31)


def derive_key(key_seed, epoch, parts, i):
    """16-byte PrfKey128 for sub-PIR i at preprocessing epoch `epoch`: two uint64 draws, little-endian,
    the analogue of RandKey128 (pianopir/util.go:25-31)."""
    a = mix64(key_seed, 2 * (epoch * parts + i))
    b = mix64(key_seed, 2 * (epoch * parts + i) + 1)
    return a.to_bytes(8, "little") + b.to_bytes(8, "little")
