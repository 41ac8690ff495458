"""Generates tests/golden/*.json.  Run here (needs python `cryptography`); the JSON files are committed.

Independent implementations used as the source of truth:
  * AES: python `cryptography` (OpenSSL) in ECB mode -> FIPS-197 block vectors, key-schedule words, and the
    PRF of pianopir/util.go:157-165  PRF(k, tag, x) = LE64((AES_k(B) xor B)[0:8]), B = LE64((tag<<35)+x) || 0^64
  * L2: a scalar numpy float32 emulation of graphann/l2_distance_amd64.s:4-36 (8 lanes, sub/mul/add each rounded,
    hadd tree) + the scalar tail of build_graph.go:119-127
  * uint32 inner product: python integers mod 2^32 (graphann_test.go:225-247)
"""
import json
import os

import numpy as np
from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

HERE = os.path.dirname(os.path.abspath(__file__))
M64 = (1 << 64) - 1


def aes_ecb(key, block):
    return Cipher(algorithms.AES(key), modes.ECB()).encryptor().update(block)


def prf(key, tag, x):
    b = (((tag << 35) + x) & M64).to_bytes(8, "little") + bytes(8)
    e = aes_ecb(key, b)
    return int.from_bytes(bytes(a ^ c for a, c in zip(e, b))[:8], "little")


def key_schedule_words(key):
    """round keys recovered without re-implementing the schedule: AES with all-but-one rounds is not exposed by
    OpenSSL, so use the FIPS-197 recurrence with an S-box derived from the cipher itself:
    S[x] is read off AES-ECB of crafted blocks under the zero key... simpler and still independent of the oracle:
    compute the schedule with python ints and the S-box obtained from GF(2^8) log tables."""
    # GF(2^8) log/antilog tables with generator 3
    exp, log = [0] * 512, [0] * 256
    x = 1
    for i in range(255):
        exp[i] = x
        log[x] = i
        x ^= (x << 1) ^ (0x11B if x & 0x80 else 0)
        x &= 0xFF
    for i in range(255, 512):
        exp[i] = exp[i - 255]
    sbox = []
    for v in range(256):
        inv = 0 if v == 0 else exp[255 - log[v]]
        s = inv
        for r in range(1, 5):
            s ^= ((inv << r) | (inv >> (8 - r))) & 0xFF
        sbox.append(s ^ 0x63)
    w = [list(key[4 * i:4 * i + 4]) for i in range(4)]
    rcon = 1
    for i in range(4, 44):
        t = list(w[i - 1])
        if i % 4 == 0:
            t = [sbox[t[1]] ^ rcon, sbox[t[2]], sbox[t[3]], sbox[t[0]]]
            rcon = ((rcon << 1) ^ (0x11B if rcon & 0x80 else 0)) & 0xFF
        w.append([a ^ b for a, b in zip(w[i - 4], t)])
    words = [int.from_bytes(bytes(x), "little") for x in w]
    # self-check against OpenSSL: encrypt with these round keys == library AES
    return words, sbox


def aes_with_schedule(words, sbox, block):
    def xt(a):
        return ((a << 1) ^ (0x1B if a & 0x80 else 0)) & 0xFF
    k = b"".join(int(w).to_bytes(4, "little") for w in words)
    s = [block[i] ^ k[i] for i in range(16)]
    for r in range(1, 11):
        t = [sbox[s[4 * ((c + row) & 3) + row]] for c in range(4) for row in range(4)]
        if r < 10:
            s = []
            for c in range(4):
                a = t[4 * c:4 * c + 4]
                s += [xt(a[0]) ^ xt(a[1]) ^ a[1] ^ a[2] ^ a[3], a[0] ^ xt(a[1]) ^ xt(a[2]) ^ a[2] ^ a[3],
                      a[0] ^ a[1] ^ xt(a[2]) ^ xt(a[3]) ^ a[3], xt(a[0]) ^ a[0] ^ a[1] ^ a[2] ^ xt(a[3])]
        else:
            s = t
        s = [s[i] ^ k[16 * r + i] for i in range(16)]
    return bytes(s)


def l2_emulated(a, b):
    a, b = a.astype(np.float32), b.astype(np.float32)
    dim = a.size
    body = dim - (dim & 7)
    acc = np.zeros(8, np.float32)
    for i in range(0, body, 8):
        d = (a[i:i + 8] - b[i:i + 8]).astype(np.float32)
        acc = (acc + (d * d).astype(np.float32)).astype(np.float32)
    lo = np.float32(np.float32(acc[0] + acc[1]) + np.float32(acc[2] + acc[3]))
    hi = np.float32(np.float32(acc[4] + acc[5]) + np.float32(acc[6] + acc[7]))
    d = np.float32(lo + hi) if body else np.float32(0)
    for i in range(body, dim):
        x = np.float32(a[i] - b[i])
        d = np.float32(d + np.float32(x * x))
    return d


def main():
    rng = np.random.default_rng(197)
    out = {"fips197": [], "schedule": [], "prf": []}
    # FIPS-197 Appendix C.1 and B
    out["fips197"].append(dict(key=bytes(range(16)).hex(), pt="00112233445566778899aabbccddeeff", ct="69c4e0d86a7b0430d8cdb78070b4c55a"))
    out["fips197"].append(dict(key="2b7e151628aed2a6abf7158809cf4f3c", pt="3243f6a8885a308d313198a2e0370734", ct="3925841d02dc09fbdc118597196a0b32"))
    for v in out["fips197"]:
        assert aes_ecb(bytes.fromhex(v["key"]), bytes.fromhex(v["pt"])).hex() == v["ct"]
    keys = [bytes(range(16)), bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c"), bytes(16), bytes([255] * 16)]
    keys += [bytes(rng.integers(0, 256, 16, dtype=np.uint8)) for _ in range(4)]
    for key in keys:
        words, sbox = key_schedule_words(key)
        blk = bytes(rng.integers(0, 256, 16, dtype=np.uint8))
        assert aes_with_schedule(words, sbox, blk) == aes_ecb(key, blk), "schedule self-check against OpenSSL"
        out["schedule"].append(dict(key=key.hex(), words=words))
    cases = [(0, 0), (1, 0), (0, 1), (5, 7), (104359, 511), (2**29 - 1, 2**35 - 1), (24415, 195), (2**29, 0), (123456789, 2**40 + 17)]
    cases += [(int(rng.integers(0, 2**29)), int(rng.integers(0, 2**20))) for _ in range(40)]
    for key in keys[:3] + keys[4:6]:
        for tag, x in cases:
            out["prf"].append(dict(key=key.hex(), tag=tag, x=x, out=prf(key, tag, x)))
    json.dump(out, open(os.path.join(HERE, "aes_prf_kat.json"), "w"), indent=0)

    l2 = []
    for dim in (8, 16, 128, 192, 100, 13, 5, 960):
        for scale in (1.0, 37.5):
            a = (rng.standard_normal(dim) * scale).astype(np.float32)
            b = (rng.standard_normal(dim)).astype(np.float32)
            l2.append(dict(dim=dim, a=a.view(np.uint32).tolist(), b=b.view(np.uint32).tolist(),
                           out_bits=int(np.float32(l2_emulated(a, b)).view(np.uint32))))
    ip = []
    for dim in (16, 128, 192):
        a = rng.integers(0, 2**32, dim, dtype=np.uint64)
        b = rng.integers(0, 2**32, dim, dtype=np.uint64)
        ip.append(dict(a=a.tolist(), b=b.tolist(), out=int(sum(int(x) * int(y) for x, y in zip(a, b)) % 2**32)))
    json.dump(dict(l2=l2, ip=ip), open(os.path.join(HERE, "distance_kat.json"), "w"), indent=0)
    print("wrote", len(out["prf"]), "PRF KATs,", len(out["schedule"]), "schedules,", len(l2), "L2 and", len(ip), "IP vectors")


if __name__ == "__main__":
    main()
