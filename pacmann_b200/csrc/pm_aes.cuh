// AES-128 on the SM integer pipe + shared-memory T-table, bit-exact with FIPS-197 and therefore with
// the reference's AES-NI routines (pianopir/aes_amd64.s:19-126, a copy of Go's crypto/aes).
//
// Layout choices (B200):
//  * one 1 KB encryption table Te0 (LE column packing: byte0 = 2S, byte1 = S, byte2 = S, byte3 = 3S);
//    Te1..Te3 are byte rotations of it (PRMT, ALU pipe) -- or, with NTAB = 4, four tables.
//  * every table is replicated 32x in shared memory, word (v, lane) at [v*32 + lane], so lane l only
//    ever touches bank l: data-dependent lookups are bank-conflict free by construction.
//  * round keys are read from kernel parameters (constant bank, uniform datapath), never from
//    shared memory or registers.
#pragma once
#include <stdint.h>

namespace pm {

__device__ __forceinline__ uint32_t byte_of(uint32_t w, int k) {  // zero-extended byte k of w
    return __byte_perm(w, 0, 0x4440 + k);
}
__device__ __forceinline__ uint32_t rotl8(uint32_t v) { return __byte_perm(v, v, 0x2103); }
__device__ __forceinline__ uint32_t rotl16(uint32_t v) { return __byte_perm(v, v, 0x1032); }
__device__ __forceinline__ uint32_t rotl24(uint32_t v) { return __byte_perm(v, v, 0x0321); }

// Lane-private view of the replicated tables.  `base` points at shared word [0*REP + lane % REP].
// NTAB = 1: Te0 replicated 32x (conflict-free);  NTAB = 4: Te0..Te3 replicated 32x;  NTAB = 8: "compact", Te0 replicated
// 4x only (4 KB instead of 32 KB: for kernels that evaluate few PRFs and need the shared memory for occupancy).
template <int NTAB>
struct AesTab {
    static constexpr int REP = NTAB == 8 ? 4 : 32;
    const uint32_t *base;
    __device__ __forceinline__ uint32_t t0(uint32_t b) const { return base[b * REP]; }
    __device__ __forceinline__ uint32_t t1(uint32_t b) const {
        return NTAB == 4 ? base[8192 + b * 32] : rotl8(base[b * REP]);
    }
    __device__ __forceinline__ uint32_t t2(uint32_t b) const {
        return NTAB == 4 ? base[2 * 8192 + b * 32] : rotl16(base[b * REP]);
    }
    __device__ __forceinline__ uint32_t t3(uint32_t b) const {
        return NTAB == 4 ? base[3 * 8192 + b * 32] : rotl24(base[b * REP]);
    }
    __device__ __forceinline__ uint32_t sbox(uint32_t b) const { return byte_of(base[b * REP], 1); }
};

// shared words needed by AesTab<NTAB>
template <int NTAB>
__host__ __device__ constexpr int aes_tab_words() { return NTAB == 8 ? 256 * 4 : NTAB * 256 * 32; }

// Fill the replicated tables from the 256-entry Te0 in constant memory.  Whole CTA participates.
template <int NTAB>
__device__ __forceinline__ void aes_tab_fill(uint32_t *smem, const uint32_t *te0_const) {
    constexpr int REP = AesTab<NTAB>::REP;
    for (int i = threadIdx.x; i < 256 * REP; i += blockDim.x) {
        uint32_t v = te0_const[i / REP];  // warp-uniform index when REP = 32
        smem[i] = v;
        if (NTAB == 4) {
            smem[8192 + i] = rotl8(v);
            smem[2 * 8192 + i] = rotl16(v);
            smem[3 * 8192 + i] = rotl24(v);
        }
    }
}

// One full round: SubBytes+ShiftRows+MixColumns via tables, then AddRoundKey(k0..k3).
template <int NTAB>
__device__ __forceinline__ void aes_round(const AesTab<NTAB> &T, uint32_t &s0, uint32_t &s1, uint32_t &s2, uint32_t &s3,
                                          uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3) {
    uint32_t t0 = T.t0(byte_of(s0, 0)) ^ T.t1(byte_of(s1, 1)) ^ T.t2(byte_of(s2, 2)) ^ T.t3(byte_of(s3, 3)) ^ k0;
    uint32_t t1 = T.t0(byte_of(s1, 0)) ^ T.t1(byte_of(s2, 1)) ^ T.t2(byte_of(s3, 2)) ^ T.t3(byte_of(s0, 3)) ^ k1;
    uint32_t t2 = T.t0(byte_of(s2, 0)) ^ T.t1(byte_of(s3, 1)) ^ T.t2(byte_of(s0, 2)) ^ T.t3(byte_of(s1, 3)) ^ k2;
    uint32_t t3 = T.t0(byte_of(s3, 0)) ^ T.t1(byte_of(s0, 1)) ^ T.t2(byte_of(s1, 2)) ^ T.t3(byte_of(s2, 3)) ^ k3;
    s0 = t0; s1 = t1; s2 = t2; s3 = t3;
}

// Full AES-128 block encryption of the LE words (s0..s3); rk = 44 words.  Used by the generic PRF
// kernel (pm_prf_batch), where no structure of the input can be assumed.
template <int NTAB, typename RK>
__device__ __forceinline__ void aes128_encrypt(const AesTab<NTAB> &T, const RK &rk, uint32_t &s0, uint32_t &s1,
                                               uint32_t &s2, uint32_t &s3) {
    s0 ^= rk[0]; s1 ^= rk[1]; s2 ^= rk[2]; s3 ^= rk[3];
#pragma unroll
    for (int r = 1; r < 10; r++) aes_round(T, s0, s1, s2, s3, rk[4 * r], rk[4 * r + 1], rk[4 * r + 2], rk[4 * r + 3]);
    uint32_t t0 = T.sbox(byte_of(s0, 0)) | (T.sbox(byte_of(s1, 1)) << 8) | (T.sbox(byte_of(s2, 2)) << 16) |
                  (T.sbox(byte_of(s3, 3)) << 24);
    uint32_t t1 = T.sbox(byte_of(s1, 0)) | (T.sbox(byte_of(s2, 1)) << 8) | (T.sbox(byte_of(s3, 2)) << 16) |
                  (T.sbox(byte_of(s0, 3)) << 24);
    uint32_t t2 = T.sbox(byte_of(s2, 0)) | (T.sbox(byte_of(s3, 1)) << 8) | (T.sbox(byte_of(s0, 2)) << 16) |
                  (T.sbox(byte_of(s1, 3)) << 24);
    uint32_t t3 = T.sbox(byte_of(s3, 0)) | (T.sbox(byte_of(s0, 1)) << 8) | (T.sbox(byte_of(s1, 2)) << 16) |
                  (T.sbox(byte_of(s2, 3)) << 24);
    s0 = t0 ^ rk[40]; s1 = t1 ^ rk[41]; s2 = t2 ^ rk[42]; s3 = t3 ^ rk[43];
}

// ---- the hint-generation PRF -------------------------------------------------------------------
// PRFEvalWithLongKeyAndTag (pianopir/util.go:157-165) evaluates AES-MMO on B = LE64((tag<<35)+x)||0^64
// and the caller keeps only `& (ChunkSize-1)` (pir.go:318,336).  With x = chunk id < 2^32:
//   word0 = x, word1 = low32(tag << 3), word2 = word3 = 0
// so (a) after round 1 the state is  X(x) xor G(tag)  with X depending on word0 only and G on the tag
// and key only: G is hoisted per hint (prf_tag_part), X costs 4 lookups per evaluation;
// (b) only the low NB bytes of output word 0 are kept, so round 9 needs NB of its 4 output columns
// and the final round NB S-box lookups.  Lookups per evaluation: 4 + 7*16 + 4*NB + NB (126 for NB=2)
// instead of 160.  The result is identical to masking the full AES-MMO output.
struct PrfTagPart { uint32_t g0, g1, g2, g3; };

template <int NTAB, typename RK>
__device__ __forceinline__ PrfTagPart prf_tag_part(const AesTab<NTAB> &T, const RK &rk, uint64_t tag) {
    uint32_t w1 = (uint32_t)(tag << 3) ^ rk[1], w2 = rk[2], w3 = rk[3];
    PrfTagPart g;
    g.g0 = T.t1(byte_of(w1, 1)) ^ T.t2(byte_of(w2, 2)) ^ T.t3(byte_of(w3, 3)) ^ rk[4];
    g.g1 = T.t0(byte_of(w1, 0)) ^ T.t1(byte_of(w2, 1)) ^ T.t2(byte_of(w3, 2)) ^ rk[5];
    g.g2 = T.t3(byte_of(w1, 3)) ^ T.t0(byte_of(w2, 0)) ^ T.t1(byte_of(w3, 1)) ^ rk[6];
    g.g3 = T.t2(byte_of(w1, 2)) ^ T.t0(byte_of(w3, 0)) ^ T.t3(byte_of(w2, 3)) ^ rk[7];
    return g;
}

// low 8*NB bits of PRF(rk, tag, x); NB in {2, 4}
template <int NTAB, int NB, typename RK>
__device__ __forceinline__ uint32_t prf_low(const AesTab<NTAB> &T, const RK &rk, const PrfTagPart &g, uint32_t x) {
    uint32_t w0 = x ^ rk[0];
    uint32_t s0 = g.g0 ^ T.t0(byte_of(w0, 0));
    uint32_t s1 = g.g1 ^ T.t3(byte_of(w0, 3));
    uint32_t s2 = g.g2 ^ T.t2(byte_of(w0, 2));
    uint32_t s3 = g.g3 ^ T.t1(byte_of(w0, 1));
#pragma unroll
    for (int r = 2; r < 9; r++) aes_round(T, s0, s1, s2, s3, rk[4 * r], rk[4 * r + 1], rk[4 * r + 2], rk[4 * r + 3]);
    // round 9: only the columns that feed the wanted bytes of final word 0
    uint32_t t0 = T.t0(byte_of(s0, 0)) ^ T.t1(byte_of(s1, 1)) ^ T.t2(byte_of(s2, 2)) ^ T.t3(byte_of(s3, 3)) ^ rk[36];
    uint32_t t1 = T.t0(byte_of(s1, 0)) ^ T.t1(byte_of(s2, 1)) ^ T.t2(byte_of(s3, 2)) ^ T.t3(byte_of(s0, 3)) ^ rk[37];
    uint32_t out = T.sbox(byte_of(t0, 0)) | (T.sbox(byte_of(t1, 1)) << 8);
    if (NB == 4) {
        uint32_t t2 = T.t0(byte_of(s2, 0)) ^ T.t1(byte_of(s3, 1)) ^ T.t2(byte_of(s0, 2)) ^ T.t3(byte_of(s1, 3)) ^ rk[38];
        uint32_t t3 = T.t0(byte_of(s3, 0)) ^ T.t1(byte_of(s0, 1)) ^ T.t2(byte_of(s1, 2)) ^ T.t3(byte_of(s2, 3)) ^ rk[39];
        out |= (T.sbox(byte_of(t2, 2)) << 16) | (T.sbox(byte_of(t3, 3)) << 24);
    }
    return out ^ rk[40] ^ x;  // final AddRoundKey, then the MMO feed-forward (xor with input word 0)
}

// ---- the same PRF as a resumable computation -----------------------------------------------------
// prf_rounds<.., RA, RB> runs AES rounds RA..RB (1-based; 10 = final round + feed-forward) on `st`, so the hint
// kernel can interleave slices of the next group's PRF with the row loads of the current group.  After round 10
// st.s0 holds what prf_low() returns.  Identical arithmetic, different schedule.
//
// XB = number of low bytes of x (the chunk id) that can be non-zero: 1 for SetSize <= 256 (the MS-MARCO and SIFT
// partitions: 196 / 124 chunks), 2 up to 65536 (pir_test's 2^20 rows: 512), 4 = no assumption.  The constant bytes
// of word 0 join the per-hint part, and so does every term of round 2
// that only reads constant columns of the round-1 state:
//   XB = 4:  rounds 1-2 cost 4 + 16 lookups per evaluation  (state after round 1 = g ^ X(x), all four columns vary)
//   XB = 2:  columns 0 and 3 vary                ->  2 +  8
//   XB = 1:  only column 0 varies                ->  1 +  4      (111 lookups per PRF instead of 126)
// Bit-identical to the generic path (tests: test_hintgen_low_bits_path_equals_full_prf, hg_xbytes knob sweeps).
struct PrfState { uint32_t s0, s1, s2, s3; };

template <int XB> struct PrfHint;
template <> struct PrfHint<4> { uint32_t g0, g1, g2, g3; };          // round-1 state without the x terms
template <> struct PrfHint<2> { uint32_t g0, g3, h0, h1, h2, h3; };  // varying columns of round 1 + constant part of round 2
template <> struct PrfHint<1> { uint32_t g0, h0, h1, h2, h3; };

template <int XB, int NTAB, typename RK>
__device__ __forceinline__ PrfHint<XB> prf_hint_part(const AesTab<NTAB> &T, const RK &rk, uint64_t tag) {
    const uint32_t w0 = rk[0], w1 = (uint32_t)(tag << 3) ^ rk[1], w2 = rk[2], w3 = rk[3];
    uint32_t g0 = T.t1(byte_of(w1, 1)) ^ T.t2(byte_of(w2, 2)) ^ T.t3(byte_of(w3, 3)) ^ rk[4];
    uint32_t g1 = T.t0(byte_of(w1, 0)) ^ T.t1(byte_of(w2, 1)) ^ T.t2(byte_of(w3, 2)) ^ rk[5];
    uint32_t g2 = T.t3(byte_of(w1, 3)) ^ T.t0(byte_of(w2, 0)) ^ T.t1(byte_of(w3, 1)) ^ rk[6];
    uint32_t g3 = T.t2(byte_of(w1, 2)) ^ T.t0(byte_of(w3, 0)) ^ T.t3(byte_of(w2, 3)) ^ rk[7];
    PrfHint<XB> g;
    if constexpr (XB == 4) {
        g.g0 = g0; g.g1 = g1; g.g2 = g2; g.g3 = g3;
    } else {
        g1 ^= T.t3(byte_of(w0, 3));   // bytes 2 and 3 of x are zero: word 0 contributes key bytes only
        g2 ^= T.t2(byte_of(w0, 2));
        if constexpr (XB == 1) g3 ^= T.t1(byte_of(w0, 1));
        g.g0 = g0;
        if constexpr (XB == 2) {
            g.g3 = g3;
            g.h0 = T.t1(byte_of(g1, 1)) ^ T.t2(byte_of(g2, 2)) ^ rk[8];
            g.h1 = T.t0(byte_of(g1, 0)) ^ T.t1(byte_of(g2, 1)) ^ rk[9];
            g.h2 = T.t0(byte_of(g2, 0)) ^ T.t3(byte_of(g1, 3)) ^ rk[10];
            g.h3 = T.t2(byte_of(g1, 2)) ^ T.t3(byte_of(g2, 3)) ^ rk[11];
        } else {
            g.h0 = T.t1(byte_of(g1, 1)) ^ T.t2(byte_of(g2, 2)) ^ T.t3(byte_of(g3, 3)) ^ rk[8];
            g.h1 = T.t0(byte_of(g1, 0)) ^ T.t1(byte_of(g2, 1)) ^ T.t2(byte_of(g3, 2)) ^ rk[9];
            g.h2 = T.t0(byte_of(g2, 0)) ^ T.t1(byte_of(g3, 1)) ^ T.t3(byte_of(g1, 3)) ^ rk[10];
            g.h3 = T.t0(byte_of(g3, 0)) ^ T.t2(byte_of(g1, 2)) ^ T.t3(byte_of(g2, 3)) ^ rk[11];
        }
    }
    return g;
}

template <int XB, int NTAB, int NB, int RA, int RB, typename RK>
__device__ __forceinline__ void prf_rounds(const AesTab<NTAB> &T, const RK &rk, const PrfHint<XB> &g, uint32_t x, PrfState &st) {
#pragma unroll
    for (int r = RA; r <= RB; r++) {
        if (r == 1) {
            const uint32_t w0 = x ^ rk[0];
            st.s0 = g.g0 ^ T.t0(byte_of(w0, 0));
            if constexpr (XB == 4) {
                st.s1 = g.g1 ^ T.t3(byte_of(w0, 3));
                st.s2 = g.g2 ^ T.t2(byte_of(w0, 2));
                st.s3 = g.g3 ^ T.t1(byte_of(w0, 1));
            } else if constexpr (XB == 2) {
                st.s3 = g.g3 ^ T.t1(byte_of(w0, 1));
            }
        } else if (r == 2 && XB != 4) {
            const uint32_t s0 = st.s0;
            if constexpr (XB == 2) {
                const uint32_t s3 = st.s3;
                st.s0 = g.h0 ^ T.t0(byte_of(s0, 0)) ^ T.t3(byte_of(s3, 3));
                st.s1 = g.h1 ^ T.t3(byte_of(s0, 3)) ^ T.t2(byte_of(s3, 2));
                st.s2 = g.h2 ^ T.t2(byte_of(s0, 2)) ^ T.t1(byte_of(s3, 1));
                st.s3 = g.h3 ^ T.t1(byte_of(s0, 1)) ^ T.t0(byte_of(s3, 0));
            } else if constexpr (XB == 1) {
                st.s0 = g.h0 ^ T.t0(byte_of(s0, 0));
                st.s1 = g.h1 ^ T.t3(byte_of(s0, 3));
                st.s2 = g.h2 ^ T.t2(byte_of(s0, 2));
                st.s3 = g.h3 ^ T.t1(byte_of(s0, 1));
            }
        } else if (r < 9) {
            aes_round(T, st.s0, st.s1, st.s2, st.s3, rk[4 * r], rk[4 * r + 1], rk[4 * r + 2], rk[4 * r + 3]);
        } else if (r == 9) {
            uint32_t s0 = st.s0, s1 = st.s1, s2 = st.s2, s3 = st.s3;
            st.s0 = T.t0(byte_of(s0, 0)) ^ T.t1(byte_of(s1, 1)) ^ T.t2(byte_of(s2, 2)) ^ T.t3(byte_of(s3, 3)) ^ rk[36];
            st.s1 = T.t0(byte_of(s1, 0)) ^ T.t1(byte_of(s2, 1)) ^ T.t2(byte_of(s3, 2)) ^ T.t3(byte_of(s0, 3)) ^ rk[37];
            if (NB == 4) {
                st.s2 = T.t0(byte_of(s2, 0)) ^ T.t1(byte_of(s3, 1)) ^ T.t2(byte_of(s0, 2)) ^ T.t3(byte_of(s1, 3)) ^ rk[38];
                st.s3 = T.t0(byte_of(s3, 0)) ^ T.t1(byte_of(s0, 1)) ^ T.t2(byte_of(s1, 2)) ^ T.t3(byte_of(s2, 3)) ^ rk[39];
            }
        } else {
            uint32_t out = T.sbox(byte_of(st.s0, 0)) | (T.sbox(byte_of(st.s1, 1)) << 8);
            if (NB == 4) out |= (T.sbox(byte_of(st.s2, 2)) << 16) | (T.sbox(byte_of(st.s3, 3)) << 24);
            st.s0 = out ^ rk[40] ^ x;
        }
    }
}
// rounds of phase PH when the 10 rounds are cut into NPH slices.  With hoisted rounds 1-2 (XB < 4) the four-slice cut
// is 1-3 | 4-5 | 6-7 | 8-10 (21..32 lookups each) instead of 1-2 | 3-5 | 6-7 | 8-10.
template <int PH, int NPH, int XB>
struct PrfPhase {
    static constexpr int cut(int i) { return (NPH == 4 && XB != 4) ? (i == 0 ? 0 : i == 1 ? 3 : i == 2 ? 5 : i == 3 ? 7 : 10) : i * 10 / NPH; }
    static constexpr int first = cut(PH) + 1;
    static constexpr int last = cut(PH + 1);
};

}  // namespace pm
