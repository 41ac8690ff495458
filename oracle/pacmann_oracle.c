/*
 * pacmann_oracle.c -- CPU restatement of the Pacmann hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle and the CPU baseline.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  Nothing under pacmann_b200/ links,
 * imports or calls it; the product path is CUDA-only and fails loudly without its .so.
 *
 * The reference (wuwuz/Pacmann) is Go + Plan-9 assembler and cannot be built here (no Go toolchain),
 * so this is a "port" oracle.  Each function cites the reference file:line it follows and uses the
 * same x86 instructions as the reference .s files through <immintrin.h> (AESENC / AESKEYGENASSIST,
 * 256-bit VPXOR, VSUBPS/VMULPS/VADDPS + VHADDPS, VPMULLD), with a portable C path selected at run
 * time when the host lacks AES-NI / AVX2 / AVX-512 (integer results are identical by construction;
 * the fp32 path keeps the exact 8-lane order).
 *
 * Pinning status: the reference's tests hold NO literal golden vector for the PRF,
 * the hint parities or the server answer ("parity unpinned" at the bit level by the reference
 * itself).  The oracle is pinned by: FIPS-197 vectors (aes_amd64.s is Go's crypto/aes) and the
 * PRF KAT table in tests/golden/ (generated with an independent AES, python `cryptography`);
 * TestXORPerf constants (pir_test.go:279-290); TestInnerProduct's exact scalar equality
 * (graphann_test.go:225-247); the report identities 212.429688 MB / 2130 KB / 3150 KB / window 23
 * (private-search-report.txt:9-21) for parameter derivation; and the functional acceptance tests
 * TestPIRBasic / TestBatchPIRBasic (pir_test.go:9-202) re-run over this code.
 *
 * Determinism: the reference seeds keys and replacement offsets from time.Now() (pir.go:132,208,305)
 * and Go's math/rand stream is not reproducible without Go.  Here keys are injected and every random
 * draw is a counter-based splitmix64 hash, so oracle and CUDA path can be fed identical inputs.
 */
#include <immintrin.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))
#define DEFAULT_PROGRAM_POINT 0x7fffffffULL /* pir.go:15 */
#define REAL_QUERY_PER_PARTITION 2          /* batch-pir.go:13 */
#define QUERY_PER_PARTITION 2               /* batch-pir.go:14 */
#define DEFAULT_VALUE 0xdeadbeefULL         /* batch-pir.go:15 */

/* ------------------------------------------------------------------------------------------- */
/* CPU feature switches                                                                        */
/* ------------------------------------------------------------------------------------------- */
static int g_has_aesni = -1, g_has_avx2 = -1, g_has_avx512 = -1, g_force_portable = 0;

static void detect_cpu(void) {
    if (g_has_aesni >= 0) return;
    __builtin_cpu_init();
    g_has_aesni = __builtin_cpu_supports("aes") && __builtin_cpu_supports("sse4.1");
    g_has_avx2 = __builtin_cpu_supports("avx2");
    g_has_avx512 = __builtin_cpu_supports("avx512f");
}
ORC_API void orc_force_portable(int on) { g_force_portable = on; }
ORC_API int orc_cpu_features(void) {
    detect_cpu();
    return (g_has_aesni ? 1 : 0) | (g_has_avx2 ? 2 : 0) | (g_has_avx512 ? 4 : 0);
}
static inline int use_aesni(void) { detect_cpu(); return g_has_aesni && !g_force_portable; }
static inline int use_avx2(void) { detect_cpu(); return g_has_avx2 && !g_force_portable; }
static inline int use_avx512(void) { detect_cpu(); return g_has_avx512 && !g_force_portable; }

/* ------------------------------------------------------------------------------------------- */
/* Deterministic counter-based randomness shared with the product host code (documented in      */
/* DESIGN.md): splitmix64 finaliser over (seed + (ctr+1) * golden gamma).                        */
/* ------------------------------------------------------------------------------------------- */
ORC_API uint64_t orc_mix64(uint64_t seed, uint64_t ctr) {
    uint64_t z = seed + (ctr + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* ------------------------------------------------------------------------------------------- */
/* Portable FIPS-197 AES-128 (S-box computed, not typed in)                                     */
/* ------------------------------------------------------------------------------------------- */
static uint8_t g_sbox[256];
static int g_sbox_ready = 0;
static uint8_t gf_mul(uint8_t a, uint8_t b) {
    uint8_t p = 0;
    for (int i = 0; i < 8; i++) {
        if (b & 1) p ^= a;
        uint8_t hi = a & 0x80;
        a <<= 1;
        if (hi) a ^= 0x1b;
        b >>= 1;
    }
    return p;
}
static void init_sbox(void) {
    if (g_sbox_ready) return;
    for (int x = 0; x < 256; x++) {
        uint8_t inv = 0;
        if (x) {
            for (int y = 1; y < 256; y++)
                if (gf_mul((uint8_t)x, (uint8_t)y) == 1) { inv = (uint8_t)y; break; }
        }
        uint8_t s = inv, r = inv;
        for (int i = 0; i < 4; i++) { r = (uint8_t)((r << 1) | (r >> 7)); s ^= r; }
        g_sbox[x] = s ^ 0x63;
    }
    g_sbox_ready = 1;
}
ORC_API void orc_sbox(uint8_t out[256]) { init_sbox(); memcpy(out, g_sbox, 256); }

static void expand_key_portable(const uint8_t key[16], uint32_t rk[44]) {
    init_sbox();
    uint8_t *w = (uint8_t *)rk; /* round keys as raw 16-byte blocks: memory order == AES byte order */
    memcpy(w, key, 16);
    uint8_t rcon = 1;
    for (int i = 4; i < 44; i++) {
        uint8_t t[4];
        memcpy(t, w + 4 * (i - 1), 4);
        if ((i & 3) == 0) {
            uint8_t t0 = t[0];
            t[0] = g_sbox[t[1]] ^ rcon; t[1] = g_sbox[t[2]]; t[2] = g_sbox[t[3]]; t[3] = g_sbox[t0];
            rcon = gf_mul(rcon, 2);
        }
        for (int b = 0; b < 4; b++) w[4 * i + b] = w[4 * (i - 4) + b] ^ t[b];
    }
}
static void encrypt_portable(const uint32_t *rk, uint8_t dst[16], const uint8_t src[16]) {
    init_sbox();
    const uint8_t *k = (const uint8_t *)rk;
    uint8_t s[16], t[16];
    for (int i = 0; i < 16; i++) s[i] = src[i] ^ k[i];
    for (int r = 1; r <= 10; r++) {
        for (int c = 0; c < 4; c++)
            for (int row = 0; row < 4; row++) t[4 * c + row] = g_sbox[s[4 * ((c + row) & 3) + row]];
        if (r < 10) {
            for (int c = 0; c < 4; c++) {
                uint8_t a0 = t[4 * c], a1 = t[4 * c + 1], a2 = t[4 * c + 2], a3 = t[4 * c + 3];
                s[4 * c + 0] = gf_mul(a0, 2) ^ gf_mul(a1, 3) ^ a2 ^ a3;
                s[4 * c + 1] = a0 ^ gf_mul(a1, 2) ^ gf_mul(a2, 3) ^ a3;
                s[4 * c + 2] = a0 ^ a1 ^ gf_mul(a2, 2) ^ gf_mul(a3, 3);
                s[4 * c + 3] = gf_mul(a0, 3) ^ a1 ^ a2 ^ gf_mul(a3, 2);
            }
        } else {
            memcpy(s, t, 16);
        }
        for (int i = 0; i < 16; i++) s[i] ^= k[16 * r + i];
    }
    memcpy(dst, s, 16);
}

/* ------------------------------------------------------------------------------------------- */
/* aes_amd64.s restated with intrinsics                                                         */
/* ------------------------------------------------------------------------------------------- */
/* _expand_key_128<> : pianopir/aes_amd64.s:117-126 */
__attribute__((target("aes,sse4.1"))) static inline __m128i expand_step(__m128i x0, __m128i x1, __m128i *x4) {
    x1 = _mm_shuffle_epi32(x1, 0xff);
    *x4 = _mm_castps_si128(_mm_shuffle_ps(_mm_castsi128_ps(*x4), _mm_castsi128_ps(x0), 0x10));
    x0 = _mm_xor_si128(x0, *x4);
    *x4 = _mm_castps_si128(_mm_shuffle_ps(_mm_castsi128_ps(*x4), _mm_castsi128_ps(x0), 0x8c));
    x0 = _mm_xor_si128(x0, *x4);
    return _mm_xor_si128(x0, x1);
}
/* expandKeyAsm : pianopir/aes_amd64.s:87-115 */
__attribute__((target("aes,sse4.1"))) static void expand_key_aesni(const uint8_t key[16], uint32_t rk[44]) {
    __m128i x0 = _mm_loadu_si128((const __m128i *)key), x4 = _mm_setzero_si128();
    __m128i *out = (__m128i *)rk;
    _mm_storeu_si128(out++, x0);
#define STEP(rc) x0 = expand_step(x0, _mm_aeskeygenassist_si128(x0, rc), &x4); _mm_storeu_si128(out++, x0);
    STEP(0x01) STEP(0x02) STEP(0x04) STEP(0x08) STEP(0x10) STEP(0x20) STEP(0x40) STEP(0x80) STEP(0x1b) STEP(0x36)
#undef STEP
}
/* encryptAes128 : pianopir/aes_amd64.s:19-48 */
__attribute__((target("aes,sse4.1"))) static inline __m128i enc_block_aesni(const __m128i *k, __m128i x0) {
    x0 = _mm_xor_si128(x0, _mm_loadu_si128(k));
    for (int r = 1; r < 10; r++) x0 = _mm_aesenc_si128(x0, _mm_loadu_si128(k + r));
    return _mm_aesenclast_si128(x0, _mm_loadu_si128(k + 10));
}

ORC_API void orc_expand_key(const uint8_t key[16], uint32_t rk[44]) {
    if (use_aesni()) expand_key_aesni(key, rk); else expand_key_portable(key, rk);
}
__attribute__((target("aes,sse4.1"))) static void encrypt_aesni(const uint32_t *rk, uint8_t *dst, const uint8_t *src) {
    _mm_storeu_si128((__m128i *)dst, enc_block_aesni((const __m128i *)rk, _mm_loadu_si128((const __m128i *)src)));
}
ORC_API void orc_encrypt_aes128(const uint32_t *rk, uint8_t dst[16], const uint8_t src[16]) {
    if (use_aesni()) encrypt_aesni(rk, dst, src); else encrypt_portable(rk, dst, src);
}
/* aes128MMO : pianopir/aes_amd64.s:51-82  (AES_k(x) xor x) */
__attribute__((target("aes,sse4.1"))) static void mmo_aesni(const uint32_t *rk, uint8_t *dst, const uint8_t *src) {
    __m128i in = _mm_loadu_si128((const __m128i *)src);
    _mm_storeu_si128((__m128i *)dst, _mm_xor_si128(enc_block_aesni((const __m128i *)rk, in), in));
}
ORC_API void orc_aes128_mmo(const uint32_t *rk, uint8_t dst[16], const uint8_t src[16]) {
    if (use_aesni()) { mmo_aesni(rk, dst, src); return; }
    uint8_t t[16];
    encrypt_portable(rk, t, src);
    for (int i = 0; i < 16; i++) dst[i] = t[i] ^ src[i];
}

/* PRFEvalWithLongKeyAndTag : pianopir/util.go:157-165 */
__attribute__((target("aes,sse4.1"))) static inline uint64_t prf_aesni(const uint32_t *rk, uint64_t tag, uint64_t x) {
    __m128i in = _mm_set_epi64x(0, (long long)((tag << 35) + x));
    __m128i o = _mm_xor_si128(enc_block_aesni((const __m128i *)rk, in), in);
    return (uint64_t)_mm_cvtsi128_si64(o);
}
static inline uint64_t prf_any(const uint32_t *rk, uint64_t tag, uint64_t x, int aesni) {
    if (aesni) return prf_aesni(rk, tag, x);
    uint8_t src[16] = {0}, dst[16];
    uint64_t v = (tag << 35) + x;
    memcpy(src, &v, 8); /* little-endian host */
    orc_aes128_mmo(rk, dst, src);
    uint64_t o;
    memcpy(&o, dst, 8);
    return o;
}
ORC_API uint64_t orc_prf(const uint32_t *rk, uint64_t tag, uint64_t x) { return prf_any(rk, tag, x, use_aesni()); }
ORC_API void orc_prf_batch(const uint32_t *rk, const uint64_t *tags, const uint64_t *xs, uint64_t n, uint64_t *out) {
    int a = use_aesni();
    for (uint64_t i = 0; i < n; i++) out[i] = prf_any(rk, tags[i], xs[i], a);
}
/* PRFEval4 : pianopir/util.go:146-155 (key expanded per call, tag-less input) */
ORC_API uint64_t orc_prf_eval4(const uint8_t key[16], uint64_t x) {
    uint32_t rk[44];
    orc_expand_key(key, rk);
    uint8_t src[16] = {0}, dst[16];
    memcpy(src, &x, 8);
    orc_aes128_mmo(rk, dst, src);
    uint64_t o;
    memcpy(&o, dst, 8);
    return o;
}

/* xorSlices : pianopir/aes_amd64.s:133-157.  The asm reads its count from n+32(FP), i.e. len(src),
 * and processes 4 u64 per iteration: the len%4 tail is NOT xored. */
__attribute__((target("avx2"))) static void xor_avx2(uint64_t *dst, const uint64_t *src, int64_t iters) {
    for (int64_t i = 0; i < iters; i++) {
        __m256i a = _mm256_loadu_si256((const __m256i *)(dst + 4 * i));
        __m256i b = _mm256_loadu_si256((const __m256i *)(src + 4 * i));
        _mm256_storeu_si256((__m256i *)(dst + 4 * i), _mm256_xor_si256(a, b));
    }
}
ORC_API void orc_xor_slices(uint64_t *dst, const uint64_t *src, int64_t len_src) {
    int64_t iters = len_src >> 2;
    if (use_avx2()) { xor_avx2(dst, src, iters); return; }
    for (int64_t i = 0; i < 4 * iters; i++) dst[i] ^= src[i];
}
/* EntryXor : pianopir/pir.go:258-265 */
static inline void entry_xor(uint64_t *a, const uint64_t *b, uint64_t e) { orc_xor_slices(a, b, (int64_t)e); }

/* ------------------------------------------------------------------------------------------- */
/* graphann distance kernels                                                                    */
/* ------------------------------------------------------------------------------------------- */
/* L2DistanceSIMD : graphann/l2_distance_amd64.s:4-36 -- 8 fp32 lanes, sub/mul/add each rounded (no
 * FMA), then VEXTRACTF128 + 3x VHADDPS = ((a0+a1)+(a2+a3)) + ((a4+a5)+(a6+a7)).  The asm loop is a
 * do-while: it always runs once, even for n == 0 (never called that way through L2Dist unless
 * dim < 8, where the reference would read out of bounds; we return 0 there). */
__attribute__((target("avx"))) static float l2_simd_avx(const float *a, const float *b, int64_t n) {
    __m256 acc = _mm256_setzero_ps();
    for (int64_t i = 0; i < n; i += 8) {
        __m256 d = _mm256_sub_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i));
        acc = _mm256_add_ps(acc, _mm256_mul_ps(d, d));
    }
    __m128 hi = _mm256_extractf128_ps(acc, 1), lo = _mm256_castps256_ps128(acc);
    lo = _mm_hadd_ps(lo, hi);
    lo = _mm_hadd_ps(lo, lo);
    lo = _mm_hadd_ps(lo, lo);
    return _mm_cvtss_f32(lo);
}
static float l2_simd_portable(const float *a, const float *b, int64_t n) {
    volatile float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; /* volatile: forbid reassociation / contraction */
    for (int64_t i = 0; i < n; i += 8)
        for (int l = 0; l < 8; l++) {
            volatile float d = a[i + l] - b[i + l];
            volatile float p = d * d;
            acc[l] = acc[l] + p;
        }
    volatile float s01 = acc[0] + acc[1], s23 = acc[2] + acc[3], s45 = acc[4] + acc[5], s67 = acc[6] + acc[7];
    volatile float lo = s01 + s23, hi = s45 + s67;
    return lo + hi;
}
ORC_API float orc_l2_distance_simd(const float *a, const float *b, int64_t n) {
    if (n <= 0) return 0.0f;
    return use_avx2() ? l2_simd_avx(a, b, n) : l2_simd_portable(a, b, n);
}
/* L2Dist : graphann/build_graph.go:119-127 (SIMD body over dim - dim%8, scalar tail) */
ORC_API float orc_l2dist(const float *v1, const float *v2, int64_t dim) {
    int64_t rem = dim & 7;
    volatile float d = orc_l2_distance_simd(v1, v2, dim - rem);
    for (int64_t i = dim - rem; i < dim; i++) {
        volatile float x = v1[i] - v2[i];
        volatile float p = x * x;
        d = d + p;
    }
    return d;
}
ORC_API void orc_l2dist_batch(const float *vecs, int64_t row_stride_f32, int64_t dim, const float *queries,
                              const int64_t *ids, int64_t nq, int64_t k, float *out) {
    for (int64_t q = 0; q < nq; q++)
        for (int64_t j = 0; j < k; j++)
            out[q * k + j] = orc_l2dist(vecs + ids[q * k + j] * row_stride_f32, queries + q * dim, dim);
}

/* InnerProduct : graphann/l2_distance_amd64.s:39-68 -- wrapping uint32 dot product, 16 lanes.  The
 * asm loop (SUBQ $16; JNZ) never terminates unless n % 16 == 0; we reject such n with 0. */
__attribute__((target("avx512f"))) static uint32_t ip_avx512(const uint32_t *a, const uint32_t *b, int64_t n) {
    __m512i acc = _mm512_setzero_si512();
    for (int64_t i = 0; i < n; i += 16)
        acc = _mm512_add_epi32(acc, _mm512_mullo_epi32(_mm512_loadu_si512(a + i), _mm512_loadu_si512(b + i)));
    return (uint32_t)_mm512_reduce_add_epi32(acc);
}
__attribute__((target("avx2"))) static uint32_t ip_avx2(const uint32_t *a, const uint32_t *b, int64_t n) {
    __m256i acc = _mm256_setzero_si256();
    for (int64_t i = 0; i < n; i += 8)
        acc = _mm256_add_epi32(acc, _mm256_mullo_epi32(_mm256_loadu_si256((const __m256i *)(a + i)),
                                                       _mm256_loadu_si256((const __m256i *)(b + i))));
    uint32_t t[8];
    _mm256_storeu_si256((__m256i *)t, acc);
    return t[0] + t[1] + t[2] + t[3] + t[4] + t[5] + t[6] + t[7];
}
ORC_API uint32_t orc_inner_product(const uint32_t *a, const uint32_t *b, int64_t n) {
    if (n <= 0 || (n & 15)) return 0;
    if (use_avx512()) return ip_avx512(a, b, n);
    if (use_avx2()) return ip_avx2(a, b, n);
    uint32_t s = 0;
    for (int64_t i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}
/* scan loop of TestInnerProduct : graphann/graphann_test.go:268-273, one checksum per query */
ORC_API void orc_ip_scan(const uint32_t *rows, int64_t n, int64_t d, const uint32_t *queries, int64_t nq,
                         uint32_t *checksum, int threads) {
    for (int64_t q = 0; q < nq; q++) {
        uint32_t sum = 0;
#pragma omp parallel for reduction(+ : sum) num_threads(threads > 0 ? threads : 1) schedule(static)
        for (int64_t i = 0; i < n; i++) sum += orc_inner_product(rows + i * d, queries + q * d, d);
        checksum[q] = sum;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* PianoPIR                                                                                      */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    /* PianoPIRConfig : pir.go:18-26 */
    uint64_t entry_bytes, entry_u64, db_size, chunk_size, set_size, thread_num, fail_log2;
    const uint64_t *raw_db; /* server view, aliased (pir.go:34-38) */
    /* PianoPIRClient : pir.go:91-122 */
    int skip_prep;
    uint8_t master_key[16];
    uint32_t long_key[44];
    uint64_t max_query_num, finished_query_num, max_query_per_chunk;
    uint64_t *query_histogram;
    uint64_t primary_hint_num;
    uint64_t *primary_short_tag, *primary_parity, *primary_program_point;
    uint64_t *replacement_idx, *replacement_val; /* flattened [set][mqpc], [set][mqpc*E] */
    uint64_t *backup_short_tag, *backup_parity;
    /* localCache map[uint64][]uint64 -> open-addressing table */
    uint64_t cache_cap, cache_cnt, *cache_keys, *cache_vals;
    uint8_t *cache_used;
    uint64_t dummy_seed, dummy_ctr;
    uint64_t n_private_queries; /* server PrivateQuery calls, for accounting */
} orc_pir;

/* GenParams : util.go:97-108 / NewPianoPIR : pir.go:487-494 */
ORC_API void orc_gen_params(uint64_t db_size, uint64_t *chunk_size, uint64_t *set_size) {
    uint64_t target = (uint64_t)(2 * sqrt((double)db_size));
    uint64_t c = 1;
    while (c < target) c *= 2;
    uint64_t s = (uint64_t)ceil((double)db_size / (double)c);
    s = (s + 3) / 4 * 4;
    *chunk_size = c;
    *set_size = s;
}
/* NewPianoPIRClient sizes : pir.go:138-142 (ThreadNum = 8 from pir.go:502) */
ORC_API void orc_client_params(uint64_t db_size, uint64_t chunk_size, uint64_t set_size, uint64_t fail_log2,
                               uint64_t *max_query_num, uint64_t *primary_hint_num, uint64_t *max_query_per_chunk) {
    const uint64_t T = 8;
    uint64_t maxq = (uint64_t)(sqrt((double)db_size) * log((double)db_size));
    uint64_t k = (uint64_t)ceil(log(2.0) * (double)(fail_log2 + 1)); /* primaryNumParam : pir.go:124-127 */
    uint64_t p = k * chunk_size;
    p = (p + T - 1) / T * T;
    uint64_t mq = 3 * (uint64_t)((double)maxq / (double)set_size);
    mq = (mq + T - 1) / T * T;
    *max_query_num = maxq;
    *primary_hint_num = p;
    *max_query_per_chunk = mq;
}

static void cache_reset(orc_pir *p) {
    free(p->cache_keys); free(p->cache_vals); free(p->cache_used);
    p->cache_cap = 1;
    while (p->cache_cap < 4 * (p->max_query_num + 16)) p->cache_cap *= 2;
    p->cache_cnt = 0;
    p->cache_keys = calloc(p->cache_cap, 8);
    p->cache_used = calloc(p->cache_cap, 1);
    p->cache_vals = NULL; /* allocated lazily: values are E u64 each */
}
static uint64_t *cache_find(orc_pir *p, uint64_t key, int insert) {
    uint64_t h = orc_mix64(0x1234, key) & (p->cache_cap - 1);
    while (p->cache_used[h]) {
        if (p->cache_keys[h] == key) return p->cache_vals + h * p->entry_u64;
        h = (h + 1) & (p->cache_cap - 1);
    }
    if (!insert) return NULL;
    if (!p->cache_vals) p->cache_vals = calloc(p->cache_cap * p->entry_u64, 8);
    p->cache_used[h] = 1;
    p->cache_keys[h] = key;
    p->cache_cnt++;
    return p->cache_vals + h * p->entry_u64;
}

/* NewPianoPIR : pir.go:479-514  +  NewPianoPIRClient : pir.go:130-175 */
ORC_API orc_pir *orc_pir_new(uint64_t db_size, uint64_t entry_bytes, const uint64_t *raw_db, uint64_t fail_log2) {
    orc_pir *p = calloc(1, sizeof(orc_pir));
    p->entry_bytes = entry_bytes;
    p->entry_u64 = entry_bytes / 8;
    p->db_size = db_size;
    orc_gen_params(db_size, &p->chunk_size, &p->set_size);
    p->thread_num = 8;
    p->fail_log2 = fail_log2;
    p->raw_db = raw_db;
    orc_client_params(db_size, p->chunk_size, p->set_size, fail_log2, &p->max_query_num, &p->primary_hint_num,
                      &p->max_query_per_chunk);
    p->dummy_seed = 0xD00D;
    return p;
}
static void free_tables(orc_pir *p) {
    free(p->query_histogram); free(p->primary_short_tag); free(p->primary_parity); free(p->primary_program_point);
    free(p->replacement_idx); free(p->replacement_val); free(p->backup_short_tag); free(p->backup_parity);
    p->query_histogram = p->primary_short_tag = p->primary_parity = p->primary_program_point = NULL;
    p->replacement_idx = p->replacement_val = p->backup_short_tag = p->backup_parity = NULL;
}
ORC_API void orc_pir_free(orc_pir *p) {
    if (!p) return;
    free_tables(p);
    free(p->cache_keys); free(p->cache_vals); free(p->cache_used);
    free(p);
}

/* Initialization : pir.go:203-255 (key injected instead of time-seeded) */
ORC_API void orc_pir_initialization(orc_pir *p, const uint8_t key[16]) {
    p->finished_query_num = 0;
    memcpy(p->master_key, key, 16);
    orc_expand_key(key, p->long_key);
    free_tables(p);
    uint64_t S = p->set_size, M = p->max_query_per_chunk, E = p->entry_u64, P = p->primary_hint_num;
    p->query_histogram = calloc(S, 8);
    uint64_t tag = 0;
    p->primary_short_tag = calloc(P, 8);
    p->primary_parity = calloc(P * E, 8);
    p->primary_program_point = calloc(P, 8);
    for (uint64_t i = 0; i < P; i++) {
        p->primary_short_tag[i] = tag++;
        p->primary_program_point[i] = DEFAULT_PROGRAM_POINT;
    }
    p->replacement_idx = calloc(S * M, 8);
    p->replacement_val = calloc(S * M * E, 8);
    p->backup_short_tag = calloc(S * M, 8);
    p->backup_parity = calloc(S * M * E, 8);
    for (uint64_t i = 0; i < S; i++)
        for (uint64_t j = 0; j < M; j++) {
            p->replacement_idx[i * M + j] = DEFAULT_PROGRAM_POINT;
            p->backup_short_tag[i * M + j] = tag++;
        }
    cache_reset(p);
}

/* UpdatePreprocessing : pir.go:303-352.  [h0,h1) restricts the hint range (primary hints numbered
 * 0..P-1, backup hints P..P+S*M-1) so OpenMP threads can split the work; the full call is [0, P+S*M).
 * Replacement offsets: counter-based hash of (repl_seed, chunk*M + j) instead of a time-seeded rng. */
static void update_preprocessing(orc_pir *p, uint64_t chunk_id, const uint64_t *chunk, uint64_t h0, uint64_t h1,
                                 uint64_t repl_seed, int do_repl) {
    const uint64_t E = p->entry_u64, C = p->chunk_size, S = p->set_size, M = p->max_query_per_chunk,
                   P = p->primary_hint_num;
    const int aesni = use_aesni();
    for (uint64_t i = h0; i < h1 && i < P; i++) {
        uint64_t off = prf_any(p->long_key, p->primary_short_tag[i], chunk_id, aesni) & (C - 1);
        entry_xor(p->primary_parity + i * E, chunk + off * E, E);
    }
    for (uint64_t i = 0; i < S; i++) {
        if (i == chunk_id) continue;
        for (uint64_t j = 0; j < M; j++) {
            uint64_t h = P + i * M + j;
            if (h < h0 || h >= h1) continue;
            uint64_t off = prf_any(p->long_key, p->backup_short_tag[i * M + j], chunk_id, aesni) & (C - 1);
            entry_xor(p->backup_parity + (i * M + j) * E, chunk + off * E, E);
        }
    }
    if (do_repl)
        for (uint64_t j = 0; j < M; j++) {
            uint64_t off = orc_mix64(repl_seed, chunk_id * M + j) & (C - 1);
            p->replacement_idx[chunk_id * M + j] = off + chunk_id * C;
            memcpy(p->replacement_val + (chunk_id * M + j) * E, chunk + off * E, E * 8);
        }
}

/* Preprocessing : pir.go:267-301 (zero-padded tmpChunk for the ragged tail) */
static void preprocessing_range(orc_pir *p, uint64_t h0, uint64_t h1, uint64_t repl_seed, int do_repl) {
    const uint64_t E = p->entry_u64, C = p->chunk_size, S = p->set_size;
    const uint64_t len_db = p->db_size * E;
    uint64_t *tmp = NULL;
    for (uint64_t i = 0; i < S; i++) {
        uint64_t start = i * C, end = (i + 1) * C;
        if (end * E > len_db) {
            if (!tmp) tmp = malloc(C * E * 8);
            for (uint64_t j = start * E; j < end * E; j++) tmp[j - start * E] = (j >= len_db) ? 0 : p->raw_db[j];
            update_preprocessing(p, i, tmp, h0, h1, repl_seed, do_repl);
        } else {
            update_preprocessing(p, i, p->raw_db + start * E, h0, h1, repl_seed, do_repl);
        }
    }
    free(tmp);
}
ORC_API void orc_pir_preprocessing(orc_pir *p, const uint8_t key[16], uint64_t repl_seed, int threads) {
    orc_pir_initialization(p, key);
    if (p->skip_prep) return;
    uint64_t H = p->primary_hint_num + p->set_size * p->max_query_per_chunk;
    if (threads <= 1) { preprocessing_range(p, 0, H, repl_seed, 1); return; }
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int t = 0; t < threads; t++) {
        uint64_t a = H * (uint64_t)t / (uint64_t)threads, b = H * (uint64_t)(t + 1) / (uint64_t)threads;
        preprocessing_range(p, a, b, repl_seed, t == 0);
    }
}
/* hint-set shard of Preprocessing: only hints [h0,h1) are computed (the others stay zero) -- what one GPU of an
 * N-GPU run owns.  Replacement values are produced only when do_repl != 0. */
ORC_API void orc_pir_preprocessing_range(orc_pir *p, const uint8_t key[16], uint64_t repl_seed, uint64_t h0, uint64_t h1,
                                         int do_repl) {
    orc_pir_initialization(p, key);
    preprocessing_range(p, h0, h1, repl_seed, do_repl);
}
/* DummyPreprocessing : pir.go:520-523 */
ORC_API void orc_pir_dummy_preprocessing(orc_pir *p, const uint8_t key[16]) {
    orc_pir_initialization(p, key);
    p->skip_prep = 1;
}

/* PrivateQuery : pir.go:65-88 */
ORC_API void orc_pir_private_query(orc_pir *p, const uint32_t *offsets, uint64_t *ret) {
    const uint64_t E = p->entry_u64;
    memset(ret, 0, E * 8);
    for (uint64_t i = 0; i < p->set_size; i++) {
        uint64_t idx = (uint64_t)offsets[i] + i * p->chunk_size;
        if (idx >= p->db_size) continue;
        entry_xor(ret, p->raw_db + idx * E, E);
    }
    p->n_private_queries++;
}
/* NonePrivateQuery : pir.go:41-62.  returns 0 ok, 1 = out of range error */
ORC_API int orc_pir_nonprivate_query(orc_pir *p, uint64_t idx, uint64_t *ret) {
    memset(ret, 0, p->entry_u64 * 8);
    if (idx >= p->db_size) return idx < p->chunk_size * p->set_size ? 0 : 1;
    memcpy(ret, p->raw_db + idx * p->entry_u64, p->entry_u64 * 8);
    return 0;
}

/* Query : pir.go:354-471.  Return codes: 0 ok, 1 out of range, 2 query budget exceeded,
 * 3 too many queries in chunk, 4 no hit hint.  `ret` receives the (possibly zero) entry.
 * If offsets_out != NULL the offset vector sent to the server is copied there. */
ORC_API int orc_pir_client_query(orc_pir *p, uint64_t idx, int real_query, uint64_t *ret, uint32_t *offsets_out) {
    const uint64_t E = p->entry_u64, C = p->chunk_size, S = p->set_size, M = p->max_query_per_chunk;
    const int aesni = use_aesni();
    memset(ret, 0, E * 8);
    uint32_t *offs = malloc(S * 4);
    if (!real_query) {
        for (uint64_t i = 0; i < S; i++) offs[i] = (uint32_t)(orc_mix64(p->dummy_seed, p->dummy_ctr++) & (C - 1));
        uint64_t *tmp = malloc(E * 8);
        orc_pir_private_query(p, offs, tmp);
        if (offsets_out) memcpy(offsets_out, offs, S * 4);
        free(tmp); free(offs);
        return 0;
    }
    if (idx >= p->db_size) { free(offs); return 1; }
    uint64_t *cached = cache_find(p, idx, 0);
    if (cached) { memcpy(ret, cached, E * 8); free(offs); return 0; }
    if (p->finished_query_num >= p->max_query_num) { free(offs); return 2; }
    uint64_t chunk_id = idx / C, offset = idx % C;
    if (p->query_histogram[chunk_id] >= M) { free(offs); return 3; }

    uint64_t hit = DEFAULT_PROGRAM_POINT;
    for (uint64_t i = 0; i < p->primary_hint_num; i++) {
        uint64_t ho = prf_any(p->long_key, p->primary_short_tag[i], chunk_id, aesni) & (C - 1);
        if (ho == offset) {
            if (p->primary_program_point[i] == DEFAULT_PROGRAM_POINT || (p->primary_program_point[i] / C != chunk_id)) {
                hit = i;
                break;
            }
        }
    }
    if (hit == DEFAULT_PROGRAM_POINT) { free(offs); return 4; }

    uint64_t *query_set = malloc(S * 8);
    for (uint64_t i = 0; i < S; i++)
        query_set[i] = i * C + (prf_any(p->long_key, p->primary_short_tag[hit], i, aesni) & (C - 1));
    if (p->primary_program_point[hit] != DEFAULT_PROGRAM_POINT)
        query_set[p->primary_program_point[hit] / C] = p->primary_program_point[hit];
    uint64_t in_group = p->query_histogram[chunk_id];
    uint64_t repl_idx = p->replacement_idx[chunk_id * M + in_group];
    const uint64_t *repl_val = p->replacement_val + (chunk_id * M + in_group) * E;
    query_set[chunk_id] = repl_idx;
    for (uint64_t i = 0; i < S; i++) offs[i] = (uint32_t)(query_set[i] & (C - 1));
    if (offsets_out) memcpy(offsets_out, offs, S * 4);

    orc_pir_private_query(p, offs, ret);
    entry_xor(ret, repl_val, E);
    entry_xor(ret, p->primary_parity + hit * E, E);

    p->primary_short_tag[hit] = p->backup_short_tag[chunk_id * M + in_group];
    memcpy(p->primary_parity + hit * E, p->backup_parity + (chunk_id * M + in_group) * E, E * 8);
    p->primary_program_point[hit] = idx;
    entry_xor(p->primary_parity + hit * E, ret, E);

    p->finished_query_num += 1;
    p->query_histogram[chunk_id] += 1;
    memcpy(cache_find(p, idx, 1), ret, E * 8);
    free(query_set); free(offs);
    return 0;
}
/* PianoPIR.Query : pir.go:525-533 (re-preprocess when the budget is exactly spent) */
ORC_API int orc_pir_query(orc_pir *p, uint64_t idx, int real_query, uint64_t *ret, const uint8_t rekey[16],
                          uint64_t repl_seed) {
    if (p->finished_query_num == p->max_query_num) orc_pir_preprocessing(p, rekey, repl_seed, 1);
    return orc_pir_client_query(p, idx, real_query, ret, NULL);
}

/* LocalStorageSize : pir.go:178-190 */
ORC_API double orc_pir_local_storage(const orc_pir *p) {
    double s = 0, P = (double)p->primary_hint_num, B = (double)p->set_size * (double)p->max_query_per_chunk,
           EB = (double)p->entry_bytes;
    s += P * 8 + P * EB + P * 8;
    s += B * 8 + B * EB + B * 8 + B * EB;
    return s;
}
/* CommCostPerQuery : pir.go:539-544 */
ORC_API double orc_pir_comm_cost(const orc_pir *p) { return (double)(p->set_size * 4 + p->entry_u64 * 8); }

/* field access for tests */
ORC_API uint64_t orc_pir_get(const orc_pir *p, int what) {
    switch (what) {
    case 0: return p->entry_u64;
    case 1: return p->db_size;
    case 2: return p->chunk_size;
    case 3: return p->set_size;
    case 4: return p->max_query_num;
    case 5: return p->primary_hint_num;
    case 6: return p->max_query_per_chunk;
    case 7: return p->finished_query_num;
    case 8: return p->n_private_queries;
    }
    return 0;
}
ORC_API uint64_t *orc_pir_table(orc_pir *p, int what) {
    switch (what) {
    case 0: return p->primary_short_tag;
    case 1: return p->primary_parity;
    case 2: return p->primary_program_point;
    case 3: return p->replacement_idx;
    case 4: return p->replacement_val;
    case 5: return p->backup_short_tag;
    case 6: return p->backup_parity;
    case 7: return p->query_histogram;
    }
    return NULL;
}
ORC_API const uint32_t *orc_pir_long_key(const orc_pir *p) { return p->long_key; }
ORC_API void orc_pir_set_dummy_seed(orc_pir *p, uint64_t seed) { p->dummy_seed = seed; p->dummy_ctr = 0; }

/* ------------------------------------------------------------------------------------------- */
/* SimpleBatchPianoPIR : batch-pir.go                                                            */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t entry_bytes, entry_u64, db_size, batch_size, partition_num, partition_size, fail_log2;
    orc_pir **sub;
    uint64_t finished_batch_num, queries_made_in_partition, support_batch_num;
    uint64_t key_seed, repl_seed;
    uint64_t *sub_epoch; /* preprocessings made so far per sub-PIR: every one draws a fresh key (pir.go:208-211) */
} orc_batch;

/* NewSimpleBatchPianoPIR : batch-pir.go:55-93 */
ORC_API orc_batch *orc_batch_new(uint64_t db_size, uint64_t entry_bytes, uint64_t batch_size, const uint64_t *raw_db,
                                 uint64_t fail_log2) {
    orc_batch *b = calloc(1, sizeof(orc_batch));
    b->entry_bytes = entry_bytes;
    b->entry_u64 = entry_bytes / 8;
    b->db_size = db_size;
    b->batch_size = batch_size;
    b->partition_num = batch_size / REAL_QUERY_PER_PARTITION;
    b->partition_size = (db_size + b->partition_num - 1) / b->partition_num;
    b->fail_log2 = fail_log2;
    b->sub = calloc(b->partition_num, sizeof(orc_pir *));
    b->sub_epoch = calloc(b->partition_num, sizeof(uint64_t));
    for (uint64_t i = 0; i < b->partition_num; i++) {
        uint64_t start = i * b->partition_size, end = (i + 1) * b->partition_size;
        if (end > db_size) end = db_size;
        b->sub[i] = orc_pir_new(end - start, entry_bytes, raw_db + start * b->entry_u64, fail_log2);
        orc_pir_set_dummy_seed(b->sub[i], orc_mix64(0xD00D, i));
    }
    return b;
}
ORC_API void orc_batch_free(orc_batch *b) {
    if (!b) return;
    for (uint64_t i = 0; i < b->partition_num; i++) orc_pir_free(b->sub[i]);
    free(b->sub);
    free(b->sub_epoch);
    free(b);
}
ORC_API orc_pir *orc_batch_sub(orc_batch *b, uint64_t i) { return b->sub[i]; }
ORC_API uint64_t orc_batch_get(const orc_batch *b, int what) {
    switch (what) {
    case 0: return b->partition_num;
    case 1: return b->partition_size;
    case 2: return b->finished_batch_num;
    case 3: return b->queries_made_in_partition;
    case 4: return b->support_batch_num;
    }
    return 0;
}
/* key for partition i at preprocessing epoch e: 16 bytes = LE(mix(seed, 2*(e*parts+i))) || LE(mix(.., +1)),
 * the analogue of RandKey128's two rng.Uint64() draws (util.go:25-31). */
ORC_API void orc_derive_key(uint64_t key_seed, uint64_t epoch, uint64_t parts, uint64_t i, uint8_t key[16]) {
    uint64_t a = orc_mix64(key_seed, 2 * (epoch * parts + i)), c = orc_mix64(key_seed, 2 * (epoch * parts + i) + 1);
    memcpy(key, &a, 8);
    memcpy(key + 8, &c, 8);
}
/* Preprocessing : batch-pir.go:119-155 + RecordStats :110-117.  `threads` = goroutine fan-out width
 * (the reference's ThreadNum const is 1); threads > partitions are spent on hint ranges. */
ORC_API void orc_batch_preprocessing(orc_batch *b, uint64_t key_seed, uint64_t repl_seed, int threads) {
    b->finished_batch_num = 0;
    b->queries_made_in_partition = 0;
    b->key_seed = key_seed;
    b->repl_seed = repl_seed;
    int P = (int)b->partition_num;
    if (threads < 1) threads = 1;
    int outer = threads < P ? threads : P, inner = threads / outer;
    if (inner < 1) inner = 1;
#ifdef _OPENMP
    omp_set_max_active_levels(2);
#endif
#pragma omp parallel for num_threads(outer) schedule(dynamic, 1)
    for (int i = 0; i < P; i++) {
        uint8_t key[16];
        const uint64_t e = b->sub_epoch[i]++;
        orc_derive_key(key_seed, e, b->partition_num, (uint64_t)i, key);
        orc_pir_preprocessing(b->sub[i], key, orc_mix64(repl_seed, e * b->partition_num + (uint64_t)i), inner);
    }
    b->support_batch_num = b->sub[0]->max_query_num / QUERY_PER_PARTITION;
}
/* DummyPreprocessing : batch-pir.go:157-166 */
ORC_API void orc_batch_dummy_preprocessing(orc_batch *b, uint64_t key_seed) {
    for (uint64_t i = 0; i < b->partition_num; i++) {
        uint8_t key[16];
        orc_derive_key(key_seed, b->sub_epoch[i]++, b->partition_num, i, key);
        orc_pir_dummy_preprocessing(b->sub[i], key);
    }
    b->support_batch_num = b->sub[0]->max_query_num / QUERY_PER_PARTITION;
}

/* Query : batch-pir.go:170-248.  out is [n][E]; responses keyed by global idx (duplicates share the
 * last stored answer), zero rows for misses.  status_out[i] (optional) = per-input sub-query return
 * code, or -1 when the index was dropped (surplus in its partition). */
ORC_API int orc_batch_query(orc_batch *b, const uint64_t *idx, uint64_t n, uint64_t *out, int *status_out) {
    const uint64_t E = b->entry_u64, PN = b->partition_num, PS = b->partition_size;
    uint64_t to_make = n / PN;
    uint64_t *cnt = calloc(PN, 8), *lists = malloc((n + 1) * PN * 8);
    for (uint64_t i = 0; i < n; i++) {
        uint64_t pi = idx[i] / PS;
        if (pi >= PN) { free(cnt); free(lists); return -1; } /* Go would panic: index out of range */
        lists[pi * n + cnt[pi]++] = idx[i];
    }
    /* responses map[uint64][]uint64 */
    uint64_t *resp_key = malloc((n + 1) * 8), *resp_val = calloc((n + 1) * E, 8);
    int *resp_code = malloc((n + 1) * sizeof(int));
    uint64_t nresp = 0;
    uint64_t *tmp = malloc(E * 8);
    for (uint64_t i = 0; i < PN; i++) {
        for (uint64_t j = 0; j < to_make; j++) {
            /* PianoPIR.Query (pir.go:525-533): a sub-PIR whose budget is exactly spent re-preprocesses first, under
             * the key of ITS next epoch */
            if (b->sub[i]->finished_query_num == b->sub[i]->max_query_num) {
                uint8_t key[16];
                const uint64_t e = b->sub_epoch[i]++;
                orc_derive_key(b->key_seed, e, PN, i, key);
                orc_pir_preprocessing(b->sub[i], key, orc_mix64(b->repl_seed, e * PN + i), 1);
            }
            if (j >= cnt[i] || lists[i * n + j] == DEFAULT_VALUE) {
                orc_pir_client_query(b->sub[i], 0, 0, tmp, NULL);
            } else {
                uint64_t g = lists[i * n + j];
                int rc = orc_pir_client_query(b->sub[i], g - i * PS, 1, tmp, NULL);
                uint64_t r = 0;
                while (r < nresp && resp_key[r] != g) r++;
                if (r == nresp) nresp++;
                resp_key[r] = g;
                resp_code[r] = rc;
                memcpy(resp_val + r * E, tmp, E * 8);
            }
        }
    }
    for (uint64_t i = 0; i < n; i++) {
        uint64_t r = 0;
        while (r < nresp && resp_key[r] != idx[i]) r++;
        if (r < nresp) {
            memcpy(out + i * E, resp_val + r * E, E * 8);
            if (status_out) status_out[i] = resp_code[r];
        } else {
            memset(out + i * E, 0, E * 8);
            if (status_out) status_out[i] = -1;
        }
    }
    int redo = 0;
    if (b->queries_made_in_partition >= b->sub[0]->max_query_num - 2) {
        orc_batch_preprocessing(b, b->key_seed, b->repl_seed, 1);
        redo = 1;
    } else {
        b->finished_batch_num += n / b->batch_size;
        b->queries_made_in_partition += to_make;
    }
    free(cnt); free(lists); free(resp_key); free(resp_val); free(resp_code); free(tmp);
    return redo;
}
/* LocalStorageSize : batch-pir.go:250-256 ; CommCostPerBatchOnline : :258-264 */
ORC_API double orc_batch_local_storage(const orc_batch *b) {
    double r = 0;
    for (uint64_t i = 0; i < b->partition_num; i++) r += orc_pir_local_storage(b->sub[i]);
    return r;
}
ORC_API uint64_t orc_batch_comm_online(const orc_batch *b) {
    double r = 0;
    for (uint64_t i = 0; i < b->partition_num; i++) r += orc_pir_comm_cost(b->sub[i]) * (double)QUERY_PER_PARTITION;
    return (uint64_t)r;
}

/* ------------------------------------------------------------------------------------------- */
/* private-search.go wire format                                                                */
/* ------------------------------------------------------------------------------------------- */
/* PIRGraphInfo.Preprocess packing : private-search.go:371-397.  entry = dim LE f32 || m LE u32 */
ORC_API void orc_pack_db(const float *vectors, const int32_t *graph, uint64_t n, uint64_t dim, uint64_t m,
                         uint64_t *raw_db) {
    uint64_t eb = dim * 4 + m * 4;
    uint8_t *o = (uint8_t *)raw_db;
    for (uint64_t i = 0; i < n; i++) {
        memcpy(o + i * eb, vectors + i * dim, dim * 4);
        for (uint64_t j = 0; j < m; j++) {
            uint32_t v = (uint32_t)graph[i * m + j];
            memcpy(o + i * eb + dim * 4 + j * 4, &v, 4);
        }
    }
}
/* Entry2VectorAndNeighbors : private-search.go:418-439 */
ORC_API void orc_unpack_entry(const uint64_t *entry, uint64_t dim, uint64_t m, float *vector, int64_t *neighbors) {
    const uint8_t *e = (const uint8_t *)entry;
    memcpy(vector, e, dim * 4);
    for (uint64_t j = 0; j < m; j++) {
        uint32_t v;
        memcpy(&v, e + dim * 4 + j * 4, 4);
        neighbors[j] = (int64_t)v;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* SearchKNN : graphann/search.go:114-234                                                        */
/* ------------------------------------------------------------------------------------------- */
/* Tie rules (Go's own order on equal distances is unspecified): the start-vertex
 * sort is stable in input order; the explore queue replicates container/heap's binary heap exactly
 * (Less = dist <); the final ranking orders by (dist, id). */
typedef struct { float dist; int64_t id; } vd_t;
typedef struct { vd_t *a; int64_t n, cap; } heap_t;
static void heap_up(heap_t *h, int64_t j) { /* container/heap.up */
    for (;;) {
        int64_t i = (j - 1) / 2;
        if (i == j || j <= 0 || !(h->a[j].dist < h->a[i].dist)) break;
        vd_t t = h->a[i]; h->a[i] = h->a[j]; h->a[j] = t;
        j = i;
    }
}
static void heap_down(heap_t *h, int64_t i0, int64_t n) { /* container/heap.down */
    int64_t i = i0;
    for (;;) {
        int64_t j1 = 2 * i + 1;
        if (j1 >= n || j1 < 0) break;
        int64_t j = j1, j2 = j1 + 1;
        if (j2 < n && h->a[j2].dist < h->a[j1].dist) j = j2;
        if (!(h->a[j].dist < h->a[i].dist)) break;
        vd_t t = h->a[i]; h->a[i] = h->a[j]; h->a[j] = t;
        i = j;
    }
}
static void heap_push(heap_t *h, vd_t v) {
    if (h->n == h->cap) { h->cap = h->cap ? 2 * h->cap : 64; h->a = realloc(h->a, h->cap * sizeof(vd_t)); }
    h->a[h->n++] = v;
    heap_up(h, h->n - 1);
}
static vd_t heap_pop(heap_t *h) {
    int64_t n = h->n - 1;
    vd_t t = h->a[0]; h->a[0] = h->a[n]; h->a[n] = t;
    heap_down(h, 0, n);
    h->n = n;
    return h->a[n];
}
static int cmp_vd_stable(const void *x, const void *y) { /* used with an index payload for stability */
    const vd_t *a = x, *b = y;
    if (a->dist < b->dist) return -1;
    if (b->dist < a->dist) return 1;
    return 0;
}
static int cmp_vd_id(const void *x, const void *y) {
    const vd_t *a = x, *b = y;
    if (a->dist < b->dist) return -1;
    if (b->dist < a->dist) return 1;
    return (a->id > b->id) - (a->id < b->id);
}
static void stable_sort_vd(vd_t *a, int64_t n) { /* insertion-merge: simple stable merge sort */
    if (n < 2) return;
    vd_t *tmp = malloc(n * sizeof(vd_t));
    for (int64_t w = 1; w < n; w *= 2) {
        for (int64_t lo = 0; lo < n; lo += 2 * w) {
            int64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n, i = lo, j = mid, k = lo;
            while (i < mid && j < hi) tmp[k++] = (cmp_vd_stable(&a[j], &a[i]) < 0) ? a[j++] : a[i++];
            while (i < mid) tmp[k++] = a[i++];
            while (j < hi) tmp[k++] = a[j++];
        }
        memcpy(a, tmp, n * sizeof(vd_t));
    }
    free(tmp);
}

/* vertex source: fills vectors [n][dim] and neighbors [n][m] for ids; mirrors GetVertexInfo (search.go:23) */
typedef int (*orc_vertex_fn)(void *ctx, const int64_t *ids, int64_t n, float *vectors, int64_t *neighbors);

typedef struct {
    const float *vectors; const int32_t *graph; int64_t n, dim, m;
} orc_basic_graph;
/* BasicGraphInfo.GetVertexInfo : search.go:43-49 */
static int basic_vertex_info(void *ctx, const int64_t *ids, int64_t n, float *vectors, int64_t *neighbors) {
    orc_basic_graph *g = ctx;
    for (int64_t i = 0; i < n; i++) {
        memcpy(vectors + i * g->dim, g->vectors + ids[i] * g->dim, g->dim * 4);
        for (int64_t j = 0; j < g->m; j++) neighbors[i * g->m + j] = g->graph[ids[i] * g->m + j];
    }
    return 0;
}
typedef struct {
    orc_batch *pir; int64_t dim, m; int64_t total_q, succ_q; const int32_t *graph;
} orc_pir_graph;
/* PIRGraphInfo.GetVertexInfo : private-search.go:441-506 (private mode) */
static int pir_vertex_info(void *ctx, const int64_t *ids, int64_t n, float *vectors, int64_t *neighbors) {
    orc_pir_graph *g = ctx;
    uint64_t E = g->pir->entry_u64;
    uint64_t *idx = calloc((size_t)n + 1, 8), *resp = malloc(((size_t)n + 1) * E * 8);
    for (int64_t i = 0; i < n; i++) idx[i] = (uint64_t)ids[i];
    g->total_q += n;
    if (orc_batch_query(g->pir, idx, (uint64_t)n, resp, NULL) < 0) { free(idx); free(resp); return -1; }
    for (int64_t i = 0; i < n; i++) {
        orc_unpack_entry(resp + i * E, g->dim, g->m, vectors + i * g->dim, neighbors + i * g->m);
        int ok = 1;
        for (int64_t j = 0; j < g->m; j++)
            if (neighbors[i * g->m + j] != (int64_t)(uint32_t)g->graph[ids[i] * g->m + j]) { ok = 0; break; }
        g->succ_q += ok;
    }
    free(idx); free(resp);
    return 0;
}

static int search_knn(orc_vertex_fn fn, void *ctx, int64_t n, int64_t dim, int64_t m, const int64_t *start_ids,
                      const float *start_vecs, const int64_t *start_nbrs, int64_t n_start, const float *query,
                      int64_t k, int64_t max_step, int64_t parallel, int benchmarking, uint64_t rand_seed,
                      int64_t *ret, int64_t *step_ret) {
    /* knownVertices / reachStep maps -> dense arrays indexed by vertex id */
    int32_t *reach = malloc(n * 4);
    for (int64_t i = 0; i < n; i++) reach[i] = -2; /* -2 = unknown */
    int64_t known_cap = n_start + max_step * parallel * m + 8, known_n = 0;
    int64_t *known_id = malloc(known_cap * 8);
    float *known_vec = malloc(known_cap * dim * 4);
    int64_t *known_nbr = malloc(known_cap * m * 8);
    int64_t *slot_of = malloc(n * 8);
    heap_t heap = {0};
    uint64_t rctr = 0;

    if (!benchmarking) { /* search.go:129-148 */
        vd_t *fs = malloc((n_start + 1) * sizeof(vd_t));
        for (int64_t i = 0; i < n_start; i++) {
            fs[i].dist = orc_l2dist(start_vecs + i * dim, query, dim);
            fs[i].id = i; /* index into the start list; stable sort keeps input order on ties */
        }
        stable_sort_vd(fs, n_start);
        for (int64_t i = 0; heap.n < parallel && i < n_start; i++) {
            int64_t s = fs[i].id, id = start_ids[s];
            if (reach[id] != -2) continue;
            slot_of[id] = known_n;
            known_id[known_n] = id;
            memcpy(known_vec + known_n * dim, start_vecs + s * dim, dim * 4);
            memcpy(known_nbr + known_n * m, start_nbrs + s * m, m * 8);
            known_n++;
            vd_t v = {fs[i].dist, id};
            heap_push(&heap, v);
            reach[id] = 0;
        }
        free(fs);
    }
    int64_t bq_cap = parallel * m;
    int64_t *batch = malloc(bq_cap * 8), *r_nbr = malloc(bq_cap * m * 8);
    float *r_vec = malloc(bq_cap * dim * 4);
    for (int64_t step = 0; step < max_step; step++) { /* search.go:150-208 */
        int64_t bn = 0;
        for (int64_t rept = 0; rept < parallel; rept++) {
            if (heap.n == 0 || benchmarking) {
                for (int64_t i = 0; i < m; i++) batch[bn++] = (int64_t)(orc_mix64(rand_seed, rctr++) % (uint64_t)n);
            } else {
                vd_t it = heap_pop(&heap);
                memcpy(batch + bn, known_nbr + slot_of[it.id] * m, m * 8);
                bn += m;
            }
        }
        if (fn(ctx, batch, bn, r_vec, r_nbr) != 0) return -1;
        if (benchmarking) continue;
        for (int64_t i = 0; i < bn; i++) {
            int64_t id = batch[i];
            if (reach[id] != -2) continue;
            int ok = 0;
            for (int64_t j = 0; j < m; j++)
                if (r_nbr[i * m + j] != 0) { ok = 1; break; }
            if (!ok) continue;
            slot_of[id] = known_n;
            known_id[known_n] = id;
            memcpy(known_vec + known_n * dim, r_vec + i * dim, dim * 4);
            memcpy(known_nbr + known_n * m, r_nbr + i * m, m * 8);
            known_n++;
            reach[id] = (int32_t)step;
            vd_t v = {orc_l2dist(r_vec + i * dim, query, dim), id};
            heap_push(&heap, v);
        }
    }
    /* search.go:210-233 */
    vd_t *all = malloc((known_n + 1) * sizeof(vd_t));
    for (int64_t i = 0; i < known_n; i++) {
        all[i].dist = orc_l2dist(known_vec + i * dim, query, dim);
        all[i].id = known_id[i];
    }
    qsort(all, known_n, sizeof(vd_t), cmp_vd_id);
    for (int64_t i = 0; i < k; i++) {
        if (i >= known_n) { ret[i] = -1; step_ret[i] = -1; }
        else { ret[i] = all[i].id; step_ret[i] = reach[all[i].id]; }
    }
    free(all); free(batch); free(r_nbr); free(r_vec); free(heap.a);
    free(reach); free(known_id); free(known_vec); free(known_nbr); free(slot_of);
    return 0;
}

/* Non-private search over a BasicGraphInfo-like dataset with explicit start vertices. */
ORC_API int orc_search_knn_basic(const float *vectors, const int32_t *graph, int64_t n, int64_t dim, int64_t m,
                                 const int64_t *start_ids, int64_t n_start, const float *queries, int64_t nq,
                                 int64_t k, int64_t max_step, int64_t parallel, int64_t *ret, int64_t *step_ret) {
    orc_basic_graph g = {vectors, graph, n, dim, m};
    float *sv = malloc(n_start * dim * 4);
    int64_t *sn = malloc(n_start * m * 8);
    basic_vertex_info(&g, start_ids, n_start, sv, sn);
    int rc = 0;
    for (int64_t q = 0; q < nq && rc == 0; q++)
        rc = search_knn(basic_vertex_info, &g, n, dim, m, start_ids, sv, sn, n_start, queries + q * dim, k, max_step,
                        parallel, 0, 0, ret + q * k, step_ret + q * k);
    free(sv); free(sn);
    return rc;
}
/* Private search: vertices fetched through the batch PIR (sequential client state, as the reference). */
ORC_API int orc_search_knn_private(orc_batch *pir, const float *vectors, const int32_t *graph, int64_t n, int64_t dim,
                                   int64_t m, const int64_t *start_ids, int64_t n_start, const float *queries,
                                   int64_t nq, int64_t k, int64_t max_step, int64_t parallel, int benchmarking,
                                   uint64_t rand_seed, int64_t *ret, int64_t *step_ret, int64_t *stats) {
    orc_basic_graph g = {vectors, graph, n, dim, m};
    orc_pir_graph pg = {pir, dim, m, 0, 0, graph};
    float *sv = malloc(n_start * dim * 4);
    int64_t *sn = malloc(n_start * m * 8);
    basic_vertex_info(&g, start_ids, n_start, sv, sn); /* start vertices are plaintext : private-search.go:508-531 */
    int rc = 0;
    for (int64_t q = 0; q < nq && rc == 0; q++)
        rc = search_knn(pir_vertex_info, &pg, n, dim, m, start_ids, sv, sn, n_start, queries + q * dim, k, max_step,
                        parallel, benchmarking, orc_mix64(rand_seed, (uint64_t)q), ret + q * k, step_ret + q * k);
    if (stats) { stats[0] = pg.total_q; stats[1] = pg.succ_q; }
    free(sv); free(sn);
    return rc;
}


/* robustPrune (graphann/build_graph.go:169-236): prune the candidates of vertex u to at most m out-neighbours.
 * vectors [n][dim].  sort.Slice is unstable in Go: candidates at equal distance from u may come out in any order; this
 * restatement (and the GPU path) keeps them in candidate order (a stable sort).  Returns the number of ids written. */
typedef struct { int64_t id; float dist; int64_t pos; } orc_idd;
static int orc_idd_cmp(const void *a, const void *b) {
    const orc_idd *x = (const orc_idd *)a, *y = (const orc_idd *)b;
    if (x->dist < y->dist) return -1;
    if (x->dist > y->dist) return 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos ? 1 : 0);
}
ORC_API int64_t orc_robust_prune(const float *vectors, int64_t dim, int64_t u, const int64_t *candidates, int64_t n_cand, int64_t m,
                                 float alpha, int64_t *out) {
    if (n_cand <= m) {                                   /* :170-172 */
        for (int64_t i = 0; i < n_cand; i++) out[i] = candidates[i];
        return n_cand;
    }
    orc_idd *d = (orc_idd *)malloc((size_t)n_cand * sizeof(orc_idd));
    orc_idd *acc = (orc_idd *)malloc((size_t)n_cand * sizeof(orc_idd)), *dis = (orc_idd *)malloc((size_t)n_cand * sizeof(orc_idd));
    for (int64_t i = 0; i < n_cand; i++) {               /* :175-181 */
        d[i].id = candidates[i];
        d[i].dist = orc_l2dist(vectors + u * dim, vectors + candidates[i] * dim, dim);
        d[i].pos = i;
    }
    qsort(d, (size_t)n_cand, sizeof(orc_idd), orc_idd_cmp);   /* :183-185 */
    int64_t na = 0, nd = 0;
    for (int64_t i = 0; i < n_cand; i++) {               /* :187-208 */
        const int64_t v = d[i].id;
        const float dist_uv = d[i].dist;
        int ok = 1;
        for (int64_t j = 0; j < na; j++) {
            const float dj = orc_l2dist(vectors + acc[j].id * dim, vectors + v * dim, dim);
            if (dj * alpha < dist_uv) { ok = 0; break; }
        }
        if (ok) {
            acc[na++] = d[i];
            if (na == m) break;
        } else {
            dis[nd++] = d[i];
        }
    }
    if (na < m)                                          /* :213-226 */
        for (int64_t i = 0; i < nd && na < m; i++) acc[na++] = dis[i];
    for (int64_t i = 0; i < na; i++) out[i] = acc[i].id;
    free(d); free(acc); free(dis);
    return na;
}
