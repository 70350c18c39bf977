/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of matter-labs/hodor's NTT / LDE / Merkle / FRI path.
 *
 * This file is the parity oracle for the CUDA kernels in hodor_b200/csrc and, through
 * bench.py's cpu_baseline / --impl reference legs, the timed "reference multicore CPU path".
 * Nothing under hodor_b200/ links, loads or calls it.
 *
 * PARITY STATUS: "parity unpinned".  The reference (@76fc894) is Rust; cargo/rustc are absent
 * from this image, so it cannot be compiled or run, and its tests contain no golden vectors for
 * this path.  What the reference does hold is pinned: the Montgomery encodings
 * MINUS_ONE / NON_RESIDUE (src/experiments/square_root_calculator/fp2.rs:10-22,
 * tests/test_oracle_pins.py) and E_PRECOMPUTED / F_PRECOMPUTED (fp2.rs:51-81), the printed
 * output of the reference's own `find_c` test: ~3000 dependent Montgomery mul / add / sub and
 * one inversion over `experiments::Fr`, reproduced bit for bit by fe_mul / fe_add / fe_sub /
 * fe_inv below (tests/test_reference_kat.py).  That pins row a1 (field arithmetic); the
 * transforms, the tree and the FRI chain built on it stay unpinned against the reference,
 * but are held to independent third-party code: sympy's ntt / intt for serial_fft, best_fft,
 * ifft and the (coset) LDE (tests/test_third_party_pins.py), hashlib for Blake2s.
 * The un-vendored dependencies whose published algorithms are
 * restated here:  ff_ce "0.7" (derive(PrimeField): 4 x u64 little-endian limbs, Montgomery
 * R = 2^256, canonical representatives, ROOT_OF_UNITY = GENERATOR^((p-1)/2^S));
 * blake2s_simd "0.5" (RFC 7693 Blake2s-256 with key and personalisation).
 *
 * The algorithms AND the thread decomposition follow the reference, deliberately without
 * improvement (running-product twiddles, the O(n*C) prologue of parallel_fft, a fresh tree
 * of threads per scope), because this code is also the timed CPU baseline.
 * All citations are relative to /root/reference.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;

typedef struct {
    fe p;
    uint64_t inv;      /* -p^-1 mod 2^64 */
    fe r;              /* R mod p = one */
    fe r2;             /* R^2 mod p */
    fe generator;      /* multiplicative generator, Montgomery */
    fe root_of_unity;  /* generator^((p-1)/2^S), Montgomery */
    uint32_t s;        /* 2-adicity */
    uint32_t num_bits;
    int ready;
} field_t;

enum { FIELD_BLS12_381_FR = 0, FIELD_BN254_FR = 1, FIELD_STARK252 = 2, NUM_FIELDS = 3 };

/* moduli: src/bn256.rs:5, (true BN254 Fr), src/experiments/mod.rs:19 */
static const uint64_t MODULI[NUM_FIELDS][4] = {
    {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL},
    {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    {0x0000000000000001ULL, 0x0000000000000000ULL, 0x0000000000000000ULL, 0x0800000000000011ULL},
};
static const uint64_t GENERATORS[NUM_FIELDS] = {7, 7, 3};  /* BN254: pairing_ce / ff_ce bn256::Fr declares 7 */

static field_t FIELDS[NUM_FIELDS];
static pthread_once_t fields_once = PTHREAD_ONCE_INIT;

/* ------------------------------------------------------------------ 256-bit helpers */
static inline int ge256(const fe *a, const fe *b) {
    for (int i = 3; i >= 0; i--) {
        if (a->l[i] > b->l[i]) return 1;
        if (a->l[i] < b->l[i]) return 0;
    }
    return 1;
}
static inline uint64_t add256(fe *r, const fe *a, const fe *b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a->l[i] + b->l[i]; r->l[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
static inline uint64_t sub256(fe *r, const fe *a, const fe *b) {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a->l[i] - b->l[i] - borrow;
        r->l[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}

/* ------------------------------------------------------------------ field ops (ff_ce derive) */
static inline void fe_add(const field_t *F, fe *r, const fe *a, const fe *b) {
    fe t; uint64_t c = add256(&t, a, b);
    if (c || ge256(&t, &F->p)) sub256(&t, &t, &F->p);
    *r = t;
}
static inline void fe_sub(const field_t *F, fe *r, const fe *a, const fe *b) {
    fe t;
    if (sub256(&t, a, b)) add256(&t, &t, &F->p);
    *r = t;
}
static inline void fe_neg(const field_t *F, fe *r, const fe *a) {
    fe z = {{0, 0, 0, 0}};
    fe_sub(F, r, &z, a);
}
/* Montgomery multiplication, CIOS over 64-bit limbs: r = a*b*R^-1 mod p, canonical */
static inline void fe_mul(const field_t *F, fe *r, const fe *a, const fe *b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c; c >>= 64;
        }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * F->inv;
        c = (u128)m * F->p.l[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * F->p.l[j] + t[j];
            t[j - 1] = (uint64_t)c; c >>= 64;
        }
        c += t[4]; t[3] = (uint64_t)c; c >>= 64;
        t[4] = t[5] + (uint64_t)c;
    }
    fe o = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || ge256(&o, &F->p)) sub256(&o, &o, &F->p);
    *r = o;
}
static void fe_pow_u64(const field_t *F, fe *r, const fe *base, uint64_t e) {
    fe acc = F->r, b = *base;
    while (e) {
        if (e & 1) fe_mul(F, &acc, &acc, &b);
        fe_mul(F, &b, &b, &b);
        e >>= 1;
    }
    *r = acc;
}
static void fe_pow_256(const field_t *F, fe *r, const fe *base, const fe *e) {
    fe acc = F->r;
    for (int i = 255; i >= 0; i--) {
        fe_mul(F, &acc, &acc, &acc);
        if ((e->l[i / 64] >> (i % 64)) & 1) fe_mul(F, &acc, &acc, base);
    }
    *r = acc;
}
static void fe_inv(const field_t *F, fe *r, const fe *a) {
    fe two = {{2, 0, 0, 0}}, e;
    sub256(&e, &F->p, &two);
    fe_pow_256(F, r, a, &e);
}
static inline int fe_eq(const fe *a, const fe *b) { return memcmp(a, b, sizeof(fe)) == 0; }

static void field_setup(field_t *F, const uint64_t mod[4], uint64_t gen) {
    memcpy(F->p.l, mod, 32);
    /* -p^-1 mod 2^64 by Newton iteration */
    uint64_t x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - mod[0] * x;
    F->inv = (uint64_t)0 - x;
    /* R mod p and R^2 mod p by repeated doubling of 1 (256 and 512 times) */
    fe v = {{1, 0, 0, 0}};
    for (int i = 0; i < 512; i++) {
        fe t; uint64_t c = add256(&t, &v, &v);
        if (c || ge256(&t, &F->p)) sub256(&t, &t, &F->p);
        v = t;
        if (i == 255) F->r = v;
    }
    F->r2 = v;
    F->num_bits = 0;
    for (int i = 255; i >= 0; i--) if ((mod[i / 64] >> (i % 64)) & 1) { F->num_bits = i + 1; break; }
    fe g = {{gen, 0, 0, 0}};
    fe_mul(F, &F->generator, &g, &F->r2);
    /* t = (p-1) >> S */
    fe t = F->p; t.l[0] -= 1;
    F->s = 0;
    while (!(t.l[0] & 1)) {
        for (int i = 0; i < 3; i++) t.l[i] = (t.l[i] >> 1) | (t.l[i + 1] << 63);
        t.l[3] >>= 1;
        F->s++;
    }
    fe_pow_256(F, &F->root_of_unity, &F->generator, &t);
    F->ready = 1;
}
static void fields_init(void) {
    for (int f = 0; f < NUM_FIELDS; f++) field_setup(&FIELDS[f], MODULI[f], GENERATORS[f]);
}
static const field_t *get_field(int id) {
    pthread_once(&fields_once, fields_init);
    if (id < 0 || id >= NUM_FIELDS) return NULL;
    return &FIELDS[id];
}

/* ------------------------------------------------------------------ Worker (src/fft/multicore.rs) */
static inline uint32_t log2_floor(uint64_t n) { uint32_t r = 0; while (n >>= 1) r++; return r; }
/* get_chunk_size :74-86 */
static inline size_t chunk_size(size_t elements, size_t cpus) { return elements < cpus ? 1 : elements / cpus; }

typedef void (*job_fn)(void *arg, size_t idx);
typedef struct { job_fn fn; void *arg; size_t idx; } job_t;
static void *job_tramp(void *p) { job_t *j = (job_t *)p; j->fn(j->arg, j->idx); return NULL; }
/* crossbeam::scope + spawn per chunk: one OS thread per job, joined at scope end (:60-72) */
static void run_jobs(size_t n, job_fn fn, void *arg) {
    if (n == 0) return;
    if (n == 1) { fn(arg, 0); return; }
    pthread_t *th = (pthread_t *)malloc(n * sizeof(pthread_t));
    job_t *jobs = (job_t *)malloc(n * sizeof(job_t));
    for (size_t i = 0; i < n; i++) {
        jobs[i].fn = fn; jobs[i].arg = arg; jobs[i].idx = i;
        if (pthread_create(&th[i], NULL, job_tramp, &jobs[i]) != 0) { th[i] = 0; fn(arg, i); }
    }
    for (size_t i = 0; i < n; i++) if (th[i]) pthread_join(th[i], NULL);
    free(th); free(jobs);
}

/* ------------------------------------------------------------------ NTT (src/fft/fft.rs) */
static inline uint32_t bitreverse32(uint32_t n, uint32_t l) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < l; i++) { r = (r << 1) | (n & 1); n >>= 1; }
    return r;
}
/* serial_fft :21-66 */
static void serial_fft(const field_t *F, fe *a, const fe *omega, uint32_t log_n) {
    uint32_t n = 1u << log_n;
    for (uint32_t k = 0; k < n; k++) {
        uint32_t rk = bitreverse32(k, log_n);
        if (k < rk) { fe t = a[rk]; a[rk] = a[k]; a[k] = t; }
    }
    uint32_t m = 1;
    for (uint32_t s = 0; s < log_n; s++) {
        fe w_m; fe_pow_u64(F, &w_m, omega, n / (2 * m));
        for (uint32_t k = 0; k < n; k += 2 * m) {
            fe w = F->r;
            for (uint32_t j = 0; j < m; j++) {
                fe t; fe_mul(F, &t, &a[k + j + m], &w);
                fe tmp; fe_sub(F, &tmp, &a[k + j], &t);
                a[k + j + m] = tmp;
                fe_add(F, &a[k + j], &a[k + j], &t);
                fe_mul(F, &w, &w, &w_m);
            }
        }
        m *= 2;
    }
}

typedef struct {
    const field_t *F; const fe *a; fe **tmp; const fe *omega; fe new_omega;
    uint32_t log_n, log_cpus, log_new_n; int radix4;
} pfft_t;
static void serial_fft_radix_4(const field_t *F, fe *a, const fe *omega, uint32_t log_n);
/* parallel_fft :86-108, the per-thread body */
static void pfft_job(void *p, size_t j) {
    pfft_t *c = (pfft_t *)p; const field_t *F = c->F;
    fe *tmp = c->tmp[j];
    fe omega_j, omega_step;
    fe_pow_u64(F, &omega_j, c->omega, j);
    fe_pow_u64(F, &omega_step, c->omega, (uint64_t)j << c->log_new_n);
    size_t num_cpus = (size_t)1 << c->log_cpus, new_n = (size_t)1 << c->log_new_n;
    size_t mask = ((size_t)1 << c->log_n) - 1;
    fe elt = F->r;
    for (size_t i = 0; i < new_n; i++) {
        for (size_t s = 0; s < num_cpus; s++) {
            size_t idx = (i + (s << c->log_new_n)) & mask;
            fe t; fe_mul(F, &t, &c->a[idx], &elt);
            fe_add(F, &tmp[i], &tmp[i], &t);
            fe_mul(F, &elt, &elt, &omega_step);
        }
        fe_mul(F, &elt, &elt, &omega_j);
    }
    if (c->radix4) serial_fft_radix_4(F, tmp, &c->new_omega, c->log_new_n);
    else serial_fft(F, tmp, &c->new_omega, c->log_new_n);
}
typedef struct { fe *a; fe **tmp; size_t chunk, n; uint32_t log_cpus; } gather_t;
/* parallel_fft :111-124 */
static void gather_job(void *p, size_t ci) {
    gather_t *g = (gather_t *)p;
    size_t idx = ci * g->chunk, end = idx + g->chunk; if (end > g->n) end = g->n;
    size_t mask = ((size_t)1 << g->log_cpus) - 1;
    for (; idx < end; idx++) g->a[idx] = g->tmp[idx & mask][idx >> g->log_cpus];
}
static void parallel_fft_any(const field_t *F, fe *a, size_t cpus, const fe *omega, uint32_t log_n,
                             uint32_t log_cpus, int radix4) {
    size_t num_cpus = (size_t)1 << log_cpus, n = (size_t)1 << log_n;
    pfft_t c; c.F = F; c.a = a; c.omega = omega; c.log_n = log_n; c.log_cpus = log_cpus;
    c.log_new_n = log_n - log_cpus; c.radix4 = radix4;
    c.tmp = (fe **)malloc(num_cpus * sizeof(fe *));
    for (size_t j = 0; j < num_cpus; j++) c.tmp[j] = (fe *)calloc((size_t)1 << c.log_new_n, sizeof(fe));
    fe_pow_u64(F, &c.new_omega, omega, num_cpus);
    run_jobs(num_cpus, pfft_job, &c);
    gather_t g; g.a = a; g.tmp = c.tmp; g.n = n; g.log_cpus = log_cpus; g.chunk = chunk_size(n, cpus);
    run_jobs((n + g.chunk - 1) / g.chunk, gather_job, &g);
    for (size_t j = 0; j < num_cpus; j++) free(c.tmp[j]);
    free(c.tmp);
}
/* best_fft :5-19.  hint < 0 means None */
static void best_fft(const field_t *F, fe *a, size_t cpus, const fe *omega, uint32_t log_n, long hint) {
    uint32_t log_cpus = hint >= 0 ? log2_floor((uint64_t)hint) : log2_floor(cpus);
    if (log_cpus == 0 || log_n <= log_cpus) serial_fft(F, a, omega, log_n);
    else parallel_fft_any(F, a, cpus, omega, log_n, log_cpus, 0);
}

/* ------------------------------------------------------------------ radix-4 (src/fft/radix4_fft/mod.rs) */
static inline uint64_t base4_digit_reverse(uint64_t n, uint64_t l) {
    uint64_t r = 0;
    for (uint64_t i = 0; i < l; i++) { r = (r << 2) | (n & 3); n >>= 2; }
    return r;
}
/* serial_fft_radix_4 :45-123 */
static void serial_fft_radix_4(const field_t *F, fe *a, const fe *omega, uint32_t log_n) {
    uint64_t n = (uint64_t)1 << log_n, num_digits = log_n / 2;
    for (uint64_t k = 0; k < n; k++) {
        uint64_t rk = base4_digit_reverse(k, num_digits);
        if (k < rk) { fe t = a[rk]; a[rk] = a[k]; a[k] = t; }
    }
    fe v; fe_pow_u64(F, &v, omega, n / 4);
    uint64_t m = 1;
    for (uint32_t s = 0; s < log_n / 2; s++) {
        fe w_m; fe_pow_u64(F, &w_m, omega, n / (4 * m));
        for (uint64_t k = 0; k < n; k += 4 * m) {
            fe w = F->r;
            for (uint64_t j = 0; j < m; j++) {
                fe u = w, x0 = a[k + j], x1, x2, x3;
                fe_mul(F, &x1, &a[k + j + m], &w);
                fe_mul(F, &u, &u, &w); fe_mul(F, &x2, &a[k + j + 2 * m], &u);
                fe_mul(F, &u, &u, &w); fe_mul(F, &x3, &a[k + j + 3 * m], &u);
                fe x0p2, x1p3, x0m2, x1m3;
                fe_add(F, &x0p2, &x0, &x2); fe_add(F, &x1p3, &x1, &x3);
                fe_add(F, &a[k + j], &x0p2, &x1p3);
                fe_sub(F, &a[k + j + 2 * m], &x0p2, &x1p3);
                fe_sub(F, &x0m2, &x0, &x2); fe_sub(F, &x1m3, &x1, &x3);
                fe_mul(F, &x1m3, &x1m3, &v);
                fe_add(F, &a[k + j + m], &x0m2, &x1m3);
                fe_sub(F, &a[k + j + 3 * m], &x0m2, &x1m3);
                fe_mul(F, &w, &w, &w_m);
            }
        }
        m *= 4;
    }
}
/* radix4 best_fft :5-20 */
static void best_fft_radix_4(const field_t *F, fe *a, size_t cpus, const fe *omega, uint32_t log_n) {
    uint32_t log_cpus = log2_floor(cpus);
    if (log_cpus % 2) log_cpus -= 1;
    if (log_n <= log_cpus) serial_fft_radix_4(F, a, omega, log_n);
    else parallel_fft_any(F, a, cpus, omega, log_n, log_cpus, 1);
}

/* ------------------------------------------------------------------ distribute_powers (src/fft/mod.rs:110-123) */
typedef struct { const field_t *F; fe *a; size_t n, chunk; fe g; } dp_t;
static void dp_job(void *p, size_t i) {
    dp_t *d = (dp_t *)p;
    size_t b = i * d->chunk, e = b + d->chunk; if (e > d->n) e = d->n;
    fe u; fe_pow_u64(d->F, &u, &d->g, (uint64_t)(i * d->chunk));
    for (size_t k = b; k < e; k++) { fe_mul(d->F, &d->a[k], &d->a[k], &u); fe_mul(d->F, &u, &u, &d->g); }
}
static void distribute_powers(const field_t *F, fe *a, size_t n, size_t cpus, const fe *g) {
    dp_t d; d.F = F; d.a = a; d.n = n; d.chunk = chunk_size(n, cpus); d.g = *g;
    run_jobs((n + d.chunk - 1) / d.chunk, dp_job, &d);
}
/* the `v *= minv` scope of Polynomial::ifft (src/polynomials/mod.rs:777-787) */
typedef struct { const field_t *F; fe *a; size_t n, chunk; fe c; } scale_t;
static void scale_job(void *p, size_t i) {
    scale_t *d = (scale_t *)p;
    size_t b = i * d->chunk, e = b + d->chunk; if (e > d->n) e = d->n;
    for (size_t k = b; k < e; k++) fe_mul(d->F, &d->a[k], &d->a[k], &d->c);
}

static void domain_generator(const field_t *F, uint32_t log_n, fe *out) {
    /* Domain::new_for_size src/domains/mod.rs:35-38 */
    fe g = F->root_of_unity;
    for (uint32_t i = log_n; i < F->s; i++) fe_mul(F, &g, &g, &g);
    *out = g;
}

/* ------------------------------------------------------------------ multi-coset LDE (src/polynomials/mod.rs:418-482, 544-609) */
typedef struct {
    const field_t *F; const fe *coeffs; fe **results; size_t n, factor, chunk, cpus;
    fe coset_omega, omega; uint32_t log_n; long hint; int coset;
} lde_t;
static void lde_job(void *p, size_t i) {
    lde_t *c = (lde_t *)p; const field_t *F = c->F;
    /* NOTE (reference quirk, kept on purpose): the start generator is coset_omega^i with i the
     * CHUNK index, :448 / :575, which is only the right coset when chunk == 1 (cpus >= factor). */
    fe gen; fe_pow_u64(F, &gen, &c->coset_omega, (uint64_t)i);
    if (c->coset) fe_mul(F, &gen, &gen, &F->generator);
    size_t b = i * c->chunk, e = b + c->chunk; if (e > c->factor) e = c->factor;
    for (size_t r = b; r < e; r++) {
        fe *v = (fe *)malloc(c->n * sizeof(fe));
        memcpy(v, c->coeffs, c->n * sizeof(fe));
        distribute_powers(F, v, c->n, c->cpus, &gen);
        best_fft(F, v, c->cpus, &c->omega, c->log_n, c->hint);
        c->results[r] = v;
        fe_mul(F, &gen, &gen, &c->coset_omega);
    }
}
typedef struct { fe *out; fe **results; size_t total, chunk, factor; } il_t;
static void il_job(void *p, size_t i) {
    il_t *c = (il_t *)p;
    size_t idx = i * c->chunk, e = idx + c->chunk; if (e > c->total) e = c->total;
    for (; idx < e; idx++) c->out[idx] = c->results[idx % c->factor][idx / c->factor];
}

/* ------------------------------------------------------------------ Blake2s-256, RFC 7693, keyed + personalised */
static const uint32_t B2S_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                   0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t B2S_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
static inline uint32_t rotr32(uint32_t x, int r) { return (x >> r) | (x << (32 - r)); }
static void b2s_compress(uint32_t h[8], const uint8_t block[64], uint64_t t, int last) {
    uint32_t m[16], v[16];
    for (int i = 0; i < 16; i++) {
        m[i] = (uint32_t)block[4 * i] | ((uint32_t)block[4 * i + 1] << 8) | ((uint32_t)block[4 * i + 2] << 16) |
               ((uint32_t)block[4 * i + 3] << 24);
    }
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = B2S_IV[i]; }
    v[12] ^= (uint32_t)t; v[13] ^= (uint32_t)(t >> 32);
    if (last) v[14] = ~v[14];
#define G(a, b, c, d, x, y)                                                  \
    v[a] = v[a] + v[b] + (x); v[d] = rotr32(v[d] ^ v[a], 16);                \
    v[c] = v[c] + v[d];       v[b] = rotr32(v[b] ^ v[c], 12);                \
    v[a] = v[a] + v[b] + (y); v[d] = rotr32(v[d] ^ v[a], 8);                 \
    v[c] = v[c] + v[d];       v[b] = rotr32(v[b] ^ v[c], 7);
    for (int r = 0; r < 10; r++) {
        const uint8_t *s = B2S_SIGMA[r];
        G(0, 4, 8, 12, m[s[0]], m[s[1]]) G(1, 5, 9, 13, m[s[2]], m[s[3]])
        G(2, 6, 10, 14, m[s[4]], m[s[5]]) G(3, 7, 11, 15, m[s[6]], m[s[7]])
        G(0, 5, 10, 15, m[s[8]], m[s[9]]) G(1, 6, 11, 12, m[s[10]], m[s[11]])
        G(2, 7, 8, 13, m[s[12]], m[s[13]]) G(3, 4, 9, 14, m[s[14]], m[s[15]])
    }
#undef G
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}
static const char B2S_KEY[] = "Squeamish Ossifrage"; /* src/iop/blake2s_trivial_iop.rs:12 */
static const char B2S_PERSONAL[] = "Shaftoe";         /* :13 */
/* Params::new().hash_length(32).key(..).personal(..).to_state(); update(data); finalize().
 * As in blake2s_simd (and on the reference CPU), this is TWO compressions per call: the key
 * block, then the data block.  len is 32 (leaf) or 64 (node). */
static void b2s_hash(uint8_t out[32], const uint8_t *data, size_t len) {
    uint32_t h[8];
    uint8_t block[64];
    size_t keylen = sizeof(B2S_KEY) - 1;
    for (int i = 0; i < 8; i++) h[i] = B2S_IV[i];
    h[0] ^= 0x01010000u ^ ((uint32_t)keylen << 8) ^ 32u; /* digest 32, key len, fanout 1, depth 1 */
    uint8_t pers[8] = {0};
    memcpy(pers, B2S_PERSONAL, sizeof(B2S_PERSONAL) - 1);
    h[6] ^= (uint32_t)pers[0] | ((uint32_t)pers[1] << 8) | ((uint32_t)pers[2] << 16) | ((uint32_t)pers[3] << 24);
    h[7] ^= (uint32_t)pers[4] | ((uint32_t)pers[5] << 8) | ((uint32_t)pers[6] << 16) | ((uint32_t)pers[7] << 24);
    memset(block, 0, 64); memcpy(block, B2S_KEY, keylen);
    if (len == 0) { b2s_compress(h, block, 64, 1); }
    else {
        b2s_compress(h, block, 64, 0);
        uint64_t t = 64;
        while (len > 64) { b2s_compress(h, data, t += 64, 0); data += 64; len -= 64; }
        memset(block, 0, 64); memcpy(block, data, len);
        b2s_compress(h, block, t + len, 1);
    }
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)h[i]; out[4 * i + 1] = (uint8_t)(h[i] >> 8);
        out[4 * i + 2] = (uint8_t)(h[i] >> 16); out[4 * i + 3] = (uint8_t)(h[i] >> 24);
    }
}
/* encode_leaf :36-42 -- raw Montgomery limbs little-endian.  (x86-64 is little-endian, so the
 * in-memory fe already is that byte string.) */
static inline void hash_leaf(uint8_t out[32], const fe *x) { b2s_hash(out, (const uint8_t *)x, 32); }

/* Blake2sIopTree::create :131-219 */
typedef struct { const fe *leaves; uint8_t *lh; size_t n, chunk; } lh_t;
static void lh_job(void *p, size_t i) {
    lh_t *c = (lh_t *)p;
    size_t b = i * c->chunk, e = b + c->chunk; if (e > c->n) e = c->n;
    for (size_t k = b; k < e; k++) hash_leaf(c->lh + 32 * k, &c->leaves[k]);
}
typedef struct { const uint8_t *in; uint8_t *out; size_t n_out, chunk; } nl_t;
static void nl_job(void *p, size_t i) {
    nl_t *c = (nl_t *)p;
    size_t b = i * c->chunk, e = b + c->chunk; if (e > c->n_out) e = c->n_out;
    for (size_t k = b; k < e; k++) b2s_hash(c->out + 32 * k, c->in + 64 * k, 64);
}
static void merkle_create(const fe *leaves, size_t n, uint8_t *nodes, size_t cpus) {
    memset(nodes, 0, 32 * n);
    uint8_t *lh = (uint8_t *)malloc(32 * n);
    lh_t a; a.leaves = leaves; a.lh = lh; a.n = n; a.chunk = chunk_size(n, cpus);
    run_jobs((n + a.chunk - 1) / a.chunk, lh_job, &a);
    nl_t b; b.in = lh; b.out = nodes + 32 * (n / 2); b.n_out = n / 2; b.chunk = chunk_size(n / 2, cpus);
    run_jobs((b.n_out + b.chunk - 1) / b.chunk, nl_job, &b);
    for (size_t width = n / 4; width >= 1; width /= 2) {
        nl_t c; c.in = nodes + 32 * (2 * width); c.out = nodes + 32 * width; c.n_out = width;
        c.chunk = chunk_size(width, cpus);
        run_jobs((width + c.chunk - 1) / c.chunk, nl_job, &c);
    }
    free(lh);
}
/* interpret_hash :48-60 -> Montgomery form (from_repr) */
static int interpret_hash(const field_t *F, const uint8_t d[32], fe *out) {
    fe v;
    for (int k = 0; k < 4; k++) {
        uint64_t x = 0;
        for (int b = 0; b < 8; b++) x = (x << 8) | d[8 * k + b];
        v.l[3 - k] = x;
    }
    uint32_t shave = (256 - (F->num_bits - 1)) % 64;
    v.l[3] &= 0xffffffffffffffffULL >> shave;
    if (ge256(&v, &F->p)) return -1;
    fe_mul(F, out, &v, &F->r2);
    return 0;
}

/* ------------------------------------------------------------------ FRI (src/fri/fri_on_values.rs:11-159) */
typedef struct { const field_t *F; fe *v; size_t n, chunk; fe base; } pw_t;
static void pw_job(void *p, size_t i) {
    pw_t *c = (pw_t *)p;
    size_t b = i * c->chunk, e = b + c->chunk; if (e > c->n) e = c->n;
    fe u; fe_pow_u64(c->F, &u, &c->base, (uint64_t)(i * c->chunk));
    for (size_t k = b; k < e; k++) { c->v[k] = u; fe_mul(c->F, &u, &u, &c->base); }
}
typedef struct {
    const field_t *F; const fe *values, *omegas_inv; fe *next; size_t next_size, chunk, stride; fe challenge, two_inv;
} fold_t;
static void fold_job(void *p, size_t i) {
    fold_t *c = (fold_t *)p; const field_t *F = c->F;
    size_t b = i * c->chunk, e = b + c->chunk; if (e > c->next_size) e = c->next_size;
    for (size_t idx = b; idx < e; idx++) {
        const fe *f0 = &c->values[idx], *f1 = &c->values[idx + c->next_size];
        fe even, odd, tmp;
        fe_add(F, &even, f0, f1);
        fe_sub(F, &odd, f0, f1);
        fe_mul(F, &odd, &odd, &c->omegas_inv[idx * c->stride]);
        fe_mul(F, &tmp, &odd, &c->challenge);
        fe_add(F, &tmp, &tmp, &even);
        fe_mul(F, &c->next[idx], &tmp, &c->two_inv);
    }
}

/* ================================================================== exported C API (ctypes) */
#define API __attribute__((visibility("default")))

API int oracle_field_constants(int field, uint64_t *p, uint64_t *r, uint64_t *r2, uint64_t *inv,
                               uint64_t *gen, uint64_t *root, uint32_t *s, uint32_t *num_bits) {
    const field_t *F = get_field(field); if (!F) return -1;
    memcpy(p, &F->p, 32); memcpy(r, &F->r, 32); memcpy(r2, &F->r2, 32); *inv = F->inv;
    memcpy(gen, &F->generator, 32); memcpy(root, &F->root_of_unity, 32); *s = F->s; *num_bits = F->num_bits;
    return 0;
}
API int oracle_mul(int field, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t count) {
    const field_t *F = get_field(field); if (!F) return -1;
    for (size_t i = 0; i < count; i++) fe_mul(F, (fe *)out + i, (const fe *)a + i, (const fe *)b + i);
    return 0;
}
API int oracle_add(int field, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t count) {
    const field_t *F = get_field(field); if (!F) return -1;
    for (size_t i = 0; i < count; i++) fe_add(F, (fe *)out + i, (const fe *)a + i, (const fe *)b + i);
    return 0;
}
API int oracle_sub(int field, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t count) {
    const field_t *F = get_field(field); if (!F) return -1;
    for (size_t i = 0; i < count; i++) fe_sub(F, (fe *)out + i, (const fe *)a + i, (const fe *)b + i);
    return 0;
}
API int oracle_inverse(int field, const uint64_t *a, uint64_t *out) {
    const field_t *F = get_field(field); if (!F) return -1;
    fe_inv(F, (fe *)out, (const fe *)a); return 0;
}
API int oracle_pow(int field, const uint64_t *a, uint64_t e, uint64_t *out) {
    const field_t *F = get_field(field); if (!F) return -1;
    fe_pow_u64(F, (fe *)out, (const fe *)a, e); return 0;
}
API int oracle_to_mont(int field, const uint64_t *plain, uint64_t *out, size_t count) {
    const field_t *F = get_field(field); if (!F) return -1;
    for (size_t i = 0; i < count; i++) fe_mul(F, (fe *)out + i, (const fe *)plain + i, &F->r2);
    return 0;
}
API int oracle_from_mont(int field, const uint64_t *mont, uint64_t *out, size_t count) {
    const field_t *F = get_field(field); if (!F) return -1;
    fe one = {{1, 0, 0, 0}};
    for (size_t i = 0; i < count; i++) fe_mul(F, (fe *)out + i, (const fe *)mont + i, &one);
    return 0;
}
API int oracle_domain_generator(int field, uint32_t log_n, uint64_t *out) {
    const field_t *F = get_field(field); if (!F) return -1;
    if (log_n > F->s) return -2; /* Domain::new_for_size -> Err(SynthesisError::Error) */
    domain_generator(F, log_n, (fe *)out); return 0;
}
API int oracle_serial_fft(int field, uint64_t *a, const uint64_t *omega, uint32_t log_n) {
    const field_t *F = get_field(field); if (!F) return -1;
    serial_fft(F, (fe *)a, (const fe *)omega, log_n); return 0;
}
API int oracle_serial_fft_radix_4(int field, uint64_t *a, const uint64_t *omega, uint32_t log_n) {
    const field_t *F = get_field(field); if (!F) return -1;
    if (log_n % 2) return -2; /* assert!(log_n % 2 == 0) */
    serial_fft_radix_4(F, (fe *)a, (const fe *)omega, log_n); return 0;
}
/* serial_DIT_fft (src/fft/dit_fft/mod.rs:4-53): decimation-in-frequency butterflies with a running
 * twiddle per block, then the bit-reversal swap; `non_zero_entries_count` prunes the butterflies of a
 * zero tail (only the first min(block_len/2, nz) of every block are executed).  Same output as
 * serial_fft -- the reference asserts it (src/fft/mod.rs:66-126, 187-279). */
static void serial_dif_fft(const field_t *F, fe *a, const fe *omega, uint32_t log_n, uint64_t nz) {
    const uint64_t n = (uint64_t)1 << log_n;
    uint64_t m = 1;
    for (uint32_t s = 0; s < log_n; s++) {
        fe w_m; fe_pow_u64(F, &w_m, omega, m);
        const uint64_t block_len = n / m, lim = block_len / 2 < nz ? block_len / 2 : nz;
        for (uint64_t block = 0; block < m; block++) {
            fe w = F->r;
            for (uint64_t k = block * block_len; k < block * block_len + lim; k++) {
                fe t = a[k + block_len / 2], tmp;
                fe_sub(F, &tmp, &a[k], &t);
                fe_mul(F, &a[k + block_len / 2], &tmp, &w);
                fe_add(F, &a[k], &a[k], &t);
                fe_mul(F, &w, &w, &w_m);
            }
        }
        m *= 2;
    }
    for (uint64_t k = 0; k < n; k++) {
        uint64_t rk = bitreverse32((uint32_t)k, log_n);
        if (k < rk) { fe t = a[rk]; a[rk] = a[k]; a[k] = t; }
    }
}
API int oracle_serial_dif_fft(int field, uint64_t *a, const uint64_t *omega, uint32_t log_n, uint64_t non_zero_entries) {
    const field_t *F = get_field(field); if (!F || log_n > 31) return -1;
    serial_dif_fft(F, (fe *)a, (const fe *)omega, log_n, non_zero_entries); return 0;
}
/* serial_lde (src/fft/lde.rs:15-126): radix-2 DIT NTT of a vector that was zero-padded by `lde_factor`;
 * after the bit reversal index idx holds a non-zero value at stage `step` iff (idx mod lde_factor) <
 * 2^step (is_non_zero, :27-31), and the butterflies with a structurally zero input are short-cut by
 * the four-way match at :90-116 until the round is dense (:33-41). */
static void serial_lde(const field_t *F, fe *a, const fe *omega, uint32_t log_n, size_t lde_factor) {
    const uint32_t n = (uint32_t)1 << log_n;
    for (uint32_t k = 0; k < n; k++) {
        uint32_t rk = bitreverse32(k, log_n);
        if (k < rk) { fe t = a[rk]; a[rk] = a[k]; a[k] = t; }
    }
    uint32_t m = 1, step = 0;
    for (uint32_t s = 0; s < log_n; s++) {
        fe w_m; fe_pow_u64(F, &w_m, omega, (uint64_t)(n / (2 * m)));
        const int dense = (lde_factor >> step) <= 1;
        for (uint32_t k = 0; k < n; k += 2 * m) {
            fe w = F->r;
            for (uint32_t j = 0; j < m; j++) {
                const size_t odd = k + j + m, even = k + j;
                const int odd_nz = dense || (odd & (lde_factor - 1)) < ((size_t)1 << step);
                const int even_nz = dense || (even & (lde_factor - 1)) < ((size_t)1 << step);
                if (odd_nz && even_nz) {
                    fe t, tmp; fe_mul(F, &t, &a[odd], &w);
                    fe_sub(F, &tmp, &a[even], &t); a[odd] = tmp; fe_add(F, &a[even], &a[even], &t);
                } else if (!odd_nz && even_nz) {
                    a[odd] = a[even];
                } else if (odd_nz && !even_nz) {
                    fe t, tmp; fe_mul(F, &t, &a[odd], &w); fe_neg(F, &tmp, &t); a[odd] = tmp; a[even] = t;
                }
                fe_mul(F, &w, &w, &w_m);
            }
        }
        step++; m *= 2;
    }
}
/* Polynomial::filtering_lde / coset_filtering_lde (src/polynomials/mod.rs:355-368, 484-499), serial path:
 * (distribute_powers by the multiplicative generator,) zero-pad to n * factor, serial_lde. */
API int oracle_filtering_lde(int field, const uint64_t *coeffs, uint32_t log_n, uint32_t factor, int coset, uint64_t *out,
                             uint32_t cpus) {
    const field_t *F = get_field(field); if (!F || cpus == 0 || factor == 0 || (factor & (factor - 1))) return -1;
    uint32_t log_f = log2_floor(factor);
    if (log_n + log_f > F->s || log_n + log_f > 31) return -2;
    const size_t n = (size_t)1 << log_n, total = n * factor;
    memset(out, 0, total * sizeof(fe));
    memcpy(out, coeffs, n * sizeof(fe));
    if (coset) distribute_powers(F, (fe *)out, n, cpus, &F->generator);
    if (factor == 1) { fe om; domain_generator(F, log_n, &om); serial_fft(F, (fe *)out, &om, log_n); return 0; }
    fe omega; domain_generator(F, log_n + log_f, &omega);
    serial_lde(F, (fe *)out, &omega, log_n + log_f, factor);
    return 0;
}
/* best_fft(a, worker{cpus}, omega, log_n, hint)  hint < 0 == None */
API int oracle_best_fft(int field, uint64_t *a, const uint64_t *omega, uint32_t log_n, uint32_t cpus, long hint) {
    const field_t *F = get_field(field); if (!F || cpus == 0) return -1;
    best_fft(F, (fe *)a, cpus, (const fe *)omega, log_n, hint); return 0;
}
API int oracle_best_fft_radix_4(int field, uint64_t *a, const uint64_t *omega, uint32_t log_n, uint32_t cpus) {
    const field_t *F = get_field(field); if (!F || cpus == 0) return -1;
    if (log_n % 2) return -2;
    best_fft_radix_4(F, (fe *)a, cpus, (const fe *)omega, log_n); return 0;
}
API int oracle_distribute_powers(int field, uint64_t *a, size_t n, const uint64_t *g, uint32_t cpus) {
    const field_t *F = get_field(field); if (!F || cpus == 0) return -1;
    distribute_powers(F, (fe *)a, n, cpus, (const fe *)g); return 0;
}
/* Polynomial::ifft / icoset_fft (src/polynomials/mod.rs:773-807) */
API int oracle_ifft(int field, uint64_t *a, uint32_t log_n, uint32_t cpus, int coset) {
    const field_t *F = get_field(field); if (!F || cpus == 0) return -1;
    if (log_n > F->s) return -2;
    size_t n = (size_t)1 << log_n;
    fe omega, omega_inv, nn = {{n, 0, 0, 0}}, minv;
    domain_generator(F, log_n, &omega); fe_inv(F, &omega_inv, &omega);
    fe_mul(F, &nn, &nn, &F->r2); fe_inv(F, &minv, &nn);
    best_fft(F, (fe *)a, cpus, &omega_inv, log_n, -1);
    scale_t s; s.F = F; s.a = (fe *)a; s.n = n; s.chunk = chunk_size(n, cpus); s.c = minv;
    run_jobs((n + s.chunk - 1) / s.chunk, scale_job, &s);
    if (coset) { fe ginv; fe_inv(F, &ginv, &F->generator); distribute_powers(F, (fe *)a, n, cpus, &ginv); }
    return 0;
}
/* Polynomial::fft / coset_fft (:611-631) */
API int oracle_fft(int field, uint64_t *a, uint32_t log_n, uint32_t cpus, int coset) {
    const field_t *F = get_field(field); if (!F || cpus == 0) return -1;
    if (log_n > F->s) return -2;
    fe omega; domain_generator(F, log_n, &omega);
    if (coset) distribute_powers(F, (fe *)a, (size_t)1 << log_n, cpus, &F->generator);
    best_fft(F, (fe *)a, cpus, &omega, log_n, -1);
    return 0;
}
/* (coset_)lde_using_multiple_cosets */
API int oracle_lde(int field, const uint64_t *coeffs, uint32_t log_n, uint32_t factor, int coset, uint64_t *out,
                   uint32_t cpus) {
    const field_t *F = get_field(field); if (!F || cpus == 0) return -1;
    if (factor == 0 || (factor & (factor - 1))) return -3;
    uint32_t log_f = log2_floor(factor);
    if (log_n + log_f > F->s) return -2;
    size_t n = (size_t)1 << log_n;
    if (factor == 1) { memcpy(out, coeffs, 32 * n); return oracle_fft(field, out, log_n, cpus, coset); }
    lde_t c; c.F = F; c.coeffs = (const fe *)coeffs; c.n = n; c.factor = factor; c.cpus = cpus;
    c.log_n = log_n; c.coset = coset;
    /* num_cpus_hint :423-432 */
    if (cpus <= factor) c.hint = 1;
    else { long t = factor / cpus; if (factor % cpus) t += 1; c.hint = t; }
    domain_generator(F, log_n + log_f, &c.coset_omega); domain_generator(F, log_n, &c.omega);
    c.results = (fe **)calloc(factor, sizeof(fe *));
    c.chunk = chunk_size(factor, cpus);
    run_jobs((factor + c.chunk - 1) / c.chunk, lde_job, &c);
    il_t il; il.out = (fe *)out; il.results = c.results; il.total = n * factor; il.factor = factor;
    il.chunk = chunk_size(il.total, cpus);
    run_jobs((il.total + il.chunk - 1) / il.chunk, il_job, &il);
    for (size_t i = 0; i < factor; i++) free(c.results[i]);
    free(c.results);
    return 0;
}
API int oracle_hash_leaf(int field, const uint64_t *x, uint8_t *out) { (void)field; hash_leaf(out, (const fe *)x); return 0; }
API int oracle_hash_node(const uint8_t *l, const uint8_t *r, uint8_t *out) {
    uint8_t buf[64]; memcpy(buf, l, 32); memcpy(buf + 32, r, 32); b2s_hash(out, buf, 64); return 0;
}
API int oracle_blake2s(const uint8_t *data, size_t len, uint8_t *out) { b2s_hash(out, data, len); return 0; }
API int oracle_merkle_create(int field, const uint64_t *leaves, size_t n, uint8_t *nodes, uint32_t cpus) {
    const field_t *F = get_field(field); if (!F || cpus == 0) return -1;
    if (n < 2 || (n & (n - 1))) return -3;
    merkle_create((const fe *)leaves, n, nodes, cpus); return 0;
}
API int oracle_interpret_hash(int field, const uint8_t *digest, uint64_t *out) {
    const field_t *F = get_field(field); if (!F) return -1;
    return interpret_hash(F, digest, (fe *)out);
}
/* proof_from_lde_by_values.  Caller provides: l0_nodes (n*32), layer_nodes[i] ((n>>(i+1))*32),
 * layer_values[i] ((n>>(i+1))*4 u64), challenges (steps*4), final_root (32), final_coeffs (out_coeffs*4).
 * Returns num_steps or a negative error. */
API int oracle_fri_commit(int field, const uint64_t *lde, size_t n, uint32_t lde_factor, uint32_t out_coeffs,
                          uint8_t *l0_nodes, uint8_t **layer_nodes, uint64_t **layer_values, uint64_t *challenges,
                          uint8_t *final_root, uint64_t *final_coeffs, uint32_t cpus) {
    const field_t *F = get_field(field); if (!F || cpus == 0) return -1;
    if (n < 2 || (n & (n - 1)) || !lde_factor || (lde_factor & (lde_factor - 1)) || !out_coeffs ||
        (out_coeffs & (out_coeffs - 1))) return -3;
    uint32_t log_n = log2_floor(n);
    if (log_n > F->s) return -2;
    size_t initial_degree_plus_one = n / lde_factor;
    if (initial_degree_plus_one / out_coeffs == 0) return -3;
    int num_steps = (int)log2_floor(initial_degree_plus_one / out_coeffs);
    if (num_steps < 1) return -4; /* reference panics: roots.pop() on empty */
    merkle_create((const fe *)lde, n, l0_nodes, cpus);
    fe two, two_inv, omega, omega_inv;
    fe_add(F, &two, &F->r, &F->r); fe_inv(F, &two_inv, &two);
    domain_generator(F, log_n, &omega); fe_inv(F, &omega_inv, &omega);
    size_t pre = n / 2;
    fe *omegas_inv = (fe *)malloc(pre * sizeof(fe));
    pw_t pw; pw.F = F; pw.v = omegas_inv; pw.n = pre; pw.chunk = chunk_size(pre, cpus); pw.base = omega_inv;
    run_jobs((pre + pw.chunk - 1) / pw.chunk, pw_job, &pw);
    fe challenge;
    if (interpret_hash(F, l0_nodes + 32, &challenge)) { free(omegas_inv); return -5; }
    const fe *values = (const fe *)lde;
    size_t next_size = n / 2;
    for (int i = 0; i < num_steps; i++) {
        memcpy(challenges + 4 * i, &challenge, 32);
        fold_t f; f.F = F; f.values = values; f.omegas_inv = omegas_inv; f.next = (fe *)layer_values[i];
        f.next_size = next_size; f.chunk = chunk_size(next_size, cpus); f.stride = (size_t)1 << i;
        f.challenge = challenge; f.two_inv = two_inv;
        run_jobs((next_size + f.chunk - 1) / f.chunk, fold_job, &f);
        merkle_create((const fe *)layer_values[i], next_size, layer_nodes[i], cpus);
        if (interpret_hash(F, layer_nodes[i] + 32, &challenge)) { free(omegas_inv); return -5; }
        values = (const fe *)layer_values[i];
        next_size >>= 1;
    }
    memcpy(final_root, layer_nodes[num_steps - 1] + 32, 32);
    size_t last = n >> num_steps;
    fe *fin = (fe *)malloc(last * sizeof(fe));
    memcpy(fin, values, last * sizeof(fe));
    oracle_ifft(field, (uint64_t *)fin, log2_floor(last), cpus, 0);
    memcpy(final_coeffs, fin, (size_t)out_coeffs * sizeof(fe));
    free(fin); free(omegas_inv);
    return num_steps;
}

/* ------------------------------------------------------------------ Polynomial::evaluate_at (src/polynomials/mod.rs:685-711) */
/* Worker::get_num_spawned_threads (src/fft/multicore.rs:87-102) */
static size_t num_spawned_threads(size_t elements, size_t cpus) {
    if (elements < cpus) return elements;
    size_t chunk = chunk_size(elements, cpus), n = elements / chunk;
    if (elements % chunk != 0) n++;
    return n;
}
typedef struct { const field_t *F; const fe *a; size_t n, chunk; fe g; fe *sub; } ev_t;
static void ev_job(void *p, size_t i) {
    ev_t *d = (ev_t *)p; const field_t *F = d->F;
    size_t b = i * d->chunk, e = b + d->chunk; if (e > d->n) e = d->n;
    fe x; fe_pow_u64(F, &x, &d->g, (uint64_t)(i * d->chunk));          /* :695 */
    fe s; memset(&s, 0, sizeof s);
    for (size_t k = b; k < e; k++) {                                    /* :696-701 */
        fe v; fe_mul(F, &v, &x, &d->a[k]); fe_add(F, &s, &s, &v); fe_mul(F, &x, &x, &d->g);
    }
    d->sub[i] = s;
}
API int oracle_evaluate_at(int field, const uint64_t *coeffs, size_t n, const uint64_t *g, uint32_t cpus, uint64_t *out) {
    const field_t *F = get_field(field); if (!F || cpus == 0) return -1;
    fe res; memset(&res, 0, sizeof res);
    if (n) {
        size_t threads = num_spawned_threads(n, cpus);
        ev_t d; d.F = F; d.a = (const fe *)coeffs; d.n = n; d.chunk = chunk_size(n, cpus); d.g = *(const fe *)g;
        d.sub = (fe *)calloc(threads, sizeof(fe));
        run_jobs(threads, ev_job, &d);
        for (size_t i = 0; i < threads; i++) fe_add(F, &res, &res, &d.sub[i]);   /* :706-709 */
        free(d.sub);
    }
    memcpy(out, &res, 32);
    return 0;
}

/* ------------------------------------------------------------------ Polynomial::batch_inversion (src/polynomials/mod.rs:889-954) */
typedef struct { const field_t *F; fe *a; fe *grand; fe *sub; size_t n, chunk; } bi_t;
static void bi_up_job(void *p, size_t i) {                              /* :897-908 */
    bi_t *d = (bi_t *)p; const field_t *F = d->F;
    size_t b = i * d->chunk, e = b + d->chunk; if (e > d->n) e = d->n;
    fe s = F->r;
    for (size_t k = b; k < e; k++) { fe_mul(F, &s, &s, &d->a[k]); d->grand[k] = s; }
    d->sub[i] = s;
}
static void bi_down_job(void *p, size_t i) {                            /* :934-951; sub[] now holds the sub-inverses */
    bi_t *d = (bi_t *)p; const field_t *F = d->F;
    size_t b = i * d->chunk, e = b + d->chunk; if (e > d->n) e = d->n;
    fe s = d->sub[i];
    for (size_t k = e; k-- > b;) {
        fe tmp = d->a[k];
        fe g = (k > b) ? d->grand[k - 1] : F->r;
        fe_mul(F, &d->a[k], &g, &s);
        fe_mul(F, &s, &s, &tmp);
    }
}
/* returns 0, or -2 (vector untouched) when an element is zero: Err(SynthesisError::Error), :919 */
API int oracle_batch_inversion(int field, uint64_t *a, size_t n, uint32_t cpus) {
    const field_t *F = get_field(field); if (!F || cpus == 0) return -1;
    if (n == 0) return 0;
    size_t threads = num_spawned_threads(n, cpus);
    bi_t d; d.F = F; d.a = (fe *)a; d.n = n; d.chunk = chunk_size(n, cpus);
    d.grand = (fe *)malloc(n * sizeof(fe)); d.sub = (fe *)malloc(threads * sizeof(fe));
    run_jobs(threads, bi_up_job, &d);
    fe full = F->r, zero; memset(&zero, 0, sizeof zero);
    for (size_t i = 0; i < threads; i++) fe_mul(F, &full, &full, &d.sub[i]);   /* :914-917 */
    if (fe_eq(&full, &zero)) { free(d.grand); free(d.sub); return -2; }
    fe pinv; fe_inv(F, &pinv, &full);
    fe *subinv = (fe *)malloc(threads * sizeof(fe));
    for (size_t i = 0; i < threads; i++) {                                      /* :922-932 */
        fe t = pinv;
        for (size_t j = 0; j < threads; j++) if (j != i) fe_mul(F, &t, &t, &d.sub[j]);
        subinv[i] = t;
    }
    memcpy(d.sub, subinv, threads * sizeof(fe)); free(subinv);
    run_jobs(threads, bi_down_job, &d);
    free(d.grand); free(d.sub);
    return 0;
}

/* SplitMix64 test-vector generator: limbs used directly as Montgomery form. */
API int oracle_random_elements(int field, uint64_t *out, size_t count, uint64_t seed) {
    const field_t *F = get_field(field); if (!F) return -1;
    uint64_t state = seed, mask = 0xffffffffffffffffULL >> (256 - F->num_bits);
    size_t k = 0;
    while (k < count) {
        fe v;
        for (int i = 0; i < 4; i++) {
            state += 0x9E3779B97F4A7C15ULL;
            uint64_t z = state;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
            v.l[i] = z ^ (z >> 31);
        }
        v.l[3] &= mask;
        if (!ge256(&v, &F->p)) memcpy(out + 4 * k++, &v, 32);
    }
    return 0;
}
