/*
 * ckks_oracle.c -- CPU restatement of the RNS-CKKS primitives behind FHEController (TEST
 * INFRASTRUCTURE ONLY; see ckks_oracle.h for the scope note and the "parity unpinned" statement).
 *
 * Each function cites the reference call site it stands behind (F.cpp = /root/reference/src/
 * FHEController.cpp) and the OpenFHE convention from SURVEY.md Appendix A (A.n) it restates.
 * OpenFHE itself (un-vendored, unpinned dependency, CMakeLists.txt:13) is not available; the
 * algorithms are the published ones: Cheon-Han-Kim-Kim-Song full-RNS CKKS, Han-Ki hybrid key
 * switching, Harvey/Shoup NTT butterflies.
 *
 * Build: see oracle/Makefile (gcc -O3 -march=x86-64-v3 -fopenmp -ffp-contract=off).
 */
#include "ckks_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef unsigned __int128 u128;
typedef __int128 i128;

#define MAXLIMBS 64

struct orc_ctx {
    int logN, N, L, K, dnum, alpha;
    u64 q[MAXLIMBS];       /* moduli: Q then P */
    u64 mu[MAXLIMBS];      /* Barrett: floor(2^(2n)/q), n = bitlen(q) */
    int nbits[MAXLIMBS];
    u64 psi[MAXLIMBS], psi_inv[MAXLIMBS], n_inv[MAXLIMBS], n_inv_sh[MAXLIMBS];
    u64 *tw[MAXLIMBS], *tw_sh[MAXLIMBS];     /* psi^bitrev(i), Shoup companions */
    u64 *itw[MAXLIMBS], *itw_sh[MAXLIMBS];   /* psi^-bitrev(i) */
    uint32_t* brev;        /* bit reversal of logN bits */
    double sf[MAXLIMBS];   /* scaling factor per level (A.8) */
};

/* ---------- modular arithmetic ---------- */
static inline u64 mulmod_slow(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }
static u64 powmod(u64 a, u64 e, u64 q) {
    u64 r = 1; a %= q;
    while (e) { if (e & 1) r = mulmod_slow(r, a, q); a = mulmod_slow(a, a, q); e >>= 1; }
    return r;
}
static u64 invmod(u64 a, u64 q) { return powmod(a, q - 2, q); } /* q prime */
static inline u64 addmod(u64 a, u64 b, u64 q) { u64 s = a + b; return s >= q ? s - q : s; }
static inline u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
static inline u64 shoup(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }
/* a*w mod q with w' = floor(w 2^64/q); result in [0,q) */
static inline u64 mulmod_shoup(u64 a, u64 w, u64 wsh, u64 q) {
    u64 hi = (u64)(((u128)a * wsh) >> 64);
    u64 r = a * w - hi * q;
    return r >= q ? r - q : r;
}
/* Barrett for x < q^2 */
static inline u64 barrett128(u128 x, u64 q, u64 mu, int n) {
    u64 x1 = (u64)(x >> (n - 1));
    u64 qh = (u64)(((u128)x1 * mu) >> (n + 1));
    u64 r = (u64)x - qh * q;
    while (r >= q) r -= q;
    return r;
}
static inline u64 mulmod(const orc_ctx* c, u64 a, u64 b, int m) {
    return barrett128((u128)a * b, c->q[m], c->mu[m], c->nbits[m]);
}
/* reduce arbitrary 128-bit value */
static inline u64 reduce128(u128 x, u64 q) { return (u64)(x % q); }

static int is_prime(u64 n) {
    if (n < 2) return 0;
    static const u64 sp[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (int i = 0; i < 12; i++) { if (n % sp[i] == 0) return n == sp[i]; }
    u64 d = n - 1; int r = 0;
    while (!(d & 1)) { d >>= 1; r++; }
    for (int i = 0; i < 12; i++) {
        u64 x = powmod(sp[i], d, n);
        if (x == 1 || x == n - 1) continue;
        int comp = 1;
        for (int j = 1; j < r; j++) { x = mulmod_slow(x, x, n); if (x == n - 1) { comp = 0; break; } }
        if (comp) return 0;
    }
    return 1;
}
/* A.2: first prime >= 2^bits + 1 congruent 1 mod m; previous / next in the same residue class */
static u64 first_prime(int bits, u64 m) { u64 q = ((u64)1 << bits) + 1; while (!is_prime(q)) q += m; return q; }
static u64 prev_prime(u64 q, u64 m) { do { q -= m; } while (!is_prime(q)); return q; }
static u64 next_prime(u64 q, u64 m) { do { q += m; } while (!is_prime(q)); return q; }

/* A.3: numerically smallest primitive m-th root of unity mod q (m a power of two) */
static u64 min_root_of_unity(u64 m, u64 q) {
    u64 g = 0;
    for (u64 x = 2;; x++) {
        u64 r = powmod(x, (q - 1) / m, q);
        if (powmod(r, m / 2, q) == q - 1) { g = r; break; }   /* r has exact order m */
    }
    u64 g2 = mulmod_slow(g, g, q), cur = g, best = g;
    for (u64 i = 1; i < m / 2; i++) { cur = mulmod_slow(cur, g2, q); if (cur < best) best = cur; }
    return best;
}

static int bitlen(u64 x) { int n = 0; while (x) { n++; x >>= 1; } return n; }

/* ---------- context ---------- */
orc_ctx* orc_create(int logN, int L, int dnum, int first_bits, int scale_bits, int aux_bits) {
    orc_ctx* c = (orc_ctx*)calloc(1, sizeof(orc_ctx));
    c->logN = logN; c->N = 1 << logN; c->L = L; c->dnum = dnum;
    int N = c->N; u64 M = 2 * (u64)N;
    /* A.2 FLEXIBLEAUTO prime chain, built last-first (reference parameters F.cpp:18-25) */
    u64* q = c->q;
    q[L - 1] = first_prime(scale_bits, M);
    if (L > 1) {
        double sf = (double)q[L - 1];
        unsigned cnt = 0;
        for (int i = L - 2; i >= 1; i--) {
            sf = sf * sf / (double)q[i + 1];
            u64 sfi = (u64)llround(sf), rem = sfi % M, cand;
            int same;
            if (cnt % 2 == 0) {
                cand = sfi - M - rem + 1;
                do { cand = prev_prime(cand, M); same = 0; for (int j = i + 1; j < L; j++) if (q[j] == cand) same = 1; } while (same);
            } else {
                cand = sfi + M - rem + 1;
                do { cand = next_prime(cand, M); same = 0; for (int j = i + 1; j < L; j++) if (q[j] == cand) same = 1; } while (same);
            }
            q[i] = cand; cnt++;
        }
        if (first_bits == scale_bits) {
            u64 cand = q[1]; int same;
            do { cand = prev_prime(cand, M); same = 0; for (int j = 1; j < L; j++) if (q[j] == cand) same = 1; } while (same);
            q[0] = cand;
        } else {
            q[0] = prev_prime(first_prime(first_bits, M), M);
        }
    }
    /* A.6 digits and P chain */
    c->alpha = (L + dnum - 1) / dnum;
    while (c->alpha * (c->dnum - 1) >= L && c->dnum > 1) c->dnum--;
    double maxbits = 0;
    for (int d = 0; d < c->dnum; d++) {
        double b = 0;
        for (int i = d * c->alpha; i < (d + 1) * c->alpha && i < L; i++) b += log2((double)q[i]);
        if (b > maxbits) maxbits = b;
    }
    c->K = (int)ceil(ceil(maxbits) / aux_bits);
    {
        u64 p = first_prime(aux_bits, M);
        for (int k = 0; k < c->K; k++) {
            int inq;
            do { p = prev_prime(p, M); inq = 0; for (int j = 0; j < L; j++) if (q[j] == p) inq = 1; } while (inq);
            q[L + k] = p;
        }
    }
    /* A.8 scaling factors: sf[0] = q_{L-1}; sf[i+1] = sf[i]^2 / q_{L-1-i} */
    c->sf[0] = (double)q[L - 1];
    for (int i = 0; i + 1 < L; i++) c->sf[i + 1] = c->sf[i] * c->sf[i] / (double)q[L - 1 - i];
    /* tables */
    c->brev = (uint32_t*)malloc(sizeof(uint32_t) * N);
    for (int i = 0; i < N; i++) { uint32_t r = 0; for (int b = 0; b < logN; b++) if (i >> b & 1) r |= 1u << (logN - 1 - b); c->brev[i] = r; }
    int T = L + c->K;
    #pragma omp parallel for schedule(dynamic)
    for (int m = 0; m < T; m++) {
        u64 qq = q[m];
        c->nbits[m] = bitlen(qq);
        c->mu[m] = (u64)((((u128)1) << (2 * c->nbits[m])) / qq);
        c->psi[m] = min_root_of_unity(M, qq);
        c->psi_inv[m] = invmod(c->psi[m], qq);
        c->n_inv[m] = invmod((u64)N, qq); c->n_inv_sh[m] = shoup(c->n_inv[m], qq);
        c->tw[m] = (u64*)malloc(8 * N); c->tw_sh[m] = (u64*)malloc(8 * N);
        c->itw[m] = (u64*)malloc(8 * N); c->itw_sh[m] = (u64*)malloc(8 * N);
        u64 p = 1, pi = 1;
        for (int i = 0; i < N; i++) {   /* A.4: tables in bit-reversed order */
            uint32_t r = c->brev[i];
            c->tw[m][r] = p; c->itw[m][r] = pi;
            p = mulmod_slow(p, c->psi[m], qq); pi = mulmod_slow(pi, c->psi_inv[m], qq);
        }
        for (int i = 0; i < N; i++) { c->tw_sh[m][i] = shoup(c->tw[m][i], qq); c->itw_sh[m][i] = shoup(c->itw[m][i], qq); }
    }
    return c;
}
void orc_destroy(orc_ctx* c) {
    if (!c) return;
    for (int m = 0; m < c->L + c->K; m++) { free(c->tw[m]); free(c->tw_sh[m]); free(c->itw[m]); free(c->itw_sh[m]); }
    free(c->brev); free(c);
}
void orc_info(const orc_ctx* c, int* info) { info[0] = c->logN; info[1] = c->L; info[2] = c->K; info[3] = c->alpha; info[4] = c->dnum; }
void orc_moduli(const orc_ctx* c, u64* out) { memcpy(out, c->q, 8 * (c->L + c->K)); }
void orc_roots(const orc_ctx* c, u64* out) { memcpy(out, c->psi, 8 * (c->L + c->K)); }
void orc_scale_factors(const orc_ctx* c, double* sf) { memcpy(sf, c->sf, 8 * c->L); }
/* bench.py's reference arm runs under torch.distributed.run, which exports OMP_NUM_THREADS=1: the thread count is set explicitly */
void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---------- NTT (A.4): forward = Cooley-Tukey natural -> bit-reversed; inverse = Gentleman-Sande ---------- */
static void ntt_limb(const orc_ctx* c, u64* a, int m) {
    const u64 q = c->q[m], q2 = 2 * q; const u64 *tw = c->tw[m], *sh = c->tw_sh[m];
    int N = c->N, t = N;
    for (int mm = 1; mm < N; mm <<= 1) {
        t >>= 1;
        for (int i = 0; i < mm; i++) {
            u64 w = tw[mm + i], ws = sh[mm + i];
            u64* x = a + 2 * i * t; u64* y = x + t;
            for (int j = 0; j < t; j++) {   /* Harvey lazy butterfly, values in [0,4q) */
                u64 u = x[j]; if (u >= q2) u -= q2;
                u64 hi = (u64)(((u128)y[j] * ws) >> 64);
                u64 v = y[j] * w - hi * q;
                x[j] = u + v; y[j] = u - v + q2;
            }
        }
    }
    for (int j = 0; j < N; j++) { u64 v = a[j]; if (v >= q2) v -= q2; if (v >= q) v -= q; a[j] = v; }
}
static void intt_limb(const orc_ctx* c, u64* a, int m) {
    const u64 q = c->q[m], q2 = 2 * q; const u64 *tw = c->itw[m], *sh = c->itw_sh[m];
    int N = c->N, t = 1;
    for (int mm = N >> 1; mm >= 1; mm >>= 1) {
        for (int i = 0; i < mm; i++) {
            u64 w = tw[mm + i], ws = sh[mm + i];
            u64* x = a + 2 * i * t; u64* y = x + t;
            for (int j = 0; j < t; j++) {   /* values in [0,2q) */
                u64 u = x[j], v = y[j];
                u64 s = u + v; if (s >= q2) s -= q2;
                u64 d = u - v + q2;
                u64 hi = (u64)(((u128)d * ws) >> 64);
                x[j] = s; y[j] = d * w - hi * q;
            }
        }
        t <<= 1;
    }
    const u64 ni = c->n_inv[m], nis = c->n_inv_sh[m];
    for (int j = 0; j < N; j++) a[j] = mulmod_shoup(a[j] >= q2 ? a[j] - q2 : a[j], ni, nis, q);
}
void orc_ntt(const orc_ctx* c, u64* a, const int* midx, int nl) {
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < nl; i++) ntt_limb(c, a + (size_t)i * c->N, midx[i]);
}
void orc_intt(const orc_ctx* c, u64* a, const int* midx, int nl) {
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < nl; i++) intt_limb(c, a + (size_t)i * c->N, midx[i]);
}

/* ---------- elementwise (F.cpp:409-431 EvalAdd / EvalMult cores) ---------- */
void orc_add(const orc_ctx* c, u64* out, const u64* a, const u64* b, const int* midx, int nl) {
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < nl; i++) { u64 q = c->q[midx[i]]; size_t o = (size_t)i * c->N; for (int j = 0; j < c->N; j++) out[o + j] = addmod(a[o + j], b[o + j], q); }
}
void orc_sub(const orc_ctx* c, u64* out, const u64* a, const u64* b, const int* midx, int nl) {
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < nl; i++) { u64 q = c->q[midx[i]]; size_t o = (size_t)i * c->N; for (int j = 0; j < c->N; j++) out[o + j] = submod(a[o + j], b[o + j], q); }
}
void orc_mul(const orc_ctx* c, u64* out, const u64* a, const u64* b, const int* midx, int nl) {
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < nl; i++) { int m = midx[i]; size_t o = (size_t)i * c->N; for (int j = 0; j < c->N; j++) out[o + j] = mulmod(c, a[o + j], b[o + j], m); }
}
static u64 signed128_mod(u64 lo, int64_t hi, u64 q) {
    i128 v = ((i128)hi << 64) | (i128)(u128)lo;
    i128 r = v % (i128)q; if (r < 0) r += q;
    return (u64)r;
}
void orc_mul_scalar(const orc_ctx* c, u64* out, const u64* a, u64 s_lo, int64_t s_hi, const int* midx, int nl) {
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < nl; i++) {
        int m = midx[i]; u64 q = c->q[m]; u64 s = signed128_mod(s_lo, s_hi, q), ss = shoup(s, q);
        size_t o = (size_t)i * c->N;
        for (int j = 0; j < c->N; j++) out[o + j] = mulmod_shoup(a[o + j], s, ss, q);
    }
}

/* ---------- automorphism (A.5; behind EvalRotate F.cpp:435) ---------- */
uint32_t orc_galois_for_rotation(const orc_ctx* c, int k) {
    u64 M = 2 * (u64)c->N, g = 1, base = 5;
    if (k < 0) { /* inverse of 5 mod 2N */
        u64 inv = 1; for (u64 x = 1; x < M; x += 2) if ((x * 5) % M == 1) { inv = x; break; }
        base = inv; k = -k;
    }
    for (int i = 0; i < k; i++) g = g * base % M;
    return (uint32_t)g;
}
uint32_t orc_galois_conj(const orc_ctx* c) { return (uint32_t)(2 * c->N - 1); }
void orc_automorph_eval(const orc_ctx* c, u64* out, const u64* in, int nl, uint32_t g) {
    int N = c->N; uint32_t M = 2 * N;
    uint32_t* map = (uint32_t*)malloc(4 * N);
    for (int j = 0; j < N; j++) { uint32_t t = (uint32_t)(((u64)(2 * j + 1) * g) % M); map[c->brev[j]] = c->brev[(t - 1) / 2]; }
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < nl; i++) { size_t o = (size_t)i * N; for (int j = 0; j < N; j++) out[o + j] = in[o + map[j]]; }
    free(map);
}
void orc_automorph_coeff(const orc_ctx* c, u64* out, const u64* in, const int* midx, int nl, uint32_t g) {
    int N = c->N; uint32_t M = 2 * N;
    for (int i = 0; i < nl; i++) {
        u64 q = c->q[midx[i]]; size_t o = (size_t)i * N;
        for (int j = 0; j < N; j++) {
            uint32_t t = (uint32_t)(((u64)j * g) % M);
            if (t < (uint32_t)N) out[o + t] = in[o + j]; else out[o + t - N] = in[o + j] ? q - in[o + j] : 0;
        }
    }
}

/* ---------- rescale (A.7; implicit under FLEXIBLEAUTO F.cpp:18) ---------- */
void orc_rescale(const orc_ctx* c, u64* out, const u64* in, int l) {
    int N = c->N, last = l - 1;
    u64* x = (u64*)malloc(8 * N);
    memcpy(x, in + (size_t)last * N, 8 * N);
    intt_limb(c, x, last);
    u64 ql = c->q[last], half = ql >> 1;
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < last; i++) {
        u64 q = c->q[i]; u64* t = (u64*)malloc(8 * N);
        u64 qlm = ql % q;
        for (int j = 0; j < N; j++) { u64 v = x[j] % q; if (x[j] > half) v = submod(v, qlm, q); t[j] = v; }  /* centred switch */
        ntt_limb(c, t, i);
        u64 inv = invmod(ql % q, q), invs = shoup(inv, q);
        for (int j = 0; j < N; j++) out[(size_t)i * N + j] = mulmod_shoup(submod(in[(size_t)i * N + j], t[j], q), inv, invs, q);
        free(t);
    }
    free(x);
}

/* ---------- hybrid key switching (A.6; behind EvalRotate/EvalMult F.cpp:431,435) ---------- */
/* fast (approximate) basis conversion of coefficient-form limbs src (moduli sm[0..ns)) to modulus index tm */
static void conv_consts(const orc_ctx* c, const int* sm, int ns, u64* hatinv /* ns */) {
    for (int i = 0; i < ns; i++) {
        u64 qi = c->q[sm[i]], h = 1;
        for (int k = 0; k < ns; k++) if (k != i) h = mulmod_slow(h, c->q[sm[k]] % qi, qi);
        hatinv[i] = invmod(h, qi);
    }
}
static u64 hat_mod(const orc_ctx* c, const int* sm, int ns, int i, u64 t) {
    u64 h = 1; for (int k = 0; k < ns; k++) if (k != i) h = mulmod_slow(h, c->q[sm[k]] % t, t); return h;
}
/* y[i][j] = src[i][j]*hatinv[i] mod q_i (precomputed by caller); out[j] = sum_i y[i][j]*(Qhat_i mod t) mod t */
static void conv_to(const orc_ctx* c, u64* out, u64* const* y, const int* sm, int ns, int tm) {
    u64 t = c->q[tm], hm[MAXLIMBS];
    for (int i = 0; i < ns; i++) hm[i] = hat_mod(c, sm, ns, i, t);
    for (int j = 0; j < c->N; j++) {
        u128 acc = 0;
        for (int i = 0; i < ns; i++) acc += (u128)y[i][j] * hm[i];
        out[j] = reduce128(acc, t);
    }
}
void orc_modup(const orc_ctx* c, u64* out, const u64* in_eval, int l, int digit) {
    int N = c->N, K = c->K, lo = digit * c->alpha, hi = lo + c->alpha; if (hi > l) hi = l;
    int ns = hi - lo, sm[MAXLIMBS]; u64 hatinv[MAXLIMBS]; u64* y[MAXLIMBS];
    for (int i = 0; i < ns; i++) sm[i] = lo + i;
    conv_consts(c, sm, ns, hatinv);
    for (int i = 0; i < ns; i++) {
        y[i] = (u64*)malloc(8 * N);
        memcpy(y[i], in_eval + (size_t)(lo + i) * N, 8 * N);
        intt_limb(c, y[i], sm[i]);
        u64 q = c->q[sm[i]], s = shoup(hatinv[i], q);
        for (int j = 0; j < N; j++) y[i][j] = mulmod_shoup(y[i][j], hatinv[i], s, q);
    }
    #pragma omp parallel for schedule(dynamic)
    for (int t = 0; t < l + K; t++) {
        int tm = t < l ? t : c->L + (t - l);
        u64* o = out + (size_t)t * N;
        if (t >= lo && t < hi) { memcpy(o, in_eval + (size_t)t * N, 8 * N); continue; }
        conv_to(c, o, y, sm, ns, tm);
        ntt_limb(c, o, tm);
    }
    for (int i = 0; i < ns; i++) free(y[i]);
}
void orc_moddown(const orc_ctx* c, u64* out, const u64* in, int l) {
    int N = c->N, K = c->K, L = c->L, sm[MAXLIMBS]; u64 hatinv[MAXLIMBS]; u64* y[MAXLIMBS];
    for (int k = 0; k < K; k++) sm[k] = L + k;
    conv_consts(c, sm, K, hatinv);
    for (int k = 0; k < K; k++) {
        y[k] = (u64*)malloc(8 * N);
        memcpy(y[k], in + (size_t)(l + k) * N, 8 * N);
        intt_limb(c, y[k], sm[k]);
        u64 q = c->q[sm[k]], s = shoup(hatinv[k], q);
        for (int j = 0; j < N; j++) y[k][j] = mulmod_shoup(y[k][j], hatinv[k], s, q);
    }
    #pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < l; i++) {
        u64 q = c->q[i]; u64* t = (u64*)malloc(8 * N);
        conv_to(c, t, y, sm, K, i);
        ntt_limb(c, t, i);
        u64 pinv = 1; for (int k = 0; k < K; k++) pinv = mulmod_slow(pinv, c->q[L + k] % q, q);
        pinv = invmod(pinv, q); u64 ps = shoup(pinv, q);
        for (int j = 0; j < N; j++) out[(size_t)i * N + j] = mulmod_shoup(submod(in[(size_t)i * N + j], t[j], q), pinv, ps, q);
        free(t);
    }
    for (int k = 0; k < K; k++) free(y[k]);
}
void orc_keyswitch(const orc_ctx* c, u64* out0, u64* out1, const u64* in, const u64* evk, int l) {
    int N = c->N, K = c->K, L = c->L, T = l + K, beta = (l + c->alpha - 1) / c->alpha;
    size_t kl = (size_t)(L + K) * N;     /* one evk polynomial */
    u64* up = (u64*)malloc(8 * (size_t)T * N);
    u64* acc0 = (u64*)calloc((size_t)T * N, 8);
    u64* acc1 = (u64*)calloc((size_t)T * N, 8);
    for (int d = 0; d < beta; d++) {
        orc_modup(c, up, in, l, d);
        const u64* kb = evk + (size_t)d * 2 * kl; const u64* ka = kb + kl;
        #pragma omp parallel for schedule(static)
        for (int t = 0; t < T; t++) {
            int tm = t < l ? t : L + (t - l); u64 q = c->q[tm];
            const u64* u = up + (size_t)t * N; const u64* b = kb + (size_t)tm * N; const u64* a = ka + (size_t)tm * N;
            u64* o0 = acc0 + (size_t)t * N; u64* o1 = acc1 + (size_t)t * N;
            for (int j = 0; j < N; j++) { o0[j] = addmod(o0[j], mulmod(c, u[j], b[j], tm), q); o1[j] = addmod(o1[j], mulmod(c, u[j], a[j], tm), q); }
        }
    }
    orc_moddown(c, out0, acc0, l);
    orc_moddown(c, out1, acc1, l);
    free(up); free(acc0); free(acc1);
}
static void iota(int* v, int n) { for (int i = 0; i < n; i++) v[i] = i; }
void orc_rotate(const orc_ctx* c, u64* out, const u64* ct, int l, uint32_t g, const u64* evk) {
    int N = c->N, idx[MAXLIMBS]; iota(idx, l); size_t pl = (size_t)l * N;
    u64* k0 = (u64*)malloc(8 * pl); u64* k1 = (u64*)malloc(8 * pl);
    orc_keyswitch(c, k0, k1, ct + pl, evk, l);      /* A.5: key-switch first ... */
    orc_add(c, k0, k0, ct, idx, l);
    orc_automorph_eval(c, out, k0, l, g);           /* ... then permute both components */
    orc_automorph_eval(c, out + pl, k1, l, g);
    free(k0); free(k1);
}
void orc_mul_relin(const orc_ctx* c, u64* out, const u64* a, const u64* b, int l, const u64* evk) {
    int N = c->N, idx[MAXLIMBS]; iota(idx, l); size_t pl = (size_t)l * N;
    u64* d0 = (u64*)malloc(8 * pl); u64* d1 = (u64*)malloc(8 * pl); u64* d2 = (u64*)malloc(8 * pl); u64* t = (u64*)malloc(8 * pl);
    orc_mul(c, d0, a, b, idx, l);
    orc_mul(c, d1, a, b + pl, idx, l); orc_mul(c, t, a + pl, b, idx, l); orc_add(c, d1, d1, t, idx, l);
    orc_mul(c, d2, a + pl, b + pl, idx, l);
    u64* k0 = t; u64* k1 = (u64*)malloc(8 * pl);
    orc_keyswitch(c, k0, k1, d2, evk, l);
    orc_add(c, out, d0, k0, idx, l); orc_add(c, out + pl, d1, k1, idx, l);
    free(d0); free(d1); free(d2); free(t); free(k1);
}
void orc_mul_plain(const orc_ctx* c, u64* out, const u64* ct, const u64* pt, int l) {
    int idx[MAXLIMBS]; iota(idx, l); size_t pl = (size_t)l * c->N;
    orc_mul(c, out, ct, pt, idx, l); orc_mul(c, out + pl, ct + pl, pt, idx, l);
}

/* ---------- sampling: splitmix64 streams (shared spec with the product; DESIGN.md "Randomness") ---------- */
static inline u64 sm64_next(u64* s) { u64 z = (*s += 0x9E3779B97F4A7C15ULL); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
static inline u64 sub_seed(u64 seed, u64 tag) { u64 s = seed + tag * 0x9E3779B97F4A7C15ULL; return sm64_next(&s); }
void orc_sample_uniform(const orc_ctx* c, u64 seed, u64* out, const int* midx, int nl) {
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < nl; i++) { u64 s = sub_seed(seed, 1000 + midx[i]), q = c->q[midx[i]]; for (int j = 0; j < c->N; j++) out[(size_t)i * c->N + j] = sm64_next(&s) % q; }
}
void orc_sample_ternary(const orc_ctx* c, u64 seed, int8_t* out) {
    u64 s = seed; for (int j = 0; j < c->N; j++) { u64 r = sm64_next(&s) % 3; out[j] = r == 2 ? -1 : (int8_t)r; }
}
void orc_sample_sparse_ternary(const orc_ctx* c, u64 seed, int h, int8_t* out) {
    u64 s = seed; memset(out, 0, c->N); int placed = 0;
    while (placed < h) { u64 pos = sm64_next(&s) % c->N; u64 sg = sm64_next(&s) & 1; if (!out[pos]) { out[pos] = sg ? -1 : 1; placed++; } }
}
static const u64 GAUSS_CDT[30] = {   /* |X| cumulative, sigma = 3.19, scaled 2^64 */
    0x2003F343659528D0ULL, 0x5CF9E7DE0F06D96BULL, 0x9194FD0BB0694AF3ULL, 0xBABAA2EF1EEC1101ULL, 0xD7E6AB30AA084360ULL,
    0xEAA5B92100F77DABULL, 0xF591040AC34992E9ULL, 0xFB54CB2CA496FFFAULL, 0xFE1702297749D972ULL, 0xFF4953F8BD4AE9D1ULL,
    0xFFC1C20EF5DE7233ULL, 0xFFECAC7F021E2BA0ULL, 0xFFFA892133378B29ULL, 0xFFFE9810099A70DBULL, 0xFFFFABC31CFFB430ULL,
    0xFFFFEE138CF385DBULL, 0xFFFFFC88B8F21ED7ULL, 0xFFFFFF641EF54A94ULL, 0xFFFFFFE7207A46BAULL, 0xFFFFFFFC6560DA3AULL,
    0xFFFFFFFF86A24B98ULL, 0xFFFFFFFFF1822EC9ULL, 0xFFFFFFFFFE6DFC66ULL, 0xFFFFFFFFFFD877EFULL, 0xFFFFFFFFFFFC7916ULL,
    0xFFFFFFFFFFFFB6EAULL, 0xFFFFFFFFFFFFFAA2ULL, 0xFFFFFFFFFFFFFFA4ULL, 0xFFFFFFFFFFFFFFFAULL, 0xFFFFFFFFFFFFFFFFULL};
void orc_sample_gauss(const orc_ctx* c, u64 seed, int8_t* out) {
    u64 s = seed;
    for (int j = 0; j < c->N; j++) {
        u64 u = sm64_next(&s), sg = sm64_next(&s) & 1; int k = 0;
        while (k < 29 && u >= GAUSS_CDT[k]) k++;
        out[j] = (int8_t)(sg ? -k : k);
    }
}
void orc_small_to_eval(const orc_ctx* c, u64* out, const int8_t* in, const int* midx, int nl) {
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < nl; i++) {
        u64 q = c->q[midx[i]]; u64* o = out + (size_t)i * c->N;
        for (int j = 0; j < c->N; j++) o[j] = in[j] >= 0 ? (u64)in[j] : q - (u64)(-in[j]);
        ntt_limb(c, o, midx[i]);
    }
}

/* ---------- keys (F.cpp:47-49 KeyGen/EvalMultKeyGen; F.cpp:248 EvalRotateKeyGen) ---------- */
static void all_idx(const orc_ctx* c, int* idx) { for (int i = 0; i < c->L + c->K; i++) idx[i] = i; }
void orc_gen_sk(const orc_ctx* c, u64 seed, int h, u64* sk_eval) {
    int idx[MAXLIMBS]; all_idx(c, idx);
    int8_t* s = (int8_t*)malloc(c->N);
    if (h > 0) orc_sample_sparse_ternary(c, sub_seed(seed, 0), h, s); else orc_sample_ternary(c, sub_seed(seed, 0), s);
    orc_small_to_eval(c, sk_eval, s, idx, c->L + c->K);
    free(s);
}
void orc_gen_pk(const orc_ctx* c, u64 seed, const u64* sk_eval, u64* pk) {
    int N = c->N, L = c->L, idx[MAXLIMBS]; iota(idx, L); size_t pl = (size_t)L * N;
    int8_t* e = (int8_t*)malloc(N); u64* ee = (u64*)malloc(8 * pl);
    orc_sample_uniform(c, sub_seed(seed, 10), pk + pl, idx, L);       /* a */
    orc_sample_gauss(c, sub_seed(seed, 11), e); orc_small_to_eval(c, ee, e, idx, L);
    orc_mul(c, pk, pk + pl, sk_eval, idx, L);                          /* a*s */
    orc_sub(c, pk, ee, pk, idx, L);                                    /* b = e - a*s */
    free(e); free(ee);
}
void orc_gen_evk(const orc_ctx* c, u64 seed, const u64* sk_old, const u64* sk_new, u64* evk) {
    int N = c->N, L = c->L, K = c->K, T = L + K, idx[MAXLIMBS]; all_idx(c, idx); size_t kl = (size_t)T * N;
    int8_t* e = (int8_t*)malloc(N); u64* ee = (u64*)malloc(8 * kl);
    for (int d = 0; d < c->dnum; d++) {
        u64* b = evk + (size_t)d * 2 * kl; u64* a = b + kl;
        orc_sample_uniform(c, sub_seed(seed, 100 + 2 * d), a, idx, T);
        orc_sample_gauss(c, sub_seed(seed, 101 + 2 * d), e); orc_small_to_eval(c, ee, e, idx, T);
        orc_mul(c, b, a, sk_new, idx, T); orc_sub(c, b, ee, b, idx, T);       /* e - a*s_new */
        int lo = d * c->alpha, hi = lo + c->alpha; if (hi > L) hi = L;
        #pragma omp parallel for schedule(static)
        for (int i = lo; i < hi; i++) {                                        /* + (P mod q_i) * s_old on the digit's limbs */
            u64 q = c->q[i], pm = 1; for (int k = 0; k < K; k++) pm = mulmod_slow(pm, c->q[L + k] % q, q);
            u64 ps = shoup(pm, q);
            for (int j = 0; j < N; j++) { size_t o = (size_t)i * N + j; b[o] = addmod(b[o], mulmod_shoup(sk_old[o], pm, ps, q), q); }
        }
    }
    free(e); free(ee);
}
void orc_gen_relin_key(const orc_ctx* c, u64 seed, const u64* sk_eval, u64* evk) {
    int T = c->L + c->K, idx[MAXLIMBS]; all_idx(c, idx);
    u64* s2 = (u64*)malloc(8 * (size_t)T * c->N);
    orc_mul(c, s2, sk_eval, sk_eval, idx, T);
    orc_gen_evk(c, seed, s2, sk_eval, evk);
    free(s2);
}
static uint32_t inv_mod_pow2(uint32_t g, uint32_t M) { uint32_t x = 1; for (int i = 0; i < 6; i++) x = x * (2 - g * x); return x & (M - 1); }
void orc_gen_galois_key(const orc_ctx* c, u64 seed, const u64* sk_eval, uint32_t g, u64* evk) {
    /* A.5 / EvalAutomorphismKeyGen: switch from s to sigma_{g^-1}(s); the permutation by g afterwards restores s */
    int T = c->L + c->K; uint32_t gi = inv_mod_pow2(g, 2 * c->N);
    u64* sp = (u64*)malloc(8 * (size_t)T * c->N);
    orc_automorph_eval(c, sp, sk_eval, T, gi);
    orc_gen_evk(c, seed, sk_eval, sp, evk);
    free(sp);
}

/* ---------- CKKS encoding (A.9; MakeCKKSPackedPlaintext F.cpp:353) ---------- */
typedef struct { double re, im; } cplx;
static void special_tables(int n, uint32_t** rot, cplx** ksi) {
    uint32_t m = 4u * n; *rot = (uint32_t*)malloc(4 * (n > 0 ? n : 1)); *ksi = (cplx*)malloc(sizeof(cplx) * (m + 1));
    uint32_t p = 1; for (int i = 0; i < n; i++) { (*rot)[i] = p; p = (uint32_t)(((u64)p * 5) % m); }
    for (uint32_t k = 0; k <= m; k++) { double a = 2.0 * M_PI * (double)k / (double)m; (*ksi)[k].re = cos(a); (*ksi)[k].im = sin(a); }
}
static void bitrev_cplx(cplx* v, int n) {
    for (int i = 1, j = 0; i < n; i++) { int bit = n >> 1; for (; j >= bit; bit >>= 1) j -= bit; j += bit; if (i < j) { cplx t = v[i]; v[i] = v[j]; v[j] = t; } }
}
static void fft_special_inv(cplx* v, int n) {
    uint32_t* rot; cplx* ksi; special_tables(n, &rot, &ksi); uint32_t m = 4u * n;
    for (int len = n; len >= 2; len >>= 1) {
        int lenh = len >> 1; uint32_t lenq = (uint32_t)len << 2;
        for (int i = 0; i < n; i += len)
            for (int j = 0; j < lenh; j++) {
                uint32_t idx = (lenq - (rot[j] % lenq)) * (m / lenq);
                cplx u = {v[i + j].re + v[i + j + lenh].re, v[i + j].im + v[i + j + lenh].im};
                cplx w = {v[i + j].re - v[i + j + lenh].re, v[i + j].im - v[i + j + lenh].im};
                cplx k = ksi[idx];
                v[i + j] = u;
                v[i + j + lenh].re = w.re * k.re - w.im * k.im; v[i + j + lenh].im = w.re * k.im + w.im * k.re;
            }
    }
    bitrev_cplx(v, n);
    for (int i = 0; i < n; i++) { v[i].re /= n; v[i].im /= n; }
    free(rot); free(ksi);
}
static void fft_special(cplx* v, int n) {
    uint32_t* rot; cplx* ksi; special_tables(n, &rot, &ksi); uint32_t m = 4u * n;
    bitrev_cplx(v, n);
    for (int len = 2; len <= n; len <<= 1) {
        int lenh = len >> 1; uint32_t lenq = (uint32_t)len << 2;
        for (int i = 0; i < n; i += len)
            for (int j = 0; j < lenh; j++) {
                uint32_t idx = (rot[j] % lenq) * (m / lenq);
                cplx k = ksi[idx], a = v[i + j], b = v[i + j + lenh];
                cplx w = {b.re * k.re - b.im * k.im, b.re * k.im + b.im * k.re};
                v[i + j].re = a.re + w.re; v[i + j].im = a.im + w.im;
                v[i + j + lenh].re = a.re - w.re; v[i + j + lenh].im = a.im - w.im;
            }
    }
    free(rot); free(ksi);
}
static void encode_coeffs128(const orc_ctx* c, i128* co, const double* vals, int slots, double scale) {
    int N = c->N, Nh = N / 2, gap = Nh / slots;
    cplx* v = (cplx*)malloc(sizeof(cplx) * slots);
    for (int i = 0; i < slots; i++) { v[i].re = vals[2 * i]; v[i].im = vals[2 * i + 1]; }
    fft_special_inv(v, slots);
    for (int i = 0; i < N; i++) co[i] = 0;
    for (int i = 0; i < slots; i++) {
        co[i * gap] = (i128)rint(v[i].re * scale);
        co[Nh + i * gap] = (i128)rint(v[i].im * scale);
    }
    free(v);
}
void orc_encode_coeffs(const orc_ctx* c, int64_t* out, const double* vals, int slots, double scale) {
    i128* co = (i128*)malloc(sizeof(i128) * c->N);
    encode_coeffs128(c, co, vals, slots, scale);
    for (int i = 0; i < c->N; i++) out[i] = (int64_t)co[i];
    free(co);
}
void orc_encode(const orc_ctx* c, u64* out, const double* vals, int slots, double scale, int l) {
    int N = c->N; i128* co = (i128*)malloc(sizeof(i128) * N);
    encode_coeffs128(c, co, vals, slots, scale);
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < l; i++) {
        i128 q = (i128)c->q[i]; u64* o = out + (size_t)i * N;
        for (int j = 0; j < N; j++) { i128 r = co[j] % q; if (r < 0) r += q; o[j] = (u64)r; }
        ntt_limb(c, o, i);
    }
    free(co);
}
void orc_decode(const orc_ctx* c, double* vals, const u64* poly, int slots, double scale, int l) {
    int N = c->N, Nh = N / 2, gap = Nh / slots, nl = l >= 2 ? 2 : 1;
    u64* x = (u64*)malloc(8 * (size_t)nl * N); memcpy(x, poly, 8 * (size_t)nl * N);
    int idx[2] = {0, 1}; orc_intt(c, x, idx, nl);
    double* co = (double*)malloc(8 * N);
    if (nl == 1) {
        u64 q = c->q[0];
        for (int j = 0; j < N; j++) co[j] = x[j] > q / 2 ? -(double)(q - x[j]) : (double)x[j];
    } else {
        u64 q0 = c->q[0], q1 = c->q[1], q0inv = invmod(q0 % q1, q1); u128 Q = (u128)q0 * q1;
        for (int j = 0; j < N; j++) {
            u64 d = submod(x[N + j], x[j] % q1, q1);
            u128 v = (u128)x[j] + (u128)q0 * mulmod_slow(d, q0inv, q1);
            co[j] = v > Q / 2 ? -(double)(Q - v) : (double)v;
        }
    }
    cplx* v = (cplx*)malloc(sizeof(cplx) * slots);
    for (int i = 0; i < slots; i++) { v[i].re = co[i * gap] / scale; v[i].im = co[Nh + i * gap] / scale; }
    fft_special(v, slots);
    for (int i = 0; i < slots; i++) { vals[2 * i] = v[i].re; vals[2 * i + 1] = v[i].im; }
    free(x); free(co); free(v);
}

/* ---------- encrypt / decrypt (F.cpp:373-404) ---------- */
void orc_encrypt(const orc_ctx* c, u64 seed, u64* ct, const u64* pt, const u64* pk, int l) {
    int N = c->N, L = c->L, idx[MAXLIMBS]; iota(idx, l); size_t pl = (size_t)l * N, pkl = (size_t)L * N;
    int8_t* s = (int8_t*)malloc(N); u64* v = (u64*)malloc(8 * pl); u64* e = (u64*)malloc(8 * pl);
    orc_sample_ternary(c, sub_seed(seed, 1), s); orc_small_to_eval(c, v, s, idx, l);
    orc_sample_gauss(c, sub_seed(seed, 2), s); orc_small_to_eval(c, e, s, idx, l);
    orc_mul(c, ct, pk, v, idx, l); orc_add(c, ct, ct, e, idx, l); orc_add(c, ct, ct, pt, idx, l);
    orc_sample_gauss(c, sub_seed(seed, 3), s); orc_small_to_eval(c, e, s, idx, l);
    orc_mul(c, ct + pl, pk + pkl, v, idx, l); orc_add(c, ct + pl, ct + pl, e, idx, l);
    free(s); free(v); free(e);
}
void orc_decrypt(const orc_ctx* c, u64* pt, const u64* ct, const u64* sk, int l, int ncomp) {
    int N = c->N, idx[MAXLIMBS]; iota(idx, l); size_t pl = (size_t)l * N;
    u64* t = (u64*)malloc(8 * pl);
    orc_mul(c, pt, ct + pl, sk, idx, l); orc_add(c, pt, pt, ct, idx, l);
    if (ncomp == 3) { orc_mul(c, t, ct + 2 * pl, sk, idx, l); orc_mul(c, t, t, sk, idx, l); orc_add(c, pt, pt, t, idx, l); }
    free(t);
}
