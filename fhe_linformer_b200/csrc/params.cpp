// params.cpp -- prime chain / root / CRT-constant generation for the B200 CKKS engine.
// Stands in for GenCryptoContext (reference FHEController.cpp:37); rules per SURVEY.md Appendix A.
#include "params.h"

#include <string>

#include <cmath>
#include <stdexcept>

namespace flk {
namespace nt {

u64 mulmod(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }

u64 powmod(u64 a, u64 e, u64 q) {
    u64 r = 1 % q;
    a %= q;
    for (; e; e >>= 1) {
        if (e & 1) r = mulmod(r, a, q);
        a = mulmod(a, a, q);
    }
    return r;
}

u64 invmod(u64 a, u64 q) { return powmod(a % q, q - 2, q); }

u64 shoup(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }

bool is_prime(u64 n) {
    if (n < 4) return n == 2 || n == 3;
    if (!(n & 1)) return false;
    u64 d = n - 1;
    int s = 0;
    while (!(d & 1)) { d >>= 1; ++s; }
    // deterministic witness set for 64-bit integers
    for (u64 a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        if (n % a == 0) return n == a;
        u64 x = powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool witness = true;
        for (int r = 1; r < s && witness; ++r) {
            x = mulmod(x, x, n);
            if (x == n - 1) witness = false;
        }
        if (witness) return false;
    }
    return true;
}

// smallest element of exact multiplicative order `order` (a power of two) in Z_q^*  (A.3)
u64 min_primitive_root(u64 order, u64 q) {
    u64 g = 0;
    for (u64 x = 2; !g; ++x) {
        u64 r = powmod(x, (q - 1) / order, q);
        if (powmod(r, order / 2, q) == q - 1) g = r;
    }
    u64 step = mulmod(g, g, q), cur = g, best = g;
    for (u64 i = 1; i < order / 2; ++i) {
        cur = mulmod(cur, step, q);
        if (cur < best) best = cur;
    }
    return best;
}

}  // namespace nt

namespace {

u64 step_prime(u64 from, u64 m, bool up) {
    u64 c = from;
    do { c = up ? c + m : c - m; } while (!nt::is_prime(c));
    return c;
}

u64 first_prime_above(int bits, u64 m) {
    u64 c = (u64(1) << bits) + 1;
    while (!nt::is_prime(c)) c += m;
    return c;
}

bool contains(const std::vector<u64>& v, int lo, int hi, u64 x) {
    for (int i = lo; i < hi; ++i)
        if (v[i] == x) return true;
    return false;
}

}  // namespace

Params::Params(const ParamSpec& s) : spec(s) {
    logN = s.logN; N = 1 << logN; L = s.L; dnum = s.dnum;
    if (L < 1 || L > 60) throw std::invalid_argument("L out of range");
    const u64 M = 2 * (u64)N;
    q.assign(L, 0);
    // A.2: scaling primes chosen last-first, alternating below/above the running scale target
    q[L - 1] = first_prime_above(s.scale_bits, M);
    if (L > 1) {
        double target = (double)q[L - 1];
        for (int i = L - 2, turn = 0; i >= 1; --i, ++turn) {
            target = target * target / (double)q[i + 1];
            u64 t = (u64)std::llround(target), rem = t % M;
            bool up = turn & 1;
            u64 c = up ? t + M - rem + 1 : t - M - rem + 1;
            do { c = step_prime(c, M, up); } while (contains(q, i + 1, L, c));
            q[i] = c;
        }
        if (s.first_bits == s.scale_bits) {
            u64 c = q[1];
            do { c = step_prime(c, M, false); } while (contains(q, 1, L, c));
            q[0] = c;
        } else {
            q[0] = step_prime(first_prime_above(s.first_bits, M), M, false);
        }
    }
    // A.6: contiguous digits of alpha limbs; P sized to cover the widest digit
    alpha = (L + dnum - 1) / dnum;
    while (dnum > 1 && alpha * (dnum - 1) >= L) --dnum;
    double widest = 0;
    for (int d = 0; d < dnum; ++d) {
        double bits = 0;
        for (int i = d * alpha; i < std::min(L, (d + 1) * alpha); ++i) bits += std::log2((double)q[i]);
        widest = std::max(widest, bits);
    }
    K = (int)std::ceil(std::ceil(widest) / s.aux_bits);
    T = L + K;
    // launch descriptors index moduli with 8 bits and address at most 256 limbs per ModUp (device_ctx.h LimbSel): reject what
    // they cannot hold here, at context creation, instead of corrupting memory in the first key switch
    if (T > 255) throw std::invalid_argument("modulus chain too long: L + K = " + std::to_string(T) + " > 255");
    if (dnum * T - L > 256)
        throw std::invalid_argument("dnum (L + K) - L = " + std::to_string(dnum * T - L) + " extended limbs per ModUp exceed the 256 a launch addresses: use fewer digits");
    u64 p = first_prime_above(s.aux_bits, M);
    for (int k = 0; k < K; ++k) {
        do { p = step_prime(p, M, false); } while (contains(q, 0, L, p));
        q.push_back(p);
    }
    // A.8 scaling factors
    sf.assign(L, 0.0);
    sf[0] = (double)q[L - 1];
    for (int i = 0; i + 1 < L; ++i) sf[i + 1] = sf[i] * sf[i] / (double)q[L - 1 - i];

    brev.resize(N);
    for (int i = 0; i < N; ++i) {
        uint32_t r = 0;
        for (int b = 0; b < logN; ++b) r |= ((i >> b) & 1u) << (logN - 1 - b);
        brev[i] = r;
    }
    psi.resize(T); psi_inv.resize(T); mu_hi.resize(T); mu_lo.resize(T); ninv.resize(T); ninv_sh.resize(T);
    for (int m = 0; m < T; ++m) {
        psi[m] = nt::min_primitive_root(M, q[m]);
        psi_inv[m] = nt::invmod(psi[m], q[m]);
        u128 mu = ~(u128)0 / q[m];
        mu_hi[m] = (u64)(mu >> 64); mu_lo[m] = (u64)mu;
        ninv[m] = nt::invmod((u64)N, q[m]);
        ninv_sh[m] = nt::shoup(ninv[m], q[m]);
    }
}

void Params::twiddles(int m, u64* tw, u64* tw_sh, u64* itw, u64* itw_sh) const {
    const u64 qq = q[m];
    u64 f = 1, b = 1;
    for (int i = 0; i < N; ++i) {
        uint32_t r = brev[i];
        tw[r] = f; itw[r] = b;
        f = nt::mulmod(f, psi[m], qq);
        b = nt::mulmod(b, psi_inv[m], qq);
    }
    for (int i = 0; i < N; ++i) {
        tw_sh[i] = nt::shoup(tw[i], qq);
        itw_sh[i] = nt::shoup(itw[i], qq);
    }
}

void Params::conv_hatinv(const int* sm, int ns, u64* hatinv) const {
    for (int i = 0; i < ns; ++i) {
        u64 qi = q[sm[i]], h = 1;
        for (int k = 0; k < ns; ++k)
            if (k != i) h = nt::mulmod(h, q[sm[k]] % qi, qi);
        hatinv[i] = nt::invmod(h, qi);
    }
}

u64 Params::conv_hat_mod(const int* sm, int ns, int i, u64 t) const {
    u64 h = 1;
    for (int k = 0; k < ns; ++k)
        if (k != i) h = nt::mulmod(h, q[sm[k]] % t, t);
    return h;
}

u64 Params::P_mod(u64 t) const {
    u64 r = 1;
    for (int k = 0; k < K; ++k) r = nt::mulmod(r, q[L + k] % t, t);
    return r;
}

uint32_t Params::galois_for_rotation(int k) const {
    const u64 M = 2 * (u64)N;
    u64 base = 5;
    if (k < 0) {   // 5^-1 mod 2N by Newton iteration on a power of two
        u64 x = 1;
        for (int i = 0; i < 6; ++i) x = x * (2 - 5 * x);
        base = x & (M - 1);
        k = -k;
    }
    u64 g = 1;
    for (int i = 0; i < k; ++i) g = g * base % M;
    return (uint32_t)g;
}

void Params::automorph_map(uint32_t g, uint32_t* map) const {
    const uint32_t M = 2u * N;
    for (int j = 0; j < N; ++j) {
        uint32_t t = (uint32_t)(((u64)(2 * j + 1) * g) % M);
        map[brev[j]] = brev[(t - 1) >> 1];
    }
}

}  // namespace flk
