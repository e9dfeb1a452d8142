// params.h -- host-side CKKS parameter set for the B200 engine: RNS prime chain, 2N-th roots,
// twiddle tables and CRT constants.  Replaces what GenCryptoContext builds for the reference
// (/root/reference/src/FHEController.cpp:4-37); conventions per SURVEY.md Appendix A.1-A.8.
#pragma once
#include <cstdint>
#include <vector>

namespace flk {

using u64 = uint64_t;
using u128 = unsigned __int128;
using i128 = __int128;

struct ParamSpec {
    int logN = 15;
    int L = 28;          // Q limbs = multiplicative depth + 1 (FLEXIBLEAUTO)
    int dnum = 4;        // key-switch digits
    int first_bits = 55;
    int scale_bits = 52;
    int aux_bits = 60;
    int sparse_h = 192;  // SPARSE_TERNARY Hamming weight (0 = uniform ternary)
};

namespace nt {
u64 mulmod(u64 a, u64 b, u64 q);
u64 powmod(u64 a, u64 e, u64 q);
u64 invmod(u64 a, u64 q);
bool is_prime(u64 n);
u64 shoup(u64 w, u64 q);
u64 min_primitive_root(u64 order, u64 q);
}  // namespace nt

struct Params {
    ParamSpec spec;
    int logN, N, L, K, dnum, alpha, T;   // T = L + K
    std::vector<u64> q;                   // moduli, Q then P
    std::vector<u64> psi, psi_inv;        // 2N-th primitive roots
    std::vector<u64> mu_hi, mu_lo;        // floor(2^128 / q)
    std::vector<u64> ninv, ninv_sh;       // N^-1 mod q
    std::vector<double> sf;               // scaling factor per level
    std::vector<uint32_t> brev;           // logN-bit reversal

    explicit Params(const ParamSpec& s);
    // twiddles for modulus m: psi^bitrev(i) (and inverse), with Shoup companions; each length N
    void twiddles(int m, u64* tw, u64* tw_sh, u64* itw, u64* itw_sh) const;
    int beta(int l) const { return (l + alpha - 1) / alpha; }
    int mod_index_ext(int l, int t) const { return t < l ? t : L + (t - l); }
    // fast basis conversion constants for source modulus set sm[0..ns): hatinv[i] = (S/s_i)^-1 mod s_i
    void conv_hatinv(const int* sm, int ns, u64* hatinv) const;
    u64 conv_hat_mod(const int* sm, int ns, int i, u64 t) const;   // (S/s_i) mod t
    u64 P_mod(u64 t) const;                                         // prod(P) mod t
    uint32_t galois_for_rotation(int k) const;
    uint32_t galois_conj() const { return 2u * N - 1; }
    void automorph_map(uint32_t g, uint32_t* map) const;            // eval-domain gather map (A.5)
};

}  // namespace flk
