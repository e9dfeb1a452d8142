// scheme.h -- CKKS scheme layer over the device engine: keys, encode/encrypt/decrypt, FLEXIBLEAUTO level and
// scale management, polynomial evaluation and bootstrapping.  This is what stands behind the reference's
// `CryptoContext<DCRTPoly> context` (FHEController.h:23); every public method names the OpenFHE call it replaces.
#pragma once
#include <complex>
#include <functional>
#include <map>
#include <memory>
#include <vector>

#include "chacha.cuh"
#include "engine.h"

namespace flk {

using cplx = std::complex<double>;

struct DevMem {   // stream-ordered HBM allocation, returned to the engine pool on destruction
    Engine* eng;
    u64* p;
    size_t words;
    DevMem(Engine* e, size_t w) : eng(e), p(e->alloc(w)), words(w) {}
    ~DevMem() { try { eng->release(p); } catch (...) {} }
    DevMem(const DevMem&) = delete;
};
using Mem = std::shared_ptr<DevMem>;

// Ciphertext (ncomp = 2) or plaintext (ncomp = 1): ncomp polynomials of l limbs, evaluation format.
// batch > 1: that many ciphertexts with identical metadata stored back to back ([batch][ncomp][l][N]); the leveled
// operations below treat them as one operand (one kernel launch per stage for the whole batch) -- this is how the
// independent row ciphertexts of the reference's `for (i < rows.size())` loops (F.cpp:872-1120) share launches.
struct Elem {
    Mem mem;
    size_t off = 0;     // first word inside mem (slices of a batch share the allocation)
    int batch = 1;
    int ncomp = 0;
    int l = 0;          // active Q limbs; GetLevel() = L - l
    int deg = 1;        // noiseScaleDeg
    double scale = 0;   // scalingFactor
    int slots = 0;
    u64* data() const { return mem->p + off; }
    size_t words_each(int N) const { return (size_t)ncomp * l * N; }
    bool valid() const { return (bool)mem; }
};

struct SplitMix {
    u64 s;
    explicit SplitMix(u64 seed) : s(seed) {}
    u64 next() { u64 z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    static u64 sub(u64 seed, u64 tag) { SplitMix t(seed + tag * 0x9E3779B97F4A7C15ull); return t.next(); }
};

struct BootPrecomp;   // bootstrap.cpp

// A slots x slots matrix given by its (generalised) diagonals, prepared for the baby-step/giant-step evaluation
// (M v)[p] = sum_d diag_d[p] v[(p + d) mod slots]  ==  sum_j Rot_{G_j}( sum_i P_{j,i} * Rot_{g i}(v) ),  P_{j,i} = Rot_{-G_j}(diag).
// This is the ct x pt matrix product of CoeffsToSlots / SlotsToCoeffs and of packed linear layers (lintrans.cpp).
struct LinTrans {
    int slots = 0, g = 1, n1 = 1, n2 = 1, off = 0, cnt = 0;
    std::vector<int> giant_rot;                          // rotation of giant step j; rotating steps first, the identity (0) last
    std::vector<std::vector<std::vector<cplx>>> host;    // [j][i] pre-rotated diagonals, empty = absent
    std::vector<uint32_t> mask;
    int ndiag = 0;
    int level = -1;                                      // level the device plaintexts are encoded at (-1: not yet)
    double pt_scale = 0;
    Mem pts;                                             // [n2][n1][l+K][N] over Q_l u P, evaluation form
};

class Scheme {
public:
    explicit Scheme(const ParamSpec& spec, int device = -1);
    ~Scheme();
    Engine eng;
    const Params& P;
    int L() const { return P.L; }

    // ---- keys (KeyGen F.cpp:47, EvalMultKeyGen :49, EvalRotateKeyGen :248, EvalBootstrapKeyGen :239) ----
    void keygen(u64 seed);
    void gen_mult_key();
    void gen_rotation_key(int k);
    void gen_galois_key(uint32_t g);
    bool has_galois_key(uint32_t g) const { return gk_.count(g) != 0; }
    void clear_rotation_keys();                       // ClearEvalAutomorphismKeys F.cpp:336
    void clear_mult_key();                            // ClearEvalMultKeys F.cpp:342
    size_t num_galois_keys() const { return gk_.size(); }
    // raw import/export (device <-> host), used by save/load and by the parity tests
    void export_sk(u64* out) const;                   // (L+K) limbs eval
    void export_pk(u64* out) const;                   // [2][L][N]
    void export_evk(uint32_t g, u64* out) const;      // g = 0: mult key
    void import_keys(const u64* sk, const u64* pk);
    void import_evk(uint32_t g, const u64* evk);

    // ---- encode / encrypt / decrypt (MakeCKKSPackedPlaintext F.cpp:353, Encrypt :380, Decrypt :389) ----
    Elem encode(const cplx* vals, int n, int level, int slots, int deg = 1);
    Elem encrypt_values_many(const double* vals, int B, int n, int level, int slots);   // encode + encrypt, one batched ciphertext
    Elem encode_many_real(const double* vals, int B, int n, int level, int slots);   // one batched plaintext
    Elem encode_real(const double* vals, int n, int level, int slots);
    Elem encode_at(const cplx* vals, int n, int l, double scale, int slots, int deg);   // explicit limb count / scale
    Elem encrypt(const Elem& pt);
    Elem encrypt_seeded(const Elem& pt, u64 seed);
    Elem encrypt_many(const std::vector<const Elem*>& pts);   // one batched operand holding Encrypt(pt_b), b < pts.size()
    void decrypt(const Elem& ct, cplx* out, int slots);
    void decode(const Elem& pt, cplx* out, int slots);

    // ---- leveled ops with FLEXIBLEAUTO bookkeeping ----
    Elem add(const Elem& a, const Elem& b);           // EvalAdd F.cpp:410,414 (ct+ct, ct+pt)
    Elem sub(const Elem& a, const Elem& b);
    Elem add_many(std::vector<Elem> v);               // EvalAddMany F.cpp:418,1067
    Elem add_const(const Elem& a, double c);
    Elem mult(const Elem& a, const Elem& b);          // EvalMult F.cpp:427 (ct*pt), :431 (ct*ct + relinearisation)
    Elem mult_const(const Elem& a, double c);         // EvalMult(ct, double)
    Elem mult_many(std::vector<Elem> v);              // EvalMultMany F.cpp:1297
    // EvalLinearWSum: out[o] = sum_t w[o * n_in + t] * in[t] for a batched operand `in` (n_in ciphertexts); result is a batch
    // of n_out ciphertexts one degree deeper.  The encrypted Linformer E / F projection is one such call.
    Elem linear_wsum(const Elem& in, const double* w, int n_out);
    Elem weighted_sum(const std::vector<Elem>& terms, const std::vector<double>& w);   // aligned terms, one kernel
    Elem weighted_sums(const std::vector<Elem>& terms, const std::vector<std::vector<double>>& w);   // several combinations, one kernel
    Elem square(const Elem& a) { return mult(a, a); }
    Elem rotate(const Elem& a, int k);                // EvalRotate F.cpp:435,833,843
    Elem conjugate(const Elem& a);
    Elem rotsum(const Elem& a, int steps, int stride);   // FHEController::rotsum / repeat ladders F.cpp:829-867
    static std::vector<int> ladder_plan(int steps);                   // doubling steps per hoisted key switch
    static std::vector<int> ladder_rotations(int steps, int stride);  // rotation keys rotsum(steps, stride) uses with that plan
    // BSGS diagonal linear transforms (EvalLinearTransform of CoeffsToSlots / SlotsToCoeffs; packed ct x pt matrix products)
    LinTrans lintrans_plan(const std::map<int, std::vector<cplx>>& diags, int slots, int max_baby = 0);
    std::vector<int> lintrans_rotations(const LinTrans& t) const;
    void lintrans_encode(LinTrans& t, int level);
    Elem lintrans_apply(LinTrans& t, const Elem& ct);          // result: deg + 1, scale = ct.scale * sf[level]
    Elem lintrans_apply_plain(LinTrans& t, const Elem& ct);    // the same transform, one EvalRotate / EvalMult / EvalAdd at a time (checker)
    Elem apply_galois(const Elem& a, uint32_t g);
    Elem clone(const Elem& a);                        // Ciphertext::Clone M:223
    Elem pack(const std::vector<Elem>& v);            // gather ciphertexts of identical level / scale into one batched operand
    Elem slice(const Elem& a, int i) const;           // zero-copy view of element i of a batched operand
    Elem range(const Elem& a, int first, int count) const;   // ... of elements first .. first + count - 1
    int max_batch(int l) const;                       // batch size that keeps the key-switch workspace within budget
    Elem rescaled(const Elem& a);                     // ModReduceInternal
    void rescale_inplace(Elem& a);
    void level_reduce_inplace(Elem& a, int levels);   // LevelReduceInternal
    void mult_scalar_core(Elem& a, double c);         // EvalMultCoreInPlace(ct, double): x round(c * sf[level]), deg+1
    void mult_int_inplace(Elem& a, i128 k);           // multiply by an integer, metadata unchanged
    void adjust_pair(Elem& a, Elem& b);               // AdjustLevelsAndDepthInPlace
    void adjust_pair_to_one(Elem& a, Elem& b);        // AdjustLevelsAndDepthToOneInPlace
    void drop_to(Elem& a, int l);                     // keep the first l limbs

    // ---- polynomial evaluation (EvalPoly F.cpp:1291, EvalChebyshevFunction F.cpp:486,1319-1335) ----
    Elem eval_poly(const Elem& x, const std::vector<double>& coeffs);
    Elem eval_chebyshev(const Elem& x, const std::vector<double>& coeffs, double a, double b);
    static std::vector<double> chebyshev_coefficients(double (*f)(double, void*), void* user, double a, double b, int degree);

    // ---- bootstrapping (EvalBootstrapSetup F.cpp:238,280; EvalBootstrapKeyGen :239; EvalBootstrap :445) ----
    void bootstrap_setup(int budget_cts, int budget_stc, int slots);
    void bootstrap_keygen(int slots);
    std::vector<int> bootstrap_rotations(int slots);
    Elem bootstrap(const Elem& ct);
    Elem bootstrap_iter(const Elem& ct, int iterations, int precision);   // EvalBootstrap(c, 2, precision) F.cpp:461
    bool has_rotation_key(int k) const { return k == 0 || gk_.count(P.galois_for_rotation(k)) != 0; }

    // ---- serialisation (stands in for Serial::SerializeToFile / DeserializeFromFile, F.cpp:59-89,1360-1394) ----
    Elem import_elem(const u64* host, int ncomp, int l, int deg, double scale, int slots);
    void save_elem(const Elem& a, const char* path);
    Elem load_elem(const char* path);
    void save_keys(const char* path, int what = 15);
    void load_keys(const char* path);

    int level_of(const Elem& a) const { return P.L - a.l; }

private:
    friend struct BootPrecomp;
    Elem make(int ncomp, int l, int deg, double scale, int slots, int batch = 1);
    std::vector<Elem> split_run(const Elem& a, const std::function<Elem(const Elem&)>& f);
    void keyswitch_gen(const u64* sk_old_dev, const u64* sk_new_dev, u64 seed, u64* evk_dev);
    void sample_to_eval(u64* dst, const std::vector<int8_t>& s, const LimbSel& sel);
    // randomness: what a stream is for (independent ChaCha20 keys) and, for the seeded test entries, its SplitMix64 seed
    enum class Use { Secret = 0, Public = 1, Error = 2, Encrypt = 3 };
    struct Draw { bool seeded; u64 seed; Use use; };
    Draw draw(Use use, u64 seed) const { return Draw{seeded_keys_, seed, use}; }
    void uniform_to_dev(u64* dst, const Draw& d, const LimbSel& sel);
    void sample_dev_to_eval(u64* dst, const Draw& d, int kind, const LimbSel& sel);   // kind 0 ternary, 1 Gaussian (device sampler + NTT)
    Elem encrypt_with(const Elem& pt, bool seeded, u64 seed);
    struct DevFft { uint32_t* rot; double* cre; double* cim; };               // special-FFT tables of one slot count, on the device
    DevFft& dev_fft(int slots);
    double* stage_slot(int& slot);                                   // next free slot of the pinned staging ring
    void upload_small(u64* dst, const u64* src, size_t words);        // host -> device through the ring, no synchronisation
    void encode_coeffs(const cplx* vals, int n, int slots, double scale, std::vector<i128>& co) const;
    void coeffs_to_dev(u64* dst, const std::vector<i128>& co, int l);
    ScalarSet scalar_set(i128 k, int l) const;
    Elem binary(const Elem& a, const Elem& b, bool subtract);
    // evaluation helpers
    Elem cheby_ps(const Elem& x, const std::vector<double>& c);
    std::vector<Elem> add_each(const std::vector<Elem>& a, const std::vector<Elem>& b);   // independent sums, batched per shape
    Elem part_of(const Elem& all, size_t g, int each) const;
    std::vector<Elem> mult_each(std::vector<Elem> a, std::vector<Elem> b, bool align);   // independent products as one batched EvalMult
    std::vector<Elem> inner_linear_many(const std::vector<Elem>& T, const std::vector<const std::vector<double>*>& polys, int kmax);
    std::vector<Elem> ps_settled_;                          // inner_linear caches, valid during one cheby_ps evaluation
    std::map<std::pair<int, int>, Elem> ps_aligned_;

    std::map<int, DevFft> dev_fft_;
    static constexpr int kStageSlots = 8;
    double* stage_ = nullptr;                 // pinned staging ring for slot values
    cudaEvent_t stage_ev_[kStageSlots] = {};
    int stage_next_ = 0;
    ChaChaKey rng_keys_[4];                   // per Use, from getrandom(); re-drawn by every unseeded keygen
    u64 rng_stream_ = 0;                      // next unused ChaCha20 stream number (never reused under one key)
    bool seeded_keys_ = false;                // keys of this context came from fl_keygen_seeded (tests only)
    u64 key_seed_ = 0;
    u64* sk_ = nullptr;      // (L+K) limbs eval
    u64* pk_ = nullptr;      // [2][L][N]
    u64* mk_ = nullptr;      // relinearisation key
    std::map<uint32_t, u64*> gk_;
    std::map<int, std::shared_ptr<BootPrecomp>> boot_;
};

}  // namespace flk
