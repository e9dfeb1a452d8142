// FHEController.cpp -- the reference's controller (src/FHEController.cpp) re-backed by the B200 CKKS engine.
//
// Same public methods, argument meaning, slot layouts and level bookkeeping as the reference class, so that
// src/main.cpp (encoder1 / pooler / classifier, main.cpp:145-475) runs on it unchanged; every OpenFHE call of the
// reference (`context->...`) becomes one call of the C-ABI in include/fl_ckks.h, which launches sm_100a kernels.
// There is no CPU evaluation path in here: without the CUDA library nothing below can run.
//
// Differences that do not change results: mask plaintexts are encoded once per (pattern, value, level) and cached
// (the reference re-encodes a 16384-slot vector for every mask call, F.cpp:1207-1286); a rotation whose key was never
// generated is generated on first use when the secret key is resident (the reference's key list at main.cpp:84 misses
// -8, -16 and -512); key / ciphertext files use the engine's own container (DESIGN.md "Files").
#include "FHEController.h"

#include <set>

#include <cstdlib>
#include <cstring>

namespace {

inline void need(int rc, const char* what) {
    if (rc) lbcrypto::fl_fail(what);
}
// the reference's behaviour on unreadable key material: message on cerr, exit(1)  (F.cpp:62-70,192-220,286-294)
[[noreturn]] void die(const std::string& msg) {
    std::cerr << msg << std::endl;
    std::exit(1);
}
bool file_exists(const std::string& p) { return std::ifstream(p).good(); }

struct ContextFile {   // crypto-context.txt: the CCParams needed to rebuild the context
    char magic[8];
    int32_t logN, L, dnum, first_bits, scale_bits, aux_bits, sparse_h, budget0, budget1, depth;
    uint64_t reserved;   // was a key seed in round 1: never written any more (a seed next to the public key gives the secret key away)
};

double cheb_trampoline(double x, void* user) { return (*static_cast<const std::function<double(double)>*>(user))(x); }

// FHECKKSRNS::GetBootstrapDepth(approxModDepth, levelBudget, SPARSE_TERNARY) = approxModDepth + budget[0] + budget[1]
int bootstrap_depth(int approx_mod_depth, const vector<uint32_t>& budget) { return approx_mod_depth + (int)budget[0] + (int)budget[1]; }

vector<fl_elem*> handles(const vector<Ctxt>& v) {
    vector<fl_elem*> h(v.size());
    for (size_t i = 0; i < v.size(); ++i) h[i] = v[i]->handle();
    return h;
}

}  // namespace

FHEController::~FHEController() {
    mask_cache_.clear();
    // ciphertexts handed out may outlive the controller (main.cpp keeps globals); the context is released at exit
}

void FHEController::release_context() {
    mask_cache_.clear();
    for (auto& kv : packed_) fl_lt_free(kv.second);
    packed_.clear();
    if (ctx_) { fl_ctx_destroy(ctx_); ctx_ = nullptr; }
}

string FHEController::key_path(const string& name) const {
    const char* root = std::getenv("FHE_LINFORMER_ROOT");   // default: the reference's "../" relative layout
    return string(root ? root : "..") + "/" + parameters_folder + "/" + name;
}

/* ------------------------------------------------------------------ context ------------------------------------------------------------------ */

void FHEController::create(int log_ring, int depth, int digits, int first_bits, int scale_bits) {
    if (ctx_) { mask_cache_.clear(); fl_ctx_destroy(ctx_); ctx_ = nullptr; }
    params_.logN = log_ring;
    params_.L = depth + 1;   // FLEXIBLEAUTO: multiplicative depth + 1 limbs
    params_.dnum = digits;
    params_.first_bits = first_bits;
    params_.scale_bits = scale_bits;
    params_.aux_bits = 60;
    params_.sparse_h = 192;  // SPARSE_TERNARY
    if (const char* ov = std::getenv("FHE_LINFORMER_LOGN")) params_.logN = std::atoi(ov);   // test / bench override only
    if (const char* ov = std::getenv("FLK_MAX_ROWS")) max_rows_per_batch = std::max(1, std::atoi(ov));
    need(fl_ctx_create(&params_, device, &ctx_), "GenCryptoContext");
    if (cache_gb > 0) need(fl_ctx_set_cache_bytes(ctx_, (uint64_t)(cache_gb * 1073741824.0)), "cache cap");
    if (const char* ov = std::getenv("FLK_AUTO_ROTATION_KEYS")) auto_rotation_keys = ov[0] == '1';
}

// reference: F.cpp:3-90
void FHEController::generate_context(bool serialize, bool secure) {
    (void)secure;   // ignored by the reference as well (HEStd_NotSet either way)
    num_slots = 1 << 14;
    level_budget = {3, 3};
    const int levels_before_bootstrap = 12;
    circuit_depth = 1 + levels_before_bootstrap + bootstrap_depth(8, level_budget);
    cout << endl << "Ciphertexts depth: " << circuit_depth << ", available multiplications: " << levels_before_bootstrap - 2 << endl;
    create(15, circuit_depth, 4, 55, 52);
    cout << "Context built, generating keys..." << endl;
    need(fl_keygen(ctx_, key_seed), "KeyGen");
    need(fl_gen_mult_key(ctx_), "EvalMultKeyGen");
    cout << "Generated." << endl;
    if (serialize) serialize_context();
}

// reference: F.cpp:92-182 (explicit parameters; nobody calls it, kept for interface parity)
void FHEController::generate_context(int log_ring, int log_scale, int log_primes, int digits_hks, int cts_levels, int stc_levels, int relu_deg,
                                     bool serialize) {
    num_slots = 1 << 14;
    level_budget = {(uint32_t)cts_levels, (uint32_t)stc_levels};
    relu_degree = relu_deg;
    const int levels_before_bootstrap = get_relu_depth(relu_deg) + 3;
    circuit_depth = levels_before_bootstrap + bootstrap_depth(8, level_budget);
    cout << endl << "Ciphertexts depth: " << circuit_depth << ", available multiplications: " << levels_before_bootstrap - 2 << endl;
    create(log_ring, circuit_depth, digits_hks, log_primes, log_scale);
    need(fl_keygen(ctx_, key_seed), "KeyGen");
    need(fl_gen_mult_key(ctx_), "EvalMultKeyGen");
    if (serialize) serialize_context();
}

void FHEController::serialize_context() {
    cout << "Now serializing keys ..." << endl;
    if (fl_keys_save_sel(ctx_, key_path("mult-keys.txt").c_str(), 4)) die("Error serializing EvalMult keys in \"" + key_path("mult-keys.txt") + "\"");
    cout << "Relinearization Keys have been serialized" << endl;
    ContextFile cf{};
    std::memcpy(cf.magic, "FLCKCTX", 8);
    cf.logN = params_.logN; cf.L = params_.L; cf.dnum = params_.dnum; cf.first_bits = params_.first_bits; cf.scale_bits = params_.scale_bits;
    cf.aux_bits = params_.aux_bits; cf.sparse_h = params_.sparse_h; cf.budget0 = (int)level_budget[0]; cf.budget1 = (int)level_budget[1];
    cf.depth = circuit_depth; cf.reserved = 0;
    std::ofstream out(key_path("crypto-context.txt"), ios::out | ios::binary);
    if (out.write(reinterpret_cast<const char*>(&cf), sizeof cf)) cout << "Crypto Context have been serialized" << endl;
    else cerr << "Error writing serialization of the crypto context to crypto-context.txt" << endl;
    if (fl_keys_save_sel(ctx_, key_path("public-key.txt").c_str(), 2)) cerr << "Error writing serialization of public key to public-key.txt" << endl;
    else cout << "Public Key has been serialized" << endl;
    if (fl_keys_save_sel(ctx_, key_path("secret-key.txt").c_str(), 1)) cerr << "Error writing serialization of secret key to secret-key.txt" << endl;
    else cout << "Secret Key has been serialized" << endl;
}

// reference: F.cpp:184-235
void FHEController::load_context(bool verbose) {
    if (verbose) cout << "Reading serialized context..." << endl;
    ContextFile cf{};
    std::ifstream in(key_path("crypto-context.txt"), ios::in | ios::binary);
    if (!in.read(reinterpret_cast<char*>(&cf), sizeof cf) || std::memcmp(cf.magic, "FLCKCTX", 8))
        die("I cannot read serialized data from: " + key_path("crypto-context.txt"));
    if (ctx_) { mask_cache_.clear(); fl_ctx_destroy(ctx_); ctx_ = nullptr; }
    params_.logN = cf.logN; params_.L = cf.L; params_.dnum = cf.dnum; params_.first_bits = cf.first_bits; params_.scale_bits = cf.scale_bits;
    params_.aux_bits = cf.aux_bits; params_.sparse_h = cf.sparse_h;
    need(fl_ctx_create(&params_, device, &ctx_), "DeserializeFromFile(crypto-context)");
    if (cache_gb > 0) need(fl_ctx_set_cache_bytes(ctx_, (uint64_t)(cache_gb * 1073741824.0)), "cache cap");
    if (fl_keys_load(ctx_, key_path("public-key.txt").c_str())) die("I cannot read serialized data from public-key.txt");
    if (fl_keys_load(ctx_, key_path("secret-key.txt").c_str())) die("I cannot read serialized data from secret-key.txt");
    if (!file_exists(key_path("mult-keys.txt"))) die("Cannot read serialization from mult-keys.txt");
    if (fl_keys_load(ctx_, key_path("mult-keys.txt").c_str())) die("Could not deserialize eval mult key file");
    level_budget = {3, 3};
    if (verbose) cout << "CtoS: " << level_budget[0] << ", StoC: " << level_budget[1] << endl;
    const int levels_before_bootstrap = 12;
    circuit_depth = levels_before_bootstrap + bootstrap_depth(8, level_budget);   // 26 here vs 27 at generation, as in the reference
    if (verbose) cout << "Circuit depth: " << circuit_depth << ", available multiplications: " << levels_before_bootstrap - 2 << endl;
    num_slots = 1 << 14;
}

/* ------------------------------------------------------------------ keys ------------------------------------------------------------------ */

void FHEController::generate_bootstrapping_keys(int bootstrap_slots) {
    need(fl_bootstrap_setup(ctx_, (int)level_budget[0], (int)level_budget[1], bootstrap_slots), "EvalBootstrapSetup");
    need(fl_bootstrap_keygen(ctx_, bootstrap_slots), "EvalBootstrapKeyGen");
}

void FHEController::generate_rotation_keys(vector<int> rotations, bool serialize, string filename) {
    if (serialize && filename.empty()) {
        cout << "Filename cannot be empty when serializing rotation keys." << endl;
        return;
    }
    need(fl_gen_rot_keys(ctx_, rotations.data(), (int)rotations.size()), "EvalRotateKeyGen");
    // The batched recipes of this backend use more indices than the reference's sequential loops: the hoisted ladder groups
    // (multiples of a ladder's stride, fl_rotsum_rotations) and the binary trees that replace rotate(., -1) / rotate(., -512)
    // chains.  They are derived from the list given here and generated NOW, so that no key is ever made from the secret key
    // in the middle of an evaluation (and none inside a timed region); rotate() on a missing index fails as OpenFHE does.
    const vector<int> extra = derived_rotations(rotations);
    if (!extra.empty()) need(fl_gen_rot_keys(ctx_, extra.data(), (int)extra.size()), "EvalRotateKeyGen");
    if (!serialize) return;
    if (fl_keys_save_sel(ctx_, key_path("rot_" + filename).c_str(), 8)) die("Error serializing Rotation keys" + key_path("rot_" + filename));
    cout << "Rotation keys \"" << filename << "\" have been serialized" << endl;
}

vector<int> FHEController::derived_rotations(const vector<int>& listed) const {
    std::set<int> have(listed.begin(), listed.end()), out;
    auto chain = [&](int stride, int steps) {   // stride * 2^i listed for every doubling step?
        for (int i = 0; i < steps; ++i) if (!have.count(stride * (1 << i))) return false;
        return true;
    };
    if (hoist_ladders) {
        // every ladder F.cpp:829-867 is called with on this circuit (rotsum 128/64/32 x stride 128 or 1, repeat 128/64 x -1, -128)
        static const int ladders[][2] = {{128, 128}, {128, 1}, {64, 1}, {32, 128}, {128, -1}, {64, -1}, {128, -128}, {64, 128}, {32, 1}};
        for (auto& ld : ladders) {
            int steps = 0;
            while ((1 << steps) < ld[0]) ++steps;
            if (!chain(ld[1], steps)) continue;
            int rots[64];
            const int nr = fl_rotsum_rotations(steps, ld[1], rots, 64);
            for (int i = 0; i < nr && i < 64; ++i) if (!have.count(rots[i])) out.insert(rots[i]);
        }
    }
    if (batch_rows) {   // shifted_sum trees: stride * 2^b for up to 256 rows (stride -1) / 32 tokens per container (stride -512)
        if (have.count(-1)) for (int b = 1; b < 8; ++b) if (!have.count(-(1 << b))) out.insert(-(1 << b));
        if (have.count(-512)) for (int b = 1; b < 5; ++b) if (!have.count(-512 * (1 << b))) out.insert(-512 * (1 << b));
    }
    return vector<int>(out.begin(), out.end());
}

/* ------------------------------------------------------------------ packed linear layers ------------------------------------------------------------------ */
// In the wrapped-expanded layout a 128 x 128 weight acts on the block index: y[128 i + t] = sum_j W[j][i] x[128 j + t].  With
// j = (i + k) mod 128 the source slot is (128 i + t) + 128 k modulo 16384 -- the slot vector wraps exactly where the block index
// does -- so the layer is the sum over k < 128 of diag_k (x) rot(x, 128 k) with diag_k[128 i + t] = W[(i + k) mod 128][i]:
// a diagonal matrix product, evaluated by Engine::linear_transform with 16 baby x 8 giant steps (double hoisting).
namespace {
fl_lt* plan_packed(fl_ctx* ctx, int slots, const std::function<double(int, int)>& weight, double scale, const vector<double>* column_scale = nullptr) {
    const int D = 128;
    vector<int> shifts(D);
    vector<double> re((size_t)D * slots, 0.0);
    for (int k = 0; k < D; ++k) {
        shifts[k] = D * k;
        double* d = re.data() + (size_t)k * slots;
        for (int i = 0; i < D; ++i) {
            const double v = scale * weight((i + k) % D, i);
            for (int t = 0; t < D; ++t) d[D * i + t] = column_scale ? v * (*column_scale)[(size_t)t] : v;
        }
    }
    fl_lt* lt = nullptr;
    if (fl_lt_create(ctx, shifts.data(), D, re.data(), nullptr, slots, -1, 0, &lt)) lbcrypto::fl_fail("packed linear layer: plan");
    return lt;
}
}  // namespace

Ctxt FHEController::packed_linear(const Ctxt& x, const string& name, const std::function<double(int, int)>& weight, double scale,
                                   const vector<double>* column_scale) {
    if (num_slots != 128 * 128) throw std::invalid_argument("packed_linear: the wrapped-expanded layout needs 128 x 128 slots");
    if (column_scale && column_scale->size() < 128) throw std::invalid_argument("packed_linear: column_scale needs 128 values");
    auto it = packed_.find(name);
    if (it == packed_.end()) it = packed_.emplace(name, plan_packed(ctx_, num_slots, weight, scale, column_scale)).first;
    int rots[64];
    const int nr = fl_lt_rotations(ctx_, it->second, rots, 64);
    for (int i = 0; i < nr && i < 64; ++i) require_rotation_key(rots[i]);
    fl_elem* e = nullptr;
    need(fl_lt_apply(ctx_, it->second, x->handle(), &e), "packed linear layer");
    return wrap(e);
}

void FHEController::generate_packed_keys() {
    // the rotations depend on the diagonal pattern only (shifts 128 k, k < 128), not on the weights
    fl_lt* probe = plan_packed(ctx_, num_slots, [](int, int) { return 1.0; }, 1.0);
    int rots[64];
    const int nr = fl_lt_rotations(ctx_, probe, rots, 64);
    fl_lt_free(probe);
    need(fl_gen_rot_keys(ctx_, rots, std::min(nr, 64)), "EvalRotateKeyGen (packed layers)");
}

double FHEController::rotation_key_bytes() const {
    return fl_rot_key_bytes(ctx_);
}

void FHEController::require_rotation_key(int index) {
    if (fl_has_rot_key(ctx_, index)) return;
    if (!auto_rotation_keys)
        throw std::runtime_error("EvalRotate: no evaluation key for rotation index " + std::to_string(index) +
                                 " (EvalRotateKeyGen was not called for it; set auto_rotation_keys to generate keys on demand)");
    need(fl_gen_rot_keys(ctx_, &index, 1), "EvalRotateKeyGen");
}

void FHEController::generate_bootstrapping_and_rotation_keys(vector<int> rotations, int bootstrap_slots, bool serialize, const string& filename) {
    if (serialize && filename.empty()) {
        cout << "Filename cannot be empty when serializing bootstrapping and rotation keys." << endl;
        return;
    }
    generate_bootstrapping_keys(bootstrap_slots);
    generate_rotation_keys(rotations, serialize, filename);
}

void FHEController::load_bootstrapping_and_rotation_keys(const string& filename, int bootstrap_slots, bool verbose) {
    if (verbose) cout << endl << "Loading bootstrapping and rotations keys from " << filename << "..." << endl;
    auto start = start_time();
    need(fl_bootstrap_setup(ctx_, (int)level_budget[0], (int)level_budget[1], bootstrap_slots), "EvalBootstrapSetup");
    if (verbose) cout << "(1/2) Bootstrapping precomputations completed!" << endl;
    if (!file_exists(key_path("rot_" + filename))) die("Cannot read serialization from " + key_path("rot_" + filename));
    if (fl_keys_load(ctx_, key_path("rot_" + filename).c_str())) die("Could not deserialize eval rot key file");
    if (verbose) {
        cout << "(2/2) Rotation keys read!" << endl;
        print_duration(start, "Loading bootstrapping pre-computations + rotations");
        cout << endl;
    }
}

void FHEController::load_rotation_keys(const string& filename, bool verbose) {
    if (verbose) cout << endl << "Loading rotations keys from " << filename << "..." << endl;
    auto start = start_time();
    if (!file_exists(key_path("rot_" + filename))) die("Cannot read serialization from " + key_path("rot_" + filename));
    if (fl_keys_load(ctx_, key_path("rot_" + filename).c_str())) die("Could not deserialize eval rot key file");
    if (verbose) {
        cout << "(1/1) Rotation keys read!" << endl;
        print_duration(start, "Loading rotation keys");
        cout << endl;
    }
}

void FHEController::clear_bootstrapping_and_rotation_keys(int) { clear_rotation_keys(); }
void FHEController::clear_rotation_keys() { need(fl_keys_clear(ctx_, 0), "ClearEvalAutomorphismKeys"); }
void FHEController::clear_context(int bootstrapping_key_slots) {
    if (bootstrapping_key_slots != 0) clear_bootstrapping_and_rotation_keys(bootstrapping_key_slots);
    else clear_rotation_keys();
    need(fl_keys_clear(ctx_, 1), "ClearEvalMultKeys");
}

/* ------------------------------------------------------------------ encode / encrypt / decrypt ------------------------------------------------------------------ */

Ptxt FHEController::encode(const vector<double>& vec, int level, int plaintext_num_slots) {
    if (plaintext_num_slots == 0) plaintext_num_slots = num_slots;
    fl_elem* e = nullptr;
    need(fl_encode(ctx_, vec.data(), nullptr, (int)std::min<size_t>(vec.size(), (size_t)plaintext_num_slots), level, plaintext_num_slots, &e),
         "MakeCKKSPackedPlaintext");
    Ptxt p = wrap_pt(e);
    p->SetLength(plaintext_num_slots);
    return p;
}

Ptxt FHEController::encode(double val, int level, int plaintext_num_slots) {
    if (plaintext_num_slots == 0) plaintext_num_slots = num_slots;
    return encode(vector<double>((size_t)plaintext_num_slots, val), level, plaintext_num_slots);
}

Ctxt FHEController::encrypt(const vector<double>& vec, int level, int plaintext_num_slots) {
    if (plaintext_num_slots == 0) plaintext_num_slots = num_slots;
    return encrypt_ptxt(encode(vec, level, plaintext_num_slots));
}

Ctxt FHEController::encrypt_ptxt(const Ptxt& p) {
    fl_elem* e = nullptr;
    need(fl_encrypt(ctx_, p->handle(), &e), "Encrypt");
    return wrap(e);
}

Ptxt FHEController::decrypt(const Ctxt& c) {
    const int n = (int)c->GetSlots();
    vector<double> re(n), im(n);
    need(fl_decrypt(ctx_, c->handle(), re.data(), im.data(), n), "Decrypt");
    vector<std::complex<double>> v(n);
    for (int i = 0; i < n; ++i) v[i] = {re[i], im[i]};
    return std::make_shared<PlaintextImpl>(std::move(v), (int)c->GetLevel());
}

vector<double> FHEController::decrypt_tovector(const Ctxt& c, int slots) {
    if (slots == 0) slots = num_slots;
    Ptxt p = decrypt(c);
    p->SetSlots(slots);
    p->SetLength(slots);
    return p->GetRealPackedValue();
}

/* ------------------------------------------------------------------ homomorphic operations ------------------------------------------------------------------ */

Ctxt FHEController::add(const Ctxt& c1, const Ctxt& c2) {
    fl_elem* e = nullptr;
    need(fl_add(ctx_, c1->handle(), c2->handle(), &e), "EvalAdd");
    return wrap(e);
}
Ctxt FHEController::add(const Ctxt& c1, const Ptxt& c2) {
    fl_elem* e = nullptr;
    need(fl_add(ctx_, c1->handle(), c2->handle(), &e), "EvalAdd");
    return wrap(e);
}
Ctxt FHEController::add(vector<Ctxt> c) {
    auto h = handles(c);
    fl_elem* e = nullptr;
    need(fl_add_many(ctx_, h.data(), (int)h.size(), &e), "EvalAddMany");
    return wrap(e);
}
Ctxt FHEController::mult(const Ctxt& c1, double d) {
    // reference: encode(d) at the ciphertext's level, then EvalMult(ct, pt)  (F.cpp:421-424)
    return mult(c1, mask_plain(4, 0, 0, d, (int)c1->GetLevel()));
}
Ctxt FHEController::mult(const Ctxt& c, const Ptxt& p) {
    fl_elem* e = nullptr;
    need(fl_mul(ctx_, c->handle(), p->handle(), &e), "EvalMult");
    return wrap(e);
}
Ctxt FHEController::mult(const Ctxt& c1, const Ctxt& c2) {
    fl_elem* e = nullptr;
    need(fl_mul(ctx_, c1->handle(), c2->handle(), &e), "EvalMult");
    return wrap(e);
}
Ctxt FHEController::rotate(const Ctxt& c, int index) {
    require_rotation_key(index);   // missing key: an error, as from OpenFHE's EvalRotate (unless auto_rotation_keys)
    fl_elem* e = nullptr;
    need(fl_rotate(ctx_, c->handle(), index, &e), "EvalRotate");
    return wrap(e);
}

Ctxt FHEController::bootstrap(const Ctxt& c, bool timing) {
    auto start = start_time();
    fl_elem* e = nullptr;
    need(fl_bootstrap(ctx_, c->handle(), &e), "EvalBootstrap");
    if (timing) {
        need(fl_sync(ctx_), "sync");
        print_duration(start, "Bootstrapping " + to_string(c->GetSlots()) + " slots");
    }
    return wrap(e);
}

Ctxt FHEController::bootstrap(const Ctxt& c, int precision, bool timing) {
    if (static_cast<int>(c->GetLevel()) + 2 < circuit_depth)
        cout << "You are bootstrapping with remaining levels! You are at " << to_string(c->GetLevel()) << "/" << circuit_depth - 2 << endl;
    auto start = start_time();
    fl_elem* e = nullptr;
    need(fl_bootstrap_iter(ctx_, c->handle(), 2, precision, &e), "EvalBootstrap");
    if (timing) {
        need(fl_sync(ctx_), "sync");
        print_duration(start, "Double Bootstrapping " + to_string(c->GetSlots()) + " slots");
    }
    return wrap(e);
}

Ctxt FHEController::chebyshev(const std::function<double(double)>& f, const Ctxt& c, double a, double b, int degree) {
    vector<double> coeffs((size_t)degree + 1);
    need(fl_chebyshev_coefficients(&cheb_trampoline, const_cast<std::function<double(double)>*>(&f), a, b, degree, coeffs.data()),
         "EvalChebyshevCoefficients");
    fl_elem* e = nullptr;
    need(fl_eval_chebyshev(ctx_, c->handle(), coeffs.data(), (int)coeffs.size(), a, b, &e), "EvalChebyshevFunction");
    return wrap(e);
}

Ctxt FHEController::relu(const Ctxt& c, double scale, bool timing) {
    auto start = start_time();
    Ctxt res = chebyshev([scale](double x) { return x < 0 ? 0.0 : x / scale; }, c, -1, 1, relu_degree);
    if (timing) {
        need(fl_sync(ctx_), "sync");
        print_duration(start, "ReLU d = " + to_string(relu_degree) + " evaluation");
    }
    return res;
}

/* ------------------------------------------------------------------ text-file readers ------------------------------------------------------------------ */
// Layouts: plain = the file as is; repeated R(v): slot[128 j + i] = v[i]; expanded E(v): slot[128 j + i] = v[j].

namespace {
vector<double> tile(const vector<double>& v, int period, int copies, double scale) {
    vector<double> out((size_t)period * copies);
    for (int j = 0; j < copies; ++j)
        for (int i = 0; i < period; ++i) out[(size_t)j * period + i] = v.at(i) * scale;
    return out;
}
vector<double> stretch(const vector<double>& v, int rows, int width, int filled, double scale) {
    vector<double> out((size_t)rows * width, 0.0);
    for (int j = 0; j < rows; ++j)
        for (int i = 0; i < filled; ++i) out[(size_t)j * width + i] = v.at(j) * scale;
    return out;
}
}  // namespace

Ctxt FHEController::read_input(const string& filename, double scale) {
    return encrypt(read_values_from_file(filename, scale), circuit_depth - 10, num_slots);
}
Ptxt FHEController::read_plain_input(const string& filename, int level, double scale) {
    return encode(read_values_from_file(filename, scale), level, num_slots);
}
vector<Ptxt> FHEController::read_plain_256_input(const string& filename, int level, double scale) {
    const vector<double> v = read_values_from_file(filename, scale);
    vector<Ptxt> halves;
    for (int h = 0; h < 2; ++h) halves.push_back(encode(vector<double>(v.begin() + 128 * h, v.begin() + 128 * (h + 1)), level, num_slots));
    return halves;
}
Ctxt FHEController::read_repeated_input(const string& filename, double scale) {
    // the reference builds the repeated vector but encrypts the raw (scaled) input (F.cpp:559-580); reproduced as is
    return encrypt(read_values_from_file(filename, scale), 0, num_slots);
}
Ptxt FHEController::read_plain_repeated_input(const string& filename, int level, double scale) {
    return encode(tile(read_values_from_file(filename), 128, 128, scale), level, num_slots);
}
Ptxt FHEController::read_plain_repeated_512_input(const string& filename, int level, double scale) {
    return encode(tile(read_values_from_file(filename), 512, 32, scale), level, num_slots);
}
Ctxt FHEController::read_expanded_input(const string& filename, double scale) {
    return encrypt(stretch(read_values_from_file(filename), 128, 128, 128, scale), 0, num_slots);
}
vector<Ctxt> FHEController::encrypt_many(const vector<Ptxt>& plaintexts) {
    if (plaintexts.empty()) return {};
    if (!batch_rows) {
        vector<Ctxt> out;
        for (const Ptxt& p : plaintexts) out.push_back(encrypt_ptxt(p));
        return out;
    }
    vector<Ctxt> out;
    for (size_t first = 0; first < plaintexts.size(); first += (size_t)max_rows_per_batch) {
        const size_t last = std::min(plaintexts.size(), first + (size_t)max_rows_per_batch);
        vector<const fl_elem*> h;
        for (size_t i = first; i < last; ++i) h.push_back(plaintexts[i]->handle());
        fl_elem* e = nullptr;
        need(fl_encrypt_many(ctx_, h.data(), (int)h.size(), &e), "Encrypt");
        const vector<Ctxt> part = unpack(wrap(e));
        out.insert(out.end(), part.begin(), part.end());
    }
    return out;
}
vector<Ctxt> FHEController::read_expanded_inputs(const vector<string>& filenames, double scale) {
    if (!batch_rows) {
        vector<Ptxt> pts;
        for (const string& f : filenames) pts.push_back(read_plain_expanded_input(f, 0, scale));
        return encrypt_many(pts);
    }
    // all rows of a chunk in one pass: one upload, one batched encoding + encryption (fl_encrypt_values_many)
    vector<Ctxt> out;
    for (size_t first = 0; first < filenames.size(); first += (size_t)max_rows_per_batch) {
        const size_t last = std::min(filenames.size(), first + (size_t)max_rows_per_batch), count = last - first;
        vector<double> slots(count * (size_t)num_slots);
        for (size_t i = 0; i < count; ++i) {
            const vector<double> v = stretch(read_values_from_file(filenames[first + i]), 128, 128, 128, scale);
            std::copy(v.begin(), v.begin() + std::min(v.size(), (size_t)num_slots), slots.begin() + i * (size_t)num_slots);
        }
        fl_elem* ct = nullptr;
        need(fl_encrypt_values_many(ctx_, slots.data(), (int)count, num_slots, 0, num_slots, &ct), "Encrypt");
        const vector<Ctxt> part = unpack(wrap(ct));
        out.insert(out.end(), part.begin(), part.end());
    }
    return out;
}
// The same for M samples at once: element t of the result is ONE batched ciphertext holding file t of every sample (sample-major
// inside), so that a whole forward can run on M samples per call (every primitive below takes batched operands).
vector<Ctxt> FHEController::read_expanded_inputs_many(const vector<vector<string>>& files_per_sample, double scale) {
    const size_t M = files_per_sample.size();
    if (M == 0) return {};
    const size_t T = files_per_sample[0].size();
    for (const auto& f : files_per_sample)
        if (f.size() != T) throw std::invalid_argument("read_expanded_inputs_many: the samples differ in their number of rows");
    if (M == 1) return read_expanded_inputs(files_per_sample[0], scale);
    const size_t per_call = std::max<size_t>(1, (size_t)max_rows_per_batch / M);      // rows t per batched call
    vector<Ctxt> out;
    for (size_t first = 0; first < T; first += per_call) {
        const size_t count = std::min(T, first + per_call) - first;
        vector<double> slots(count * M * (size_t)num_slots);
        for (size_t t = 0; t < count; ++t)
            for (size_t m = 0; m < M; ++m) {
                const vector<double> v = stretch(read_values_from_file(files_per_sample[m][first + t]), 128, 128, 128, scale);
                std::copy(v.begin(), v.begin() + std::min(v.size(), (size_t)num_slots), slots.begin() + (t * M + m) * (size_t)num_slots);
            }
        fl_elem* ct = nullptr;
        need(fl_encrypt_values_many(ctx_, slots.data(), (int)(count * M), num_slots, 0, num_slots, &ct), "Encrypt");
        const vector<Ctxt> part = unpack(wrap(ct), (int)M);
        out.insert(out.end(), part.begin(), part.end());
    }
    return out;
}
Ptxt FHEController::read_plain_expanded_input(const string& filename, int level, double scale) {
    return encode(stretch(read_values_from_file(filename), 128, 128, 128, scale), level, num_slots);
}
Ptxt FHEController::read_plain_expanded_input(const string& filename, int level, double scale, int num_inputs) {
    return encode(stretch(read_values_from_file(filename), 128, 128, num_inputs, scale), level, num_slots);
}

/* ------------------------------------------------------------------ debug printing (secret-key decrypts, as in the reference) ------------------------------------------------------------------ */

namespace {
void print_slots(const vector<double>& v, int first, int last, int step, int precision, const char* zero, bool close_on_index, int close_index) {
    cout << setprecision(precision) << fixed << "[ ";
    for (int i = first; i < last; i += step) {
        const double a = std::abs(v[i]);
        const char* sign = v[i] > 0 ? " " : "-";
        if (close_on_index && i == close_index) cout << sign << a << " ]";
        else if (a < 1e-8) cout << zero << ", ";
        else cout << sign << a << ", ";
    }
}
}  // namespace

void FHEController::print(const Ctxt& c, int slots, string prefix) {
    if (slots == 0) slots = num_slots;
    cout << prefix << " (Lv. " << c->GetLevel() << ") ";
    print_slots(decrypt_tovector(c, num_slots), 0, slots, 1, 4, " 0.0000", true, slots - 1);
    cout << endl;
}
void FHEController::print_expanded(const Ctxt& c, int slots, int expansion_factor, string prefix) {
    if (slots == 0) slots = num_slots;
    cout << prefix << " (Lv. " << c->GetLevel() << ") ";
    print_slots(decrypt_tovector(c, num_slots), 0, slots, expansion_factor, 4, " 0.000", true, slots - 1);
    cout << " ]" << endl;
}
void FHEController::print_padded(const Ctxt& c, int slots, int padding, string prefix) {
    if (slots == 0) slots = num_slots;
    cout << prefix;
    print_slots(decrypt_tovector(c, num_slots), 0, slots * padding, padding, 10, " 0.000", true, slots - 1);
    cout << endl;
}
void FHEController::print_min_max(const Ctxt& c) {
    const vector<double> v = decrypt(c)->GetRealPackedValue();
    cout << "min: " << *min_element(v.begin(), v.end()) << ", max: " << *max_element(v.begin(), v.end()) << endl;
}

/* ------------------------------------------------------------------ rotate-and-add ladders ------------------------------------------------------------------ */
// result = sum_{t < slots} rot(in, t * stride), as log2(slots) doubling steps (F.cpp:829-867).  The whole ladder is one
// C-ABI call so the engine keeps the running ciphertext on the device and reuses its key-switch workspace.

Ctxt FHEController::ladder(const Ctxt& in, int slots, int stride) {
    int steps = 0;
    while ((1 << steps) < slots) ++steps;
    // the doubling keys, plus (hoist_ladders) the extra multiples the engine's hoisted groups use -- up to four doubling steps
    // share one ModUp / ModDown (r + rot(r,k) + ... + rot(r,15k)); generated on first use
    int rots[64];
    const int nr = fl_rotsum_rotations(steps, stride, rots, 64);
    for (int i = 0; i < nr && i < 64; ++i) {
        if (i >= steps && !hoist_ladders) break;
        require_rotation_key(rots[i]);
    }
    fl_elem* e = nullptr;
    need(fl_rotsum(ctx_, in->handle(), steps, stride, &e), "EvalRotate ladder");
    return wrap(e);
}
Ctxt FHEController::rotsum(const Ctxt& in, int slots, int padding) { return ladder(in, slots, padding); }
Ctxt FHEController::rotsum_padded(const Ctxt& in, int slots) { return ladder(in, slots, slots); }
Ctxt FHEController::repeat(const Ctxt& in, int slots) { return ladder(in, slots, -1); }
Ctxt FHEController::repeat(const Ctxt& in, int slots, int padding) { return ladder(in, slots, -padding); }

/* ------------------------------------------------------------------ row batching ------------------------------------------------------------------ */
// The reference walks its row ciphertexts one by one (`for (i < rows.size())`, F.cpp:872-1120) although the iterations are
// independent.  Here rows of identical level / degree / scale are packed into one batched operand, the per-row recipe
// runs once on the pack (each engine stage is a single launch for all rows) and the results are handed back as views.

Ctxt FHEController::pack(const vector<Ctxt>& rows) const {
    if (rows.size() == 1) return rows[0];
    auto h = handles(rows);
    fl_elem* e = nullptr;
    need(fl_batch_pack(ctx_, h.data(), (int)h.size(), &e), "pack");
    return wrap(e);
}

vector<Ctxt> FHEController::unpack(const Ctxt& packed) const {
    const int n = fl_elem_batch(packed->handle());
    if (n == 1) return {packed};
    vector<Ctxt> out((size_t)n);
    for (int i = 0; i < n; ++i) {
        fl_elem* e = nullptr;
        need(fl_batch_slice(ctx_, packed->handle(), i, &e), "slice");
        out[i] = wrap(e);
    }
    return out;
}

// groups of `group` consecutive elements of a batched operand, as views (group = 1: the single elements)
vector<Ctxt> FHEController::unpack(const Ctxt& packed, int group) const {
    const int n = fl_elem_batch(packed->handle());
    if (group < 1 || n % group != 0) throw std::invalid_argument("unpack: the batch does not split into groups of that size");
    if (group == 1) return unpack(packed);
    if (n == group) return {packed};
    vector<Ctxt> out((size_t)(n / group));
    for (int i = 0; i < n / group; ++i) {
        fl_elem* e = nullptr;
        need(fl_batch_range(ctx_, packed->handle(), i * group, group, &e), "slice");
        out[(size_t)i] = wrap(e);
    }
    return out;
}

vector<Ctxt> FHEController::per_row(const vector<Ctxt>& rows, const std::function<Ctxt(const Ctxt&)>& recipe) const {
    vector<Ctxt> out(rows.size());
    if (!batch_rows) {
        for (size_t i = 0; i < rows.size(); ++i) out[i] = recipe(rows[i]);
        return out;
    }
    auto same = [](const Ctxt& a, const Ctxt& b) {
        return a->GetLevel() == b->GetLevel() && a->GetNoiseScaleDeg() == b->GetNoiseScaleDeg() && a->GetScalingFactor() == b->GetScalingFactor() &&
               a->GetSlots() == b->GetSlots() && fl_elem_batch(a->handle()) == fl_elem_batch(b->handle());
    };
    size_t first = 0;
    while (first < rows.size()) {
        // a row may itself be a batch (one element per sample, LinformerForward::add_sample): it comes back as the same batch
        const int each = fl_elem_batch(rows[first]->handle());
        size_t last = first + 1;
        while (last < rows.size() && (last - first + 1) * (size_t)each <= (size_t)std::max(max_rows_per_batch, each) && same(rows[first], rows[last])) ++last;
        const vector<Ctxt> part = unpack(recipe(pack(vector<Ctxt>(rows.begin() + first, rows.begin() + last))), each);
        for (size_t i = 0; i < part.size(); ++i) out[first + i] = part[i];
        first = last;
    }
    return out;
}

vector<Ctxt> FHEController::project_rows(const vector<Ctxt>& rows, const vector<vector<double>>& weights, const vector<Ptxt>& bias) {
    const int n_in = (int)rows.size(), n_out = (int)weights.size();
    vector<double> w((size_t)n_out * n_in);
    for (int o = 0; o < n_out; ++o)
        for (int t = 0; t < n_in; ++t) w[(size_t)o * n_in + t] = weights[o].at(t);
    fl_elem* e = nullptr;
    need(fl_linear_wsum(ctx_, pack(rows)->handle(), w.data(), n_out, &e), "EvalLinearWSum");
    vector<Ctxt> out = unpack(wrap(e));
    if (!bias.empty())
        for (int o = 0; o < n_out; ++o) out[o] = add(out[o], bias.at(o));
    return out;
}

// FLEXIBLEAUTO rescales a degree-2 ciphertext right before its next multiplication.  When many rows are about to be
// multiplied by DIFFERENT plaintexts (the mask loops of wrapUp*), the pending rescales are done here for all rows at once;
// the multiplications that follow find degree-1 operands and produce the same limbs as if each had rescaled on its own.
vector<Ctxt> FHEController::settle_rows(const vector<Ctxt>& rows) const {
    if (!batch_rows) return rows;
    return per_row(rows, [&](const Ctxt& r) {
        if (r->GetNoiseScaleDeg() < 2) return r;
        fl_elem* e = nullptr;
        need(fl_rescale(ctx_, r->handle(), &e), "ModReduce");
        return wrap(e);
    });
}

/* ------------------------------------------------------------------ packed matrix products ------------------------------------------------------------------ */
// "RE": rows arrive Expanded, weight is the row-major 128x128 matrix, summing over the 128 blocks (stride 128) leaves the
// product Repeated.  "CR": rows arrive Repeated, summing inside each block (stride 1) leaves product entry j at slot 128 j.

vector<Ctxt> FHEController::matmulRE(vector<Ctxt> rows, const Ptxt& weight, const Ptxt& bias) { return matmulRE(rows, weight, bias, 128, 128); }

vector<Ctxt> FHEController::matmulRE(vector<Ctxt> rows, const Ptxt& weight, const Ptxt& bias, int row_size, int padding) {
    return per_row(rows, [&](const Ctxt& r) {
        Ctxt acc = rotsum(mult(r, weight), row_size, padding);
        return bias != nullptr ? add(acc, bias) : acc;
    });
}

vector<Ctxt> FHEController::matmulRE(vector<Ctxt> rows, const Ctxt& weight, int row_size, int padding) {
    return per_row(rows, [&](const Ctxt& r) { return rotsum(mult(r, weight), row_size, padding); });
}

// 128 -> 512: four 128x128 blocks, each product masked to its first 128 slots and shifted into place by two rotations of
// -64 (the reference's stand-in for -128, F.cpp:930-931); block 3 is produced first and ends up highest.
vector<Ctxt> FHEController::matmulRElarge(vector<Ctxt>& rows, const vector<Ptxt>& weight, const Ptxt& bias, double mask_value) {
    const int nb = (int)weight.size();
    return per_row(rows, [&](const Ctxt& r) {
        Ctxt acc;
        for (int j = nb - 1; j >= 0; --j) {
            Ctxt part = mask_first_n(rotsum(mult(r, weight[j]), 128, 128), 128, mask_value);
            if (j == nb - 1) acc = part;
            else acc = add(rotate(rotate(acc, -64), -64), part);
        }
        return add(acc, bias);
    });
}

vector<Ctxt> FHEController::matmulCR(vector<Ctxt> rows, const Ptxt& weight, const Ptxt& bias) {
    return per_row(rows, [&](const Ctxt& r) {
        Ctxt acc = rotsum(mult(r, weight), 128, 1);
        return bias != nullptr ? add(acc, bias) : acc;
    });
}

vector<Ctxt> FHEController::matmulCR(vector<Ctxt> rows, const Ctxt& matrix) {
    return per_row(rows, [&](const Ctxt& r) { return rotsum(mult(r, matrix), 64, 1); });
}

Ctxt FHEController::matmulCR_128(Ctxt row, const Ctxt& matrix) { return rotsum(mult(row, matrix), 128, 1); }

vector<Ctxt> FHEController::matmulCR_128(vector<Ctxt> rows, const Ctxt& matrix) {
    return per_row(rows, [&](const Ctxt& r) { return matmulCR_128(r, matrix); });
}

// 512 -> 128: the four block products are summed first, then one stride-1 ladder (F.cpp:998-1026)
vector<Ctxt> FHEController::matmulCRlarge(vector<vector<Ctxt>> rows, vector<Ptxt> weights, const Ptxt& bias) {
    // per_row over the row index, with the four block operands of a row travelling together
    vector<Ctxt> index(rows.size());
    vector<vector<Ctxt>> column(4, vector<Ctxt>(rows.size()));
    for (size_t i = 0; i < rows.size(); ++i)
        for (int b = 0; b < 4; ++b) column[b][i] = rows[i][b];
    size_t at = 0;   // per_row hands out contiguous, ordered groups of column[0]; the other columns are cut the same way
    return per_row(column[0], [&](const Ctxt& first) {
        const size_t n = (size_t)fl_elem_batch(first->handle());
        vector<Ctxt> parts = {mult(first, weights[0])};
        for (int b = 1; b < 4; ++b) parts.push_back(mult(pack(vector<Ctxt>(column[b].begin() + at, column[b].begin() + at + n)), weights[b]));
        at += n;
        Ctxt acc = rotsum(add(parts), 128, 1);
        return bias != nullptr ? add(acc, bias) : acc;
    });
}

// scores of up to 128 queries against the wrapped keys; each query's 128 scores sit at slots 128 j and are scaled by
// 1/64 (1/8 for sqrt(d_head) times r = 1/8, undone later by raising exp to the 8th power), F.cpp:1028-1058
Ctxt FHEController::matmulScores(vector<Ctxt> queries, const Ctxt& key) {
    const double scale = (1 / 8.0) * (1 / 8.0);
    vector<Ctxt> scores = matmulCR_128(queries, key);
    if (scores.size() == 1) return mask_heads_128(scores[0], scale);
    Ctxt packed = rotate(mask_heads_128(scores.back(), scale), -1);
    for (int i = (int)scores.size() - 2; i >= 0; --i) {
        packed = add(packed, mask_heads_128(scores[i], scale));
        if (i > 0) packed = rotate(packed, -1);
    }
    return packed;
}

Ctxt FHEController::matmulScores(Ctxt query, const Ctxt& key) { return mask_heads_128(matmulCR_128(query, key), (1 / 8.0) * (1 / 8.0)); }

/* ------------------------------------------------------------------ layout conversions ------------------------------------------------------------------ */

// vector j (Repeated) keeps only its block j: the result holds vector j in slots 128 j .. 128 j + 127
Ctxt FHEController::wrapUpRepeated(vector<Ctxt> vectors) {
    vectors = settle_rows(vectors);
    vector<Ctxt> blocks;
    blocks.reserve(vectors.size());
    for (size_t i = 0; i < vectors.size(); ++i) blocks.push_back(mask_block(vectors[i], 128 * (int)i, 128 * ((int)i + 1), 1));
    return add(blocks);
}

// vector t has entry j at slot 128 j; interleave so that slot 128 j + t holds (vector t)[j].  Horner over rotate(-1).
Ctxt FHEController::wrapUpExpanded(vector<Ctxt> vectors) {
    // every vector gets the same mask: one batched multiplication (and one batched pending rescale) for all of them
    const vector<Ctxt> masked = per_row(vectors, [&](const Ctxt& v) { return mask_mod_n(v, 128); });
    // the reference's Horner chain acc = rot(acc, -1) + masked[i] equals sum_i rot(masked[i], -i)
    return shifted_sum(masked, -1);
}

// sum_i rot(items[i], stride * i) as a binary tree: level b rotates every odd survivor by stride * 2^b -- one batched key
// switch per level with the power-of-two key -- and adds it to its even neighbour.  Same result as the sequential Horner
// chain the reference writes (F.cpp:1070-1084, 1193-1205) with log2(n) instead of n dependent rotations per item.
Ctxt FHEController::shifted_sum(vector<Ctxt> items, int stride) {
    if (!batch_rows) {
        Ctxt acc = items.back();
        for (int i = (int)items.size() - 2; i >= 0; --i) acc = add(rotate(acc, stride), items[i]);
        return acc;
    }
    for (int step = stride; items.size() > 1; step *= 2) {
        vector<Ctxt> even, odd;
        for (size_t i = 0; i < items.size(); ++i) (i % 2 ? odd : even).push_back(items[i]);
        const vector<Ctxt> moved = per_row(odd, [&](const Ctxt& r) { return rotate(r, step); });
        for (size_t j = 0; j < moved.size(); ++j) even[j] = add(even[j], moved[j]);
        items = even;
    }
    return items[0];
}

// inverse of wrapUpExpanded: vector t comes back Expanded (entry j replicated over block j)
// rot(c, t) for t = 0 .. count-1.  The reference walks c = rotate(c, 1) count-1 times (F.cpp:1086-1100); here the set doubles:
// step 2^b rotates everything produced so far by 2^b in one batched key switch, so rot(c, t) costs popcount(t) key switches
// (less noise than t of them) and the chain is log2(count) launches deep.
vector<Ctxt> FHEController::all_shifts(const Ctxt& c, int count) {
    vector<Ctxt> out = {c};
    if (!batch_rows) {
        for (int t = 1; t < count; ++t) out.push_back(rotate(out.back(), 1));
        return out;
    }
    for (int step = 1; step < count; step *= 2) {
        const int take = std::min(step, count - step);
        const vector<Ctxt> moved = per_row(vector<Ctxt>(out.begin(), out.begin() + take), [&](const Ctxt& r) { return rotate(r, step); });
        out.insert(out.end(), moved.begin(), moved.end());
    }
    return out;
}

vector<Ctxt> FHEController::unwrapExpanded(Ctxt c, int inputs_num) {
    // the rotate(c, 1) chain is sequential; the 7-step replication ladders of the picked columns are independent.  The
    // pending rescale of c is taken once up front instead of once per mask (rescaling commutes with rotation).
    c = settle_rows({c})[0];
    const vector<Ctxt> shifted = all_shifts(c, inputs_num);
    return per_row(shifted, [&](const Ctxt& r) { return repeat(mask_mod_n(r, 128, 0, inputs_num * 128), 128); });
}

vector<Ctxt> FHEController::unwrapScoresExpanded(Ctxt c, int inputs_num) {
    c = settle_rows({c})[0];
    const vector<Ctxt> shifted = all_shifts(c, inputs_num);
    vector<Ctxt> lo = per_row(shifted, [&](const Ctxt& r) { return repeat(mask_mod_n(r, 128, 0, inputs_num * 128), 64); });
    vector<Ctxt> hi = per_row(shifted, [&](const Ctxt& r) { return repeat(mask_mod_n(r, 128, 64, inputs_num * 128), 64); });
    vector<Ctxt> out;
    for (int t = 0; t < inputs_num; ++t) out.push_back(add(lo[t], hi[t]));
    return out;
}

// token `index` of a container (512 slots per token) -> four Repeated 128-vectors
vector<Ctxt> FHEController::unwrap_512_in_4_128(const Ctxt& c, int index) {
    vector<Ctxt> quad;
    for (int b = 0; b < 4; ++b) {
        const int from = index * 512 + 128 * b;
        quad.push_back(repeat(mask_block(c, from, from + 128, 1), 128, -128));
    }
    return quad;
}

vector<vector<Ctxt>> FHEController::unwrapRepeatedLarge(vector<Ctxt> containers, int input_number) {
    // every (token, 128-block) ladder of unwrap_512_in_4_128 is independent: mask them all, replicate them as batches
    containers = settle_rows(containers);
    vector<Ctxt> blocks;
    for (size_t i = 0; i < containers.size(); ++i) {
        const int held = std::min(32, input_number - 32 * (int)i);
        for (int j = 0; j < held; ++j)
            for (int b = 0; b < 4; ++b) blocks.push_back(mask_block(containers[i], j * 512 + 128 * b, j * 512 + 128 * (b + 1), 1));
    }
    blocks = per_row(blocks, [&](const Ctxt& r) { return repeat(r, 128, -128); });
    vector<vector<Ctxt>> out;
    for (size_t i = 0; i + 3 < blocks.size(); i += 4) out.push_back({blocks[i], blocks[i + 1], blocks[i + 2], blocks[i + 3]});
    return out;
}

// 32 tokens (512 slots each) per ciphertext; token order inside a container follows the reference's reversed Horner chain
vector<Ctxt> FHEController::generate_containers(vector<Ctxt> inputs, const Ptxt& bias) {
    vector<Ctxt> containers;
    const int n = (int)inputs.size();
    for (int first = 0; first < n; first += 32) {
        vector<Ctxt> group = slicing(inputs, first, first + 32);
        reverse(group.begin(), group.end());
        const int held = std::min(32, n - first);
        Ctxt packed = wrap_containers(group, held);
        containers.push_back(bias != nullptr ? add(packed, bias) : packed);
    }
    return containers;
}

Ctxt FHEController::wrap_containers(vector<Ctxt> c, int inputs_number) {
    // result = rot(... rot(c[0], -512) + c[1] ..., -512) + c[n-1] = sum_k rot(c[n-1-k], -512 k)
    vector<Ctxt> items(c.rend() - inputs_number, c.rend());
    for (int b = 0; (1 << b) < inputs_number; ++b) {
        int k = -512 * (1 << b);
        if (batch_rows) require_rotation_key(k);
    }
    return shifted_sum(items, -512);
}

/* ------------------------------------------------------------------ masks ------------------------------------------------------------------ */
// kind 0: value on [a, b)   1: value where i % a == b   2: value on i < a   4: value everywhere   5: exp correction (0 inside
// the a x a score region, -1 outside).  Encoded at `level` once, then served from the cache.

Ptxt FHEController::mask_plain(int kind, int a, int b, double value, int level) {
    const auto key = std::make_tuple(kind, a, b, value, level);
    auto hit = mask_cache_.find(key);
    if (hit != mask_cache_.end()) return hit->second;
    vector<double> m((size_t)num_slots, 0.0);
    for (int i = 0; i < num_slots; ++i) {
        bool on = false;
        switch (kind) {
            case 0: on = i >= a && i < b; break;
            case 1: on = i % a == b; break;
            case 2: on = i < a; break;
            case 4: on = true; break;
            case 5: on = !(i % 128 < a && i < 128 * a); break;
        }
        if (on) m[i] = value;
    }
    Ptxt p = encode(m, level, num_slots);
    mask_cache_.emplace(key, p);
    return p;
}

Ctxt FHEController::mask_block(const Ctxt& c, int from, int to, double mask_value) { return mult(c, mask_plain(0, from, to, mask_value, (int)c->GetLevel())); }
Ctxt FHEController::mask_heads(const Ctxt& c, double mask_value) { return mult(c, mask_plain(1, 64, 0, mask_value, (int)c->GetLevel())); }
Ctxt FHEController::mask_heads_128(const Ctxt& c, double mask_value) { return mult(c, mask_plain(1, 128, 0, mask_value, (int)c->GetLevel())); }
Ctxt FHEController::mask_mod_n(const Ctxt& c, int n) { return mult(c, mask_plain(1, n, 0, 1, (int)c->GetLevel())); }
Ctxt FHEController::mask_mod_n(const Ctxt& c, int n, int padding, int) { return mult(c, mask_plain(1, n, padding, 1, (int)c->GetLevel())); }
Ctxt FHEController::mask_first_n(const Ctxt& c, int n, double mask_value) { return mult(c, mask_plain(2, n, 0, mask_value, (int)c->GetLevel())); }

/* ------------------------------------------------------------------ activations ------------------------------------------------------------------ */

// exp(8 x) ~ (T6(x))^8 with T6 the degree-6 Taylor polynomial (F.cpp:1289-1311); slots outside the inputs x inputs score
// region would become 1 and are pulled back to 0
Ctxt FHEController::eval_exp(const Ctxt& c, int inputs_number) {
    const double taylor[7] = {1, 1, 1 / 2.0, 1 / 6.0, 1 / 24.0, 1 / 120.0, 1 / 720.0};
    fl_elem* e = nullptr;
    need(fl_eval_poly(ctx_, c->handle(), taylor, 7, &e), "EvalPoly");
    Ctxt res = wrap(e);
    if ((int)res->GetLevel() + 4 > circuit_depth) res = bootstrap(res);
    fl_elem* eight[8];
    for (auto& h : eight) h = res->handle();
    need(fl_mul_many(ctx_, eight, 8, &e), "EvalMultMany");
    res = wrap(e);
    if (inputs_number <= 0) return res;
    return add(res, mask_plain(5, inputs_number, 0, -1, (int)res->GetLevel()));
}

Ctxt FHEController::eval_inverse(const Ctxt& c, double min, double max) {
    const double middle = (max - min) / 2;
    Ctxt res = add(c, encode(-middle - min, (int)c->GetLevel(), num_slots));
    res = mult(res, encode(1 / middle, (int)res->GetLevel(), num_slots));
    return chebyshev([](double x) { return 1 / ((x * 9895) + 9995); }, res, -1, 1, 200);
}
Ctxt FHEController::eval_inverse_naive(const Ctxt& c, double min, double max) {
    return chebyshev([](double x) { return 1 / x; }, c, min, max, 119);
}
Ctxt FHEController::eval_inverse_naive_2(const Ctxt& c, double min, double max, double mult) {
    return chebyshev([mult](double x) { return mult / x; }, c, min, max, 200);
}
Ctxt FHEController::eval_gelu_function(const Ctxt& c, double min, double max, double mult, int degree) {
    return chebyshev([mult](double x) { const double y = x * (1 / mult); return 0.5 * y * (1 + erf(y / 1.41421356237)); }, c, min, max, degree);
}
Ctxt FHEController::eval_tanh_function(const Ctxt& c, double min, double max, double mult, int degree) {
    return chebyshev([mult](double x) { return tanh(x * (1 / mult)); }, c, min, max, degree);
}

vector<Ctxt> FHEController::slicing(vector<Ctxt>& arr, int X, int Y) {
    if (Y - X >= (int)arr.size()) return arr;
    Y = std::min(Y, (int)arr.size());
    return vector<Ctxt>(arr.begin() + X, arr.begin() + Y);
}

/* ------------------------------------------------------------------ ciphertext files ------------------------------------------------------------------ */

void FHEController::save(Ctxt v, string filename) {
    if (fl_elem_save(ctx_, v->handle(), filename.c_str())) cerr << "Could not write \"" << filename << "\": " << fl_last_error() << endl;
}

void FHEController::save(vector<Ctxt> v, string filename) {
    // a vector is stored as <filename> holding the count, plus <filename>.<i> per element
    std::ofstream idx(filename, ios::out | ios::binary);
    const uint64_t n = v.size();
    idx.write("FLCKVEC", 8);
    idx.write(reinterpret_cast<const char*>(&n), 8);
    for (size_t i = 0; i < v.size(); ++i) save(v[i], filename + "." + to_string(i));
}

vector<Ctxt> FHEController::load_vector(string filename) {
    vector<Ctxt> result;
    std::ifstream idx(filename, ios::in | ios::binary);
    char magic[8] = {0};
    uint64_t n = 0;
    if (!idx.read(magic, 8) || std::memcmp(magic, "FLCKVEC", 8) || !idx.read(reinterpret_cast<char*>(&n), 8)) {
        cerr << "Could not find \"" << filename << "\"" << endl;
        return result;
    }
    for (uint64_t i = 0; i < n; ++i) result.push_back(load_ciphertext(filename + "." + to_string(i)));
    return result;
}

Ctxt FHEController::load_ciphertext(string filename) {
    fl_elem* e = nullptr;
    if (fl_elem_load(ctx_, filename.c_str(), &e)) {
        cerr << "Could not find \"" << filename << "\"" << endl;
        return Ctxt();
    }
    return wrap(e);
}

Ctxt FHEController::adopt(fl_elem* e) const { return wrap(e); }
