// scheme.cpp -- CKKS scheme layer: keys, encoding, encryption and the leveled operations with OpenFHE's
// FLEXIBLEAUTO bookkeeping, all arithmetic on the device engine.
#include "scheme.h"

#include <sys/random.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace flk {

namespace {

// sparse ternary secret with h non-zero coefficients (rejection loop, host side).  next() is the 64-bit stream to draw from:
// SplitMix64 for the seeded test entry, ChaCha20 key stream otherwise.
template <class Next>
std::vector<int8_t> sample_sparse(int N, int h, Next&& next) {
    std::vector<int8_t> s(N, 0);
    for (int placed = 0; placed < h;) {
        u64 pos = next() % N, sg = next() & 1;
        if (!s[pos]) { s[pos] = sg ? -1 : 1; ++placed; }
    }
    return s;
}

// 32 bytes from the operating system's CSPRNG
ChaChaKey os_random_key() {
    ChaChaKey k;
    size_t got = 0;
    while (got < sizeof k.k) {
        const ssize_t r = getrandom(reinterpret_cast<char*>(k.k) + got, sizeof k.k - got, 0);
        if (r <= 0) break;
        got += (size_t)r;
    }
    if (got < sizeof k.k) {
        FILE* f = std::fopen("/dev/urandom", "rb");
        if (!f || std::fread(reinterpret_cast<char*>(k.k) + got, 1, sizeof k.k - got, f) != sizeof k.k - got) {
            if (f) std::fclose(f);
            throw std::runtime_error("no operating-system randomness available (getrandom and /dev/urandom both failed)");
        }
        std::fclose(f);
    }
    return k;
}

// special FFT of CKKS encoding (A.9), tables cached per slot count
struct FftTables {
    std::vector<uint32_t> rot;
    std::vector<double> cre, cim;
};
// process-wide cache shared by every Scheme (one controller per host thread is a supported mode): guarded, and std::map never
// moves its nodes, so the reference handed out stays valid while other threads insert
const FftTables& fft_tables(int n) {
    static std::map<int, FftTables> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(n);
    if (it != cache.end()) return it->second;
    FftTables t;
    const uint32_t m = 4u * n;
    t.rot.resize(n);
    uint32_t p = 1;
    for (int i = 0; i < n; ++i) { t.rot[i] = p; p = (uint32_t)(((u64)p * 5) % m); }
    t.cre.resize(m + 1); t.cim.resize(m + 1);
    for (uint32_t k = 0; k <= m; ++k) {
        double a = 2.0 * M_PI * (double)k / (double)m;
        t.cre[k] = std::cos(a); t.cim[k] = std::sin(a);
    }
    return cache.emplace(n, std::move(t)).first->second;
}
void bit_reverse(double* re, double* im, int n) {
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j >= bit; bit >>= 1) j -= bit;
        j += bit;
        if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
    }
}
void fft_special_inv(double* re, double* im, int n) {
    const FftTables& t = fft_tables(n);
    const uint32_t m = 4u * n;
    for (int len = n; len >= 2; len >>= 1) {
        const int lenh = len >> 1;
        const uint32_t lenq = (uint32_t)len << 2;
        for (int i = 0; i < n; i += len)
            for (int j = 0; j < lenh; ++j) {
                const uint32_t idx = (lenq - (t.rot[j] % lenq)) * (m / lenq);
                const double ur = re[i + j] + re[i + j + lenh], ui = im[i + j] + im[i + j + lenh];
                const double wr = re[i + j] - re[i + j + lenh], wi = im[i + j] - im[i + j + lenh];
                const double kr = t.cre[idx], ki = t.cim[idx];
                re[i + j] = ur; im[i + j] = ui;
                re[i + j + lenh] = wr * kr - wi * ki;
                im[i + j + lenh] = wr * ki + wi * kr;
            }
    }
    bit_reverse(re, im, n);
    for (int i = 0; i < n; ++i) { re[i] /= n; im[i] /= n; }
}
void fft_special(double* re, double* im, int n) {
    const FftTables& t = fft_tables(n);
    const uint32_t m = 4u * n;
    bit_reverse(re, im, n);
    for (int len = 2; len <= n; len <<= 1) {
        const int lenh = len >> 1;
        const uint32_t lenq = (uint32_t)len << 2;
        for (int i = 0; i < n; i += len)
            for (int j = 0; j < lenh; ++j) {
                const uint32_t idx = (t.rot[j] % lenq) * (m / lenq);
                const double kr = t.cre[idx], ki = t.cim[idx];
                const double ar = re[i + j], ai = im[i + j], br = re[i + j + lenh], bi = im[i + j + lenh];
                const double wr = br * kr - bi * ki, wi = br * ki + bi * kr;
                re[i + j] = ar + wr; im[i + j] = ai + wi;
                re[i + j + lenh] = ar - wr; im[i + j + lenh] = ai - wi;
            }
    }
}

}  // namespace

Scheme::Scheme(const ParamSpec& spec, int device) : eng(spec, device), P(eng.P) {
    // independent keys for the secret key, the public `a` polynomials, the key errors and the encryption randomness
    for (ChaChaKey& k : rng_keys_) k = os_random_key();
}

Scheme::~Scheme() {
    try {
        eng.sync();
        for (auto& kv : dev_fft_) { cudaFree(kv.second.rot); cudaFree(kv.second.cre); cudaFree(kv.second.cim); }
        if (stage_) { cudaFreeHost(stage_); for (auto& ev : stage_ev_) cudaEventDestroy(ev); }
        boot_.clear();
        eng.release(sk_); eng.release(pk_); eng.release(mk_);
        for (auto& kv : gk_) eng.release(kv.second);
        eng.sync();
    } catch (...) {}
}

Elem Scheme::make(int ncomp, int l, int deg, double scale, int slots, int batch) {
    Elem e;
    e.mem = std::make_shared<DevMem>(&eng, (size_t)batch * ncomp * l * P.N);
    e.batch = batch; e.ncomp = ncomp; e.l = l; e.deg = deg; e.scale = scale; e.slots = slots;
    return e;
}

Elem Scheme::slice(const Elem& a, int i) const {
    if (i < 0 || i >= a.batch) throw std::out_of_range("batch slice out of range");
    Elem e = a;
    e.off = a.off + (size_t)i * a.words_each(P.N);
    e.batch = 1;
    return e;
}

Elem Scheme::range(const Elem& a, int first, int count) const {
    if (first < 0 || count < 1 || first + count > a.batch) throw std::out_of_range("batch range out of range");
    Elem e = a;
    e.off = a.off + (size_t)first * a.words_each(P.N);
    e.batch = count;
    return e;
}

Elem Scheme::pack(const std::vector<Elem>& v) {
    if (v.empty()) throw std::invalid_argument("pack: empty input");
    const Elem& f = v[0];
    size_t total = 0;
    bool contiguous = true;
    for (const Elem& e : v) {
        if (e.ncomp != f.ncomp || e.l != f.l || e.deg != f.deg || e.scale != f.scale || e.slots != f.slots)
            throw std::invalid_argument("pack: ciphertexts differ in level, degree or scale");
        contiguous = contiguous && e.mem == f.mem && e.off == f.off + total * f.words_each(P.N);
        total += e.batch;
    }
    if (contiguous) {   // slices of one batch, in order: a view is enough
        Elem r = f;
        r.batch = (int)total;
        return r;
    }
    Elem r = make(f.ncomp, f.l, f.deg, f.scale, f.slots, (int)total);
    size_t at = 0;
    for (const Elem& e : v) {
        eng.copy(r.data() + at, e.data(), e.batch * e.words_each(P.N));
        at += e.batch * e.words_each(P.N);
    }
    return r;
}

int Scheme::max_batch(int l) const {
    const double per = (double)(l + P.beta(l) * (l + P.K) + 2 * (l + P.K) + 2 * l + 4 * l) * P.N * 8.0;   // key-switch workspace + in/out
    return std::max(1, std::min(64, (int)(6.0e9 / per)));
}

// run a single-ciphertext routine over every element of a batched operand
std::vector<Elem> Scheme::split_run(const Elem& a, const std::function<Elem(const Elem&)>& f) {
    std::vector<Elem> out;
    for (int i = 0; i < a.batch; ++i) out.push_back(f(slice(a, i)));
    return out;
}

// ---------------------------------------------------------------- keys
void Scheme::sample_to_eval(u64* dst, const std::vector<int8_t>& s, const LimbSel& sel) {
    int8_t* d8 = (int8_t*)eng.alloc((P.N + 7) / 8);
    FLK_CUDA(cudaMemcpyAsync(d8, s.data(), P.N, cudaMemcpyHostToDevice, eng.stream));
    launch_reduce_i8(eng.T, dst, d8, sel, eng.stream);
    eng.ntt(dst, sel);
    eng.sync();   // s is a host temporary
    eng.release((u64*)d8);
}

// Uniform, ternary and Gaussian polynomials are generated on the device (encode.cu): no host loop, no upload, no
// synchronisation.  A Draw names the stream: production draws (seeded == false) are ChaCha20 key stream under the key of their
// purpose with a fresh 64-bit stream number each; seeded draws are the SplitMix64 streams the CPU restatement walks, reachable
// only through fl_keygen_seeded / fl_encrypt_seeded (parity tests).
void Scheme::uniform_to_dev(u64* dst, const Draw& d, const LimbSel& sel) {
    if (!d.seeded) {
        launch_uniform_limbs_csprng(eng.T, dst, rng_keys_[(int)d.use], rng_stream_, sel, eng.stream);
        rng_stream_ += (u64)sel.n;   // one stream per limb
        return;
    }
    u64 seeds[kMaxLimbSel];
    for (int i = 0; i < sel.n; ++i) seeds[i] = SplitMix::sub(d.seed, 1000 + sel.m[i]);
    launch_uniform_limbs(eng.T, dst, seeds, sel, eng.stream);
}

void Scheme::sample_dev_to_eval(u64* dst, const Draw& d, int kind, const LimbSel& sel) {
    if (d.seeded) launch_sample_limbs(eng.T, dst, d.seed, kind, sel, eng.stream);
    else launch_sample_limbs_csprng(eng.T, dst, rng_keys_[(int)d.use], rng_stream_++, kind, sel, eng.stream);
    eng.ntt(dst, sel);
}

// seed != 0: the deterministic test entry (every stream derives from `seed`; never use outside tests).  seed == 0: operating-
// system randomness, nothing to persist, nothing from which a public polynomial could be traced back to the secret key.
void Scheme::keygen(u64 seed) {
    seeded_keys_ = seed != 0;
    key_seed_ = seed;
    if (!seeded_keys_) for (ChaChaKey& k : rng_keys_) k = os_random_key();
    const int T = P.T, N = P.N;
    if (!sk_) sk_ = eng.alloc((size_t)T * N);
    if (!pk_) pk_ = eng.alloc((size_t)2 * P.L * N);
    if (P.spec.sparse_h > 0) {   // rejection loop: host
        std::vector<int8_t> s;
        if (seeded_keys_) {
            SplitMix r(SplitMix::sub(seed, 0));
            s = sample_sparse(N, P.spec.sparse_h, [&] { return r.next(); });
        } else {
            uint32_t blk[16];
            u64 ctr = 0;
            int at = 8;
            const u64 stream = rng_stream_++;
            s = sample_sparse(N, P.spec.sparse_h, [&] {
                if (at == 8) { chacha20_block(rng_keys_[(int)Use::Secret], ctr++, stream, blk); at = 0; }
                return chacha_u64(blk, at++);
            });
        }
        sample_to_eval(sk_, s, sel_range(0, T));
        std::fill(s.begin(), s.end(), 0);
    } else {
        sample_dev_to_eval(sk_, draw(Use::Secret, SplitMix::sub(seed, 0)), 0, sel_range(0, T));
    }
    // pk = (e - a*s, a)
    const LimbSel q = sel_range(0, P.L);
    const size_t pl = (size_t)P.L * N;
    const u64 pseed = seed + 1;
    uniform_to_dev(pk_ + pl, draw(Use::Public, SplitMix::sub(pseed, 10)), q);
    u64* e = eng.alloc(pl);
    sample_dev_to_eval(e, draw(Use::Error, SplitMix::sub(pseed, 11)), 1, q);
    launch_ew(eng.T, EwOp::Mul, pk_, pk_ + pl, sk_, q, 1, 1, 0, 0, 0, eng.stream);
    launch_ew(eng.T, EwOp::Sub, pk_, e, pk_, q, 1, 1, 0, 0, 0, eng.stream);
    eng.release(e);
}

// hybrid key-switching key from s_old to s_new (A.6): per digit d, (b_d, a_d) with
// b_d = e_d - a_d s_new + [limb in digit d] (P mod q_i) s_old
void Scheme::keyswitch_gen(const u64* sk_old, const u64* sk_new, u64 seed, u64* evk) {
    const int T = P.T, N = P.N;
    const size_t kl = (size_t)T * N;
    const LimbSel all = sel_range(0, T);
    u64* e = eng.alloc(kl);
    u64* tmp = eng.alloc((size_t)P.alpha * N);
    for (int d = 0; d < P.dnum; ++d) {
        u64* b = evk + (size_t)d * 2 * kl;
        u64* a = b + kl;
        uniform_to_dev(a, draw(Use::Public, SplitMix::sub(seed, 100 + 2 * d)), all);
        sample_dev_to_eval(e, draw(Use::Error, SplitMix::sub(seed, 101 + 2 * d)), 1, all);
        launch_ew(eng.T, EwOp::Mul, b, a, sk_new, all, 1, 1, 0, 0, 0, eng.stream);
        launch_ew(eng.T, EwOp::Sub, b, e, b, all, 1, 1, 0, 0, 0, eng.stream);
        const int lo = d * P.alpha, hi = std::min(lo + P.alpha, P.L), ns = hi - lo;
        ScalarSet sc;
        for (int i = 0; i < ns; ++i) { sc.c[i] = P.P_mod(P.q[lo + i]); sc.c_sh[i] = nt::shoup(sc.c[i], P.q[lo + i]); }
        const LimbSel ds = sel_range(lo, ns);
        launch_mul_scalar(eng.T, tmp, sk_old + (size_t)lo * N, sc, ds, 1, eng.stream);
        launch_ew(eng.T, EwOp::Add, b + (size_t)lo * N, b + (size_t)lo * N, tmp, ds, 1, 1, 0, 0, 0, eng.stream);
    }
    eng.release(e); eng.release(tmp);
}

void Scheme::gen_mult_key() {
    if (!sk_) throw std::runtime_error("EvalMultKeyGen: no secret key");
    const size_t kl = (size_t)P.T * P.N;
    u64* s2 = eng.alloc(kl);
    launch_ew(eng.T, EwOp::Mul, s2, sk_, sk_, sel_range(0, P.T), 1, 1, 0, 0, 0, eng.stream);
    if (!mk_) mk_ = eng.alloc(eng.evk_words());
    keyswitch_gen(s2, sk_, key_seed_ + 2, mk_);
    eng.release(s2);
}

void Scheme::gen_galois_key(uint32_t g) {
    if (!sk_) throw std::runtime_error("EvalAutomorphismKeyGen: no secret key");
    if (gk_.count(g)) return;
    // switch from s to sigma_{g^-1}(s); the permutation by g that follows restores s (A.5)
    uint32_t gi = 1;
    for (int i = 0; i < 6; ++i) gi = gi * (2 - g * gi);
    gi &= (uint32_t)(2 * P.N - 1);
    u64* sp = eng.alloc((size_t)P.T * P.N);
    eng.automorph(sp, sk_, gi, P.T);
    u64* evk = eng.alloc(eng.evk_words());
    keyswitch_gen(sk_, sp, key_seed_ + 1000 + g, evk);
    eng.release(sp);
    gk_[g] = evk;
}
void Scheme::gen_rotation_key(int k) { gen_galois_key(P.galois_for_rotation(k)); }

void Scheme::clear_rotation_keys() {
    for (auto& kv : gk_) eng.release(kv.second);
    gk_.clear();
}
void Scheme::clear_mult_key() { eng.release(mk_); mk_ = nullptr; }

void Scheme::export_sk(u64* out) const { const_cast<Engine&>(eng).download(out, sk_, (size_t)P.T * P.N); }
void Scheme::export_pk(u64* out) const { const_cast<Engine&>(eng).download(out, pk_, (size_t)2 * P.L * P.N); }
void Scheme::export_evk(uint32_t g, u64* out) const {
    const u64* src = g == 0 ? mk_ : (gk_.count(g) ? gk_.at(g) : nullptr);
    if (!src) throw std::runtime_error("export_evk: key not present");
    const_cast<Engine&>(eng).download(out, src, eng.evk_words());
}
void Scheme::import_keys(const u64* sk, const u64* pk) {
    if (sk) { if (!sk_) sk_ = eng.alloc((size_t)P.T * P.N); eng.upload(sk_, sk, (size_t)P.T * P.N); }
    if (pk) { if (!pk_) pk_ = eng.alloc((size_t)2 * P.L * P.N); eng.upload(pk_, pk, (size_t)2 * P.L * P.N); }
    eng.sync();
}
void Scheme::import_evk(uint32_t g, const u64* evk) {
    u64*& dst = g == 0 ? mk_ : gk_[g];
    if (!dst) dst = eng.alloc(eng.evk_words());
    eng.upload(dst, evk, eng.evk_words());
    eng.sync();
}

// ---------------------------------------------------------------- encode / decode
void Scheme::encode_coeffs(const cplx* vals, int n, int slots, double scale, std::vector<i128>& co) const {
    const int N = P.N, Nh = N / 2;
    if (slots < 1 || slots > Nh || (slots & (slots - 1))) throw std::invalid_argument("encode: slots must be a power of two <= N/2");
    const int gap = Nh / slots;
    std::vector<double> re(slots, 0.0), im(slots, 0.0);
    for (int i = 0; i < std::min(n, slots); ++i) { re[i] = vals[i].real(); im[i] = vals[i].imag(); }
    fft_special_inv(re.data(), im.data(), slots);
    co.assign(N, 0);
    for (int i = 0; i < slots; ++i) {
        co[(size_t)i * gap] = (i128)std::rint(re[i] * scale);
        co[(size_t)Nh + (size_t)i * gap] = (i128)std::rint(im[i] * scale);
    }
}

void Scheme::coeffs_to_dev(u64* dst, const std::vector<i128>& co, int l) {
    const int N = P.N;
    std::vector<int64_t> lohi((size_t)2 * N);
    for (int j = 0; j < N; ++j) { lohi[2 * j] = (int64_t)(u64)co[j]; lohi[2 * j + 1] = (int64_t)(co[j] >> 64); }
    int64_t* d = (int64_t*)eng.alloc((size_t)2 * N);
    FLK_CUDA(cudaMemcpyAsync(d, lohi.data(), (size_t)16 * N, cudaMemcpyHostToDevice, eng.stream));
    const LimbSel sel = sel_range(0, l);
    launch_reduce_i128(eng.T, dst, d, sel, eng.stream);
    eng.ntt(dst, sel);
    eng.sync();
    eng.release((u64*)d);
}

// MakeCKKSPackedPlaintext on the device: the slot values are staged through pinned memory, the special inverse FFT, the
// scaling / rounding and the reduction into the RNS limbs run as kernels (encode.cu), then the limbs are transformed.
Elem Scheme::encode_at(const cplx* vals, int n, int l, double scale, int slots, int deg) {
    if (l < 1 || l > P.L) throw std::invalid_argument("encode: level out of range");
    const int Nh = P.N / 2;
    if (slots < 1 || slots > Nh || (slots & (slots - 1))) throw std::invalid_argument("encode: slots must be a power of two <= N/2");
    DevFft& f = dev_fft(slots);
    int slot;
    double* h = stage_slot(slot);
    const int m = std::min(n, slots);
    for (int i = 0; i < m; ++i) { h[i] = vals[i].real(); h[slots + i] = vals[i].imag(); }
    for (int i = m; i < slots; ++i) { h[i] = 0.0; h[slots + i] = 0.0; }
    double* d = (double*)eng.alloc((size_t)2 * slots);
    FLK_CUDA(cudaMemcpyAsync(d, h, (size_t)2 * slots * sizeof(double), cudaMemcpyHostToDevice, eng.stream));
    FLK_CUDA(cudaEventRecord(stage_ev_[slot], eng.stream));
    Elem e = make(1, l, deg, scale, slots);
    launch_encode(eng.T, e.data(), d, d + slots, slots, scale, l, f.rot, f.cre, f.cim, eng.stream);
    eng.ntt(e.data(), sel_range(0, l));
    eng.release((u64*)d);
    return e;
}

// pinned staging ring for small host -> device uploads: a slot is reused only after the copy that read it has completed, so
// the host never waits for the device in the steady state
double* Scheme::stage_slot(int& slot) {
    const size_t each = (size_t)P.N;   // doubles per slot (2 x N/2 slot values)
    if (!stage_) {
        FLK_CUDA(cudaMallocHost(&stage_, kStageSlots * each * sizeof(double)));
        for (auto& ev : stage_ev_) FLK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
    slot = stage_next_++ % kStageSlots;
    FLK_CUDA(cudaEventSynchronize(stage_ev_[slot]));
    return stage_ + (size_t)slot * each;
}
void Scheme::upload_small(u64* dst, const u64* src, size_t words) {
    if (words > (size_t)P.N) { eng.upload(dst, src, words); eng.sync(); return; }
    int slot;
    u64* h = reinterpret_cast<u64*>(stage_slot(slot));
    std::memcpy(h, src, words * 8);
    FLK_CUDA(cudaMemcpyAsync(dst, h, words * 8, cudaMemcpyHostToDevice, eng.stream));
    FLK_CUDA(cudaEventRecord(stage_ev_[slot], eng.stream));
}

Scheme::DevFft& Scheme::dev_fft(int slots) {
    auto it = dev_fft_.find(slots);
    if (it != dev_fft_.end()) return it->second;
    const FftTables& t = fft_tables(slots);
    DevFft f{};
    FLK_CUDA(cudaMalloc(&f.rot, t.rot.size() * sizeof(uint32_t)));
    FLK_CUDA(cudaMalloc(&f.cre, t.cre.size() * sizeof(double)));
    FLK_CUDA(cudaMalloc(&f.cim, t.cim.size() * sizeof(double)));
    FLK_CUDA(cudaMemcpyAsync(f.rot, t.rot.data(), t.rot.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, eng.stream));
    FLK_CUDA(cudaMemcpyAsync(f.cre, t.cre.data(), t.cre.size() * sizeof(double), cudaMemcpyHostToDevice, eng.stream));
    FLK_CUDA(cudaMemcpyAsync(f.cim, t.cim.data(), t.cim.size() * sizeof(double), cudaMemcpyHostToDevice, eng.stream));
    eng.sync();
    return dev_fft_.emplace(slots, f).first->second;
}

Elem Scheme::encode(const cplx* vals, int n, int level, int slots, int deg) {
    if (level < 0 || level >= P.L) throw std::invalid_argument("encode: level out of range");
    double scale = P.sf[level];
    if (deg == 2) scale *= scale;
    return encode_at(vals, n, P.L - level, scale, slots, deg);
}

// B real slot vectors (n values each, row-major) -> ONE batched plaintext: one upload, one special FFT, one finish and one
// transform for all of them instead of seven launches and a staged copy per vector (a forward encodes S + 64 input rows).
// Element b is bit-identical to encode_real(vals + b n, ...): same kernels, one more grid dimension.
Elem Scheme::encode_many_real(const double* vals, int B, int n, int level, int slots) {
    if (level < 0 || level >= P.L) throw std::invalid_argument("encode: level out of range");
    if (B < 1) throw std::invalid_argument("encode (batched): no vectors");
    const int Nh = P.N / 2, l = P.L - level;
    if (slots < 1 || slots > Nh || (slots & (slots - 1))) throw std::invalid_argument("encode: slots must be a power of two <= N/2");
    if (n < 0 || n > slots) throw std::invalid_argument("encode (batched): more values than slots");
    DevFft& f = dev_fft(slots);
    const size_t each = (size_t)slots;
    double* d = (double*)eng.alloc(2 * each * B);                 // [B][slots] real parts, then [B][slots] imaginary parts (zero)
    FLK_CUDA(cudaMemsetAsync(d, 0, 2 * each * B * sizeof(double), eng.stream));
    // the source is the caller's (pageable) memory: the runtime stages it and returns once it has been read
    FLK_CUDA(cudaMemcpy2DAsync(d, each * sizeof(double), vals, (size_t)n * sizeof(double), (size_t)n * sizeof(double), B, cudaMemcpyHostToDevice, eng.stream));
    Elem e = make(1, l, 1, P.sf[level], slots, B);
    launch_encode(eng.T, e.data(), d, d + each * B, slots, P.sf[level], l, f.rot, f.cre, f.cim, eng.stream, 0, B);
    eng.ntt(e.data(), sel_range(0, l), B, (size_t)l * P.N);
    eng.release((u64*)d);
    return e;
}

Elem Scheme::encode_real(const double* vals, int n, int level, int slots) {
    std::vector<cplx> v(n);
    for (int i = 0; i < n; ++i) v[i] = cplx(vals[i], 0.0);
    return encode(v.data(), n, level, slots, 1);
}

void Scheme::decode(const Elem& pt, cplx* out, int slots) {
    const int N = P.N, Nh = N / 2;
    if (slots <= 0) slots = pt.slots;
    const int gap = Nh / slots;
    const int nl = pt.l >= 2 ? 2 : 1;
    u64* d = eng.alloc((size_t)nl * N);
    eng.copy(d, pt.data(), (size_t)nl * N);
    eng.intt(d, sel_range(0, nl));
    std::vector<u64> x((size_t)nl * N);
    eng.download(x.data(), d, x.size());
    eng.release(d);
    std::vector<double> re(slots), im(slots);
    auto coef = [&](int j) -> double {
        if (nl == 1) {
            const u64 q = P.q[0];
            return x[j] > q / 2 ? -(double)(q - x[j]) : (double)x[j];
        }
        const u64 q0 = P.q[0], q1 = P.q[1];
        static thread_local u64 cq0 = 0, cq1 = 0, inv = 0;
        if (cq0 != q0 || cq1 != q1) { cq0 = q0; cq1 = q1; inv = nt::invmod(q0 % q1, q1); }
        const u64 x0 = x[j], x1 = x[(size_t)N + j];
        const u64 x0m = x0 % q1;
        const u64 dlt = x1 >= x0m ? x1 - x0m : x1 + q1 - x0m;
        const u128 v = (u128)x0 + (u128)q0 * nt::mulmod(dlt, inv, q1), Q = (u128)q0 * q1;
        return v > Q / 2 ? -(double)(Q - v) : (double)v;
    };
    for (int i = 0; i < slots; ++i) { re[i] = coef(i * gap) / pt.scale; im[i] = coef(Nh + i * gap) / pt.scale; }
    fft_special(re.data(), im.data(), slots);
    for (int i = 0; i < slots; ++i) out[i] = cplx(re[i], im[i]);
}

// ---------------------------------------------------------------- encrypt / decrypt
Elem Scheme::encrypt_seeded(const Elem& pt, u64 seed) { return encrypt_with(pt, true, seed); }
// every encryption draws (v, e0, e1) from fresh ChaCha20 streams under the encryption key of this process: no two
// encryptions, in this or any other process, share randomness
Elem Scheme::encrypt(const Elem& pt) { return encrypt_with(pt, false, 0); }

// Encrypt(pk, pt_b) for B plaintexts of one level as ONE batched operand: the ternary / Gaussian polynomials of all of them come
// out of one sampler launch each (ChaCha20 stream per polynomial), one batched transform, one batched product / sum per component --
// a dozen launches in all instead of a dozen per ciphertext (a forward encrypts S + 64 rows: FHEController.cpp:628-649 per file).
Elem Scheme::encrypt_many(const std::vector<const Elem*>& pts) {
    if (!pk_) throw std::runtime_error("Encrypt: no public key");
    if (pts.empty()) throw std::invalid_argument("Encrypt: no plaintexts");
    const Elem& f = *pts[0];
    const bool one_batched = pts.size() == 1 && f.batch > 1;      // a batched plaintext (encode_many_real) instead of a list
    for (const Elem* p : pts)
        if (p->ncomp != 1 || (p->batch != 1 && !one_batched) || p->l != f.l || p->deg != f.deg || p->scale != f.scale || p->slots != f.slots)
            throw std::invalid_argument("Encrypt (batched): plaintexts must share level, degree, scale and slot count");
    const int B = one_batched ? f.batch : (int)pts.size(), l = f.l, N = P.N;
    const size_t pl = (size_t)l * N, pkl = (size_t)P.L * N, cs = 2 * pl;
    const LimbSel sel = sel_range(0, l);
    Elem ct = make(2, l, f.deg, f.scale, f.slots, B);
    u64* v = eng.alloc(pl * B);
    u64* e = eng.alloc(pl * B);
    u64* t = eng.alloc(pl * B);
    auto sample = [&](u64* dst, int kind) {
        launch_sample_limbs_csprng(eng.T, dst, rng_keys_[(int)Use::Encrypt], rng_stream_, kind, sel, eng.stream, B, pl);
        rng_stream_ += (u64)B;
        eng.ntt(dst, sel, B, pl);
    };
    sample(v, 0);
    sample(e, 1);
    // c0 = pk0 v + e0 + pt
    launch_ew(eng.T, EwOp::Mul, t, v, pk_, sel, 1, B, pl, 0, 0, eng.stream);
    launch_ew(eng.T, EwOp::Add, t, t, e, sel, 1, B, pl, pl, 0, eng.stream);
    bool contiguous = true;                                       // slices of one batched plaintext, in order: one launch adds them all
    for (int b = 1; b < B && !one_batched && contiguous; ++b) contiguous = pts[b]->data() == pts[0]->data() + (size_t)b * pl;
    if (contiguous) {
        FLK_CUDA(cudaMemcpy2DAsync(ct.data(), cs * 8, t, pl * 8, pl * 8, B, cudaMemcpyDeviceToDevice, eng.stream));
        launch_ew(eng.T, EwOp::Add, ct.data(), ct.data(), f.data(), sel, 1, B, cs, pl, 0, eng.stream);
    } else {
        for (int b = 0; b < B; ++b) launch_ew(eng.T, EwOp::Add, ct.data() + (size_t)b * cs, t + (size_t)b * pl, pts[b]->data(), sel, 1, 1, 0, 0, 0, eng.stream);
    }
    // c1 = pk1 v + e1
    sample(e, 1);
    launch_ew(eng.T, EwOp::Mul, t, v, pk_ + pkl, sel, 1, B, pl, 0, 0, eng.stream);
    launch_ew(eng.T, EwOp::Add, t, t, e, sel, 1, B, pl, pl, 0, eng.stream);
    FLK_CUDA(cudaMemcpy2DAsync(ct.data() + pl, cs * 8, t, pl * 8, pl * 8, B, cudaMemcpyDeviceToDevice, eng.stream));
    eng.release(v); eng.release(e); eng.release(t);
    return ct;
}

// Encode and encrypt B real vectors in one pass: the encoded message stays in coefficient form until the error e0 has been added to
// it, so message and error share ONE forward transform (three per ciphertext -- v, e0 + m, e1 -- instead of four), and v lives in the
// c1 slot of the result, so both components are finished in place (no staging buffer, no strided copies).  Same distribution as
// Encrypt(pk, MakeCKKSPackedPlaintext(values)); used for the S + 64 input rows of a forward.
Elem Scheme::encrypt_values_many(const double* vals, int B, int n, int level, int slots) {
    if (!pk_) throw std::runtime_error("Encrypt: no public key");
    if (level < 0 || level >= P.L) throw std::invalid_argument("encode: level out of range");
    if (B < 1) throw std::invalid_argument("Encrypt (batched): no vectors");
    const int Nh = P.N / 2, l = P.L - level, N = P.N;
    if (slots < 1 || slots > Nh || (slots & (slots - 1))) throw std::invalid_argument("encode: slots must be a power of two <= N/2");
    if (n < 0 || n > slots) throw std::invalid_argument("encode (batched): more values than slots");
    DevFft& f = dev_fft(slots);
    const size_t each = (size_t)slots, pl = (size_t)l * N, pkl = (size_t)P.L * N, cs = 2 * pl;
    const LimbSel sel = sel_range(0, l);
    double* d = (double*)eng.alloc(2 * each * B);
    FLK_CUDA(cudaMemsetAsync(d, 0, 2 * each * B * sizeof(double), eng.stream));
    FLK_CUDA(cudaMemcpy2DAsync(d, each * sizeof(double), vals, (size_t)n * sizeof(double), (size_t)n * sizeof(double), B, cudaMemcpyHostToDevice, eng.stream));
    u64* m = eng.alloc(pl * B);                                   // message, then message + e0, coefficient form -> evaluation form
    launch_encode(eng.T, m, d, d + each * B, slots, P.sf[level], l, f.rot, f.cre, f.cim, eng.stream, 0, B);
    eng.release((u64*)d);
    auto sample = [&](u64* dst, int kind, size_t stride, bool accumulate) {
        launch_sample_limbs_csprng(eng.T, dst, rng_keys_[(int)Use::Encrypt], rng_stream_, kind, sel, eng.stream, B, stride, accumulate);
        rng_stream_ += (u64)B;
    };
    sample(m, 1, pl, true);                                       // m + e0, still in coefficient form
    eng.ntt(m, sel, B, pl);
    Elem ct = make(2, l, 1, P.sf[level], slots, B);
    u64* c0 = ct.data(); u64* c1 = c0 + pl;
    sample(c1, 0, cs, false);                                     // v, in the c1 slot of every ciphertext
    eng.ntt(c1, sel, B, cs);
    launch_ew_muladd(eng.T, c0, c1, pk_, m, sel, B, cs, 0, pl, eng.stream);               // c0 = pk0 v + (e0 + m)
    u64* e = m;                                                   // (the message buffer is free again)
    sample(e, 1, pl, false);
    eng.ntt(e, sel, B, pl);
    launch_ew_muladd(eng.T, c1, c1, pk_ + pkl, e, sel, B, cs, 0, pl, eng.stream);         // c1 = pk1 v + e1
    eng.release(m);
    return ct;
}

Elem Scheme::encrypt_with(const Elem& pt, bool seeded, u64 seed) {
    if (!pk_) throw std::runtime_error("Encrypt: no public key");
    if (pt.ncomp != 1) throw std::invalid_argument("Encrypt: plaintext expected");
    const int l = pt.l, N = P.N;
    const size_t pl = (size_t)l * N, pkl = (size_t)P.L * N;
    const LimbSel sel = sel_range(0, l);
    Elem ct = make(2, l, pt.deg, pt.scale, pt.slots);
    u64* v = eng.alloc(pl);
    u64* e = eng.alloc(pl);
    const auto dr = [&](u64 tag) { return Draw{seeded, SplitMix::sub(seed, tag), Use::Encrypt}; };
    sample_dev_to_eval(v, dr(1), 0, sel);
    sample_dev_to_eval(e, dr(2), 1, sel);
    u64* c0 = ct.data(); u64* c1 = c0 + pl;
    launch_ew(eng.T, EwOp::Mul, c0, pk_, v, sel, 1, 1, 0, 0, 0, eng.stream);
    launch_ew(eng.T, EwOp::Add, c0, c0, e, sel, 1, 1, 0, 0, 0, eng.stream);
    launch_ew(eng.T, EwOp::Add, c0, c0, pt.data(), sel, 1, 1, 0, 0, 0, eng.stream);
    sample_dev_to_eval(e, dr(3), 1, sel);
    launch_ew(eng.T, EwOp::Mul, c1, pk_ + pkl, v, sel, 1, 1, 0, 0, 0, eng.stream);
    launch_ew(eng.T, EwOp::Add, c1, c1, e, sel, 1, 1, 0, 0, 0, eng.stream);
    eng.release(v); eng.release(e);
    return ct;
}

void Scheme::decrypt(const Elem& ct_in, cplx* out, int slots) {
    if (ct_in.batch != 1) throw std::invalid_argument("Decrypt: take a slice of the batched ciphertext first");
    if (!sk_) throw std::runtime_error("Decrypt: no secret key");
    Elem ct = ct_in;
    if (ct.deg >= 2 && ct.l >= 2) rescale_inplace(ct);      // bring the message under two limbs before decoding
    const int l = ct.l;
    const size_t pl = (size_t)l * P.N;
    Elem m = make(1, l, ct.deg, ct.scale, ct.slots);
    launch_ew(eng.T, EwOp::Mul, m.data(), ct.data() + pl, sk_, sel_range(0, l), 1, 1, 0, 0, 0, eng.stream);
    launch_ew(eng.T, EwOp::Add, m.data(), m.data(), ct.data(), sel_range(0, l), 1, 1, 0, 0, 0, eng.stream);
    decode(m, out, slots > 0 ? slots : ct.slots);
}

// ---------------------------------------------------------------- level / scale plumbing
Elem Scheme::clone(const Elem& a) {
    Elem r = make(a.ncomp, a.l, a.deg, a.scale, a.slots, a.batch);
    eng.copy(r.data(), a.data(), (size_t)a.batch * a.words_each(P.N));
    return r;
}

void Scheme::drop_to(Elem& a, int l) {
    if (l == a.l) return;
    if (l > a.l || l < 1) throw std::invalid_argument("LevelReduce: bad target");
    Elem r = make(a.ncomp, l, a.deg, a.scale, a.slots, a.batch);
    FLK_CUDA(cudaMemcpy2DAsync(r.data(), (size_t)l * P.N * 8, a.data(), (size_t)a.l * P.N * 8, (size_t)l * P.N * 8, (size_t)a.ncomp * a.batch,
                               cudaMemcpyDeviceToDevice, eng.stream));
    a = r;
}
void Scheme::level_reduce_inplace(Elem& a, int levels) { if (levels > 0) drop_to(a, a.l - levels); }

void Scheme::rescale_inplace(Elem& a) {
    if (a.l < 2) throw std::runtime_error("rescale: ciphertext is at the last level");
    Elem r = make(a.ncomp, a.l - 1, a.deg - 1, a.scale / (double)P.q[a.l - 1], a.slots, a.batch);
    eng.rescale(r.data(), a.data(), a.l, a.ncomp * a.batch);
    a = r;
}
Elem Scheme::rescaled(const Elem& a) { Elem r = a; rescale_inplace(r); return r; }

ScalarSet Scheme::scalar_set(i128 k, int l) const {
    ScalarSet sc;
    for (int i = 0; i < l; ++i) {
        i128 r = k % (i128)P.q[i];
        if (r < 0) r += P.q[i];
        sc.c[i] = (u64)r; sc.c_sh[i] = nt::shoup((u64)r, P.q[i]);
    }
    return sc;
}

void Scheme::mult_int_inplace(Elem& a, i128 k) {
    Elem r = make(a.ncomp, a.l, a.deg, a.scale, a.slots, a.batch);
    launch_mul_scalar(eng.T, r.data(), a.data(), scalar_set(k, a.l), sel_range(0, a.l), a.ncomp * a.batch, eng.stream);
    a = r;
}

void Scheme::mult_scalar_core(Elem& a, double c) {
    const double sf = P.sf[level_of(a)];
    const i128 k = (i128)std::rint(c * sf);
    mult_int_inplace(a, k);
    a.deg += 1;
    a.scale *= sf;
}

// OpenFHE LeveledSHECKKSRNS::AdjustLevelsAndDepthInPlace (FLEXIBLEAUTO), restated (A.8)
void Scheme::adjust_pair(Elem& a, Elem& b) {
    const int la = level_of(a), lb = level_of(b);
    if (la == lb) {
        if (a.deg < b.deg) mult_scalar_core(a, 1.0);
        else if (b.deg < a.deg) mult_scalar_core(b, 1.0);
        return;
    }
    Elem& lo = la < lb ? a : b;          // fewer dropped limbs: must be brought down
    const Elem& hi = la < lb ? b : a;
    const int l1 = level_of(lo), l2 = level_of(hi);
    const double scf = P.sf[l1];
    const double q1 = (double)P.q[lo.l - 1];
    if (lo.deg == 2) {
        if (hi.deg == 2) {
            mult_scalar_core(lo, hi.scale / lo.scale * q1 / scf);
            rescale_inplace(lo);
            if (l1 + 1 < l2) level_reduce_inplace(lo, l2 - l1 - 1);
            lo.scale = hi.scale; lo.deg = 2;
        } else {
            if (l1 + 1 == l2) {
                rescale_inplace(lo);
            } else {
                const double scf2 = P.sf[l2 - 1] * P.sf[l2 - 1];
                mult_scalar_core(lo, scf2 / lo.scale * q1 / scf);
                rescale_inplace(lo);
                if (l1 + 2 < l2) level_reduce_inplace(lo, l2 - l1 - 2);
                rescale_inplace(lo);
                lo.scale = hi.scale; lo.deg = 1;
            }
        }
    } else {
        if (hi.deg == 2) {
            mult_scalar_core(lo, hi.scale / lo.scale / scf);
            level_reduce_inplace(lo, l2 - l1);
            lo.scale = hi.scale; lo.deg = 2;
        } else {
            const double scf2 = P.sf[l2 - 1] * P.sf[l2 - 1];
            mult_scalar_core(lo, scf2 / lo.scale / scf);
            if (l1 + 1 < l2) level_reduce_inplace(lo, l2 - l1 - 1);
            rescale_inplace(lo);
            lo.scale = hi.scale; lo.deg = 1;
        }
    }
}

void Scheme::adjust_pair_to_one(Elem& a, Elem& b) {
    adjust_pair(a, b);
    if (a.deg == 2) { rescale_inplace(a); rescale_inplace(b); }
}

// ---------------------------------------------------------------- leveled operations
Elem Scheme::binary(const Elem& a_in, const Elem& b_in, bool subtract) {
    if (a_in.ncomp == 1 && b_in.ncomp == 2 && !subtract) return binary(b_in, a_in, false);
    if (a_in.ncomp < b_in.ncomp) throw std::invalid_argument("EvalSub(plaintext, ciphertext) is not supported");
    if (a_in.batch == 1 && b_in.batch > 1 && !subtract) return binary(b_in, a_in, false);
    if (b_in.batch != 1 && b_in.batch != a_in.batch) throw std::invalid_argument("EvalAdd: batch sizes differ");
    Elem a = a_in, b = b_in;
    adjust_pair(a, b);
    Elem r = make(a.ncomp, a.l, a.deg, a.scale, a.slots, a.batch);
    const size_t pl = (size_t)a.l * P.N, ca = a.ncomp * pl;
    const LimbSel sel = sel_range(0, a.l);
    const EwOp op = subtract ? EwOp::Sub : EwOp::Add;
    if (b.ncomp == a.ncomp) {   // batch b against batch b, or one operand broadcast over the batch
        launch_ew(eng.T, op, r.data(), a.data(), b.data(), sel, a.ncomp, a.batch, ca, b.batch > 1 ? ca : 0, pl, eng.stream);
    } else {                    // ciphertext + plaintext: only c0 changes
        launch_ew(eng.T, op, r.data(), a.data(), b.data(), sel, 1, a.batch, ca, 0, 0, eng.stream);
        FLK_CUDA(cudaMemcpy2DAsync(r.data() + pl, ca * 8, a.data() + pl, ca * 8, pl * (a.ncomp - 1) * 8, a.batch, cudaMemcpyDeviceToDevice, eng.stream));
    }
    if (eng.ledger_on) eng.ledger.add("add", a.l, 48.0 * P.N * a.l, a.batch);
    return r;
}
Elem Scheme::add(const Elem& a, const Elem& b) { return binary(a, b, false); }
Elem Scheme::sub(const Elem& a, const Elem& b) { return binary(a, b, true); }

Elem Scheme::add_many(std::vector<Elem> v) {
    if (v.empty()) throw std::invalid_argument("EvalAddMany: empty input");
    const size_t n = v.size();
    for (size_t j = 1; j < n; j *= 2)
        for (size_t i = 0; i + j < n; i += 2 * j) v[i] = add(v[i], v[i + j]);
    return v[0];
}

Elem Scheme::add_const(const Elem& a, double c) {
    // constant polynomial round(c * scale): every evaluation-format entry of c0 receives the same residue
    const i128 k = (i128)std::rint(c * a.scale);
    Elem r = clone(a);
    launch_add_scalar(eng.T, r.data(), a.data(), scalar_set(k, a.l), sel_range(0, a.l), a.batch, a.words_each(P.N), eng.stream);
    return r;
}

Elem Scheme::mult(const Elem& a_in, const Elem& b_in) {
    if (a_in.ncomp == 1 && b_in.ncomp == 2) return mult(b_in, a_in);
    if (a_in.ncomp != 2) throw std::invalid_argument("EvalMult: ciphertext expected");
    if (a_in.batch == 1 && b_in.batch > 1) return mult(b_in, a_in);
    if (b_in.batch != 1 && b_in.batch != a_in.batch) throw std::invalid_argument("EvalMult: batch sizes differ");
    Elem a = a_in, b = b_in;
    adjust_pair_to_one(a, b);
    Elem r = make(2, a.l, a.deg + b.deg, a.scale * b.scale, a.slots, a.batch);
    if (b.ncomp == 2) {
        if (!mk_) throw std::runtime_error("EvalMult: relinearisation key missing");
        for (int b0 = 0, mb = max_batch(a.l); b0 < a.batch; b0 += mb) {
            const size_t o = (size_t)b0 * a.words_each(P.N);
            eng.mul_relin_batch(r.data() + o, a.data() + o, b.data() + (b.batch > 1 ? o : 0), a.l, mk_, std::min(mb, a.batch - b0),
                                b.batch > 1 ? a.words_each(P.N) : 0);
        }
    } else {
        const size_t pl = (size_t)a.l * P.N;
        launch_ew(eng.T, EwOp::Mul, r.data(), a.data(), b.data(), sel_range(0, a.l), 2, a.batch, 2 * pl, 0, 0, eng.stream);
        if (eng.ledger_on) eng.ledger.add("mul_plain", a.l, 40.0 * P.N * a.l, a.batch);
    }
    return r;
}

Elem Scheme::mult_const(const Elem& a_in, double c) {
    Elem a = a_in;
    if (a.deg == 2) rescale_inplace(a);
    mult_scalar_core(a, c);
    return a;
}

Elem Scheme::linear_wsum(const Elem& in_raw, const double* w, int n_out) {
    if (in_raw.ncomp != 2) throw std::invalid_argument("EvalLinearWSum: ciphertexts expected");
    Elem in = in_raw;
    if (in.deg == 2) rescale_inplace(in);
    const int n_in = in.batch, l = in.l;
    const double sf = P.sf[level_of(in)];
    std::vector<u64> k((size_t)n_out * n_in * l * 2);
    for (int o = 0; o < n_out; ++o)
        for (int t = 0; t < n_in; ++t) {
            const i128 v = (i128)std::rint(w[(size_t)o * n_in + t] * sf);
            for (int i = 0; i < l; ++i) {
                i128 r = v % (i128)P.q[i];
                if (r < 0) r += P.q[i];
                u64* e = &k[(((size_t)o * n_in + t) * l + i) * 2];
                e[0] = (u64)r; e[1] = nt::shoup((u64)r, P.q[i]);
            }
        }
    u64* kd = eng.alloc(k.size());
    eng.upload(kd, k.data(), k.size());
    Elem r = make(2, l, in.deg + 1, in.scale * sf, in.slots, n_out);
    launch_lincomb(eng.T, r.data(), in.data(), kd, l, 2 * l, n_in, n_out, eng.stream);
    eng.sync();   // k is a host temporary
    eng.release(kd);
    if (eng.ledger_on) eng.ledger.add("linear_wsum", l, (double)(n_in + n_out) * 16.0 * l * P.N);
    return r;
}

// sum_i w[i] * terms[i] for terms of identical level / degree 1 / scale (each possibly a batched operand): one kernel.
// The result is one degree deeper, like EvalMult by a scalar.
Elem Scheme::weighted_sum(const std::vector<Elem>& terms, const std::vector<double>& w) { return weighted_sums(terms, {w}); }

// n_out plaintext-weighted sums of the same aligned terms as ONE kernel: out[o] = sum_t w[o][t] terms[t].  The terms are read
// where they lie (a table of pointers travels with the scalars), the result is one batched operand: combination o is elements
// [o B, (o + 1) B) of it, B the batch of a term.
Elem Scheme::weighted_sums(const std::vector<Elem>& terms, const std::vector<std::vector<double>>& w) {
    const Elem& f = terms.at(0);
    const int n_in = (int)terms.size(), n_out = (int)w.size(), l = f.l, rows = 2 * l * f.batch;
    if (n_out < 1) throw std::invalid_argument("weighted_sums: no combinations");
    const double sf = P.sf[level_of(f)];
    const size_t nk = (size_t)n_out * n_in * l * 2;
    std::vector<u64> k(nk + (size_t)n_in);               // scalars with Shoup companions, then the pointer table
    for (int t = 0; t < n_in; ++t) {
        if (terms[t].ncomp != 2 || terms[t].l != l || terms[t].deg != 1 || terms[t].scale != f.scale || terms[t].batch != f.batch)
            throw std::invalid_argument("weighted_sum: misaligned terms");
        k[nk + (size_t)t] = (u64)(uintptr_t)terms[t].data();
    }
    for (int o = 0; o < n_out; ++o) {
        if ((int)w[o].size() != n_in) throw std::invalid_argument("weighted_sums: one weight per term expected");
        for (int t = 0; t < n_in; ++t) {
            const i128 v = (i128)std::rint(w[o][t] * sf);
            for (int i = 0; i < l; ++i) {
                i128 r = v % (i128)P.q[i];
                if (r < 0) r += P.q[i];
                const size_t at = (((size_t)o * n_in + t) * l + i) * 2;
                k[at] = (u64)r; k[at + 1] = nt::shoup((u64)r, P.q[i]);
            }
        }
    }
    u64* kd = eng.alloc(k.size());
    upload_small(kd, k.data(), k.size());
    Elem r = make(2, l, 2, f.scale * sf, f.slots, f.batch * n_out);
    launch_lincomb(eng.T, r.data(), nullptr, kd, l, rows, n_in, n_out, eng.stream, reinterpret_cast<const u64* const*>(kd + nk));
    eng.release(kd);
    return r;
}

Elem Scheme::mult_many(std::vector<Elem> v) {
    const size_t n = v.size();
    if (n == 0) throw std::invalid_argument("EvalMultMany: empty input");
    if (n == 1) return v[0];
    std::vector<Elem> m(n - 1);
    size_t ctr = 0;
    for (size_t i = 0; i < 2 * n - 2; i += 2) {
        const Elem& x = i < n ? v[i] : m[i - n];
        const Elem& y = i + 1 < n ? v[i + 1] : m[i + 1 - n];
        m[ctr++] = mult(x, y);
    }
    return m.back();
}

Elem Scheme::apply_galois(const Elem& a, uint32_t g) {
    if (a.ncomp != 2) throw std::invalid_argument("EvalAutomorphism: ciphertext expected");
    auto it = gk_.find(g);
    if (it == gk_.end()) throw std::runtime_error("EvalAutomorphism: no evaluation key for Galois element " + std::to_string(g));
    Elem r = make(2, a.l, a.deg, a.scale, a.slots, a.batch);
    for (int b0 = 0, mb = max_batch(a.l); b0 < a.batch; b0 += mb) {
        const size_t o = (size_t)b0 * a.words_each(P.N);
        eng.rotate_batch(r.data() + o, a.data() + o, a.l, g, it->second, std::min(mb, a.batch - b0), false);
    }
    return r;
}
Elem Scheme::rotate(const Elem& a, int k) {
    if (k == 0) return clone(a);
    return apply_galois(a, P.galois_for_rotation(k));
}
Elem Scheme::conjugate(const Elem& a) { return apply_galois(a, P.galois_conj()); }

// How many doubling steps each hoisted key switch of a ladder takes.  A group of g steps costs one ModUp + ModDown (measured
// ~ 12 evaluation-key products, scripts/prof_rotsum.py) plus 2^g - 1 key products, so the cheapest partition of `steps` is
// found by dynamic programming over group sizes 1 .. 4 (15 rotations); smaller groups first.  FLK_HOIST_GMAX overrides the cap.
std::vector<int> Scheme::ladder_plan(int steps) {
    static const int gmax = [] { const char* e = std::getenv("FLK_HOIST_GMAX"); return e ? std::max(1, std::min(4, std::atoi(e))) : 4; }();
    const long C = 12;
    std::vector<long> best(steps + 1, 0);
    std::vector<int> pick(steps + 1, 0);
    for (int s = 1; s <= steps; ++s) {
        best[s] = -1;
        for (int g = 1; g <= std::min(gmax, s); ++g) {
            const long c = best[s - g] + C + ((1L << g) - 1);
            if (best[s] < 0 || c < best[s]) { best[s] = c; pick[s] = g; }
        }
    }
    std::vector<int> groups;
    for (int s = steps; s > 0; s -= pick[s]) groups.push_back(pick[s]);
    std::sort(groups.begin(), groups.end());
    return groups;
}
// rotation amounts a ladder uses with that plan (the doubling keys stride 2^i first)
std::vector<int> Scheme::ladder_rotations(int steps, int stride) {
    std::vector<int> r;
    for (int i = 0; i < steps; ++i) r.push_back(stride * (1 << i));
    int i = 0;
    for (int g : ladder_plan(steps)) {
        for (int t = 3; t < (1 << g); ++t)
            if (t & (t - 1)) r.push_back(t * stride * (1 << i));
        i += g;
    }
    return r;
}

// r <- r + rot(r, stride 2^i), i = 0 .. steps-1: the rotate-and-add ladders of FHEController::rotsum / repeat
// (F.cpp:829-867) as one call; the running ciphertext never leaves the device.
Elem Scheme::rotsum(const Elem& a, int steps, int stride) {
    if (a.ncomp != 2) throw std::invalid_argument("rotsum: ciphertext expected");
    static const bool no_hoist = [] { const char* e = std::getenv("FLK_NO_HOIST"); return e && e[0] == '1'; }();
    Elem r = a;
    auto key_of = [&](int k) -> const u64* {
        auto it = gk_.find(P.galois_for_rotation(k));
        return it == gk_.end() ? nullptr : it->second;
    };
    const std::vector<int> plan = ladder_plan(steps);
    size_t pi = 0;
    int planned_left = 0;
    for (int i = 0; i < steps;) {
        const int k = stride * (1 << i);
        if (!key_of(k)) throw std::runtime_error("rotsum: no evaluation key for rotation " + std::to_string(k));
        // g doubling steps at once: r + sum_{t = 1 .. 2^g - 1} rot(r, t k), all rotations hoisted on one ModUp / one ModDown.
        // The group sizes come from ladder_plan(); a group whose extra keys (t k, t not a power of two) are missing is
        // taken in smaller groups.
        if (planned_left == 0 && pi < plan.size()) planned_left = plan[pi++];
        int g = 1;
        if (!no_hoist) {
            auto have = [&](int gg) {
                for (int t = 1; t < (1 << gg); ++t)
                    if (!key_of(t * k)) return false;
                return true;
            };
            g = std::max(1, std::min(planned_left, steps - i));
            while (g > 1 && !have(g)) --g;
        }
        planned_left = std::max(0, planned_left - g);
        const int nk = (1 << g) - 1;
        uint32_t gs[kHoistMax];
        const u64* evks[kHoistMax];
        for (int t = 1; t <= nk; ++t) { gs[t - 1] = P.galois_for_rotation(t * k); evks[t - 1] = key_of(t * k); }
        Elem nx = make(2, r.l, r.deg, r.scale, r.slots, r.batch);
        for (int b0 = 0, mb = max_batch(r.l); b0 < r.batch; b0 += mb) {
            const size_t o = (size_t)b0 * r.words_each(P.N);
            const int nb = std::min(mb, r.batch - b0);
            if (g > 1) eng.rotate_sum_batch(nx.data() + o, r.data() + o, r.l, gs, evks, nk, nb, true);
            else eng.rotate_batch(nx.data() + o, r.data() + o, r.l, gs[0], evks[0], nb, true);
        }
        r = nx;
        i += g;
    }
    return steps == 0 ? clone(a) : r;
}

// EvalBootstrap(ct, numIterations = 2, precision) (F.cpp:461): bootstrap, then bootstrap the 2^precision-amplified
// residual error and subtract it
Elem Scheme::bootstrap_iter(const Elem& ct, int iterations, int precision) {
    Elem first = bootstrap(ct);
    if (iterations <= 1) return first;
    Elem err = sub(ct, first);                      // aligned to ct's (deeper) level by the FLEXIBLEAUTO adjustment
    mult_int_inplace(err, (i128)1 << precision);    // integer scaling: no level is spent
    Elem eb = bootstrap(err);
    Elem fix = mult_const(eb, std::ldexp(1.0, -precision));
    return add(first, fix);
}

}  // namespace flk
