// capi.cpp -- extern "C" boundary (include/fl_ckks.h) over the device engine.
#include "../../include/fl_ckks.h"

#include <chrono>
#include <cstring>
#include <map>
#include <string>

#include <cstdio>
#include <array>
#include <vector>

#include "scheme.h"

using namespace flk;

struct ProfRec { const char* name; cudaEvent_t a, b; double host_s; };
struct fl_ctx {
    Scheme* sch;
    Engine* eng;   // = &sch->eng
    bool prof_on = false;
    std::vector<ProfRec> prof;
};
// GPU time of one C-ABI call: events on the engine stream around everything the call enqueues (fl_prof_*)
struct ProfScope {
    fl_ctx* c; ProfRec r; std::chrono::steady_clock::time_point t0;
    ProfScope(fl_ctx* ctx, const char* name) : c(ctx && ctx->prof_on ? ctx : nullptr) {
        // the current device is per host thread: a caller that drives several contexts (or one context from a worker thread)
        // lands on the engine's device here, whichever device its thread last used
        if (ctx && ctx->eng) cudaSetDevice(ctx->eng->device_id);
        if (!c) return;
        r.name = name;
        cudaEventCreate(&r.a); cudaEventCreate(&r.b);
        cudaEventRecord(r.a, c->eng->stream);
        t0 = std::chrono::steady_clock::now();
    }
    ~ProfScope() {
        if (!c) return;
        cudaEventRecord(r.b, c->eng->stream);
        r.host_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        c->prof.push_back(r);
    }
};
struct fl_elem {
    Elem e;
};

static thread_local std::string g_err;

const char* fl_last_error(void) { return g_err.c_str(); }

#define FL_TRY(...)                      \
    try {                                \
        ProfScope prof_scope_(prof_ctx(c), __func__); \
        __VA_ARGS__;                     \
        return 0;                        \
    } catch (const std::exception& e) {  \
        g_err = e.what();                \
        return 1;                        \
    } catch (...) {                      \
        g_err = "unknown error";         \
        return 2;                        \
    }

#define FL_TRY0(...)                      \
    try {                                \
        __VA_ARGS__;                     \
        return 0;                        \
    } catch (const std::exception& e) {  \
        g_err = e.what();                \
        return 1;                        \
    } catch (...) {                      \
        g_err = "unknown error";         \
        return 2;                        \
    }

static inline fl_ctx* prof_ctx(fl_ctx* c) { return c; }
static inline fl_ctx* prof_ctx(const void*) { return nullptr; }   // entry points without a context argument are not timed

static LimbSel make_sel(const int* midx, int nl) {
    if (nl < 0 || nl > kMaxLimbSel) throw std::invalid_argument("limb count out of range");
    LimbSel s;
    for (int i = 0; i < nl; ++i) s.push(midx[i], i);
    return s;
}

extern "C" {

int fl_ctx_create(const fl_params* p, int device, fl_ctx** out) {
    FL_TRY0({
        ParamSpec s;
        s.logN = p->logN; s.L = p->L; s.dnum = p->dnum; s.first_bits = p->first_bits; s.scale_bits = p->scale_bits;
        s.aux_bits = p->aux_bits; s.sparse_h = p->sparse_h;
        fl_ctx* c = new fl_ctx{nullptr, nullptr};
        try { c->sch = new Scheme(s, device); } catch (...) { delete c; throw; }
        c->eng = &c->sch->eng;
        *out = c;
    })
}
void fl_ctx_destroy(fl_ctx* c) {
    if (!c) return;
    delete c->sch;
    delete c;
}
int fl_ctx_info(fl_ctx* c, int* info) {
    FL_TRY({ const Params& P = c->eng->P; info[0] = P.logN; info[1] = P.L; info[2] = P.K; info[3] = P.alpha; info[4] = P.dnum;
             info[5] = (int)c->eng->pool_allocs; info[6] = (int)c->eng->cache_trims; info[7] = (int)(c->eng->cached_bytes() >> 20); })
}
int fl_ctx_moduli(fl_ctx* c, uint64_t* out) { FL_TRY(std::memcpy(out, c->eng->P.q.data(), 8 * c->eng->P.T)) }
int fl_ctx_roots(fl_ctx* c, uint64_t* out) { FL_TRY(std::memcpy(out, c->eng->P.psi.data(), 8 * c->eng->P.T)) }
int fl_ctx_scale_factors(fl_ctx* c, double* out) { FL_TRY(std::memcpy(out, c->eng->P.sf.data(), 8 * c->eng->P.L)) }
uint32_t fl_galois_for_rotation(fl_ctx* c, int k) { return c->eng->P.galois_for_rotation(k); }
uint32_t fl_galois_conj(fl_ctx* c) { return c->eng->P.galois_conj(); }
void* fl_ctx_stream(fl_ctx* c) { return (void*)c->eng->stream; }
int fl_sync(fl_ctx* c) { FL_TRY(c->eng->sync()) }
int fl_ctx_set_cache_bytes(fl_ctx* c, uint64_t bytes) { FL_TRY(c->eng->set_cache_cap((size_t)bytes)) }

int fl_dev_alloc(fl_ctx* c, size_t words, uint64_t** out) { FL_TRY(*out = c->eng->alloc(words)) }
int fl_dev_free(fl_ctx* c, uint64_t* p) { FL_TRY(c->eng->release(p)) }
int fl_dev_upload(fl_ctx* c, uint64_t* dst, const uint64_t* src, size_t words) { FL_TRY(c->eng->upload(dst, src, words)) }
int fl_dev_download(fl_ctx* c, uint64_t* dst, const uint64_t* src, size_t words) { FL_TRY(c->eng->download(dst, src, words)) }

int fl_raw_ntt(fl_ctx* c, uint64_t* d, const int* midx, int nl) { FL_TRY(c->eng->ntt(d, make_sel(midx, nl))) }
int fl_raw_intt(fl_ctx* c, uint64_t* d, const int* midx, int nl) { FL_TRY(c->eng->intt(d, make_sel(midx, nl))) }
int fl_raw_ntt_batch(fl_ctx* c, uint64_t* d, const int* midx, int nl, int batch, int inverse) {
    FL_TRY({
        const LimbSel sel = make_sel(midx, nl);
        const size_t bs = (size_t)nl * c->eng->P.N;
        if (inverse) c->eng->intt(d, sel, batch, bs); else c->eng->ntt(d, sel, batch, bs);
    })
}
int fl_raw_add(fl_ctx* c, uint64_t* out, const uint64_t* a, const uint64_t* b, const int* midx, int nl) {
    FL_TRY(c->eng->ew_sel(EwOp::Add, out, a, b, make_sel(midx, nl)))
}
int fl_raw_sub(fl_ctx* c, uint64_t* out, const uint64_t* a, const uint64_t* b, const int* midx, int nl) {
    FL_TRY(c->eng->ew_sel(EwOp::Sub, out, a, b, make_sel(midx, nl)))
}
int fl_raw_mul(fl_ctx* c, uint64_t* out, const uint64_t* a, const uint64_t* b, const int* midx, int nl) {
    FL_TRY(c->eng->ew_sel(EwOp::Mul, out, a, b, make_sel(midx, nl)))
}
int fl_raw_automorph(fl_ctx* c, uint64_t* out, const uint64_t* in, int nl, uint32_t g) { FL_TRY(c->eng->automorph(out, in, g, nl)) }
int fl_raw_rescale(fl_ctx* c, uint64_t* out, const uint64_t* in, int l, int polys) { FL_TRY(c->eng->rescale(out, in, l, polys)) }
int fl_raw_modup(fl_ctx* c, uint64_t* out_ext, const uint64_t* c_eval, int l, int digit) { FL_TRY(c->eng->modup(out_ext, c_eval, l, digit)) }
int fl_raw_moddown(fl_ctx* c, uint64_t* out, const uint64_t* in_ext, int l) { FL_TRY(c->eng->moddown(out, in_ext, l)) }
int fl_raw_keyswitch(fl_ctx* c, uint64_t* out2, const uint64_t* poly, const uint64_t* evk, int l) {
    FL_TRY(c->eng->keyswitch(out2, poly, evk, l, nullptr, nullptr, 0))
}
int fl_raw_rotate(fl_ctx* c, uint64_t* out, const uint64_t* ct, int l, uint32_t g, const uint64_t* evk) {
    FL_TRY(c->eng->rotate(out, ct, l, g, evk))
}
int fl_raw_rotate_batch(fl_ctx* c, uint64_t* out, const uint64_t* ct, int l, uint32_t g, const uint64_t* evk, int batch) {
    FL_TRY(c->eng->rotate_batch(out, ct, l, g, evk, batch, false))
}
int fl_raw_mul_relin(fl_ctx* c, uint64_t* out, const uint64_t* a, const uint64_t* b, int l, const uint64_t* evk) {
    FL_TRY(c->eng->mul_relin(out, a, b, l, evk))
}
int fl_raw_mul_plain(fl_ctx* c, uint64_t* out, const uint64_t* ct, const uint64_t* pt, int l) {
    FL_TRY(c->eng->ew(EwOp::Mul, out, ct, pt, l, 2, true))
}

int fl_raw_ks_digits(fl_ctx* c, uint64_t* dco, const uint64_t* poly, int l) { FL_TRY(c->eng->ks_digits(dco, poly, l)) }
int fl_raw_ks_digits_part(fl_ctx* c, uint64_t* dco, const uint64_t* poly, int l, int first, int count) {
    FL_TRY(c->eng->ks_digits_part(dco, poly, l, first, count))
}
int fl_raw_ks_modup(fl_ctx* c, uint64_t* up, const uint64_t* dco, int l, int first, int count) { FL_TRY(c->eng->ks_modup_part(up, dco, l, first, count)) }
int fl_raw_ks_inner(fl_ctx* c, uint64_t* acc, const uint64_t* up, const uint64_t* poly, const uint64_t* evk, int l, int first, int count) {
    FL_TRY(c->eng->ks_inner_part(acc, up, poly, evk, l, first, count))
}
int fl_raw_ks_pcoef(fl_ctx* c, uint64_t* acc, int l, int first, int count) { FL_TRY(c->eng->ks_pcoef_part(acc, l, first, count)) }
int fl_raw_ks_moddown(fl_ctx* c, uint64_t* out, uint64_t* tq, const uint64_t* acc, int l, int first, int count, const uint64_t* add0,
                      const uint64_t* add1, uint32_t g) {
    FL_TRY(c->eng->ks_moddown_part(out, tq, acc, l, first, count, add0, add1, g))
}
int fl_host_ntt(fl_ctx* c, uint64_t* poly_host, int l, int inverse) {
    FL_TRY({
        Engine& e = *c->eng;
        const size_t w = (size_t)l * e.P.N;
        u64* d = e.alloc(w);
        e.upload(d, poly_host, w);
        if (inverse) e.intt(d, sel_range(0, l)); else e.ntt(d, sel_range(0, l));
        e.download(poly_host, d, w);
        e.release(d);
    })
}
int fl_host_rotate(fl_ctx* c, uint64_t* out_host, const uint64_t* ct_host, int l, uint32_t g, const uint64_t* evk_dev) {
    FL_TRY({
        Engine& e = *c->eng;
        const size_t w = (size_t)2 * l * e.P.N;
        u64* in = e.alloc(w); u64* out = e.alloc(w);
        e.upload(in, ct_host, w);
        e.rotate(out, in, l, g, evk_dev);
        e.download(out_host, out, w);
        e.release(in); e.release(out);
    })
}
// ciphertexts per pipeline stage of the host-operand calls (FLK_HOST_CHUNK overrides).  Measured (scripts/e2e_sweep.py): a call that
// waits for its results is shortest with one ciphertext per stage (smallest drain), overlapped calls with two.
static int host_chunk(bool wait) {
    static const int v = [] { const char* e = std::getenv("FLK_HOST_CHUNK"); return e ? std::max(1, std::atoi(e)) : 0; }();
    return v ? v : (wait ? 1 : 2);
}
int fl_host_rotate_batch(fl_ctx* c, uint64_t* out_host, const uint64_t* ct_host, int l, uint32_t g, const uint64_t* evk_dev, int batch) {
    FL_TRY(c->eng->rotate_batch_host(out_host, ct_host, l, g, evk_dev, batch, host_chunk(true), true))
}
int fl_host_rotate_batch_async(fl_ctx* c, uint64_t* out_host, const uint64_t* ct_host, int l, uint32_t g, const uint64_t* evk_dev, int batch) {
    FL_TRY(c->eng->rotate_batch_host(out_host, ct_host, l, g, evk_dev, batch, host_chunk(false), false))
}
int fl_host_mul_relin(fl_ctx* c, uint64_t* out_host, const uint64_t* a_host, const uint64_t* b_host, int l, const uint64_t* evk_dev) {
    FL_TRY({
        Engine& e = *c->eng;
        const size_t w = (size_t)2 * l * e.P.N;
        u64* a = e.alloc(w); u64* b = e.alloc(w); u64* out = e.alloc(w);
        e.upload(a, a_host, w); e.upload(b, b_host, w);
        e.mul_relin(out, a, b, l, evk_dev);
        e.download(out_host, out, w);
        e.release(a); e.release(b); e.release(out);
    })
}


// ======================= scheme level =======================
static fl_elem* wrap(Elem&& e) { return new fl_elem{std::move(e)}; }

int fl_keygen(fl_ctx* c, uint64_t seed) { FL_TRY(c->sch->keygen(seed)) }
int fl_keygen_seeded(fl_ctx* c, uint64_t seed) {
    FL_TRY(if (seed == 0) throw std::invalid_argument("fl_keygen_seeded: the seed must be non-zero (0 selects operating-system randomness)"); c->sch->keygen(seed))
}
int fl_gen_mult_key(fl_ctx* c) { FL_TRY(c->sch->gen_mult_key()) }
int fl_gen_rot_keys(fl_ctx* c, const int* idx, int n) { FL_TRY(for (int i = 0; i < n; ++i) c->sch->gen_rotation_key(idx[i])) }
int fl_gen_conj_key(fl_ctx* c) { FL_TRY(c->sch->gen_galois_key(c->sch->P.galois_conj())) }
int fl_keys_clear(fl_ctx* c, int kind) { FL_TRY(if (kind == 0) c->sch->clear_rotation_keys(); else c->sch->clear_mult_key()) }
int fl_num_rot_keys(fl_ctx* c) { return (int)c->sch->num_galois_keys(); }
double fl_rot_key_bytes(fl_ctx* c) { return (double)c->sch->num_galois_keys() * (double)c->sch->eng.evk_words() * 8.0; }
int fl_export_sk(fl_ctx* c, uint64_t* out) { FL_TRY(c->sch->export_sk(out)) }
int fl_export_pk(fl_ctx* c, uint64_t* out) { FL_TRY(c->sch->export_pk(out)) }
int fl_export_evk(fl_ctx* c, uint32_t g, uint64_t* out) { FL_TRY(c->sch->export_evk(g, out)) }
int fl_import_keys(fl_ctx* c, const uint64_t* sk, const uint64_t* pk) { FL_TRY(c->sch->import_keys(sk, pk)) }
int fl_import_evk(fl_ctx* c, uint32_t g, const uint64_t* evk) { FL_TRY(c->sch->import_evk(g, evk)) }
int fl_keys_save(fl_ctx* c, const char* path) { FL_TRY(c->sch->save_keys(path, 15)) }
int fl_keys_save_sel(fl_ctx* c, const char* path, int what) { FL_TRY(c->sch->save_keys(path, what)) }
int fl_keys_load(fl_ctx* c, const char* path) { FL_TRY(c->sch->load_keys(path)) }

int fl_encode(fl_ctx* c, const double* re, const double* im, int n, int level, int slots, fl_pt** out) {
    FL_TRY({
        std::vector<cplx> v(n);
        for (int i = 0; i < n; ++i) v[i] = cplx(re[i], im ? im[i] : 0.0);
        *out = wrap(c->sch->encode(v.data(), n, level, slots, 1));
    })
}
int fl_encode_many(fl_ctx* c, const double* re, int count, int n, int level, int slots, fl_pt** out) {
    FL_TRY(*out = wrap(c->sch->encode_many_real(re, count, n, level, slots)))
}
int fl_encrypt_values_many(fl_ctx* c, const double* re, int count, int n, int level, int slots, fl_ct** out) {
    FL_TRY(*out = wrap(c->sch->encrypt_values_many(re, count, n, level, slots)))
}
int fl_encrypt(fl_ctx* c, const fl_pt* p, fl_ct** out) { FL_TRY(*out = wrap(c->sch->encrypt(p->e))) }
int fl_encrypt_many(fl_ctx* c, const fl_pt* const* pts, int n, fl_ct** out) {
    FL_TRY({
        std::vector<const Elem*> v((size_t)n);
        for (int i = 0; i < n; ++i) v[i] = &pts[i]->e;
        *out = wrap(c->sch->encrypt_many(v));
    })
}
int fl_encrypt_seeded(fl_ctx* c, const fl_pt* p, uint64_t seed, fl_ct** out) { FL_TRY(*out = wrap(c->sch->encrypt_seeded(p->e, seed))) }
static void split(const std::vector<cplx>& v, double* re, double* im) {
    for (size_t i = 0; i < v.size(); ++i) { re[i] = v[i].real(); if (im) im[i] = v[i].imag(); }
}
int fl_decrypt(fl_ctx* c, const fl_ct* a, double* re, double* im, int slots) {
    FL_TRY({ std::vector<cplx> v(slots); c->sch->decrypt(a->e, v.data(), slots); split(v, re, im); })
}
int fl_decode(fl_ctx* c, const fl_pt* p, double* re, double* im, int slots) {
    FL_TRY({ std::vector<cplx> v(slots); c->sch->decode(p->e, v.data(), slots); split(v, re, im); })
}
int fl_add(fl_ctx* c, const fl_elem* a, const fl_elem* b, fl_ct** out) { FL_TRY(*out = wrap(c->sch->add(a->e, b->e))) }
int fl_sub(fl_ctx* c, const fl_elem* a, const fl_elem* b, fl_ct** out) { FL_TRY(*out = wrap(c->sch->sub(a->e, b->e))) }
int fl_add_many(fl_ctx* c, fl_ct* const* v, int n, fl_ct** out) {
    FL_TRY({ std::vector<Elem> e(n); for (int i = 0; i < n; ++i) e[i] = v[i]->e; *out = wrap(c->sch->add_many(std::move(e))); })
}
int fl_add_const(fl_ctx* c, const fl_ct* a, double k, fl_ct** out) { FL_TRY(*out = wrap(c->sch->add_const(a->e, k))) }
int fl_mul(fl_ctx* c, const fl_elem* a, const fl_elem* b, fl_ct** out) { FL_TRY(*out = wrap(c->sch->mult(a->e, b->e))) }
int fl_mul_const(fl_ctx* c, const fl_ct* a, double k, fl_ct** out) { FL_TRY(*out = wrap(c->sch->mult_const(a->e, k))) }
int fl_mul_many(fl_ctx* c, fl_ct* const* v, int n, fl_ct** out) {
    FL_TRY({ std::vector<Elem> e(n); for (int i = 0; i < n; ++i) e[i] = v[i]->e; *out = wrap(c->sch->mult_many(std::move(e))); })
}
int fl_linear_wsum(fl_ctx* c, const fl_ct* in, const double* w, int n_out, fl_ct** out) { FL_TRY(*out = wrap(c->sch->linear_wsum(in->e, w, n_out))) }
int fl_rotate(fl_ctx* c, const fl_ct* a, int k, fl_ct** out) { FL_TRY(*out = wrap(c->sch->rotate(a->e, k))) }
int fl_has_rot_key(fl_ctx* c, int k) { return c->sch->has_rotation_key(k) ? 1 : 0; }
int fl_rotsum_rotations(int steps, int stride, int* out, int cap) {
    const std::vector<int> r = Scheme::ladder_rotations(steps, stride);
    for (int i = 0; i < (int)r.size() && i < cap; ++i) out[i] = r[i];
    return (int)r.size();
}
int fl_rotsum(fl_ctx* c, const fl_ct* a, int steps, int stride, fl_ct** out) { FL_TRY(*out = wrap(c->sch->rotsum(a->e, steps, stride))) }
int fl_bootstrap_iter(fl_ctx* c, const fl_ct* a, int iterations, int precision, fl_ct** out) {
    FL_TRY(*out = wrap(c->sch->bootstrap_iter(a->e, iterations, precision)))
}
int fl_conjugate(fl_ctx* c, const fl_ct* a, fl_ct** out) { FL_TRY(*out = wrap(c->sch->conjugate(a->e))) }
int fl_rescale(fl_ctx* c, const fl_ct* a, fl_ct** out) { FL_TRY(*out = wrap(c->sch->rescaled(a->e))) }
int fl_eval_poly(fl_ctx* c, const fl_ct* a, const double* coeffs, int n, fl_ct** out) {
    FL_TRY(*out = wrap(c->sch->eval_poly(a->e, std::vector<double>(coeffs, coeffs + n))))
}
int fl_eval_chebyshev(fl_ctx* c, const fl_ct* x, const double* coeffs, int n, double a, double b, fl_ct** out) {
    FL_TRY(*out = wrap(c->sch->eval_chebyshev(x->e, std::vector<double>(coeffs, coeffs + n), a, b)))
}
int fl_chebyshev_coefficients(double (*f)(double, void*), void* user, double a, double b, int degree, double* out) {
    FL_TRY0({ auto v = Scheme::chebyshev_coefficients(f, user, a, b, degree); std::memcpy(out, v.data(), 8 * v.size()); })
}
struct fl_lt {
    LinTrans t;
};
int fl_lt_create(fl_ctx* c, const int* shifts, int ndiag, const double* re, const double* im, int slots, int level, int max_baby, fl_lt** out) {
    FL_TRY({
        std::map<int, std::vector<cplx>> d;
        for (int k = 0; k < ndiag; ++k) {
            std::vector<cplx> v(slots);
            for (int p = 0; p < slots; ++p) v[p] = cplx(re[(size_t)k * slots + p], im ? im[(size_t)k * slots + p] : 0.0);
            d[shifts[k]] = std::move(v);
        }
        auto* h = new fl_lt{c->sch->lintrans_plan(d, slots, max_baby)};
        if (level >= 0) c->sch->lintrans_encode(h->t, level);
        *out = h;
    })
}
int fl_lt_rotations(fl_ctx* c, const fl_lt* t, int* out, int cap) {
    const std::vector<int> r = c->sch->lintrans_rotations(t->t);
    for (int i = 0; i < (int)r.size() && i < cap; ++i) out[i] = r[i];
    return (int)r.size();
}
int fl_lt_shape(const fl_lt* t, int* n1, int* n2, int* stride, int* ndiag) {
    *n1 = t->t.n1; *n2 = t->t.n2; *stride = t->t.g; *ndiag = t->t.ndiag;
    return 0;
}
int fl_lt_apply(fl_ctx* c, fl_lt* t, const fl_ct* a, fl_ct** out) { FL_TRY(*out = wrap(c->sch->lintrans_apply(t->t, a->e))) }
int fl_lt_apply_plain(fl_ctx* c, fl_lt* t, const fl_ct* a, fl_ct** out) { FL_TRY(*out = wrap(c->sch->lintrans_apply_plain(t->t, a->e))) }
void fl_lt_free(fl_lt* t) { delete t; }

int fl_bootstrap_setup(fl_ctx* c, int b0, int b1, int slots) { FL_TRY(c->sch->bootstrap_setup(b0, b1, slots)) }
int fl_bootstrap_keygen(fl_ctx* c, int slots) { FL_TRY(c->sch->bootstrap_keygen(slots)) }
int fl_bootstrap(fl_ctx* c, const fl_ct* a, fl_ct** out) { FL_TRY(*out = wrap(c->sch->bootstrap(a->e))) }

int fl_elem_level(const fl_elem* a) { return a->e.valid() ? (int)(a->e.mem->eng->P.L - a->e.l) : -1; }
int fl_elem_limbs(const fl_elem* a) { return a->e.l; }
int fl_elem_deg(const fl_elem* a) { return a->e.deg; }
int fl_elem_slots(const fl_elem* a) { return a->e.slots; }
int fl_elem_ncomp(const fl_elem* a) { return a->e.ncomp; }
double fl_elem_scale(const fl_elem* a) { return a->e.scale; }
int fl_elem_batch(const fl_elem* a) { return a->e.batch; }
int fl_batch_pack(fl_ctx* c, fl_elem* const* v, int n, fl_elem** out) {
    FL_TRY({ std::vector<Elem> e(n); for (int i = 0; i < n; ++i) e[i] = v[i]->e; *out = wrap(c->sch->pack(e)); })
}
int fl_batch_slice(fl_ctx* c, const fl_elem* a, int i, fl_elem** out) { FL_TRY(*out = wrap(c->sch->slice(a->e, i))) }
int fl_batch_range(fl_ctx* c, const fl_elem* a, int first, int count, fl_elem** out) {
    FL_TRY({
        *out = wrap(c->sch->range(a->e, first, count));
    })
}
int fl_elem_clone(fl_ctx* c, const fl_elem* a, fl_elem** out) { FL_TRY(*out = wrap(c->sch->clone(a->e))) }
void fl_elem_free(fl_elem* a) { delete a; }
int fl_elem_export(fl_ctx* c, const fl_elem* a, uint64_t* host) {
    FL_TRY(c->eng->download(host, a->e.data(), (size_t)a->e.batch * a->e.ncomp * a->e.l * c->eng->P.N))
}
int fl_elem_import(fl_ctx* c, const uint64_t* host, int ncomp, int limbs, int deg, double scale, int slots, fl_elem** out) {
    FL_TRY(*out = wrap(c->sch->import_elem(host, ncomp, limbs, deg, scale, slots)))
}
int fl_elem_save(fl_ctx* c, const fl_elem* a, const char* path) { FL_TRY(c->sch->save_elem(a->e, path)) }
int fl_elem_load(fl_ctx* c, const char* path, fl_elem** out) { FL_TRY(*out = wrap(c->sch->load_elem(path))) }

int fl_prof_enable(fl_ctx* c, int on) { c->prof_on = on != 0; return 0; }
int fl_prof_dump(fl_ctx* c, char* buf, size_t cap) {
    c->eng->sync();
    std::map<std::string, std::array<double, 3>> acc;   // calls, gpu ms, host s
    for (auto& r : c->prof) {
        float ms = 0;
        cudaEventElapsedTime(&ms, r.a, r.b);
        auto& e = acc[r.name];
        e[0] += 1; e[1] += ms; e[2] += r.host_s;
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    c->prof.clear();
    std::string s;
    for (auto& kv : acc) {
        char line[200];
        std::snprintf(line, sizeof line, "%s %.0f %.3f %.3f\n", kv.first.c_str(), kv.second[0], kv.second[1], kv.second[2] * 1e3);
        s += line;
    }
    if (s.size() + 1 > cap) { g_err = "profile buffer too small"; return 1; }
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return 0;
}
int fl_ledger_enable(fl_ctx* c, int on) { c->eng->ledger_on = on != 0; return 0; }
int fl_ledger_reset(fl_ctx* c) { c->eng->ledger.reset(); return 0; }
int fl_ledger_dump(fl_ctx* c, char* buf, size_t cap) {
    FL_TRY({
        std::string s;
        for (auto& kv : c->eng->ledger.rows) {
            char line[160];
            std::snprintf(line, sizeof line, "%s %ld %.0f\n", kv.first.c_str(), kv.second.count, kv.second.bytes);
            s += line;
        }
        if (s.size() + 1 > cap) throw std::runtime_error("ledger buffer too small");
        std::memcpy(buf, s.c_str(), s.size() + 1);
    })
}

}  // extern "C"
