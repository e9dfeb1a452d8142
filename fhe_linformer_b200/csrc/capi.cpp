// capi.cpp -- extern "C" boundary (include/fl_ckks.h) over the device engine.
#include "../../include/fl_ckks.h"

#include <cstring>
#include <string>

#include "engine.h"

using namespace flk;

struct fl_ctx {
    Engine* eng;
};

static thread_local std::string g_err;

const char* fl_last_error(void) { return g_err.c_str(); }

#define FL_TRY(...)                      \
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

static LimbSel make_sel(const int* midx, int nl) {
    if (nl < 0 || nl > kMaxLimbSel) throw std::invalid_argument("limb count out of range");
    LimbSel s; s.n = nl;
    for (int i = 0; i < nl; ++i) { s.m[i] = (uint8_t)midx[i]; s.pos[i] = (uint8_t)i; }
    return s;
}

extern "C" {

int fl_ctx_create(const fl_params* p, int device, fl_ctx** out) {
    FL_TRY({
        ParamSpec s;
        s.logN = p->logN; s.L = p->L; s.dnum = p->dnum; s.first_bits = p->first_bits; s.scale_bits = p->scale_bits;
        s.aux_bits = p->aux_bits; s.sparse_h = p->sparse_h;
        fl_ctx* c = new fl_ctx{nullptr};
        try { c->eng = new Engine(s, device); } catch (...) { delete c; throw; }
        *out = c;
    })
}
void fl_ctx_destroy(fl_ctx* c) {
    if (!c) return;
    delete c->eng;
    delete c;
}
int fl_ctx_info(fl_ctx* c, int* info) {
    FL_TRY({ const Params& P = c->eng->P; info[0] = P.logN; info[1] = P.L; info[2] = P.K; info[3] = P.alpha; info[4] = P.dnum; })
}
int fl_ctx_moduli(fl_ctx* c, uint64_t* out) { FL_TRY(std::memcpy(out, c->eng->P.q.data(), 8 * c->eng->P.T)) }
int fl_ctx_roots(fl_ctx* c, uint64_t* out) { FL_TRY(std::memcpy(out, c->eng->P.psi.data(), 8 * c->eng->P.T)) }
int fl_ctx_scale_factors(fl_ctx* c, double* out) { FL_TRY(std::memcpy(out, c->eng->P.sf.data(), 8 * c->eng->P.L)) }
uint32_t fl_galois_for_rotation(fl_ctx* c, int k) { return c->eng->P.galois_for_rotation(k); }
uint32_t fl_galois_conj(fl_ctx* c) { return c->eng->P.galois_conj(); }
void* fl_ctx_stream(fl_ctx* c) { return (void*)c->eng->stream; }
int fl_sync(fl_ctx* c) { FL_TRY(c->eng->sync()) }

int fl_dev_alloc(fl_ctx* c, size_t words, uint64_t** out) { FL_TRY(*out = c->eng->alloc(words)) }
int fl_dev_free(fl_ctx* c, uint64_t* p) { FL_TRY(c->eng->release(p)) }
int fl_dev_upload(fl_ctx* c, uint64_t* dst, const uint64_t* src, size_t words) { FL_TRY(c->eng->upload(dst, src, words)) }
int fl_dev_download(fl_ctx* c, uint64_t* dst, const uint64_t* src, size_t words) { FL_TRY(c->eng->download(dst, src, words)) }

int fl_raw_ntt(fl_ctx* c, uint64_t* d, const int* midx, int nl) { FL_TRY(c->eng->ntt(d, make_sel(midx, nl))) }
int fl_raw_intt(fl_ctx* c, uint64_t* d, const int* midx, int nl) { FL_TRY(c->eng->intt(d, make_sel(midx, nl))) }
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
int fl_raw_mul_relin(fl_ctx* c, uint64_t* out, const uint64_t* a, const uint64_t* b, int l, const uint64_t* evk) {
    FL_TRY(c->eng->mul_relin(out, a, b, l, evk))
}
int fl_raw_mul_plain(fl_ctx* c, uint64_t* out, const uint64_t* ct, const uint64_t* pt, int l) {
    FL_TRY(c->eng->ew(EwOp::Mul, out, ct, pt, l, 2, true))
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

}  // extern "C"
