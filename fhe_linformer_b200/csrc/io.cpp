// io.cpp -- ciphertext / key files.  Stands in for the cereal-BINARY Serial::SerializeToFile calls of the reference
// (FHEController.cpp:59-89 keys, :251 automorphism keys, :1360-1394 ciphertext checkpoint).  The container is our own
// ("FLCK" header + raw little-endian limbs); reading OpenFHE's cereal archives is SURVEY.md row F3 (needs artifacts).
#include <cstdio>
#include <cstring>

#include "scheme.h"

namespace flk {
namespace {
struct FileHdr {
    char magic[4];
    uint32_t kind;     // 1 = element, 2 = key bundle
    int32_t logN, L, K, ncomp, l, deg, slots, nkeys;
    double scale;
};
struct File {
    FILE* f;
    File(const char* path, const char* mode) : f(std::fopen(path, mode)) {
        if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    }
    ~File() { if (f) std::fclose(f); }
    void write(const void* p, size_t n) { if (std::fwrite(p, 1, n, f) != n) throw std::runtime_error("short write"); }
    void read(void* p, size_t n) { if (std::fread(p, 1, n, f) != n) throw std::runtime_error("short read"); }
};
}  // namespace

Elem Scheme::import_elem(const u64* host, int ncomp, int l, int deg, double scale, int slots) {
    if (ncomp < 1 || ncomp > 2 || l < 1 || l > P.L) throw std::invalid_argument("import: bad shape");
    Elem e = make(ncomp, l, deg, scale, slots);
    eng.upload(e.data(), host, (size_t)ncomp * l * P.N);
    eng.sync();
    return e;
}

void Scheme::save_elem(const Elem& a, const char* path) {
    if (!a.valid()) throw std::invalid_argument("save: empty ciphertext");
    std::vector<u64> h((size_t)a.ncomp * a.l * P.N);
    eng.download(h.data(), a.data(), h.size());
    FileHdr hd{};
    std::memcpy(hd.magic, "FLCK", 4);
    hd.kind = 1; hd.logN = P.logN; hd.L = P.L; hd.K = P.K; hd.ncomp = a.ncomp; hd.l = a.l; hd.deg = a.deg; hd.slots = a.slots; hd.scale = a.scale;
    File f(path, "wb");
    f.write(&hd, sizeof hd);
    f.write(h.data(), h.size() * 8);
}

Elem Scheme::load_elem(const char* path) {
    File f(path, "rb");
    FileHdr hd{};
    f.read(&hd, sizeof hd);
    if (std::memcmp(hd.magic, "FLCK", 4) || hd.kind != 1 || hd.logN != P.logN || hd.L != P.L) throw std::runtime_error("not a ciphertext file of this context");
    std::vector<u64> h((size_t)hd.ncomp * hd.l * P.N);
    f.read(h.data(), h.size() * 8);
    return import_elem(h.data(), hd.ncomp, hd.l, hd.deg, hd.scale, hd.slots);
}

void Scheme::save_keys(const char* path) {
    if (!sk_ || !pk_) throw std::runtime_error("save_keys: no key pair");
    FileHdr hd{};
    std::memcpy(hd.magic, "FLCK", 4);
    hd.kind = 2; hd.logN = P.logN; hd.L = P.L; hd.K = P.K; hd.nkeys = (int)gk_.size() + (mk_ ? 1 : 0);
    File f(path, "wb");
    f.write(&hd, sizeof hd);
    std::vector<u64> buf(std::max((size_t)2 * P.L * P.N, eng.evk_words()));
    export_sk(buf.data()); f.write(buf.data(), (size_t)P.T * P.N * 8);
    export_pk(buf.data()); f.write(buf.data(), (size_t)2 * P.L * P.N * 8);
    auto put = [&](uint32_t g) { export_evk(g, buf.data()); f.write(&g, 4); f.write(buf.data(), eng.evk_words() * 8); };
    if (mk_) put(0);
    for (auto& kv : gk_) put(kv.first);
}

void Scheme::load_keys(const char* path) {
    File f(path, "rb");
    FileHdr hd{};
    f.read(&hd, sizeof hd);
    if (std::memcmp(hd.magic, "FLCK", 4) || hd.kind != 2 || hd.logN != P.logN || hd.L != P.L || hd.K != P.K) throw std::runtime_error("not a key file of this context");
    std::vector<u64> buf(std::max((size_t)2 * P.L * P.N, eng.evk_words())), sk((size_t)P.T * P.N);
    f.read(sk.data(), sk.size() * 8);
    f.read(buf.data(), (size_t)2 * P.L * P.N * 8);
    import_keys(sk.data(), buf.data());
    for (int i = 0; i < hd.nkeys; ++i) {
        uint32_t g;
        f.read(&g, 4);
        f.read(buf.data(), eng.evk_words() * 8);
        import_evk(g, buf.data());
    }
}

}  // namespace flk
