// io.cpp -- ciphertext / key files.  Stands in for the cereal-BINARY Serial::SerializeToFile calls of the reference
// (FHEController.cpp:59-89 keys, :251 automorphism keys, :1360-1394 ciphertext checkpoint).  The container is our own
// ("FLCK" header + raw little-endian limbs); reading OpenFHE's cereal archives is SURVEY.md row F3 (needs artifacts).
#include <cstdio>
#include <cstring>

#include "scheme.h"

namespace flk {
namespace {
constexpr uint32_t kFormatVersion = 2;       // 2: version / byte-order mark / chain fingerprint / batch count added to the header
constexpr uint32_t kByteOrderMark = 0x01020304u;
struct FileHdr {
    char magic[4];
    uint32_t kind;     // 1 = element, 2 = key bundle
    int32_t logN, L, K, ncomp, l, deg, slots, nkeys;
    double scale;
    uint32_t version, bom;
    uint64_t chain;    // fingerprint of the modulus chain (moduli and scaling rule): same logN / L with other primes must not load
    int32_t batch, pad;
};
// FNV-1a over the moduli: a file written under another prime chain decrypts to garbage, so it is refused instead
uint64_t chain_fingerprint(const Params& P) {
    uint64_t h = 0xcbf29ce484222325ull;
    auto mix = [&](uint64_t v) { for (int i = 0; i < 8; ++i) { h ^= (v >> (8 * i)) & 0xff; h *= 0x100000001b3ull; } };
    for (u64 q : P.q) mix(q);
    mix((uint64_t)P.spec.first_bits); mix((uint64_t)P.spec.scale_bits); mix((uint64_t)P.dnum);
    return h;
}
void fill_header(FileHdr& hd, uint32_t kind, const Params& P) {
    std::memcpy(hd.magic, "FLCK", 4);
    hd.kind = kind; hd.logN = P.logN; hd.L = P.L; hd.K = P.K;
    hd.version = kFormatVersion; hd.bom = kByteOrderMark; hd.chain = chain_fingerprint(P);
}
void check_header(const FileHdr& hd, uint32_t kind, const Params& P, const char* what) {
    if (std::memcmp(hd.magic, "FLCK", 4) || hd.kind != kind) throw std::runtime_error(std::string("not a ") + what + " file");
    if (hd.bom != kByteOrderMark) throw std::runtime_error(std::string(what) + " file written with another byte order or by format version 1");
    if (hd.version != kFormatVersion) throw std::runtime_error(std::string(what) + " file has format version " + std::to_string(hd.version) + ", expected " + std::to_string(kFormatVersion));
    if (hd.logN != P.logN || hd.L != P.L || hd.K != P.K || hd.chain != chain_fingerprint(P))
        throw std::runtime_error(std::string(what) + " file belongs to another context (ring, chain length or moduli differ)");
}
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
    if (a.batch != 1) throw std::invalid_argument("save: a batched operand holds " + std::to_string(a.batch) + " ciphertexts; save its slices one by one");
    std::vector<u64> h((size_t)a.ncomp * a.l * P.N);
    eng.download(h.data(), a.data(), h.size());
    FileHdr hd{};
    fill_header(hd, 1, P);
    hd.ncomp = a.ncomp; hd.l = a.l; hd.deg = a.deg; hd.slots = a.slots; hd.scale = a.scale; hd.batch = 1;
    File f(path, "wb");
    f.write(&hd, sizeof hd);
    f.write(h.data(), h.size() * 8);
}

Elem Scheme::load_elem(const char* path) {
    File f(path, "rb");
    FileHdr hd{};
    f.read(&hd, sizeof hd);
    check_header(hd, 1, P, "ciphertext");
    // validate every field before it sizes an allocation
    if (hd.ncomp < 1 || hd.ncomp > 2 || hd.l < 1 || hd.l > P.L || hd.deg < 1 || hd.deg > 4 || hd.batch != 1 || hd.slots < 1 || hd.slots > P.N / 2 ||
        (hd.slots & (hd.slots - 1)) || !(hd.scale > 0))
        throw std::runtime_error("corrupt ciphertext header");
    std::vector<u64> h((size_t)hd.ncomp * hd.l * P.N);
    f.read(h.data(), h.size() * 8);
    return import_elem(h.data(), hd.ncomp, hd.l, hd.deg, hd.scale, hd.slots);
}

// Key bundle = header + tagged records {tag, galois, words} + payload.  tag: 1 secret key, 2 public key, 3 evaluation key
// (galois 0 = relinearisation key).  `what` selects what is written: 1 sk | 2 pk | 4 mult key | 8 automorphism keys,
// mirroring the reference's separate secret-key / public-key / mult-keys / rot_* files (FHEController.cpp:59-89,251).
void Scheme::save_keys(const char* path, int what) {
    FileHdr hd{};
    fill_header(hd, 2, P);
    std::vector<uint32_t> evks;
    if ((what & 4) && mk_) evks.push_back(0);
    if (what & 8) for (auto& kv : gk_) evks.push_back(kv.first);
    if ((what & 3) && (!sk_ || !pk_)) throw std::runtime_error("save_keys: no key pair");
    hd.nkeys = (int)evks.size() + ((what & 1) ? 1 : 0) + ((what & 2) ? 1 : 0);
    File f(path, "wb");
    f.write(&hd, sizeof hd);
    std::vector<u64> buf(std::max((size_t)2 * P.L * P.N, eng.evk_words()));
    auto rec = [&](uint32_t tag, uint32_t g, size_t words) { uint64_t w = words; f.write(&tag, 4); f.write(&g, 4); f.write(&w, 8); f.write(buf.data(), words * 8); };
    if (what & 1) { export_sk(buf.data()); rec(1, 0, (size_t)P.T * P.N); }
    if (what & 2) { export_pk(buf.data()); rec(2, 0, (size_t)2 * P.L * P.N); }
    for (uint32_t g : evks) { export_evk(g, buf.data()); rec(3, g, eng.evk_words()); }
}

void Scheme::load_keys(const char* path) {
    File f(path, "rb");
    FileHdr hd{};
    f.read(&hd, sizeof hd);
    check_header(hd, 2, P, "key");
    if (hd.nkeys < 0 || hd.nkeys > (1 << 20)) throw std::runtime_error("corrupt key file");
    std::vector<u64> buf;
    for (int i = 0; i < hd.nkeys; ++i) {
        uint32_t tag, g; uint64_t words;
        f.read(&tag, 4); f.read(&g, 4); f.read(&words, 8);
        if (words > std::max((size_t)2 * P.L * P.N, eng.evk_words())) throw std::runtime_error("corrupt key file");
        buf.resize(words);
        f.read(buf.data(), words * 8);
        if (tag == 1 && words == (size_t)P.T * P.N) import_keys(buf.data(), nullptr);
        else if (tag == 2 && words == (size_t)2 * P.L * P.N) import_keys(nullptr, buf.data());
        else if (tag == 3 && words == eng.evk_words()) import_evk(g, buf.data());
        else throw std::runtime_error("corrupt key record");
    }
}

}  // namespace flk
