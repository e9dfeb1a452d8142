// lintrans.cpp -- baby-step/giant-step diagonal linear transforms on ciphertexts: the ct x pt matrix product behind
// CoeffsToSlots / SlotsToCoeffs (OpenFHE EvalLinearTransform inside EvalBootstrap, reference FHEController.cpp:445) and the
// packed linear layers (E/F projections, Q/K/V/FFN weights as diagonals; BASELINE.json north star).
//
//   (M v)[p] = sum_d diag_d[p] v[(p + d) mod n],  d = g (n1 j + i - off):
//   M v = sum_j Rot_{G_j}( sum_i P_{j,i} * Rot_{g i}(v) ),  G_j = g (n1 j - off),  P_{j,i} = Rot_{-G_j}(diag_d).
//
// lintrans_apply runs the whole transform as one engine call with double hoisting (Engine::linear_transform): the baby
// rotations share one ModUp and are never ModDown'ed, the plaintexts live in the extended basis Q_l u P, every giant step
// costs one ModDown, and the giant rotations share the final ModDown.  lintrans_apply_plain is the textbook sequence of
// EvalRotate / EvalMult / EvalAdd over the same plan, kept as the checker (tests/test_gpu_lintrans.py).
#include <algorithm>
#include <cmath>
#include <numeric>
#include <stdexcept>

#include "scheme.h"

namespace flk {

namespace {
std::vector<cplx> rotated(const std::vector<cplx>& v, int k) {
    const int n = (int)v.size();
    std::vector<cplx> r(n);
    for (int p = 0; p < n; ++p) r[p] = v[(((p + k) % n) + n) % n];
    return r;
}
}  // namespace

LinTrans Scheme::lintrans_plan(const std::map<int, std::vector<cplx>>& diags, int slots, int max_baby) {
    const int n = slots;
    if (n < 1 || n > P.N / 2 || (n & (n - 1))) throw std::invalid_argument("linear transform: slots must be a power of two <= N/2");
    if (diags.empty()) throw std::invalid_argument("linear transform: no diagonals");
    LinTrans t;
    t.slots = n;
    // signed shifts in (-n/2, n/2]; common stride g
    std::map<int, const std::vector<cplx>*> by_shift;
    int g = 0;
    for (auto& kv : diags) {
        if ((int)kv.second.size() != n) throw std::invalid_argument("linear transform: diagonal length must equal slots");
        int d = ((kv.first % n) + n) % n;
        if (d > n / 2) d -= n;
        by_shift[d] = &kv.second;
        g = std::gcd(g, std::abs(d));
    }
    if (g == 0) g = 1;
    t.g = g;
    const int lo = by_shift.begin()->first / g, hi = by_shift.rbegin()->first / g;
    t.off = -std::min(lo, 0);
    const int top = std::max(hi, 0);
    t.cnt = top + t.off + 1;
    t.n1 = 1;
    while (t.n1 * t.n1 < t.cnt) t.n1 <<= 1;
    if (max_baby > 0) t.n1 = std::min(t.n1, max_baby);
    t.n1 = std::min(t.n1, kBsgsMax);
    t.n2 = (t.cnt + t.n1 - 1) / t.n1;
    if (t.n2 > kBsgsMax) throw std::invalid_argument("linear transform: more than 256 diagonal positions between the extreme shifts");
    // giant steps: rotating ones first, the non-rotating one (if present) last
    std::vector<int> order;
    for (int j = 0; j < t.n2; ++j) if (t.n1 * j != t.off) order.push_back(j);
    for (int j = 0; j < t.n2; ++j) if (t.n1 * j == t.off) order.push_back(j);
    std::vector<int> kept;
    for (int j : order) {
        const int G = g * (t.n1 * j - t.off);
        std::vector<std::vector<cplx>> row(t.n1);
        uint32_t m = 0;
        for (int i = 0; i < t.n1; ++i) {
            auto it = by_shift.find(g * (t.n1 * j + i - t.off));
            if (it == by_shift.end()) continue;
            row[i] = rotated(*it->second, -G);
            m |= 1u << i;
            ++t.ndiag;
        }
        if (!m) continue;                 // giant step without diagonals
        t.giant_rot.push_back(G);
        t.host.push_back(std::move(row));
        t.mask.push_back(m);
    }
    t.n2 = (int)t.giant_rot.size();
    return t;
}

std::vector<int> Scheme::lintrans_rotations(const LinTrans& t) const {
    std::vector<int> r;
    uint32_t used = 0;
    for (uint32_t m : t.mask) used |= m;
    for (int i = 1; i < t.n1; ++i) if ((used >> i) & 1u) r.push_back(t.g * i);
    for (int G : t.giant_rot) if (G) r.push_back(G);
    std::sort(r.begin(), r.end());
    r.erase(std::unique(r.begin(), r.end()), r.end());
    return r;
}

void Scheme::lintrans_encode(LinTrans& t, int level) {
    if (level < 0 || level >= P.L) throw std::invalid_argument("linear transform: level out of range");
    const int l = P.L - level, ext = l + P.K, n = t.slots;
    const double scale = P.sf[level];
    t.pts = std::make_shared<DevMem>(&eng, (size_t)t.n2 * t.n1 * ext * P.N);
    DevFft& f = dev_fft(n);
    LimbSel se;
    for (int k = 0; k < ext; ++k) se.push(P.mod_index_ext(l, k), k);
    for (int j = 0; j < t.n2; ++j)
        for (int i = 0; i < t.n1; ++i) {
            if (!((t.mask[j] >> i) & 1u)) continue;
            const std::vector<cplx>& v = t.host[j][i];
            int slot;
            double* h = stage_slot(slot);
            for (int p = 0; p < n; ++p) { h[p] = v[p].real(); h[n + p] = v[p].imag(); }
            double* d = (double*)eng.alloc((size_t)2 * n);
            FLK_CUDA(cudaMemcpyAsync(d, h, (size_t)2 * n * sizeof(double), cudaMemcpyHostToDevice, eng.stream));
            FLK_CUDA(cudaEventRecord(stage_ev_[slot], eng.stream));
            u64* dst = t.pts->p + ((size_t)j * t.n1 + i) * ext * P.N;
            launch_encode(eng.T, dst, d, d + n, n, scale, l, f.rot, f.cre, f.cim, eng.stream, P.K);
            eng.ntt(dst, se);
            eng.release((u64*)d);
        }
    t.level = level;
    t.pt_scale = scale;
}

Elem Scheme::lintrans_apply(LinTrans& t, const Elem& in) {
    if (in.ncomp != 2) throw std::invalid_argument("linear transform: ciphertext expected");
    if (in.slots != t.slots) throw std::invalid_argument("linear transform: slot count differs from the plan");
    Elem ct = in;
    if (ct.deg == 2) rescale_inplace(ct);
    const int lvl = level_of(ct);
    if (lvl != t.level) lintrans_encode(t, lvl);     // plaintexts follow the level the ciphertext actually has
    LtPlan p;
    p.n1 = t.n1; p.n2 = t.n2; p.l = ct.l; p.pts = t.pts->p; p.ndiag = t.ndiag;
    uint32_t used = 0;
    for (int j = 0; j < t.n2; ++j) { p.mask[j] = t.mask[j]; used |= t.mask[j]; }
    auto key_of = [&](int k, uint32_t& g) -> const u64* {
        g = P.galois_for_rotation(k);
        auto it = gk_.find(g);
        if (it == gk_.end()) throw std::runtime_error("linear transform: no evaluation key for rotation " + std::to_string(k));
        return it->second;
    };
    p.baby_g[0] = 1;
    for (int i = 1; i < t.n1; ++i)
        if ((used >> i) & 1u) p.baby_evk[i] = key_of(t.g * i, p.baby_g[i]);
    for (int j = 0; j < t.n2; ++j) {
        if (t.giant_rot[j]) p.giant_evk[j] = key_of(t.giant_rot[j], p.giant_g[j]);
        else p.giant_g[j] = 1;
    }
    Elem r = make(2, ct.l, ct.deg + 1, ct.scale * t.pt_scale, ct.slots, ct.batch);
    const int mb = std::max(1, max_batch(ct.l) / std::max(1, std::max(t.n1 - 1, t.n2)));
    for (int b0 = 0; b0 < ct.batch; b0 += mb) {
        const size_t o = (size_t)b0 * ct.words_each(P.N);
        eng.linear_transform(r.data() + o, ct.data() + o, std::min(mb, ct.batch - b0), p);
    }
    return r;
}

Elem Scheme::lintrans_apply_plain(LinTrans& t, const Elem& in) {
    Elem ct = in;
    if (ct.deg == 2) rescale_inplace(ct);
    const int lvl = level_of(ct), n = t.slots;
    std::vector<Elem> baby(t.n1);
    baby[0] = ct;
    uint32_t used = 0;
    for (uint32_t m : t.mask) used |= m;
    for (int i = 1; i < t.n1; ++i)
        if ((used >> i) & 1u) baby[i] = rotate(ct, t.g * i);
    Elem acc;
    for (int j = 0; j < t.n2; ++j) {
        Elem inner;
        for (int i = 0; i < t.n1; ++i) {
            if (!((t.mask[j] >> i) & 1u)) continue;
            Elem pt = encode(t.host[j][i].data(), n, lvl, n, 1);
            Elem term = mult(baby[i], pt);
            inner = inner.valid() ? add(inner, term) : term;
        }
        if (t.giant_rot[j]) inner = rotate(inner, t.giant_rot[j]);
        acc = acc.valid() ? add(acc, inner) : inner;
    }
    return acc;
}

}  // namespace flk
