// poly.cpp -- polynomial evaluation on ciphertexts: power-basis (EvalPoly, reference FHEController.cpp:1291)
// and Chebyshev series (EvalChebyshevFunction, FHEController.cpp:486,1319-1335) with a depth-optimal
// baby-step / giant-step (Paterson-Stockmeyer) schedule: depth ceil(log2(deg+1)), the same budget OpenFHE's
// EvalChebyshevSeriesPS needs (reference table Utils.h:127-153).
#include <algorithm>
#include <cmath>
#include <functional>
#include <map>

#include "scheme.h"

namespace flk {

// OpenFHE EvalChebyshevCoefficients (A.10): f sampled at the deg+1 Chebyshev nodes of [a,b]
std::vector<double> Scheme::chebyshev_coefficients(double (*f)(double, void*), void* user, double a, double b, int degree) {
    if (degree < 1) throw std::invalid_argument("Chebyshev degree must be positive");
    const int n = degree + 1;
    const double bma = 0.5 * (b - a), bpa = 0.5 * (b + a), pin = M_PI / n;
    std::vector<double> fx(n), c(n, 0.0);
    for (int i = 0; i < n; ++i) fx[i] = f(std::cos(pin * (i + 0.5)) * bma + bpa, user);
    const double mf = 2.0 / n;
    for (int i = 0; i < n; ++i) {
        double s = 0;
        for (int j = 0; j < n; ++j) s += fx[j] * std::cos(pin * i * (j + 0.5));
        c[i] = s * mf;
    }
    return c;
}

namespace {
bool negligible(double c) { return std::fabs(c) < 1e-300; }
bool same_shape(const Elem& x, const Elem& y) { return x.l == y.l && x.deg == y.deg && x.scale == y.scale && x.batch == y.batch && x.slots == y.slots; }
int poly_degree(const std::vector<double>& c) {
    int d = (int)c.size() - 1;
    while (d > 0 && negligible(c[d])) --d;
    return d;
}
}  // namespace

// elements [g * each, (g + 1) * each) of a batched operand, as a view
Elem Scheme::part_of(const Elem& all, size_t g, int each) const {
    Elem e = all;
    e.off = all.off + g * (size_t)each * all.words_each(P.N);
    e.batch = each;
    return e;
}

// Sums of independent pairs, one batched EvalAdd per group of pairs that share their shapes (the level adjustment of the
// shallower side, a scalar product and a rescale, then happens once per group instead of once per pair).
std::vector<Elem> Scheme::add_each(const std::vector<Elem>& a, const std::vector<Elem>& b) {
    const size_t n = a.size();
    if (b.size() != n) throw std::invalid_argument("add_each: operand counts differ");
    std::vector<Elem> out(n);
    std::vector<char> done(n, 0);
    for (size_t first = 0; first < n; ++first) {
        if (done[first]) continue;
        std::vector<size_t> grp;
        for (size_t i = first; i < n; ++i)
            if (!done[i] && same_shape(a[i], a[first]) && same_shape(b[i], b[first])) { grp.push_back(i); done[i] = 1; }
        if (grp.size() == 1) { out[first] = add(a[first], b[first]); continue; }
        std::vector<Elem> ga, gb;
        for (size_t i : grp) { ga.push_back(a[i]); gb.push_back(b[i]); }
        const Elem sum = add(pack(ga), pack(gb));
        for (size_t g = 0; g < grp.size(); ++g) out[grp[g]] = part_of(sum, g, a[first].batch);
    }
    return out;
}

// Baby polynomials sum_{1 <= i <= upto} c[i] T[i] (T[0] is the constant 1: c[0] is left to the caller), ALL of an evaluation at
// once.  The terms a combination uses are brought to the deepest level among them (never deeper, so no level is wasted) and to
// one scale -- OpenFHE's EvalChebyshevSeriesPS aligns levels the same way before its EvalLinearWSum -- and every group of
// polynomials that ends up on the same level is ONE kernel (weighted_sums: the 2^m baby polynomials of a Paterson-Stockmeyer
// tree share their terms) instead of a scalar multiplication, a rescale and an addition per term.  `settled` holds T[i] after
// its pending rescale, `aligned` caches (i, target limbs) -> adjusted copy, so each T[i] is adjusted once per target level.
// An entry of the result is invalid when its polynomial has no term beyond c[0].
std::vector<Elem> Scheme::inner_linear_many(const std::vector<Elem>& T, const std::vector<const std::vector<double>*>& polys, int kmax) {
    std::vector<Elem> out(polys.size());
    if (ps_settled_.size() != T.size()) { ps_settled_.assign(T.size(), Elem()); ps_aligned_.clear(); }
    auto settled = [&](int i) -> const Elem& {
        if (!ps_settled_[i].valid()) {
            ps_settled_[i] = T[i];
            if (ps_settled_[i].deg == 2) rescale_inplace(ps_settled_[i]);
        }
        return ps_settled_[i];
    };
    // the level each polynomial lands on: that of the deepest term it uses
    std::map<int, std::vector<size_t>> by_limbs;
    for (size_t p = 0; p < polys.size(); ++p) {
        const std::vector<double>& c = *polys[p];
        int limbs = 0;
        for (int i = 1; i <= kmax && i < (int)c.size(); ++i)
            if (!negligible(c[i])) limbs = limbs ? std::min(limbs, settled(i).l) : settled(i).l;
        if (limbs) by_limbs[limbs].push_back(p);
    }
    for (const auto& grp : by_limbs) {
        const int limbs = grp.first;
        std::vector<int> idx;                              // union of the terms the group uses
        for (int i = 1; i <= kmax; ++i)
            for (size_t p : grp.second)
                if (i < (int)polys[p]->size() && !negligible((*polys[p])[i])) { idx.push_back(i); break; }
        int deepest = -1;
        for (int i : idx)
            if (settled(i).l == limbs) { deepest = i; break; }
        const Elem ref = settled(deepest);
        // terms above the target level are adjusted to it -- all those that share a level in one go (T_(2^(d-1)+1) .. T_(2^d) do)
        std::vector<int> todo;
        for (int i : idx) {
            const Elem& si = settled(i);
            if (!(si.l == ref.l && si.scale == ref.scale) && !ps_aligned_.count(std::make_pair(i, ref.l))) todo.push_back(i);
        }
        while (!todo.empty()) {
            const Elem first = settled(todo[0]);
            std::vector<int> same, rest;
            for (int i : todo) (same_shape(settled(i), first) ? same : rest).push_back(i);
            std::vector<Elem> src;
            for (int i : same) src.push_back(settled(i));
            Elem a = src.size() == 1 ? src[0] : pack(src), r = ref;
            adjust_pair(a, r);
            for (size_t g = 0; g < same.size(); ++g) ps_aligned_.emplace(std::make_pair(same[g], ref.l), part_of(a, g, first.batch));
            todo.swap(rest);
        }
        std::vector<Elem> terms;
        for (int i : idx) {
            const Elem& si = settled(i);
            terms.push_back(si.l == ref.l && si.scale == ref.scale ? si : ps_aligned_.at(std::make_pair(i, ref.l)));
        }
        std::vector<std::vector<double>> w(grp.second.size(), std::vector<double>(idx.size(), 0.0));
        for (size_t g = 0; g < grp.second.size(); ++g) {
            const std::vector<double>& c = *polys[grp.second[g]];
            for (size_t t = 0; t < idx.size(); ++t)
                if (idx[t] < (int)c.size() && !negligible(c[idx[t]])) w[g][t] = c[idx[t]];
        }
        const Elem all = weighted_sums(terms, w);
        for (size_t g = 0; g < grp.second.size(); ++g) out[grp.second[g]] = part_of(all, g, ref.batch);
    }
    return out;
}

// Products of independent pairs as ONE batched EvalMult per group of operands that share level, degree and scale (one tensor
// product, one relinearisation for the whole group).  At the sizes of the forward (N = 2^15, a dozen limbs, one ciphertext) a
// key switch is a string of launches that each fill a fraction of the chip, so n products in one call cost little more than one.
// b may hold a single element: it multiplies all.  With `align` every side is first brought to the deepest level among its
// elements (what EvalMult would do pairwise when the partner is that deep anyway); without it nothing is adjusted here, so no
// product lands deeper than its own EvalMult would put it.
std::vector<Elem> Scheme::mult_each(std::vector<Elem> a, std::vector<Elem> b, bool align) {
    const size_t n = a.size();
    if (b.size() != n && b.size() != 1) throw std::invalid_argument("mult_each: operand counts differ");
    std::vector<Elem> out(n);
    if (n == 0) return out;
    auto deepen = [&](std::vector<Elem>& v) {
        size_t ref = 0;
        for (size_t i = 1; i < v.size(); ++i)
            if (v[i].l < v[ref].l || (v[i].l == v[ref].l && v[i].deg > v[ref].deg)) ref = i;
        const Elem target = v[ref];
        for (Elem& e : v)
            if (e.l != target.l || e.deg != target.deg) { Elem r = target; adjust_pair(e, r); }
    };
    if (align) { deepen(a); deepen(b); }
    auto same = [](const Elem& x, const Elem& y) { return x.l == y.l && x.deg == y.deg && x.scale == y.scale && x.batch == y.batch && x.slots == y.slots; };
    std::vector<char> done(n, 0);
    for (size_t first = 0; first < n; ++first) {
        if (done[first]) continue;
        std::vector<size_t> grp;
        for (size_t i = first; i < n; ++i)
            if (!done[i] && same(a[i], a[first]) && (b.size() == 1 || same(b[i], b[first]))) { grp.push_back(i); done[i] = 1; }
        if (grp.size() == 1) { out[first] = mult(a[first], b[b.size() == 1 ? 0 : first]); continue; }
        std::vector<Elem> ga, gb;
        const int each = a[first].batch;
        for (size_t i : grp) {
            ga.push_back(a[i]);
            if (b.size() > 1) gb.push_back(b[i]);
            else if (each > 1) gb.push_back(b[0]);       // a batched multiplier is repeated; a single one is broadcast by EvalMult
        }
        const Elem prod = mult(pack(ga), gb.empty() ? b[0] : pack(gb));
        for (size_t g = 0; g < grp.size(); ++g) out[grp[g]] = part_of(prod, g, each);
    }
    return out;
}

// Chebyshev series sum c[i] T_i(x) (true coefficients, c[0] not halved), x already mapped to [-1,1].
// Level-synchronous Paterson-Stockmeyer: the baby steps of one depth (T_i, 2^(d-1) < i <= 2^d) are independent and go
// through mult_each together, and so do the products q T_g of all nodes at one height of the recursive split; a degree-300
// series is 12 batched EvalMult calls instead of 49 single ones, on the same tree and with the same level budget.
Elem Scheme::cheby_ps(const Elem& x, const std::vector<double>& c_in) {
    std::vector<double> c = c_in;
    const int n = poly_degree(c);
    c.resize(n + 1);
    if (n == 0) { Elem z = mult_const(x, 0.0); return add_const(z, c[0]); }
    int D = 0;
    while ((1 << D) < n + 1) ++D;
    const int kl = (D + 1) / 2, k = 1 << kl, m = D - kl;
    // baby steps T_1 .. T_k  (depth ceil(log2 i)): T_i = 2 T_a T_b - T_(b-a) with a = floor(i/2), b = i - a, so b - a is 0 or 1
    std::vector<Elem> T(k + 1);
    T[1] = x;
    if (T[1].deg == 2) rescale_inplace(T[1]);
    for (int lo = 1; lo < k; lo *= 2) {
        std::vector<Elem> as, bs;
        for (int i = lo + 1; i <= std::min(2 * lo, k); ++i) { as.push_back(T[i / 2]); bs.push_back(T[i - i / 2]); }
        std::vector<Elem> prod = mult_each(as, bs, true);
        // doubling and the pending rescale on the whole round at once when it came back as one batch
        bool one = prod.size() > 1;
        for (size_t j = 1; j < prod.size() && one; ++j)
            one = prod[j].mem == prod[0].mem && prod[j].off == prod[0].off + j * (size_t)prod[0].batch * prod[0].words_each(P.N);
        Elem all;
        if (one) {
            all = prod[0];
            all.batch = prod[0].batch * (int)prod.size();
            all = add(all, all);
            rescale_inplace(all);
        }
        Elem t1;                                                 // T_1 at the level of this round, adjusted once
        for (size_t j = 0; j < prod.size(); ++j) {
            const int i = lo + 1 + (int)j;
            Elem p;
            if (one) {
                p = all;
                p.off = all.off + j * (size_t)prod[0].batch * all.words_each(P.N);
                p.batch = prod[0].batch;
            } else {
                p = add(prod[j], prod[j]);
                rescale_inplace(p);
            }
            if (i % 2 == 0) { T[i] = add_const(p, -1.0); continue; }
            if (!t1.valid()) { t1 = T[1]; Elem r = p; adjust_pair(t1, r); }
            T[i] = sub(p, t1);
        }
    }
    // giant steps T_{k 2^j}
    std::vector<Elem> Gs(m + 1);
    Gs[0] = T[k];
    for (int j = 1; j < m; ++j) {
        Elem p = square(Gs[j - 1]);
        p = add(p, p);
        Gs[j] = add_const(p, -1.0);
    }
    ps_settled_.clear(); ps_aligned_.clear();   // per-evaluation caches of inner_linear
    // recursive split p = q T_g + r, g = k 2^(j-1), laid out as a tree first and evaluated height by height
    struct Node { std::vector<double> p; int j = 0, q = -1, r = -1; double c0 = 0; Elem val; };
    std::vector<Node> nodes;
    std::function<int(const std::vector<double>&, int)> build = [&](const std::vector<double>& p, int j) -> int {
        const int d = poly_degree(p);
        if (j == 0 || d < k) {                   // baby polynomial (deg < k)
            Node nd; nd.p = p; nd.j = 0;
            nodes.push_back(std::move(nd));
            return (int)nodes.size() - 1;
        }
        const int g = k << (j - 1);
        if (d < g) return build(p, j - 1);
        std::vector<double> q(d - g + 1, 0.0), r(g, 0.0);
        for (int i = 0; i < g && i <= d; ++i) r[i] = p[i];
        q[0] = p[g];
        for (int i = g + 1; i <= d; ++i) { q[i - g] = 2.0 * p[i]; r[2 * g - i] -= p[i]; }
        const int qi = build(q, j - 1), ri = build(r, j - 1);
        Node nd; nd.j = j; nd.q = qi; nd.r = ri;
        nodes.push_back(std::move(nd));
        return (int)nodes.size() - 1;
    };
    const int root = build(c, m);
    {
        std::vector<size_t> leaves;
        std::vector<const std::vector<double>*> polys;
        for (size_t i = 0; i < nodes.size(); ++i)
            if (nodes[i].j == 0) { leaves.push_back(i); polys.push_back(&nodes[i].p); }
        const std::vector<Elem> vals = inner_linear_many(T, polys, k - 1);
        for (size_t i = 0; i < leaves.size(); ++i) {
            Node& nd = nodes[leaves[i]];
            nd.c0 = nd.p.empty() ? 0.0 : nd.p[0];
            nd.val = vals[i];
        }
    }
    for (int h = 1; h <= m; ++h) {
        std::vector<int> with_q;                 // nodes of this height whose quotient is a ciphertext: (q + q0) T_g in one call
        std::vector<Elem> qs;
        for (size_t i = 0; i < nodes.size(); ++i) {
            if (nodes[i].j != h) continue;
            const Node& q = nodes[nodes[i].q];
            if (!q.val.valid()) continue;
            with_q.push_back((int)i);
            qs.push_back(negligible(q.c0) ? q.val : add_const(q.val, q.c0));
        }
        const std::vector<Elem> prods = mult_each(qs, {Gs[h - 1]}, false);
        std::vector<size_t> both;                // nodes whose product and remainder are both ciphertexts: summed together below
        std::vector<Elem> sum_a, sum_b;
        for (size_t i = 0; i < nodes.size(); ++i) {
            Node& nd = nodes[i];
            if (nd.j != h) continue;
            const Node& q = nodes[nd.q];
            const Node& r = nodes[nd.r];
            Elem prod;
            const auto at = std::find(with_q.begin(), with_q.end(), (int)i);
            if (at != with_q.end()) prod = prods[(size_t)(at - with_q.begin())];
            else if (!negligible(q.c0)) prod = mult_const(Gs[h - 1], q.c0);
            nd.c0 = r.c0;
            if (prod.valid() && r.val.valid()) { both.push_back(i); sum_a.push_back(prod); sum_b.push_back(r.val); }
            else nd.val = prod.valid() ? prod : r.val;
        }
        const std::vector<Elem> sums = add_each(sum_a, sum_b);
        for (size_t g = 0; g < both.size(); ++g) nodes[both[g]].val = sums[g];
    }
    double c0 = nodes[root].c0;
    Elem res = nodes[root].val;
    ps_settled_.clear(); ps_aligned_.clear();
    if (!res.valid()) res = mult_const(x, 0.0);
    return negligible(c0) ? res : add_const(res, c0);
}

Elem Scheme::eval_chebyshev(const Elem& x, const std::vector<double>& coeffs, double a, double b) {
    if (coeffs.empty()) throw std::invalid_argument("EvalChebyshevSeries: no coefficients");
    std::vector<double> c = coeffs;
    c[0] *= 0.5;                                   // series is c0/2 + sum c_i T_i (A.10)
    Elem y = x;
    if (!(a == -1.0 && b == 1.0)) {                // affine map of [a,b] onto [-1,1]
        y = mult_const(x, 2.0 / (b - a));
        y = add_const(y, -(a + b) / (b - a));
    }
    return cheby_ps(y, c);
}

// power basis -> Chebyshev basis (exact for the small degrees the reference uses), then the same evaluator
Elem Scheme::eval_poly(const Elem& x, const std::vector<double>& a) {
    const int n = (int)a.size() - 1;
    if (n < 0) throw std::invalid_argument("EvalPoly: no coefficients");
    std::vector<double> c(n + 1, 0.0);
    // x^i = 2^(1-i) * sum'_{j = i, i-2, ...} binom(i, (i-j)/2) T_j   (the j = 0 term is halved)
    for (int i = 0; i <= n; ++i) {
        if (a[i] == 0.0) continue;
        if (i == 0) { c[0] += a[0]; continue; }
        double binom = 1.0;
        for (int t = 0; 2 * t <= i; ++t) {
            const int j = i - 2 * t;
            double w = std::ldexp(binom, 1 - i);
            if (j == 0) w *= 0.5;
            c[j] += a[i] * w;
            binom = binom * (i - t) / (t + 1);
        }
    }
    return cheby_ps(x, c);
}

}  // namespace flk
