// poly.cpp -- polynomial evaluation on ciphertexts: power-basis (EvalPoly, reference FHEController.cpp:1291)
// and Chebyshev series (EvalChebyshevFunction, FHEController.cpp:486,1319-1335) with a depth-optimal
// baby-step / giant-step (Paterson-Stockmeyer) schedule: depth ceil(log2(deg+1)), the same budget OpenFHE's
// EvalChebyshevSeriesPS needs (reference table Utils.h:127-153).
#include <cmath>

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
int poly_degree(const std::vector<double>& c) {
    int d = (int)c.size() - 1;
    while (d > 0 && negligible(c[d])) --d;
    return d;
}
}  // namespace

// sum_{i<=upto} c[i] T[i]; T[0] is the constant 1 (T[0] unused).  Returns an invalid Elem when everything is zero
// except possibly c[0], which the caller adds as a constant.
// sum_{1 <= i <= upto} c[i] T[i] as ONE kernel (weighted_sum) instead of a scalar multiplication, a rescale and an addition per
// term: the terms a combination uses are brought to the deepest level among them (never deeper, so no level is wasted) and
// to one scale -- OpenFHE's EvalChebyshevSeriesPS aligns levels the same way before its EvalLinearWSum.  `settled` holds
// T[i] after its pending rescale, `aligned` caches (i, target limbs) -> adjusted copy, so each polynomial is adjusted once
// per target level rather than once per use.
Elem Scheme::inner_linear(const std::vector<Elem>& T, const std::vector<double>& c, int upto) {
    std::vector<int> idx;
    std::vector<double> w;
    for (int i = 1; i <= upto && i < (int)c.size(); ++i) {
        if (negligible(c[i])) continue;
        idx.push_back(i);
        w.push_back(c[i]);
    }
    if (idx.empty()) return Elem();
    if (ps_settled_.size() != T.size()) { ps_settled_.assign(T.size(), Elem()); ps_aligned_.clear(); }
    int deepest = idx[0];
    for (int i : idx) {
        if (!ps_settled_[i].valid()) {
            ps_settled_[i] = T[i];
            if (ps_settled_[i].deg == 2) rescale_inplace(ps_settled_[i]);
        }
        if (ps_settled_[i].l < ps_settled_[deepest].l) deepest = i;
    }
    const Elem& ref = ps_settled_[deepest];
    std::vector<Elem> terms;
    for (int i : idx) {
        const Elem& si = ps_settled_[i];
        if (si.l == ref.l && si.scale == ref.scale) { terms.push_back(si); continue; }
        auto key = std::make_pair(i, ref.l);
        auto it = ps_aligned_.find(key);
        if (it == ps_aligned_.end()) {
            Elem a = si, r = ref;
            adjust_pair(a, r);
            it = ps_aligned_.emplace(key, a).first;
        }
        terms.push_back(it->second);
    }
    return weighted_sum(terms, w);
}

// Chebyshev series sum c[i] T_i(x) (true coefficients, c[0] not halved), x already mapped to [-1,1].
Elem Scheme::cheby_ps(const Elem& x, const std::vector<double>& c_in) {
    std::vector<double> c = c_in;
    const int n = poly_degree(c);
    c.resize(n + 1);
    if (n == 0) { Elem z = mult_const(x, 0.0); return add_const(z, c[0]); }
    int D = 0;
    while ((1 << D) < n + 1) ++D;
    const int kl = (D + 1) / 2, k = 1 << kl, m = D - kl;
    // baby steps T_1 .. T_k  (depth ceil(log2 i))
    std::vector<Elem> T(k + 1);
    T[1] = x;
    for (int i = 2; i <= k; ++i) {
        const int a = i / 2, b = i - a;
        Elem p = mult(T[a], T[b]);
        p = add(p, p);
        T[i] = a == b ? add_const(p, -1.0) : sub(p, T[1]);
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
    // recursive split p = q T_g + r
    struct Rec {
        Scheme* s; const std::vector<Elem>& T; const std::vector<Elem>& Gs; int k;
        Elem run(const std::vector<double>& p, int j, double& c0) {
            const int d = poly_degree(p);
            if (j == 0 || d < k) {          // baby polynomial (deg < k)
                c0 = p.empty() ? 0.0 : p[0];
                return s->inner_linear(T, p, std::min(d, k - 1));
            }
            const int g = k << (j - 1);
            if (d < g) return run(p, j - 1, c0);
            std::vector<double> q(d - g + 1, 0.0), r(g, 0.0);
            for (int i = 0; i < g && i <= d; ++i) r[i] = p[i];
            q[0] = p[g];
            for (int i = g + 1; i <= d; ++i) { q[i - g] = 2.0 * p[i]; r[2 * g - i] -= p[i]; }
            double q0 = 0, r0 = 0;
            Elem qe = run(q, j - 1, q0);
            Elem re = run(r, j - 1, r0);
            // (qe + q0) * T_g + re + r0
            Elem prod;
            if (qe.valid()) {
                Elem qq = negligible(q0) ? qe : s->add_const(qe, q0);
                prod = s->mult(qq, Gs[j - 1]);
            } else if (!negligible(q0)) {
                prod = s->mult_const(Gs[j - 1], q0);
            }
            c0 = r0;
            if (prod.valid() && re.valid()) return s->add(prod, re);
            return prod.valid() ? prod : re;
        }
    } rec{this, T, Gs, k};
    double c0 = 0;
    Elem res = rec.run(c, m, c0);
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
