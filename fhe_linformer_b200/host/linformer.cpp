// linformer.cpp -- encrypted forward pass of the 1-layer Linformer classifier (d = 128, k = 32 projected keys, FFN 512,
// CLS-only attention) on FHEController.  Restates the circuit of the reference's src/main.cpp:145-475 stage by stage;
// the line each step follows is cited as M:<line>.  Slot layouts are those of SURVEY.md section 3.3.
#include "linformer.h"

#include <cstdio>
#include <cstdlib>

#include <cmath>
#include <filesystem>
#include <tuple>

namespace flh {

namespace fs = std::filesystem;

LinformerForward::LinformerForward(FHEController& controller, LinformerFiles files, bool verbose)
    : fc_(controller), files_(std::move(files)), verbose_(verbose), t0_(utils::start_time()) {}

double LinformerForward::scalar(const std::string& path) const {
    std::ifstream in(path);
    double v;
    if (!(in >> v)) {   // M:30-38
        std::cerr << "cannot read a value from " << path << std::endl;
        std::exit(1);
    }
    return v;
}

void LinformerForward::checkpoint(const std::string& name, const Ctxt& c) {
    static const bool trace = std::getenv("FLH_TRACE_LEVELS") != nullptr;
    if (trace) std::fprintf(stderr, "  [levels] %-28s level %2d  noise degree %d\n", name.c_str(), (int)c->GetLevel(), (int)c->GetNoiseScaleDeg());
    if (sink_) sink_(name, fc_.decrypt_tovector(c, fc_.num_slots), (int)c->GetLevel());
}

void LinformerForward::lap(const std::string& name) {
    fl_sync(fc_.native());
    const double s = std::chrono::duration<double>(utils::clock_type::now() - t0_).count();
    times_.push_back({name, s});
    if (verbose_) std::cout << "The evaluation of " << name << " took: " << s << " seconds." << std::endl;
    t0_ = utils::start_time();
}

std::vector<Ctxt> LinformerForward::load_expanded(const std::string& dir, const std::string& stem, int count) {
    // M:159-173 encrypts file by file; the rows are independent, so they are encoded one by one and encrypted as one batch
    std::vector<std::string> files;
    files.reserve(count);
    for (int i = 0; i < count; ++i) files.push_back(dir + "/" + stem + std::to_string(i) + ".txt");
    return fc_.read_expanded_inputs(files);
}

std::vector<Ctxt> LinformerForward::load_expanded(const std::vector<std::string>& dirs, const std::string& stem, int count) {
    std::vector<std::vector<std::string>> files(dirs.size());
    for (size_t m = 0; m < dirs.size(); ++m)
        for (int i = 0; i < count; ++i) files[m].push_back(dirs[m] + "/" + stem + std::to_string(i) + ".txt");
    return fc_.read_expanded_inputs_many(files);
}

// ---- Linformer projection under encryption (F1): row i of X_E is sum_t E[i][t] rows[t] + E_b[i], still in the Expanded layout --
std::vector<Ctxt> LinformerForward::project(const std::vector<Ctxt>& rows, const std::string& which) {
    const std::vector<double> flat = utils::read_values_from_file(layer("selfAttn_" + which + "_weight.txt"));
    const std::vector<double> b = utils::read_values_from_file(layer("selfAttn_" + which + "_bias.txt"));
    const size_t S = rows.size(), width = flat.size() / 32;   // 32 x 701 in the reference's checkpoint
    if (flat.size() % 32 != 0 || width < S || b.size() < 32) throw std::runtime_error("projection weights: expected 32 x (>= S) and 32 biases");
    std::vector<std::vector<double>> w(32, std::vector<double>(S));
    for (int i = 0; i < 32; ++i)
        for (size_t t = 0; t < S; ++t) w[i][t] = flat[(size_t)i * width + t];
    std::vector<Ctxt> out = fc_.project_rows(rows, w, {});
    for (int i = 0; i < 32; ++i) out[i] = fc_.add(out[i], fc_.encode(b[i], (int)out[i]->GetLevel() + 1, 0));   // product is rescaled lazily: bias one level lower
    return out;
}

// ---- attention for the CLS query only (M:176-215) -------------------------------------------------------------------------
// K = X_E W_K, V = X_F W_V on the 32 client-projected rows; scores = softmax-like exp / sum over the 32 keys; context = scores V.
Ctxt LinformerForward::attend_cls(const std::vector<Ctxt>& rows, const std::vector<Ctxt>& xe, const std::vector<Ctxt>& xf) {
    const Ptxt wq = fc_.read_plain_input(layer("selfAttn_WQ_weight_T.txt"));            // M:177
    const Ptxt bq = fc_.read_plain_repeated_input(layer("selfAttn_WQ_bias.txt"));        // M:178
    const Ptxt wk = fc_.read_plain_input(layer("selfAttn_WK_weight_T.txt"));            // M:179
    const Ptxt bk = fc_.read_plain_repeated_input(layer("selfAttn_WK_bias.txt"));        // M:180

    // The reference projects every row to a query (M:182) and then uses row 0 only (M:196).
    const std::vector<Ctxt> queries = fc_.matmulRE(dead_work_ && !packed_ ? rows : std::vector<Ctxt>{rows[0]}, wq, bq);
    const Ctxt keys = fc_.wrapUpRepeated(fc_.matmulRE(xe, wk, bk));                     // M:183-185
    checkpoint("query_cls", queries[0]);
    checkpoint("keys_wrapped", keys);

    Ctxt scores = fc_.matmulScores(queries[0], keys);                                    // M:196
    checkpoint("scores_raw", scores);
    scores = fc_.eval_exp(scores, 32);                                                   // M:197
    checkpoint("scores_exp", scores);
    const Ctxt total = fc_.rotsum(scores, 32, 128);                                      // M:201
    const Ctxt inverse = fc_.eval_inverse_naive(total, -1, 128);                         // M:203
    checkpoint("scores_inverse", inverse);
    scores = fc_.mult(scores, inverse);                                                  // M:205
    checkpoint("scores_normalised", scores);
    const std::vector<Ctxt> weights = fc_.unwrapExpanded(scores, 1);                     // M:207

    const Ptxt wv = fc_.read_plain_input(layer("selfAttn_WV_weight_T.txt"));            // M:209
    const Ptxt bv = fc_.read_plain_repeated_input(layer("selfAttn_WV_bias.txt"));        // M:210
    const Ctxt values = fc_.wrapUpRepeated(fc_.matmulRE(xf, wv, bv));                   // M:212-213
    checkpoint("values_wrapped", values);
    return fc_.matmulRE(weights, values, 128, 128)[0];                                   // M:215-216
}

// ---- packed mode: the same attention block, never leaving the wrapped-expanded layout --------------------------------------------
// The 32 projected rows of X_E (X_F) enter ONE ciphertext, row t in column t (position masks, no rotation; columns 32..127 zero):
//   K = X_E W_K, V = X_F W_V, q = W_Q x_0 / 64 and finally W_O are FHEController::packed_linear products (BSGS, 23 key switches each)
//     instead of 32 + 32 + 1 + 1 rotate-and-sum ladders of 7 rotations;
//   scores  = sum over the 128 features (stride-128 ladder) of q (.) K: score t in column t of every block, 0 in columns >= 32;
//   exp     = the reference's polynomial, and -1 in the columns that hold no key (its eval_exp does the same for its layout);
//   total   = stride-1 ladder over 32 columns: column t receives the sum over keys t..31 -- the SAME partial sums the reference's
//             stride-128 ladder over its 32 blocks produces (M:201: block t sums blocks t..t+31, of which t..31 hold keys), so
//             weight t = exp(s_t) / sum_{t' >= t} exp(s_t') exactly as there;
//   context = stride-1 ladder of weights (.) V: the attention output of the CLS query in column 0 of every block.
// Same polynomial for exp, same Chebyshev interpolant for 1/x on the same totals as M:197-205: the logits agree up to CKKS noise.
// Returns row 0 after W_O, bias and residual (M:217-239) in the wrapped-expanded layout: column 0 is what "attended_row0" holds in
// every column; the other columns are never read (wrap_rows_packed keeps column 0 of row 0).
Ctxt LinformerForward::attend_cls_packed(const std::vector<Ctxt>& rows, const std::vector<Ctxt>& xe, const std::vector<Ctxt>& xf) {
    const int slots = fc_.num_slots;
    auto expanded = [&](const std::vector<double>& v, double scale, int columns) {        // slot 128 j + t = v[j] for t < columns
        std::vector<double> e((size_t)slots, 0.0);
        for (int j = 0; j < 128; ++j)
            for (int t = 0; t < columns; ++t) e[(size_t)128 * j + t] = v.at((size_t)j) * scale;
        return e;
    };
    auto linear = [&](const Ctxt& x, const std::string& weight_file, const std::string& bias_file, bool transposed_file, double scale, int columns) {
        std::vector<double> f;
        auto weight = [&](int j, int i) {                                                 // coefficient of input feature j in output feature i
            if (f.empty()) { f = utils::read_values_from_file(layer(weight_file)); if (f.size() < 128 * 128) throw std::runtime_error(weight_file + ": expected 128 x 128 values"); }
            return transposed_file ? f[(size_t)128 * j + i] : f[(size_t)128 * i + j];
        };
        const Ctxt y = fc_.packed_linear(x, layer(weight_file) + (scale == 1.0 ? "" : "@*" + std::to_string(scale)), weight, scale);
        return fc_.add(y, fc_.encode(expanded(utils::read_values_from_file(layer(bias_file)), scale, columns), (int)y->GetLevel() + 1, slots));
    };
    const Ctxt keys = linear(wrap_rows_packed(xe, 0), "selfAttn_WK_weight_T.txt", "selfAttn_WK_bias.txt", true, 1.0, 32);   // M:179-185
    const Ctxt query = linear(rows[0], "selfAttn_WQ_weight_T.txt", "selfAttn_WQ_bias.txt", true, 1.0 / 64.0, 128);          // M:177-182, 1/64 of matmulScores folded in
    checkpoint("packed_keys", keys);
    checkpoint("packed_query", query);
    Ctxt scores = fc_.rotsum(fc_.mult(query, keys), 128, 128);                            // M:196
    checkpoint("packed_scores", scores);
    scores = fc_.eval_exp(scores, 0);                                                     // M:197
    std::vector<double> no_key((size_t)slots, 0.0);
    for (int p = 0; p < slots; ++p)
        if (p % 128 >= 32) no_key[(size_t)p] = -1.0;
    scores = fc_.add(scores, fc_.encode(no_key, (int)scores->GetLevel(), slots));
    checkpoint("packed_scores_exp", scores);
    const Ctxt inverse = fc_.eval_inverse_naive(fc_.rotsum(scores, 32, 1), -1, 128);      // M:201-203
    checkpoint("packed_scores_inverse", inverse);
    scores = fc_.mult(scores, inverse);                                                   // M:205
    checkpoint("packed_scores_normalised", scores);
    const Ctxt values = linear(wrap_rows_packed(xf, 0), "selfAttn_WV_weight_T.txt", "selfAttn_WV_bias.txt", true, 1.0, 32);   // M:209-213
    const Ctxt context = fc_.rotsum(fc_.mult(scores, values), 32, 1);                     // M:215-216
    checkpoint("packed_attention_cls", context);
    lap("Self-Attention");
    const Ctxt out = fc_.add(linear(context, "selfAttn_WO_weight.txt", "selfAttn_WO_bias.txt", false, 1.0, 128), rows[0]);   // M:231-239
    checkpoint("packed_attended_row0", out);
    return out;
}

// ---- attention for every row (src/main_2.cpp:187-229, "M2:") ---------------------------------------------------------------
// Queries go in two halves of at most 128; matmulScores(vector) packs query t's 32 scores at slots 128 j + t.
std::vector<Ctxt> LinformerForward::attend_all(const std::vector<Ctxt>& rows, const std::vector<Ctxt>& xe, const std::vector<Ctxt>& xf) {
    const Ptxt wq = fc_.read_plain_input(layer("selfAttn_WQ_weight_T.txt"));
    const Ptxt bq = fc_.read_plain_repeated_input(layer("selfAttn_WQ_bias.txt"));
    const Ptxt wk = fc_.read_plain_input(layer("selfAttn_WK_weight_T.txt"));
    const Ptxt bk = fc_.read_plain_repeated_input(layer("selfAttn_WK_bias.txt"));
    const std::vector<Ctxt> queries = fc_.matmulRE(rows, wq, bq);                         // M2:182
    const Ctxt keys = fc_.wrapUpRepeated(fc_.matmulRE(xe, wk, bk));                      // M2:183-185
    std::vector<Ctxt> weights;
    for (int half = 0; half < 2; ++half) {                                               // M2:187-216
        const size_t lo = half ? 128 : 0, hi = half ? queries.size() : std::min<size_t>(128, queries.size());
        if (lo >= hi) break;
        const std::vector<Ctxt> q(queries.begin() + lo, queries.begin() + hi);
        Ctxt scores = fc_.eval_exp(fc_.matmulScores(q, keys), (int)q.size());            // M2:196-200
        if (half == 0) checkpoint("all_scores_exp_0", scores);
        const Ctxt inverse = fc_.eval_inverse_naive(fc_.rotsum(scores, 32, 128), -1, 190000);   // M2:202-211
        scores = fc_.mult(scores, inverse);                                              // M2:213-214
        if (half == 0) checkpoint("all_scores_normalised_0", scores);
        const std::vector<Ctxt> un = fc_.unwrapExpanded(scores, (int)q.size());          // M2:216-217
        weights.insert(weights.end(), un.begin(), un.end());
    }
    const Ptxt wv = fc_.read_plain_input(layer("selfAttn_WV_weight_T.txt"));
    const Ptxt bv = fc_.read_plain_repeated_input(layer("selfAttn_WV_bias.txt"));
    const Ctxt values = fc_.wrapUpRepeated(fc_.matmulRE(xf, wv, bv));                   // M2:226-227
    return fc_.matmulRE(weights, values, 128, 128);                                      // M2:229
}

// ---- W_O, bias and residual (M:217-239): only row 0 carries attention output, the other rows are encryptions of zero ------
std::vector<Ctxt> LinformerForward::self_output(const Ctxt& cls_context, const std::vector<Ctxt>& rows) {
    const int level = (int)cls_context->GetLevel();
    std::vector<Ctxt> out;
    out.reserve(rows.size());
    out.push_back(cls_context);
    const Ctxt zero = fc_.encrypt_ptxt(fc_.encode(0, level, 0));                         // M:220-221
    for (size_t i = 1; i < rows.size(); ++i) out.push_back(zero->Clone());               // M:222-224
    const Ptxt wo = fc_.read_plain_input(layer("selfAttn_WO_weight.txt"), level);        // M:231
    const Ptxt bo = fc_.read_plain_expanded_input(layer("selfAttn_WO_bias.txt"), level + 1);   // M:232
    if (dead_work_) {
        out = fc_.matmulCR(out, wo, nullptr);                                            // M:235
    } else {
        // W_O applied to an encryption of zero is an encryption of zero at the product's level: do one and share it
        std::vector<Ctxt> two = fc_.matmulCR({out[0], out.size() > 1 ? out[1] : out[0]}, wo, nullptr);
        for (size_t i = 0; i < out.size(); ++i) out[i] = i == 0 ? two[0] : two[1];
    }
    out[0] = fc_.add(out[0], bo);                                                        // M:236
    for (size_t i = 0; i < out.size(); ++i) out[i] = fc_.add(out[i], rows[i]);           // M:237-239
    return out;
}

// ---- rows -> two wrapped ciphertexts (first 128 rows, the rest), the affine stand-in for LayerNorm, optional bootstrap ----
// M:292-320 (affine1 + bootstrap) and M:382-414 (affine2, no bootstrap; the caller adds the residual in between)
std::pair<Ctxt, Ctxt> LinformerForward::affine_and_refresh(const std::vector<Ctxt>& rows, const std::string& which, bool refresh) {
    if (rows.size() <= 128 || rows.size() > 256) throw std::invalid_argument("the circuit packs rows as 128 + (S - 128): need 129 <= S <= 256");
    const std::vector<Ctxt> first(rows.begin(), rows.begin() + 128), rest(rows.begin() + 128, rows.end());
    // packed mode builds the same two wrapped ciphertexts from position masks alone (no rotate(-1) chain)
    Ctxt w0 = packed_ ? wrap_rows_packed(first, 0) : fc_.wrapUpExpanded(first);          // M:307-308 / M:392-393
    const bool first_half_only = packed_ && !dead_work_;      // packed + lean: nothing downstream reads the second half (see encoder())
    Ctxt w1 = first_half_only ? Ctxt() : packed_ ? wrap_rows_packed(rest, 0) : fc_.wrapUpExpanded(rest);
    if (!refresh) return {w0, w1};
    const double s = (double)rows.size();
    const double f = scalar(layer("ffn_" + which + "_c0.txt")) + scalar(layer("ffn_" + which + "_c1.txt")) / std::sqrt(s) +
                     scalar(layer("ffn_" + which + "_c2.txt")) / s;                      // M:292-297
    // (in packed mode the second half consists of fresh rows only and sits at a shallower level than the first: each half
    // gets the plaintexts of its own level; in the reference's flow both levels coincide)
    auto affine = [&](const Ctxt& c) {
        const Ptxt a = fc_.read_plain_repeated_input(layer("ffn_" + which + "_a.txt"), (int)c->GetLevel(), f);       // M:311
        const Ptxt b = fc_.read_plain_repeated_input(layer("ffn_" + which + "_b.txt"), (int)c->GetLevel() + 1, f);   // M:312
        return fc_.add(fc_.mult(c, a), b);
    };
    w0 = affine(w0);                                                                     // M:314-315
    checkpoint(which + "_0", w0);
    if (first_half_only) return {fc_.bootstrap(w0), Ctxt()};
    w1 = affine(w1);                                                                     // M:316-317
    checkpoint(which + "_1", w1);
    if (samples() > 1) return {fc_.bootstrap(w0), fc_.bootstrap(w1)};                    // each half is already a batch over the samples
    const std::vector<Ctxt> fresh = fc_.per_row({w0, w1}, [&](const Ctxt& c) { return fc_.bootstrap(c); });   // M:319-320, as one batch
    return {fresh[0], fresh[1]};
}

// ---- packed mode --------------------------------------------------------------------------------------------------------------
// Expanded rows (slot 128 j + i = x_t[j] for every i) -> wrapped-expanded (slot 128 j + t = x_t[j]): keep column t of row t.  The
// reference gets there with one mask at column 0 and a rotate(-1) Horner chain (F.cpp:1070-1084); the masks at column t need no
// rotation at all.  Row 0 of the first half comes out of the attention block at a deeper level than the fresh rows: the fresh
// ones are summed first so that only one level alignment happens.
Ctxt LinformerForward::wrap_rows_packed(const std::vector<Ctxt>& rows, int) {
    std::vector<Ctxt> fresh;
    for (size_t t = 1; t < rows.size(); ++t) fresh.push_back(fc_.mask_mod_n(rows[t], 128, (int)t, 0));
    const Ctxt head = fc_.mask_mod_n(rows[0], 128, 0, 0);
    if (fresh.empty()) return head;
    return fc_.add(head, fresh.size() == 1 ? fresh[0] : fc_.add(fresh));
}

// FFN on the two wrapped-expanded halves, never leaving that layout: 128 -> 512 as four packed_linear transforms (the four
// 128-column blocks of W0, scaled by 1/8), GELU and bootstrap on the 2 x 4 hidden ciphertexts as one batched operand, 512 -> 128 as
// four packed_linear transforms summed.  Replaces unwrapExpanded + matmulRElarge + generate_containers + unwrapRepeatedLarge +
// matmulCRlarge + wrapUpExpanded (M:325-393; ~100 S rotations) by 8 BSGS products per half (~22 hoisted rotations each).
std::pair<Ctxt, Ctxt> LinformerForward::feed_forward_packed(const Ctxt& half0, const Ctxt& half1, const std::vector<double>* cls_column_scale) {
    const double gelu_scale = 1.0 / 8.0;                                                 // M:334
    const int slots = fc_.num_slots;
    const int halves = cls_column_scale ? 1 : 2;                                         // lean tail: the half with the CLS row only
    const Ctxt x = halves == 1 ? half0 : fc_.pack({half0, half1});
    const std::vector<double> b0 = utils::read_values_from_file(layer("ffn_Wffn_0_bias.txt"));   // 512 values (M:345)
    if (b0.size() < 512) throw std::runtime_error("ffn_Wffn_0_bias.txt: expected 512 values");
    std::vector<Ctxt> hidden;
    for (int b = 0; b < 4; ++b) {
        const std::string name = "ffn_W0_transposed_block_" + std::to_string(b);
        std::vector<double> f;
        auto weight = [&](int j, int i) {
            if (f.empty()) { f = utils::read_values_from_file(w(name + ".txt")); if (f.size() < 128 * 128) throw std::runtime_error(name + ": expected 128 x 128 values"); }
            return f[(size_t)128 * j + i];                                               // y = x . F  (M:338-341)
        };
        Ctxt u = fc_.packed_linear(x, w(name), weight, gelu_scale);                      // plan cached per weight FILE
        std::vector<double> bias((size_t)slots);
        for (int i = 0; i < 128; ++i)
            for (int t = 0; t < 128; ++t) bias[(size_t)128 * i + t] = b0[(size_t)128 * b + i] * gelu_scale;
        hidden.push_back(fc_.add(u, fc_.encode(bias, (int)u->GetLevel() + 1, slots)));   // product rescaled lazily: bias one level lower
    }
    if (sink_) checkpoint("packed_hidden_block0", halves == 1 ? hidden[0] : fc_.unpack(hidden[0])[0]);
    // GELU (M:362) on all 2 x 4 hidden ciphertexts at once.  The reference refreshes every container right here (M:363) because its
    // unwrap / W2 / re-wrap chain still costs five levels; the packed chain needs two (W2, affine2), so the refresh moves behind
    // the second affine, where ONE ciphertext (the half that holds the CLS row) is left to bootstrap instead of eight.
    const Ctxt act = fc_.eval_gelu_function(fc_.pack(hidden), -1, 1, gelu_scale, 119);
    const std::vector<Ctxt> parts = fc_.unpack(act, samples());                          // [block b][half h] -> index halves b + h (one element per sample each)
    checkpoint("packed_gelu_block0", parts[0]);
    lap("Intermediate");
    Ctxt sum;
    for (int b = 0; b < 4; ++b) {
        const std::string name = "ffn_W2_block_" + std::to_string(b);
        std::vector<double> f;
        auto weight = [&](int j, int i) {
            if (f.empty()) { f = utils::read_values_from_file(w(name + ".txt")); if (f.size() < 128 * 128) throw std::runtime_error(name + ": expected 128 x 128 values"); }
            return f[(size_t)128 * i + j];                                               // y = F . x  (M:373-376)
        };
        // lean tail (cls_column_scale): only the half that holds the CLS row goes through W2, with the second affine's factor and
        // the column mask folded into the diagonals (one plan per value of the factor: it depends on S through c1, c2)
        const Ctxt y = cls_column_scale
                           ? fc_.packed_linear(parts[b], w(name) + "@cls*" + std::to_string((*cls_column_scale)[0]), weight, 1.0, cls_column_scale)
                           : fc_.packed_linear(fc_.pack({parts[2 * b], parts[2 * b + 1]}), w(name), weight, 1.0);
        sum = b == 0 ? y : fc_.add(sum, y);
    }
    const std::vector<double> b2 = utils::read_values_from_file(layer("ffn_Wffn_2_bias.txt"));   // M:378
    std::vector<double> bias((size_t)slots);
    for (int i = 0; i < 128; ++i)
        for (int t = 0; t < 128; ++t) bias[(size_t)128 * i + t] = b2.at((size_t)i) * (cls_column_scale ? (*cls_column_scale)[(size_t)t] : 1.0);
    sum = fc_.add(sum, fc_.encode(bias, (int)sum->GetLevel() + 1, slots));
    if (cls_column_scale) return {sum, Ctxt()};
    const std::vector<Ctxt> out = fc_.unpack(sum, samples());
    return {out[0], out[1]};
}

// ---- FFN: 128 -> 512 (scaled by 1/8 so GELU's argument lies in [-1, 1]), GELU, bootstrap, 512 -> 128 (M:325-380) ----------
std::vector<Ctxt> LinformerForward::feed_forward(const Ctxt& half0, const Ctxt& half1, int rows) {
    const double gelu_scale = 1.0 / 8.0;                                                 // M:334
    std::vector<Ctxt> x0 = fc_.unwrapExpanded(half0, 128);                               // M:325
    std::vector<Ctxt> x1 = fc_.unwrapExpanded(half1, rows - 128);                        // M:326
    checkpoint("self_output_row0", x0[0]);
    lap("Self-Output");

    const int level = (int)half0->GetLevel();
    std::vector<Ptxt> w0;
    for (int b = 0; b < 4; ++b) w0.push_back(fc_.read_plain_input(w("ffn_W0_transposed_block_" + std::to_string(b) + ".txt"), level, gelu_scale));   // M:338-341
    const Ptxt b0 = fc_.read_plain_input(layer("ffn_Wffn_0_bias.txt"), level + 1, gelu_scale);   // M:345
    std::vector<Ctxt> hidden = fc_.matmulRElarge(x0, w0, b0);                            // M:347
    std::vector<Ctxt> hidden1 = fc_.matmulRElarge(x1, w0, b0);                           // M:348
    hidden.insert(hidden.end(), hidden1.begin(), hidden1.end());                         // M:351-354
    checkpoint("hidden_row0", hidden[0]);

    std::vector<Ctxt> containers = fc_.generate_containers(hidden, nullptr);             // M:358
    checkpoint("container0_pre_gelu", containers[0]);
    // M:360-364 loops over the containers; they are independent and share level and scale, so GELU and the bootstrap run on
    // all of them as one batched operand
    containers = fc_.per_row(containers, [&](const Ctxt& c) { return fc_.bootstrap(fc_.eval_gelu_function(c, -1, 1, gelu_scale, 119)); });
    checkpoint("container0_gelu", containers[0]);
    std::vector<std::vector<Ctxt>> quads = fc_.unwrapRepeatedLarge(containers, rows);    // M:366
    lap("Intermediate");

    const int l2 = (int)quads[0][0]->GetLevel();
    std::vector<Ptxt> w2;
    for (int b = 0; b < 4; ++b) w2.push_back(fc_.read_plain_input(w("ffn_W2_block_" + std::to_string(b) + ".txt"), l2));   // M:373-376
    const Ptxt b2 = fc_.read_plain_expanded_input(layer("ffn_Wffn_2_bias.txt"), l2 + 1); // M:378
    return fc_.matmulCRlarge(quads, w2, b2);                                             // M:380
}

Ctxt LinformerForward::encoder() {
    t0_ = utils::start_time();
    int found = 0;                                                                       // M:147-154: one row per file, plus CLS
    for (const auto& entry : fs::directory_iterator(files_.tokens))
        if (entry.path().filename().string().rfind("input_", 0) == 0) ++found;
    if (token_limit_ > 0 && found > token_limit_) found = token_limit_;
    tokens_ = found + 1;
    if (packed_ && all_tokens_) throw std::invalid_argument("packed mode evaluates the CLS-query attention of main.cpp; the all-token circuit has no packed form");
    if (verbose_) std::cout << tokens_ << " inputs found!" << std::endl << std::endl;

    std::vector<Ctxt> rows;
    rows.push_back(fc_.read_expanded_input(w("cls_token.txt")));                         // M:170
    std::vector<Ctxt> xe, xf;
    if (more_.empty()) {
        const std::vector<Ctxt> embedded = load_expanded(files_.tokens, "input_", found);    // M:171-173
        rows.insert(rows.end(), embedded.begin(), embedded.end());
        if (!encrypted_projection_) {
            xe = load_expanded(files_.input, "XE_", 32);                                 // M:159-162
            xf = load_expanded(files_.input, "XF_", 32);                                 // M:164-167
        }
    } else {
        // several samples per call: row t of every sample in one batched ciphertext (the CLS row is a model weight: one for all)
        if (!packed_ || all_tokens_ || encrypted_projection_ || sink_)
            throw std::invalid_argument("several samples per call: packed mode only (no all-token attention, no encrypted projection, no checkpoints)");
        std::vector<std::string> token_dirs{files_.tokens}, input_dirs{files_.input};
        for (const auto& s : more_) { input_dirs.push_back(s.first); token_dirs.push_back(s.second); }
        for (const std::string& d : token_dirs) {
            int n = 0;
            for (const auto& entry : fs::directory_iterator(d))
                if (entry.path().filename().string().rfind("input_", 0) == 0) ++n;
            if (std::min(n, token_limit_ > 0 ? token_limit_ : n) != found) throw std::invalid_argument("several samples per call: the samples differ in their number of rows");
        }
        const std::vector<Ctxt> embedded = load_expanded(token_dirs, "input_", found);
        rows.insert(rows.end(), embedded.begin(), embedded.end());
        xe = load_expanded(input_dirs, "XE_", 32);
        xf = load_expanded(input_dirs, "XF_", 32);
    }
    lap("Encrypt");
    if (encrypted_projection_) {
        xe = project(rows, "E");
        xf = project(rows, "F");
        checkpoint("projected_E0", xe[0]);
        lap("Projection");
    }

    std::vector<Ctxt> attended;
    if (all_tokens_) {
        std::vector<Ctxt> context = attend_all(rows, xe, xf);
        checkpoint("attention_cls", context[0]);
        checkpoint("attention_row1", context[1]);
        lap("Self-Attention");
        const int level = (int)context[0]->GetLevel();
        const Ptxt wo = fc_.read_plain_input(layer("selfAttn_WO_weight.txt"), level);            // M2:239
        const Ptxt bo = fc_.read_plain_expanded_input(layer("selfAttn_WO_bias.txt"), level + 1); // M2:240
        attended = fc_.matmulCR(context, wo, bo);                                                // M2:242
        for (size_t i = 0; i < attended.size(); ++i) attended[i] = fc_.add(attended[i], rows[i]); // M2:244-246
    } else if (packed_) {
        attended = rows;                                                                 // the other rows go on as the fresh inputs they are
        attended[0] = attend_cls_packed(rows, xe, xf);
    } else {
        const Ctxt context = attend_cls(rows, xe, xf);
        checkpoint("attention_cls", context);
        lap("Self-Attention");
        attended = self_output(context, rows);
    }
    if (!packed_) checkpoint("attended_row0", attended[0]);                              // (packed: "packed_attended_row0" above)
    checkpoint("attended_row1", attended[1]);
    auto [half0, half1] = affine_and_refresh(attended, "affine1", true);
    checkpoint("affine1_refreshed_0", half0);
    const Ctxt residual0 = half0->Clone(), residual1 = half1 ? half1->Clone() : Ctxt();  // M:322-323

    Ctxt o0, o1;
    if (packed_ && !dead_work_) {
        // Packed AND lean: from W2 on only the CLS row is read (M:416-424), i.e. column 0 of the first half.  The second affine --
        // (ffn + h) a + b with a, b indexed by position t -- and the column mask of unwrapExpanded(., 1) are folded into the W2
        // diagonals and into one plaintext product of the residual, so the chain GELU (7 levels) + W2 + pooler product fits the levels
        // left after the first refresh: the pooler's own bootstrap (M:441) is the only other one of the forward (3 instead of 10).
        lap("Self-Output");
        const double sl = (double)tokens_;
        const double f2l = scalar(layer("ffn_affine2_c0.txt")) + scalar(layer("ffn_affine2_c1.txt")) / std::sqrt(sl) + scalar(layer("ffn_affine2_c2.txt")) / sl;   // M:398-403
        const std::vector<double> a = utils::read_values_from_file(layer("ffn_affine2_a.txt")), b = utils::read_values_from_file(layer("ffn_affine2_b.txt"));
        std::vector<double> col(128, 0.0);
        col[0] = a.at(0) * f2l;                                                          // R(a)[128 j + t] = a[t] meets column t = 0
        std::vector<double> am((size_t)fc_.num_slots, 0.0), bm((size_t)fc_.num_slots, 0.0);
        for (int j = 0; j < 128; ++j) { am[(size_t)128 * j] = col[0]; bm[(size_t)128 * j] = b.at(0) * f2l; }
        Ctxt y = feed_forward_packed(half0, half1, &col).first;                          // a[0] f (W2 g + b2) on column 0
        const Ctxt res = fc_.mult(residual0, fc_.encode(am, (int)residual0->GetLevel(), fc_.num_slots));   // a[0] f h on column 0
        y = fc_.add(fc_.add(y, res), fc_.encode(bm, (int)y->GetLevel() + 1, fc_.num_slots));
        checkpoint("packed_affine2_cls", y);
        const Ctxt cls = fc_.repeat(y, 128);                                             // F.cpp:1086-1100 for index 0: replicate the kept column
        checkpoint("encoder_out", cls);
        lap("Output");
        return cls;
    }
    if (packed_) {
        lap("Self-Output");
        std::tie(o0, o1) = feed_forward_packed(half0, half1, nullptr);
        checkpoint("packed_ffn_0", o0);
    } else {
        const std::vector<Ctxt> ffn = feed_forward(half0, half1, tokens_);
        checkpoint("ffn_row0", ffn[0]);
        std::tie(o0, o1) = affine_and_refresh(ffn, "affine2", false);
    }
    o0 = fc_.add(o0, residual0);                                                         // M:395
    o1 = fc_.add(o1, residual1);                                                         // M:396
    const double s = (double)tokens_;
    const double f2 = scalar(layer("ffn_affine2_c0.txt")) + scalar(layer("ffn_affine2_c1.txt")) / std::sqrt(s) + scalar(layer("ffn_affine2_c2.txt")) / s;   // M:398-403
    const Ptxt a2 = fc_.read_plain_repeated_input(layer("ffn_affine2_a.txt"), (int)o0->GetLevel(), f2);       // M:407
    const Ptxt b2 = fc_.read_plain_repeated_input(layer("ffn_affine2_b.txt"), (int)o1->GetLevel() + 1, f2);   // M:408
    o0 = fc_.add(fc_.mult(o0, a2), b2);                                                  // M:410,413
    o1 = fc_.add(fc_.mult(o1, a2), b2);                                                  // M:411,414
    checkpoint("affine2_0", o0);
    if (packed_) o0 = fc_.bootstrap(o0);                                                  // the refresh of M:363, moved here (see feed_forward_packed)
    std::vector<Ctxt> final0 = fc_.unwrapExpanded(o0, dead_work_ && !packed_ ? 128 : 1);  // M:416 (only element 0 is used)
    if (dead_work_ && !packed_) (void)fc_.unwrapExpanded(o1, tokens_ - 128);             // M:417
    checkpoint("encoder_out", final0[0]);
    lap("Output");
    return final0[0];                                                                    // M:424 (the caller may save() it, M:422)
}

Ctxt LinformerForward::pooler(const Ctxt& encoded) {
    const double tanh_scale = all_tokens_ ? 1.0 / 18.0 : 1.0 / 50;                       // M:430 (main_2.cpp:386: 1/18)
    const int level = (int)encoded->GetLevel();
    const Ptxt weight = fc_.read_plain_input(w("pooler_dense_weight_T.txt"), level, tanh_scale);          // M:432
    const Ptxt bias = fc_.read_plain_repeated_input(w("pooler_dense_bias.txt"), level + 1, tanh_scale);   // M:433
    Ctxt y = fc_.add(fc_.rotsum(fc_.mult(encoded, weight), 128, 128), bias);             // M:435-439
    checkpoint("pooler_pre_tanh", y);
    y = fc_.eval_tanh_function(fc_.bootstrap(y), -1, 1, tanh_scale, 300);                // M:441-445
    checkpoint("pooler_out", y);
    lap("Pooler");
    return y;
}

Ctxt LinformerForward::classifier(const Ctxt& pooled) {
    const int level = (int)pooled->GetLevel();
    const Ptxt weight = fc_.read_plain_input(w("fcLinear_0_weight.txt"), level);         // M:454
    const Ptxt bias = fc_.read_plain_expanded_input(w("fcLinear_0_bias.txt"), level);    // M:455
    Ctxt y = fc_.add(fc_.rotsum(fc_.mult(pooled, weight), 128, 1), bias);                // M:457-461
    std::vector<double> pick((size_t)fc_.num_slots, 0.0);                                // M:463-470: keep slots 128 i, i < 20
    for (int i = 0; i < 20; ++i) pick[(size_t)i * 128] = 1;
    if (all_tokens_) y = fc_.mult(y, fc_.encode(pick, (int)y->GetLevel(), fc_.num_slots));   // main_2.cpp:427: plaintext mask
    else y = fc_.mult(y, fc_.encrypt(pick, (int)y->GetLevel()));                         // M:472 (ciphertext mask, as in main.cpp)
    lap("Classifier");
    return y;
}

std::vector<double> LinformerForward::logits(const Ctxt& classified, int classes) {
    const std::vector<double> slots = fc_.decrypt_tovector(classified, fc_.num_slots);   // M:115
    std::vector<double> out((size_t)classes);
    for (int i = 0; i < classes; ++i) out[i] = slots[(size_t)i * 128];                   // M:120-123
    return out;
}

std::vector<std::vector<double>> LinformerForward::logits_many(const Ctxt& classified, int classes) {
    std::vector<std::vector<double>> out;
    for (const Ctxt& c : fc_.unpack(classified)) out.push_back(logits(c, classes));
    return out;
}

std::vector<std::vector<double>> LinformerForward::run_many(int classes) {
    times_.clear();
    return logits_many(classifier(pooler(encoder())), classes);
}

std::vector<double> LinformerForward::run(int classes) {
    if (samples() > 1) throw std::invalid_argument("run(): several samples were added, call run_many()");
    times_.clear();
    return logits(classifier(pooler(encoder())), classes);
}

int LinformerForward::argmax_softmax(const std::vector<double>& z, std::vector<double>* prob) {
    const double top = *std::max_element(z.begin(), z.end());
    std::vector<double> p(z.size());
    double sum = 0;
    for (size_t i = 0; i < z.size(); ++i) sum += (p[i] = std::exp(z[i] - top));
    for (double& v : p) v /= sum;
    const int best = (int)(std::max_element(p.begin(), p.end()) - p.begin());
    if (prob) *prob = std::move(p);
    return best;
}

}  // namespace flh
