// flhost.cpp -- extern "C" face of the host layer (see flhost.h).
#include "flhost.h"

#include <cstring>
#include <map>

#include "linformer.h"

struct flh_controller {
    FHEController fc;
};

namespace {
thread_local std::string g_err;

template <class F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    } catch (...) {
        g_err = "unknown error";
        return 2;
    }
}

Ctxt borrow_ct(FHEController& fc, fl_elem* h) {
    fl_elem* copy = nullptr;   // the veneer's handles own their element: work on a device copy
    if (fl_elem_clone(fc.native(), h, &copy)) throw std::runtime_error(fl_last_error());
    return fc.adopt(copy);
}
Ptxt borrow_pt(FHEController& fc, fl_elem* h) {
    if (!h) return nullptr;
    fl_elem* copy = nullptr;
    if (fl_elem_clone(fc.native(), h, &copy)) throw std::runtime_error(fl_last_error());
    return std::make_shared<lbcrypto::PlaintextImpl>(fc.native(), copy);
}
fl_elem* release(FHEController& fc, const Ctxt& c) {
    fl_elem* copy = nullptr;
    if (fl_elem_clone(fc.native(), c->handle(), &copy)) throw std::runtime_error(fl_last_error());
    return copy;
}
}  // namespace

extern "C" {

const char* flh_last_error(void) { return g_err.c_str(); }

flh_controller* flh_new(int device, unsigned long long key_seed) {
    auto* c = new flh_controller();
    c->fc.device = device;
    c->fc.key_seed = key_seed;
    return c;
}
void flh_free(flh_controller* c) {
    if (!c) return;
    c->fc.release_context();     // the handles this face hands out are device copies owned by the caller's context wrapper: none are left here
    delete c;
}
fl_ctx* flh_native(flh_controller* c) { return c->fc.native(); }
int flh_set_option(flh_controller* c, const char* name, double value) {
    return guarded([&] {
        const std::string n = name;
        FHEController& fc = c->fc;
        if (n == "cache_gb") { fc.cache_gb = value; if (fc.native() && value > 0 && fl_ctx_set_cache_bytes(fc.native(), (uint64_t)(value * 1073741824.0))) throw std::runtime_error(fl_last_error()); }
        else if (n == "auto_rotation_keys") fc.auto_rotation_keys = value != 0;
        else if (n == "packed_keys") { if (value != 0) fc.generate_packed_keys(); }
        else if (n == "batch_rows") fc.batch_rows = value != 0;
        else if (n == "hoist_ladders") fc.hoist_ladders = value != 0;
        else if (n == "max_rows_per_batch") fc.max_rows_per_batch = std::max(1, (int)value);
        else throw std::invalid_argument("flh_set_option: unknown option " + n);
    });
}
double flh_rotation_key_bytes(flh_controller* c) { return c->fc.native() ? c->fc.rotation_key_bytes() : 0.0; }

int flh_generate(flh_controller* c, int log_ring, const int* rotations, int n_rot, int bootstrap_slots, int serialize) {
    return guarded([&] {
        if (log_ring) setenv("FHE_LINFORMER_LOGN", std::to_string(log_ring).c_str(), 1);
        else unsetenv("FHE_LINFORMER_LOGN");
        c->fc.generate_context(serialize != 0, false);
        c->fc.generate_bootstrapping_and_rotation_keys(std::vector<int>(rotations, rotations + n_rot), bootstrap_slots, serialize != 0, "rotation_keys.txt");
    });
}
int flh_load(flh_controller* c, const char* rotation_file, int bootstrap_slots) {
    return guarded([&] {
        c->fc.load_context(false);
        c->fc.load_bootstrapping_and_rotation_keys(rotation_file, bootstrap_slots, false);
    });
}
int flh_info(flh_controller* c, int* circuit_depth, int* num_slots) {
    *circuit_depth = c->fc.circuit_depth;
    *num_slots = c->fc.num_slots;
    return 0;
}

int flh_forward(flh_controller* c, const char* weights_dir, const char* input_dir, const char* tokens_dir, int token_limit, int dead_work,
                int classes, double* logits, flh_checkpoint_fn sink, void* user, char* timing_names, int names_cap, double* timing_seconds,
                int* n_timings, int* tokens) {
    return guarded([&] {
        flh::LinformerForward fwd(c->fc, {weights_dir, input_dir, tokens_dir}, false);
        fwd.set_token_limit(token_limit);
        fwd.set_dead_work((dead_work & 1) != 0);
        fwd.set_encrypted_projection((dead_work & 2) != 0);
        fwd.set_all_token_attention((dead_work & 4) != 0);
        fwd.set_packed((dead_work & 8) != 0);
        if (sink) fwd.set_checkpoint_sink([&](const std::string& name, const std::vector<double>& v, int level) { sink(name.c_str(), v.data(), (int)v.size(), level, user); });
        const std::vector<double> z = fwd.run(classes);
        std::memcpy(logits, z.data(), sizeof(double) * z.size());
        if (tokens) *tokens = fwd.tokens();
        if (n_timings) {
            std::string names;
            int n = 0;
            for (const auto& t : fwd.timings()) {
                if (n >= *n_timings || (int)(names.size() + t.name.size() + 2) > names_cap) break;
                names += t.name + "\n";
                timing_seconds[n++] = t.seconds;
            }
            *n_timings = n;
            if (timing_names && names_cap > 0) std::strncpy(timing_names, names.c_str(), (size_t)names_cap);
        }
    });
}

int flh_forward_many(flh_controller* c, const char* weights_dir, const char* const* input_dirs, const char* const* tokens_dirs, int samples,
                     int token_limit, int flags, int classes, double* logits, int* tokens) {
    return guarded([&] {
        if (samples < 1) throw std::invalid_argument("flh_forward_many: no samples");
        flh::LinformerForward fwd(c->fc, {weights_dir, input_dirs[0], tokens_dirs[0]}, false);
        fwd.set_token_limit(token_limit);
        fwd.set_dead_work((flags & 1) != 0);
        fwd.set_encrypted_projection((flags & 2) != 0);
        fwd.set_all_token_attention((flags & 4) != 0);
        fwd.set_packed((flags & 8) != 0);
        for (int m = 1; m < samples; ++m) fwd.add_sample(input_dirs[m], tokens_dirs[m]);
        const std::vector<std::vector<double>> z = fwd.run_many(classes);
        for (int m = 0; m < samples; ++m) std::memcpy(logits + (size_t)m * classes, z[(size_t)m].data(), sizeof(double) * (size_t)classes);
        if (tokens) *tokens = fwd.tokens();
    });
}

int flh_invoke(flh_controller* c, const char* method, fl_elem* const* cts, int n_cts, fl_elem* const* pts, int n_pts, const int* ints, int n_ints,
               const double* reals, int n_reals, fl_elem** out, int out_cap, int* n_out) {
    return guarded([&] {
        FHEController& fc = c->fc;
        const std::string m = method;
        std::vector<Ctxt> in;
        for (int i = 0; i < n_cts; ++i) in.push_back(borrow_ct(fc, cts[i]));
        std::vector<Ptxt> pin;
        for (int i = 0; i < n_pts; ++i) pin.push_back(borrow_pt(fc, pts[i]));
        auto I = [&](int k) { if (k >= n_ints) throw std::invalid_argument(m + ": missing integer argument"); return ints[k]; };
        auto R = [&](int k) { if (k >= n_reals) throw std::invalid_argument(m + ": missing real argument"); return reals[k]; };
        auto P = [&](int k) -> Ptxt { if (k >= n_pts) throw std::invalid_argument(m + ": missing plaintext argument"); return pin[k]; };
        auto rows_but_last = [&] { return std::vector<Ctxt>(in.begin(), in.end() - 1); };
        std::vector<Ctxt> res;
        if (m == "rotsum") res = {fc.rotsum(in.at(0), I(0), I(1))};
        else if (m == "rotsum_padded") res = {fc.rotsum_padded(in.at(0), I(0))};
        else if (m == "repeat") res = {n_ints > 1 ? fc.repeat(in.at(0), I(0), I(1)) : fc.repeat(in.at(0), I(0))};
        else if (m == "rotate") res = {fc.rotate(in.at(0), I(0))};
        else if (m == "add") res = {n_pts ? fc.add(in.at(0), P(0)) : (in.size() == 2 ? fc.add(in[0], in[1]) : fc.add(in))};
        else if (m == "mult") res = {n_pts ? fc.mult(in.at(0), P(0)) : (n_reals ? fc.mult(in.at(0), R(0)) : fc.mult(in.at(0), in.at(1)))};
        else if (m == "bootstrap") res = {n_ints ? fc.bootstrap(in.at(0), I(0)) : fc.bootstrap(in.at(0))};
        else if (m == "relu") res = {fc.relu(in.at(0), R(0))};
        else if (m == "matmulRE") {
            if (n_pts) res = n_ints ? fc.matmulRE(in, P(0), P(1), I(0), I(1)) : fc.matmulRE(in, P(0), P(1));
            else res = fc.matmulRE(rows_but_last(), in.back(), I(0), I(1));
        } else if (m == "matmulRElarge") {
            res = fc.matmulRElarge(in, std::vector<Ptxt>(pin.begin(), pin.begin() + 4), P(4), n_reals ? R(0) : 1.0);
        } else if (m == "matmulCR") {
            res = n_pts ? fc.matmulCR(in, P(0), P(1)) : fc.matmulCR(rows_but_last(), in.back());
        } else if (m == "matmulCR_128") res = fc.matmulCR_128(rows_but_last(), in.back());
        else if (m == "matmulCRlarge") {
            std::vector<std::vector<Ctxt>> quads;
            for (size_t i = 0; i + 3 < in.size(); i += 4) quads.push_back({in[i], in[i + 1], in[i + 2], in[i + 3]});
            res = fc.matmulCRlarge(quads, std::vector<Ptxt>(pin.begin(), pin.begin() + 4), P(4));
        } else if (m == "matmulScores") {
            res = {in.size() == 2 ? fc.matmulScores(in[0], in[1]) : fc.matmulScores(rows_but_last(), in.back())};
        } else if (m == "matmulScoresVec") res = {fc.matmulScores(rows_but_last(), in.back())};
        else if (m == "wrapUpRepeated") res = {fc.wrapUpRepeated(in)};
        else if (m == "wrapUpExpanded") res = {fc.wrapUpExpanded(in)};
        else if (m == "unwrapExpanded") res = fc.unwrapExpanded(in.at(0), I(0));
        else if (m == "unwrapScoresExpanded") res = fc.unwrapScoresExpanded(in.at(0), I(0));
        else if (m == "unwrap_512_in_4_128") res = fc.unwrap_512_in_4_128(in.at(0), I(0));
        else if (m == "unwrapRepeatedLarge") {
            for (auto& quad : fc.unwrapRepeatedLarge(in, I(0))) res.insert(res.end(), quad.begin(), quad.end());
        } else if (m == "generate_containers") res = fc.generate_containers(in, n_pts ? P(0) : nullptr);
        else if (m == "wrap_containers") res = {fc.wrap_containers(in, I(0))};
        else if (m == "mask_block") res = {fc.mask_block(in.at(0), I(0), I(1), n_reals ? R(0) : 1.0)};
        else if (m == "mask_heads") res = {fc.mask_heads(in.at(0), n_reals ? R(0) : 1.0)};
        else if (m == "mask_heads_128") res = {fc.mask_heads_128(in.at(0), n_reals ? R(0) : 1.0)};
        else if (m == "mask_mod_n") res = {n_ints > 1 ? fc.mask_mod_n(in.at(0), I(0), I(1), I(2)) : fc.mask_mod_n(in.at(0), I(0))};
        else if (m == "mask_first_n") res = {fc.mask_first_n(in.at(0), I(0), n_reals ? R(0) : 1.0)};
        else if (m == "eval_exp") res = {fc.eval_exp(in.at(0), I(0))};
        else if (m == "eval_inverse") res = {fc.eval_inverse(in.at(0), R(0), R(1))};
        else if (m == "eval_inverse_naive") res = {fc.eval_inverse_naive(in.at(0), R(0), R(1))};
        else if (m == "eval_inverse_naive_2") res = {fc.eval_inverse_naive_2(in.at(0), R(0), R(1), R(2))};
        else if (m == "eval_gelu_function") res = {fc.eval_gelu_function(in.at(0), R(0), R(1), R(2), I(0))};
        else if (m == "eval_tanh_function") res = {fc.eval_tanh_function(in.at(0), R(0), R(1), R(2), I(0))};
        else if (m == "project_rows") {
            std::vector<std::vector<double>> w((size_t)I(0), std::vector<double>(in.size()));
            for (size_t o = 0; o < w.size(); ++o)
                for (size_t t = 0; t < in.size(); ++t) w[o][t] = R((int)(o * in.size() + t));
            res = fc.project_rows(in, w, {});
        } else if (m == "slicing") res = fc.slicing(in, I(0), I(1));
        else throw std::invalid_argument("flh_invoke: unknown method " + m);
        if ((int)res.size() > out_cap) throw std::runtime_error(m + ": result count exceeds the output capacity");
        for (size_t i = 0; i < res.size(); ++i) out[i] = release(fc, res[i]);
        *n_out = (int)res.size();
    });
}

}  // extern "C"
