// FHEController.h -- the reference's controller interface (src/FHEController.h:22-162: same public methods,
// fields, default arguments and leaked using-directives main.cpp depends on), re-backed by the B200 engine
// through include/fl_ckks.h instead of OpenFHE.  main.cpp compiles against this header unmodified.
#ifndef FLB200_FHECONTROLLER_H
#define FLB200_FHECONTROLLER_H

#include <algorithm>
#include <chrono>
#include <cmath>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <map>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "Utils.h"
#include "openfhe_compat.h"

using namespace lbcrypto;
using namespace std;
using namespace std::chrono;

using namespace utils;

using Ptxt = Plaintext;
using Ctxt = Ciphertext<DCRTPoly>;

class FHEController {
public:
    int circuit_depth;
    int num_slots;

    FHEController() {}
    ~FHEController();

    /* context generating / loading */
    void generate_context(bool serialize = false, bool secure = false);
    void generate_context(int log_ring, int log_scale, int log_primes, int digits_hks, int cts_levels, int stc_levels, int relu_deg, bool serialize = false);
    void load_context(bool verbose = true);

    /* bootstrapping and rotation keys */
    void generate_bootstrapping_keys(int bootstrap_slots);
    void generate_rotation_keys(vector<int> rotations, bool serialize = false, string filename = "");
    void generate_bootstrapping_and_rotation_keys(vector<int> rotations, int bootstrap_slots, bool serialize, const string& filename);
    void load_bootstrapping_and_rotation_keys(const string& filename, int bootstrap_slots, bool verbose);
    void load_rotation_keys(const string& filename, bool verbose);
    void clear_bootstrapping_and_rotation_keys(int bootstrap_num_slots);
    void clear_rotation_keys();
    void clear_context(int bootstrapping_key_slots);

    /* CKKS encoding / decoding / encryption / decryption */
    Ptxt encode(const vector<double>& vec, int level, int plaintext_num_slots);
    Ptxt encode(double val, int level, int plaintext_num_slots);
    Ctxt encrypt(const vector<double>& vec, int level = 0, int plaintext_num_slots = 0);
    Ctxt encrypt_ptxt(const Ptxt& p);
    Ptxt decrypt(const Ctxt& c);
    vector<double> decrypt_tovector(const Ctxt& c, int slots);

    /* homomorphic operations */
    Ctxt add(const Ctxt& c1, const Ctxt& c2);
    Ctxt add(const Ctxt& c1, const Ptxt& c2);
    Ctxt add(vector<Ctxt> c);
    Ctxt mult(const Ctxt& c1, const Ctxt& c2);
    Ctxt mult(const Ctxt& c, double d);
    Ctxt mult(const Ctxt& c, const Ptxt& p);
    Ctxt rotate(const Ctxt& c, int index);
    Ctxt bootstrap(const Ctxt& c, bool timing = false);
    Ctxt bootstrap(const Ctxt& c, int precision, bool timing = false);
    Ctxt relu(const Ctxt& c, double scale, bool timing = false);

    /* I/O */
    Ctxt read_input(const string& filename, double scale = 1);
    Ctxt read_repeated_input(const string& filename, double scale = 1);
    Ctxt read_expanded_input(const string& filename, double scale = 1);
    Ptxt read_plain_input(const string& filename, int level = 0, double scale = 1);
    vector<Ptxt> read_plain_256_input(const string& filename, int level = 0, double scale = 1);
    Ptxt read_plain_repeated_input(const string& filename, int level = 0, double scale = 1);
    Ptxt read_plain_repeated_512_input(const string& filename, int level = 0, double scale = 1);
    Ptxt read_plain_expanded_input(const string& filename, int level = 0, double scale = 1);
    Ptxt read_plain_expanded_input(const string& filename, int level, double scale, int num_inputs);

    void print(const Ctxt& c, int slots = 0, string prefix = "");
    void print_padded(const Ctxt& c, int slots = 0, int padding = 1, string prefix = "");
    void print_expanded(const Ctxt& c, int slots = 0, int expansion_factor = 1, string prefix = "");
    void print_min_max(const Ctxt& c);

    Ctxt rotsum(const Ctxt& in, int slots, int padding);
    Ctxt rotsum_padded(const Ctxt& in, int slots);
    Ctxt repeat(const Ctxt& in, int slots);
    Ctxt repeat(const Ctxt& in, int slots, int padding);

    vector<Ctxt> matmulRE(vector<Ctxt> rows, const Ptxt& weight, const Ptxt& bias);
    vector<Ctxt> matmulRE(vector<Ctxt> rows, const Ptxt& weight, const Ptxt& bias, int row_size, int padding);
    vector<Ctxt> matmulRE(vector<Ctxt> rows, const Ctxt& weight, int row_size, int padding);
    vector<Ctxt> matmulRElarge(vector<Ctxt>& rows, const vector<Ptxt>& weight, const Ptxt& bias, double mask_value = 1);
    vector<Ctxt> matmulCR(vector<Ctxt> rows, const Ptxt& weight, const Ptxt& bias);
    vector<Ctxt> matmulCR(vector<Ctxt> rows, const Ctxt& matrix);
    vector<Ctxt> matmulCR_128(vector<Ctxt> rows, const Ctxt& matrix);
    Ctxt matmulCR_128(Ctxt row, const Ctxt& matrix);
    vector<Ctxt> matmulCRlarge(vector<vector<Ctxt>> rows, vector<Ptxt> weights, const Ptxt& bias);

    Ctxt matmulScores(vector<Ctxt> queries, const Ctxt& key);
    Ctxt matmulScores(Ctxt query, const Ctxt& key);

    Ctxt wrapUpRepeated(vector<Ctxt> vectors);
    Ctxt wrapUpExpanded(vector<Ctxt> vectors);

    vector<Ctxt> unwrapExpanded(Ctxt c, int inputs_num);
    vector<vector<Ctxt>> unwrapRepeatedLarge(vector<Ctxt> c, int input_number);
    vector<Ctxt> unwrapScoresExpanded(Ctxt c, int inputs_num);
    vector<Ctxt> unwrap_512_in_4_128(const Ctxt& c, int index);

    vector<Ctxt> generate_containers(vector<Ctxt> inputs, const Ptxt& bias);
    Ctxt wrap_containers(vector<Ctxt> inputs, int inputs_number);

    Ctxt mask_block(const Ctxt& c, int from, int to, double mask_value = 1);
    Ctxt mask_heads(const Ctxt& c, double mask_value = 1);
    Ctxt mask_heads_128(const Ctxt& c, double mask_value = 1);
    Ctxt mask_mod_n(const Ctxt& c, int n);
    Ctxt mask_mod_n(const Ctxt& c, int n, int padding, int max_slots);
    Ctxt mask_first_n(const Ctxt& c, int n, double mask_value = 1);

    Ctxt eval_exp(const Ctxt& c, int inputs_number);
    Ctxt eval_inverse(const Ctxt& c, double min, double max);
    Ctxt eval_inverse_naive(const Ctxt& c, double min, double max);
    Ctxt eval_inverse_naive_2(const Ctxt& c, double min, double max, double mult);
    Ctxt eval_gelu_function(const Ctxt& c, double min, double max, double mult, int degree);
    Ctxt eval_tanh_function(const Ctxt& c, double min, double max, double mult, int degree);

    vector<Ctxt> slicing(vector<Ctxt>& arr, int X, int Y);

    void save(Ctxt v, string filename);
    void save(vector<Ctxt> v, string filename);
    vector<Ctxt> load_vector(string filename);
    Ctxt load_ciphertext(string filename);

    int relu_degree = 119;
    string parameters_folder = "keys";

    /* ---- additions of the B200 backend (not in the reference interface) ---- */
    int device = 0;                 // CUDA device of this controller (one context per GPU)
    unsigned long long key_seed = 20261018ULL;
    fl_ctx* native() const { return ctx_; }
    Ctxt chebyshev(const std::function<double(double)>& f, const Ctxt& c, double a, double b, int degree);
    Ctxt ladder(const Ctxt& in, int slots, int stride);   // shared body of rotsum / rotsum_padded / repeat
    // out[o] = sum_t weights[o][t] * rows[t] + bias[o] (bias: one plaintext per output, may be empty): the Linformer E / F
    // projection evaluated on the row ciphertexts instead of by the client
    vector<Ctxt> project_rows(const vector<Ctxt>& rows, const vector<vector<double>>& weights, const vector<Ptxt>& bias);
    Ctxt adopt(fl_elem* e) const;                          // take ownership of a raw C-ABI handle
    // row batching (FHEController.cpp "row batching"): independent rows share kernel launches
    bool hoist_ladders = true;      // generate the extra 3 * stride * 4^i keys that let ladders take two steps per key switch
    bool batch_rows = true;
    int max_rows_per_batch = 64;    // measured: same speed as 256 (1.698 vs 1.693 s at S = 256) with 50 GB instead of 85 GB of cached blocks
    Ctxt pack(const vector<Ctxt>& rows) const;
    vector<Ctxt> unpack(const Ctxt& packed) const;
    vector<Ctxt> per_row(const vector<Ctxt>& rows, const std::function<Ctxt(const Ctxt&)>& recipe) const;
    vector<Ctxt> settle_rows(const vector<Ctxt>& rows) const;
    Ctxt shifted_sum(vector<Ctxt> items, int stride);          // sum_i rot(items[i], stride * i), tree of batched rotations
    vector<Ctxt> all_shifts(const Ctxt& c, int count);          // rot(c, t), t < count, by batched doubling   // pending FLEXIBLEAUTO rescales of many rows, as one batch

private:
    void create(int log_ring, int depth, int digits, int first_bits, int scale_bits);
    void serialize_context();
    Ctxt wrap(fl_elem* e) const { return std::make_shared<CiphertextImpl<DCRTPoly>>(ctx_, e); }
    Ptxt wrap_pt(fl_elem* e) const { return std::make_shared<PlaintextImpl>(ctx_, e); }
    Ptxt mask_plain(int kind, int a, int b, double value, int level);
    string key_path(const string& name) const;

    fl_ctx* ctx_ = nullptr;
    fl_params params_{};
    vector<uint32_t> level_budget = {4, 4};
    map<std::tuple<int, int, int, double, int>, Ptxt> mask_cache_;
};

#endif  // FLB200_FHECONTROLLER_H
