// FHEController.h -- the controller interface the reference's pipeline is written against (src/FHEController.h:22-162),
// re-backed by the B200 CKKS engine through include/fl_ckks.h instead of OpenFHE.
//
// Drop-in contract: every public method the reference defines exists here with a type-identical signature (same names,
// argument order, default arguments), the public data members main.cpp reads (circuit_depth, num_slots, relu_degree,
// parameters_folder) are present, and the using-directives / aliases the reference's header leaks into its includers are
// leaked the same way -- so src/main.cpp compiles against this file unmodified (`make reference-main-check`).
// Declarations are grouped by what they do on the engine, not in the reference's order; `Rows` and `RowGroups` are aliases.
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

using Ptxt = Plaintext;              // encoded plaintext on the device, or decoded values on the host
using Ctxt = Ciphertext<DCRTPoly>;   // ciphertext handle; may carry a batch of ciphertexts with identical metadata
using Rows = vector<Ctxt>;           // one ciphertext per token / row
using RowGroups = vector<Rows>;

class FHEController {
public:
    FHEController() {}
    ~FHEController();
    // Destroys the context (keys, caches, device memory).  The destructor leaves it alive because ciphertexts handed out may
    // outlive the controller (main.cpp keeps globals); call this only when none are left.
    void release_context();

    // ---- data members the pipeline reads ------------------------------------------------------------------------------
    int circuit_depth;                       // multiplicative depth chosen at generation (27) / recomputed at load (26)
    int num_slots;                           // 2^14
    int relu_degree = 119;
    string parameters_folder = "keys";

    // ---- context and key material (F.cpp:3-343) ------------------------------------------------------------------------
    void generate_context(bool serialize = false, bool secure = false);
    void generate_context(int log_ring, int log_scale, int log_primes, int digits_hks, int cts_levels, int stc_levels, int relu_deg, bool serialize = false);
    void load_context(bool verbose = true);
    void clear_context(int bootstrapping_key_slots);
    void generate_rotation_keys(vector<int> rotations, bool serialize = false, string filename = "");
    void generate_bootstrapping_keys(int bootstrap_slots);
    void generate_bootstrapping_and_rotation_keys(vector<int> rotations, int bootstrap_slots, bool serialize, const string& filename);
    void load_rotation_keys(const string& filename, bool verbose);
    void load_bootstrapping_and_rotation_keys(const string& filename, int bootstrap_slots, bool verbose);
    void clear_rotation_keys();
    void clear_bootstrapping_and_rotation_keys(int bootstrap_num_slots);

    // ---- plaintexts, encryption, decryption (F.cpp:348-404) ------------------------------------------------------------
    Ptxt encode(double val, int level, int plaintext_num_slots);
    Ptxt encode(const vector<double>& vec, int level, int plaintext_num_slots);
    Ctxt encrypt_ptxt(const Ptxt& p);
    Ctxt encrypt(const vector<double>& vec, int level = 0, int plaintext_num_slots = 0);
    Ptxt decrypt(const Ctxt& c);
    vector<double> decrypt_tovector(const Ctxt& c, int slots);

    // ---- text files -> plaintexts / ciphertexts in the plain, repeated or expanded layout (F.cpp:501-698) ---------------
    Ptxt read_plain_input(const string& filename, int level = 0, double scale = 1);
    Ptxt read_plain_repeated_input(const string& filename, int level = 0, double scale = 1);
    Ptxt read_plain_repeated_512_input(const string& filename, int level = 0, double scale = 1);
    Ptxt read_plain_expanded_input(const string& filename, int level = 0, double scale = 1);
    Ptxt read_plain_expanded_input(const string& filename, int level, double scale, int num_inputs);
    vector<Ptxt> read_plain_256_input(const string& filename, int level = 0, double scale = 1);
    Ctxt read_input(const string& filename, double scale = 1);
    Ctxt read_repeated_input(const string& filename, double scale = 1);
    Ctxt read_expanded_input(const string& filename, double scale = 1);

    // ---- primitive homomorphic operations (F.cpp:409-469) ---------------------------------------------------------------
    Ctxt add(const Ctxt& c1, const Ctxt& c2);
    Ctxt add(const Ctxt& c1, const Ptxt& c2);
    Ctxt add(Rows c);
    Ctxt mult(const Ctxt& c, const Ptxt& p);
    Ctxt mult(const Ctxt& c, double d);
    Ctxt mult(const Ctxt& c1, const Ctxt& c2);
    Ctxt rotate(const Ctxt& c, int index);
    Ctxt bootstrap(const Ctxt& c, bool timing = false);
    Ctxt bootstrap(const Ctxt& c, int precision, bool timing = false);

    // ---- rotate-and-add ladders (F.cpp:829-867) -------------------------------------------------------------------------
    Ctxt rotsum(const Ctxt& in, int slots, int padding);
    Ctxt rotsum_padded(const Ctxt& in, int slots);
    Ctxt repeat(const Ctxt& in, int slots);
    Ctxt repeat(const Ctxt& in, int slots, int padding);

    // ---- masks: multiply by a cached 0 / value plaintext (F.cpp:1207-1286) ----------------------------------------------
    Ctxt mask_first_n(const Ctxt& c, int n, double mask_value = 1);
    Ctxt mask_block(const Ctxt& c, int from, int to, double mask_value = 1);
    Ctxt mask_mod_n(const Ctxt& c, int n);
    Ctxt mask_mod_n(const Ctxt& c, int n, int padding, int max_slots);
    Ctxt mask_heads(const Ctxt& c, double mask_value = 1);
    Ctxt mask_heads_128(const Ctxt& c, double mask_value = 1);

    // ---- packed matrix products, one row ciphertext per token (F.cpp:869-1058) -----------------------------------------
    Rows matmulRE(Rows rows, const Ptxt& weight, const Ptxt& bias);
    Rows matmulRE(Rows rows, const Ptxt& weight, const Ptxt& bias, int row_size, int padding);
    Rows matmulRE(Rows rows, const Ctxt& weight, int row_size, int padding);
    Rows matmulRElarge(Rows& rows, const vector<Ptxt>& weight, const Ptxt& bias, double mask_value = 1);
    Rows matmulCR(Rows rows, const Ptxt& weight, const Ptxt& bias);
    Rows matmulCR(Rows rows, const Ctxt& matrix);
    Rows matmulCR_128(Rows rows, const Ctxt& matrix);
    Ctxt matmulCR_128(Ctxt row, const Ctxt& matrix);
    Rows matmulCRlarge(RowGroups rows, vector<Ptxt> weights, const Ptxt& bias);
    Ctxt matmulScores(Ctxt query, const Ctxt& key);
    Ctxt matmulScores(Rows queries, const Ctxt& key);

    // ---- layout conversions between "one vector per ciphertext" and "many vectors per ciphertext" (F.cpp:1060-1205, 1338) -
    Ctxt wrapUpRepeated(Rows vectors);
    Ctxt wrapUpExpanded(Rows vectors);
    Ctxt wrap_containers(Rows inputs, int inputs_number);
    Rows generate_containers(Rows inputs, const Ptxt& bias);
    Rows unwrapExpanded(Ctxt c, int inputs_num);
    Rows unwrapScoresExpanded(Ctxt c, int inputs_num);
    Rows unwrap_512_in_4_128(const Ctxt& c, int index);
    RowGroups unwrapRepeatedLarge(Rows c, int input_number);
    Rows slicing(Rows& arr, int X, int Y);

    // ---- polynomial approximations of the activations (F.cpp:471-495, 1289-1336) ----------------------------------------
    Ctxt eval_exp(const Ctxt& c, int inputs_number);
    // inputs_number <= 0: the bare ((Taylor_6)^8) without the "-1 outside the first inputs_number x inputs_number region" plaintext
    // (packed attention keeps meaningful scores in every slot)
    Ctxt eval_inverse(const Ctxt& c, double min, double max);
    Ctxt eval_inverse_naive(const Ctxt& c, double min, double max);
    Ctxt eval_inverse_naive_2(const Ctxt& c, double min, double max, double mult);
    Ctxt eval_gelu_function(const Ctxt& c, double min, double max, double mult, int degree);
    Ctxt eval_tanh_function(const Ctxt& c, double min, double max, double mult, int degree);
    Ctxt relu(const Ctxt& c, double scale, bool timing = false);

    // ---- ciphertext files and debug printing (secret-key decrypts, as in the reference) (F.cpp:700-826, 1360-1394) -------
    void save(Ctxt v, string filename);
    void save(Rows v, string filename);
    Ctxt load_ciphertext(string filename);
    Rows load_vector(string filename);
    void print(const Ctxt& c, int slots = 0, string prefix = "");
    void print_padded(const Ctxt& c, int slots = 0, int padding = 1, string prefix = "");
    void print_expanded(const Ctxt& c, int slots = 0, int expansion_factor = 1, string prefix = "");
    void print_min_max(const Ctxt& c);

    // =====================================================================================================================
    // Additions of the B200 backend (not part of the reference interface)
    // =====================================================================================================================
    int device = 0;                               // CUDA device of this controller: one context per GPU
    unsigned long long key_seed = 0;              // 0: operating-system randomness (production); non-zero: reproducible TEST keys
    double cache_gb = 0;                          // device block-cache cap of this controller's context in GB (0: engine default, 96)
    bool auto_rotation_keys = false;              // false: rotate() on an index without a key fails, as OpenFHE's EvalRotate does
    bool batch_rows = true;                       // independent rows share kernel launches (FHEController.cpp "row batching")
    bool hoist_ladders = true;                    // extra rotation keys (fl_rotsum_rotations): ladders take up to four doubling steps per hoisted key switch
    int max_rows_per_batch = 64;                  // same speed as 256 (1.698 vs 1.693 s at S = 256) with 50 GB instead of 85 GB cached
    fl_ctx* native() const { return ctx_; }
    Ctxt adopt(fl_elem* e) const;                 // take ownership of a raw C-ABI handle
    Ctxt chebyshev(const std::function<double(double)>& f, const Ctxt& c, double a, double b, int degree);   // EvalChebyshevFunction
    Ctxt ladder(const Ctxt& in, int slots, int stride);              // shared body of rotsum / rotsum_padded / repeat
    Ctxt pack(const Rows& rows) const;                               // rows of identical level / degree / scale -> one batched operand
    Rows unpack(const Ctxt& packed) const;                           // zero-copy views of a batched operand
    Rows unpack(const Ctxt& packed, int group) const;                // ... in groups of `group` consecutive elements
    Rows per_row(const Rows& rows, const std::function<Ctxt(const Ctxt&)>& recipe) const;   // run a per-row recipe on batches
    Rows settle_rows(const Rows& rows) const;                        // pending FLEXIBLEAUTO rescales of many rows, as one batch
    Rows encrypt_many(const vector<Ptxt>& plaintexts);                // Encrypt of many plaintexts of one level as one batched call
    Rows read_expanded_inputs(const vector<string>& filenames, double scale = 1);   // read_expanded_input for a list of files, encrypted together
    // file t of M samples as ONE batched ciphertext per t (a forward over M samples per call, LinformerForward::add_sample)
    Rows read_expanded_inputs_many(const vector<vector<string>>& files_per_sample, double scale = 1);
    Ctxt shifted_sum(Rows items, int stride);                        // sum_i rot(items[i], stride * i) as a tree of batched rotations
    Rows all_shifts(const Ctxt& c, int count);                       // rot(c, t), t < count, by batched doubling
    // out[o] = sum_t weights[o][t] * rows[t] (+ bias[o]): the Linformer E / F projection on the row ciphertexts
    Rows project_rows(const Rows& rows, const vector<vector<double>>& weights, const vector<Ptxt>& bias);

    // ---- packed linear layers (BASELINE north star: BSGS diagonal ct x pt matrix product behind the Q/K/V/FFN linears) ----
    // x holds up to 128 vectors in the wrapped-expanded layout (slot 128 j + t = element j of vector t, F.cpp:1070-1084);
    // the result holds y_t = x_t W for every t in the same layout: ONE baby-step/giant-step transform over the 128 diagonals
    // 128 k (fl_lt_apply, ~22 hoisted rotations) instead of one (x) + 7-step ladder per vector (F.cpp:869-883, 982-996).
    // weight(j, i) = W[j][i] with y_i = sum_j x_j W[j][i]; `name` keys the cached plan (plaintext diagonals follow x's level).
    // column_scale (optional, 128 values): output column t (the vector of position t) is additionally scaled by column_scale[t] --
    // a slot-wise factor of the wrapped-expanded layout folded into the diagonals (e.g. the position-indexed second affine, or a
    // column mask); it belongs to the plan, so give such a plan its own name.
    Ctxt packed_linear(const Ctxt& x, const string& name, const std::function<double(int, int)>& weight, double scale = 1.0,
                       const vector<double>* column_scale = nullptr);
    void generate_packed_keys();                                     // rotation keys of the packed transforms (before any packed forward)
    vector<int> derived_rotations(const vector<int>& listed) const;  // extra indices the batched / hoisted recipes use for a key list
    double rotation_key_bytes() const;                               // device memory held by automorphism keys

private:
    void require_rotation_key(int index);
    void create(int log_ring, int depth, int digits, int first_bits, int scale_bits);
    void serialize_context();
    string key_path(const string& name) const;
    Ptxt mask_plain(int kind, int a, int b, double value, int level);
    Ctxt wrap(fl_elem* e) const { return std::make_shared<CiphertextImpl<DCRTPoly>>(ctx_, e); }
    Ptxt wrap_pt(fl_elem* e) const { return std::make_shared<PlaintextImpl>(ctx_, e); }

    fl_ctx* ctx_ = nullptr;
    fl_params params_{};
    vector<uint32_t> level_budget = {4, 4};
    map<std::tuple<int, int, int, double, int>, Ptxt> mask_cache_;
    map<string, fl_lt*> packed_;                                     // BSGS plans of the packed linear layers, by weight name
};

#endif  // FLB200_FHECONTROLLER_H
