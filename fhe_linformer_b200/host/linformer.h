// linformer.h -- the encrypted Linformer forward pass (encoder layer -> pooler -> classifier) driven through
// FHEController, i.e. the circuit of the reference's src/main.cpp:145-475, organised as stages so that tests and the
// bench can time and checkpoint each of them.  Reads the same text files, in the same layouts, as the reference.
#pragma once
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "FHEController.h"

namespace flh {

struct LinformerFiles {
    std::string weights;   // the reference's ../weights-20NG
    std::string input;     // ../input          (XE_0..31.txt, XF_0..31.txt: client-side Linformer projections)
    std::string tokens;    // <input_folder>    (input_0.txt .. input_{S-2}.txt: token embeddings)
};

struct StageTime {
    std::string name;
    double seconds;
};

class LinformerForward {
public:
    LinformerForward(FHEController& controller, LinformerFiles files, bool verbose = false);

    // decrypted slots of named intermediate ciphertexts are handed to `sink` (tests compare them with the slot simulator);
    // the reference prints a few of the same intermediates (main.cpp:198-199,227,329,369,420,448)
    void set_checkpoint_sink(std::function<void(const std::string&, const std::vector<double>&, int level)> sink) { sink_ = std::move(sink); }
    void set_token_limit(int n) { token_limit_ = n; }   // use only the first n token files (0 = all)
    // true (default): issue every operation main.cpp issues, including the ones whose results it never reads (queries of
    // rows 1.., W_O on the S-1 zero rows, the final unwrap of all rows).  false: skip those; the logits are identical.
    void set_dead_work(bool on) { dead_work_ = on; }
    // true: the server computes X_E = E[:, :S] X + E_b and X_F on the encrypted rows (fl_linear_wsum) from
    // <weights>/..._selfAttn_{E,F}_{weight,bias}.txt instead of reading the client's XE_i / XF_i uploads (M:159-167,
    // src/python/dimReduce.py:153-160) -- SURVEY.md F1
    void set_encrypted_projection(bool on) { encrypted_projection_ = on; }
    // true: the circuit of the reference's src/main_2.cpp (not built by its CMakeLists): attention for EVERY row (two halves of
    // up to 128 queries packed by matmulScores(vector)), 1/x fitted on [-1, 190000], W_O bias on every row, tanh scale 1/18,
    // plaintext class mask -- SURVEY.md F4
    void set_all_token_attention(bool on) { all_tokens_ = on; }
    // true: packed mode (BASELINE north star: BSGS diagonal ct x pt matmul behind the linear layers).  Same network, same weights,
    // same logits up to CKKS noise; the 32 projected rows of the attention block and the S rows from the first affine to the
    // second stay in the wrapped-expanded layout (128 rows per ciphertext), each 128 x 128 weight block (W_K, W_V, W_Q, W_O, the
    // eight FFN blocks) is ONE FHEController::packed_linear instead of a (x) + 7-step ladder per row, and the unwrap / container /
    // re-wrap round trips between them (F.cpp:1086-1205) disappear: ~0.8 k rotations per forward instead of ~21 k at S = 200.
    // Together with set_dead_work(false) only what the logits read is evaluated (first half of the rows, CLS column after W2).
    // Needs FHEController::generate_packed_keys().
    void set_packed(bool on) { packed_ = on; }
    // One more sample (its XE_/XF_ folder and its token folder; same weights, same number of rows) evaluated by the SAME calls:
    // every ciphertext of the forward then carries one element per sample (BASELINE config 5: a batch of samples,
    // ciphertext-parallel).  Packed mode only; the results come from run_many() / logits_many().
    void add_sample(const std::string& input_dir, const std::string& tokens_dir) { more_.push_back({input_dir, tokens_dir}); }
    int samples() const { return 1 + (int)more_.size(); }

    Ctxt encoder();                       // main.cpp:145-425
    Ctxt pooler(const Ctxt& encoded);     // main.cpp:427-451
    Ctxt classifier(const Ctxt& pooled);  // main.cpp:453-475
    std::vector<double> logits(const Ctxt& classified, int classes = 20);   // main.cpp:115-123: slot 128 i holds class i
    std::vector<double> run(int classes = 20);                               // the three stages + decrypt
    std::vector<std::vector<double>> logits_many(const Ctxt& classified, int classes = 20);   // one vector per sample
    std::vector<std::vector<double>> run_many(int classes = 20);
    static int argmax_softmax(const std::vector<double>& logits, std::vector<double>* prob = nullptr);   // main.cpp:125-142

    const std::vector<StageTime>& timings() const { return times_; }
    int tokens() const { return tokens_; }

private:
    struct Attention;   // K/V side of the CLS-only attention
    std::string w(const std::string& name) const { return files_.weights + "/" + name; }
    std::string layer(const std::string& name) const { return w("linformer_transformerLayers_transformer0_" + name); }
    double scalar(const std::string& path) const;
    void checkpoint(const std::string& name, const Ctxt& c);
    void lap(const std::string& name);
    std::vector<Ctxt> load_expanded(const std::string& dir, const std::string& stem, int count);
    std::vector<Ctxt> load_expanded(const std::vector<std::string>& dirs, const std::string& stem, int count);

    Ctxt attend_cls(const std::vector<Ctxt>& rows, const std::vector<Ctxt>& xe, const std::vector<Ctxt>& xf);
    Ctxt attend_cls_packed(const std::vector<Ctxt>& rows, const std::vector<Ctxt>& xe, const std::vector<Ctxt>& xf);
    std::vector<Ctxt> self_output(const Ctxt& cls_context, const std::vector<Ctxt>& rows);
    std::pair<Ctxt, Ctxt> affine_and_refresh(const std::vector<Ctxt>& rows, const std::string& which, bool refresh);
    std::vector<Ctxt> feed_forward(const Ctxt& half0, const Ctxt& half1, int rows);
    std::pair<Ctxt, Ctxt> feed_forward_packed(const Ctxt& half0, const Ctxt& half1, const std::vector<double>* cls_column_scale);
    Ctxt wrap_rows_packed(const std::vector<Ctxt>& rows, int first_position);

    FHEController& fc_;
    LinformerFiles files_;
    bool verbose_;
    int token_limit_ = 0;
    bool dead_work_ = true;
    bool encrypted_projection_ = false;
    bool all_tokens_ = false;
    bool packed_ = false;
    std::vector<std::pair<std::string, std::string>> more_;   // (input folder, token folder) of the samples after the first
    std::vector<Ctxt> attend_all(const std::vector<Ctxt>& rows, const std::vector<Ctxt>& xe, const std::vector<Ctxt>& xf);
    std::vector<Ctxt> project(const std::vector<Ctxt>& rows, const std::string& which);
    int tokens_ = 0;
    std::function<void(const std::string&, const std::vector<double>&, int)> sink_;
    std::vector<StageTime> times_;
    utils::time_point t0_;
};

}  // namespace flh
