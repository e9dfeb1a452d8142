// fhe_linformer_main.cpp -- command line of the encrypted classifier on the B200 engine.  Same two modes as the
// reference binary (src/main.cpp:40-143): `--generate_keys [--secure]` writes the key files, anything else loads them and
// evaluates encoder -> pooler -> classifier on the text files of one sample.  Directory layout defaults to the
// reference's (../keys, ../weights-20NG, ../input, ../checkpoint relative to the working directory).
#include <cstdlib>
#include <iostream>

#include "linformer.h"

namespace {
std::string arg_value(int argc, char** argv, const std::string& flag, const std::string& fallback) {
    for (int i = 1; i + 1 < argc; ++i)
        if (flag == argv[i]) return argv[i + 1];
    return fallback;
}
bool has_flag(int argc, char** argv, const std::string& flag) {
    for (int i = 1; i < argc; ++i)
        if (flag == argv[i]) return true;
    return false;
}
}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) {
        std::cout << "Encrypted Linformer text classifier (B200 CKKS engine).\n\nUsage: fhe_linformer [--generate_keys [--secure]] [--verbose]\n"
                     "       [--root DIR] [--tokens DIR] [--lean] [--packed] [--resume]\n\n--generate_keys  create context, key pair, relinearisation, rotation and bootstrapping keys under <root>/keys\n"
                     "--verbose        per-stage timings and decrypted intermediates\n--root DIR       parent of keys/ weights-20NG/ input/ checkpoint/ (default ..)\n"
                     "--tokens DIR     folder with input_<i>.txt token embeddings (default <root>/tokens)\n--lean           skip operations whose results the circuit never reads\n"
                     "--encrypted-projection  compute the Linformer E/F projections on the server from the encrypted rows\n"
                     "--all-tokens     attention for every row (the circuit of the reference's main_2.cpp)\n"
                     "--packed         every linear layer (attention and feed-forward) on 128 rows per ciphertext through BSGS diagonal products\n"
                     "                 (same logits, ~6x faster; with --lean ~8x);\n"
                     "                 give it to --generate_keys as well so that the keys of the packed transforms are written\n"
                     "--resume         start from <root>/checkpoint/encodered.bin instead of running the encoder\n";
        return 0;
    }
    const std::string root = arg_value(argc, argv, "--root", "..");
    setenv("FHE_LINFORMER_ROOT", root.c_str(), 1);
    FHEController controller;
    if (std::string(argv[1]) == "--generate_keys") {
        if (std::system(("mkdir -p " + root + "/keys").c_str()) != 0) return 1;
        controller.generate_context(true, has_flag(argc, argv, "--secure"));
        // every index the circuit rotates by (the reference's own list misses -8, -16, -512)
        std::vector<int> rotations;
        for (int i = 0; i <= 13; ++i) rotations.push_back(1 << i);
        for (int i = 0; i <= 6; ++i) rotations.push_back(-(1 << i));
        rotations.push_back(-512);
        if (has_flag(argc, argv, "--packed")) {
            controller.generate_bootstrapping_keys(16384);
            controller.generate_packed_keys();                                  // stored with the other automorphism keys below
            controller.generate_rotation_keys(rotations, true, "rotation_keys.txt");
        } else {
            controller.generate_bootstrapping_and_rotation_keys(rotations, 16384, true, "rotation_keys.txt");
        }
        return 0;
    }
    const bool verbose = has_flag(argc, argv, "--verbose");
    controller.load_context(false);
    controller.load_bootstrapping_and_rotation_keys("rotation_keys.txt", 16384, false);
    if (std::system(("mkdir -p " + root + "/checkpoint").c_str()) != 0) return 1;
    if (verbose) std::cout << "\nSERVER-SIDE\nThe evaluation of the circuit started." << std::endl;
    const auto start = utils::start_time();
    flh::LinformerForward forward(controller, {root + "/weights-20NG", root + "/input", arg_value(argc, argv, "--tokens", root + "/tokens")}, verbose);
    forward.set_dead_work(!has_flag(argc, argv, "--lean"));
    forward.set_encrypted_projection(has_flag(argc, argv, "--encrypted-projection"));
    forward.set_all_token_attention(has_flag(argc, argv, "--all-tokens"));
    forward.set_packed(has_flag(argc, argv, "--packed"));
    Ctxt encoded;
    if (has_flag(argc, argv, "--resume")) {
        encoded = controller.load_ciphertext(root + "/checkpoint/encodered.bin");
    } else {
        encoded = forward.encoder();
        controller.save(encoded, root + "/checkpoint/encodered.bin");
    }
    const Ctxt classified = forward.classifier(forward.pooler(encoded));
    if (verbose) std::cout << "The circuit has been evaluated, the results are sent back to the client\n\nCLIENT-SIDE" << std::endl;
    const std::vector<double> z = forward.logits(classified);
    if (verbose) std::cout << "\nThe evaluation of the FHE circuit took: " << std::chrono::duration<double>(utils::clock_type::now() - start).count() << " seconds." << std::endl;
    std::vector<double> prob;
    const int pred = flh::LinformerForward::argmax_softmax(z, &prob);
    for (double p : prob) std::cout << "Softmax Prob: " << p << std::endl;
    std::cout << "Pred: " << pred << std::endl;
    return 0;
}
