// Utils.h -- host helpers with the names the reference's sources expect in namespace `utils`
// (timers, text-file reader, approximation error; reference Utils.h:21-125), written against the compat types.
#pragma once
#include <chrono>
#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "openfhe_compat.h"

#define YELLOW_TEXT "\033[1;33m"
#define RESET_COLOR "\033[0m"

namespace utils {

using clock_type = std::chrono::steady_clock;
using time_point = std::chrono::time_point<clock_type, std::chrono::nanoseconds>;

inline std::chrono::nanoseconds& total_time() {
    static std::chrono::nanoseconds t{0};
    return t;
}
static inline time_point start_time() { return clock_type::now(); }

static inline void print_duration(time_point start, const std::string& title) {
    const auto d = clock_type::now() - start;
    total_time() += d;
    const double s = std::chrono::duration<double>(d).count();
    const double tot = std::chrono::duration<double>(total_time()).count();
    std::cout << "(" << title << "): " << YELLOW_TEXT << std::fixed << std::setprecision(3) << s << " s" << RESET_COLOR << " (total " << tot << " s)"
              << std::endl;
}

// comma- and/or newline-separated decimals (the %.18e text files of src/python/dimReduce.py:11-14)
static inline std::vector<double> read_values_from_file(const std::string& filename, double scale = 1) {
    std::vector<double> values;
    std::ifstream in(filename);
    if (!in.is_open()) {
        std::cerr << "Can not open " << filename << std::endl;
        return values;
    }
    std::string line, tok;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        while (std::getline(ls, tok, ',')) {
            try {
                values.push_back(std::stod(tok) * scale);
            } catch (const std::exception&) {
                if (tok.find_first_not_of(" \t\r") != std::string::npos) std::cerr << "Can not convert: " << tok << std::endl;
            }
        }
    }
    return values;
}

// infinity-norm distance of the real parts, reported as |log2| like the reference's helper
static inline double compute_approx_error(lbcrypto::Plaintext expected, lbcrypto::Plaintext actual) {
    const auto a = actual->GetCKKSPackedValue();
    const auto e = expected->GetCKKSPackedValue();
    if (a.size() != e.size()) throw std::runtime_error("Cannot compare vectors with different numbers of elements");
    double worst = 0;
    for (size_t i = 0; i < a.size(); ++i) worst = std::max(worst, std::abs(a[i].real() - e[i].real()));
    return std::abs(std::log2(worst));
}

static inline int get_relu_depth(int degree) {   // OpenFHE FUNCTION_EVALUATION.md depth table
    const int bound[] = {5, 13, 27, 59, 119, 247, 495, 1007, 2031};
    for (int i = 0; i < 9; ++i)
        if (degree <= bound[i]) return 3 + i;
    std::cerr << "Set a valid degree for ReLU" << std::endl;
    std::exit(1);
}

static inline void write_to_file(const std::string& filename, const std::string& content) {
    std::ofstream f(filename);
    f << content;
}
static inline std::string read_from_file(const std::string& filename) {
    std::ifstream f(filename);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

}  // namespace utils
