// openfhe_compat.h -- the handful of OpenFHE names the reference's FHEController.h / main.cpp use
// (`Plaintext`, `Ciphertext<DCRTPoly>`, ->GetLevel(), ->Clone(), ->GetSlots(), ->SetLength(), ->SetSlots(),
// ->GetRealPackedValue(), ->GetCKKSPackedValue()), re-backed by handles of the
// B200 engine's C-ABI (include/fl_ckks.h).  Value semantics are the reference's: handles are shared pointers,
// Clone() is a device copy, every operation returns a new object.
#pragma once
#include <complex>
#include <cstdint>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fl_ckks.h"

namespace lbcrypto {

struct DCRTPoly {};

// owns an fl_elem*; shared by Plaintext / Ciphertext wrappers
class ElemHandle {
public:
    ElemHandle(fl_ctx* c, fl_elem* e) : ctx_(c), e_(e) {}
    ~ElemHandle() { if (e_) fl_elem_free(e_); }
    ElemHandle(const ElemHandle&) = delete;
    fl_ctx* ctx() const { return ctx_; }
    fl_elem* get() const { return e_; }
private:
    fl_ctx* ctx_;
    fl_elem* e_;
};

[[noreturn]] inline void fl_fail(const char* what) {
    // the reference reports errors on cerr and exits (FHEController.cpp:62-70); OpenFHE exceptions reach main uncaught
    throw std::runtime_error(std::string(what) + ": " + fl_last_error());
}

class PlaintextImpl {
public:
    // encoded plaintext living on the device
    PlaintextImpl(fl_ctx* c, fl_elem* e) : h_(std::make_shared<ElemHandle>(c, e)), slots_(fl_elem_slots(e)), length_(slots_) {}
    // decrypted / decoded values held on the host
    PlaintextImpl(std::vector<std::complex<double>> v, int level) : values_(std::move(v)), slots_((int)values_.size()), length_(values_.size()), level_(level) {}

    void SetLength(size_t n) { length_ = n; }
    void SetSlots(uint32_t n) { slots_ = (int)n; }
    uint32_t GetSlots() const { return (uint32_t)slots_; }
    size_t GetLength() const { return length_; }
    uint32_t GetLevel() const { return h_ ? (uint32_t)fl_elem_level(h_->get()) : (uint32_t)level_; }
    std::vector<double> GetRealPackedValue() const {
        fetch();
        std::vector<double> r(std::min(length_, values_.size()));
        for (size_t i = 0; i < r.size(); ++i) r[i] = values_[i].real();
        return r;
    }
    std::vector<std::complex<double>> GetCKKSPackedValue() const {
        fetch();
        return std::vector<std::complex<double>>(values_.begin(), values_.begin() + std::min(length_, values_.size()));
    }
    fl_elem* handle() const { return h_ ? h_->get() : nullptr; }

private:
    void fetch() const {
        if (!values_.empty() || !h_) return;
        std::vector<double> re(slots_), im(slots_);
        if (fl_decode(h_->ctx(), h_->get(), re.data(), im.data(), slots_)) fl_fail("decode");
        values_.resize(slots_);
        for (int i = 0; i < slots_; ++i) values_[i] = {re[i], im[i]};
    }
    std::shared_ptr<ElemHandle> h_;
    mutable std::vector<std::complex<double>> values_;
    int slots_ = 0;
    size_t length_ = 0;
    int level_ = 0;
};
using Plaintext = std::shared_ptr<PlaintextImpl>;

template <class Element>
class CiphertextImpl {
public:
    CiphertextImpl(fl_ctx* c, fl_elem* e) : h_(std::make_shared<ElemHandle>(c, e)) {}
    uint32_t GetLevel() const { return (uint32_t)fl_elem_level(h_->get()); }
    uint32_t GetSlots() const { return (uint32_t)fl_elem_slots(h_->get()); }
    uint32_t GetNoiseScaleDeg() const { return (uint32_t)fl_elem_deg(h_->get()); }
    double GetScalingFactor() const { return fl_elem_scale(h_->get()); }
    std::shared_ptr<CiphertextImpl<Element>> Clone() const {
        fl_elem* out = nullptr;
        if (fl_elem_clone(h_->ctx(), h_->get(), &out)) fl_fail("Clone");
        return std::make_shared<CiphertextImpl<Element>>(h_->ctx(), out);
    }
    fl_elem* handle() const { return h_->get(); }
    fl_ctx* ctx() const { return h_->ctx(); }

private:
    std::shared_ptr<ElemHandle> h_;
};
template <class Element>
using Ciphertext = std::shared_ptr<CiphertextImpl<Element>>;
template <class Element>
using ConstCiphertext = std::shared_ptr<const CiphertextImpl<Element>>;

}  // namespace lbcrypto
