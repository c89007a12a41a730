// sm/Lazy.h -- op-chain fusion front end.  NOT part of the reference's
// surface: an opt-in addition beside it.
//
// Every SMArray operator of the reference materialises a full temporary
// (include/SMArray.h:217-305: `new T[totalSize]` per operator, and again in sm::pow,
// include/UserFunctions.h:42-48), so `(a + b) * c - d` moves 9 arrays' worth of bytes for
// 4 inputs and 1 output.  sm::lazy(a) starts a chain that records the operators instead and
// evaluates them in ONE kernel (smb_chain) when it is converted to an SMArray:
//
//     sm::SMArray<float> r = (sm::lazy(a) + b) * c - d;        // one pass, 5 arrays of traffic
//     sm::SMArray<float> p = sm::pow(sm::lazy(a) + b, 2.5f);    // sm::pow(a + b, e) fused
//     sm::SMArray<float> q = 1.0f / (sm::lazy(a) * 2.0f + b);   // scalars on either side
//
// The same Op structs drive it (Operation::device_op, see math/add.h); the chain applies them
// left to right on an accumulator, each intermediate rounded to T, so + - * / results are
// bit-identical to the unfused operators.  A right operand that is itself a chain is evaluated
// first and joins as a leaf; a chain longer than SMB_CHAIN_MAX is cut the same way.
// Arrays named in a chain must outlive its evaluation (they are referenced, not copied).
#pragma once
#include <memory>
#include <stdexcept>
#include <vector>

#include "SMArray.h"
#include "math/pow.h"

namespace sm {
    template<typename T>
    class Lazy {
    public:
        explicit Lazy(const SMArray<T> &first) { push(0, false, &first, T{}); }

        // acc = acc (Operation) leaf   /   acc = leaf (Operation) acc
        template<typename Operation> Lazy &&then(const SMArray<T> &leaf, bool leaf_on_left = false) && {
            push(smb::OpTag<Operation>::id(), leaf_on_left, &leaf, T{});
            return std::move(*this);
        }
        template<typename Operation> Lazy &&then(T value, bool leaf_on_left = false) && {
            push(smb::OpTag<Operation>::id(), leaf_on_left, nullptr, value);
            return std::move(*this);
        }
        template<typename Operation> Lazy &&then(Lazy &&rhs, bool leaf_on_left = false) && {
            owned_.push_back(std::make_unique<SMArray<T> >(rhs.eval())); // a nested chain becomes a leaf
            push(smb::OpTag<Operation>::id(), leaf_on_left, owned_.back().get(), T{});
            return std::move(*this);
        }

        // One kernel for the whole chain; the result is a fresh dense array of the broadcast shape.
        SMArray<T> eval() const {
            std::vector<size_t> shape = result_shape();
            const size_t n = calculateTotalSize(shape);
            if (shape.size() > MAX_NDIM) throw std::runtime_error("smb200: rank exceeds MAX_NDIM");
            std::vector<smb_chain_step> steps(steps_.size());
            for (size_t i = 0; i < steps_.size(); ++i) {
                const Step &s = steps_[i];
                smb_chain_step &c = steps[i];
                c = smb_chain_step{};
                c.op = s.op;
                c.swap = s.leaf_on_left ? 1 : 0;
                if (s.arr) {
                    c.data = s.arr->data;
                    const auto &sh = s.arr->shape();
                    const auto &st = s.arr->strides();
                    const size_t pad = shape.size() - sh.size();
                    for (size_t k = 0; k < sh.size(); ++k)
                        c.stride[pad + k] = (sh[k] == 1 && shape[pad + k] > 1) ? 0 : st[k];
                } else if constexpr (std::is_same_v<T, float>) c.value.f32 = s.value;
                else if constexpr (std::is_same_v<T, double>) c.value.f64 = s.value;
                else c.value.i32 = s.value;
            }
            T *result = storage::acquire<T>(n);
            try {
                smb::check(smb_chain(smb::DTypeTag<T>::value, steps.data(), static_cast<int>(steps.size()),
                                     smb::u64(shape), static_cast<int>(shape.size()), n, result, nullptr));
            } catch (...) {
                storage::release(result);
                throw;
            }
            return SMArray<T>(result, std::move(shape));
        }

        operator SMArray<T>() const { return eval(); }

        Lazy(Lazy &&) noexcept = default;
        Lazy &operator=(Lazy &&) noexcept = default;

    private:
        struct Step {
            int op;
            bool leaf_on_left;
            const SMArray<T> *arr; // nullptr: the constant
            T value;
        };
        std::vector<Step> steps_;
        std::vector<std::unique_ptr<SMArray<T> > > owned_; // materialised sub-chains

        void push(int op, bool leaf_on_left, const SMArray<T> *arr, T value) {
            if (steps_.size() == SMB_CHAIN_MAX) { // cut: what we have becomes the first leaf of a new chain
                owned_.push_back(std::make_unique<SMArray<T> >(eval()));
                steps_.clear();
                steps_.push_back(Step{0, false, owned_.back().get(), T{}});
            }
            steps_.push_back(Step{op, leaf_on_left, arr, value});
        }

        // sm::broadcast over all array leaves (SMUtils.h:34-99 semantics; throws on mismatch).
        std::vector<size_t> result_shape() const {
            std::vector<size_t> shape, strides;
            for (const Step &s: steps_) {
                if (!s.arr) continue;
                if (shape.empty()) { shape = s.arr->shape(); strides = s.arr->strides(); continue; }
                auto bc = sm::broadcast(shape, strides, s.arr->shape(), s.arr->strides());
                shape = std::move(bc.resultShape);
                strides.assign(shape.size(), 1);
            }
            return shape;
        }
    };

    template<typename T> Lazy<T> lazy(const SMArray<T> &a) { return Lazy<T>(a); }

    // chain (op) array | scalar | chain, and the mirrored forms with the chain on the right
#define SM_LAZY_OPERATOR(SYM, OP)                                                                                     \
    template<typename T> Lazy<T> operator SYM(Lazy<T> &&l, const SMArray<T> &r) { return std::move(l).template then<OP<T> >(r); }        \
    template<typename T> Lazy<T> operator SYM(Lazy<T> &&l, std::type_identity_t<T> r) { return std::move(l).template then<OP<T> >(r); }  \
    template<typename T> Lazy<T> operator SYM(Lazy<T> &&l, Lazy<T> &&r) { return std::move(l).template then<OP<T> >(std::move(r)); }     \
    template<typename T> Lazy<T> operator SYM(const SMArray<T> &l, Lazy<T> &&r) { return std::move(r).template then<OP<T> >(l, true); }  \
    template<typename T> Lazy<T> operator SYM(std::type_identity_t<T> l, Lazy<T> &&r) { return std::move(r).template then<OP<T> >(l, true); }
    SM_LAZY_OPERATOR(+, AddOp)
    SM_LAZY_OPERATOR(-, SubtractOp)
    SM_LAZY_OPERATOR(*, MultiplyOp)
    SM_LAZY_OPERATOR(/, DivideOp)
#undef SM_LAZY_OPERATOR

    // sm::pow(chain, e): the power joins the chain (array ^ scalar, like sm::pow(SMArray&, T)).
    template<typename T> Lazy<T> pow(Lazy<T> &&l, std::type_identity_t<T> e) { return std::move(l).template then<PowOp<T> >(e); }
} // namespace sm
