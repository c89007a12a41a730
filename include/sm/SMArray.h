// sm/SMArray.h -- drop-in for the reference's include/SMArray.h.
//
// Same public surface (SMArray.h:33-346 of the reference): public `data` /
// `totalSize`, the three constructors, move, element assignment, operator()
// for indices and slices, transpose / repeat, operator% and the eight
// arithmetic operators, toString / shape / strides.
//
// What changed underneath:
//   * storage comes from the pooled allocator of libsmb200.so (smb_alloc,
//     managed memory: HBM-resident while kernels run, still dereferenceable
//     through `data` on the host, which the reference's tests rely on --
//     tests/pow.cpp:48-51, tests/add.cpp:67-71) instead of new[] / delete[];
//   * operators + - * / launch the sm_100a kernels through element_wise_op /
//     array_scalar_op (sm/math/calculate.h) instead of the OpenMP/AVX loops.
// Views keep aliasing interior pointers without reference counts, exactly as in
// the reference (SMArray.h:121-136, 397-437).
#pragma once
#include <cassert>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

#include <smb200.h>

#include "Slice.h"
#include "SMUtils.h"
#include "math/add.h"
#include "math/subtract.h"
#include "math/multiply.h"
#include "math/division.h"
#include "math/calculate.h"
#include "math/product.h"

namespace sm {
    template<typename T>
    concept ArithmeticOrComplex =
            std::is_arithmetic_v<T> ||
            (requires { typename T::value_type; } &&
             std::is_arithmetic_v<typename T::value_type> &&
             std::same_as<T, std::complex<typename T::value_type> >);

    namespace storage {
        // The host is about to read or write array memory: wait for kernels an sm::async_scope left
        // in flight (one atomic load when there are none).
        inline void host_access() {
            if (smb_wait_pending() != SMB_OK) throw std::runtime_error(std::string("smb200: ") + smb_last_error());
        }

        // One block of n elements of T: managed memory from the pool.  Throws
        // std::runtime_error when there is no device (no CPU fallback).
        template<typename T>
        inline T *acquire(size_t n) {
            void *p = smb_alloc((n ? n : 1) * sizeof(T), SMB_MEM_MANAGED);
            if (!p) throw std::runtime_error(std::string("smb200: ") + smb_last_error());
            return static_cast<T *>(p);
        }

        // Blocks handed to the adopting constructor by user code may be plain
        // new[] memory (what the reference expects, SMArray.h:70-76,342-346).
        template<typename T>
        inline void release(T *p) {
            if (!p) return;
            if (smb_owns(p)) smb_free(p);
            else delete[] p;
        }
    }

    template<ArithmeticOrComplex T>
    class SMArray {
    public:
        T *data = nullptr;
        size_t totalSize = 0;

        SMArray(const std::initializer_list<T> &list) {
            totalSize = list.size();
            _shape = {totalSize};
            _strides = {1};
            ndim = 1;
            data = storage::acquire<T>(totalSize);
            storage::host_access(); // a recycled block may still be in use by kernels an async scope left in flight
            std::memcpy(data, list.begin(), totalSize * sizeof(T));
            smb_host_written(data); // pages are on the host now: prefetch before the first kernel
        }

        // Rows given as arrays: shape = {rows, child shape...}; children are dense.
        SMArray(const std::initializer_list<SMArray> &list) {
            const SMArray &first = *list.begin();
            _shape.reserve(first._shape.size() + 1);
            _shape.push_back(list.size());
            _shape.insert(_shape.end(), first._shape.begin(), first._shape.end());
            ndim = _shape.size();
            totalSize = calculateTotalSize(_shape);
            data = storage::acquire<T>(totalSize);
            storage::host_access();
            size_t at = 0;
            for (const SMArray &row: list) {
                std::memcpy(data + at, row.data, row.totalSize * sizeof(T));
                at += row.totalSize;
            }
            smb_host_written(data);
            calculateStride();
        }

        // Adopts `data` (ownership passes to the array).
        SMArray(T *data, std::vector<size_t> &&shape) {
            _shape = std::move(shape);
            ndim = _shape.size();
            this->data = data;
            totalSize = calculateTotalSize(_shape);
            calculateStride();
        }

        SMArray(SMArray &&other) noexcept {
            data = other.data;
            totalSize = other.totalSize;
            ndim = other.ndim;
            isView = other.isView;
            _shape = std::move(other._shape);
            _strides = std::move(other._strides);
            other.data = nullptr;
        }

        // Element-wise copy into an existing array of the same shape.
        SMArray &operator=(const SMArray &&other) {
            assert(_shape.size() == other._shape.size() && "Shape mismatch in assignment");
            for (size_t i = 0; i < _shape.size(); ++i)
                assert(_shape[i] == other._shape[i] && "Shape mismatch in assignment");
            storage::host_access();
            for (size_t i = 0; i < totalSize; ++i) data[i] = other.data[i];
            noteHostTouch();
            return *this;
        }

        template<typename... Args>
            requires ((std::is_integral_v<std::remove_cvref_t<Args> > || std::is_same_v<Args, Slice>) && ...)
        auto operator()(Args &&... args) const {
            if constexpr ((std::is_integral_v<std::remove_cvref_t<Args> > && ...)) {
                auto indices = {static_cast<std::size_t>(args)...};
                return accessByValue(indices);
            } else {
                std::initializer_list<Slice> slices = {processIndex(std::forward<Args>(args))...};
                return accessByArray(slices);
            }
        }

        template<typename... Args>
            requires ((std::is_integral_v<std::remove_cvref_t<Args> > && ...))
        ALWAYS_INLINE T &operator()(Args &&... args) {
            auto indices = {static_cast<size_t>(args)...};
            return accessByValueRef(indices);
        }

        // Reversed shape and strides over the same buffer.
        SMArray transpose() const {
            SMArray view;
            view.data = data;
            view.isView = true;
            view.ndim = ndim;
            view._shape.assign(_shape.rbegin(), _shape.rend());
            view._strides.assign(_strides.rbegin(), _strides.rend());
            view.totalSize = totalSize;
            return view;
        }

        // np.repeat semantics (each element repeated in place along the flattened array / along
        // `axis`); the reference's 1-D loop (SMArray.h:138-160) writes newData[i + j] and leaves most
        // of its result uninitialised.  Runs on the device: the result is
        // the input broadcast along an inserted stride-0 dim, materialised by a one-leaf smb_chain
        // (the reference's serial host loop would also pull the managed pages back to the host).
        SMArray repeat(int numberOfRepeats) const {
            assert(numberOfRepeats > 1);
            const size_t reps = static_cast<size_t>(numberOfRepeats);
            if constexpr (requires { smb::DTypeTag<T>::value; }) {
                if (isDense()) return broadcastCopy({totalSize, reps}, {1, 0}, {totalSize * reps});
            }
            T *fresh = storage::acquire<T>(totalSize * reps);
            storage::host_access();
            for (size_t i = 0; i < totalSize; ++i)
                for (size_t j = 0; j < reps; ++j) fresh[i * reps + j] = data[i];
            smb_host_written(fresh);
            return SMArray(fresh, {totalSize * reps});
        }

        SMArray repeat(int numberOfRepeats, int axis) const {
            assert(axis >= 0 && axis < static_cast<int>(ndim));
            if (ndim == 1) return repeat(numberOfRepeats);
            const size_t reps = static_cast<size_t>(numberOfRepeats);
            std::vector<size_t> newShape = _shape;
            newShape[axis] *= reps;
            if constexpr (requires { smb::DTypeTag<T>::value; }) {
                if (ndim < MAX_NDIM) { // {.., shape[axis], reps, ..} with stride 0 on the inserted dim; views welcome
                    std::vector<size_t> shape = _shape, strides = _strides;
                    shape.insert(shape.begin() + axis + 1, reps);
                    strides.insert(strides.begin() + axis + 1, 0);
                    return broadcastCopy(std::move(shape), std::move(strides), std::move(newShape));
                }
            }
            size_t inner = 1;
            for (size_t k = axis + 1; k < ndim; ++k) inner *= _shape[k];
            const size_t outer = totalSize / (inner * _shape[axis]);
            T *fresh = storage::acquire<T>(totalSize * reps);
            storage::host_access();
            smb_host_written(fresh);
            T *w = fresh;
            for (size_t o = 0; o < outer; ++o)
                for (size_t a = 0; a < _shape[axis]; ++a) {
                    const T *src = data + (o * _shape[axis] + a) * inner;
                    for (size_t j = 0; j < reps; ++j, w += inner) std::memcpy(w, src, inner * sizeof(T));
                }
            return SMArray(fresh, std::move(newShape));
        }

        // Dot product over the dense data (reference SMArray.h:213-215 -> math/product.h): the
        // first row widened after the elementwise path; device reduction for float/double/int32.
        T operator%(SMArray &arr) const { return dot_product(data, arr.data, arr.totalSize); }

        SMArray operator+(const SMArray &arr) const { return binary<AddOp<T> >(arr); }
        SMArray operator+(const T val) const { return withScalar<AddOp<T> >(val); }
        SMArray operator-(const SMArray &arr) const { return binary<SubtractOp<T> >(arr); }
        SMArray operator-(const T val) const { return withScalar<SubtractOp<T> >(val); }
        SMArray operator*(const SMArray &arr) const { return binary<MultiplyOp<T> >(arr); }
        SMArray operator*(const T val) const { return withScalar<MultiplyOp<T> >(val); }
        SMArray operator/(const SMArray &arr) const { return binary<DivideOp<T> >(arr); }
        SMArray operator/(const T val) const { return withScalar<DivideOp<T> >(val); }

        // array (op) scalar through any Op struct; what sm::pow uses.
        template<typename Operation>
        SMArray applyScalar(const T val) const { return withScalar<Operation>(val); }

        [[nodiscard]] std::string toString() const {
            storage::host_access();
            noteHostTouch();
            std::ostringstream os;
            std::function<void(size_t, size_t)> emit = [&](size_t offset, size_t dim) {
                os << "[";
                for (size_t i = 0; i < _shape[dim]; ++i) {
                    if (dim + 1 == ndim) {
                        if (i) os << ", ";
                        os << data[offset + i * _strides[dim]];
                    } else {
                        if (i) os << ",\n";
                        emit(offset + i * _strides[dim], dim + 1);
                    }
                }
                os << "]";
            };
            emit(0, 0);
            return os.str();
        }

        [[nodiscard]] const std::vector<size_t> &shape() const { return _shape; }

        [[nodiscard]] const std::vector<size_t> &strides() const { return _strides; }

        ~SMArray() {
            if (!isView) storage::release(data);
        }

    private:
        std::vector<size_t> _shape;
        std::vector<size_t> _strides;
        size_t ndim = 0;
        bool isView = false;
        // The library has been told that the host touched this array since it last went to a kernel.
        // Element access through operator() -- the way the reference's tests fill and check arrays,
        // tests/add.cpp:67-71 -- pulls managed pages to host memory, reads as well as writes; the next
        // kernel then prefetches the block back in one transfer instead of demand-paging it.
        // (smb_host_written is a lock and a map lookup: once per fill loop, not once per element.)
        mutable bool hostTouchNoted = false;

        SMArray() = default;

        void noteHostTouch() const {
            if (hostTouchNoted) return;
            smb_host_written(data);
            hostTouchNoted = true;
        }

        // Row-major strides in elements (reference SMArray.h:357-364).
        void calculateStride() {
            _strides.assign(ndim, 1);
            for (size_t k = ndim; k-- > 1;) _strides[k - 1] = _strides[k] * _shape[k];
        }

        bool isDense() const { return is_contiguous(_shape, _strides); }

        // Dense copy of this array's data seen through (shape, strides) -- stride 0 repeats -- made on
        // the device by a one-leaf chain; the result gets `resultShape` (same element count).
        SMArray broadcastCopy(std::vector<size_t> shape, std::vector<size_t> strides, std::vector<size_t> resultShape) const {
            const size_t n = calculateTotalSize(shape);
            smb_chain_step leaf{};
            leaf.data = data;
            for (size_t k = 0; k < shape.size(); ++k) leaf.stride[k] = strides[k];
            T *fresh = storage::acquire<T>(n);
            try {
                smb::check(smb_chain(smb::DTypeTag<T>::value, &leaf, 1, smb::u64(shape), static_cast<int>(shape.size()), n, fresh,
                                     nullptr));
            } catch (...) {
                storage::release(fresh);
                throw;
            }
            return SMArray(fresh, std::move(resultShape));
        }

        // SMArray (op) SMArray: broadcast -> fresh dense result -> element_wise_op
        // (reference SMArray.h:217-225 and siblings).
        template<typename Operation>
        SMArray binary(const SMArray &arr) const {
            auto bc = sm::broadcast(_shape, _strides, arr._shape, arr._strides);
            T *result = storage::acquire<T>(bc.totalSize);
            hostTouchNoted = arr.hostTouchNoted = false; // both go to a kernel now
            try {
                element_wise_op<T, Operation>(data, bc.newStrides1, arr.data, bc.newStrides2, bc.totalSize, result,
                                              bc.resultShape);
            } catch (...) {
                storage::release(result);
                throw;
            }
            return SMArray(result, std::move(bc.resultShape));
        }

        // SMArray (op) scalar (reference SMArray.h:226-237 and siblings): dense
        // arrays take array_scalar_op over data[0..totalSize).  The reference
        // does the same for views and then reads the wrong elements (it ignores
        // strides, calculate.h:137-169); here a non-dense view goes through a
        // two-step smb_chain (the view, then the constant), so its result is the view's.
        template<typename Operation>
        SMArray withScalar(const T val) const {
            T *result = storage::acquire<T>(totalSize);
            hostTouchNoted = false;
            try {
                if (isDense()) {
                    array_scalar_op<T, Operation>(data, val, totalSize, result);
                } else {
                    // a strided view: one two-step chain (view leaf, then the constant) -- no temporary
                    smb_chain_step steps[2] = {};
                    steps[0].data = data;
                    for (size_t k = 0; k < ndim; ++k) steps[0].stride[k] = _strides[k];
                    steps[1].op = smb::OpTag<Operation>::id();
                    if constexpr (std::is_same_v<T, float>) steps[1].value.f32 = val;
                    else if constexpr (std::is_same_v<T, double>) steps[1].value.f64 = val;
                    else steps[1].value.i32 = val;
                    if (ndim > MAX_NDIM) throw std::runtime_error("smb200: rank exceeds MAX_NDIM");
                    smb::check(smb_chain(smb::DTypeTag<T>::value, steps, 2, smb::u64(_shape), static_cast<int>(ndim), totalSize,
                                         result, nullptr));
                }
            } catch (...) {
                storage::release(result);
                throw;
            }
            std::vector<size_t> shape = _shape;
            return SMArray(result, std::move(shape));
        }

        T *locate(const std::initializer_list<std::size_t> &indices) const {
            T *p = data;
            size_t axis = 0;
            for (size_t index: indices) {
                assert(index < _shape[axis] && "Index out of bounds");
                p += index * _strides[axis];
                ++axis;
            }
            return p;
        }

        T accessByValue(const std::initializer_list<std::size_t> &indices) const {
            assert(indices.size() <= ndim && "Number of indices exceeds number of dimensions");
            storage::host_access();
            noteHostTouch();
            return *locate(indices);
        }

        T &accessByValueRef(const std::initializer_list<std::size_t> &indices) const {
            assert(indices.size() == ndim && "Number of indices exceeds number of dimensions");
            storage::host_access();
            noteHostTouch();
            return *locate(indices);
        }

        // Index / slice view (reference SMArray.h:397-437): an INDEX moves the
        // base pointer and drops the axis, a SLICE keeps it (step is always 1).
        const SMArray accessByArray(const std::initializer_list<Slice> &slices) const {
            SMArray view;
            view.data = data;
            view.isView = true;
            for (size_t axis = 0; axis < ndim; ++axis) {
                const Slice s = axis < slices.size() ? *(slices.begin() + axis) : Slice(0, -1);
                view.data += s.start * _strides[axis];
                if (s.sliceType == Slice::INDEX) continue;
                const size_t stop = s.end == static_cast<size_t>(-1) ? _shape[axis] : s.end;
                view._shape.push_back(stop > s.start ? stop - s.start : s.start - stop);
                view._strides.push_back(_strides[axis]);
            }
            view.ndim = view._shape.size();
            view.totalSize = calculateTotalSize(view._shape);
            return view;
        }
    };
} // namespace sm
