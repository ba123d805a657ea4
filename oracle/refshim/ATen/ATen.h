// TEST INFRASTRUCTURE -- stand-in for <ATen/ATen.h>, just enough to compile the reference's
// sdf/sdf/csrc/sdf_cuda_kernel.cu UNCHANGED (its device code is plain templated CUDA; only the host launcher
// sdf_cuda() at :307-335 touches ATen: .size(), .type(), .data<T>(), AT_DISPATCH_FLOATING_TYPES).
// Used by oracle/build_ref_sdf.sh only; never part of the product.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
namespace at {
enum class ScalarType { Float, Double, Int };
struct Tensor {
    void* ptr = nullptr;
    int64_t sizes[4] = {0, 0, 0, 0};
    ScalarType st = ScalarType::Float;
    int64_t size(int i) const { return sizes[i]; }
    ScalarType type() const { return st; }                       // the removed Tensor::type(): only its dtype is used
    template <class T> T* data() const { return static_cast<T*>(ptr); }
};
}  // namespace at
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...)                                        \
    [&] {                                                                                  \
        switch (TYPE) {                                                                    \
        case at::ScalarType::Float: { using scalar_t = float; return __VA_ARGS__(); }      \
        case at::ScalarType::Double: { using scalar_t = double; return __VA_ARGS__(); }    \
        default: fprintf(stderr, "%s: not a floating type\n", NAME); abort();              \
        }                                                                                  \
    }()
