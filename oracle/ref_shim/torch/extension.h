// oracle shim: the few torch::Tensor members the reference's *_api wrappers touch
// (data_ptr, sizes, scalar_type), backed by caller-owned raw pointers so the
// reference's host wrappers run unmodified on CPU. Test infrastructure only.
#pragma once
#include <cstdint>
#include <vector>
namespace at {
enum class ScalarType { Float, Half, Double, Int, Byte };
struct Half { _Float16 v; Half() = default; Half(float f) : v((_Float16)f) {} operator float() const { return (float)v; } };
}
namespace torch {
struct Tensor {
    void *p = nullptr; std::vector<int64_t> shape; at::ScalarType st = at::ScalarType::Float;
    Tensor() = default;
    Tensor(void *p_, std::vector<int64_t> s, at::ScalarType t = at::ScalarType::Float) : p(p_), shape(std::move(s)), st(t) {}
    void *data_ptr() const { return p; }
    const std::vector<int64_t> &sizes() const { return shape; }
    at::ScalarType scalar_type() const { return st; }
};
}
