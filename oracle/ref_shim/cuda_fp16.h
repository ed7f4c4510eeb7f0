// oracle shim: minimal stand-in for <cuda_fp16.h>. The reference's half paths are
// unreachable but must parse. Test infrastructure only.
#pragma once
struct __half { _Float16 v; __half() = default; __half(float f) : v((_Float16)f) {} operator float() const { return (float)v; } };
typedef __half half;
