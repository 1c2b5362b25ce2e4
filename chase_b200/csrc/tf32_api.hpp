// chase_b200 — what kernels_capi.cu needs to know about the tcgen05 kind::tf32 path (hemm_tf32.cu)
#pragma once
#include <cstddef>
#include <cstdint>

namespace cb2
{
// a single-precision matrix registered with the lo part of its TF32 split (chase_b200_tf32_register)
struct Tf32Reg
{
    void* lo;
    int64_t ld, rows, cols;
    int kind; // 0: Hermitian (A B = A^H B), 1: pseudo-Hermitian (H B = S H^H S B), 2: general (only A^H B)
    void* scratch;
    size_t scratch_bytes;
};
bool tf32_lookup(const void* A, Tf32Reg* out);
int tf32_terms();
} // namespace cb2
