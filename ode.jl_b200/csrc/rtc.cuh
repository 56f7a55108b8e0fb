// rtc.cuh -- run-time registered right-hand sides (see rtc.cu)
#pragma once
namespace odeb200 {
bool user_rhs_info(int rhs, int* dof, int* np, bool* has_jac);
const void* user_kernel(int rhs, int method, int mode, int literal, int jac);
const void* user_init_kernel(int rhs, int method);  // hinit pre-pass compiled with the mode-0 adaptive kernel
bool user_stencil_info(int rhs, int* radius, int* np);
// kernels of the large-system path for a registered stencil; false + last_error on failure
bool user_stencil_kernels(int rhs, int method, const void** attempt, const void** attempt_literal, const void** hinit_f0,
                          const void** hinit_f1);
// registered general large-system right-hand sides (field.cuh): parameter count; kernels for one tableau
bool user_field_info(int rhs, int* np);
bool user_field_kernels(int rhs, int method, int n_stages, const void** out);
// fixed-step stage kernels for a registered field or stencil (out[s-1] = stage s)
bool user_fixed_kernels(int rhs, int method, int n_stages, const void** out);
}  // namespace odeb200
