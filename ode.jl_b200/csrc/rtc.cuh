// rtc.cuh -- run-time registered right-hand sides (see rtc.cu)
#pragma once
namespace odeb200 {
bool user_rhs_info(int rhs, int* dof, int* np, bool* has_jac);
const void* user_kernel(int rhs, int method, int mode, int literal, int jac);
}  // namespace odeb200
