// rtc.cu -- right-hand sides registered at run time.
//
// The reference takes any Julia function as `F`.  A closure cannot cross the C ABI, so the built-in
// functors of rhs.cuh are compiled in; for everything else a caller registers the BODY of the device
// functor as CUDA C++ source (ode_b200_register_rhs).  On first use with a given method the library
// compiles, with NVRTC, the very same kernel templates (ensemble_*.cuh) instantiated for that functor --
// same arithmetic rules (-fmad=false), same controller, same bits -- and caches the loaded kernel.
// NVRTC is opened with dlopen, so the library itself does not depend on it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ode_b200.h"
#include "host_util.cuh"
#include "rtc.cuh"

namespace odeb200 {

struct UserRhs {
  std::string name, eval_body, jac_body;
  int dof, np;
};
static std::vector<UserRhs> g_user;
static std::vector<UserRhs> g_fields;    // eval_body = body of rhs(i, n, y, p, t); large-system gather regime
static std::vector<UserRhs> g_stencils;  // eval_body = body of point(um, u, up, p0, p1)
static std::map<std::string, const void*> g_kernels;  // key -> cudaKernel_t
static std::mutex g_mu;

struct Nvrtc {
  void* h = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
  nvrtcResult (*AddNameExpression)(nvrtcProgram, const char*);
  nvrtcResult (*GetLoweredName)(nvrtcProgram, const char*, const char**);
  nvrtcResult (*DestroyProgram)(nvrtcProgram*);
  const char* (*GetErrorString)(nvrtcResult);
};
static Nvrtc g_nvrtc;

static int load_nvrtc() {
  if (g_nvrtc.h) return 0;
  const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
  void* h = nullptr;
  for (const char* n : names)
    if ((h = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
  if (!h) return set_err(ODE_B200_E_UNSUPPORTED, "NVRTC (libnvrtc.so.12) not found: run-time RHS registration unavailable");
#define LOAD(sym)                                                                   \
  *(void**)(&g_nvrtc.sym) = dlsym(h, "nvrtc" #sym);                                 \
  if (!g_nvrtc.sym) return set_err(ODE_B200_E_UNSUPPORTED, "nvrtc" #sym " missing")
  LOAD(CreateProgram);
  LOAD(CompileProgram);
  LOAD(GetProgramLogSize);
  LOAD(GetProgramLog);
  LOAD(GetCUBINSize);
  LOAD(GetCUBIN);
  LOAD(AddNameExpression);
  LOAD(GetLoweredName);
  LOAD(DestroyProgram);
  LOAD(GetErrorString);
#undef LOAD
  g_nvrtc.h = h;
  return 0;
}

static std::string library_dir() {
  Dl_info info;
  if (dladdr((void*)&ode_b200_version, &info) && info.dli_fname) {
    std::string p(info.dli_fname);
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? std::string(".") : p.substr(0, k);
  }
  return ".";
}

bool user_rhs_info(int rhs, int* dof, int* np, bool* has_jac) {
  std::lock_guard<std::mutex> lk(g_mu);
  int i = rhs - ODE_B200_RHS_USER_BASE;
  if (i < 0 || i >= (int)g_user.size()) return false;
  if (dof) *dof = g_user[i].dof;
  if (np) *np = g_user[i].np;
  if (has_jac) *has_jac = !g_user[i].jac_body.empty();
  return true;
}

bool user_stencil_info(int rhs, int* radius, int* np) {
  std::lock_guard<std::mutex> lk(g_mu);
  int i = rhs - ODE_B200_STENCIL_USER_BASE;
  if (i < 0 || i >= (int)g_stencils.size()) return false;
  if (radius) *radius = g_stencils[i].dof;  // (the dof slot of a stencil entry holds its radius)
  if (np) *np = g_stencils[i].np;
  return true;
}

static const char* tab_name(int method) {
  switch (method) {
    case ODE_B200_M_FEULER: return "odeb200::TabFEuler";
    case ODE_B200_M_MIDPOINT: return "odeb200::TabMidpoint";
    case ODE_B200_M_HEUN: return "odeb200::TabHeun";
    case ODE_B200_M_RK4: return "odeb200::TabRK4";
    case ODE_B200_M_RK21: return "odeb200::TabRK21";
    case ODE_B200_M_RK23: return "odeb200::TabRK23";
    case ODE_B200_M_RK45_FE: return "odeb200::TabRK45";
    case ODE_B200_M_DOPRI5: return "odeb200::TabDopri5";
    case ODE_B200_M_FEH78: return "odeb200::TabFeh78";
    case ODE_B200_M_CK45: return "odeb200::TabCK45";
    default: return nullptr;
  }
}

// NVRTC-compile `src` (which includes this library's kernel headers), load the cubin and resolve the kernels
// named by the template-instantiation expressions `exprs`.
static bool compile_and_load(const std::string& what, const std::string& src, const char* const* exprs, int n,
                             const void** out) {
  if (load_nvrtc()) return false;
  nvrtcProgram prog;
  nvrtcResult r = g_nvrtc.CreateProgram(&prog, src.c_str(), "user_rhs.cu", 0, nullptr, nullptr);
  if (r != NVRTC_SUCCESS) {
    set_err(ODE_B200_E_CUDA, "nvrtcCreateProgram: %s", g_nvrtc.GetErrorString(r));
    return false;
  }
  for (int i = 0; i < n; i++) g_nvrtc.AddNameExpression(prog, exprs[i]);
  const std::string dir = library_dir();
  const std::string i1 = "-I" + dir + "/csrc", i2 = "-I" + dir + "/../include";
  // -fmad=false: the reference never fuses multiply-add; -default-device: lambdas inside the kernels
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-fmad=false", "-lineinfo", "-default-device",
                        i1.c_str(), i2.c_str()};
  r = g_nvrtc.CompileProgram(prog, 7, opts);
  if (r != NVRTC_SUCCESS) {
    size_t nl = 0;
    g_nvrtc.GetProgramLogSize(prog, &nl);
    std::string log(nl + 1, '\0');
    if (nl) g_nvrtc.GetProgramLog(prog, &log[0]);
    set_err(ODE_B200_E_INVALID, "NVRTC compilation of rhs '%s' failed: %.380s", what.c_str(), log.c_str());
    g_nvrtc.DestroyProgram(&prog);
    return false;
  }
  size_t nb = 0;
  g_nvrtc.GetCUBINSize(prog, &nb);
  std::vector<char> cubin(nb);
  g_nvrtc.GetCUBIN(prog, cubin.data());
  cudaLibrary_t lib = nullptr;
  cudaError_t e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  for (int i = 0; i < n && e == cudaSuccess; i++) {
    const char* lowered = nullptr;
    g_nvrtc.GetLoweredName(prog, exprs[i], &lowered);
    cudaKernel_t kern = nullptr;
    e = cudaLibraryGetKernel(&kern, lib, lowered ? lowered : "");
    out[i] = (const void*)kern;
  }
  g_nvrtc.DestroyProgram(&prog);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_err(ODE_B200_E_CUDA, "loading the compiled rhs '%s' failed: %s", what.c_str(), cudaGetErrorString(e));
    return false;
  }
  return true;
}

// Kernel of `method` instantiated for the registered RHS; nullptr + last_error on failure.
const void* user_kernel(int rhs, int method, int mode, int literal, int jac) {
  UserRhs u;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    int i = rhs - ODE_B200_RHS_USER_BASE;
    if (i < 0 || i >= (int)g_user.size()) {
      set_err(ODE_B200_E_INVALID, "unknown user rhs %d", rhs);
      return nullptr;
    }
    u = g_user[i];
  }
  char expr[512], expr_init[256] = "";
  const char* header = "ensemble_adapt.cuh";
  if (method_is_adaptive_rk(method)) {
    snprintf(expr, sizeof(expr), "odeb200::ens_adapt_kernel<%s, odeb200::RhsUser, %d, %s%s>", tab_name(method), mode ? 1 : 0,
             literal == 1 ? "true" : "false", literal == 2 ? ", true" : "");  // 2: uniform-parameter fast kernel
    if (!mode && !literal)  // the hinit pre-pass travels with the mode-0 fast kernel
      snprintf(expr_init, sizeof(expr_init), "odeb200::ens_adapt_init_kernel<%s, odeb200::RhsUser>", tab_name(method));
  } else if (method <= ODE_B200_M_RK4) {
    header = "ensemble_fixed.cuh";
    snprintf(expr, sizeof(expr), "odeb200::ens_fixed_kernel<%s, odeb200::RhsUser>", tab_name(method));
  } else if (method == ODE_B200_M_ODE23S) {
    header = "ensemble_ode23s.cuh";
    snprintf(expr, sizeof(expr), "odeb200::ens_ode23s_kernel<odeb200::RhsUser, %d, %d>", mode ? 1 : 0, jac ? 1 : 0);
  } else if (method == ODE_B200_M_ODE4S_S || method == ODE_B200_M_ODE4S_KR) {
    header = "ensemble_ros.cuh";
    snprintf(expr, sizeof(expr), "odeb200::ens_rosenbrock_kernel<odeb200::RhsUser, odeb200::%s, %d>",
             method == ODE_B200_M_ODE4S_KR ? "RosKapsRentrop" : "RosShampine", jac ? 1 : 0);
  } else if (ms_order(method) > 0) {
    header = "ensemble_ros.cuh";
    snprintf(expr, sizeof(expr), "odeb200::ens_ms_kernel<odeb200::RhsUser>");
  } else {
    set_err(ODE_B200_E_UNSUPPORTED, "method %d has no kernel for a run-time RHS", method);
    return nullptr;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  char keyb[640];
  snprintf(keyb, sizeof(keyb), "%d|%d|%s", dev, rhs, expr);
  const std::string key(keyb);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_kernels.find(key);
    if (it != g_kernels.end()) return it->second;
  }
  const int npa = u.np > 0 ? u.np : 1;
  std::string src = "#include \"" + std::string(header) + "\"\nnamespace odeb200 {\nstruct RhsUser {\n";
  char buf[512];
  snprintf(buf, sizeof(buf), "  static constexpr int DOF = %d, NP = %d, ID = %d;\n", u.dof, u.np, rhs);
  src += buf;
  snprintf(buf, sizeof(buf),
           "  __device__ __forceinline__ static void eval(double t, const double (&y)[%d], double (&f)[%d], const double (&p)[%d]) {\n",
           u.dof, u.dof, npa);
  src += buf;
  src += u.eval_body + "\n  }\n";
  snprintf(buf, sizeof(buf),
           "  __device__ __forceinline__ static void jac(double t, const double (&y)[%d], double (&J)[%d][%d], const double (&p)[%d]) {\n",
           u.dof, u.dof, u.dof, npa);
  src += buf;
  src += u.jac_body + "\n  }\n};\n}  // namespace odeb200\n";
  const char* exprs[2] = {expr, expr_init};
  const void* out[2] = {nullptr, nullptr};
  if (!compile_and_load(u.name, src, exprs, expr_init[0] ? 2 : 1, out)) return nullptr;
  std::lock_guard<std::mutex> lk(g_mu);
  g_kernels[key] = out[0];
  if (expr_init[0]) g_kernels[key + "|init"] = out[1];
  return out[0];
}

// hinit pre-pass kernel that was compiled together with the mode-0 fast kernel of (rhs, method); nullptr if none.
const void* user_init_kernel(int rhs, int method) {
  if (!method_is_adaptive_rk(method)) return nullptr;
  if (!user_kernel(rhs, method, 0, 0, 0)) return nullptr;
  int dev = 0;
  cudaGetDevice(&dev);
  char keyb[640];
  snprintf(keyb, sizeof(keyb), "%d|%d|odeb200::ens_adapt_kernel<%s, odeb200::RhsUser, 0, false>|init", dev, rhs, tab_name(method));
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_kernels.find(keyb);
  return it == g_kernels.end() ? nullptr : it->second;
}

bool user_stencil_kernels(int rhs, int method, const void** attempt, const void** attempt_literal, const void** hinit_f0,
                          const void** hinit_f1) {
  UserRhs u;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    int i = rhs - ODE_B200_STENCIL_USER_BASE;
    if (i < 0 || i >= (int)g_stencils.size()) {
      set_err(ODE_B200_E_INVALID, "unknown user stencil %d", rhs);
      return false;
    }
    u = g_stencils[i];
  }
  const char* tab = tab_name(method);
  if (!tab || !method_is_adaptive_rk(method)) {
    set_err(ODE_B200_E_NOT_ADAPTIVE, "Can only use this solver with an adaptive RK Butcher table");
    return false;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  char e0[256], e1[128], e2[128], e3[256], keyb[512];
  snprintf(e0, sizeof(e0), "odeb200::large_attempt_kernel<%s, odeb200::StencilUser, false>", tab);
  snprintf(e3, sizeof(e3), "odeb200::large_attempt_kernel<%s, odeb200::StencilUser, true>", tab);
  snprintf(e1, sizeof(e1), "odeb200::hinit_f0_kernel<odeb200::StencilUser>");
  snprintf(e2, sizeof(e2), "odeb200::hinit_f1_kernel<odeb200::StencilUser>");
  snprintf(keyb, sizeof(keyb), "%d|%d|%s", dev, rhs, e0);
  const std::string key(keyb);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_kernels.find(key);
    if (it != g_kernels.end()) {
      *attempt = it->second;
      *attempt_literal = g_kernels[key + "|lit"];
      *hinit_f0 = g_kernels[key + "|f0"];
      *hinit_f1 = g_kernels[key + "|f1"];
      return true;
    }
  }
  // point(w, p, t, i, n): w[k] = u at grid index i-R+k.  For a 3-point stencil the body may also use the
  // names um, u, up, and p0, p1 for the first two parameters.
  char decl[512];
  snprintf(decl, sizeof(decl),
           "  static constexpr int NP = %d, RADIUS = %d, ID = 0;\n"
           "  __device__ __forceinline__ static double point(const double (&w)[%d], const double (&p)[8], double t, "
           "long long i, long long n) {\n",
           u.np, u.dof, 2 * u.dof + 1);
  std::string src = "#include \"large.cuh\"\nnamespace odeb200 {\nstruct StencilUser {\n" + std::string(decl);
  if (u.dof == 1) src += "    const double um = w[0], u = w[1], up = w[2]; (void)um; (void)u; (void)up;\n";
  src += "    const double p0 = p[0], p1 = p[1]; (void)p0; (void)p1; (void)t; (void)i; (void)n;\n";
  src += u.eval_body + "\n  }\n};\n}  // namespace odeb200\n";
  const char* exprs[4] = {e0, e1, e2, e3};
  const void* out[4] = {nullptr, nullptr, nullptr, nullptr};
  if (!compile_and_load(u.name, src, exprs, 4, out)) return false;
  std::lock_guard<std::mutex> lk(g_mu);
  g_kernels[key] = out[0];
  g_kernels[key + "|f0"] = out[1];
  g_kernels[key + "|f1"] = out[2];
  g_kernels[key + "|lit"] = out[3];
  *attempt_literal = out[3];
  *attempt = out[0];
  *hinit_f0 = out[1];
  *hinit_f1 = out[2];
  return true;
}


bool user_field_info(int rhs, int* np) {
  std::lock_guard<std::mutex> lk(g_mu);
  int i = rhs - ODE_B200_FIELD_USER_BASE;
  if (i < 0 || i >= (int)g_fields.size()) return false;
  if (np) *np = g_fields[i].np;
  return true;
}

// Kernels of the stage-per-kernel path (field.cuh) for a registered field functor and one tableau:
// out[0] = field_first_kernel, out[1 .. S-1] = field_stage_kernel<s = 2 .. S>, then field_f1_kernel,
// field_hinit_f0_kernel, field_hinit_f1_kernel (n_stages = S).
bool user_field_kernels(int rhs, int method, int n_stages, const void** out) {
  UserRhs u;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    int i = rhs - ODE_B200_FIELD_USER_BASE;
    if (i < 0 || i >= (int)g_fields.size()) {
      set_err(ODE_B200_E_INVALID, "unknown user field %d", rhs);
      return false;
    }
    u = g_fields[i];
  }
  const char* tab = tab_name(method);
  if (!tab || !method_is_adaptive_rk(method)) {
    set_err(ODE_B200_E_NOT_ADAPTIVE, "Can only use this solver with an adaptive RK Butcher table");
    return false;
  }
  const int nk = n_stages + 3;
  int dev = 0;
  cudaGetDevice(&dev);
  char keyb[256];
  snprintf(keyb, sizeof(keyb), "%d|%d|field|%s", dev, rhs, tab);
  const std::string key(keyb);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_kernels.find(key + "|0");
    if (it != g_kernels.end()) {
      for (int k = 0; k < nk; k++) out[k] = g_kernels[key + "|" + std::to_string(k)];
      return true;
    }
  }
  std::vector<std::string> names;
  char e[256];
  snprintf(e, sizeof(e), "odeb200::field_first_kernel<%s>", tab);
  names.push_back(e);
  for (int s = 2; s <= n_stages; s++) {
    snprintf(e, sizeof(e), "odeb200::field_stage_kernel<%s, odeb200::FieldUser, %d>", tab, s);
    names.push_back(e);
  }
  names.push_back("odeb200::field_f1_kernel<odeb200::FieldUser>");
  names.push_back("odeb200::field_hinit_f0_kernel<odeb200::FieldUser>");
  names.push_back("odeb200::field_hinit_f1_kernel<odeb200::FieldUser>");
  char decl[512];
  snprintf(decl, sizeof(decl),
           "  static constexpr int NP = %d, ID = 0;\n"
           "  __device__ __forceinline__ static double rhs(long long i, long long n, const double* __restrict__ y, "
           "const double (&p)[8], double t) {\n    (void)i; (void)n; (void)y; (void)p; (void)t;\n",
           u.np);
  const std::string src = "#include \"field.cuh\"\nnamespace odeb200 {\nstruct FieldUser {\n" + std::string(decl) + u.eval_body +
                          "\n  }\n};\n}  // namespace odeb200\n";
  std::vector<const char*> exprs;
  for (const std::string& nme : names) exprs.push_back(nme.c_str());
  std::vector<const void*> got(names.size(), nullptr);
  if (!compile_and_load(u.name, src, exprs.data(), (int)exprs.size(), got.data())) return false;
  std::lock_guard<std::mutex> lk(g_mu);
  for (int k = 0; k < nk; k++) {
    g_kernels[key + "|" + std::to_string(k)] = got[k];
    out[k] = got[k];
  }
  return true;
}


// Fixed-step stage kernels (field_fixed_stage_kernel<Tab, F, s>, s = 1 .. S) for a registered field, or for a
// registered stencil through the FieldFromStencil adapter.  out[s - 1] = stage s.
bool user_fixed_kernels(int rhs, int method, int n_stages, const void** out) {
  const bool is_field = rhs >= ODE_B200_FIELD_USER_BASE;
  UserRhs u;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    std::vector<UserRhs>& reg = is_field ? g_fields : g_stencils;
    int i = rhs - (is_field ? ODE_B200_FIELD_USER_BASE : ODE_B200_STENCIL_USER_BASE);
    if (i < 0 || i >= (int)reg.size()) {
      set_err(ODE_B200_E_INVALID, "unknown large-system right-hand side %d", rhs);
      return false;
    }
    u = reg[i];
  }
  const bool ms4 = ms_order(method) > 0;  // ode_ms, any order: one kernel per step (field_ms_kernel)
  const char* tab = ms4 ? "ms" : tab_name(method);
  if (!tab || (!ms4 && method > ODE_B200_M_RK4)) {
    set_err(ODE_B200_E_UNSUPPORTED, "method %d is not a fixed-step explicit RK tableau", method);
    return false;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  char keyb[256];
  snprintf(keyb, sizeof(keyb), "%d|%d|fixed|%s", dev, rhs, tab);
  const std::string key(keyb);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_kernels.find(key + "|0");
    if (it != g_kernels.end()) {
      for (int k = 0; k < n_stages; k++) out[k] = g_kernels[key + "|" + std::to_string(k)];
      return true;
    }
  }
  std::string src = "#include \"field.cuh\"\nnamespace odeb200 {\n";
  const char* functor;
  char decl[512];
  if (is_field) {
    snprintf(decl, sizeof(decl),
             "struct FieldUser {\n  static constexpr int NP = %d, ID = 0;\n"
             "  __device__ __forceinline__ static double rhs(long long i, long long n, const double* __restrict__ y, "
             "const double (&p)[8], double t) {\n    (void)i; (void)n; (void)y; (void)p; (void)t;\n",
             u.np);
    src += std::string(decl) + u.eval_body + "\n  }\n};\n";
    functor = "odeb200::FieldUser";
  } else {
    snprintf(decl, sizeof(decl),
             "struct StencilUser {\n  static constexpr int NP = %d, RADIUS = %d, ID = 0;\n"
             "  __device__ __forceinline__ static double point(const double (&w)[%d], const double (&p)[8], double t, "
             "long long i, long long n) {\n",
             u.np, u.dof, 2 * u.dof + 1);
    src += decl;
    if (u.dof == 1) src += "    const double um = w[0], u = w[1], up = w[2]; (void)um; (void)u; (void)up;\n";
    src += "    const double p0 = p[0], p1 = p[1]; (void)p0; (void)p1; (void)t; (void)i; (void)n;\n";
    src += u.eval_body + "\n  }\n};\n";
    functor = "odeb200::FieldFromStencil<odeb200::StencilUser>";
  }
  src += "}  // namespace odeb200\n";
  std::vector<std::string> names;
  for (int st = 1; st <= n_stages; st++) {
    char e[256];
    if (ms4)
      snprintf(e, sizeof(e), "odeb200::field_ms_kernel<%s>", functor);
    else
      snprintf(e, sizeof(e), "odeb200::field_fixed_stage_kernel<%s, %s, %d>", tab, functor, st);
    names.push_back(e);
  }
  std::vector<const char*> exprs;
  for (const std::string& nme : names) exprs.push_back(nme.c_str());
  std::vector<const void*> got(names.size(), nullptr);
  if (!compile_and_load(u.name, src, exprs.data(), (int)exprs.size(), got.data())) return false;
  std::lock_guard<std::mutex> lk(g_mu);
  for (int k = 0; k < n_stages; k++) {
    g_kernels[key + "|" + std::to_string(k)] = got[k];
    out[k] = got[k];
  }
  return true;
}

}  // namespace odeb200

using namespace odeb200;

extern "C" int ode_b200_register_rhs(const char* name, int32_t dof, int32_t n_params, const char* eval_body,
                                     const char* jac_body) {
  if (!name || !eval_body || dof < 1 || dof > 16 || n_params < 0 || n_params > 64)
    return set_err(ODE_B200_E_INVALID, "register_rhs: need a name, an eval body, 1 <= dof <= 16, 0 <= n_params <= 64");
  std::lock_guard<std::mutex> lk(g_mu);
  UserRhs u;
  u.name = name;
  u.eval_body = eval_body;
  u.jac_body = jac_body ? jac_body : "";
  u.dof = dof;
  u.np = n_params;
  g_user.push_back(u);
  return ODE_B200_RHS_USER_BASE + (int)g_user.size() - 1;
}

extern "C" int ode_b200_register_stencil(const char* name, int32_t radius, int32_t n_params, const char* point_body) {
  if (!name || !point_body || radius < 1 || radius > ODE_B200_STENCIL_MAX_RADIUS || n_params < 0 ||
      n_params > ODE_B200_STENCIL_MAX_NP)
    return set_err(ODE_B200_E_INVALID, "register_stencil: need a name, the body of point(), 1 <= radius <= %d, 0 <= n_params <= %d",
                   ODE_B200_STENCIL_MAX_RADIUS, ODE_B200_STENCIL_MAX_NP);
  std::lock_guard<std::mutex> lk(g_mu);
  UserRhs u;
  u.name = name;
  u.eval_body = point_body;
  u.dof = radius;  // a stencil entry keeps its radius here
  u.np = n_params;
  g_stencils.push_back(u);
  return ODE_B200_STENCIL_USER_BASE + (int)g_stencils.size() - 1;
}

extern "C" int ode_b200_register_field(const char* name, int32_t n_params, const char* rhs_body) {
  if (!name || !rhs_body || n_params < 0 || n_params > ODE_B200_STENCIL_MAX_NP)
    return set_err(ODE_B200_E_INVALID, "register_field: need a name, the body of rhs(), 0 <= n_params <= %d",
                   ODE_B200_STENCIL_MAX_NP);
  std::lock_guard<std::mutex> lk(g_mu);
  UserRhs u;
  u.name = name;
  u.eval_body = rhs_body;
  u.dof = 0;
  u.np = n_params;
  g_fields.push_back(u);
  return ODE_B200_FIELD_USER_BASE + (int)g_fields.size() - 1;
}
