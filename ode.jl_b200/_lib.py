"""ctypes binding of libode_b200.so (include/ode_b200.h).

The shared library is the product; this module only marshals arguments.  It
fails loudly if the library is missing or cannot be loaded -- there is no CPU
or pure-Python fallback for the solvers.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ODE_B200_LIB lets a developer point at an experimental build of the same ABI
LIB_PATH = os.environ.get("ODE_B200_LIB") or os.path.join(_HERE, "libode_b200.so")

# include/ode_b200.h enums
METHODS = dict(ode1=0, ode2_midpoint=1, ode2_heun=2, ode4=3, ode21=4, ode23=5, ode45_fe=6, ode45_dp=7, ode45=7,
               ode78=8, ode23s=9, ode4s=10, ode4s_s=10, ode4s_kr=11, ode4ms=12, ode5ms=13, ode_ms1=14, ode_ms2=15,
               ode_ms3=16, ode_ms4=12, ode_ms5=13, ode45_ck=17)
RHS = dict(linear=0, const=1, timelin=2, rotation=3, lorenz=4, vdp=5, robertson=6, kepler=7, reaction_diffusion=8,
           reaction_diffusion_field=9)
POINTS = {"all": 0, "specified": 1, "final": 2}
STATUS = {0: "OK", 1: "MINSTEP", 2: "MAXSTEPS", 3: "NONFINITE", 4: "SINGULAR", 5: "OVERFLOW", 6: "EXCHANGE"}
E_ZERO_SPAN, E_INITSTEP, E_POINTS, E_NOT_ADAPTIVE, E_NO_DEVICE, E_UNSUPPORTED = -2, -3, -4, -5, -6, -8
XCHG_DOUBLES = 264
PEER_HANDLE_BYTES = 64
MAX_PEERS = 16
FP_STRICT, FP_FMA = 0, 1

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)
lp = C.POINTER(C.c_int64)


class Opts(C.Structure):
    _fields_ = [("method", C.c_int32), ("rhs", C.c_int32), ("dof", C.c_int32), ("n_params", C.c_int32),
                ("params_stride", C.c_int32), ("points", C.c_int32), ("jacobian", C.c_int32),
                ("device", C.c_int32), ("max_steps", C.c_int64),
                ("reltol", C.c_double), ("abstol", C.c_double), ("minstep", C.c_double),
                ("maxstep", C.c_double), ("initstep", C.c_double),
                ("mathmode", C.c_int32), ("host_output", C.c_int32), ("n_gpus", C.c_int32), ("fp_mode", C.c_int32),
                ("reserved", C.c_int32 * 4)]


class EnsembleIO(C.Structure):
    _fields_ = [("n_traj", C.c_int64), ("y0", C.c_void_p), ("params", C.c_void_p), ("tspan", C.c_void_p),
                ("n_tspan", C.c_int32), ("pad0", C.c_int32),
                ("t_final", C.c_void_p), ("y_final", C.c_void_p), ("n_accept", C.c_void_p),
                ("n_reject", C.c_void_p), ("n_rhs", C.c_void_p), ("status", C.c_void_p),
                ("pts_t", C.c_void_p), ("pts_y", C.c_void_p), ("pts_n", C.c_void_p), ("pts_cap", C.c_int64),
                ("trace_traj", C.c_int64), ("trace", C.c_void_p), ("trace_cap", C.c_int64),
                ("trace_n", C.c_void_p)]


class LargeStats(C.Structure):
    _fields_ = [("n_accept", C.c_int64), ("n_reject", C.c_int64), ("n_rhs", C.c_int64), ("status", C.c_int32),
                ("pad", C.c_int32), ("t_final", C.c_double), ("kernel_ms", C.c_double)]


class OdeB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libode_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Load libode_b200.so; raise if it is not there (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) and not os.environ.get("ODE_B200_LIB"):
        # compile the CUDA library in-tree if a toolchain is present (nvcc cross-compiles without a GPU)
        import shutil
        import subprocess
        if shutil.which("nvcc") and shutil.which("make"):
            subprocess.call(["make", "-C", os.path.join(_HERE, "csrc"), "-j", str(min(16, os.cpu_count() or 1)), "-s"])
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C ode.jl_b200/csrc`; there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.ode_b200_version.restype = C.c_int
    L.ode_b200_last_error.restype = C.c_char_p
    L.ode_b200_device_count.restype = C.c_int
    L.ode_b200_rhs_id.argtypes = [C.c_char_p]
    L.ode_b200_method_id.argtypes = [C.c_char_p]
    for f in ("ode_b200_rhs_dof", "ode_b200_rhs_nparams", "ode_b200_method_stages", "ode_b200_method_is_adaptive",
              "ode_b200_method_is_fsal"):
        getattr(L, f).argtypes = [C.c_int]
        getattr(L, f).restype = C.c_int
    L.ode_b200_tableau.argtypes = [C.c_int] * 4
    L.ode_b200_tableau.restype = C.c_double
    if hasattr(L, "ode_b200_ms_coefficient"):  # (absent from experimental builds of an older ABI revision)
        L.ode_b200_method_ms_order.argtypes = [C.c_int]
        L.ode_b200_ms_coefficient.argtypes = [C.c_int, C.c_int]
        L.ode_b200_ms_coefficient.restype = C.c_double
    L.ode_b200_solve_ensemble.argtypes = [C.POINTER(Opts), C.POINTER(EnsembleIO)]
    L.ode_b200_solve_ensemble_dev.argtypes = [C.POINTER(Opts), C.POINTER(EnsembleIO), C.c_void_p]
    L.ode_b200_last_kernel_ms.restype = C.c_double
    L.ode_b200_launch_count.argtypes = [C.c_int]
    L.ode_b200_launch_count.restype = C.c_int64
    L.ode_b200_solve_single.argtypes = [C.POINTER(Opts), dp, dp, dp, C.c_int32, C.POINTER(C.c_void_p)]
    L.ode_b200_result_len.argtypes = [C.c_void_p]
    L.ode_b200_result_len.restype = C.c_int64
    L.ode_b200_result_dof.argtypes = [C.c_void_p]
    L.ode_b200_result_status.argtypes = [C.c_void_p]
    L.ode_b200_result_counts.argtypes = [C.c_void_p, C.c_int]
    L.ode_b200_result_counts.restype = C.c_int64
    L.ode_b200_result_copy.argtypes = [C.c_void_p, dp, dp]
    L.ode_b200_result_free.argtypes = [C.c_void_p]
    L.ode_b200_result_free.restype = None
    L.ode_b200_solve_large.argtypes = [C.POINTER(Opts), C.c_int64, dp, dp, C.c_double, C.c_double, dp,
                                       C.POINTER(LargeStats), dp, C.c_int64, lp]
    L.ode_b200_solve_large_points.argtypes = [C.POINTER(Opts), C.c_int64, dp, dp, dp, C.c_int32, dp,
                                              C.POINTER(LargeStats), dp, C.c_int64, lp]
    L.ode_b200_slab_set_output.argtypes = [C.c_void_p, dp, C.c_int32, C.c_void_p]
    L.ode_b200_slab_set_output_all.argtypes = [C.c_void_p, dp, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]
    L.ode_b200_solve_large_all.argtypes = [C.POINTER(Opts), C.c_int64, dp, dp, dp, C.c_int32, C.c_int64, dp, dp, lp,
                                           C.POINTER(LargeStats), dp, C.c_int64, lp]
    L.ode_b200_solve_large_fixed.argtypes = [C.POINTER(Opts), C.c_int64, dp, dp, dp, C.c_int32, dp, dp, C.POINTER(LargeStats)]
    L.ode_b200_trim.restype = None
    L.ode_b200_register_stencil.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_char_p]
    L.ode_b200_register_field.argtypes = [C.c_char_p, C.c_int32, C.c_char_p]
    L.ode_b200_register_rhs.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_char_p, C.c_char_p]
    L.ode_b200_slab_create.argtypes = [C.POINTER(Opts), C.c_int64, C.c_int64, C.c_int32, C.c_int32, dp, C.c_double,
                                       C.c_double, C.c_void_p, C.POINTER(C.c_void_p)]
    L.ode_b200_slab_set_state.argtypes = [C.c_void_p, C.c_void_p]
    L.ode_b200_slab_get_state.argtypes = [C.c_void_p, C.c_void_p]
    L.ode_b200_slab_reset.argtypes = [C.c_void_p, C.c_double, C.c_double]
    L.ode_b200_slab_set_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.ode_b200_slab_set_offset.argtypes = [C.c_void_p, C.c_int64]
    L.ode_b200_slab_set_fused.argtypes = [C.c_void_p, C.c_int32]
    L.ode_b200_slab_init_phase.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.ode_b200_slab_attempt.argtypes = [C.c_void_p, C.c_void_p]
    L.ode_b200_slab_control.argtypes = [C.c_void_p, C.c_void_p]
    L.ode_b200_slab_poll.argtypes = [C.c_void_p, dp]
    L.ode_b200_solve_large_multi.argtypes = [C.POINTER(Opts), C.c_int32, ip, C.c_int64, dp, dp, C.c_double, C.c_double, dp,
                                             C.POINTER(LargeStats), dp, C.c_int64, lp]
    L.ode_b200_slab_peer_alloc.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    L.ode_b200_slab_peer_connect.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    L.ode_b200_slab_run.argtypes = [C.c_void_p, dp, dp]
    L.ode_b200_slab_free.argtypes = [C.c_void_p]
    L.ode_b200_slab_free.restype = None
    L.ode_b200_host_alloc.argtypes = [C.c_size_t]
    L.ode_b200_host_alloc.restype = C.c_void_p
    L.ode_b200_host_free.argtypes = [C.c_void_p]
    L.ode_b200_host_free.restype = None
    L.ode_b200_fp64_peak.argtypes = [C.c_int, C.c_int]
    L.ode_b200_fp64_peak.restype = C.c_double
    L.ode_b200_selftest_detmath.argtypes = [C.c_int, dp, dp, dp, dp, C.c_int64]
    L.ode_b200_selftest_div.argtypes = [C.c_int, C.c_int64, C.c_uint64]
    L.ode_b200_selftest_div.restype = C.c_int64
    if L.ode_b200_version() != 1:
        raise ImportError("libode_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise OdeB200Error(rc, lib().ode_b200_last_error().decode())
