"""ctypes binding of libcafe_gpu.so (the C ABI in include/cafe_gpu.h).

The library must be built (cafe_mpc_b200/build.py or __graft_entry__.build()); there is no
Python / CPU fallback: if it is missing, importing this module raises.
CAFE_HOST_ONLY=1 (set by bench.py's CPU baseline arm before the import) loads libcafe_host.so instead: the problem builders and settings
readers alone, built by g++ - no solver entry point exists in that process and no CUDA code is mapped."""
import ctypes as C
import os

from ._ctypes_defs import Deck, Info, Options, CAFE_NKERNELS

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_ONLY = os.environ.get("CAFE_HOST_ONLY", "0") == "1"
LIB_PATH = os.path.join(HERE, "libcafe_host.so" if HOST_ONLY else "libcafe_gpu.so")


class CafeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cafe error %d: %s" % (code, msg))
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "the library is not built (%s). Run `python cafe_mpc_b200/build.py`; "
            "the product path has no fallback implementation." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.cafe_last_error.restype = C.c_char_p
    lib.cafe_options_load.argtypes = [C.c_char_p, C.POINTER(Options)]
    lib.cafe_deck_build_hkd.argtypes = [C.c_char_p, C.c_char_p, C.c_float, C.c_float, C.c_int, C.c_int, C.POINTER(vp)]
    lib.cafe_deck_build_mhpc.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(vp)]
    lib.cafe_deck_build_loco.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(vp)]
    lib.cafe_deck_build_barrel_to.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp)]
    lib.cafe_barrel_to_initial_guess.argtypes = [C.POINTER(Deck), vp, C.c_int, vp]
    lib.cafe_deck_mark_mpc_update.argtypes = [vp, C.c_int, ip]
    lib.cafe_deck_get.restype = C.POINTER(Deck)
    lib.cafe_deck_get.argtypes = [vp]
    lib.cafe_deck_free.argtypes = [vp]
    lib.cafe_deck_free.restype = None
    lib.cafe_hkd_state.argtypes = [dp, dp, ip, dp]
    lib.cafe_deck_lq_pattern.argtypes = [C.POINTER(Deck), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_ulonglong)]
    lib.cafe_solution_size.restype = C.c_long
    lib.cafe_solution_size.argtypes = [C.POINTER(Deck)]
    lib.cafe_command_size.restype = C.c_long
    lib.cafe_command_size.argtypes = [C.POINTER(Deck), C.c_int]
    if HOST_ONLY:
        return lib
    lib.cafe_lcm_command_size.restype = C.c_long
    lib.cafe_lcm_command_size.argtypes = [C.c_int]
    lib.cafe_hkd_lcm_command_size.restype = C.c_long
    lib.cafe_hkd_lcm_command_size.argtypes = [C.c_int]
    lib.cafe_gpu_get_hkd_lcm_commands.argtypes = [vp, C.c_int, vp]
    lib.cafe_gpu_get_hkd_lcm_commands_device.argtypes = [vp, C.c_int, vp]
    lib.cafe_gpu_create.argtypes = [C.POINTER(Deck), C.c_int, C.c_int, C.POINTER(vp)]
    lib.cafe_gpu_destroy.argtypes = [vp]
    lib.cafe_gpu_solve_batch.argtypes = [vp, vp, C.c_int, C.POINTER(Options)]
    lib.cafe_gpu_solve_batch_device.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(Options)]
    lib.cafe_gpu_get_info.argtypes = [vp, C.POINTER(Info)]
    lib.cafe_gpu_get_history.argtypes = [vp, vp, C.c_int]
    lib.cafe_gpu_get_trace.argtypes = [vp, vp, C.c_int]
    lib.cafe_gpu_get_solution.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.cafe_gpu_get_commands.argtypes = [vp, C.c_int, vp]
    lib.cafe_gpu_get_commands_device.argtypes = [vp, C.c_int, vp]
    lib.cafe_gpu_shift_guess.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
    lib.cafe_gpu_get_planned_state.argtypes = [vp, C.c_int, vp]
    lib.cafe_gpu_update_deck.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.cafe_gpu_get_solve_ms.argtypes = [vp, dp]
    lib.cafe_gpu_get_timing.argtypes = [vp, C.POINTER(C.c_double * CAFE_NKERNELS), C.POINTER(C.c_long * CAFE_NKERNELS), ip]
    lib.cafe_gpu_set_profiling.argtypes = [vp, C.c_int]
    lib.cafe_gpu_get_units.argtypes = [vp, C.POINTER(C.c_double * CAFE_NKERNELS)]
    lib.cafe_gpu_debug_get.restype = C.c_long
    lib.cafe_gpu_debug_get.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, vp]
    lib.cafe_gpu_measure_fp64_peak.argtypes = [C.c_int, dp]
    # multi-GPU
    lib.cafe_gpu_shard_range.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip]
    lib.cafe_gpu_create_multi.argtypes = [C.POINTER(Deck), C.c_int, ip, C.c_int, C.POINTER(vp)]
    lib.cafe_gpu_multi_destroy.argtypes = [vp]
    lib.cafe_gpu_multi_ndev.argtypes = [vp]
    lib.cafe_gpu_multi_handle.restype = vp
    lib.cafe_gpu_multi_handle.argtypes = [vp, C.c_int]
    lib.cafe_gpu_multi_solve_batch.argtypes = [vp, vp, C.c_int, C.POINTER(Options)]
    lib.cafe_gpu_multi_update_deck.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.cafe_gpu_multi_get_info.argtypes = [vp, C.POINTER(Info)]
    lib.cafe_gpu_multi_get_commands.argtypes = [vp, C.c_int, vp]
    lib.cafe_gpu_nccl_unique_id.argtypes = [C.c_char_p]
    lib.cafe_gpu_comm_init_rank.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    lib.cafe_gpu_comm_destroy.argtypes = [vp]
    lib.cafe_gpu_gather_commands.argtypes = [vp, C.c_int, C.c_int, vp]
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        raise CafeError(rc, lib.cafe_last_error().decode())


EXPORTED = [
    "cafe_last_error", "cafe_options_load", "cafe_info_get_number", "cafe_info_get_string", "cafe_deck_build_mhpc_config", "cafe_deck_phase_times", "cafe_deck_build_hkd", "cafe_deck_build_mhpc", "cafe_deck_build_loco", "cafe_deck_build_barrel_to", "cafe_barrel_to_initial_guess", "cafe_deck_mark_mpc_update", "cafe_deck_get",
    "cafe_deck_free", "cafe_hkd_state", "cafe_deck_lq_pattern", "cafe_solution_size", "cafe_command_size", "cafe_gpu_create",
    "cafe_gpu_destroy", "cafe_gpu_solve_batch", "cafe_gpu_solve_batch_device", "cafe_gpu_get_info",
    "cafe_gpu_get_history", "cafe_gpu_get_trace", "cafe_gpu_get_solution", "cafe_gpu_get_commands", "cafe_gpu_get_commands_device", "cafe_gpu_get_commands_async", "cafe_gpu_commands_wait", "cafe_gpu_gather_commands_async", "cafe_gpu_get_solve_ms",
    "cafe_gpu_set_references", "cafe_gpu_set_initial_guess", "cafe_gpu_set_al_params", "cafe_gpu_get_al_params", "cafe_gpu_shift_guess", "cafe_gpu_get_planned_state", "cafe_gpu_update_deck", "cafe_lcm_command_size", "cafe_gpu_get_lcm_commands", "cafe_gpu_get_lcm_commands_device", "cafe_hkd_lcm_command_size", "cafe_gpu_get_hkd_lcm_commands", "cafe_gpu_get_hkd_lcm_commands_device",
    "cafe_gpu_get_timing", "cafe_gpu_get_units", "cafe_gpu_set_profiling", "cafe_gpu_debug_get", "cafe_gpu_measure_fp64_peak",
    "cafe_gpu_shard_range", "cafe_gpu_create_multi", "cafe_gpu_multi_destroy", "cafe_gpu_multi_ndev", "cafe_gpu_multi_handle", "cafe_gpu_multi_solve_batch", "cafe_gpu_multi_update_deck",
    "cafe_gpu_multi_get_info", "cafe_gpu_multi_get_commands", "cafe_gpu_nccl_unique_id", "cafe_gpu_comm_init_rank", "cafe_gpu_comm_destroy", "cafe_gpu_gather_commands",
]
