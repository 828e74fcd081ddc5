"""ctypes binding of libgymcook.so (include/gymcook.h).  Thin marshalling only: every array is a
torch CUDA tensor owned by the caller, passed by `data_ptr()`; kernels run on torch's current
stream.  There is no CPU fallback - a missing library or a missing GPU raises."""
import ctypes as C
import os

import torch

from . import build as _build

MAX_AGENTS, MAX_OBJECTS, MAX_GOALS, MAX_CELLS, MAX_SUBTASKS, MAX_LEVELS = 4, 6, 4, 64, 32, 16
STATS_LEN = 134
SLOT_DEAD = 0xE000
PLACE_HELD, PLACE_DEAD = 0x40, 0x47
PLAN_JOINT_ACTIONS = 1


class GcError(RuntimeError):
    pass


class Subtask(C.Structure):
    _fields_ = [("kind", C.c_uint8), ("a", C.c_uint8), ("b", C.c_uint8), ("goal", C.c_uint8)]


class Level(C.Structure):
    """struct gc_level (256 bytes)."""
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("n_agent_starts", C.c_int32),
        ("n_objects", C.c_int32), ("n_goals", C.c_int32), ("delivery_cell", C.c_int32),
        ("max_timesteps", C.c_int32), ("n_subtasks", C.c_int32),
        ("cell_type", C.c_uint8 * MAX_CELLS), ("agent_cell", C.c_uint8 * MAX_AGENTS),
        ("object_init", C.c_uint16 * MAX_OBJECTS), ("goal_mask", C.c_uint8 * MAX_GOALS),
        ("subtask", Subtask * MAX_SUBTASKS), ("recipe_code", C.c_uint8 * MAX_GOALS),
        ("reserved", C.c_uint8 * 8),
    ]


assert C.sizeof(Level) == 256

_lib = None

_VOIDP = C.c_void_p
_SIGNATURES = {
    "gc_version": (C.c_int, []),
    "gc_last_error": (C.c_char_p, []),
    "gc_device_count": (C.c_int, []),
    "gc_level_parse": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(Level)]),
    "gc_level_set_subtasks": (C.c_int, [C.POINTER(Level), C.POINTER(Subtask), C.c_int]),
    "gc_env_reset": (C.c_int, [C.POINTER(Level), C.c_int, _VOIDP, _VOIDP, C.c_int64, C.c_int, _VOIDP]),
    "gc_env_prepare": (C.c_int, [C.POINTER(Level), C.c_int, C.c_int]),
    "gc_env_step": (C.c_int, [C.POINTER(Level), C.c_int, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP,
                              _VOIDP, C.c_int64, C.c_int, _VOIDP]),
    "gc_env_step_host": (C.c_int, [C.POINTER(Level), C.c_int, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP,
                                   _VOIDP, _VOIDP, C.c_int64, C.c_int, _VOIDP]),
    "gc_step_plan_create": (C.c_int, [C.POINTER(Level), _VOIDP, _VOIDP, C.c_int64, C.c_int, C.c_int, C.POINTER(_VOIDP)]),
    "gc_step_plan_run": (C.c_int, [_VOIDP, _VOIDP, _VOIDP]),
    "gc_step_plan_run_host": (C.c_int, [_VOIDP, _VOIDP, _VOIDP, _VOIDP]),
    "gc_step_plan_enqueue_host": (C.c_int, [_VOIDP, _VOIDP, _VOIDP, _VOIDP]),
    "gc_step_plan_destroy": (None, [_VOIDP]),
    "gc_env_rollout": (C.c_int, [C.POINTER(Level), C.c_int, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP,
                                 C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_uint64, _VOIDP]),
    "gc_fill_random_actions": (C.c_int, [_VOIDP, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64,
                                         C.c_uint64, _VOIDP]),
    "gc_state_hash": (C.c_int, [_VOIDP, _VOIDP, C.c_int64, C.c_int, _VOIDP]),
    "gc_stats_reduce": (C.c_int, [_VOIDP, _VOIDP, C.POINTER(Level), C.c_int, _VOIDP, _VOIDP, C.c_int64, _VOIDP]),
    "gc_bd_posterior_f32": (C.c_int, [_VOIDP] * 7 + [C.c_float, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _VOIDP]),
    "gc_bd_posterior_f64": (C.c_int, [_VOIDP] * 7 + [C.c_double, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _VOIDP]),
    "gc_bd_likelihood_rows_f32": (C.c_int, [_VOIDP, _VOIDP, C.c_int, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, C.c_int,
                                            C.c_float, C.c_float, _VOIDP, _VOIDP, _VOIDP, C.c_int64, C.c_int, C.c_int,
                                            C.c_int, _VOIDP]),
    "gc_bd_likelihood_rows_f64": (C.c_int, [_VOIDP, _VOIDP, C.c_int, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, C.c_int,
                                            C.c_double, C.c_double, _VOIDP, _VOIDP, _VOIDP, C.c_int64, C.c_int, C.c_int,
                                            C.c_int, _VOIDP]),
    "gc_bd_update_lists_f32": (C.c_int, [_VOIDP, _VOIDP, _VOIDP, C.c_int, _VOIDP, C.c_int, C.c_int, _VOIDP, _VOIDP, _VOIDP,
                                         C.c_int, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, C.c_int, C.c_float,
                                         C.c_float, C.c_float, C.c_int64, C.c_int, C.c_int, _VOIDP]),
    "gc_bd_update_lists_f64": (C.c_int, [_VOIDP, _VOIDP, _VOIDP, C.c_int, _VOIDP, C.c_int, C.c_int, _VOIDP, _VOIDP, _VOIDP,
                                         C.c_int, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, C.c_int, C.c_double,
                                         C.c_double, C.c_double, C.c_int64, C.c_int, C.c_int, _VOIDP]),
    "gc_offered_actions": (C.c_int, [C.POINTER(Level), C.c_int, _VOIDP, _VOIDP, _VOIDP, C.c_int64, C.c_int, _VOIDP]),
    "gc_subtasks_completed": (C.c_int, [C.POINTER(Level), C.c_int, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, C.c_int64, C.c_int,
                                        _VOIDP]),
    "gc_lower_bound": (C.c_int, [C.POINTER(Level), C.c_int, _VOIDP, _VOIDP, _VOIDP, C.c_int, _VOIDP,
                                 C.c_int64, C.c_int, _VOIDP]),
    "gc_subtask_q": (C.c_int, [C.POINTER(Level), C.c_int, _VOIDP, _VOIDP, _VOIDP, C.c_int, _VOIDP, _VOIDP,
                               _VOIDP, C.c_int64, C.c_int, _VOIDP]),
    "gc_joint_q_scratch_bytes": (C.c_int64, [C.c_int64, C.c_int, C.POINTER(C.c_int)]),
    "gc_joint_q": (C.c_int, [C.POINTER(Level), C.c_int, _VOIDP, _VOIDP, _VOIDP, C.c_int, _VOIDP, _VOIDP, _VOIDP,
                             _VOIDP, C.c_int64, C.c_int64, C.c_int, _VOIDP]),
    "gc_render": (C.c_int, [C.POINTER(Level), C.c_int, _VOIDP, _VOIDP, _VOIDP, _VOIDP, C.c_int64, C.c_int, _VOIDP]),
}
# entry points declared in include/gymcook.h; tests check that each one is exported
ABI_SYMBOLS = tuple(_SIGNATURES)


def lib_path():
    return _build.LIB_PATH


def load():
    """dlopen libgymcook.so (building it first if nvcc is present and the sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if _build.is_stale():
        try:
            _build.build()
        except Exception as exc:  # no nvcc on this box: a prebuilt .so must exist
            if not os.path.exists(path):
                raise GcError("libgymcook.so is missing and could not be built: %s" % exc)
    if not os.path.exists(path):
        raise GcError("libgymcook.so not found at %s - run `python __graft_entry__.py build`" % path)
    L = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError here = header and library disagree
        fn.restype, fn.argtypes = res, args
    if L.gc_version() != 2:
        raise GcError("libgymcook.so ABI version %d, expected 2" % L.gc_version())
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise GcError("libgymcook error %d: %s" % (rc, load().gc_last_error().decode()))


def ptr(t, dtype=None):
    """device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise GcError("expected a CUDA tensor: libgymcook has no CPU path")
    if not t.is_contiguous():
        raise GcError("expected a contiguous tensor")
    if dtype is not None and t.dtype != dtype:
        raise GcError("expected dtype %s, got %s" % (dtype, t.dtype))
    return t.data_ptr()


try:  # ~0.2 us instead of ~1 us for torch.cuda.current_stream(device).cuda_stream: it is on the step path
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:  # pragma: no cover
    _raw_stream = None


def stream_ptr(device=None):
    if _raw_stream is not None and device is not None and device.index is not None:
        return _raw_stream(device.index)
    return torch.cuda.current_stream(device).cuda_stream


def parse_level(text, max_timesteps=100):
    lv = Level()
    data = text.encode()
    check(load().gc_level_parse(data, len(data), max_timesteps, C.byref(lv)))
    return lv


def level_array(levels):
    arr = (Level * len(levels))()
    for i, lv in enumerate(levels):
        C.memmove(C.byref(arr, i * C.sizeof(Level)), C.byref(lv), C.sizeof(Level))
    return arr
